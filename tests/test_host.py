"""CPU-side coverage: C ABI surface, bit-exact tuple indices, graph batching, pack tables, module
state_dict layout, flat parameter buffers.  No kernel is launched here."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from util import LEVELS, load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    from grappa_b200 import _lib
    lib = _lib.lib()
    header = open(os.path.join(ROOT, "include", "grappa_b200.h")).read()
    names = sorted(set(re.findall(r"\b(grappa_b200_[a-z0-9_]+)\s*\(", header)))
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"libgrappa_b200.so does not export {n}"
    assert lib.grappa_b200_abi_version() == 3
    assert isinstance(_lib.launch_count(), int)


def test_abi_reports_errors_instead_of_crashing():
    from grappa_b200 import _lib
    from grappa_b200._lib_ops import GemmArgs
    lib = _lib.lib()
    g = GemmArgs(M=4, N=4, K=4, lda=4, ldb=4, ldc=4)          # NULL operands
    rc = lib.grappa_b200_gemm(ctypes.byref(g), None)
    assert rc == -1 and b"NULL" in lib.grappa_b200_last_error()
    with pytest.raises(_lib.GrappaB200Error):
        _lib.check(rc, "gemm")
    bad = np.array([[0, 0]], dtype=np.int64)                  # self-bond
    na, npr = ctypes.c_int64(), ctypes.c_int64()
    assert lib.grappa_b200_tuples_count(bad.ctypes.data, 1, ctypes.byref(na), ctypes.byref(npr)) == -1
    assert b"self-bond" in lib.grappa_b200_last_error()


def test_struct_layouts_match_the_header():
    """sizeof of the ctypes mirrors == what the C compiler lays out (computed from the field lists: all
    fields are naturally aligned 4/8-byte scalars, so ctypes and C agree; this guards field-count drift)."""
    from grappa_b200._lib import EnergyArgs, EnergyBwdArgs
    from grappa_b200._lib_ops import GemmArgs, HeadOutArgs, LossArgs, Perms
    assert ctypes.sizeof(Perms) == 4 + 6 * 4 * 4
    assert ctypes.sizeof(GemmArgs) % 8 == 0 and ctypes.sizeof(EnergyArgs) % 8 == 0
    assert GemmArgs._fields_[-1][0] == "max_sms" and GemmArgs._fields_[-3][0] == "colsum"
    assert ctypes.sizeof(EnergyBwdArgs) > ctypes.sizeof(EnergyArgs)
    # 27 4-byte scalars (108 bytes), 4 bytes of padding, then the four statistics pointers
    assert HeadOutArgs.stat.offset == 112 and ctypes.sizeof(HeadOutArgs) == 112 + 4 * 8
    from grappa_b200._lib_ops import HeadStatGrads
    assert ctypes.sizeof(HeadStatGrads) == 4 * 8 + 8
    assert ctypes.sizeof(LossArgs) == 9 * 8 + 4 * 4 + 4 * 4 + 9 * 8
    from grappa_b200._lib_ops import ParamLossArgs
    assert ctypes.sizeof(ParamLossArgs) == 2 * 4 + 3 * 5 * 8 + 3 * 5 * 4 + 4 + 2 * 8 + 5 * 8 + 8   # 4 bytes of padding after fac[5]
    from grappa_b200._lib_ops import IpcHandle, PeerAllreduceArgs
    assert ctypes.sizeof(IpcHandle) == 64                                   # cudaIpcMemHandle_t
    # data[8], flags[8], rank, world, start, count, epoch, ctas, split
    assert ctypes.sizeof(PeerAllreduceArgs) == 16 * 8 + 2 * 4 + 2 * 8 + 8 + 4 + 4 and PeerAllreduceArgs.start.offset == 136


@pytest.mark.parametrize("name", ["dipeptide", "peptide4", "tree", "rna", "protein30"])
def test_tuple_indices_bit_exact_vs_reference_fixture(name):
    from grappa_b200 import tuples
    z = load_golden("tuple_indices.npz")
    out = tuples.build_tuples(0, z[f"{name}.in_bonds"], z[f"{name}.in_improper_candidates"])
    for k in ("bonds", "angles", "propers", "impropers"):
        assert out[k].dtype == np.int64 and np.array_equal(out[k], z[f"{name}.{k}"]), k


def test_tuple_indices_vs_bruteforce_path_enumeration():
    import networkx as nx
    from grappa_b200 import synthetic, tuples
    rng = np.random.default_rng(0)
    el, bonds, _ = synthetic.random_tree_topology(rng, 30, 2)
    out = tuples.get_idx_tuples(bonds)
    G = nx.Graph([tuple(b) for b in bonds.tolist()])
    angles = {(a, b, c) if a < c else (c, b, a) for b in G for a in G[b] for c in G[b] if a != c}
    propers = set()
    for b, c in G.edges:
        for a in G[b]:
            for d in G[c]:
                if a != c and d != b and a != d:
                    propers.add((a, b, c, d) if a < d else (d, c, b, a))
    assert {tuple(r) for r in out["angles"].tolist()} == angles
    assert {tuple(r) for r in out["propers"].tolist()} == propers
    assert len(out["angles"]) == len(angles) and len(out["propers"]) == len(propers)


def test_synthetic_shapes_match_the_survey_table():
    from grappa_b200 import synthetic
    for n_res, want in [(1, (22, 21, 36, 41, 12)), (4, (52, 51, 90, 116, 30))]:
        g = synthetic.make_molecule(np.random.default_rng(0), "peptide", n_confs=2, n_res=n_res)
        got = (g.num_nodes("n1"),) + tuple(g.num_nodes(l) for l in LEVELS)
        assert got == want
    g = synthetic.protein(seed=0)
    assert (g.num_nodes("n1"), g.num_nodes("n2"), g.num_nodes("n3"), g.num_nodes("n4"), g.num_nodes("n4_improper")) == \
        (1502, 1501, 2700, 3741, 900)
    assert g.nodes["n1"].data["atomic_number"].shape == (1502, 53)
    gb = synthetic.peptide_batch(seed=0, batch_size=32)
    assert gb.nodes["n1"].data["xyz"].shape == (32 * 52, 50, 3)


def test_batch_unbatch_identity_and_offsets():
    """reference tests/dgl_utils.py:34-53: batch -> unbatch is the identity on every feature; idxs are shifted."""
    from grappa_b200 import graph as gbg, synthetic
    rng = np.random.default_rng(1)
    mols = [synthetic.make_molecule(rng, k, n_confs=3, **kw) for k, kw in
            [("peptide", dict(n_res=1)), ("small", dict(n_atoms=12)), ("rna", dict(n_atoms=92))]]
    b = gbg.batch(mols)
    assert b.batch_size == 3 and b.num_nodes("g") == 3
    assert b.nodes["n2"].data["idxs"].max() < b.num_nodes("n1")
    off = mols[0].num_nodes("n1")
    assert torch.equal(b.nodes["n3"].data["idxs"][mols[0].num_nodes("n3"):mols[0].num_nodes("n3") + mols[1].num_nodes("n3")],
                       mols[1].nodes["n3"].data["idxs"] + off)
    back = gbg.unbatch(b)
    for a, c in zip(mols, back):
        for nt in a.ntypes:
            assert a.num_nodes(nt) == c.num_nodes(nt)
            for k, v in a.nodes[nt].data.items():
                assert torch.equal(v, c.nodes[nt].data[k]), (nt, k)
        assert sorted(zip(*[t.tolist() for t in a.edges()])) == sorted(zip(*[t.tolist() for t in c.edges()]))
    with pytest.raises(ValueError):
        gbg.batch([mols[0], synthetic.make_molecule(rng, "small", n_confs=4, n_atoms=5)])


def test_pack_tables():
    from grappa_b200 import synthetic
    from grappa_b200.pack import PackedBatch
    g = synthetic.espaloma_mix_batch(seed=2, batch_size=5, n_confs=2)
    p = PackedBatch(g, device="cpu")
    n = g.num_nodes("n1")
    src, dst = [t.numpy() for t in g.edges()]
    indptr, esrc, erev = p.host["indptr"], p.host["esrc"], p.host["erev"]
    edst = np.repeat(np.arange(n), np.diff(indptr))
    assert sorted(zip(esrc.tolist(), edst.tolist())) == sorted(zip(src.tolist(), dst.tolist()))
    assert np.array_equal(esrc[erev], edst) and np.array_equal(edst[erev], esrc)      # reverse edges
    for l, lvl in enumerate(LEVELS):
        idx = g.nodes[lvl].data["idxs"].numpy()
        L = idx.shape[1]
        assert np.array_equal(p.host[f"idx{l}"], idx.astype(np.int32))
        ptr, ent = p.host[f"inv_ptr{l}"], p.host[f"inv_ent{l}"]
        assert ptr[-1] == idx.size
        for atom in (0, n // 2, n - 1):
            assert all(idx.reshape(-1)[e] == atom for e in ent[ptr[atom]:ptr[atom + 1]])
        assert np.array_equal(np.diff(p.host[f"tup_off{l}"]), g.batch_num_nodes(lvl).numpy())
    g.nodes["n2"].data["idxs"][0, 0] = n + 3
    with pytest.raises(AssertionError):
        PackedBatch(g, device="cpu")
    g.nodes["n2"].data["idxs"] = g.nodes["n2"].data["idxs"].float()
    with pytest.raises(IndexError):
        PackedBatch(g, device="cpu")


def test_conflict_free_rounds_schedule():
    """The energy kernel's force scatter is atomics-free because no two tuples of one round share an atom: check that,
    that every tuple is scheduled exactly once inside its molecule's rounds, and that the C++ scheduler is the first-fit
    greedy it documents (same schedule as a plain Python restatement)."""
    from grappa_b200 import synthetic
    from grappa_b200.pack import PackedBatch, conflict_free_rounds
    g = synthetic.espaloma_mix_batch(seed=3, batch_size=6, n_confs=1)
    p = PackedBatch(g, device="cpu")
    G = p.sched_groups
    for l, lvl in enumerate(LEVELS):
        idx, tup_off = p.host[f"idx{l}"], p.host[f"tup_off{l}"]
        ro, sc = p.host[f"round_off{l}"], p.host[f"sched{l}"]
        assert sc.shape == (ro[-1], G) and ro[0] == 0
        used = sc[sc >= 0]
        assert sorted(used.tolist()) == list(range(idx.shape[0]))
        want = []
        for b in range(p.n_mols):
            rounds = sc[ro[b]:ro[b + 1]]
            mine = rounds[rounds >= 0]
            assert ((mine >= tup_off[b]) & (mine < tup_off[b + 1])).all()
            for r in rounds:
                atoms = idx[r[r >= 0]].reshape(-1)
                assert len(set(atoms.tolist())) == atoms.size
                assert (r[:int((r >= 0).sum())] >= 0).all()                    # filled slots come first
            # first-fit: the lowest round with a free slot whose tuples share no atom with this one
            mol_rounds = []
            for t in range(tup_off[b], tup_off[b + 1]):
                atoms = set(idx[t].tolist())
                for rr in mol_rounds:
                    if len(rr) < G and not any(atoms & set(idx[u].tolist()) for u in rr):
                        rr.append(t)
                        break
                else:
                    mol_rounds.append([t])
            want += [rr + [-1] * (G - len(rr)) for rr in mol_rounds]
        assert np.array_equal(sc, np.array(want, dtype=np.int32).reshape(-1, G))
    # molecules without tuples, and an empty level
    ro, sc = conflict_free_rounds(np.array([[0, 1], [1, 2], [5, 6]], np.int32), np.array([0, 2, 2, 3], np.int32), 3, 2, 4)
    assert ro.tolist() == [0, 2, 2, 3] and sc.tolist() == [[0, -1, -1, -1], [1, -1, -1, -1], [2, -1, -1, -1]]
    ro, sc = conflict_free_rounds(np.zeros((0, 4), np.int32), np.zeros(3, np.int32), 2, 4, 8)
    assert ro.tolist() == [0, 0, 0] and sc.shape == (0, 8)


def test_ring_encoding_known_answers_and_python_restatement():
    """Ring-membership features without rdkit (reference utils/rdkit_utils.py:7-24: IsInRing, IsInRingSize(3..8)): the
    host C++ routine against known answers and against the search written out in Python (synthetic.ring_encoding_py)."""
    from grappa_b200 import GrappaB200Error, synthetic, tuples

    def ring(n, o=0):
        return [(o + i, o + (i + 1) % n) for i in range(n)]
    for n in range(3, 13):
        e = tuples.ring_encoding(n, ring(n))
        want = [1.0] + [float(n == s) for s in range(3, 9)]          # a 9+-ring is a ring of no listed size
        assert (e == np.array(want, dtype=np.float32)).all(), n
    # fused 6-6 system (atoms 0-9), a substituent (10) that carries a cyclopropane (11-13): bridges are not ring bonds
    bonds = ring(6) + [(0, 6), (6, 7), (7, 8), (8, 9), (9, 5)] + [(3, 10), (10, 11)] + ring(3, 11)
    e = tuples.ring_encoding(14, bonds)
    assert e[:10].tolist() == [[1, 0, 0, 0, 1, 0, 0]] * 10 and e[10].tolist() == [0] * 7
    assert e[11:].tolist() == [[1, 1, 0, 0, 0, 0, 0]] * 3
    # norbornane: two five-rings (the six-membered envelope is not a smallest ring of any bond)
    nb = [(0, 1), (1, 2), (2, 3), (3, 4), (4, 5), (5, 0), (0, 6), (6, 3)]
    assert tuples.ring_encoding(7, nb).tolist() == [[1, 0, 0, 1, 0, 0, 0]] * 7
    rng = np.random.default_rng(0)
    for kind, kw in (("peptide", dict(n_res=2)), ("small", dict(n_atoms=30)), ("rna", dict(n_atoms=93))):
        for _ in range(8):
            m = synthetic.make_molecule(rng, kind, n_confs=1, **kw)
            src, dst = [t.numpy() for t in m.edges()]
            b = np.stack([src, dst], 1)[src < dst]
            n = m.num_nodes("n1")
            assert np.array_equal(tuples.ring_encoding(n, b), synthetic.ring_encoding_py(n, b))
    assert tuples.ring_encoding(4, np.zeros((0, 2), np.int64)).tolist() == [[0] * 7] * 4
    with pytest.raises(GrappaB200Error):
        tuples.ring_encoding(3, [(0, 5)])


def test_module_tree_state_dict_and_error_behaviour():
    from grappa_b200 import GrappaB200Error, models, synthetic
    from grappa_b200.energy import Energy
    z = load_golden("dipeptide_grappa12.npz")
    m = models.model_from_config(models.grappa_1_2_model_config())
    sd = m.state_dict()
    assert sorted(sd.keys()) == list(z["meta.state_dict_keys"])
    assert [",".join(map(str, sd[k].shape)) for k in sorted(sd)] == list(z["meta.state_dict_shapes"])
    assert sd["gnn.blocks.3.head_reducer.weight"].data_ptr() == sd["gnn.att_blocks.3.head_reducer.weight"].data_ptr()
    assert m.field_of_view == 10 and sd["gnn.pre_dense.0.weight"].shape == (512, 85)
    # Lightning checkpoints prefix keys with 'model.0.' (training/lightning_model.py); stripping it must load
    m.load_state_dict({k: v for k, v in {("model.0." + k)[8:]: v for k, v in sd.items()}.items()})
    gnn10 = models.GrappaGNN(n_conv=2, n_att=1, out_feats=32, node_feats=64, n_heads=4)     # grappa-1.0 layout
    assert "conv_blocks.1.graph_module.fc_neigh.weight" in gnn10.state_dict() and len(gnn10.blocks) == 3
    # learnable_statistics turns the statistics into parameters under the same state_dict names; layer_norm=False /
    # self_interaction=False drop the corresponding sub-modules (reference final_layer.py:37-39, graph_attention.py:257-274)
    tp = models.ToPositive(1.0, 0.5, learnable_statistics=True)
    assert sorted(k for k, _ in tp.named_parameters()) == ["mean_over_std", "std"] and float(tp.mean_over_std) == 2.0
    assert sorted(models.ToPositive(1.0, 0.5).state_dict()) == sorted(tp.state_dict())
    blk = models.ResidualAttentionBlock(64, num_heads=4, layer_norm=False, self_interaction=False)
    assert sorted(blk.state_dict()) == ["graph_module.fc.weight", "head_reducer.bias", "head_reducer.weight"]
    g = synthetic.dipeptide(seed=0, n_confs=2)
    with pytest.raises(ValueError):
        Energy(terms="n2")
    with pytest.raises(GrappaB200Error):
        m(g)                                   # CPU tensors: no fallback
    del g.nodes["n1"].data["xyz"]
    with pytest.raises(ValueError):
        Energy()(g)


def test_flat_params_are_views_and_buckets_tile_the_buffer():
    import grappa_oracle as orc
    from grappa_b200 import models
    from grappa_b200.training import FlatParams
    m = models.model_from_config(orc.small_model_config())
    before = {k: v.clone() for k, v in m.state_dict().items()}
    fp = FlatParams(m, "cpu")
    for k, v in m.state_dict().items():
        assert torch.equal(v, before[k])
    p = m.gnn.att_blocks[0].head_reducer.weight
    assert p.data_ptr() >= fp.flat.data_ptr() and p._gb_sink.data_ptr() >= fp.grad.data_ptr()
    spans = [fp.span(w) for w in (m.parameter_writer.bond_writer, m.parameter_writer.angle_writer,
                                  m.parameter_writer.proper_writer, m.parameter_writer.improper_writer)]
    spans.append(fp.span(m.gnn))
    spans.sort()
    assert spans[0][0] == 0 and spans[-1][1] == fp.total
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    fp.flat.zero_()
    assert all(float(q.abs().sum()) == 0 for q in m.parameters())


def test_trainer_checkpoint_roundtrip_and_optimizer_restart():
    """Host bookkeeping of training.Trainer on CPU (no kernels): the flat buffers stay the parameters' storage through
    load_state_dict, Adam moments / counters / lr round-trip by parameter name, reset_optimizer clears exactly the
    optimizer state (reference training/lightning_model.py:142-150, 300-304)."""
    import grappa_oracle as orc
    from grappa_b200 import models
    from grappa_b200.training import Trainer
    torch.manual_seed(0)
    tr = Trainer(models.model_from_config(orc.small_model_config()), None, None, lr=3e-4, device="cpu", distributed=False)
    tr.fp.m.normal_()
    tr.fp.v.uniform_()
    tr.counters[0], tr.counters[1] = 17, 4242
    tr._host_steps = 17
    ck = tr.state_dict()
    assert set(ck["exp_avg"]) == {k for k, _ in tr.model.named_parameters()}
    flat0, m0, v0 = tr.fp.flat.clone(), tr.fp.m.clone(), tr.fp.v.clone()
    torch.manual_seed(1)
    tr2 = Trainer(models.model_from_config(orc.small_model_config()), None, None, lr=1.0, device="cpu", distributed=False)
    assert not torch.equal(tr2.fp.flat, flat0)
    ptr = tr2.fp.flat.data_ptr()
    tr2.load_state_dict(ck)
    real = torch.zeros_like(flat0, dtype=torch.bool)                    # the slices are padded to 16-byte multiples
    for q, o in zip(tr.fp.params, tr.fp.offsets):
        real[o:o + q.numel()] = True
    assert torch.equal(tr2.fp.flat, flat0) and torch.equal(tr2.fp.m[real], m0[real]) and torch.equal(tr2.fp.v[real], v0[real])
    assert tr2.counters.tolist() == [17, 4242] and tr2.lr == 3e-4 and float(tr2.lr_dev) == pytest.approx(3e-4)
    assert tr2.step_count == 17 and tr2.fp.flat.data_ptr() == ptr
    p = next(tr2.model.parameters())
    assert p.data_ptr() == ptr + 4 * tr2.fp.offsets[0]                 # still a view of the flat buffer
    tr2.reset_optimizer(lr=1e-5)
    assert float(tr2.fp.m.abs().max()) == 0 and float(tr2.fp.v.abs().max()) == 0
    assert tr2.counters.tolist() == [0, 4242] and tr2.lr == 1e-5 and torch.equal(tr2.fp.flat, flat0)
