"""Host-side data path: collate / conformation handling / packed storage / loader against the
fixture the reference's own set_number_confs + batch + MolwiseLoss produced (tests/golden/ragged_confs.npz)."""
import numpy as np
import pytest
import torch

import grappa_oracle as orc
from util import LEVELS, graph_from_fixture, load_golden, rel_err


def _mols(z):
    return [graph_from_fixture(z, prefix=f"mol{i}.") for i in range(int(z["meta.n_mols"]))]


def test_collate_matches_reference_set_number_confs_and_batch():
    from grappa_b200 import dataset
    z = load_golden("ragged_confs.npz")
    mols = _mols(z)
    g = dataset.collate(mols, conf_strategy="max", build_pack=False)
    assert np.array_equal(g.nodes["n1"].data["xyz"].numpy(), z["batched.xyz"])
    assert np.array_equal(g.nodes["g"].data["is_dummy"].numpy(), z["batched.is_dummy"])
    assert np.array_equal(g.nodes["g"].data["energy_ref"].numpy(), z["batched.energy_ref"])
    assert np.array_equal(g.nodes["n1"].data["gradient_ref"].numpy(), z["batched.gradient_ref"])
    assert np.array_equal(g.nodes["n4"].data["idxs"].numpy(), z["batched.n4.idxs"])
    assert g.nodes["g"].data["n_valid"].tolist() == z["meta.n_confs"].tolist()
    # conf_strategy semantics of the reference collate_fn (data/GraphDataLoader.py:52-66)
    confs = z["meta.n_confs"].tolist()
    assert dataset.batch_n_confs(confs, 4) == 4 and dataset.batch_n_confs(confs, 100) == max(confs)
    assert dataset.batch_n_confs(confs, "min") == min(confs) and dataset.batch_n_confs(confs, "mean") == int(np.mean(confs))
    with pytest.raises(ValueError):
        dataset.batch_n_confs(confs, "median")
    # sub-sampling keeps a subset without replacement, padding repeats the last conformation
    sub = dataset.conformation_indices(7, 3, np.random.default_rng(0))
    assert len(set(sub.tolist())) == 3 and sub.max() < 7
    assert dataset.conformation_indices(2, 5, None).tolist() == [0, 1, 1, 1, 1]


def test_oracle_loss_ignores_padding_like_the_reference():
    z = load_golden("ragged_confs.npz")
    from grappa_b200 import dataset
    g = dataset.collate(_mols(z), conf_strategy="max", build_pack=False)
    e = torch.from_numpy(z["in.energy"]).double().requires_grad_(True)
    gr = torch.from_numpy(z["in.gradient"]).double().requires_grad_(True)
    kp = torch.from_numpy(z["in.k_proper"]).double().requires_grad_(True)
    ki = torch.from_numpy(z["in.k_improper"]).double().requires_grad_(True)
    loss = orc.molwise_loss({"energy": e, "gradient": gr}, {"n4": {"k": kp}, "n4_improper": {"k": ki}}, g,
                            n_valid=z["meta.n_confs"].tolist())
    assert abs(float(loss) - float(z["out.loss"])) < 1e-6 * abs(float(z["out.loss"]))
    ge, gg, gkp, gki = torch.autograd.grad(loss, [e, gr, kp, ki])
    assert rel_err(ge.numpy(), z["grad.energy"]) < 1e-5 and rel_err(gg.numpy(), z["grad.gradient"]) < 1e-5
    assert rel_err(gkp.numpy(), z["grad.k_proper"]) < 1e-5 and rel_err(gki.numpy(), z["grad.k_improper"]) < 1e-5


def test_packed_dataset_roundtrip_and_loader(tmp_path):
    from grappa_b200 import dataset, synthetic
    rng = np.random.default_rng(3)
    mols = [synthetic.make_molecule(rng, kind, n_confs=c, **kw) for kind, c, kw in
            (("peptide", 6, dict(n_res=1)), ("small", 2, dict(n_atoms=9)), ("rna", 9, {}), ("small", 4, dict(n_atoms=30)),
             ("peptide", 5, dict(n_res=2)), ("small", 7, dict(n_atoms=14)))]
    ds = dataset.PackedDataset.from_graphs(mols, dsnames=[f"set{i % 2}" for i in range(len(mols))])
    ds.save(str(tmp_path / "ds"))
    ds2 = dataset.PackedDataset.load(str(tmp_path / "ds"), mmap=True)
    assert len(ds2) == len(mols) and ds2.dsname(3) == "set1"
    # a molecule comes back unchanged
    m = ds2.molecule(2)
    for nt in m.ntypes:
        for k, v in mols[2].nodes[nt].data.items():
            assert torch.equal(m.nodes[nt].data[k], v), (nt, k)
    assert torch.equal(m.edges()[0], mols[2].edges()[0])
    # flat-array collate == per-graph collate, including the random conformation subsets (same generator stream)
    idx = [4, 1, 2, 5]
    a = ds2.collate(idx, conf_strategy=5, rng=np.random.default_rng(7), build_pack=False)
    b = dataset.collate([mols[i] for i in idx], conf_strategy=5, rng=np.random.default_rng(7), build_pack=False)
    for nt in b.ntypes:
        assert a.batch_num_nodes(nt).tolist() == b.batch_num_nodes(nt).tolist()
        for k, v in b.nodes[nt].data.items():
            assert torch.equal(a.nodes[nt].data[k], v), (nt, k)
    assert torch.equal(a.edges()[0], b.edges()[0]) and torch.equal(a.edges()[1], b.edges()[1])
    assert a.nodes["n1"].data["xyz"].shape[1] == 5 and a.nodes["g"].data["n_valid"].tolist() == [5, 2, 5, 5]
    # loader: every batch of the sampler arrives, in order, with index tables and dataset names attached
    batches = list(dataset.batch_sampler(dataset.shard_indices(len(ds2), 0, 1), 2, np.random.default_rng(1)))
    got = list(dataset.PrefetchLoader(ds2, batches, conf_strategy="min", seed=0, depth=2, pin=False))
    assert len(got) == 3
    for g, idxs in zip(got, batches):
        assert g.batch_size == 2 and g.dsnames == [ds2.dsname(i) for i in idxs]
        assert g.num_nodes("n1") == sum(mols[i].num_nodes("n1") for i in idxs)
        assert g._pack_cache is not None
    # the data do not depend on the number of worker threads (per-batch generator streams), and a worker's error surfaces
    for w in (1, 3):
        again = list(dataset.PrefetchLoader(ds2, batches, conf_strategy=3, seed=5, depth=3, pin=False, workers=w))
        if w == 1:
            first = again
        else:
            for ga, gb_ in zip(first, again):
                assert torch.equal(ga.nodes["n1"].data["xyz"], gb_.nodes["n1"].data["xyz"])
                assert torch.equal(ga.nodes["g"].data["energy_ref"], gb_.nodes["g"].data["energy_ref"])
    with pytest.raises(IndexError):
        list(dataset.PrefetchLoader(ds2, [[0, 1], [0, 10 ** 6]], pin=False))
    assert dataset.shard_indices(10, 1, 4) == [1, 5, 9]


def test_conformation_shards_tile_the_conformation_axis():
    """SURVEY.md 8e, energy / force sweep with few molecules and many conformations: the ranks' conformation slices are
    contiguous, disjoint, cover every conformation once (uneven counts included), share topology and parameters, and the
    oracle's energies / forces of the shards concatenate to the unsharded result -- no collective on the data path."""
    from grappa_b200 import dataset, synthetic
    g = synthetic.peptide_batch(seed=3, batch_size=2, n_res=1, n_confs=7)
    gen = torch.Generator().manual_seed(0)
    for l in LEVELS:
        T = g.num_nodes(l)
        if l in ("n2", "n3"):
            g.nodes[l].data["k"] = 100 + 300 * torch.rand(T, generator=gen)
            g.nodes[l].data["eq"] = 1.2 + 0.6 * torch.rand(T, generator=gen)
        else:
            g.nodes[l].data["k"] = torch.randn(T, 3, generator=gen)
    world = 3
    shards = [dataset.shard_conformations(g, r, world) for r in range(world)]
    assert [s.nodes["n1"].data["xyz"].shape[1] for s in shards] == [2, 2, 3]
    for k in ("xyz", "gradient_ref"):
        assert torch.equal(torch.cat([s.nodes["n1"].data[k] for s in shards], dim=1), g.nodes["n1"].data[k])
        assert all(s.nodes["n1"].data[k].is_contiguous() for s in shards)
    assert torch.equal(torch.cat([s.nodes["g"].data["energy_ref"] for s in shards], dim=1), g.nodes["g"].data["energy_ref"])
    assert all(s.nodes["n4"].data["idxs"] is g.nodes["n4"].data["idxs"] for s in shards)

    def evaluate(graph):
        idxs = {l: graph.nodes[l].data["idxs"] for l in LEVELS}
        params = {l: {k: graph.nodes[l].data[k] for k in (("k", "eq") if l in ("n2", "n3") else ("k",))} for l in LEVELS}
        counts = {l: graph.batch_num_nodes(l) for l in LEVELS}
        return orc.energy_forward(graph.nodes["n1"].data["xyz"], idxs, params, counts)
    whole, parts = evaluate(g), [evaluate(s) for s in shards]
    # fp32 torch reductions vectorise differently for different shapes: equal up to rounding
    assert rel_err(torch.cat([p["energy"] for p in parts], dim=1).detach().numpy(), whole["energy"].detach().numpy()) < 1e-6
    assert rel_err(torch.cat([p["gradient"] for p in parts], dim=1).numpy(), whole["gradient"].numpy()) < 1e-6
    with pytest.raises(ValueError):
        dataset.shard_conformations(g, 3, 3)
