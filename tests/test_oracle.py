"""The CPU oracle (oracle/grappa_oracle.py) is pinned against fixtures generated from the UNMODIFIED
reference (tests/golden/make_golden.py) and -- when /root/reference is mounted -- against the reference
itself on fresh inputs."""
import numpy as np
import pytest
import torch

import grappa_oracle as orc
from util import sampled_gradient_errors, LEVELS, graph_from_fixture, load_golden, rel_err


def _sd(cfg, seed):
    from grappa_b200 import models, synthetic
    m = models.model_from_config(dict(cfg))
    return synthetic.deterministic_state_dict(m.state_dict(), seed=seed), m


def _compare(h, params, en, z, tol):
    assert rel_err(h.detach().numpy(), z["out.h"]) < tol
    for l in LEVELS:
        assert rel_err(params[l]["k"].detach().numpy(), z[f"out.{l}.k"]) < tol
        if l in ("n2", "n3"):
            assert rel_err(params[l]["eq"].detach().numpy(), z[f"out.{l}.eq"]) < tol
        assert rel_err(en["term_energy"][l].numpy(), z[f"out.g.energy_{l}"]) < 20 * tol
    assert rel_err(en["energy"].detach().numpy(), z["out.g.energy"]) < tol
    assert rel_err(en["gradient"].detach().numpy(), z["out.n1.gradient"]) < tol


def test_oracle_matches_reference_fixture_grappa12_dipeptide():
    z = load_golden("dipeptide_grappa12.npz")
    cfg = orc.grappa_1_2_model_config()
    sd, model = _sd(cfg, seed=3)
    assert sorted(sd.keys()) == list(z["meta.state_dict_keys"])          # 410 reference state_dict keys
    assert len(sd) == 410 and sum(p.numel() for p in model.parameters()) == 40_801_805
    g = graph_from_fixture(z)
    with torch.no_grad():
        pass
    h, params, en = orc.path_forward(sd, g, cfg)
    _compare(h, params, en, z, 2e-5)


def test_oracle_matches_reference_fixture_loss_and_gradients():
    z = load_golden("mixed_batch_small_model.npz")
    cfg = orc.small_model_config()
    sd, _ = _sd(cfg, seed=7)
    leaves = {k: v.requires_grad_(True) for k, v in sd.items() if ("grad." + k) in z.files}
    g = graph_from_fixture(z)
    h, params, en = orc.path_forward(sd, g, cfg, create_graph=True)
    _compare(h, params, en, z, 2e-5)
    loss = orc.molwise_loss(en, params, g)
    assert abs(float(loss) - float(z["out.loss"])) < 1e-5 * abs(float(z["out.loss"]))
    grads = torch.autograd.grad(loss, list(leaves.values()))
    for (k, _), gr in zip(leaves.items(), grads):
        assert rel_err(gr.numpy(), z["grad." + k]) < 1e-4, k


def test_oracle_energy_fixture_and_double_backward():
    z = load_golden("energy_mixed_batch.npz")
    g = graph_from_fixture(z)
    prm = {l: {n: torch.from_numpy(z[f"in.{l}.{n}"]).requires_grad_(True) for n in ("k", "eq") if f"in.{l}.{n}" in z.files}
           for l in LEVELS}
    idxs = {l: g.nodes[l].data["idxs"] for l in LEVELS}
    counts = {l: g.batch_num_nodes(l).tolist() for l in LEVELS}
    en = orc.energy_forward(g.nodes["n1"].data["xyz"], idxs, prm, counts, create_graph=True)
    assert rel_err(en["energy"].detach().numpy(), z["out.g.energy"]) < 1e-5
    assert rel_err(en["gradient"].detach().numpy(), z["out.n1.gradient"]) < 1e-5
    obj = (en["energy"] * torch.from_numpy(z["in.gE"])).sum() + (en["gradient"] * torch.from_numpy(z["in.gF"])).sum()
    leaves = [(l, n, prm[l][n]) for l in LEVELS for n in sorted(prm[l])]
    grads = torch.autograd.grad(obj, [t for _, _, t in leaves])
    for (l, n, _), gr in zip(leaves, grads):
        assert rel_err(gr.numpy(), z[f"grad.{l}.{n}"]) < 1e-4, (l, n)


def test_oracle_known_answers_geometry():
    """SURVEY.md appendix A.5 known answers."""
    p = lambda *v: torch.tensor(v, dtype=torch.float64)
    assert abs(float(orc.dihedral_angle(p(1, 0, 0), p(0, 0, 0), p(0, 0, 1), p(1, 0, 1)))) < 1e-12
    assert abs(float(orc.dihedral_angle(p(1, 0, 0), p(0, 0, 0), p(0, 0, 1), p(0, 1, 1))) + np.pi / 2) < 1e-12
    assert abs(float(orc.bond_angle(p(1, 0, 0), p(0, 0, 0), p(0, 1, 0))) - np.pi / 2) < 1e-12
    assert abs(float(orc.bond_length(p(1, 2, 2), p(0, 0, 0))) - 3.0) < 1e-12


def test_oracle_vs_live_reference_on_fresh_inputs():
    """Only where the reference sources are mounted (build container)."""
    from ref_import import import_reference, no_dihedral_noise, reference_available, to_reference_graph
    if not reference_available():
        pytest.skip("/root/reference not mounted (GPU box)")
    from grappa_b200 import synthetic
    ns = import_reference()
    cfg = orc.small_model_config()
    torch.manual_seed(0)
    ref = ns.deploy.model_from_config(dict(cfg), param_statistics=ns.graph_utils.get_default_statistics()).eval()
    sd = {k: v.clone() for k, v in ref.state_dict().items()}          # the reference's own random init
    # n_confs != 3: the reference's angle() calls torch.cross WITHOUT dim (internal_coordinates.py:159), which picks the
    # first axis of size 3 -- with exactly 3 conformations (or 3 angles) it crosses over the wrong axis.
    g = synthetic.espaloma_mix_batch(seed=21, batch_size=4, n_confs=5)
    dg = to_reference_graph(ns, g)
    with no_dihedral_noise():
        dg = torch.nn.Sequential(ref, ns.energy.Energy())(dg)
    h, params, en = orc.path_forward(sd, g, cfg)
    assert rel_err(h.detach().numpy(), dg.nodes["n1"].data["h"].detach().numpy()) < 2e-5
    assert rel_err(en["energy"].detach().numpy(), dg.nodes["g"].data["energy"].detach().numpy()) < 1e-5
    assert rel_err(en["gradient"].numpy(), dg.nodes["n1"].data["gradient"].detach().numpy()) < 1e-5
    for l in LEVELS:
        assert rel_err(params[l]["k"].detach().numpy(), dg.nodes[l].data["k"].detach().numpy()) < 2e-5


def test_oracle_param_loss_vs_reference_fixture():
    """Classical-parameter loss term: oracle restatement vs values / gradients the reference's MolwiseLoss produced."""
    z = load_golden("param_loss.npz")
    g = graph_from_fixture(z)
    counts = {l: g.batch_num_nodes(l).tolist() for l in LEVELS}
    dsw = {"spice": 0.5, "rna": 2.0}
    mw = [dsw.get(str(d), 1e-3) for d in z["meta.dsnames"]]
    for v in ("a", "b"):
        prm = {l: {n: torch.from_numpy(z[f"{v}.in.{l}.{n}"]).double().requires_grad_(True)
                   for n in ("k", "eq") if f"{v}.in.{l}.{n}" in z.files} for l in LEVELS}
        ref = {l: {n: torch.from_numpy(z[f"{v}.ref.{l}.{n}"]) for n in ("k", "eq") if f"{v}.ref.{l}.{n}" in z.files} for l in LEVELS}
        loss = orc.param_loss(prm, ref, counts, mol_weights=mw)
        assert abs(float(loss) - float(z[f"{v}.loss"])) < 1e-6 * abs(float(z[f"{v}.loss"]))
        uni = orc.param_loss(prm, ref, counts, param_weight=1e-3)
        assert abs(float(uni) - float(z[f"{v}.loss_uniform"])) < 1e-6 * abs(float(z[f"{v}.loss_uniform"]))
        keys = [k[len(v) + 6:] for k in z.files if k.startswith(f"{v}.grad.")]
        grads = torch.autograd.grad(loss, [prm[k.split(".")[0]][k.split(".")[1]] for k in keys])
        for k, gr in zip(keys, grads):
            assert rel_err(gr.numpy(), z[f"{v}.grad.{k}"]) < 1e-5, k


def test_oracle_conv_blocks_vs_live_reference():
    """grappa-1.0 style GNN (gnn_convolutions > 0: ResidualConvBlock around SAGEConv('mean')):
    oracle restatement vs the unmodified reference running on the dgl shim; also the state_dict key set."""
    from ref_import import import_reference, no_dihedral_noise, reference_available, to_reference_graph
    if not reference_available():
        pytest.skip("/root/reference not mounted (GPU box)")
    from grappa_b200 import models, synthetic
    ns = import_reference()
    cfg = orc.small_model_config()
    cfg.update(gnn_convolutions=2, gnn_attentional_layers=1, wrong_symmetry=True)   # + the improper-symmetry ablation switch
    torch.manual_seed(0)
    ref = ns.deploy.model_from_config(dict(cfg), param_statistics=ns.graph_utils.get_default_statistics()).eval()
    sd = {k: v.clone() for k, v in ref.state_dict().items()}
    ours = models.model_from_config(dict(cfg))
    assert sorted(ours.state_dict().keys()) == sorted(sd.keys())
    assert all(tuple(ours.state_dict()[k].shape) == tuple(v.shape) for k, v in sd.items())
    g = synthetic.espaloma_mix_batch(seed=4, batch_size=3, n_confs=4)
    dg = to_reference_graph(ns, g)
    with no_dihedral_noise():
        dg = torch.nn.Sequential(ref, ns.energy.Energy())(dg)
    h, params, en = orc.path_forward(sd, g, cfg)
    assert rel_err(h.detach().numpy(), dg.nodes["n1"].data["h"].detach().numpy()) < 2e-5
    assert rel_err(en["energy"].detach().numpy(), dg.nodes["g"].data["energy"].detach().numpy()) < 1e-5
    for l in LEVELS:
        assert rel_err(params[l]["k"].detach().numpy(), dg.nodes[l].data["k"].detach().numpy()) < 2e-5


def test_oracle_ablation_switches_vs_live_reference():
    """Constructor switches of GrappaModel that grappa-1.x leaves at their defaults (the whole argument list
    is part of the surface): layer_norm=False, self_interaction=False, learnable_statistics=True, ungated torsions.  The
    oracle restatement vs the unmodified reference on the dgl shim -- forward, state_dict key set (the statistics become
    parameters) and the gradients of the learnable statistics."""
    from ref_import import import_reference, no_dihedral_noise, reference_available, to_reference_graph
    if not reference_available():
        pytest.skip("/root/reference not mounted (GPU box)")
    from grappa_b200 import models, synthetic
    ns = import_reference()
    for switches in (dict(layer_norm=False, learnable_statistics=True, gated_torsion=False),
                     dict(self_interaction=False, learnable_statistics=True, gated_torsion=True, gnn_convolutions=1)):
        cfg = orc.small_model_config()
        cfg.update(switches)
        torch.manual_seed(0)
        ref = ns.deploy.model_from_config(dict(cfg), param_statistics=ns.graph_utils.get_default_statistics()).eval()
        sd = {k: v.clone() for k, v in ref.state_dict().items()}
        ours = models.model_from_config(dict(cfg))
        assert sorted(ours.state_dict().keys()) == sorted(sd.keys())
        assert all(tuple(ours.state_dict()[k].shape) == tuple(v.shape) for k, v in sd.items())
        assert sorted(k for k, _ in ours.named_parameters()) == sorted(k for k, _ in ref.named_parameters())
        g = synthetic.espaloma_mix_batch(seed=8, batch_size=3, n_confs=4)
        dg = to_reference_graph(ns, g)
        with no_dihedral_noise():
            dg = torch.nn.Sequential(ref, ns.energy.Energy())(dg)
        stat_keys = [k for k, _ in ref.named_parameters()
                     if k.rsplit(".", 1)[-1] in ("mean_over_std", "std", "std_over_max", "k_mean", "k_std")]
        assert len(stat_keys) == 4 + 3 + 2 + 2
        sd_o = {k: (v.clone().requires_grad_(True) if k in stat_keys else v) for k, v in sd.items()}
        h, params, en = orc.path_forward(sd_o, g, cfg, create_graph=True)
        assert rel_err(h.detach().numpy(), dg.nodes["n1"].data["h"].detach().numpy()) < 2e-5
        assert rel_err(en["energy"].detach().numpy(), dg.nodes["g"].data["energy"].detach().numpy()) < 1e-5
        for l in LEVELS:
            assert rel_err(params[l]["k"].detach().numpy(), dg.nodes[l].data["k"].detach().numpy()) < 2e-5
        # a scalar that touches energies, forces and every parameter type
        def scalar(energy, gradient, p):
            return (energy ** 2).mean() + 1e-2 * (gradient ** 2).mean() + sum((p[l]["k"] ** 2).mean() for l in LEVELS) \
                + p["n2"]["eq"].sum() + p["n3"]["eq"].sum()
        ref_params = {l: {k: dg.nodes[l].data[k] for k in (("k", "eq") if l in ("n2", "n3") else ("k",))} for l in LEVELS}
        named = dict(ref.named_parameters())
        g_ref = torch.autograd.grad(scalar(dg.nodes["g"].data["energy"], dg.nodes["n1"].data["gradient"], ref_params),
                                    [named[k] for k in stat_keys], allow_unused=True)
        g_orc = torch.autograd.grad(scalar(en["energy"], en["gradient"], params), [sd_o[k] for k in stat_keys],
                                    allow_unused=True)
        for k, a, b in zip(stat_keys, g_orc, g_ref):
            if b is None:                       # gated torsions: k_mean does not enter (interaction_parameters.py:546-550)
                assert a is None and k.endswith("k_mean") and cfg["gated_torsion"]
                continue
            assert rel_err(a.numpy(), b.numpy()) < 1e-4, k


@pytest.mark.parametrize("case,switches", [
    ("noln_learnable_ungated", dict(layer_norm=False, learnable_statistics=True, gated_torsion=False)),
    ("noselfint_learnable_conv", dict(self_interaction=False, learnable_statistics=True, gated_torsion=True, gnn_convolutions=1)),
])
def test_oracle_matches_reference_fixture_constructor_switches(case, switches):
    """tests/golden/switches_small_model.npz (make_golden.py --only-switches, unmodified reference): outputs, loss and the
    gradients of the learnable statistics for the constructor switches -- available on the GPU box too, where the live
    reference is not."""
    z = load_golden("switches_small_model.npz")
    cfg = orc.small_model_config()
    cfg.update(switches)
    sd, ours = _sd(cfg, seed=31)
    assert sorted(sd.keys()) == list(z[f"{case}.meta.state_dict_keys"])
    assert sorted(k for k, _ in ours.named_parameters()) == list(z[f"{case}.meta.parameter_names"])
    pre = case + ".grad."
    leaves = {k: v.requires_grad_(True) for k, v in sd.items() if (pre + k) in z.files}
    assert len(leaves) >= 12
    g = graph_from_fixture(z)
    h, params, en = orc.path_forward(sd, g, cfg, create_graph=True)
    assert rel_err(h.detach().numpy(), z[f"{case}.out.h"]) < 2e-5
    for l in LEVELS:
        assert rel_err(params[l]["k"].detach().numpy(), z[f"{case}.out.{l}.k"]) < 2e-5
        if l in ("n2", "n3"):
            assert rel_err(params[l]["eq"].detach().numpy(), z[f"{case}.out.{l}.eq"]) < 2e-5
    assert rel_err(en["energy"].detach().numpy(), z[f"{case}.out.g.energy"]) < 2e-5
    assert rel_err(en["gradient"].detach().numpy(), z[f"{case}.out.n1.gradient"]) < 2e-5
    loss = orc.molwise_loss(en, params, g)
    assert abs(float(loss) - float(z[f"{case}.out.loss"])) < 1e-5 * abs(float(z[f"{case}.out.loss"]))
    grads = torch.autograd.grad(loss, list(leaves.values()), allow_unused=True)
    for (k, _), gr in zip(leaves.items(), grads):
        if bool(z[f"{case}.gradnone.{k}"]):
            assert gr is None, k               # gated torsions: k_mean does not enter
        else:
            assert rel_err(gr.numpy(), z[pre + k]) < 1e-4, k


def test_graph_convolutions_vs_dense_adjacency_restatement_fp64():
    """DGL is absent from /root/reference and unpinned: the two DGL layers on the path are restated twice,
    independently -- edge-list form (oracle.dot_gat, oracle/dgl_shim's DotGatConv / SAGEConv, which the unmodified
    reference runs on here) and a dense N x N adjacency form written from the published definitions: DotGatConv =
    softmax over the in-neighbours of <ft_u, ft_v> / sqrt(d) per head, aggregation of ft_u; SAGEConv('mean') =
    fc_self(h) + fc_neigh(mean of in-neighbour features).  fp64, random graphs with the bonded-graph structure (both
    directions of every bond, no self loops, no isolated atoms)."""
    import os
    import sys
    shim = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "dgl_shim")
    from grappa_b200 import synthetic
    rng = np.random.default_rng(2)
    g = synthetic.make_molecule(rng, "rna", n_confs=1, n_atoms=60)
    src, dst = [t.long() for t in g.edges()]
    n, heads, d = g.num_nodes("n1"), 4, 8
    torch.manual_seed(0)
    A = torch.zeros(n, n, dtype=torch.bool)
    A[dst, src] = True                                         # A[v, u]: edge u -> v
    assert not A.diagonal().any() and A.any(1).all()
    ft = torch.randn(n, heads, d, dtype=torch.float64)
    scores = torch.einsum("vhd,uhd->hvu", ft, ft) / d ** 0.5
    scores = scores.masked_fill(~A[None], -float("inf"))
    dense = torch.einsum("hvu,uhd->vhd", torch.softmax(scores, dim=-1), ft)
    assert rel_err(orc.dot_gat(ft, src, dst).numpy(), dense.numpy()) < 1e-13
    sys.path.insert(0, shim)
    try:
        saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "dgl" or k.startswith("dgl.")}
        from dgl.nn.pytorch.conv import DotGatConv, SAGEConv

        class G:                                               # the only graph method the layers use
            @staticmethod
            def edges():
                return src, dst
        h = torch.randn(n, 24, dtype=torch.float64)
        conv = DotGatConv(24, d, heads).double()
        ftc = conv.fc(h).view(n, heads, d)
        sc = (torch.einsum("vhd,uhd->hvu", ftc, ftc) / d ** 0.5).masked_fill(~A[None], -float("inf"))
        want = torch.einsum("hvu,uhd->vhd", torch.softmax(sc, dim=-1), ftc)
        assert rel_err(conv(G, h).detach().numpy(), want.detach().numpy()) < 1e-13
        assert [k for k, _ in conv.named_parameters()] == ["fc.weight"]          # DGL's parameter name
        sage = SAGEConv(24, 16, "mean").double()
        mean = (A.double() @ h) / A.sum(1, keepdim=True)
        want = sage.fc_self(h) + sage.fc_neigh(mean)
        assert rel_err(sage(G, h).detach().numpy(), want.detach().numpy()) < 1e-13
    finally:
        sys.path.remove(shim)
        for k in [k for k in sys.modules if k == "dgl" or k.startswith("dgl.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_oracle_matches_reference_fixture_full_size_training_batch():
    """BASELINE configs[1] at full size (32 x ACE-(ALA)4-NME, 50 conformations, grappa-1.2): energies, parameters, loss
    and per-parameter gradient norms of the unmodified reference (tests/golden/train_batch_grappa12.npz; the inputs are
    regenerated from the seed and guarded by a checksum)."""
    from grappa_b200 import synthetic
    z = load_golden("train_batch_grappa12.npz")
    g = synthetic.peptide_batch(seed=int(z["meta.seed"]), batch_size=32, n_res=4, n_confs=50)
    xyz = g.nodes["n1"].data["xyz"].double()
    assert abs(float(xyz.sum()) - float(z["meta.xyz_checksum"])) < 1e-6 * float(z["meta.xyz_abs_checksum"])
    assert abs(float(xyz.abs().sum()) - float(z["meta.xyz_abs_checksum"])) < 1e-9 * float(z["meta.xyz_abs_checksum"])
    cfg = orc.grappa_1_2_model_config()
    sd, _ = _sd(cfg, seed=int(z["meta.weights_seed"]))
    keys = list(z["meta.grad_norms_keys"])
    leaves = {k: sd[k].requires_grad_(True) for k in keys}
    h, params, en = orc.path_forward(sd, g, cfg, create_graph=True)
    assert rel_err(en["energy"].detach().numpy(), z["out.g.energy"]) < 2e-5
    assert rel_err(en["gradient"].detach().norm(dim=(1, 2)).numpy(), z["out.gradient_norm_per_atom"]) < 2e-5
    for l in LEVELS:
        assert rel_err(params[l]["k"].detach().numpy(), z[f"out.{l}.k"]) < 2e-5, l
        if l in ("n2", "n3"):
            assert rel_err(params[l]["eq"].detach().numpy(), z[f"out.{l}.eq"]) < 2e-5, l
    loss = orc.molwise_loss(en, params, g)
    assert abs(float(loss) - float(z["out.loss"])) < 1e-5 * abs(float(z["out.loss"]))
    grads = dict(zip(leaves, torch.autograd.grad(loss, list(leaves.values()), allow_unused=True)))
    norms = np.array([0.0 if grads[k] is None else float(grads[k].norm()) for k in keys])
    ref = z["meta.grad_norms"]
    rel = np.abs(norms - ref) / np.maximum(ref, 1e-6 * ref.max())
    assert rel.max() < 1e-3, (keys[int(rel.argmax())], rel.max())
    for k in z.files:
        if k.startswith("grad."):
            assert rel_err(grads[k[5:]].numpy(), z[k]) < 1e-4, k
    # 16 sampled entries of every gradient tensor, relative to the tensor's max |grad|
    worst = sampled_gradient_errors(z, {k: (None if grads[k] is None else grads[k].numpy()) for k in keys})
    assert max(worst.values()) < 1e-4, max(worst.items(), key=lambda kv: kv[1])


def test_oracle_matches_reference_fixtures_protein_and_espaloma_mix():
    """BASELINE configs[2] (1,502-atom protein, grappa-1.2 parametrisation) and configs[4] (Espaloma-shaped mix, narrow
    model, energies + forces) at full size: oracle vs outputs of the unmodified reference, same seeds as
    tests/test_baseline_configs_gpu.py (which compares the CUDA path with the oracle on the same inputs)."""
    from grappa_b200 import synthetic
    z = load_golden("protein_grappa12.npz")
    g = synthetic.protein(seed=int(z["meta.seed"]))
    assert abs(float(g.nodes["n1"].data["partial_charge"].double().abs().sum()) - float(z["meta.charge_checksum"])) < 1e-9
    cfg = orc.grappa_1_2_model_config()
    sd, _ = _sd(cfg, seed=int(z["meta.weights_seed"]))
    with torch.no_grad():
        h, params = orc.model_forward(sd, g, cfg)
    assert rel_err(h.norm(dim=1).numpy(), z["out.h_norm_per_atom"]) < 2e-5
    assert rel_err(h[:8].numpy(), z["out.h_first_atoms"]) < 2e-5
    for l in LEVELS:
        assert params[l]["k"].shape == z[f"out.{l}.k"].shape
        assert rel_err(params[l]["k"].numpy(), z[f"out.{l}.k"]) < 2e-5, l
        if l in ("n2", "n3"):
            assert rel_err(params[l]["eq"].numpy(), z[f"out.{l}.eq"]) < 2e-5, l

    z = load_golden("espaloma_mix_small_model.npz")
    g = synthetic.espaloma_mix_batch(seed=int(z["meta.seed"]), batch_size=32, n_confs=32)
    assert g.batch_num_nodes("n1").tolist() == z["meta.atom_counts"].tolist()
    xyz = g.nodes["n1"].data["xyz"].double()
    assert abs(float(xyz.abs().sum()) - float(z["meta.xyz_abs_checksum"])) < 1e-9 * float(z["meta.xyz_abs_checksum"])
    cfg = orc.small_model_config()
    sd, _ = _sd(cfg, seed=int(z["meta.weights_seed"]))
    h, params, en = orc.path_forward(sd, g, cfg)
    assert rel_err(en["energy"].detach().numpy(), z["out.g.energy"]) < 2e-5
    assert rel_err(en["gradient"].detach().numpy(), z["out.n1.gradient"]) < 2e-5
    for l in LEVELS:
        assert rel_err(params[l]["k"].detach().numpy(), z[f"out.{l}.k"]) < 2e-5, l
        if l in ("n2", "n3"):
            assert rel_err(params[l]["eq"].detach().numpy(), z[f"out.{l}.eq"]) < 2e-5, l
