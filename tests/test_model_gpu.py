"""Whole-path parity on the GPU: GrappaModel + Energy (+ MolwiseLoss and its gradients) through the
C ABI versus fixtures generated from the UNMODIFIED reference (tests/golden/make_golden.py).

Tolerances (BASELINE.json north_star): parameters / energies / forces 1e-5 relative in fp32 (FFMA
GEMMs), 1e-3 with tensor-core GEMMs, gradients 1e-4.  'relative' = max|a-b| / max|b| per tensor.
"""
import numpy as np
import pytest
import torch

from util import LEVELS, graph_from_fixture, load_golden, rel_err

pytestmark = pytest.mark.gpu


def _model(cfg, seed):
    from grappa_b200 import models, synthetic
    m = models.model_from_config(dict(cfg))
    m.load_state_dict(synthetic.deterministic_state_dict(m.state_dict(), seed=seed))
    return m.cuda()


def _check_forward(g, z, tol):
    """Every output tensor of the path against the reference fixture at ONE tolerance (max |a - b| / max |b| per tensor) --
    no per-tensor carve-outs: h, all k / eq (the gated torsion amplitudes included), energy, gradient."""
    errs = {"h": rel_err(g.nodes["n1"].data["h"].detach().cpu().numpy(), z["out.h"])}
    for l in LEVELS:
        errs[f"{l}.k"] = rel_err(g.nodes[l].data["k"].detach().cpu().numpy(), z[f"out.{l}.k"])
        if l in ("n2", "n3"):
            errs[f"{l}.eq"] = rel_err(g.nodes[l].data["eq"].detach().cpu().numpy(), z[f"out.{l}.eq"])
    errs["energy"] = rel_err(g.nodes["g"].data["energy"].detach().cpu().numpy(), z["out.g.energy"])
    errs["gradient"] = rel_err(g.nodes["n1"].data["gradient"].detach().cpu().numpy(), z["out.n1.gradient"])
    bad = {k: v for k, v in errs.items() if v > tol}
    assert not bad, f"relative errors above tolerance {tol}: {bad} (all: {errs})"
    return errs


# Tolerance of the tensor-core arithmetic bench.py runs (ops.BENCH_PRECISION = bf16x3): north_star allows 1e-3 where
# reduced-precision GEMMs are used; the split-operand GEMMs reach 1e-4 on every output and on the gradients, so that is
# what is asserted.  (The raw TF32 mode misses 1e-3 on the gated torsion amplitudes -- profiles/r2_error_budget.md -- and is
# no longer a model-level mode under test; its GEMM kernel is still unit-tested in tests/test_ops_gpu.py.)
TENSOR_TOL = 1e-4


def test_grappa12_dipeptide_matches_reference_fp32():
    """BASELINE config 1: grappa-1.2 architecture, capped dipeptide, 50 conformations."""
    import grappa_oracle as orc
    from grappa_b200 import ops
    from grappa_b200.energy import Energy
    z = load_golden("dipeptide_grappa12.npz")
    ops.set_matmul_precision("fp32")
    model = _model(orc.grappa_1_2_model_config(), seed=3).eval()
    assert sorted(model.state_dict().keys()) == list(z["meta.state_dict_keys"])
    shapes = [",".join(map(str, model.state_dict()[k].shape)) for k in sorted(model.state_dict().keys())]
    assert shapes == list(z["meta.state_dict_shapes"])
    g = graph_from_fixture(z).to("cuda")
    with torch.no_grad():
        g = torch.nn.Sequential(model, Energy())(g)
    errs = _check_forward(g, z, 1e-5)
    print("fp32 relative errors:", errs)
    for l in LEVELS:   # per-term energies and internal coordinates are written like the reference does
        # torsion terms are sums of ~40 cancelling contributions (|sum| ~ 0.1 vs |k| ~ 1): 1e-5 on k gives ~1e-4 here
        tol = 1e-5 if l in ("n2", "n3") else 2e-4
        assert rel_err(g.nodes["g"].data[f"energy_{l}"].cpu().numpy(), z[f"out.g.energy_{l}"]) < tol
        assert g.nodes[l].data["x"].shape == z[f"out.{l}.x"].shape


def test_grappa12_dipeptide_matches_reference_tensor_cores():
    """Same fixture through the tcgen05 GEMMs at the arithmetic bench.py measures (bf16x3 split operands)."""
    import grappa_oracle as orc
    from grappa_b200 import ops
    from grappa_b200.energy import Energy
    z = load_golden("dipeptide_grappa12.npz")
    ops.set_matmul_precision(ops.BENCH_PRECISION)
    try:
        model = _model(orc.grappa_1_2_model_config(), seed=3).eval()
        g = graph_from_fixture(z).to("cuda")
        with torch.no_grad():
            g = torch.nn.Sequential(model, Energy())(g)
        errs = _check_forward(g, z, TENSOR_TOL)
        print(ops.BENCH_PRECISION, "relative errors:", errs)
    finally:
        ops.set_matmul_precision("fp32")


def test_small_model_gradients_tensor_cores():
    import grappa_oracle as orc
    from grappa_b200 import ops
    from grappa_b200.energy import Energy
    from grappa_b200.loss import MolwiseLoss
    z = load_golden("mixed_batch_small_model.npz")
    ops.set_matmul_precision(ops.BENCH_PRECISION)
    try:
        model = _model(orc.small_model_config(), seed=7).eval()
        g = graph_from_fixture(z).to("cuda")
        g = torch.nn.Sequential(model, Energy(write_tuple_terms=False))(g)
        _check_forward(g, z, TENSOR_TOL)
        loss = MolwiseLoss(gradient_weight=0.8, energy_weight=1.0, param_weight=1e-3, proper_regularisation=1e-3,
                           improper_regularisation=1e-3)(g)
        assert abs(loss.item() - float(z["out.loss"])) < TENSOR_TOL * abs(float(z["out.loss"]))
        loss.backward()
        named = dict(model.named_parameters())
        worst = 0.0
        for k in z.files:
            if k.startswith("grad."):
                worst = max(worst, rel_err(named[k[5:]].grad.cpu().numpy(), z[k]))
        print(ops.BENCH_PRECISION, "worst picked-gradient error", worst)
        assert worst < 2 * TENSOR_TOL
    finally:
        ops.set_matmul_precision("fp32")


def test_small_model_mixed_batch_loss_and_gradients_fp32():
    import grappa_oracle as orc
    from grappa_b200 import ops
    from grappa_b200.energy import Energy
    from grappa_b200.loss import MolwiseLoss
    z = load_golden("mixed_batch_small_model.npz")
    ops.set_matmul_precision("fp32")
    model = _model(orc.small_model_config(), seed=7).eval()
    g = graph_from_fixture(z).to("cuda")
    g = torch.nn.Sequential(model, Energy(write_tuple_terms=False))(g)
    _check_forward(g, z, 1e-5)
    loss = MolwiseLoss(gradient_weight=0.8, energy_weight=1.0, param_weight=1e-3, proper_regularisation=1e-3,
                       improper_regularisation=1e-3)(g)
    assert abs(loss.item() - float(z["out.loss"])) < 1e-5 * abs(float(z["out.loss"]))
    loss.backward()
    named = dict(model.named_parameters())
    worst = 0.0
    for k in z.files:
        if k.startswith("grad."):
            name = k[5:]
            e = rel_err(named[name].grad.cpu().numpy(), z[k])
            worst = max(worst, e)
            assert e < 1e-4, (name, e)
    keys = list(z["meta.grad_norms_keys"])
    norms = np.array([float(named[k].grad.norm()) if named[k].grad is not None else 0.0 for k in keys])
    ref = z["meta.grad_norms"]
    rel = np.abs(norms - ref) / np.maximum(ref, 1e-6 * ref.max())
    assert rel.max() < 1e-3, (keys[int(rel.argmax())], rel.max())
    print("worst picked-gradient error", worst, "worst norm error", rel.max())


def test_batch_invariance_and_permutation_symmetry():
    """Properties the reference tests (tests/unbatch.py: batch-size invariance) and the architecture
    guarantees (k invariant under the writer's permutations)."""
    import grappa_oracle as orc
    from grappa_b200 import graph as gbg, ops, synthetic
    ops.set_matmul_precision("fp32")
    model = _model(orc.small_model_config(), seed=1).eval()
    rng = np.random.default_rng(3)
    mols = [synthetic.make_molecule(rng, "peptide", n_confs=2, n_res=1), synthetic.make_molecule(rng, "small", n_confs=2, n_atoms=17),
            synthetic.make_molecule(rng, "rna", n_confs=2, n_atoms=91)]
    with torch.no_grad():
        gb = model(gbg.batch(mols).to("cuda"))
        singles = [model(m.to("cuda")) for m in mols]
    for l in LEVELS:
        cat = torch.cat([s.nodes[l].data["k"] for s in singles])
        assert rel_err(gb.nodes[l].data["k"].cpu().numpy(), cat.cpu().numpy()) < 1e-5
    # reversing a proper torsion / swapping the outer atoms of an improper leaves its parameters unchanged
    m = mols[0]
    m2 = gbg.batch([m])
    m2.nodes["n4"].data["idxs"] = m.nodes["n4"].data["idxs"].flip(1).contiguous()
    m2.nodes["n4_improper"].data["idxs"] = m.nodes["n4_improper"].data["idxs"][:, [3, 1, 2, 0]].contiguous()
    m2.nodes["n3"].data["idxs"] = m.nodes["n3"].data["idxs"].flip(1).contiguous()
    m2.nodes["n2"].data["idxs"] = m.nodes["n2"].data["idxs"].flip(1).contiguous()
    with torch.no_grad():
        a, b = singles[0], model(m2.to("cuda"))
    for l in LEVELS:
        assert rel_err(b.nodes[l].data["k"].cpu().numpy(), a.nodes[l].data["k"].cpu().numpy()) < 1e-5


def test_train_mode_dropout_gradients_are_consistent():
    """With dropout on, the backward pass regenerates the forward masks: the directional derivative of
    the loss (central differences, same seed) matches <grad, direction>."""
    import grappa_oracle as orc
    from grappa_b200 import ops, synthetic
    from grappa_b200.energy import Energy
    from grappa_b200.loss import MolwiseLoss
    ops.set_matmul_precision("fp32")
    model = _model(orc.small_model_config(), seed=2).train()
    g0 = synthetic.peptide_batch(seed=5, batch_size=3, n_res=1, n_confs=4).to("cuda")
    loss_fn = MolwiseLoss(proper_regularisation=1e-3, improper_regularisation=1e-3)
    energy = Energy(write_tuple_terms=False)

    def run():
        torch.manual_seed(123)
        g = g0.to("cuda")
        return loss_fn(energy(model(g)))

    l1, l2 = run(), run()
    assert l1.item() == l2.item()                     # same seed -> same masks
    model.eval()
    l_eval = run()
    model.train()
    assert abs(l_eval.item() - l1.item()) > 1e-6 * abs(l1.item())   # dropout is really active
    model.zero_grad()
    run().backward()
    p = model.gnn.att_blocks[1].self_interaction[2].weight
    q = model.parameter_writer.angle_writer.angle_model.grappa_transformer.transformer[0].attn.out_proj.weight
    for w in (p, q):
        # direction of steepest ascent: the loss is an fp32 number of order 1e4-1e5 (ulp ~ 1e-2), a random direction
        # changes it by about one ulp at this step size and the difference quotient is pure rounding noise
        d = w.grad / w.grad.norm()
        analytic = float((w.grad * d).sum())
        eps = 1e-2
        with torch.no_grad():
            w.add_(eps * d); lp = run().item()
            w.sub_(2 * eps * d); lm = run().item()
            w.add_(eps * d)
        fd = (lp - lm) / (2 * eps)
        assert abs(fd - analytic) < 5e-2 * max(abs(fd), abs(analytic), 1e-3), (fd, analytic)


def test_product_path_has_no_cpu_fallback():
    import grappa_oracle as orc
    from grappa_b200 import GrappaB200Error, models, synthetic
    model = models.model_from_config(orc.small_model_config())
    with pytest.raises(GrappaB200Error):
        model(synthetic.dipeptide(seed=0, n_confs=2))          # CPU graph -> loud failure


def test_conv_block_model_matches_oracle_forward_and_gradients():
    """GNN with SAGEConv('mean') convolution blocks (grappa-1.0 layout) -- forward and parameter
    gradients against the CPU oracle (which tests/test_oracle.py pins against the live reference on the dgl shim)."""
    import grappa_oracle as orc
    from grappa_b200 import ops, synthetic
    from grappa_b200.energy import Energy
    from grappa_b200.loss import MolwiseLoss
    ops.set_matmul_precision("fp32")
    cfg = orc.small_model_config()
    cfg.update(gnn_convolutions=2, gnn_attentional_layers=1, wrong_symmetry=True)   # + the improper-symmetry ablation switch
    model = _model(cfg, seed=11).eval()
    g = synthetic.espaloma_mix_batch(seed=6, batch_size=4, n_confs=5)
    sd = {k: v.detach().cpu().double().requires_grad_(v.is_floating_point() and "conv_blocks" in k)
          for k, v in model.state_dict().items()}
    h, params, en = orc.path_forward(sd, g, cfg, dtype=torch.float64, create_graph=True)
    ref_loss = orc.molwise_loss(en, params, g)
    leaves = {k: v for k, v in sd.items() if v.requires_grad}
    ref_grads = dict(zip(leaves, torch.autograd.grad(ref_loss, list(leaves.values()))))
    gd = torch.nn.Sequential(model, Energy(write_tuple_terms=False))(g.to("cuda"))
    assert rel_err(gd.nodes["n1"].data["h"].detach().cpu().numpy(), h.detach().numpy()) < 1e-5
    assert rel_err(gd.nodes["g"].data["energy"].detach().cpu().numpy(), en["energy"].detach().numpy()) < 1e-5
    for l in LEVELS:
        assert rel_err(gd.nodes[l].data["k"].detach().cpu().numpy(), params[l]["k"].detach().numpy()) < 1e-5, l
    loss = MolwiseLoss(gradient_weight=0.8, energy_weight=1.0, param_weight=0.0, proper_regularisation=1e-3,
                       improper_regularisation=1e-3)(gd)
    assert abs(loss.item() - float(ref_loss)) < 1e-5 * abs(float(ref_loss))
    model.zero_grad()
    loss.backward()
    named = dict(model.named_parameters())
    checked = 0
    for k, gr in ref_grads.items():
        if k in named:      # `blocks.*` aliases of conv_blocks.* share the parameter
            assert rel_err(named[k].grad.cpu().numpy(), gr.numpy()) < 1e-4, k
            checked += 1
    assert checked >= 16


@pytest.mark.parametrize("switches", [
    dict(layer_norm=False, learnable_statistics=True, gated_torsion=False),
    dict(self_interaction=False, learnable_statistics=True, gated_torsion=True, gnn_convolutions=1),
])
def test_constructor_switches_match_oracle_forward_and_gradients(switches):
    """the whole GrappaModel argument list is part of the surface -- layer_norm=False (blocks, transformer
    layers and symmetriser without LayerNorm), self_interaction=False (blocks end after the attention / convolution
    half), learnable_statistics=True (the output maps' statistics are parameters: read from device memory by the
    head-output kernels, gradients from the one-CTA reduction kernel), ungated torsions (mean shift).  Forward, loss and
    parameter gradients against the CPU oracle, which tests/test_oracle.py pins against the live reference for exactly
    these switches."""
    import grappa_oracle as orc
    from grappa_b200 import ops, synthetic
    from grappa_b200.energy import Energy
    from grappa_b200.loss import MolwiseLoss
    ops.set_matmul_precision("fp32")
    cfg = orc.small_model_config()
    cfg.update(switches)
    model = _model(cfg, seed=13).eval()
    g = synthetic.espaloma_mix_batch(seed=9, batch_size=4, n_confs=5)
    stat_names = ("mean_over_std", "std", "std_over_max", "k_mean", "k_std")
    param_names = {k for k, _ in model.named_parameters()}

    def wanted(k):
        return k in param_names and (k.rsplit(".", 1)[-1] in stat_names or ".0." in k or "pre_dense" in k)
    sd = {k: v.detach().cpu().double().requires_grad_(wanted(k)) for k, v in model.state_dict().items()}
    h, params, en = orc.path_forward(sd, g, cfg, dtype=torch.float64, create_graph=True)
    ref_loss = orc.molwise_loss(en, params, g)
    leaves = {k: v for k, v in sd.items() if v.requires_grad}
    ref_grads = dict(zip(leaves, torch.autograd.grad(ref_loss, list(leaves.values()), allow_unused=True)))
    gd = torch.nn.Sequential(model, Energy(write_tuple_terms=False))(g.to("cuda"))
    assert rel_err(gd.nodes["n1"].data["h"].detach().cpu().numpy(), h.detach().numpy()) < 1e-5
    assert rel_err(gd.nodes["g"].data["energy"].detach().cpu().numpy(), en["energy"].detach().numpy()) < 1e-5
    for l in LEVELS:
        assert rel_err(gd.nodes[l].data["k"].detach().cpu().numpy(), params[l]["k"].detach().numpy()) < 1e-5, l
        if l in ("n2", "n3"):
            assert rel_err(gd.nodes[l].data["eq"].detach().cpu().numpy(), params[l]["eq"].detach().numpy()) < 1e-5, l
    loss = MolwiseLoss(gradient_weight=0.8, energy_weight=1.0, param_weight=0.0, proper_regularisation=1e-3,
                       improper_regularisation=1e-3)(gd)
    assert abs(loss.item() - float(ref_loss)) < 1e-5 * abs(float(ref_loss))
    model.zero_grad()
    loss.backward()
    named = dict(model.named_parameters())
    n_stats = 0
    for k, gr in ref_grads.items():
        got = named[k].grad
        if gr is None:          # gated torsions: k_mean does not enter; the kernel writes zeros
            assert k.endswith("k_mean") and (got is None or float(got.abs().max()) == 0.0), k
            continue
        assert rel_err(got.cpu().numpy(), gr.numpy()) < 1e-4, k
        n_stats += k.rsplit(".", 1)[-1] in stat_names
    assert n_stats >= 9 and len(ref_grads) >= 30
    # a second evaluation after changing a statistic in place sees the new value (nothing host-cached)
    with torch.no_grad():
        model.parameter_writer.bond_writer.to_eq.std.mul_(2.0)
        g2 = model(g.to("cuda"))
        sd2 = {k: v.detach().cpu().double() for k, v in model.state_dict().items()}
        _, params2, _ = orc.path_forward(sd2, g, cfg, dtype=torch.float64, gradients=False)
    assert rel_err(g2.nodes["n2"].data["eq"].cpu().numpy(), params2["n2"]["eq"].numpy()) < 1e-5
