"""Generate tests/golden/*.npz from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py

Imports /root/reference's grappa.models through oracle/ref_import.py (dgl shim), gives the
reference model weights from grappa_b200.synthetic.deterministic_state_dict (a pure function of key,
shape and seed, so no checkpoint is shipped), runs Sequential(GrappaModel, Energy) in eval mode with
the dihedral noise patched out, and stores inputs + outputs.  The GPU box has no /root/reference;
tests there only read the .npz files.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

from grappa_b200 import graph as gbgraph, synthetic  # noqa: E402
import grappa_oracle as orc  # noqa: E402
from ref_import import import_reference, no_dihedral_noise, to_reference_graph  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
LEVELS = ("n2", "n3", "n4", "n4_improper")


def graph_inputs(g, prefix="in."):
    d = {}
    src, dst = g.edges()
    d[prefix + "src"] = src.numpy()
    d[prefix + "dst"] = dst.numpy()
    for nt in g.ntypes:
        d[prefix + f"count.{nt}"] = g.batch_num_nodes(nt).numpy()
        for k, v in g.nodes[nt].data.items():
            d[prefix + f"{nt}.{k}"] = v.numpy()
    return d


def run_reference(ns, cfg, g, seed, with_loss):
    torch.manual_seed(1234)
    model = ns.deploy.model_from_config(dict(cfg), param_statistics=ns.graph_utils.get_default_statistics())
    sd = synthetic.deterministic_state_dict(model.state_dict(), seed=seed)
    model.load_state_dict(sd)
    model.eval()
    full = torch.nn.Sequential(model, ns.energy.Energy())
    dg = to_reference_graph(ns, g)
    with no_dihedral_noise():
        dg = full(dg)
    out = {"out.h": dg.nodes["n1"].data["h"].detach().numpy()}
    for lvl in LEVELS:
        out[f"out.{lvl}.k"] = dg.nodes[lvl].data["k"].detach().numpy()
        if lvl in ("n2", "n3"):
            out[f"out.{lvl}.eq"] = dg.nodes[lvl].data["eq"].detach().numpy()
        out[f"out.{lvl}.x"] = dg.nodes[lvl].data["x"].detach().numpy()
        out[f"out.{lvl}.energy"] = dg.nodes[lvl].data["energy"].detach().numpy()
        out[f"out.g.energy_{lvl}"] = dg.nodes["g"].data[f"energy_{lvl}"].detach().numpy()
    out["out.g.energy"] = dg.nodes["g"].data["energy"].detach().numpy()
    out["out.n1.gradient"] = dg.nodes["n1"].data["gradient"].detach().numpy()
    keys = sorted(model.state_dict().keys())
    out["meta.state_dict_keys"] = np.array(keys)
    out["meta.state_dict_shapes"] = np.array([",".join(map(str, model.state_dict()[k].shape)) for k in keys])
    if with_loss:
        loss_fn = ns.loss.MolwiseLoss(gradient_weight=0.8, energy_weight=1.0, param_weight=1e-3,
                                      proper_regularisation=1e-3, improper_regularisation=1e-3)
        loss = loss_fn(dg)
        model.zero_grad()
        loss.backward()
        out["out.loss"] = np.array(loss.item(), dtype=np.float64)
        named = dict(model.named_parameters())
        picks = [k for k in named if any(s in k for s in (
            "gnn.pre_dense.0.weight", "gnn.att_blocks.0.graph_module.fc.weight", "gnn.att_blocks.1.layer_norm.weight",
            "gnn.att_blocks.1.self_interaction.2.bias", "gnn.post_dense.0.weight",
            "bond_writer.rep_projector.mlp.0.weight", "angle_writer.angle_model.grappa_transformer.transformer.0.attn.in_proj_weight",
            "proper_writer.torsion_model.symmetriser.mlp.0.linear1.weight", "improper_writer.torsion_model.symmetriser.mlp.2.linear2.weight",
            "proper_writer.torsion_model.grappa_transformer.transformer.1.ff.norm1.bias"))]
        for k in picks:
            out[f"grad.{k}"] = named[k].grad.detach().numpy()
        out["meta.grad_norms_keys"] = np.array(sorted(named.keys()))
        out["meta.grad_norms"] = np.array([float(named[k].grad.norm()) if named[k].grad is not None else 0.0
                                           for k in sorted(named.keys())])
    return out, sd


def param_loss_case(ns):
    """case 5: the classical-parameter term of MolwiseLoss (training/loss.py:70-113) from the reference itself:
    predicted parameters as leaves, references with NaN holes, a torsion reference with MORE periodicities than the
    model (correct_torsion_shape truncates) -- and a second variant with fewer (zero padding) --, per-dataset weights."""
    rng = np.random.default_rng(17)
    mols = [synthetic.make_molecule(rng, "peptide", n_confs=4, n_res=1), synthetic.make_molecule(rng, "small", n_confs=4, n_atoms=17),
            synthetic.make_molecule(rng, "peptide", n_confs=4, n_res=2), synthetic.make_molecule(rng, "rna", n_confs=4)]
    mols = [m for m in mols if m.num_nodes("n4_improper") > 0]
    g = gbgraph.batch(mols)
    gen = torch.Generator().manual_seed(33)
    out = {}
    for variant, ref_per in (("a", 6), ("b", 2)):
        dg = to_reference_graph(ns, g)
        leaves = {}
        for lvl in LEVELS:
            T = g.num_nodes(lvl)
            names = ("k", "eq") if lvl in ("n2", "n3") else ("k",)
            for name in names:
                shape = (T,) if lvl in ("n2", "n3") else (T, 3)
                v = torch.randn(shape, generator=gen) * (50.0 if name == "k" and lvl in ("n2", "n3") else 1.0)
                v.requires_grad_(True)
                dg.nodes[lvl].data[name] = v
                leaves[f"{lvl}.{name}"] = v
                rshape = shape if lvl in ("n2", "n3") else (T, ref_per if lvl == "n4" else 3)
                r = torch.randn(rshape, generator=gen) * (50.0 if name == "k" and lvl in ("n2", "n3") else 1.0)
                r[torch.rand(rshape, generator=gen) < 0.15] = float("nan")
                dg.nodes[lvl].data[name + "_ref"] = r
                out[f"{variant}.ref.{lvl}.{name}"] = r.numpy()
                out[f"{variant}.in.{lvl}.{name}"] = v.detach().numpy()
        # energies / gradients are not part of this case, but unbatch() needs consistent graph-level data
        dsnames = ["spice", "other", "spice", "rna"][:len(mols)]
        loss_fn = ns.loss.MolwiseLoss(gradient_weight=0., energy_weight=0., param_weight=1e-3,
                                      param_weights_by_dataset={"spice": 0.5, "rna": 2.0})
        loss = loss_fn(dg, dsnames=dsnames)
        keys = [k for k in sorted(leaves) if "improper" not in k]
        grads = torch.autograd.grad(loss, [leaves[k] for k in keys])
        out[f"{variant}.loss"] = np.array(loss.item(), dtype=np.float64)
        for k, gr in zip(keys, grads):
            out[f"{variant}.grad.{k}"] = gr.numpy()
        loss2 = ns.loss.MolwiseLoss(gradient_weight=0., energy_weight=0., param_weight=1e-3)(dg)
        out[f"{variant}.loss_uniform"] = np.array(loss2.item(), dtype=np.float64)
    out["meta.dsnames"] = np.array(dsnames)
    np.savez_compressed(os.path.join(OUT, "param_loss.npz"), **graph_inputs(g), **out)
    print("param_loss: loss a/b =", out["a.loss"], out["b.loss"], "mols =", len(mols))


def ragged_conformations_case(ns):
    """case 6: molecules with different numbers of conformations through the reference's own collate steps
    (dgl_utils.set_number_confs -> padding + 'is_dummy', dgl_utils.batch) and MolwiseLoss, whose unbatch() drops the padding
    again (utils/dgl_utils.py:63-118,132-171; data/GraphDataLoader.py:23-73).  Stores the single molecules, the batched
    fields, loss and gradients w.r.t. energy / gradient / torsion amplitudes."""
    rng = np.random.default_rng(23)
    confs = [3, 7, 5, 7, 1]
    kinds = [("peptide", dict(n_res=1)), ("small", dict(n_atoms=12)), ("peptide", dict(n_res=2)), ("rna", {}), ("small", dict(n_atoms=25))]
    mols = [synthetic.make_molecule(rng, k, n_confs=c, **kw) for (k, kw), c in zip(kinds, confs)]
    keep = [i for i, m in enumerate(mols) if m.num_nodes("n4_improper") > 0]
    mols, confs = [mols[i] for i in keep], [confs[i] for i in keep]
    n_confs = max(confs)                     # conf_strategy 'max' / int >= max: only padding, no random sub-sampling
    graphs = [ns.dgl_utils.set_number_confs(to_reference_graph(ns, m), n_confs) for m in mols]
    bg = ns.dgl_utils.batch(graphs, deep_copies_of_same_n_atoms=False)
    gen = torch.Generator().manual_seed(5)
    B, N = bg.num_nodes("g"), bg.num_nodes("n1")
    e = (torch.randn(B, n_confs, generator=gen) * 3).requires_grad_(True)
    gr = (torch.randn(N, n_confs, 3, generator=gen) * 5).requires_grad_(True)
    kp = torch.randn(bg.num_nodes("n4"), 3, generator=gen).requires_grad_(True)
    ki = torch.randn(bg.num_nodes("n4_improper"), 3, generator=gen).requires_grad_(True)
    bg.nodes["g"].data["energy"] = e
    bg.nodes["n1"].data["gradient"] = gr
    bg.nodes["n4"].data["k"] = kp
    bg.nodes["n4_improper"].data["k"] = ki
    for lvl in ("n2", "n3"):
        T = bg.num_nodes(lvl)
        bg.nodes[lvl].data["k"] = torch.rand(T, generator=gen)
        bg.nodes[lvl].data["eq"] = torch.rand(T, generator=gen)
    loss = ns.loss.MolwiseLoss(gradient_weight=0.8, energy_weight=1.0, param_weight=0., proper_regularisation=1e-3,
                               improper_regularisation=1e-3)(bg)
    ge, gg, gkp, gki = torch.autograd.grad(loss, [e, gr, kp, ki])
    out = {"meta.n_confs": np.array(confs), "meta.n_mols": np.array(len(mols)),
           "batched.xyz": bg.nodes["n1"].data["xyz"].numpy(), "batched.is_dummy": bg.nodes["g"].data["is_dummy"].numpy(),
           "batched.energy_ref": bg.nodes["g"].data["energy_ref"].numpy(),
           "batched.gradient_ref": bg.nodes["n1"].data["gradient_ref"].numpy(),
           "batched.n4.idxs": bg.nodes["n4"].data["idxs"].numpy(),
           "in.energy": e.detach().numpy(), "in.gradient": gr.detach().numpy(), "in.k_proper": kp.detach().numpy(),
           "in.k_improper": ki.detach().numpy(), "out.loss": np.array(loss.item(), dtype=np.float64),
           "grad.energy": ge.numpy(), "grad.gradient": gg.numpy(), "grad.k_proper": gkp.numpy(), "grad.k_improper": gki.numpy()}
    for i, m in enumerate(mols):
        out.update(graph_inputs(m, prefix=f"mol{i}."))
    np.savez_compressed(os.path.join(OUT, "ragged_confs.npz"), **out)
    print("ragged_confs: loss =", out["out.loss"], "confs =", confs)


SWITCH_CASES = {
    "noln_learnable_ungated": dict(layer_norm=False, learnable_statistics=True, gated_torsion=False),
    "noselfint_learnable_conv": dict(self_interaction=False, learnable_statistics=True, gated_torsion=True, gnn_convolutions=1),
}


def switches_case(ns):
    """GrappaModel constructor switches that grappa-1.x leaves at their defaults: outputs, loss and the gradients of the
    (learnable) statistics + a few weights from the unmodified reference -> switches_small_model.npz."""
    rng = np.random.default_rng(17)
    mols = [synthetic.make_molecule(rng, "peptide", n_confs=5, n_res=1), synthetic.make_molecule(rng, "small", n_confs=5, n_atoms=24),
            synthetic.make_molecule(rng, "rna", n_confs=5, n_atoms=91)]
    mols = [m for m in mols if m.num_nodes("n4_improper") > 0]
    g = gbgraph.batch(mols)
    out_all = dict(graph_inputs(g))
    stat_names = ("mean_over_std", "std", "std_over_max", "k_mean", "k_std")
    for name, switches in SWITCH_CASES.items():
        cfg = orc.small_model_config()
        cfg.update(switches)
        torch.manual_seed(99)
        model = ns.deploy.model_from_config(dict(cfg), param_statistics=ns.graph_utils.get_default_statistics())
        sd = synthetic.deterministic_state_dict(model.state_dict(), seed=31)
        model.load_state_dict(sd)
        model.eval()
        dg = to_reference_graph(ns, g)
        with no_dihedral_noise():
            dg = torch.nn.Sequential(model, ns.energy.Energy())(dg)
        loss = ns.loss.MolwiseLoss(gradient_weight=0.8, energy_weight=1.0, param_weight=0.0, proper_regularisation=1e-3,
                                   improper_regularisation=1e-3)(dg)
        model.zero_grad()
        loss.backward()
        pre = f"{name}."
        out_all[pre + "out.h"] = dg.nodes["n1"].data["h"].detach().numpy()
        for lvl in LEVELS:
            out_all[pre + f"out.{lvl}.k"] = dg.nodes[lvl].data["k"].detach().numpy()
            if lvl in ("n2", "n3"):
                out_all[pre + f"out.{lvl}.eq"] = dg.nodes[lvl].data["eq"].detach().numpy()
        out_all[pre + "out.g.energy"] = dg.nodes["g"].data["energy"].detach().numpy()
        out_all[pre + "out.n1.gradient"] = dg.nodes["n1"].data["gradient"].detach().numpy()
        out_all[pre + "out.loss"] = np.array(loss.item(), dtype=np.float64)
        named = dict(model.named_parameters())
        out_all[pre + "meta.parameter_names"] = np.array(sorted(named))
        out_all[pre + "meta.state_dict_keys"] = np.array(sorted(model.state_dict().keys()))
        for k, p in named.items():
            if k.rsplit(".", 1)[-1] in stat_names or k in ("gnn.pre_dense.0.weight", "gnn.att_blocks.0.head_reducer.weight",
                                                           "parameter_writer.angle_writer.angle_model.symmetriser.mlp.0.linear1.weight"):
                out_all[pre + f"grad.{k}"] = (np.zeros(p.shape, np.float32) if p.grad is None else p.grad.detach().numpy())
                out_all[pre + f"gradnone.{k}"] = np.array(p.grad is None)
        # the restatement must agree before the fixture is written
        h, params, en = orc.path_forward(sd, g, cfg)
        err = np.abs(h.detach().numpy() - out_all[pre + "out.h"]).max() / np.abs(out_all[pre + "out.h"]).max()
        assert err < 2e-5, (name, err)
        print(name, "loss =", loss.item(), "oracle-vs-reference h error", err)
    np.savez_compressed(os.path.join(OUT, "switches_small_model.npz"), **out_all)


def train_batch_case(ns):
    """BASELINE configs[1] at full size -- 32 x ACE-(ALA)4-NME, 50 conformations, grappa-1.2 architecture -- through the
    unmodified reference: energies, parameters, loss and per-parameter gradient norms -> train_batch_grappa12.npz.  The
    inputs are NOT stored (1 MB of coordinates): synthetic.peptide_batch(seed=100, ...) regenerates them; a checksum
    guards the regeneration."""
    g = synthetic.peptide_batch(seed=100, batch_size=32, n_res=4, n_confs=50)
    cfg = orc.grappa_1_2_model_config()
    torch.manual_seed(5)
    model = ns.deploy.model_from_config(dict(cfg), param_statistics=ns.graph_utils.get_default_statistics())
    sd = synthetic.deterministic_state_dict(model.state_dict(), seed=11)
    model.load_state_dict(sd)
    model.eval()
    dg = to_reference_graph(ns, g)
    with no_dihedral_noise():
        dg = torch.nn.Sequential(model, ns.energy.Energy())(dg)
    loss = ns.loss.MolwiseLoss(gradient_weight=0.8, energy_weight=1.0, param_weight=0.0, proper_regularisation=1e-3,
                               improper_regularisation=1e-3)(dg)
    model.zero_grad()
    loss.backward()
    named = dict(model.named_parameters())
    keys = sorted(named)
    out = {"meta.seed": np.array(100), "meta.weights_seed": np.array(11),
           "meta.xyz_checksum": np.array(float(g.nodes["n1"].data["xyz"].double().sum())),
           "meta.xyz_abs_checksum": np.array(float(g.nodes["n1"].data["xyz"].double().abs().sum())),
           "out.g.energy": dg.nodes["g"].data["energy"].detach().numpy(),
           "out.gradient_norm_per_atom": dg.nodes["n1"].data["gradient"].detach().norm(dim=(1, 2)).numpy(),
           "out.loss": np.array(loss.item(), dtype=np.float64),
           "meta.grad_norms_keys": np.array(keys),
           "meta.grad_norms": np.array([float(named[k].grad.norm()) if named[k].grad is not None else 0.0 for k in keys])}
    for lvl in LEVELS:
        out[f"out.{lvl}.k"] = dg.nodes[lvl].data["k"].detach().numpy()
        if lvl in ("n2", "n3"):
            out[f"out.{lvl}.eq"] = dg.nodes[lvl].data["eq"].detach().numpy()
    for k in ("gnn.att_blocks.6.head_reducer.bias", "parameter_writer.proper_writer.torsion_model.symmetriser.mlp.2.linear2.weight"):
        out[f"grad.{k}"] = named[k].grad.detach().numpy()
    # every parameter tensor: max |grad| and 16 sampled gradient entries (seeded positions), so that the GPU test can bound
    # the max-norm relative gradient error of ALL 305 tensors at full width without shipping 163 MB of gradients
    rs = np.random.default_rng(2024)
    sidx = np.zeros((len(keys), 16), dtype=np.int64)
    sval = np.zeros((len(keys), 16), dtype=np.float32)
    gmax = np.zeros(len(keys), dtype=np.float64)
    for i, k in enumerate(keys):
        gk = named[k].grad
        n = named[k].numel()
        sidx[i] = rs.integers(0, n, size=16)
        if gk is not None:
            flat = gk.detach().reshape(-1)
            sval[i] = flat[torch.from_numpy(sidx[i])].numpy()
            gmax[i] = float(flat.abs().max())
    out["meta.grad_sample_idx"], out["meta.grad_sample_val"], out["meta.grad_maxabs"] = sidx, sval, gmax
    np.savez_compressed(os.path.join(OUT, "train_batch_grappa12.npz"), **out)
    print("train_batch_grappa12: loss =", loss.item())


def protein_and_mix_cases(ns):
    """BASELINE configs[2] (ACE-(ALA)149-NME, 1,502 atoms, grappa-1.2 parametrisation) and configs[4] (Espaloma-shaped
    mix of 32 molecules, 32 conformations, narrow model, energies + forces) through the unmodified reference, with the
    seeds tests/test_baseline_configs_gpu.py uses -> protein_grappa12.npz, espaloma_mix_small_model.npz.  Inputs are
    regenerated from the seeds (checksummed).  No loss for the mix: the reference's MolwiseLoss is NaN for molecules
    without impropers (training/loss.py:130-132)."""
    g = synthetic.protein(seed=0)
    cfg = orc.grappa_1_2_model_config()
    torch.manual_seed(5)
    model = ns.deploy.model_from_config(dict(cfg), param_statistics=ns.graph_utils.get_default_statistics())
    model.load_state_dict(synthetic.deterministic_state_dict(model.state_dict(), seed=12))
    model.eval()
    with torch.no_grad():
        dg = model(to_reference_graph(ns, g))
    q = g.nodes["n1"].data["partial_charge"].double()
    out = {"meta.seed": np.array(0), "meta.weights_seed": np.array(12), "meta.charge_checksum": np.array(float(q.abs().sum())),
           "out.h_norm_per_atom": dg.nodes["n1"].data["h"].norm(dim=1).numpy(),
           "out.h_first_atoms": dg.nodes["n1"].data["h"][:8].numpy()}
    for lvl in LEVELS:
        out[f"out.{lvl}.k"] = dg.nodes[lvl].data["k"].numpy()
        if lvl in ("n2", "n3"):
            out[f"out.{lvl}.eq"] = dg.nodes[lvl].data["eq"].numpy()
    np.savez_compressed(os.path.join(OUT, "protein_grappa12.npz"), **out)
    print("protein_grappa12: atoms", g.num_nodes("n1"), "k[0] =", out["out.n2.k"][0])

    g = synthetic.espaloma_mix_batch(seed=4, batch_size=32, n_confs=32)
    cfg = orc.small_model_config()
    torch.manual_seed(5)
    model = ns.deploy.model_from_config(dict(cfg), param_statistics=ns.graph_utils.get_default_statistics())
    model.load_state_dict(synthetic.deterministic_state_dict(model.state_dict(), seed=13))
    model.eval()
    dg = to_reference_graph(ns, g)
    with no_dihedral_noise():
        dg = torch.nn.Sequential(model, ns.energy.Energy())(dg)
    xyz = g.nodes["n1"].data["xyz"].double()
    out = {"meta.seed": np.array(4), "meta.weights_seed": np.array(13), "meta.xyz_abs_checksum": np.array(float(xyz.abs().sum())),
           "meta.atom_counts": g.batch_num_nodes("n1").numpy(),
           "out.g.energy": dg.nodes["g"].data["energy"].detach().numpy(),
           "out.n1.gradient": dg.nodes["n1"].data["gradient"].detach().numpy()}
    for lvl in LEVELS:
        out[f"out.{lvl}.k"] = dg.nodes[lvl].data["k"].detach().numpy()
        if lvl in ("n2", "n3"):
            out[f"out.{lvl}.eq"] = dg.nodes[lvl].data["eq"].detach().numpy()
    np.savez_compressed(os.path.join(OUT, "espaloma_mix_small_model.npz"), **out)
    print("espaloma_mix_small_model: atoms", g.num_nodes("n1"), "E[0,0] =", out["out.g.energy"][0, 0])


def tiny_model_config():
    """A very small architecture whose reference-saved state_dict (~0.4 MB) can be committed."""
    cfg = orc.grappa_1_2_model_config()
    cfg.update(graph_node_features=32, gnn_width=64, gnn_attentional_layers=2, gnn_attention_heads=4)
    for w in ("bond", "angle", "proper", "improper"):
        cfg.update({f"{w}_transformer_depth": 1, f"{w}_n_heads": 4, f"{w}_transformer_width": 64,
                    f"{w}_symmetriser_depth": 2, f"{w}_symmetriser_width": 32})
    return cfg


def import_reference_parameters():
    """The reference's output dataclass `grappa.data.Parameters` (data/Parameters.py:19-140).  It imports matplotlib at
    module top (absent here; only its plotting helpers use it) -> stubbed; everything else imports unmodified."""
    import importlib
    import types
    if "matplotlib" not in sys.modules:
        class _Stub(types.ModuleType):
            __path__ = []

            def __getattr__(self, name):
                return _Stub(name)
        for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.colors", "matplotlib.cm"):
            sys.modules[name] = _Stub(name)
    pkg = sys.modules["grappa"]
    pkg.units = importlib.import_module("grappa.units")
    return importlib.import_module("grappa.data.Parameters").Parameters


def dropin_case(ns):
    """Drop-in fixtures (task 8): (i) a state_dict SAVED BY THE REFERENCE model with the reference's own
    initialisation (torch.save -> reference_tiny_state_dict.pt), (ii) the reference's outputs, loss and
    Parameters.from_dgl (grappa.py:36-57 flow: eval, no_grad, to cpu) for it on a small batch -> dropin_tiny.npz."""
    cfg = tiny_model_config()
    torch.manual_seed(21)
    model = ns.deploy.model_from_config(dict(cfg), param_statistics=ns.graph_utils.get_default_statistics())
    with torch.no_grad():       # LayerNorm / bias parameters away from their 1 / 0 defaults, like a trained checkpoint
        gen = torch.Generator().manual_seed(22)
        for k, p in model.named_parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn(p.shape, generator=gen))
    model.eval()
    torch.save({k: v.clone() for k, v in model.state_dict().items()}, os.path.join(OUT, "reference_tiny_state_dict.pt"))
    rng = np.random.default_rng(23)
    mols = [synthetic.make_molecule(rng, "peptide", n_confs=5, n_res=1), synthetic.make_molecule(rng, "small", n_confs=5, n_atoms=14),
            synthetic.make_molecule(rng, "peptide", n_confs=5, n_res=2)]
    mols = [m for m in mols if m.num_nodes("n4_improper") > 0]
    for i, m in enumerate(mols):     # atom ids as a topology would carry them: not consecutive, not starting at zero
        m.nodes["n1"].data["ids"] = torch.arange(m.num_nodes("n1")) * 3 + 7 + 100 * i
    g = gbgraph.batch(mols)
    dg = to_reference_graph(ns, g)
    with no_dihedral_noise():
        dg = torch.nn.Sequential(model, ns.energy.Energy())(dg)
    loss = ns.loss.MolwiseLoss(gradient_weight=0.8, energy_weight=1.0, param_weight=0.0, proper_regularisation=1e-3,
                               improper_regularisation=1e-3)(dg)
    out = {"out.h": dg.nodes["n1"].data["h"].detach().numpy(), "out.g.energy": dg.nodes["g"].data["energy"].detach().numpy(),
           "out.n1.gradient": dg.nodes["n1"].data["gradient"].detach().numpy(), "out.loss": np.array(loss.item(), dtype=np.float64),
           "meta.config_keys": np.array(sorted(cfg.keys())),
           "meta.config_vals": np.array([repr(cfg[k]) for k in sorted(cfg.keys())])}
    for lvl in LEVELS:
        out[f"out.{lvl}.k"] = dg.nodes[lvl].data["k"].detach().numpy()
        if lvl in ("n2", "n3"):
            out[f"out.{lvl}.eq"] = dg.nodes[lvl].data["eq"].detach().numpy()
    # Grappa.predict flow on the first molecule alone (grappa.py:36-57): model in eval mode, no_grad, to cpu, from_dgl
    Parameters = import_reference_parameters()
    d1 = to_reference_graph(ns, mols[0])
    with torch.no_grad():
        d1 = model(d1)
    d1 = d1.to("cpu")
    prm = Parameters.from_dgl(d1)
    for f in ("atoms", "bonds", "bond_k", "bond_eq", "angles", "angle_k", "angle_eq", "propers", "proper_ks", "proper_phases",
              "impropers", "improper_ks", "improper_phases"):
        out[f"params.{f}"] = np.asarray(getattr(prm, f))
    np.savez_compressed(os.path.join(OUT, "dropin_tiny.npz"), **graph_inputs(g), **out)
    print("dropin_tiny: loss =", loss.item(), "state_dict entries", len(model.state_dict()),
          "bytes", os.path.getsize(os.path.join(OUT, "reference_tiny_state_dict.pt")))


def main():
    ns = import_reference()
    torch.set_num_threads(8)
    if "--only-dropin" in sys.argv:
        return dropin_case(ns)
    if "--only-protein-mix" in sys.argv:
        return protein_and_mix_cases(ns)
    if "--only-train-batch" in sys.argv:
        return train_batch_case(ns)
    if "--only-switches" in sys.argv:
        return switches_case(ns)
    if "--only-param-loss" in sys.argv:
        return param_loss_case(ns)
    if "--only-ragged" in sys.argv:
        return ragged_conformations_case(ns)
    dropin_case(ns)
    param_loss_case(ns)
    ragged_conformations_case(ns)
    switches_case(ns)
    train_batch_case(ns)
    protein_and_mix_cases(ns)

    # ---- case 1: BASELINE config 1 -- grappa-1.2 architecture, capped dipeptide, 50 conformations
    g = synthetic.dipeptide(seed=11, n_confs=50)
    cfg = orc.grappa_1_2_model_config()
    out, sd = run_reference(ns, cfg, g, seed=3, with_loss=False)
    # the restatement must agree with the reference before the fixture is written
    h, params, en = orc.path_forward(sd, g, cfg)
    assert np.allclose(h.detach().numpy(), out["out.h"], rtol=1e-4, atol=1e-5), "oracle h != reference"
    assert np.allclose(en["energy"].detach().numpy(), out["out.g.energy"], rtol=1e-5, atol=1e-3)
    np.savez_compressed(os.path.join(OUT, "dipeptide_grappa12.npz"), **graph_inputs(g), **out)
    print("dipeptide_grappa12: energy[0,:3] =", out["out.g.energy"][0, :3])

    # ---- case 2: narrow architecture, mixed batch (peptide + small molecules + ring systems), with loss + grads
    rng = np.random.default_rng(5)
    mols = [synthetic.make_molecule(rng, "peptide", n_confs=7, n_res=1),
            synthetic.make_molecule(rng, "small", n_confs=7, n_atoms=9),
            synthetic.make_molecule(rng, "small", n_confs=7, n_atoms=31),
            synthetic.make_molecule(rng, "peptide", n_confs=7, n_res=2)]
    # the reference loss NaNs on molecules without impropers (training/loss.py:130-132): keep only those with >= 1
    mols = [m for m in mols if m.num_nodes("n4_improper") > 0]
    g = gbgraph.batch(mols)
    cfg = orc.small_model_config()
    out, sd = run_reference(ns, cfg, g, seed=7, with_loss=True)
    np.savez_compressed(os.path.join(OUT, "mixed_batch_small_model.npz"), **graph_inputs(g), **out)
    print("mixed_batch_small_model: loss =", out["out.loss"], "mols =", len(mols))

    # ---- case 3: Energy alone with given parameters on a larger mixed batch (incl. rna-like rings),
    #      plus the double backward (dL/dk, dL/deq for random upstream gE, gF) that K14 replaces
    rng = np.random.default_rng(9)
    mols = [synthetic.make_molecule(rng, "rna", n_confs=13), synthetic.make_molecule(rng, "peptide", n_confs=13, n_res=3),
            synthetic.make_molecule(rng, "small", n_confs=13, n_atoms=3), synthetic.make_molecule(rng, "small", n_confs=13, n_atoms=44)]
    g = gbgraph.batch(mols)
    dg = to_reference_graph(ns, g)
    prm = {}
    gen = torch.Generator().manual_seed(21)
    for lvl in LEVELS:
        T = g.num_nodes(lvl)
        if lvl == "n2":
            prm[lvl] = {"k": 500 + 300 * torch.rand(T, generator=gen), "eq": 1.0 + 0.5 * torch.rand(T, generator=gen)}
        elif lvl == "n3":
            prm[lvl] = {"k": 60 + 80 * torch.rand(T, generator=gen), "eq": 1.7 + 0.5 * torch.rand(T, generator=gen)}
        else:
            prm[lvl] = {"k": torch.randn(T, 3 if lvl == "n4" else 2, generator=gen)}
        for name, v in prm[lvl].items():
            v.requires_grad_(True)
            dg.nodes[lvl].data[name] = v
    with no_dihedral_noise():
        dg = ns.energy.Energy()(dg)
    E = dg.nodes["g"].data["energy"]
    Gd = dg.nodes["n1"].data["gradient"]
    gE = torch.randn(E.shape, generator=gen)
    gF = torch.randn(Gd.shape, generator=gen)
    leaves = [prm[l][n] for l in LEVELS for n in sorted(prm[l])]
    grads = torch.autograd.grad((E * gE).sum() + (Gd * gF).sum(), leaves)
    out = {"out.g.energy": E.detach().numpy(), "out.n1.gradient": Gd.detach().numpy(), "in.gE": gE.numpy(), "in.gF": gF.numpy()}
    i = 0
    for l in LEVELS:
        out[f"out.{l}.x"] = dg.nodes[l].data["x"].detach().numpy()
        out[f"out.g.energy_{l}"] = dg.nodes["g"].data[f"energy_{l}"].detach().numpy()
        for n in sorted(prm[l]):
            out[f"in.{l}.{n}"] = prm[l][n].detach().numpy()
            out[f"grad.{l}.{n}"] = grads[i].numpy()
            i += 1
    np.savez_compressed(os.path.join(OUT, "energy_mixed_batch.npz"), **graph_inputs(g), **out)
    print("energy_mixed_batch: atoms", g.num_nodes("n1"), "E[0,:2]", out["out.g.energy"][0, :2])

    # ---- case 4: tuple indices of shuffled bond lists (bit-exact contract)
    rng = np.random.default_rng(13)
    tup = {}
    for name, (el, b, imp) in {
        "dipeptide": synthetic.polyalanine_topology(1), "peptide4": synthetic.polyalanine_topology(4),
        "tree": synthetic.random_tree_topology(rng, 27, 2), "rna": synthetic.rna_like_topology(rng, 93),
        "protein30": synthetic.polyalanine_topology(30)}.items():
        b = b[rng.permutation(len(b))]
        flip = rng.random(len(b)) < 0.5
        b[flip] = b[flip][:, ::-1]
        ref = ns.tuple_indices.get_idx_tuples([tuple(x) for x in b.tolist()])
        nd = ns.tuple_indices.get_neighbor_dict([tuple(x) for x in b.tolist()])
        _, ri = ns.tuple_indices.get_torsions([tuple(x) for x in imp], nd)
        tup[f"{name}.in_bonds"] = b
        tup[f"{name}.in_improper_candidates"] = np.array(imp, dtype=np.int64).reshape(-1, 4)
        tup[f"{name}.bonds"] = np.array(ref["bonds"], dtype=np.int64).reshape(-1, 2)
        tup[f"{name}.angles"] = np.array(ref["angles"], dtype=np.int64).reshape(-1, 3)
        tup[f"{name}.propers"] = np.array(ref["propers"], dtype=np.int64).reshape(-1, 4)
        tup[f"{name}.impropers"] = np.array(ri, dtype=np.int64).reshape(-1, 4)
    np.savez_compressed(os.path.join(OUT, "tuple_indices.npz"), **tup)
    print("tuple_indices:", sorted({k.split('.')[0] for k in tup}))


if __name__ == "__main__":
    main()
