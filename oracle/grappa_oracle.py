"""CPU oracle: a functional torch restatement of Grappa's model hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; it is the CHECKER, never the product.  The product package `grappa_b200` never imports
it and has no CPU path.

Pinning: this restatement is checked against the UNMODIFIED reference (`/root/reference`, imported
through oracle/ref_import.py + oracle/dgl_shim) by tests/test_oracle_vs_reference.py in the build
container, and against the committed fixtures tests/golden/*.npz (generated from the reference by
tests/golden/make_golden.py) everywhere else.  The reference itself holds no golden vectors for this
path, so the fixtures generated from the reference's own code ARE the pin.

Everything is a pure function of (state_dict with the reference's key names, config dict, tensors);
dtype follows the inputs, so the same code gives the fp32 baseline and an fp64 tie-breaker.
Dropout is omitted (eval mode); the random noise the reference adds inside `dihedral`
(internal_coordinates.py:194-196) is omitted (SURVEY.md section 8c hazard 1).
"""
from __future__ import annotations

import math
from typing import Dict, Sequence

import torch
import torch.nn.functional as F

LEVELS = ("n2", "n3", "n4", "n4_improper")
WRITER = {"n2": ("bond_writer", "bond_model", 2), "n3": ("angle_writer", "angle_model", 3),
          "n4": ("proper_writer", "torsion_model", 4), "n4_improper": ("improper_writer", "torsion_model", 4)}
# symmetriser permutations: reference interaction_parameters.py:238,330,494,500
PERMS = {"n2": [[0, 1], [1, 0]], "n3": [[0, 1, 2], [2, 1, 0]],
         "n4": [[0, 1, 2, 3], [3, 2, 1, 0]], "n4_improper": [[0, 1, 2, 3], [3, 1, 2, 0]]}
# positional encodings: reference perm_equiv_transformer.py:172-181
POS_ENC = {"n2": None, "n3": [0.0, 1.0, 0.0], "n4": [0.0, 1.0, 1.0, 0.0], "n4_improper": [0.0, 1.0, 1.0, 0.0]}


def grappa_1_2_model_config() -> dict:
    """experiments/train-grappa-1.2/grappa_config.yaml:64-111 (model_config)."""
    cfg = dict(
        graph_node_features=256, in_feats=None,
        in_feat_name=["atomic_number", "partial_charge", "ring_encoding", "degree", "charge_model"],
        in_feat_dims={}, gnn_width=512, gnn_attentional_layers=7, gnn_convolutions=0, gnn_attention_heads=16,
        gnn_dropout_attention=0.3, gnn_dropout_initial=0.0, gnn_dropout_conv=0.1, gnn_dropout_final=0.1,
        parameter_dropout=0.5, n_periodicity_proper=3, n_periodicity_improper=3, gated_torsion=True,
        wrong_symmetry=False, positional_encoding=True, layer_norm=True, self_interaction=True,
        learnable_statistics=False, torsion_cutoff=1e-4)
    for w in ("bond", "angle", "proper", "improper"):
        cfg.update({f"{w}_transformer_depth": 3, f"{w}_n_heads": 8, f"{w}_transformer_width": 512,
                    f"{w}_symmetriser_depth": 3, f"{w}_symmetriser_width": 256})
    return cfg


def small_model_config() -> dict:
    """A narrow architecture (same structure, fewer/narrower layers) for fast CPU parity cases."""
    cfg = grappa_1_2_model_config()
    cfg.update(graph_node_features=64, gnn_width=128, gnn_attentional_layers=2, gnn_attention_heads=4)
    for w in ("bond", "angle", "proper", "improper"):
        cfg.update({f"{w}_transformer_depth": 2, f"{w}_n_heads": 4, f"{w}_transformer_width": 128,
                    f"{w}_symmetriser_depth": 3, f"{w}_symmetriser_width": 64})
    return cfg


# --------------------------------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------------------------------
def linear(sd, key, x):
    b = sd.get(key + ".bias")
    return F.linear(x, sd[key + ".weight"].to(x.dtype), None if b is None else b.to(x.dtype))


def layer_norm(sd, key, x, on: bool = True):
    # torch.nn.LayerNorm defaults: eps 1e-5, biased variance; `on=False`: the model was built with
    # layer_norm=False and the norm is skipped (graph_attention.py:279,298,398; network_utils.py:46,115)
    if not on:
        return x
    return F.layer_norm(x, (x.shape[-1],), sd[key + ".weight"].to(x.dtype), sd[key + ".bias"].to(x.dtype), 1e-5)


def charge_encoding(q: torch.Tensor, dim: int = 16) -> torch.Tensor:
    """PositionalEncoding(dimension=16, min=-2, max=2): reference models/graph_attention.py:418-444."""
    s = (torch.clamp(q, -2.0, 2.0) + 2.0) / 4.0
    i = torch.arange(0, dim // 2, dtype=torch.float32, device=q.device)
    freq = torch.exp(i * -(math.log(10000.0)) / (dim // 2)).to(q.dtype)
    enc = torch.zeros(len(q), dim, dtype=q.dtype, device=q.device)
    enc[:, 0::2] = torch.sin(s[:, None] * freq)
    enc[:, 1::2] = torch.cos(s[:, None] * freq)
    return enc


def input_features(g, in_feat_name: Sequence[str], dtype=torch.float32) -> torch.Tensor:
    """Feature concat in config order + charge sinusoid: models/graph_attention.py:157-164."""
    cols = []
    for f in in_feat_name:
        t = g.nodes["n1"].data[f].to(dtype)
        cols.append(t if t.dim() >= 2 else t[:, None])
    cols.append(charge_encoding(g.nodes["n1"].data["partial_charge"].to(dtype)))
    return torch.cat(cols, dim=-1)


def dot_gat(ft: torch.Tensor, src: torch.Tensor, dst: torch.Tensor) -> torch.Tensor:
    """DGL DotGatConv body (published semantics; see oracle/dgl_shim): ft is (N, H, D)."""
    n, h, d = ft.shape
    a = (ft[src] * ft[dst]).sum(-1) / math.sqrt(d)
    amax = torch.full((n, h), -float("inf"), dtype=a.dtype).scatter_reduce(
        0, dst[:, None].expand_as(a), a, reduce="amax", include_self=True)
    ex = torch.exp(a - amax[dst])
    den = torch.zeros((n, h), dtype=a.dtype).index_add(0, dst, ex)
    return torch.zeros_like(ft).index_add(0, dst, (ex / den[dst])[:, :, None] * ft[src])


def gnn_forward(sd, g, cfg, prefix: str = "gnn.", dtype=torch.float32, taps: dict | None = None) -> torch.Tensor:
    """GrappaGNN.forward: models/graph_attention.py:142-183 with blocks :276-310."""
    x = input_features(g, cfg["in_feat_name"], dtype)
    if taps is not None:
        taps["features"] = x
    h = F.elu(linear(sd, prefix + "pre_dense.0", x))
    src, dst = g.edges(etype="n1_edge")
    src, dst = src.long(), dst.long()
    heads = cfg["gnn_attention_heads"]
    n_conv = cfg.get("gnn_convolutions", 0)
    ln, self_int = cfg.get("layer_norm", True), cfg.get("self_interaction", True)
    for i in range(n_conv):
        # ResidualConvBlock.forward (graph_attention.py:378-415) around dgl.nn.SAGEConv(in, out, 'mean') as restated in
        # oracle/dgl_shim (fc_self with bias + bias-free fc_neigh on the mean over in-neighbours)
        p = f"{prefix}conv_blocks.{i}."
        u = layer_norm(sd, p + "layer_norm", h, ln)
        agg = torch.zeros_like(u).index_add(0, dst, u[src])
        deg = torch.zeros(len(u), dtype=u.dtype).index_add(0, dst, torch.ones(len(dst), dtype=u.dtype)).clamp(min=1)
        conv = linear(sd, p + "graph_module.fc_self", u) + F.linear(agg / deg[:, None], sd[p + "graph_module.fc_neigh.weight"].to(dtype))
        y = F.elu(conv) + u
        if self_int:
            z = layer_norm(sd, p + "interaction_norm", y, ln)
            h = F.elu(linear(sd, p + "self_interaction.0", z)) + z
        else:
            h = y
        if taps is not None:
            taps[f"conv{i}"] = h
    for i in range(cfg["gnn_attentional_layers"]):
        p = f"{prefix}att_blocks.{i}."          # same parameters as blocks.{n_conv + i} (graph_attention.py:129)
        u = layer_norm(sd, p + "layer_norm", h, ln)
        ft = F.linear(u, sd[p + "graph_module.fc.weight"].to(dtype)).view(len(u), heads, -1)
        m = dot_gat(ft, src, dst).flatten(1)
        y = linear(sd, p + "head_reducer", m) + u          # skip adds the LayerNorm-ed u
        if self_int:
            z = layer_norm(sd, p + "interaction_norm", y, ln)
            w = F.elu(linear(sd, p + "self_interaction.2", F.elu(linear(sd, p + "self_interaction.0", z))))
            h = w + z                                      # skip adds z
        else:
            h = y
        if taps is not None:
            taps[f"block{i}"] = h
    return linear(sd, prefix + "post_dense.0", h)


def _mha(sd, p, x, n_heads):
    """nn.MultiheadAttention(E, n_heads, batch_first=False) self-attention over dim 0 (appendix A.1)."""
    L, T, E = x.shape
    qkv = F.linear(x, sd[p + "in_proj_weight"].to(x.dtype), sd[p + "in_proj_bias"].to(x.dtype))
    q, k, v = qkv.split(E, dim=-1)
    hd = E // n_heads
    q = q.reshape(L, T, n_heads, hd).permute(1, 2, 0, 3)
    k = k.reshape(L, T, n_heads, hd).permute(1, 2, 0, 3)
    v = v.reshape(L, T, n_heads, hd).permute(1, 2, 0, 3)
    att = torch.softmax((q / math.sqrt(hd)) @ k.transpose(-1, -2), dim=-1)
    o = (att @ v).permute(2, 0, 1, 3).reshape(L, T, E)
    return linear(sd, p + "out_proj", o)


def writer_scores(sd, h, idxs, level, cfg, prefix="parameter_writer.", taps: dict | None = None):
    """RepProjector + GrappaTransformer + Symmetriser -> raw coefficients (T, out).

    interaction_parameters.py:155-180; perm_equiv_transformer.py:127-151,239-276; network_utils.py:44-54,112-133.
    """
    wname, mname, L = WRITER[level]
    short = {"n2": "bond", "n3": "angle", "n4": "proper", "n4_improper": "improper"}[level]
    p = f"{prefix}{wname}."
    proj = F.elu(linear(sd, p + "rep_projector.mlp.0", h))
    x = proj[idxs.long()].transpose(0, 1)                                  # (L, T, F)
    T = x.shape[1]
    perms, pos_enc = PERMS[level], POS_ENC[level]
    if level == "n4_improper" and cfg.get("wrong_symmetry", False):
        # ablation switch (interaction_parameters.py:499-505): all permutations that keep the central atom fixed
        perms = [[0, 1, 2, 3], [3, 1, 2, 0], [1, 3, 2, 0], [0, 3, 2, 1], [3, 0, 2, 1], [1, 0, 2, 3]]
        pos_enc = [0.0, 0.0, 1.0, 0.0]
    if pos_enc is not None and cfg.get("positional_encoding", True):
        pe = torch.tensor(pos_enc, dtype=x.dtype)[:, None, None].expand(L, T, 1)
        x = torch.cat([x, pe], dim=-1)
    n_heads = cfg[f"{short}_n_heads"]
    ln = cfg.get("layer_norm", True)
    for i in range(cfg[f"{short}_transformer_depth"]):
        q = f"{p}{mname}.grappa_transformer.transformer.{i}."
        x = layer_norm(sd, q + "norm1", x, ln)
        x = _mha(sd, q + "attn.", x, n_heads) + x                          # residual is the post-LN x
        xn = layer_norm(sd, q + "ff.norm1", x, ln)
        x = linear(sd, q + "ff.linear2", F.elu(linear(sd, q + "ff.linear1", xn))) + xn
        if taps is not None:
            taps[f"{level}_layer{i}"] = x
    out = 0.0
    depth = cfg[f"{short}_symmetriser_depth"]
    for perm in perms:
        s = torch.cat([x[j] for j in perm], dim=-1)                        # (T, L*E)
        for i in range(depth):
            q = f"{p}{mname}.symmetriser.mlp.{i}."
            sn = layer_norm(sd, q + "norm1", s, ln)
            y = linear(sd, q + "linear2", F.elu(linear(sd, q + "linear1", sn)))
            s = y + sn if (0 < i < depth - 1) else y                       # skip only on middle layers
        out = out + s
    return out


def to_positive(sd, p, x):
    """ToPositive: final_layer.py:48-52."""
    return sd[p + ".std"].to(x.dtype) * (F.elu(sd[p + ".mean_over_std"].to(x.dtype) + x - 1.0) + 1.0) + sd[p + ".min_"].to(x.dtype)


def to_range(sd, p, x):
    """ToRange: final_layer.py:91-97."""
    return sd[p + ".max"].to(x.dtype) * torch.sigmoid(sd[p + ".std_over_max"].to(x.dtype) * x)


def writer_params(sd, scores, level, cfg, prefix="parameter_writer."):
    """Output maps: interaction_parameters.py:244-266 (bond), :337-362 (angle), :519-562 (torsion)."""
    wname = WRITER[level][0]
    p = f"{prefix}{wname}."
    if level in ("n2", "n3"):
        eq = to_positive(sd, p + "to_eq", scores[:, 0]) if level == "n2" else to_range(sd, p + "to_eq", scores[:, 0])
        k = to_positive(sd, p + "to_k", scores[:, 1])
        return {"eq": eq, "k": k}
    n = int(sd[p + "n_periodicity"])
    if scores.shape[0] == 0:
        return {"k": torch.zeros((0, n), dtype=scores.dtype)}
    if cfg.get("gated_torsion", False):
        k = scores[:, :n] * torch.sigmoid(scores[:, n:]) * sd[p + "k_std"].to(scores.dtype)
    else:
        k = scores * sd[p + "k_std"].to(scores.dtype) + sd[p + "k_mean"].to(scores.dtype)
    cutoff = cfg.get("torsion_cutoff", 1e-4)
    if cutoff > 0:
        k = torch.where(torch.abs(k) > cutoff, k, torch.zeros_like(k))
    return {"k": k}


# --------------------------------------------------------------------------------------------------
# geometry and energy
# --------------------------------------------------------------------------------------------------
def bond_length(x0, x1):
    return torch.norm(x0 - x1, p=2, dim=-1)


def bond_angle(x0, x1, x2):
    left, right = x1 - x0, x1 - x2
    return torch.atan2(torch.norm(torch.cross(left, right, dim=-1), p=2, dim=-1), (left * right).sum(-1))


def dihedral_angle(x0, x1, x2, x3):
    """internal_coordinates.py:178-210 without the randn*1e-5 perturbation."""
    r01, r21, r23 = x1 - x0, x1 - x2, x3 - x2
    n1 = torch.cross(r01, r21, dim=-1)
    n2 = torch.cross(r21, r23, dim=-1)
    rkj = r21 / torch.norm(r21, dim=-1, keepdim=True)
    y = (torch.cross(n1, n2, dim=-1) * rkj).sum(-1)
    x = (n1 * n2).sum(-1)
    return torch.atan2(y, x)


def segment_sum(x, counts):
    seg = torch.repeat_interleave(torch.arange(len(counts)), torch.as_tensor(counts).long())
    return torch.zeros((len(counts),) + tuple(x.shape[1:]), dtype=x.dtype).index_add(0, seg, x)


def energy_forward(xyz, idxs: Dict[str, torch.Tensor], params: Dict[str, Dict[str, torch.Tensor]],
                   counts: Dict[str, Sequence[int]], terms=LEVELS, gradients=True, create_graph=False):
    """Energy.forward: models/energy.py:99-145 (+ :8-71).  Returns dict with energy (B,C), per-term
    energies, per-tuple x / energy and gradient (N,C,3) = +dE/dxyz."""
    xyz = xyz.detach().clone().requires_grad_(gradients)
    out = {"x": {}, "tuple_energy": {}, "term_energy": {}}
    n_mols = len(next(iter(counts.values())))
    total = torch.zeros((n_mols, xyz.shape[1]), dtype=xyz.dtype)
    for lvl in terms:
        idx = idxs[lvl].long()
        if idx.shape[0] == 0:
            q = torch.zeros((0, xyz.shape[1]), dtype=xyz.dtype)
        else:
            pos = xyz[idx]                                                 # (T, L, C, 3)
            if lvl == "n2":
                q = bond_length(pos[:, 0], pos[:, 1])
            elif lvl == "n3":
                q = bond_angle(pos[:, 0], pos[:, 1], pos[:, 2])
            else:
                q = dihedral_angle(pos[:, 0], pos[:, 1], pos[:, 2], pos[:, 3])
        k = params[lvl]["k"]
        if lvl in ("n2", "n3"):
            e = 0.5 * k[:, None] * torch.square(q - params[lvl]["eq"][:, None])
        else:
            n = torch.arange(1, k.shape[1] + 1, dtype=xyz.dtype)[None, :, None]
            e = (k[:, :, None] * torch.cos(n * q[:, None, :])).sum(1)
        out["x"][lvl] = q.detach()
        out["tuple_energy"][lvl] = e
        pooled = segment_sum(e, counts[lvl])
        out["term_energy"][lvl] = pooled.detach()
        total = total + pooled
    out["energy"] = total
    if gradients:
        out["gradient"] = torch.autograd.grad(total.sum(), xyz, create_graph=create_graph, retain_graph=True)[0]
    return out


# --------------------------------------------------------------------------------------------------
# whole path
# --------------------------------------------------------------------------------------------------
def model_forward(sd, g, cfg, dtype=torch.float32, taps: dict | None = None):
    """GrappaModel.forward (models/grappa.py:111-132): returns h and the parameter dict per level."""
    h = gnn_forward(sd, g, cfg, dtype=dtype, taps=taps)
    params = {}
    for lvl in LEVELS:
        idx = g.nodes[lvl].data["idxs"]
        if idx.shape[0] == 0 and lvl in ("n4", "n4_improper"):
            n = int(sd[f"parameter_writer.{WRITER[lvl][0]}.n_periodicity"])
            params[lvl] = {"k": torch.zeros((0, n), dtype=dtype)}
            continue
        scores = writer_scores(sd, h, idx, lvl, cfg, taps=taps)
        if taps is not None:
            taps[f"{lvl}_scores"] = scores
        params[lvl] = writer_params(sd, scores, lvl, cfg)
    return h, params


def path_forward(sd, g, cfg, dtype=torch.float32, gradients=True, create_graph=False):
    """Sequential(GrappaModel, Energy) as in training/trainrun.py:112-116."""
    h, params = model_forward(sd, g, cfg, dtype=dtype)
    idxs = {l: g.nodes[l].data["idxs"] for l in LEVELS}
    counts = {l: g.batch_num_nodes(l).tolist() for l in LEVELS}
    en = energy_forward(g.nodes["n1"].data["xyz"].to(dtype), idxs, params, counts, gradients=gradients,
                        create_graph=create_graph)
    return h, params, en


def molwise_loss(en, params, g, energy_weight=1.0, gradient_weight=0.8, proper_reg=1e-3, improper_reg=1e-3, n_valid=None):
    """MolwiseLoss.forward (training/loss.py:45-167) restricted to the terms active in the grappa-1.2
    config with no reference parameters on the graph: centred-energy MSE + gradient MSE + torsion L2
    (the improper regulariser enters twice, loss.py:127-132).  Mean over molecules."""
    a_counts = g.batch_num_nodes("n1").tolist()
    p_counts = g.batch_num_nodes("n4").tolist()
    i_counts = g.batch_num_nodes("n4_improper").tolist()
    e, e_ref = en["energy"], g.nodes["g"].data["energy_ref"].to(en["energy"].dtype)
    gr, gr_ref = en["gradient"], g.nodes["n1"].data["gradient_ref"].to(en["energy"].dtype)
    loss = 0.0
    a0 = p0 = i0 = 0
    nb = len(a_counts)
    for b in range(nb):
        # padding conformations (is_dummy) are removed per molecule by unbatch() -> delete_dummy_confs
        # (utils/dgl_utils.py:63-118) before any term is evaluated; they are the trailing ones (:155-157)
        nv = e.shape[1] if n_valid is None else int(n_valid[b])
        eb = e[b, :nv] - e[b, :nv].mean()
        rb = e_ref[b, :nv] - e_ref[b, :nv].mean()
        term = energy_weight * torch.mean((eb - rb) ** 2)
        term = term + gradient_weight * torch.mean((gr[a0:a0 + a_counts[b], :nv] - gr_ref[a0:a0 + a_counts[b], :nv]) ** 2)
        kp = params["n4"]["k"][p0:p0 + p_counts[b]]
        if len(kp) > 0:
            term = term + proper_reg * torch.mean(kp ** 2)
        ki = params["n4_improper"]["k"][i0:i0 + i_counts[b]]
        if len(ki) > 0:
            term = term + 2.0 * improper_reg * torch.mean(ki ** 2)
        loss = loss + term / nb
        a0 += a_counts[b]; p0 += p_counts[b]; i0 += i_counts[b]
    return loss


def param_loss(params, params_ref, counts, param_weight=1e-3, mol_weights=None,
               weights={"n2_k": 1e-3, "n3_k": 1e-2, "n4_k": 1e-4}):
    """Classical-parameter term of MolwiseLoss.forward (training/loss.py:70-113), mean over molecules:
    per molecule, the concatenation (reference key order n2_k, n2_eq, n3_k, n3_eq, n4_k; impropers skipped, :91-92) of
    fac * parameter with NaN references zeroed on both sides (:101-103), torsion references padded / cut to the model's
    periodicity (correct_torsion_shape, :170-182), mean of squared differences times the molecule's weight.
    `params[lvl][name]`, `params_ref[lvl][name]`: tensors; `counts[lvl]`: tuples per molecule."""
    nb = len(counts["n2"])
    off = {l: [0] * (nb + 1) for l in counts}
    for l in counts:
        for b in range(nb):
            off[l][b + 1] = off[l][b] + counts[l][b]
    loss = 0.0
    for b in range(nb):
        pt, rt = [], []
        for lvl, name in (("n2", "k"), ("n2", "eq"), ("n3", "k"), ("n3", "eq"), ("n4", "k")):
            if lvl not in params_ref or name not in params_ref[lvl]:
                continue
            p = params[lvl][name][off[lvl][b]:off[lvl][b + 1]]
            r = params_ref[lvl][name][off[lvl][b]:off[lvl][b + 1]].to(p.dtype)
            if lvl == "n4":
                if r.shape[1] < p.shape[1]:
                    r = torch.cat([r, torch.zeros_like(r[:, :(p.shape[1] - r.shape[1])])], dim=1)
                elif r.shape[1] > p.shape[1]:
                    r = r[:, :p.shape[1]]
            nan = torch.isnan(r)
            p = torch.where(nan, torch.zeros_like(p), p)
            r = torch.where(nan, torch.zeros_like(r), r)
            fac = weights.get(f"{lvl}_{name}", 1.0)
            pt.append(p.flatten() * fac)
            rt.append(r.flatten() * fac)
        if pt:
            w = param_weight if mol_weights is None else mol_weights[b]
            loss = loss + w * torch.mean((torch.cat(pt) - torch.cat(rt)) ** 2) / nb
    return loss
