"""Drop-in torch.nn.Module replacements for grappa.models (GNN, tuple heads, model assembly).

Boundary: same constructor arguments, same attribute tree and therefore the
same `state_dict` keys / shapes as the reference (410 entries for grappa-1.2, including DGL's
`graph_module.fc.weight`, torch's `attn.in_proj_weight` and the aliased `gnn.blocks.*`), same graph
fields read and written.  The torch sub-modules below are PARAMETER CONTAINERS only: their
`forward` is never called; the arithmetic runs in the sm_100a kernels behind the C ABI, orchestrated
by `tape.py`, and gradients come from hand-written backward kernels.

Reference classes mirrored (file:line in /root/reference/src/grappa/models/):
  GrappaGNN graph_attention.py:11-183 | ResidualAttentionBlock :188-310 | PositionalEncoding :418-444
  WriteParameters interaction_parameters.py:10-135 | RepProjector :140-180 | WriteBondParameters :183-266
  WriteAngleParameters :270-362 | WriteTorsionParameters :368-562
  SymmetrisedTransformer perm_equiv_transformer.py:13-70 | GrappaTransformer :75-190 | Symmetriser :194-323
  FeedForwardLayer network_utils.py:5-54 | DottedAttWithMLP :57-133 | HardCutoff :136-145
  ToPositive final_layer.py:11-52 | ToRange :54-97 | GrappaModel grappa.py:7-132 | deploy.py:8-64
"""
from __future__ import annotations

import copy
import os
from typing import Dict, List, Union

import torch
from torch import nn

from . import _lib, ops, tape as T_
from ._lib_ops import HeadOutArgs
from .pack import get_pack
from .tape import Tape, Var

MAX_ELEMENT = 53                      # reference constants.py:38
CHARGE_MODELS = ["am1BCC", "amber99"]  # reference constants.py:44
N_PERIODICITY_PROPER = 6
N_PERIODICITY_IMPROPER = 6
IMPROPER_CENTRAL_IDX = 2
ELU = 1


def get_default_statistics():
    """Default parameter statistics (reference utils/graph_utils.py:233-241; plain numbers)."""
    return {
        "mean": {"n2_k": torch.Tensor([763.2819]), "n2_eq": torch.Tensor([1.2353]), "n3_k": torch.Tensor([105.6576]),
                 "n3_eq": torch.Tensor([1.9750]),
                 "n4_k": torch.Tensor([1.5617e-01, -5.8312e-01, 7.0820e-02, -6.3840e-04, 4.7139e-04, -4.1655e-04]),
                 "n4_improper_k": torch.Tensor([0.0000, -2.3933, 0.0000])},
        "std": {"n2_k": torch.Tensor([161.2278]), "n2_eq": torch.Tensor([0.1953]), "n3_k": torch.Tensor([26.5965]),
                "n3_eq": torch.Tensor([0.0917]),
                "n4_k": torch.Tensor([0.4977, 1.2465, 0.1466, 0.0192, 0.0075, 0.0066]),
                "n4_improper_k": torch.Tensor([0.0000, 4.0571, 0.0000])}}


# ==================================================================================================
# stage runner: one torch.autograd.Function per fused stage, backward = tape replay
# ==================================================================================================
_BACKWARD_HOOK = None


def set_backward_hook(fn):
    """fn(tag) is called during backward when all gradients of a stage (('writer', module)), of one GNN
    block (('gnn_block', i)) or of the rest of the GNN (('gnn_rest', None)) are final -- used by
    training.Trainer to start the bucketed gradient all-reduce while backward is still running.
    Inside a writer, ('writer_part', (module, k, events)) fires when part k of `writer_parts(module)` is final once
    `events` have completed (the backward chain itself does not wait for them)."""
    global _BACKWARD_HOOK
    _BACKWARD_HOOK = fn


def _fire(tag):
    if _BACKWARD_HOOK is not None:
        _BACKWARD_HOOK(tag)


class _StageFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, runner, record, train, seed, n_inputs, *tensors):
        tape = Tape(record, train, seed)
        ins = [Var(t.detach().contiguous().float(), needs=bool(t.requires_grad)) for t in tensors[:n_inputs]]
        params = [t.detach() for t in tensors[n_inputs:]]
        for raw, det in zip(tensors[n_inputs:], params):
            sink = getattr(raw, "_gb_sink", None)
            if sink is not None:
                tape.sinks[id(det)] = sink
        ctx.tag = getattr(runner, "tag", None)
        with ops.stage_tag(ctx.tag):
            outs = runner(tape, ins, params)
        ctx.tape, ctx.ins, ctx.params, ctx.outs = tape, ins, params, outs
        ctx.concurrent = bool(tensors) and tensors[0].is_cuda and T_.concurrency()
        ctx.set_materialize_grads(False)
        return tuple(o.v for o in outs)

    @staticmethod
    def backward(ctx, *gouts):
        tape = ctx.tape
        for o, g in zip(ctx.outs, gouts):
            o.g = None if g is None else g.contiguous().float()
        # the backward of every stage runs next to its weight-gradient branch (and the writers next to each other):
        # leave part of the machine to the other streams' kernels
        with ops.stage_tag(ctx.tag), ops.gemm_sm_limit(ops.CONCURRENT_GEMM_SMS if ctx.concurrent else 0):
            tape.backward()
        gin = [v.g if v.needs else None for v in ctx.ins]
        gp = [None if id(p) in tape.sinks else tape.pgrads.get(id(p)) for p in ctx.params]
        ctx.tape = ctx.outs = None
        return (None, None, None, None, None, *gin, *gp)


def _run_stage(runner, inputs: List[torch.Tensor], params: List[torch.Tensor], train: bool):
    record = torch.is_grad_enabled() and any(t.requires_grad for t in list(inputs) + list(params))
    seed = int(torch.randint(0, 2 ** 31 - 1, (1,)).item()) if train else 0
    return _StageFn.apply(runner, record, train, seed, len(inputs), *inputs, *params)


# ==================================================================================================
# parameter containers (attribute names = reference state_dict keys)
# ==================================================================================================
class DotGatConv(nn.Module):
    """Holds DGL DotGatConv's single parameter `fc.weight` (bias-free, shared by source and destination)."""

    def __init__(self, in_feats, out_feats, num_heads):
        super().__init__()
        self._out_feats, self._num_heads = out_feats, num_heads
        self.fc = nn.Linear(in_feats, out_feats * num_heads, bias=False)


class ResidualAttentionBlock(nn.Module):
    def __init__(self, in_feats, out_feats=None, num_heads=8, self_interaction=True, layer_norm=True,
                 dropout_behind_mha=0.1, dropout_behind_self_interaction=0.1, skip_connection=True):
        super().__init__()
        out_feats = in_feats if out_feats is None else out_feats
        if not (skip_connection and out_feats == in_feats):
            raise NotImplementedError("grappa_b200 implements the block GrappaGNN builds (graph_attention.py:113-120): "
                                      "skip connection enabled, out_feats == in_feats")
        assert out_feats % num_heads == 0
        self.in_feats, self.out_feats, self.num_heads = in_feats, out_feats, num_heads
        self.feats_per_head = out_feats // num_heads
        self.graph_module = DotGatConv(in_feats, out_feats // num_heads, num_heads)
        self.p1, self.p2 = dropout_behind_mha, dropout_behind_self_interaction
        self.do_layer_norm = layer_norm
        if layer_norm:
            self.layer_norm = nn.LayerNorm(in_feats)
        self.head_reducer = nn.Linear(num_heads * self.feats_per_head, out_feats)
        if self_interaction:
            if layer_norm:
                self.interaction_norm = nn.LayerNorm(out_feats)
            self.self_interaction = nn.Sequential(nn.Linear(out_feats, 4 * out_feats), nn.ELU(),
                                                  nn.Linear(4 * out_feats, out_feats), nn.ELU())
        else:
            self.self_interaction = None

    def tape_forward(self, t: Tape, pack, h: Var, P) -> Var:
        """graph_attention.py:276-310: u=LN(h); y=W_r attn(u W_fc)+u; z=LN(y); out=ELU(W2 ELU(W1 z))+z  (the two
        LayerNorms only with layer_norm=True, the second half only with self_interaction=True)."""
        u = T_.layernorm(t, h, P(self.layer_norm.weight), P(self.layer_norm.bias)) if self.do_layer_norm else h
        ft = T_.linear(t, u, P(self.graph_module.fc.weight), None)
        m = T_.edge_attention(t, ft, pack, self.num_heads)
        y = T_.linear(t, m, P(self.head_reducer.weight), P(self.head_reducer.bias), dropout_p=self.p1, residual=u)
        if self.self_interaction is None:
            return y
        z = T_.layernorm(t, y, P(self.interaction_norm.weight), P(self.interaction_norm.bias)) if self.do_layer_norm else y
        a1 = T_.linear(t, z, P(self.self_interaction[0].weight), P(self.self_interaction[0].bias), act=ELU,
                       fuse_elu_into_consumer=True)
        return T_.linear(t, a1, P(self.self_interaction[2].weight), P(self.self_interaction[2].bias), act=ELU,
                         dropout_p=self.p2, residual=z)


class SAGEConv(nn.Module):
    """Parameters of dgl.nn.SAGEConv(in, out, 'mean') as the oracle's dgl shim restates it (DGL is unpinned and absent:
    `fc_self` Linear with bias, bias-free `fc_neigh`; rst = fc_self(h) + fc_neigh(mean_{u in N(v)} h_u)).  DGL >= 1.1
    keeps the bias as a separate `bias` parameter: such state_dicts are remapped on load."""

    def __init__(self, in_feats, out_feats, aggregator_type="mean"):
        super().__init__()
        if aggregator_type != "mean":
            raise NotImplementedError("only the 'mean' aggregator (the one grappa uses) is implemented")
        self.fc_self = nn.Linear(in_feats, out_feats)
        self.fc_neigh = nn.Linear(in_feats, out_feats, bias=False)

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        if prefix + "bias" in state_dict and prefix + "fc_self.bias" not in state_dict:
            state_dict[prefix + "fc_self.bias"] = state_dict.pop(prefix + "bias")
        return super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)


def to_dgl_state_dict(state_dict: Dict[str, torch.Tensor], separate_sage_bias: bool = True) -> Dict[str, torch.Tensor]:
    """Checkpoint export for the reference running on DGL >= 0.8, whose SAGEConv keeps its bias as a separate
    `graph_module.bias` parameter (older DGL: `graph_module.fc_self.bias`, the layout `state_dict()` emits here; loading
    accepts both).  Only convolution models (gnn_convolutions > 0, grappa-1.0) are affected -- grappa-1.1 / 1.2 have none."""
    if not separate_sage_bias:
        return dict(state_dict)
    out = {}
    for k, v in state_dict.items():
        if k.endswith("graph_module.fc_self.bias"):
            k = k[:-len("fc_self.bias")] + "bias"
        out[k] = v
    return out


class ResidualConvBlock(nn.Module):
    """grappa-1.0 convolution block (reference models/graph_attention.py:314-415): u = LN(h);
    y = dropout(ELU(SAGEConv(u))) + u; z = LN(y); out = dropout(ELU(W z + b)) + z."""

    def __init__(self, in_feats, out_feats=None, self_interaction=True, layer_norm=True, dropout=0.0, skip_connection=True):
        super().__init__()
        out_feats = in_feats if out_feats is None else out_feats
        if not (skip_connection and out_feats == in_feats):
            raise NotImplementedError("grappa_b200 implements the block GrappaGNN builds (graph_attention.py:106-109): "
                                      "skip connection enabled, out_feats == in_feats")
        self.in_feats, self.out_feats, self.p = in_feats, out_feats, dropout
        self.graph_module = SAGEConv(in_feats, out_feats, "mean")
        self.do_layer_norm = layer_norm
        if layer_norm:
            self.layer_norm = nn.LayerNorm(in_feats)
        if self_interaction:
            self.self_interaction = nn.Sequential(nn.Linear(out_feats, out_feats), nn.ELU())
            if layer_norm:
                self.interaction_norm = nn.LayerNorm(out_feats)
        else:
            self.self_interaction = None

    def tape_forward(self, t: Tape, pack, h: Var, P) -> Var:
        u = T_.layernorm(t, h, P(self.layer_norm.weight), P(self.layer_norm.bias)) if self.do_layer_norm else h
        s_self = T_.linear(t, u, P(self.graph_module.fc_self.weight), None)
        hn = T_.neighbor_mean(t, u, pack)
        y = T_.linear(t, hn, P(self.graph_module.fc_neigh.weight), P(self.graph_module.fc_self.bias), act=ELU,
                      dropout_p=self.p, residual=u, pre_add=s_self)
        if self.self_interaction is None:
            return y
        z = T_.layernorm(t, y, P(self.interaction_norm.weight), P(self.interaction_norm.bias)) if self.do_layer_norm else y
        return T_.linear(t, z, P(self.self_interaction[0].weight), P(self.self_interaction[0].bias), act=ELU,
                         dropout_p=self.p, residual=z)


class GrappaGNN(nn.Module):
    """Atom featurisation -> residual graph-attention blocks -> atom embedding g.nodes['n1'].data['h']."""

    def __init__(self, out_feats: int = 512, in_feats: int = None, node_feats: int = None, n_conv: int = 3, n_att: int = 3,
                 n_heads: int = 8, in_feat_name: Union[str, List[str]] = ["atomic_number", "ring_encoding", "partial_charge"],
                 in_feat_dims: Dict[str, int] = {}, conv_dropout: float = 0., attention_dropout: float = 0.,
                 final_dropout: float = 0., initial_dropout: float = 0., layer_norm: bool = True,
                 self_interaction: bool = True, charge_encoding=True):
        super().__init__()
        self.charge_encoding = charge_encoding
        if not isinstance(in_feat_name, list):
            in_feat_name = [in_feat_name]
        if in_feats is None:
            dims = {"atomic_number": MAX_ELEMENT, "ring_encoding": 7, "partial_charge": 1, "sp_hybridization": 6,
                    "mass": 2, "degree": 6, "is_radical": 1, "laplacian_positional_encoding": 5,
                    "charge_model": len(CHARGE_MODELS)}
            dims.update(in_feat_dims)
            in_feats = sum(dims[f] for f in in_feat_name)
        node_feats = out_feats if node_feats is None else node_feats
        self.in_feats = in_feats + (16 if charge_encoding else 0)
        self.in_feat_name = in_feat_name
        self.p_initial, self.p_final = initial_dropout, final_dropout
        self.pre_dense = nn.Sequential(nn.Linear(self.in_feats, node_feats), nn.ELU())
        self.no_convs = (n_conv + n_att) == 0
        if not self.no_convs:
            self.conv_blocks = nn.ModuleList([
                ResidualConvBlock(node_feats, node_feats, self_interaction, layer_norm, conv_dropout) for _ in range(n_conv)])
            self.att_blocks = nn.ModuleList([
                ResidualAttentionBlock(node_feats, node_feats, n_heads, self_interaction, layer_norm, attention_dropout,
                                       attention_dropout, True) for _ in range(n_att)])
            self.post_dense = nn.Sequential(nn.Linear(node_feats, out_feats))
            self.blocks = self.conv_blocks + self.att_blocks       # alias, as reference graph_attention.py:129
        else:
            self.post_dense = nn.Sequential(nn.Linear(node_feats, out_feats))

    def _runner(self, pack, plist):
        index = {id(p): i for i, p in enumerate(plist)}

        def run(t: Tape, ins, params):
            P = lambda p: params[index[id(p)]]
            x = ins[0]
            t.push(lambda: (t.join_side(), _fire(("gnn_rest", None))))          # runs last in backward
            h = T_.linear(t, x, P(self.pre_dense[0].weight), P(self.pre_dense[0].bias), act=ELU,
                          dropout_p=self.p_initial, k=self.in_feats, x_pad_is_zero=True)   # featurize zero-fills the pad
            if not self.no_convs:
                for blk in self.conv_blocks:                     # grappa-1.0 only; their gradients are final with 'gnn_rest'
                    h = blk.tape_forward(t, pack, h, P)
                for i, blk in enumerate(self.att_blocks):
                    if _BACKWARD_HOOK is not None:               # runs after block i's backward ops
                        t.push(lambda i=i: (t.join_side(), _fire(("gnn_block", i))))
                    h = blk.tape_forward(t, pack, h, P)
            h = T_.linear(t, h, P(self.post_dense[0].weight), P(self.post_dense[0].bias), dropout_p=self.p_final)
            return [h]
        run.tag = "gnn"
        return run

    def forward(self, g, in_feature=None):
        data = g.nodes["n1"].data
        if in_feature is None:
            feats = [data[f] for f in self.in_feat_name]
        else:
            feats = [in_feature]
        _lib.require_cuda(*feats)
        for f in feats:
            if f.dim() > 2:
                raise AssertionError(f"the input features must be of shape (n_nodes, n_features), but got {f.shape}")
        ld = (self.in_feats + 3) // 4 * 4
        x = ops.featurize(feats, data["partial_charge"] if self.charge_encoding else None, ld)
        pack = get_pack(g)
        plist = list(self.parameters())
        (h,) = _run_stage(self._runner(pack, plist), [x], plist, self.training)
        data["h"] = h
        return g


# ---- tuple heads ---------------------------------------------------------------------------------
class ToPositive(nn.Module):
    """std * (ELU(mean/std + x - 1) + 1) + min  (buffers only; evaluated inside head_output kernels)."""

    def __init__(self, mean, std, min_=0., learnable_statistics=False):
        super().__init__()
        # the reference divides by the fp32-rounded std (final_layer.py:30-41): keep the same rounding
        std_t = torch.tensor(float(std)).float()
        self.learnable = bool(learnable_statistics)
        if self.learnable:      # final_layer.py:37-39: trained along with the weights, read from device memory by the kernels
            self.mean_over_std = nn.Parameter(torch.tensor(float(mean / std_t)))
            self.std = nn.Parameter(torch.tensor(float(std)))
        else:
            self.register_buffer("mean_over_std", torch.tensor(float(mean / std_t)))
            self.register_buffer("std", torch.tensor(float(std)))
        self.register_buffer("min_", torch.tensor(float(min_)))


class ToRange(nn.Module):
    """max * sigmoid(std/max * x)."""

    def __init__(self, max_, std, learnable_statistics=False):
        super().__init__()
        self.learnable = bool(learnable_statistics)
        if self.learnable:      # final_layer.py:84-85
            self.std_over_max = nn.Parameter(torch.tensor(float(std / max_)).float())
        else:
            self.register_buffer("std_over_max", torch.tensor(float(std / max_)).float())
        self.register_buffer("max", torch.tensor(float(max_)).float())


class HardCutoff(nn.Module):
    def __init__(self, cutoff=0.):
        super().__init__()
        self.cutoff = cutoff


class RepProjector(nn.Module):
    def __init__(self, dim_tupel, in_feats, out_feats, improper: bool = False):
        super().__init__()
        self.dim_tupel, self.improper = dim_tupel, improper
        self.mlp = nn.Sequential(nn.Linear(in_feats, out_feats), nn.ELU())


class FeedForwardLayer(nn.Module):
    def __init__(self, in_feats, hidden_feats=None, out_feats=None, dropout=0., skip=False, layer_norm=True):
        super().__init__()
        hidden_feats = in_feats if hidden_feats is None else hidden_feats
        out_feats = in_feats if out_feats is None else out_feats
        self.in_feats, self.hidden_feats, self.out_feats = in_feats, hidden_feats, out_feats
        self.linear1 = nn.Linear(in_feats, hidden_feats)
        self.linear2 = nn.Linear(hidden_feats, out_feats)
        self.p, self.skip = dropout, skip
        self.layer_norm = layer_norm
        if layer_norm:
            self.norm1 = nn.LayerNorm(in_feats)
        assert (out_feats == in_feats) or not skip, \
            f"Skip connection is not possible with {in_feats} input features and {out_feats} output features."

    def tape_forward(self, t: Tape, x: Var, P) -> Var:
        """network_utils.py:44-54: xn = LN(x) (x itself with layer_norm=False); W2 ELU(W1 xn) (+ xn if skip)."""
        xn = T_.layernorm(t, x, P(self.norm1.weight), P(self.norm1.bias)) if self.layer_norm else x
        f1 = T_.linear(t, xn, P(self.linear1.weight), P(self.linear1.bias), act=ELU, fuse_elu_into_consumer=True)
        return T_.linear(t, f1, P(self.linear2.weight), P(self.linear2.bias), dropout_p=self.p,
                         residual=xn if self.skip else None)


class DottedAttWithMLP(nn.Module):
    def __init__(self, n_feats, num_heads, hidden_feats=None, layer_norm=True, dropout=0.):
        super().__init__()
        hidden_feats = 4 * n_feats if hidden_feats is None else hidden_feats
        assert n_feats % num_heads == 0, \
            f"Number of features ({n_feats}) must be divisible by the number of heads ({num_heads})."
        self.n_feats, self.num_heads, self.p = n_feats, num_heads, dropout
        self.layer_norm = layer_norm
        if layer_norm:
            self.norm1 = nn.LayerNorm(n_feats)
        self.attn = nn.MultiheadAttention(n_feats, num_heads, dropout=0)
        self.ff = FeedForwardLayer(n_feats, hidden_feats, out_feats=n_feats, dropout=dropout, skip=True, layer_norm=layer_norm)

    def tape_forward(self, t: Tape, x: Var, n_tuples: int, L: int, P) -> Var:
        """network_utils.py:112-133: xn = LN(x) (x itself with layer_norm=False); x = MHA(xn) + xn; x = FF(x)."""
        xn = T_.layernorm(t, x, P(self.norm1.weight), P(self.norm1.bias)) if self.layer_norm else x
        qkv = T_.linear(t, xn, P(self.attn.in_proj_weight), P(self.attn.in_proj_bias))
        att = T_.tuple_attention(t, qkv, n_tuples, L, self.num_heads)
        x1 = T_.linear(t, att, P(self.attn.out_proj.weight), P(self.attn.out_proj.bias), dropout_p=self.p, residual=xn)
        return self.ff.tape_forward(t, x1, P)


class GrappaTransformer(nn.Module):
    def __init__(self, n_feats, n_heads, hidden_feats, n_layers, out_feats, permutations, layer_norm=True, dropout=0.0,
                 positional_encoding=None):
        super().__init__()
        self.n_layers, self.out_feats = n_layers, out_feats
        self.n_seq = permutations.shape[1]
        pe = None
        if isinstance(positional_encoding, torch.Tensor):
            pe = positional_encoding.float().clone()
        elif isinstance(positional_encoding, bool) and positional_encoding:
            plist = [p.tolist() for p in permutations]
            if self.n_seq == 3 and all(p in [[0, 1, 2], [2, 1, 0]] for p in plist):
                pe = torch.tensor([[0], [1], [0]], dtype=torch.float32)
            elif self.n_seq == 4 and all(p in [[0, 1, 2, 3], [3, 2, 1, 0], [0, 2, 1, 3], [3, 1, 2, 0]] for p in plist):
                pe = torch.tensor([[0], [1], [1], [0]], dtype=torch.float32)
        if pe is not None:
            if pe.shape[1] != 1:
                raise NotImplementedError("positional encodings wider than one column are not supported")
            self.register_buffer("positional_encoding", pe)
        else:
            self.positional_encoding = None
        self.n_feats = n_feats + (pe.shape[1] if pe is not None else 0)
        if self.n_feats % n_heads != 0:
            raise ValueError(f"The number of input features cannot be divided by the number of heads: "
                             f"Number of input features: {self.n_feats}. Number of heads: {n_heads}")
        self.transformer = nn.Sequential(*[DottedAttWithMLP(self.n_feats, n_heads, hidden_feats, layer_norm, dropout)
                                           for _ in range(n_layers)])


class Symmetriser(nn.Module):
    def __init__(self, in_feats, out_feats, permutations, permutation_prefactors=None, hidden_feats=None, n_layers=1,
                 skip=True, layer_norm=True):
        super().__init__()
        assert n_layers >= 1, "n_layers must be >= 1"
        hidden_feats = in_feats if hidden_feats is None else hidden_feats
        perms = permutations.int() if isinstance(permutations, torch.Tensor) else torch.tensor(permutations, dtype=torch.int32)
        assert perms.dim() == 2 and perms.shape[0] > 0
        self.n_perm, self.n_seq = perms.shape
        assert torch.all(perms[0] == torch.arange(self.n_seq).int()), \
            "permutations must include the identity permutation at the zeroth entry."
        if permutation_prefactors is None:
            permutation_prefactors = torch.ones(self.n_perm, dtype=torch.float32)
        elif not torch.all(torch.as_tensor(permutation_prefactors) == 1):
            raise NotImplementedError("permutation prefactors other than 1 are not supported")
        self.n_feats, self.out_feats = in_feats, out_feats
        self.register_buffer("permutation_prefactors", torch.as_tensor(permutation_prefactors).float().view(self.n_perm, 1, 1))
        self.register_buffer("permutations", perms)
        self.mlp = nn.Sequential(
            FeedForwardLayer(in_feats * self.n_seq, hidden_feats, hidden_feats if n_layers > 1 else out_feats, skip=False,
                             layer_norm=layer_norm),
            *[FeedForwardLayer(hidden_feats, hidden_feats, hidden_feats if i != n_layers - 1 else out_feats,
                               skip=skip if i != n_layers - 1 else False, layer_norm=layer_norm) for i in range(1, n_layers)])
        self._perms_c = ops.make_perms(perms.tolist())


class SymmetrisedTransformer(nn.Module):
    def __init__(self, n_feats, n_heads, hidden_feats, n_layers, out_feats, permutations, layer_norm=True, dropout=0.0,
                 symmetriser_layers=1, symmetriser_hidden_feats=None, permutation_prefactors=None, positional_encoding=None):
        super().__init__()
        if n_layers > 0:
            self.grappa_transformer = GrappaTransformer(n_feats, n_heads, hidden_feats, n_layers, n_feats, permutations,
                                                        layer_norm, dropout, positional_encoding)
            trafo_out = self.grappa_transformer.n_feats
        else:
            self.grappa_transformer = None
            trafo_out = n_feats
        assert symmetriser_layers >= 1, "symmetriser_layers must be >= 1"
        self.symmetriser = Symmetriser(trafo_out, out_feats, permutations, permutation_prefactors, symmetriser_hidden_feats,
                                       symmetriser_layers, layer_norm=layer_norm)


class _TupleWriter(nn.Module):
    """Shared forward of the four writers: projector -> gather(+PE) -> transformer -> symmetriser -> maps."""
    level_id: int = 0
    level: str = "n2"
    model_attr: str = "bond_model"

    def _model(self) -> SymmetrisedTransformer:
        return getattr(self, self.model_attr)

    def _head_args(self, T: int) -> HeadOutArgs:
        raise NotImplementedError

    def parts(self):
        """Sub-modules whose gradients become final one after the other during this writer's backward pass (the
        'writer_part' hook): symmetriser first, then the transformer layers from the last to the first.  Whatever else
        the writer owns (projector, learnable statistics) is final at the ('writer', module) hook."""
        m = self._model()
        layers = list(m.grappa_transformer.transformer) if m.grappa_transformer is not None else []
        return [m.symmetriser] + layers[::-1]

    def _stat_params(self):
        """learnable_statistics: the statistics parameters in the slot order of gb_head_out_args.stat, else None."""
        return None

    def _host(self, name: str, buf: torch.Tensor):
        """Host copy of a small statistics buffer, refreshed when the buffer is modified."""
        cache = self.__dict__.setdefault("_host_cache", {})
        key = (name, buf.data_ptr(), buf._version)
        if cache.get(name, (None,))[0] != key:
            cache[name] = (key, buf.detach().cpu().flatten().tolist())
        return cache[name][1]

    def _runner(self, pack, plist, T):
        index = {id(p): i for i, p in enumerate(plist)}
        model = self._model()
        L = self.rep_projector.dim_tupel
        F = self.rep_projector.mlp[0].out_features
        gt = model.grappa_transformer
        E = gt.n_feats if gt is not None else F
        pe = gt.positional_encoding.flatten().contiguous() if (gt is not None and gt.positional_encoding is not None) else None
        sym = model.symmetriser
        args = self._head_args(T)
        stat_params = self._stat_params()

        n_layers = len(gt.transformer) if gt is not None else 0

        def part_done(t: Tape, k: int):
            # backward closures run in reverse order of their push: this one right after part k's backward ops
            if _BACKWARD_HOOK is not None:
                t.push(lambda: _fire(("writer_part", (self, k, t.flush_events()))))

        def run(t: Tape, ins, params):
            P = lambda p: params[index[id(p)]]
            h = ins[0]
            t.push(lambda: (t.join_side(), _fire(("writer", self))))
            proj = T_.linear(t, h, P(self.rep_projector.mlp[0].weight), P(self.rep_projector.mlp[0].bias), act=ELU,
                             out_ld=(E if E % 4 == 0 else (E + 3) // 4 * 4))
            x = T_.tuple_gather(t, proj, pack, self.level_id, pe, F, E)
            if gt is not None:
                for li, layer in enumerate(gt.transformer):
                    part_done(t, n_layers - li)          # parts in backward order: 0 = symmetriser, 1 = last layer, ...
                    x = layer.tape_forward(t, x, T, L, P)
            part_done(t, 0)
            s = T_.perm_concat(t, x, sym._perms_c, T, L, E)
            for ff in sym.mlp:
                s = ff.tape_forward(t, s, P)
            stats = None
            if stat_params is not None:
                # the kernels read the (trainable) statistics from the parameters' own device memory
                stats = [None if p is None else P(p) for p in stat_params]
                for i, sp in enumerate(stats):
                    args.stat[i] = None if sp is None else sp.data_ptr()
            k, eq = ops.head_output_fwd(args, s.v)
            kv, eqv = Var(k), (Var(eq) if eq is not None else None)

            def bwd():
                if kv.g is None and (eqv is None or eqv.g is None):
                    return
                deq = None if eqv is None else eqv.g
                if stats is not None:
                    targets, accs = [], []
                    for sp in stats:
                        if sp is None:
                            targets.append(None)
                            continue
                        tgt, acc = t.grad_target(sp)
                        if tgt is None:
                            tgt, acc = torch.empty_like(sp), False
                            t.pgrads[id(sp)] = tgt
                        targets.append(tgt)
                        accs.append(acc)
                    assert len(set(accs)) == 1
                    ops.head_output_stats_bwd(args, s.v, kv.g, deq, targets, accs[0])
                T_.add_grad(s, ops.head_output_bwd(args, s.v, kv.g, deq))
            t.push(bwd)
            return [kv] if eqv is None else [kv, eqv]
        run.tag = "writer"
        return run

    def _write(self, g, h):
        pack = get_pack(g)
        T = pack.n_tuples[self.level_id]
        plist = list(self.parameters())
        return _run_stage(self._runner(pack, plist, T), [h], plist, self.training)


def _stat(d, key):
    v = d[key]
    return v.item() if isinstance(v, torch.Tensor) and v.numel() == 1 else v


class WriteBondParameters(_TupleWriter):
    level_id, level, model_attr = 0, "n2", "bond_model"

    def __init__(self, rep_feats, between_feats, suffix="", param_statistics=None, n_att=2, n_heads=8, dense_layers=2,
                 dropout=0., layer_norm=True, symmetriser_feats=None, attention_hidden_feats=None, positional_encoding=True,
                 learnable_statistics: bool = False, gate: bool = False):
        super().__init__()
        EPS = 1e-6
        st = get_default_statistics() if param_statistics is None else param_statistics
        k_mean, k_std = _stat(st["mean"], "n2_k"), _stat(st["std"], "n2_k") + EPS
        eq_mean, eq_std = _stat(st["mean"], "n2_eq"), _stat(st["std"], "n2_eq") + EPS
        self.suffix, self.gate = suffix, gate
        self.rep_projector = RepProjector(2, rep_feats, between_feats)
        symmetriser_feats = between_feats if symmetriser_feats is None else symmetriser_feats
        attention_hidden_feats = 4 * between_feats if attention_hidden_feats is None else attention_hidden_feats
        self.bond_model = SymmetrisedTransformer(between_feats, n_heads, attention_hidden_feats, n_att, 2 + int(gate),
                                                 torch.tensor([[0, 1], [1, 0]], dtype=torch.int32), layer_norm, dropout,
                                                 dense_layers, symmetriser_feats, positional_encoding=False)
        self.to_k = ToPositive(k_mean, k_std, 0, learnable_statistics)
        self.to_eq = ToPositive(eq_mean, eq_std, learnable_statistics=learnable_statistics)

    def _head_args(self, T):
        a = HeadOutArgs(kind=0, T=T, n_perm=2, n_out=2 + int(self.gate), n_per=0, gated=0)
        a.k_min, a.eq_min = self._host("kmin", self.to_k.min_)[0], self._host("emin", self.to_eq.min_)[0]
        if not self.to_k.learnable:
            a.k_mean_over_std, a.k_std = self._host("kmos", self.to_k.mean_over_std)[0], self._host("kstd", self.to_k.std)[0]
            a.eq_mean_over_std, a.eq_std = (self._host("emos", self.to_eq.mean_over_std)[0],
                                            self._host("estd", self.to_eq.std)[0])
        return a

    def _stat_params(self):
        if not self.to_k.learnable:
            return None
        return [self.to_k.mean_over_std, self.to_k.std, self.to_eq.mean_over_std, self.to_eq.std]

    def forward(self, g):
        h = g.nodes["n1"].data["h"]
        if get_pack(g).n_tuples[0] == 0:
            z = torch.zeros(0, device=h.device)
            g.nodes["n2"].data["eq" + self.suffix], g.nodes["n2"].data["k" + self.suffix] = z, z.clone()
            return g
        # the harmonic gate is computed and then discarded by the reference (interaction_parameters.py:255-264):
        # the written k is the un-gated one, which is what the output kernel produces.
        k, eq = self._write(g, h)
        g.nodes["n2"].data["eq" + self.suffix] = eq
        g.nodes["n2"].data["k" + self.suffix] = k
        return g


class WriteAngleParameters(_TupleWriter):
    level_id, level, model_attr = 1, "n3", "angle_model"

    def __init__(self, rep_feats, between_feats, suffix="", param_statistics=None, n_att=2, n_heads=8, dense_layers=2,
                 dropout=0., layer_norm=True, symmetriser_feats=None, attention_hidden_feats=None, positional_encoding=True,
                 learnable_statistics: bool = False, gate: bool = False):
        super().__init__()
        EPS = 1e-6
        st = get_default_statistics() if param_statistics is None else param_statistics
        k_mean, k_std = _stat(st["mean"], "n3_k"), _stat(st["std"], "n3_k") + EPS
        eq_std = _stat(st["std"], "n3_eq") + EPS
        self.suffix, self.gate = suffix, gate
        proj_feats = between_feats - 1 if positional_encoding else between_feats
        self.rep_projector = RepProjector(3, rep_feats, proj_feats)
        symmetriser_feats = between_feats if symmetriser_feats is None else symmetriser_feats
        attention_hidden_feats = 4 * between_feats if attention_hidden_feats is None else attention_hidden_feats
        self.angle_model = SymmetrisedTransformer(proj_feats, n_heads, attention_hidden_feats, n_att, 2 + int(gate),
                                                  torch.tensor([[0, 1, 2], [2, 1, 0]], dtype=torch.int32), layer_norm,
                                                  dropout, dense_layers, symmetriser_feats,
                                                  positional_encoding=copy.deepcopy(positional_encoding))
        self.to_k = ToPositive(k_mean, k_std, 0, learnable_statistics)
        self.to_eq = ToRange(torch.pi, eq_std, learnable_statistics)

    def _head_args(self, T):
        a = HeadOutArgs(kind=1, T=T, n_perm=2, n_out=2 + int(self.gate), n_per=0, gated=0)
        a.k_min, a.eq_max = self._host("kmin", self.to_k.min_)[0], self._host("emax", self.to_eq.max)[0]
        if not self.to_k.learnable:
            a.k_mean_over_std, a.k_std = self._host("kmos", self.to_k.mean_over_std)[0], self._host("kstd", self.to_k.std)[0]
            a.eq_std_over_max = self._host("esom", self.to_eq.std_over_max)[0]
        return a

    def _stat_params(self):
        if not self.to_k.learnable:
            return None
        return [self.to_k.mean_over_std, self.to_k.std, self.to_eq.std_over_max, None]

    def forward(self, g):
        if "n3" not in g.ntypes:
            return g
        h = g.nodes["n1"].data["h"]
        if get_pack(g).n_tuples[1] == 0:
            z = torch.zeros(0, device=h.device)
            g.nodes["n3"].data["eq" + self.suffix], g.nodes["n3"].data["k" + self.suffix] = z, z.clone()
            return g
        k, eq = self._write(g, h)
        g.nodes["n3"].data["eq" + self.suffix] = eq
        g.nodes["n3"].data["k" + self.suffix] = k
        return g


class WriteTorsionParameters(_TupleWriter):
    model_attr = "torsion_model"

    def __init__(self, rep_feats, between_feats, suffix="", n_periodicity=None, improper=False, n_att=2, n_heads=8,
                 dense_layers=2, dropout=0., layer_norm=True, symmetriser_feats=None, attention_hidden_feats=None,
                 param_statistics=None, positional_encoding=True, gated: bool = False, learnable_statistics: bool = False,
                 wrong_symmetry: bool = False, cutoff=1e-4):
        super().__init__()
        self.wrong_symmetry = wrong_symmetry
        self.learnable = bool(learnable_statistics)
        EPS = 1e-1 if gated else 1e-2
        st = get_default_statistics() if param_statistics is None else param_statistics
        self.gated, self.improper, self.suffix = gated, improper, suffix
        self.level_id, self.level = (3, "n4_improper") if improper else (2, "n4")
        if n_periodicity is None:
            n_periodicity = N_PERIODICITY_IMPROPER if improper else N_PERIODICITY_PROPER
        self.register_buffer("n_periodicity", torch.tensor(n_periodicity).long())
        self._n_per = int(n_periodicity)
        if not improper:
            k_mean, k_std = st["mean"]["n4_k"], st["std"]["n4_k"] + EPS
        elif "n4_improper_k" not in st["mean"]:
            k_mean, k_std = torch.zeros(n_periodicity), torch.ones(n_periodicity)
        else:
            k_mean, k_std = st["mean"]["n4_improper_k"], st["std"]["n4_improper_k"] + EPS
            if len(k_mean) < n_periodicity or len(k_std) < n_periodicity:
                raise ValueError(f"n_periodicity is {n_periodicity} but the param_statistics contains {len(k_mean)} "
                                 f"values for the improper torsion parameters.")
        if self.learnable:      # interaction_parameters.py:465-467
            self.k_mean = nn.Parameter(k_mean[:n_periodicity].unsqueeze(0).clone().float())
            self.k_std = nn.Parameter(k_std[:n_periodicity].unsqueeze(0).clone().float())
        else:
            self.register_buffer("k_mean", k_mean[:n_periodicity].unsqueeze(0).clone())
            self.register_buffer("k_std", k_std[:n_periodicity].unsqueeze(0).clone())
        proj_feats = between_feats - 1 if positional_encoding else between_feats
        self.rep_projector = RepProjector(4, rep_feats, proj_feats, improper=improper)
        symmetriser_feats = between_feats if symmetriser_feats is None else symmetriser_feats
        attention_hidden_feats = 4 * between_feats if attention_hidden_feats is None else attention_hidden_feats
        perms = torch.tensor([[0, 1, 2, 3], [3, 1, 2, 0]] if improper else [[0, 1, 2, 3], [3, 2, 1, 0]], dtype=torch.int32)
        if improper and wrong_symmetry:
            # ablation (interaction_parameters.py:499-505): every permutation that keeps the central atom (index 2) fixed
            perms = torch.tensor([[0, 1, 2, 3], [3, 1, 2, 0], [1, 3, 2, 0], [0, 3, 2, 1], [3, 0, 2, 1], [1, 0, 2, 3]], dtype=torch.int32)
            positional_encoding = torch.tensor([[0], [0], [1], [0]], dtype=torch.float32)
        self._n_perm = int(perms.shape[0])
        n_out = 2 * n_periodicity if gated else n_periodicity
        self.torsion_model = SymmetrisedTransformer(proj_feats, n_heads, attention_hidden_feats, n_att, n_out, perms,
                                                    layer_norm, dropout, dense_layers, symmetriser_feats,
                                                    positional_encoding=copy.deepcopy(positional_encoding))
        self.cutoff = HardCutoff(cutoff) if cutoff > 0 else None

    def _head_args(self, T):
        n = self._n_per
        a = HeadOutArgs(kind=2, T=T, n_perm=self._n_perm, n_out=(2 * n if self.gated else n), n_per=n, gated=int(self.gated))
        if not self.learnable:
            std, mean = self._host("kstd", self.k_std), self._host("kmean", self.k_mean)
            for i in range(n):
                a.tk_std[i], a.tk_mean[i] = std[i], mean[i]
        a.cutoff = float(self.cutoff.cutoff) if self.cutoff is not None else 0.0
        return a

    def _stat_params(self):
        return [self.k_std, self.k_mean, None, None] if self.learnable else None

    def forward(self, g):
        if self.level not in g.ntypes:
            return g
        h = g.nodes["n1"].data["h"]
        if get_pack(g).n_tuples[self.level_id] == 0:
            g.nodes[self.level].data["k" + self.suffix] = torch.zeros((0, self._n_per), dtype=h.dtype, device=h.device)
            return g
        (k,) = self._write(g, h)
        g.nodes[self.level].data["k" + self.suffix] = k
        return g


# CUDA stream priorities of the (proper, angle, bond, improper) writer streams: OFF.  They were meant to finish the small
# writers first so that their gradient buckets are exchanged early under data parallelism, but serialising the writers
# costs more than the earlier exchange gains -- measured on B200: 1 GPU 9.8 -> 10.5 ms per step, 2 GPUs 8.74 -> 9.19 ms
# (profiles/r2_summary.md).  GRAPPA_B200_WRITER_PRIO=1 turns them on (tuning aid).
_WRITER_PRIORITIES = (0, -1, -2, -3) if os.environ.get("GRAPPA_B200_WRITER_PRIO") == "1" else None


def set_writer_priorities(on: bool):
    global _WRITER_PRIORITIES
    _WRITER_PRIORITIES = (0, -1, -2, -3) if on else None


class WriteParameters(nn.Module):
    def __init__(self, graph_node_features=256, parameter_dropout=0, layer_norm=True, positional_encoding=True,
                 bond_transformer_depth=2, bond_n_heads=8, bond_transformer_width=512, bond_symmetriser_depth=2,
                 bond_symmetriser_width=256, angle_transformer_depth=2, angle_n_heads=8, angle_transformer_width=512,
                 angle_symmetriser_depth=2, angle_symmetriser_width=256, proper_transformer_depth=2, proper_n_heads=8,
                 proper_transformer_width=512, proper_symmetriser_depth=2, proper_symmetriser_width=256,
                 improper_transformer_depth=2, improper_n_heads=8, improper_transformer_width=512,
                 improper_symmetriser_depth=2, improper_symmetriser_width=256, n_periodicity_proper=6,
                 n_periodicity_improper=3, gated_torsion: bool = False, suffix="", wrong_symmetry=False,
                 learnable_statistics: bool = False, param_statistics: dict = None, torsion_cutoff=1.e-4,
                 harmonic_gate: bool = False):
        super().__init__()
        st = get_default_statistics() if param_statistics is None else param_statistics
        for m in ("mean", "std"):   # NaN statistics fall back to the defaults (interaction_parameters.py:41-45)
            for k, v in st[m].items():
                if torch.isnan(torch.as_tensor(v)).any():
                    st[m][k] = get_default_statistics()[m][k]
        self.bond_writer = WriteBondParameters(graph_node_features, bond_transformer_width, suffix, st, bond_transformer_depth,
                                               bond_n_heads, bond_symmetriser_depth, parameter_dropout, layer_norm,
                                               bond_symmetriser_width, bond_transformer_width,
                                               learnable_statistics=learnable_statistics, gate=harmonic_gate)
        self.angle_writer = WriteAngleParameters(graph_node_features, angle_transformer_width, suffix, st,
                                                 angle_transformer_depth, angle_n_heads, angle_symmetriser_depth,
                                                 parameter_dropout, layer_norm, angle_symmetriser_width, angle_transformer_width,
                                                 positional_encoding, learnable_statistics, harmonic_gate)
        self.proper_writer = WriteTorsionParameters(graph_node_features, proper_transformer_width, suffix, n_periodicity_proper,
                                                    False, proper_transformer_depth, proper_n_heads, proper_symmetriser_depth,
                                                    parameter_dropout, layer_norm, proper_symmetriser_width,
                                                    proper_transformer_width, st, positional_encoding, gated_torsion,
                                                    learnable_statistics, cutoff=torsion_cutoff)
        self.improper_writer = WriteTorsionParameters(graph_node_features, improper_transformer_width, suffix,
                                                      n_periodicity_improper, True, improper_transformer_depth,
                                                      improper_n_heads, improper_symmetriser_depth, parameter_dropout,
                                                      layer_norm, improper_symmetriser_width, improper_transformer_width, st,
                                                      positional_encoding, gated_torsion, learnable_statistics,
                                                      wrong_symmetry, torsion_cutoff)

    def forward(self, g):
        writers = (self.proper_writer, self.angle_writer, self.bond_writer, self.improper_writer)   # largest first
        h = g.nodes["n1"].data.get("h")
        if h is None or not h.is_cuda or not T_.concurrency():
            for w in (self.bond_writer, self.angle_writer, self.proper_writer, self.improper_writer):
                g = w(g)
            return g
        # The four writers are independent (the reference runs them one after the other, see its note at
        # interaction_parameters.py:126-128): each one gets its own stream, so their kernels -- most of which cannot
        # fill 148 SMs alone -- overlap; autograd replays each writer's backward on the same stream.
        main = torch.cuda.current_stream()
        start = torch.cuda.Event()
        start.record(main)
        done = []
        # Optional stream priorities, smallest writer first (off by default, see _WRITER_PRIORITIES above)
        prios = _WRITER_PRIORITIES if torch.is_grad_enabled() else None
        for w, st in zip(writers, T_.helper_streams(len(writers), "writer", prios)):
            st.wait_event(start)
            with torch.cuda.stream(st), ops.gemm_sm_limit(ops.CONCURRENT_GEMM_SMS):
                g = w(g)
                ev = torch.cuda.Event()
                ev.record(st)
            done.append(ev)
        for ev in done:
            main.wait_event(ev)
        return g


class GrappaModel(nn.Module):
    """GNN feature extraction followed by the four parameter writers (reference models/grappa.py:7-132)."""

    def __init__(self, graph_node_features: int = 512, in_feats: int = None,
                 in_feat_name: Union[str, List[str]] = ["atomic_number", "ring_encoding", "partial_charge"],
                 in_feat_dims: Dict[str, int] = {}, gnn_width: int = None, gnn_attentional_layers: int = 3,
                 gnn_convolutions: int = 3, gnn_attention_heads: int = 8, gnn_dropout_attention: float = 0.,
                 gnn_dropout_initial: float = 0., gnn_dropout_conv: float = 0., gnn_dropout_final: float = 0.,
                 parameter_dropout: float = 0., bond_transformer_depth=2, bond_n_heads=8, bond_transformer_width=512,
                 bond_symmetriser_depth=2, bond_symmetriser_width=256, angle_transformer_depth=2, angle_n_heads=8,
                 angle_transformer_width=512, angle_symmetriser_depth=2, angle_symmetriser_width=256,
                 proper_transformer_depth=2, proper_n_heads=8, proper_transformer_width=512, proper_symmetriser_depth=2,
                 proper_symmetriser_width=256, improper_transformer_depth=2, improper_n_heads=8,
                 improper_transformer_width=512, improper_symmetriser_depth=2, improper_symmetriser_width=256,
                 n_periodicity_proper=6, n_periodicity_improper=3, gated_torsion: bool = False, wrong_symmetry=False,
                 positional_encoding=True, layer_norm=True, self_interaction=True, learnable_statistics: bool = False,
                 param_statistics: dict = None, torsion_cutoff=1.e-4, harmonic_gate: bool = False):
        super().__init__()
        self.gnn = GrappaGNN(out_feats=graph_node_features, in_feats=in_feats, node_feats=gnn_width, n_conv=gnn_convolutions,
                             n_att=gnn_attentional_layers, n_heads=gnn_attention_heads, in_feat_name=in_feat_name,
                             in_feat_dims=in_feat_dims, conv_dropout=gnn_dropout_conv, attention_dropout=gnn_dropout_attention,
                             final_dropout=gnn_dropout_final, initial_dropout=gnn_dropout_initial,
                             self_interaction=self_interaction, layer_norm=layer_norm)
        self.parameter_writer = WriteParameters(
            graph_node_features=graph_node_features, parameter_dropout=parameter_dropout, layer_norm=layer_norm,
            positional_encoding=positional_encoding, bond_transformer_depth=bond_transformer_depth, bond_n_heads=bond_n_heads,
            bond_transformer_width=bond_transformer_width, bond_symmetriser_depth=bond_symmetriser_depth,
            bond_symmetriser_width=bond_symmetriser_width, angle_transformer_depth=angle_transformer_depth,
            angle_n_heads=angle_n_heads, angle_transformer_width=angle_transformer_width,
            angle_symmetriser_depth=angle_symmetriser_depth, angle_symmetriser_width=angle_symmetriser_width,
            proper_transformer_depth=proper_transformer_depth, proper_n_heads=proper_n_heads,
            proper_transformer_width=proper_transformer_width, proper_symmetriser_depth=proper_symmetriser_depth,
            proper_symmetriser_width=proper_symmetriser_width, improper_transformer_depth=improper_transformer_depth,
            improper_n_heads=improper_n_heads, improper_transformer_width=improper_transformer_width,
            improper_symmetriser_depth=improper_symmetriser_depth, improper_symmetriser_width=improper_symmetriser_width,
            n_periodicity_proper=n_periodicity_proper, n_periodicity_improper=n_periodicity_improper,
            wrong_symmetry=wrong_symmetry, param_statistics=param_statistics, gated_torsion=gated_torsion,
            learnable_statistics=learnable_statistics, torsion_cutoff=torsion_cutoff, harmonic_gate=harmonic_gate)
        # + 3 to reach dihedrals and ring membership (reference models/grappa.py:108-109)
        self.field_of_view = gnn_attentional_layers + gnn_convolutions + 3

    def forward(self, g):
        get_pack(g)   # validates idx ranges / dtypes on the host (replaces the device syncs at grappa.py:122-128)
        g = self.gnn(g)
        g = self.parameter_writer(g)
        return g


def get_default_model_config():
    """reference models/deploy.py:18-64."""
    args = {"graph_node_features": 256, "in_feats": None,
            "in_feat_name": ["atomic_number", "partial_charge", "ring_encoding", "degree", "charge_model"],
            "in_feat_dims": {}, "gnn_width": 512, "gnn_attentional_layers": 7, "gnn_convolutions": 0,
            "gnn_attention_heads": 16, "gnn_dropout_attention": 0.3, "gnn_dropout_initial": 0.0, "gnn_dropout_conv": 0.1,
            "gnn_dropout_final": 0.1, "parameter_dropout": 0.5, "n_periodicity_proper": 6, "n_periodicity_improper": 3,
            "gated_torsion": True, "wrong_symmetry": False, "positional_encoding": True, "layer_norm": True,
            "self_interaction": True, "learnable_statistics": False, "torsion_cutoff": 1e-4}
    for w in ("bond", "angle", "proper", "improper"):
        args.update({f"{w}_transformer_depth": 3, f"{w}_n_heads": 8, f"{w}_transformer_width": 512,
                     f"{w}_symmetriser_depth": 3, f"{w}_symmetriser_width": 256})
    return args


def grappa_1_2_model_config():
    """experiments/train-grappa-1.2/grappa_config.yaml:64-111."""
    cfg = get_default_model_config()
    cfg["n_periodicity_proper"] = 3
    return cfg


def model_from_config(model_config: Dict, param_statistics: Dict = None):
    """reference models/deploy.py:8-16."""
    return GrappaModel(param_statistics=param_statistics, **model_config)
