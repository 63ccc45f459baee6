"""A small reverse-mode tape over the CUDA operators (replaces torch autograd inside the modules).

The reference relies on torch autograd over ~150 ATen/DGL ops per forward; here each fused stage
(GNN, one writer) records its handful of kernel-level ops on a `Tape` and the stage's
torch.autograd.Function replays the tape backwards with hand-written backward kernels:

    linear      dgrad / wgrad GEMMs with fused residual-gradient add and ELU'-mask epilogues
    layernorm   dx kernel + deterministic column reductions for gamma / beta
    edge / tuple attention, tuple gather, permuted concat, output maps: dedicated backward kernels

Gradient accumulation for values with several consumers is fused into the producing GEMM epilogue
where possible (`residual=` existing gradient) instead of separate add kernels.
"""
from __future__ import annotations

import contextlib
import os
from typing import Callable, Dict, List, Optional

import torch

from . import ops

# Weight / bias gradient kernels do not feed the backward chain, so they are issued on a per-stream helper
# stream and overlap with the dgrad kernels (inside a captured CUDA graph this becomes a parallel branch).
_CONCURRENT = True
_FUSE_COLSUM = os.environ.get("GRAPPA_B200_FUSE_COLSUM", "1") != "0"   # bias-gradient column sums inside the dgrad epilogue
_side_streams: Dict[tuple, "torch.cuda.Stream"] = {}


def set_concurrency(on: bool):
    """Enable / disable multi-stream overlap (weight-gradient branch, concurrent writers)."""
    global _CONCURRENT
    _CONCURRENT = bool(on)


def concurrency() -> bool:
    return _CONCURRENT


def helper_streams(n: int, tag: str, priorities=None):
    """`n` cached helper streams tied to the current stream (one set per (device, current stream, tag)).  `priorities`:
    optional CUDA stream priority per stream (0 = default, negative = scheduled first)."""
    cur = torch.cuda.current_stream()
    out = []
    for i in range(n):
        prio = 0 if priorities is None else int(priorities[i])
        key = (cur.device, cur.cuda_stream, tag, i, prio)
        st = _side_streams.get(key)
        if st is None:
            st = torch.cuda.Stream(device=cur.device, priority=prio)
            _side_streams[key] = st
        out.append(st)
    return out


class Var:
    """A value on the tape with its (lazily created) gradient."""
    __slots__ = ("v", "g", "needs", "elu_fusable", "g_is_pre", "g_colsum")

    def __init__(self, v: torch.Tensor, needs: bool = True):
        self.v = v
        self.g: Optional[torch.Tensor] = None
        self.needs = needs
        self.elu_fusable = False   # v == ELU(pre) exactly and the single consumer is a linear layer
        self.g_is_pre = False      # g already holds d/d(pre-activation)
        self.g_colsum = None       # per-32-row column sums of g emitted by the GEMM that produced it (bias-gradient partials)


class Tape:
    def __init__(self, record: bool, train: bool, seed: int):
        self.record = record
        self.train = train
        self.seed = int(seed)
        self._n = 0
        self.fns: List[Callable[[], None]] = []
        self.pgrads: Dict[int, torch.Tensor] = {}   # id(param tensor) -> grad
        self.sinks: Dict[int, torch.Tensor] = {}    # id(param tensor) -> preallocated gradient buffer (flat-buffer view)
        self._sink_written = set()
        self._side = None          # helper stream of the weight-gradient branch (created on first use)
        self._side_dirty = False
        self._keep: List[torch.Tensor] = []   # tensors the helper stream still reads (released at the next join)
        self.colsums = ops.ColumnSums()       # deferred bias / LayerNorm parameter-gradient reductions
        self._wgrads: List[tuple] = []        # deferred weight-gradient GEMMs (a, b, kwargs): issued in groups of <= 4

    def next_seed(self) -> int:
        self._n += 1
        return (self.seed * 0x9E3779B1 + self._n * 0x85EBCA77) & 0xFFFFFFFFFFFFFFFF

    def push(self, fn):
        if self.record:
            self.fns.append(fn)

    def backward(self):
        for fn in reversed(self.fns):
            fn()
        self.fns = []
        self.join_side()

    @contextlib.contextmanager
    def side_branch(self, *tensors):
        """Run the enclosed kernel launches on the helper stream, ordered after everything enqueued so far on the
        current stream.  `tensors` (allocated on the current stream) stay referenced until `join_side`."""
        if not _CONCURRENT or not tensors[0].is_cuda:
            yield
            return
        if self._side is None:
            cur = torch.cuda.current_stream()
            self._side = helper_streams(1, "wgrad", [getattr(cur, "priority", 0)])[0]   # the branch inherits its stage's priority
        ev = torch.cuda.Event()
        ev.record()
        self._side.wait_event(ev)
        self._keep.extend(tensors)
        self._side_dirty = True
        with torch.cuda.stream(self._side):
            yield

    def defer_wgrad(self, dpre: torch.Tensor, x: torch.Tensor, out: torch.Tensor, accumulate: bool, **kw):
        """Queue dW = dpre^T x for a grouped launch.  The weight gradients of consecutive layers of one stage are
        independent of the backward chain and of each other, so four of them share ONE persistent kernel whose work
        items are sized to fill the SMs once (grappa_b200_gemm_grouped): fewer, longer K slices, fewer partials, a
        quarter of the launches.  A second write to a buffer that is already queued flushes first (ordering)."""
        if any(o.data_ptr() == out.data_ptr() for _, _, o, _ in self._wgrads):
            self.flush_wgrads()
        self._wgrads.append((dpre, x, out, dict(trans_a=True, trans_b=True, out=out, accumulate=accumulate, **kw)))
        if len(self._wgrads) >= 4:
            self.flush_wgrads()

    def flush_wgrads(self):
        if not self._wgrads:
            return
        batch, self._wgrads = self._wgrads, []
        tensors = [t for d, x, _, _ in batch for t in (d, x)]
        with self.side_branch(*tensors):
            ops.gemm_grouped([(d, x, kw) for d, x, _, kw in batch])

    def flush_events(self):
        """Issue every deferred weight-gradient / column-sum kernel recorded so far and return CUDA events (current
        stream, weight-gradient stream) after which the parameter gradients written so far are final -- WITHOUT making
        the current stream wait for the weight-gradient branch (`join_side` does).  Used by the data-parallel trainer
        to exchange a stage's gradients part by part while its backward pass is still running."""
        self.flush_wgrads()
        self.colsums.flush()
        if not torch.cuda.is_available():
            return ()
        evs = []
        ev = torch.cuda.Event()
        ev.record()
        evs.append(ev)
        if self._side is not None and self._side_dirty:
            ev2 = torch.cuda.Event()
            ev2.record(self._side)
            evs.append(ev2)
        return tuple(evs)

    def join_side(self):
        """Fold the deferred column sums and make the current stream wait for the weight-gradient branch (before
        parameter gradients are consumed)."""
        self.flush_wgrads()
        self.colsums.flush()
        if self._side_dirty:
            ev = torch.cuda.Event()
            ev.record(self._side)
            torch.cuda.current_stream().wait_event(ev)
            self._side_dirty = False
        self._keep = []

    def grad_target(self, p: torch.Tensor):
        """(buffer, accumulate) if the parameter's gradient is written in place by the kernels, else (None, False)."""
        s = self.sinks.get(id(p))
        if s is None:
            return None, False
        acc = id(p) in self._sink_written
        self._sink_written.add(id(p))
        return s, acc

    def deferred_target(self, p: torch.Tensor):
        """(buffer, accumulate) for a parameter gradient that a later `colsums.flush()` writes: the flat-buffer view if
        the trainer installed one, else a tensor handed to autograd (created here on first use)."""
        tgt, acc = self.grad_target(p)
        if tgt is not None:
            return tgt, acc
        k = id(p)
        if k in self.pgrads:
            return self.pgrads[k], True
        g = torch.empty_like(p)
        self.pgrads[k] = g
        return g, False

    def add_pgrad(self, p: torch.Tensor, g: torch.Tensor):
        k = id(p)
        if k in self.pgrads:
            ops.axpby(g, self.pgrads[k], 1.0, 1.0)
        else:
            self.pgrads[k] = g


def add_grad(var: Var, g: torch.Tensor):
    if not var.needs:
        return
    if var.g is None:
        var.g = g
    else:
        out = var.g.clone()
        ops.axpby(g, out, 1.0, 1.0)
        var.g = out


# --------------------------------------------------------------------------------------------------
# primitives
# --------------------------------------------------------------------------------------------------
def linear(t: Tape, x: Var, W: torch.Tensor, b: Optional[torch.Tensor], *, act: int = 0, dropout_p: float = 0.0,
           residual: Optional[Var] = None, k: Optional[int] = None, out_ld: Optional[int] = None,
           fuse_elu_into_consumer: bool = False, x_pad_is_zero: bool = False, pre_add: Optional[Var] = None) -> Var:
    """y = dropout(act(x[:, :k] W^T + b + pre_add)) + residual.  `out_ld` > N gives a padded output buffer.
    `x_pad_is_zero`: the columns of x beyond k (up to the next multiple of 4) hold zeros.  `pre_add`: a value added
    under the activation (the output of another bias-free linear: SAGEConv's fc_self + fc_neigh); its buffer is reused
    as the output, so it must have no other consumer."""
    xv = x.v
    N = W.shape[0]
    K = k if k is not None else W.shape[1]
    M = xv.shape[0]
    p = dropout_p if t.train else 0.0
    seed = t.next_seed() if p > 0.0 else 0
    need_act_out = t.record and act != 0 and (p > 0.0 or residual is not None)
    if out_ld is not None and out_ld != N:
        buf = torch.zeros((M, out_ld), device=xv.device, dtype=torch.float32)
        out = buf[:, :N]
        assert p == 0.0, "dropout on a padded output is not supported"
    elif pre_add is not None:
        assert tuple(pre_add.v.shape) == (M, N) and pre_add.v.is_contiguous()
        buf = out = pre_add.v
    else:
        buf = torch.empty((M, N), device=xv.device, dtype=torch.float32)
        out = buf
    act_out = torch.empty((M, N), device=xv.device, dtype=torch.float32) if need_act_out else None
    Wf, Kf = W, K
    K4 = (K + 3) // 4 * 4
    if (x_pad_is_zero and W.stride(0) % 4 != 0 and ops.matmul_precision() != ops.FP32 and xv.shape[1] >= K4
            and xv.stride(0) % 4 == 0 and M >= 128):
        # rows of W are not 16-byte multiples (pre_dense: 85 features): the tensor-core path needs a zero-padded copy;
        # the activation buffer already carries zero pad columns (featurize), so the product is unchanged
        Wf, Kf = ops.pad_rows(W, K4), K4
    ops.gemm(xv, Wf, bias=b, act=act, dropout_p=p, dropout_seed=seed, residual=None if residual is None else residual.v,
             out=out, k=Kf, m=M, n=N, act_out=act_out, accumulate=2 if pre_add is not None else False)
    y = Var(buf)
    if act != 0 and p == 0.0 and residual is None and fuse_elu_into_consumer:
        y.elu_fusable = True

    def bwd():
        dy_full = y.g
        if dy_full is None:
            return
        if residual is not None:
            add_grad(residual, dy_full)
        # elementwise part on the full (possibly padded, contiguous) buffers; padded columns carry zeros
        bias_partial = None
        if y.g_is_pre:
            dpre_full = dy_full
            if b is not None and y.g_colsum is not None and dy_full.shape[1] == N:
                bias_partial = y.g_colsum          # the dgrad GEMM that wrote dy also summed its columns
        elif act != 0 or p > 0.0:
            saved = act_out if need_act_out else (y.v if act != 0 else None)
            dyc = dy_full.contiguous()
            if b is not None and dyc.shape[1] == N and N % 4 == 0 and N <= 8192 and M > 0:
                # one pass: dpre and the per-CTA partial sums of the bias gradient (folded later by tape.colsums)
                dpre_full, bias_partial = ops.act_dropout_bwd_fused(dyc, saved, p, seed)
            else:
                dpre_full = ops.act_dropout_bwd(dyc, saved, p, seed)
        else:
            dpre_full = dy_full
        dpre = dpre_full[:, :N] if dpre_full.shape[1] != N else dpre_full
        if pre_add is not None:
            add_grad(pre_add, dpre_full)
        # weight / bias gradients (parallel branch; the GEMMs are queued and issued four at a time)
        tgt, acc = t.grad_target(W)
        if tgt is None:
            if id(W) in t.pgrads:       # second use of a weight inside one stage: keep the simple ordered path
                t.flush_wgrads()
                with t.side_branch(dpre_full, xv):
                    t.add_pgrad(W, ops.gemm(dpre, xv, trans_a=True, trans_b=True, m=N, n=K, k=M, family="wgrad"))
            else:
                gw = torch.empty((N, K), device=xv.device, dtype=torch.float32)
                t.pgrads[id(W)] = gw
                t.defer_wgrad(dpre, xv, gw, False, m=N, n=K, k=M, family="wgrad")
        else:
            t.defer_wgrad(dpre, xv, tgt, acc, m=N, n=K, k=M, family="wgrad")
        with t.side_branch(dpre_full, xv) if (b is not None and bias_partial is None) else contextlib.nullcontext():
            if b is not None and bias_partial is None:
                tgt, acc = t.grad_target(b)
                if tgt is None:
                    db, _ = ops.col_reduce(dpre, cols=N)
                    t.add_pgrad(b, db)
                else:
                    ops.col_reduce(dpre, out_sum=tgt, cols=N, accumulate=acc)
        if bias_partial is not None:
            tgt, acc = t.deferred_target(b)
            t.colsums.add(bias_partial, bias_partial.shape[0], N, N, tgt, acc)
        if x.needs:
            fuse = x.elu_fusable and x.g is None
            if K == xv.shape[1]:
                if fuse and _FUSE_COLSUM:
                    # dx is the pre-activation gradient of the producing layer: let the epilogue also emit its column
                    # sums (that layer's bias gradient) instead of re-reading dx in a separate reduction
                    dx, x.g_colsum = ops.gemm_with_colsum(dpre, W, trans_b=True, m=M, n=K, k=N, residual=x.g, mul_elu_out=xv,
                                                          family="dgrad")
                else:
                    dx = ops.gemm(dpre, W, trans_b=True, m=M, n=K, k=N, residual=x.g, mul_elu_out=xv if fuse else None,
                                  family="dgrad")
                x.g = dx
                x.g_is_pre = fuse
            else:   # consumer read only the first K columns of a wider buffer
                dx = torch.zeros_like(xv)
                ops.gemm(dpre, W, trans_b=True, m=M, n=K, k=N, out=dx[:, :K], family="dgrad")
                add_grad(x, dx)

    t.push(bwd)
    return y


def layernorm(t: Tape, x: Var, gamma: torch.Tensor, beta: torch.Tensor) -> Var:
    yv, mean, rstd = ops.layernorm_fwd(x.v, gamma, beta)
    y = Var(yv)

    def bwd():
        dy = y.g
        if dy is None:
            return
        cols = x.v.shape[1]
        if x.needs and cols <= 512 and cols % 4 == 0 and x.v.shape[0] > 0:
            # one pass over dy / x: dx and per-CTA partials of the gamma / beta gradients
            dx, partial = ops.layernorm_bwd_fused(dy.contiguous(), x.v, mean, rstd, gamma)
            tg, acc_g = t.deferred_target(gamma)
            tb, acc_b = t.deferred_target(beta)
            n = partial.shape[0]
            t.colsums.add(partial, n, 2 * cols, cols, tb, acc_b)
            t.colsums.add(partial[0, 1], n, 2 * cols, cols, tg, acc_g)
            add_grad(x, dx)
            return
        with t.side_branch(dy, x.v, mean, rstd):
            tg, acc = t.grad_target(gamma)
            tb, _ = t.grad_target(beta)
            if tg is None or tb is None:
                dbeta, dgamma = ops.col_reduce(dy, x=x.v, mean=mean, rstd=rstd)
                t.add_pgrad(gamma, dgamma)
                t.add_pgrad(beta, dbeta)
            else:
                ops.col_reduce(dy, out_sum=tb, x=x.v, mean=mean, rstd=rstd, out_xhat=tg, accumulate=acc)
        if x.needs:
            add_grad(x, ops.layernorm_bwd(dy, x.v, mean, rstd, gamma))

    t.push(bwd)
    return y


def edge_attention(t: Tape, ft: Var, pack, heads: int) -> Var:
    out, alpha = ops.edge_attention_fwd(ft.v, pack, heads)
    y = Var(out)

    def bwd():
        if y.g is None:
            return
        add_grad(ft, ops.edge_attention_bwd(ft.v, alpha, y.g, pack, heads))

    t.push(bwd)
    return y


def neighbor_mean(t: Tape, x: Var, pack) -> Var:
    """SAGEConv('mean') aggregation: y[v] = mean over bonded neighbours of x."""
    y = Var(ops.neighbor_mean(x.v, pack))

    def bwd():
        if y.g is None:
            return
        add_grad(x, ops.neighbor_mean(y.g.contiguous(), pack, transpose=True))

    t.push(bwd)
    return y


def tuple_attention(t: Tape, qkv: Var, T: int, L: int, heads: int) -> Var:
    y = Var(ops.tuple_attention_fwd(qkv.v, T, L, heads))

    def bwd():
        if y.g is None:
            return
        add_grad(qkv, ops.tuple_attention_bwd(qkv.v, y.g, T, L, heads))

    t.push(bwd)
    return y


def tuple_gather(t: Tape, p: Var, pack, level: int, pe: Optional[torch.Tensor], F: int, E: int) -> Var:
    L = (2, 3, 4, 4)[level]
    T = pack.n_tuples[level]
    y = Var(ops.tuple_gather_fwd(p.v, pack[f"idx{level}"], pe, T, L, F, E))

    def bwd():
        if y.g is None or not p.needs:
            return
        add_grad(p, ops.tuple_gather_bwd(y.g, pack[f"inv_ptr{level}"], pack[f"inv_ent{level}"], pack.n_atoms,
                                         p.v.shape[1], T, L, F, E))

    t.push(bwd)
    return y


def perm_concat(t: Tape, x: Var, perms, T: int, L: int, E: int) -> Var:
    y = Var(ops.perm_concat_fwd(x.v, perms, T, L, E))

    def bwd():
        if y.g is None:
            return
        add_grad(x, ops.perm_concat_bwd(y.g, perms, T, L, E))

    t.push(bwd)
    return y
