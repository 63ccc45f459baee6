"""Tensor-level wrappers over the C ABI: torch owns memory and streams, the kernels do the math.

Every function takes contiguous fp32 CUDA tensors, allocates its outputs with torch (caching
allocator, graph-capture safe) and enqueues on torch's current stream.  No function here computes
anything with torch ops.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib_ops import COLSUM_MAX, ColsumBatch, FeaturizeArgs, GemmArgs, HeadOutArgs, HeadStatGrads, Perms

FP32, TF32, AUTO, BF16X3 = 0, 1, 2, 3
_PREC_CODES = {"fp32": FP32, "tf32": AUTO, "bf16x3": BF16X3}
_PREC_NAMES = {FP32: "fp32", AUTO: "tf32", TF32: "tf32", BF16X3: "bf16x3"}
FAMILIES = ("fwd", "dgrad", "wgrad")
BENCH_PRECISION = "bf16x3"      # what bench.py measures and the full-size parity tests check (bench.py --precision overrides)
# GEMM precision per family (forward Linear / input gradient / weight gradient), optionally overridden per stage
# ("gnn.fwd", "writer.dgrad", ...): the error budget of profiles/r2_error_budget.md is measured by switching one entry.
_policy = {f: FP32 for f in FAMILIES}
_stage_tag = None


def set_matmul_precision(mode):
    """GEMM arithmetic of the Linear layers.

    'fp32'    FFMA on the CUDA cores (1e-5 parity path)
    'tf32'    tcgen05 kind::tf32 on the raw fp32 operands (10-bit mantissa; misses the 1e-3 contract on gated torsion k)
    'bf16x3'  tcgen05 kind::f16: every fp32 operand is split in-kernel into bf16 hi + bf16 lo and the product is
              hi*hi + lo*hi + hi*lo with fp32 accumulation in TMEM (16-bit mantissa: fp32-class parity on tensor cores)
    or a policy string / dict per family and stage: 'fwd=bf16x3,dgrad=tf32,wgrad=tf32', 'tf32,gnn.fwd=fp32', ...
    """
    global _policy
    pol = {}
    if isinstance(mode, dict):
        items = list(mode.items())
    else:
        items = []
        for part in str(mode).split(","):
            part = part.strip()
            if not part:
                continue
            if "=" in part:
                k, v = part.split("=")
                items.append((k.strip(), v.strip()))
            else:
                items.extend((f, part) for f in FAMILIES)
    for k, v in items:
        if v not in _PREC_CODES:
            raise ValueError("matmul precision must be 'fp32', 'tf32' or 'bf16x3'")
        if k.split(".")[-1] not in FAMILIES:
            raise ValueError(f"unknown GEMM family {k!r} (families: {FAMILIES}, optionally prefixed by 'gnn.' / 'writer.')")
        pol[k] = _PREC_CODES[v]
    for f in FAMILIES:
        pol.setdefault(f, FP32)
    _policy = pol


def get_matmul_precision() -> str:
    vals = {_PREC_NAMES[_policy[f]] for f in FAMILIES}
    if len(vals) == 1 and len(_policy) == len(FAMILIES):
        return vals.pop()
    return ",".join(f"{k}={_PREC_NAMES[v]}" for k, v in sorted(_policy.items()))


class stage_tag:
    """Context manager naming the stage ('gnn' / 'writer') whose GEMMs are being issued (per-stage precision overrides)."""

    def __init__(self, tag):
        self.tag = tag

    def __enter__(self):
        global _stage_tag
        self.prev, _stage_tag = _stage_tag, self.tag
        return self

    def __exit__(self, *exc):
        global _stage_tag
        _stage_tag = self.prev


def matmul_precision(family: str = "fwd") -> int:
    if _stage_tag is not None:
        v = _policy.get(f"{_stage_tag}.{family}")
        if v is not None:
            return v
    return _policy[family]


_rng_offset = None     # device uint64 counter mixed into every dropout seed (None = static seeds)


def set_rng_offset(t):
    """Register a 1-element device int64 tensor whose value is added to every dropout seed at RUN time.
    A captured CUDA graph then draws fresh masks on each replay once the counter is advanced (`tick`)."""
    global _rng_offset
    _rng_offset = t


def _rng_ptr() -> int:
    return 0 if _rng_offset is None else _rng_offset.data_ptr()


_gemm_profile = None   # list collecting (flops, start_event, stop_event, kernel) when profiling is on


# SMs a tensor-core GEMM may occupy (0 = all).  Sections that run several streams side by side set it to CONCURRENT_GEMM_SMS:
# a persistent GEMM on all 148 SMs makes every kernel of the other streams wait for a free SM (148 / 132 / 116 / 100 / 88
# SMs -> 6.02 / 5.97 / 5.92 / 6.10 / 6.08 ms per training step); single-stream code (GNN forward, inference) uses all SMs.
CONCURRENT_GEMM_SMS = 116
_gemm_sm_limit = 0


class gemm_sm_limit:
    """Context manager: tensor-core GEMMs issued inside use at most `n` SMs (0 = all)."""

    def __init__(self, n: int):
        self.n = int(n)

    def __enter__(self):
        global _gemm_sm_limit
        self.prev, _gemm_sm_limit = _gemm_sm_limit, self.n
        return self

    def __exit__(self, *exc):
        global _gemm_sm_limit
        _gemm_sm_limit = self.prev


_gemm_record = None    # list collecting every GEMM call (argument structs + the tensors they point to) of a step


def set_gemm_recorder(sink):
    """bench.py: pass a list to record every GEMM call (the exact gb_gemm_args, tensors kept alive); None = off.
    `replay_gemms(sink)` re-issues them back to back, so the GEMM family can be timed without anything in between."""
    global _gemm_record
    _gemm_record = sink


def replay_gemms(records) -> float:
    """Issue the recorded GEMM calls on the current stream; returns the FLOPs issued."""
    lib = _lib.lib()
    flops = 0.0
    for kind, arr, n, _keep in records:
        if kind == "single":
            _lib.check(lib.grappa_b200_gemm(C.byref(arr), _s()), "gemm")
            flops += 2.0 * arr.M * arr.N * arr.K
        else:
            _lib.check(lib.grappa_b200_gemm_grouped(arr, n, _s()), "gemm_grouped")
            flops += sum(2.0 * arr[i].M * arr[i].N * arr[i].K for i in range(n))
    return flops


def set_gemm_profiler(sink):
    """bench.py: pass a list to time every GEMM launch with CUDA events on the launching stream; None = off."""
    global _gemm_profile
    _gemm_profile = sink


def _p(t: Optional[torch.Tensor]) -> int:
    return 0 if t is None or t.numel() == 0 else t.data_ptr()


def _s() -> int:
    return torch.cuda.current_stream().cuda_stream


_ws_cache = {}


def workspace(nbytes: int, device, tag: str = "ws", zero: bool = False) -> torch.Tensor:
    """A reusable scratch buffer per (device, stream, tag); grows monotonically.  `zero`: zero-filled when created
    (kernels that keep self-resetting counters in it)."""
    key = (device, torch.cuda.current_stream().cuda_stream, tag)
    t = _ws_cache.get(key)
    if t is None or t.numel() * 4 < nbytes:
        n = max((nbytes + 3) // 4, 1 << 20)
        t = (torch.zeros if zero else torch.empty)(n, dtype=torch.float32, device=device)
        _ws_cache[key] = t
    return t


def drop_workspaces():
    """Forget cached scratch buffers (they are re-created on demand)."""
    _ws_cache.clear()


def _gemm_args(a: torch.Tensor, b: torch.Tensor, *, trans_a=False, trans_b=False, bias=None, act=0, mul_elu_out=None,
               dropout_p=0.0, dropout_seed=0, residual=None, out=None, accumulate=False, m=None, n=None, k=None,
               precision=None, act_out=None, colsum=None, family="fwd"):
    """Fill a gb_gemm_args for C[M,N] = opA(a) opB(b)^T (+ fused epilogue); returns (args, out)."""
    _lib.require_cuda(a, b)
    assert a.dim() == 2 and b.dim() == 2 and a.stride(1) == 1 and b.stride(1) == 1
    M = m if m is not None else (a.shape[1] if trans_a else a.shape[0])
    K = k if k is not None else (a.shape[0] if trans_a else a.shape[1])
    N = n if n is not None else (b.shape[1] if trans_b else b.shape[0])
    if out is None:
        out = torch.empty((M, N), device=a.device, dtype=torch.float32)
    g = GemmArgs()
    g.A, g.B, g.C = a.data_ptr(), b.data_ptr(), out.data_ptr()
    g.M, g.N, g.K = M, N, K
    g.lda, g.ldb, g.ldc = a.stride(0), b.stride(0), out.stride(0)
    g.trans_a, g.trans_b = int(trans_a), int(trans_b)
    g.bias = _p(bias)
    g.act = act
    g.mul_elu_out = _p(mul_elu_out)
    g.ldm = mul_elu_out.stride(0) if mul_elu_out is not None else 0
    g.dropout_p = float(dropout_p)
    g.dropout_seed = int(dropout_seed) & 0xFFFFFFFFFFFFFFFF
    g.dropout_offset = _rng_ptr() if dropout_p > 0.0 else 0
    g.residual = _p(residual)
    g.ldr = residual.stride(0) if residual is not None else 0
    g.accumulate = int(accumulate)          # 0 overwrite, 1 C += result, 2 old C added before bias / activation
    g.act_out = _p(act_out)
    g.ldact = act_out.stride(0) if act_out is not None else 0
    g.precision = matmul_precision(family) if precision is None else precision
    g.max_sms = _gemm_sm_limit
    if colsum is not None:
        g.colsum, g.ld_colsum = colsum.data_ptr(), colsum.stride(0)
    if (trans_a and M * N <= (1 << 22) and K >= 1024) or (M * N <= (1 << 20) and K >= 2048):
        # weight gradients (tiny output, very long K) and the 2048-deep feed-forward GEMMs of the GNN (13 row tiles):
        # split-K slices fill the SMs, a reduce kernel applies the epilogue
        ws = workspace(64 << 20, a.device, "splitk")
        g.workspace, g.workspace_bytes = ws.data_ptr(), ws.numel() * 4
    return g, out


def gemm(a: torch.Tensor, b: torch.Tensor, **kw) -> torch.Tensor:
    """C[M,N] = opA(a) opB(b)^T (+ fused epilogue), see gb_gemm_args.  a/b may be column-sliced views
    (last-dim stride 1); leading dimensions are taken from the row strides."""
    lib = _lib.lib()
    g, out = _gemm_args(a, b, **kw)
    if _gemm_record is not None:
        _gemm_record.append(("single", g, 1, (a, b, out, kw)))
    if _gemm_profile is not None:
        # inside a stream capture the events become graph nodes that are re-recorded by every replay
        ext = torch.cuda.is_current_stream_capturing()
        e0, e1 = torch.cuda.Event(enable_timing=True, external=ext), torch.cuda.Event(enable_timing=True, external=ext)
        e0.record()
        _lib.check(lib.grappa_b200_gemm(C.byref(g), _s()), "gemm")
        e1.record()
        _gemm_profile.append((2.0 * g.M * g.N * g.K, e0, e1, (g.M, g.N, g.K, int(g.trans_a), int(g.trans_b))))
        return out
    _lib.check(lib.grappa_b200_gemm(C.byref(g), _s()), "gemm")
    return out


def gemm_with_colsum(a: torch.Tensor, b: torch.Tensor, **kw):
    """`gemm` whose epilogue also emits per-32-row column sums of the stored result: returns (out, partial[ceil(M/32), N])
    or (out, None) when this call cannot run on the fused tensor-core path (the caller then reduces separately)."""
    lib = _lib.lib()
    g, out = _gemm_args(a, b, **kw)
    partial = torch.empty(((g.M + 31) // 32, g.N), device=a.device, dtype=torch.float32)
    g.colsum, g.ld_colsum = partial.data_ptr(), partial.stride(0)
    if g.workspace or not lib.grappa_b200_gemm_can_fuse_colsum(C.byref(g)):
        return gemm(a, b, **kw), None
    if _gemm_record is not None:
        _gemm_record.append(("single", g, 1, (a, b, out, kw, partial)))
    _lib.check(lib.grappa_b200_gemm(C.byref(g), _s()), "gemm")
    return out, partial


def gemm_grouped(problems) -> None:
    """`problems`: list of (a, b, kwargs) as for `gemm` with `out=` given.  Independent GEMMs issued through
    grappa_b200_gemm_grouped: runs of up to four tensor-core problems share one persistent launch."""
    if not problems:
        return
    lib = _lib.lib()
    arr = (GemmArgs * len(problems))()
    flops, shapes = 0.0, []
    for i, (a, b, kw) in enumerate(problems):
        g, _ = _gemm_args(a, b, **kw)
        arr[i] = g
        flops += 2.0 * g.M * g.N * g.K
        shapes.append((g.M, g.N, g.K))
    if _gemm_record is not None:
        _gemm_record.append(("grouped", arr, len(problems), problems))
    if _gemm_profile is not None:
        ext = torch.cuda.is_current_stream_capturing()
        e0, e1 = torch.cuda.Event(enable_timing=True, external=ext), torch.cuda.Event(enable_timing=True, external=ext)
        e0.record()
        _lib.check(lib.grappa_b200_gemm_grouped(arr, len(problems), _s()), "gemm_grouped")
        e1.record()
        _gemm_profile.append((flops, e0, e1, ("grouped",) + tuple(shapes)))
        return
    _lib.check(lib.grappa_b200_gemm_grouped(arr, len(problems), _s()), "gemm_grouped")


def neighbor_mean(x: torch.Tensor, pack, transpose: bool = False) -> torch.Tensor:
    """SAGEConv('mean') aggregation over the bonded graph (mode 0) or its transpose (backward, mode 1)."""
    lib = _lib.lib()
    _lib.require_cuda(x)
    assert x.dim() == 2 and x.stride(1) == 1
    out = torch.empty((x.shape[0], x.shape[1]), device=x.device, dtype=torch.float32)
    _lib.check(lib.grappa_b200_neighbor_mean(x.data_ptr(), x.stride(0), pack.ptr("indptr"), pack.ptr("esrc"), out.data_ptr(),
                                             out.stride(0), x.shape[0], x.shape[1], int(transpose), _s()), "neighbor_mean")
    return out


def pad_rows(w: torch.Tensor, ld: int) -> torch.Tensor:
    """Zero-padded copy [rows, ld] of a 2-D matrix (rows made 16-byte multiples for TMA)."""
    lib = _lib.lib()
    out = torch.empty((w.shape[0], ld), device=w.device, dtype=torch.float32)
    _lib.check(lib.grappa_b200_pad_rows(w.data_ptr(), w.shape[0], w.shape[1], w.stride(0), out.data_ptr(), ld, _s()), "pad_rows")
    return out


def layernorm_fwd(x, gamma, beta, eps=1e-5):
    lib = _lib.lib()
    rows, cols = x.shape
    y = torch.empty_like(x)
    mean = torch.empty(rows, device=x.device, dtype=torch.float32)
    rstd = torch.empty(rows, device=x.device, dtype=torch.float32)
    _lib.check(lib.grappa_b200_layernorm_fwd(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), y.data_ptr(),
                                             mean.data_ptr(), rstd.data_ptr(), rows, cols, eps, _s()), "layernorm_fwd")
    return y, mean, rstd


def layernorm_bwd(dy, x, mean, rstd, gamma):
    lib = _lib.lib()
    rows, cols = x.shape
    dx = torch.empty_like(x)
    _lib.check(lib.grappa_b200_layernorm_bwd(dy.data_ptr(), x.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                             gamma.data_ptr(), dx.data_ptr(), rows, cols, _s()), "layernorm_bwd")
    return dx


def col_reduce(dy, out_sum=None, x=None, mean=None, rstd=None, out_xhat=None, accumulate=False, cols=None):
    """out_sum[c] = sum_r dy[r,c];  out_xhat[c] = sum_r dy[r,c] * xhat[r,c]  (LayerNorm gamma grad)."""
    lib = _lib.lib()
    rows = dy.shape[0]
    cols = cols if cols is not None else dy.shape[1]
    if out_sum is None:
        out_sum = torch.empty(cols, device=dy.device, dtype=torch.float32)
        accumulate = False
    if x is not None and out_xhat is None:
        out_xhat = torch.empty(cols, device=dy.device, dtype=torch.float32)
    if rows == 0:
        if not accumulate:
            out_sum.zero_()
            if out_xhat is not None:
                out_xhat.zero_()
        return out_sum, out_xhat
    nb = lib.grappa_b200_col_reduce_workspace(rows, cols)
    ws = workspace(nb, dy.device, "colred", zero=True)
    _lib.check(lib.grappa_b200_col_reduce(dy.data_ptr(), dy.stride(0), _p(x), _p(mean), _p(rstd), out_sum.data_ptr(),
                                          _p(out_xhat), ws.data_ptr(), rows, cols, int(accumulate), _s()), "col_reduce")
    return out_sum, out_xhat


def edge_attention_fwd(ft, pack, heads):
    lib = _lib.lib()
    n, hd = ft.shape
    out = torch.empty_like(ft)
    alpha = torch.empty((pack.n_edges, heads), device=ft.device, dtype=torch.float32)
    _lib.check(lib.grappa_b200_edge_attention_fwd(ft.data_ptr(), pack.ptr("indptr"), pack.ptr("esrc"), out.data_ptr(),
                                                  _p(alpha), n, heads, hd // heads, _s()), "edge_attention_fwd")
    return out, alpha


def edge_attention_bwd(ft, alpha, dout, pack, heads):
    lib = _lib.lib()
    n, hd = ft.shape
    ds = torch.empty_like(alpha)
    dft = torch.empty_like(ft)
    _lib.check(lib.grappa_b200_edge_attention_bwd(ft.data_ptr(), alpha.data_ptr(), dout.data_ptr(), pack.ptr("indptr"),
                                                  pack.ptr("esrc"), pack.ptr("erev"), _p(ds), dft.data_ptr(), n, heads,
                                                  hd // heads, _s()), "edge_attention_bwd")
    return dft


def tuple_attention_fwd(qkv, T, L, heads):
    lib = _lib.lib()
    E = qkv.shape[1] // 3
    out = torch.empty((L * T, E), device=qkv.device, dtype=torch.float32)
    _lib.check(lib.grappa_b200_tuple_attention_fwd(_p(qkv), _p(out), T, L, heads, E // heads, _s()), "tuple_attention_fwd")
    return out


def tuple_attention_bwd(qkv, dout, T, L, heads):
    lib = _lib.lib()
    E = qkv.shape[1] // 3
    dqkv = torch.empty_like(qkv)
    _lib.check(lib.grappa_b200_tuple_attention_bwd(_p(qkv), _p(dout), _p(dqkv), T, L, heads, E // heads, _s()),
               "tuple_attention_bwd")
    return dqkv


def tuple_gather_fwd(p, idx, pe, T, L, F, E):
    lib = _lib.lib()
    x = torch.empty((L * T, E), device=p.device, dtype=torch.float32)
    _lib.check(lib.grappa_b200_tuple_gather_fwd(_p(p), p.stride(0), _p(idx), _p(pe), _p(x), T, L, F, E, _s()),
               "tuple_gather_fwd")
    return x


def tuple_gather_bwd(dx, inv_ptr, inv_ent, n_atoms, ldp, T, L, F, E, out=None, accumulate=False):
    lib = _lib.lib()
    if out is None:
        out = torch.empty((n_atoms, ldp), device=dx.device, dtype=torch.float32)
    _lib.check(lib.grappa_b200_tuple_gather_bwd(_p(dx), _p(inv_ptr), _p(inv_ent), out.data_ptr(), out.stride(0), n_atoms, T,
                                                L, F, E, int(accumulate), _s()), "tuple_gather_bwd")
    return out


def make_perms(perms: Sequence[Sequence[int]]) -> Perms:
    p = Perms()
    p.n_perm = len(perms)
    for i, row in enumerate(perms):
        for j, v in enumerate(row):
            p.perm[i][j] = int(v)
    return p


def perm_concat_fwd(x, perms: Perms, T, L, E):
    lib = _lib.lib()
    s = torch.empty((perms.n_perm * T, L * E), device=x.device, dtype=torch.float32)
    _lib.check(lib.grappa_b200_perm_concat_fwd(_p(x), _p(s), C.byref(perms), T, L, E, _s()), "perm_concat_fwd")
    return s


def perm_concat_bwd(ds, perms: Perms, T, L, E):
    lib = _lib.lib()
    dx = torch.empty((L * T, E), device=ds.device, dtype=torch.float32)
    _lib.check(lib.grappa_b200_perm_concat_bwd(_p(ds), _p(dx), C.byref(perms), T, L, E, _s()), "perm_concat_bwd")
    return dx


def featurize(feats: Sequence[torch.Tensor], charge: Optional[torch.Tensor], ld: int, enc_dim: int = 16):
    lib = _lib.lib()
    n = feats[0].shape[0]
    a = FeaturizeArgs()
    a.n_feats = len(feats)
    keep = []
    for i, f in enumerate(feats):
        f = f.float().contiguous()
        keep.append(f)
        a.feats[i] = f.data_ptr()
        a.width[i] = 1 if f.dim() == 1 else f.shape[1]
    if charge is not None:
        charge = charge.float().contiguous()
    a.charge = _p(charge)
    a.enc_dim = enc_dim
    out = torch.empty((n, ld), device=feats[0].device, dtype=torch.float32)
    _lib.check(lib.grappa_b200_featurize(C.byref(a), out.data_ptr(), n, ld, _s()), "featurize")
    return out


def head_output_fwd(args: HeadOutArgs, scores):
    lib = _lib.lib()
    T = args.T
    dev = scores.device
    if args.kind == 2:
        k = torch.empty((T, args.n_per), device=dev, dtype=torch.float32)
        eq = None
    else:
        k = torch.empty(T, device=dev, dtype=torch.float32)
        eq = torch.empty(T, device=dev, dtype=torch.float32)
    _lib.check(lib.grappa_b200_head_output_fwd(C.byref(args), _p(scores), _p(k), _p(eq), _s()), "head_output_fwd")
    return k, eq


def head_output_bwd(args: HeadOutArgs, scores, dk, deq):
    lib = _lib.lib()
    ds = torch.empty_like(scores)
    _lib.check(lib.grappa_b200_head_output_bwd(C.byref(args), _p(scores), _p(dk), _p(deq), _p(ds), _s()), "head_output_bwd")
    return ds


def head_output_stats_bwd(args: HeadOutArgs, scores, dk, deq, targets, accumulate: bool):
    """Gradients of the learnable statistics (gb_head_out_args.stat slot order) written into `targets` (None = skip)."""
    lib = _lib.lib()
    out = HeadStatGrads()
    for i, tgt in enumerate(targets):
        out.d[i] = None if tgt is None else tgt.data_ptr()
    out.accumulate = int(bool(accumulate))
    _lib.check(lib.grappa_b200_head_output_stats_bwd(C.byref(args), _p(scores), _p(dk), _p(deq), C.byref(out), _s()),
               "head_output_stats_bwd")


def dropout(x, p, seed, out=None):
    lib = _lib.lib()
    if out is None:
        out = torch.empty_like(x)
    _lib.check(lib.grappa_b200_dropout(_p(x), _p(out), x.numel(), float(p), int(seed) & 0xFFFFFFFFFFFFFFFF, _rng_ptr(),
                                       _s()), "dropout")
    return out


def act_dropout_bwd(dy, act_out, p, seed):
    lib = _lib.lib()
    dx = torch.empty_like(dy)
    _lib.check(lib.grappa_b200_act_dropout_bwd(_p(dy), _p(act_out), _p(dx), dy.numel(), float(p),
                                               int(seed) & 0xFFFFFFFFFFFFFFFF, _rng_ptr() if p > 0.0 else 0, _s()),
               "act_dropout_bwd")
    return dx


def axpby(x, y, a=1.0, b=1.0):
    lib = _lib.lib()
    _lib.check(lib.grappa_b200_axpby(_p(x), _p(y), x.numel(), float(a), float(b), _s()), "axpby")
    return y


def sumsq(x, out):
    lib = _lib.lib()
    _lib.check(lib.grappa_b200_sumsq(_p(x), x.numel(), out.data_ptr(), _s()), "sumsq")
    return out


def sumsq_det(x, out, ws):
    """out[0] = sum x^2, bit-reproducible; ws: zero-initialised fp32 scratch of >= 1024 elements (reusable)."""
    lib = _lib.lib()
    _lib.check(lib.grappa_b200_sumsq_det(_p(x), x.numel(), out.data_ptr(), ws.data_ptr(), _s()), "sumsq_det")
    return out


def adam_step(p, g, m, v, lr, beta1, beta2, eps, step, gnorm_sq=None, clip=0.0, grad_scale=1.0):
    lib = _lib.lib()
    _lib.check(lib.grappa_b200_adam_step(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), lr, beta1, beta2,
                                         eps, step, _p(gnorm_sq), float(clip), float(grad_scale), _s()), "adam_step")


def adam_step_dev(p, g, m, v, lr_dev, beta1, beta2, eps, step_dev, gnorm_sq=None, clip=0.0, grad_scale=1.0):
    """Adam with learning rate / step count read from device tensors (CUDA-graph friendly)."""
    lib = _lib.lib()
    _lib.check(lib.grappa_b200_adam_step_dev(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(),
                                             lr_dev.data_ptr(), beta1, beta2, eps, step_dev.data_ptr(), _p(gnorm_sq),
                                             float(clip), float(grad_scale), _s()), "adam_step_dev")


def tick(counters):
    """counters[:] += 1 on the device (int64 tensor of <= 32 elements)."""
    lib = _lib.lib()
    _lib.check(lib.grappa_b200_tick(counters.data_ptr(), counters.numel(), _s()), "tick")


def _n_cta(rows: int) -> int:
    """CTAs of the fused backward kernels: every CTA gets >= 8 rows, at most two CTAs per SM."""
    sms = _lib.lib().grappa_b200_sm_count() if _SMS[0] is None else _SMS[0]
    return max(1, min(2 * sms, (rows + 7) // 8))


_SMS = [None]


def layernorm_bwd_fused(dy, x, mean, rstd, gamma):
    """dx plus per-CTA partial sums [n_cta, 2, cols] (set 0: beta gradient, set 1: gamma gradient); see `ColumnSums`."""
    lib = _lib.lib()
    if _SMS[0] is None:
        _SMS[0] = lib.grappa_b200_sm_count()
    rows, cols = x.shape
    n = _n_cta(rows)
    dx = torch.empty_like(x)
    partial = torch.empty((n, 2, cols), device=x.device, dtype=torch.float32)
    _lib.check(lib.grappa_b200_layernorm_bwd_fused(dy.data_ptr(), x.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                                   gamma.data_ptr(), dx.data_ptr(), partial.data_ptr(), n, rows, cols, _s()),
               "layernorm_bwd_fused")
    return dx, partial


def act_dropout_bwd_fused(dy, act_out, p, seed):
    """dx = dy * mask * elu'(act_out) plus per-CTA column sums of dx [n_cta, cols] (bias gradient partials)."""
    lib = _lib.lib()
    if _SMS[0] is None:
        _SMS[0] = lib.grappa_b200_sm_count()
    rows, cols = dy.shape
    n = _n_cta(rows)
    dx = torch.empty_like(dy)
    partial = torch.empty((n, cols), device=dy.device, dtype=torch.float32)
    _lib.check(lib.grappa_b200_act_dropout_bwd_fused(_p(dy), _p(act_out), _p(dx), partial.data_ptr(), n, rows, cols, float(p),
                                                     int(seed) & 0xFFFFFFFFFFFFFFFF, _rng_ptr() if p > 0.0 else 0, _s()),
               "act_dropout_bwd_fused")
    return dx, partial


class ColumnSums:
    """Collects deferred column reductions (partial rows -> out vector) and folds them with one launch per
    COLSUM_MAX entries (`flush`).  Partial buffers are kept alive until the flush has been enqueued."""

    def __init__(self):
        self.items = []

    def add(self, partial, n_part, stride, cols, out, accumulate):
        self.items.append((partial, int(n_part), int(stride), int(cols), out, bool(accumulate)))

    def flush(self):
        if not self.items:
            return
        lib = _lib.lib()
        for i0 in range(0, len(self.items), COLSUM_MAX):
            chunk = self.items[i0:i0 + COLSUM_MAX]
            b = ColsumBatch()
            b.n = len(chunk)
            for j, (partial, n_part, stride, cols, out, acc) in enumerate(chunk):
                d = b.desc[j]
                d.partial, d.out = partial.data_ptr(), out.data_ptr()
                d.n_part, d.stride, d.cols, d.accumulate = n_part, stride, cols, int(acc)
            _lib.check(lib.grappa_b200_finalize_colsums(C.byref(b), _s()), "finalize_colsums")
        self.items = []
