"""Time one GEMM shape through grappa_b200_gemm (B200): python tools/gemm_one.py M N K [ta tb] (env: GRAPPA_B200_GEMM_BN/PAIR)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from grappa_b200 import ops
M, N, K = (int(x) for x in sys.argv[1:4])
ta, tb = (int(sys.argv[4]), int(sys.argv[5])) if len(sys.argv) > 5 else (0, 0)
ops.set_matmul_precision(os.environ.get("GRAPPA_B200_PREC", "tf32"))
dev = torch.device("cuda")
a = torch.randn((K, M) if ta else (M, K), device=dev)
b = torch.randn((K, N) if tb else (N, K), device=dev)
out = torch.empty(M, N, device=dev)
for _ in range(3):
    ops.gemm(a, b, trans_a=bool(ta), trans_b=bool(tb), out=out)
torch.cuda.synchronize()
ref = (a.t() if ta else a).double() @ (b if tb else b.t()).double()
err = ((out.double() - ref).abs().max() / ref.abs().max()).item()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 10
e0.record()
for _ in range(reps):
    ops.gemm(a, b, trans_a=bool(ta), trans_b=bool(tb), out=out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"{os.environ.get('GRAPPA_B200_PREC', 'tf32'):7s} M={M} N={N} K={K} ta={ta} tb={tb} BN={os.environ.get('GRAPPA_B200_GEMM_BN','auto')} PAIR={os.environ.get('GRAPPA_B200_GEMM_PAIR','auto')}: "
      f"{ms * 1e3:8.1f} us  {2.0 * M * N * K / ms / 1e9:7.1f} TFLOP/s  max rel err {err:.2e}", flush=True)
