"""Launch one GEMM shape a few times (for ncu): python tools/gemm_one.py M N K ta tb [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from grappa_b200 import ops
M, N, K, ta, tb = [int(x) for x in sys.argv[1:6]]
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 3
ops.set_matmul_precision("tf32")
dev = torch.device("cuda")
a = torch.randn((K, M) if ta else (M, K), device=dev)
b = torch.randn((K, N) if tb else (N, K), device=dev)
bias = torch.randn(N, device=dev)
out = torch.empty(M, N, device=dev)
for _ in range(reps):
    ops.gemm(a, b, trans_a=bool(ta), trans_b=bool(tb), bias=bias, act=1, out=out)
torch.cuda.synchronize()
