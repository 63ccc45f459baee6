"""One fused-epilogue GEMM (bias + dropout + residual) for ncu: python tools/gemm_epi_one.py [M N K]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from grappa_b200 import ops
M, N, K = (int(x) for x in sys.argv[1:4]) if len(sys.argv) >= 4 else (14848, 512, 512)
ops.set_matmul_precision("tf32")
dev = torch.device("cuda")
a = torch.randn(M, K, device=dev); b = torch.randn(N, K, device=dev)
bias = torch.randn(N, device=dev); res = torch.randn(M, N, device=dev); out = torch.empty(M, N, device=dev)
for _ in range(3):
    ops.gemm(a, b, bias=bias, dropout_p=0.5, dropout_seed=3, residual=res, out=out)
torch.cuda.synchronize()
