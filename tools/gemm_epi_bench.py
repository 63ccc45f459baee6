"""Epilogue-variant micro-benchmark of the tcgen05 GEMM (B200 only): which fused epilogue terms cost what.
    python tools/gemm_epi_bench.py [M N K]
"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from grappa_b200 import ops

def main():
    M, N, K = (int(x) for x in sys.argv[1:4]) if len(sys.argv) >= 4 else (14848, 512, 512)
    ops.set_matmul_precision("tf32")
    dev = torch.device("cuda")
    a = torch.randn(M, K, device=dev); b = torch.randn(N, K, device=dev)
    bias = torch.randn(N, device=dev); res = torch.randn(M, N, device=dev); out = torch.empty(M, N, device=dev)
    act_out = torch.empty(M, N, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    variants = {
        "plain": dict(),
        "bias": dict(bias=bias),
        "bias+elu": dict(bias=bias, act=1),
        "bias+drop": dict(bias=bias, dropout_p=0.5, dropout_seed=3),
        "bias+res": dict(bias=bias, residual=res),
        "bias+drop+res": dict(bias=bias, dropout_p=0.5, dropout_seed=3, residual=res),
        "bias+elu+act_out+drop+res": dict(bias=bias, act=1, dropout_p=0.3, dropout_seed=3, residual=res, act_out=act_out),
    }
    for name, kw in variants.items():
        def call():
            ops.gemm(a, b, out=out, **kw)
        for _ in range(3):
            call()
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            call()
        for cold in (True, False):
            ts = []
            for _ in range(15):
                if cold:
                    flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); gr.replay(); e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ts.sort()
            ms = ts[len(ts) // 2]
            print(f"{name:28s} {'cold' if cold else 'warm'}  {ms * 1e3:7.1f} us  {2.0 * M * N * K / ms / 1e9:7.1f} TFLOP/s", flush=True)

if __name__ == "__main__":
    main()
