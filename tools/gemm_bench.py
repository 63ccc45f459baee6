"""Micro-benchmark of grappa_b200's GEMM entry point over the shapes of one training step (B200 only).

    python tools/gemm_bench.py [--precision tf32|fp32] [--check]

Prints per-shape CUDA-event time (L2 flushed between launches) and TFLOP/s; --check compares against torch.matmul (fp32).
"""
import argparse
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from grappa_b200 import ops

SHAPES = [  # (M, N, K, trans_a, trans_b, tag)
    (1664, 512, 512, 0, 0, "gnn fc/head_reducer"),
    (1664, 2048, 512, 0, 0, "gnn ff1"),
    (1664, 512, 2048, 0, 0, "gnn ff2"),
    (1664, 2048, 256, 0, 0, "rep projectors (stacked)"),
    (3264, 512, 512, 0, 0, "bond tokens"),
    (8640, 512, 512, 0, 0, "angle tokens"),
    (14848, 512, 512, 0, 0, "proper tokens"),
    (14848, 1536, 512, 0, 0, "proper in_proj"),
    (7424, 256, 2048, 0, 0, "proper symmetriser l1"),
    (7424, 256, 256, 0, 0, "symmetriser 256"),
    (14848, 512, 512, 0, 1, "proper dgrad"),
    (14848, 512, 1536, 0, 1, "proper in_proj dgrad"),
    (1664, 512, 2048, 0, 1, "gnn ff1 dgrad"),
    (512, 512, 14848, 1, 1, "proper wgrad"),
    (1536, 512, 14848, 1, 1, "proper in_proj wgrad"),
    (2048, 512, 1664, 1, 1, "gnn ff1 wgrad"),
    (256, 2048, 7424, 1, 1, "symmetriser l1 wgrad"),
    (512, 512, 1664, 1, 1, "gnn 512 wgrad"),
    (512, 512, 3264, 1, 1, "bond wgrad"),
    (512, 512, 8640, 1, 1, "angle wgrad"),
    (256, 256, 7424, 1, 1, "symmetriser 256 wgrad"),
    (256, 256, 1920, 1, 1, "improper symmetriser 256 wgrad"),
    (3264, 512, 512, 0, 0, "bond tokens fwd"),
    (3840, 512, 512, 0, 1, "improper dgrad"),
    (8640, 512, 512, 0, 1, "angle dgrad"),
    (1664, 2048, 512, 0, 1, "gnn ff2 dgrad"),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="tf32")
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--only", default="", help="substring filter on the shape tag")
    args = ap.parse_args()
    ops.set_matmul_precision(args.precision)
    dev = torch.device("cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    tot_ms, tot_fl = 0.0, 0.0
    for M, N, K, ta, tb, tag in SHAPES:
        if args.only and args.only not in tag:
            continue
        a = torch.randn((K, M) if ta else (M, K), device=dev)
        b = torch.randn((K, N) if tb else (N, K), device=dev)
        bias = torch.randn(N, device=dev)
        out = torch.empty(M, N, device=dev)
        def call():
            ops.gemm(a, b, trans_a=bool(ta), trans_b=bool(tb), bias=None if ta else bias, act=0 if ta else 1, out=out)
        for _ in range(3):
            call()
        torch.cuda.synchronize()
        # one launch per graph replay: no host launch latency inside the timed region; L2 flushed before each
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            call()
        ts = []
        for _ in range(args.reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            gr.replay()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        ms = ts[len(ts) // 2]
        fl = 2.0 * M * N * K
        tot_ms += ms; tot_fl += fl
        err = ""
        if args.check:
            A = a.t() if ta else a
            Bm = b if tb else b.t()
            ref = A.double() @ Bm.double()
            if not ta:
                ref = torch.nn.functional.elu(ref + bias.double())
            err = f" max rel err {((out.double() - ref).abs().max() / ref.abs().max()).item():.2e}"
        print(f"{tag:28s} M={M:6d} N={N:5d} K={K:6d} ta={ta} tb={tb}  {ms * 1e3:8.1f} us  {fl / ms / 1e9:8.1f} TFLOP/s{err}", flush=True)
    print(f"TOTAL {tot_ms * 1e3:.1f} us, {tot_fl / tot_ms / 1e9:.1f} TFLOP/s")


if __name__ == "__main__":
    main()
