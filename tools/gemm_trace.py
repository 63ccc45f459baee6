"""Pipeline timeline of ONE tensor-core GEMM launch (debug build of the library with -DGB_GEMM_TRACE):
clock64() stamps per k-block of the TMA producer, the converter warps, the MMA issuer and the epilogue of a few CTAs.

    python tools/gemm_trace.py build                 # here: compiles tools/_trace/libgrappa_b200_trace.so
    GRAPPA_B200_LIB=tools/_trace/libgrappa_b200_trace.so GRAPPA_B200_PREC=bf16x3 python tools/gemm_trace.py M N K [ta tb]   # B200
"""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tools", "_trace")


def build(extra=(), name="libgrappa_b200_trace.so"):
    from grappa_b200 import build as b
    os.makedirs(OUT, exist_ok=True)
    objs = []
    for src in b._sources():
        obj = os.path.join(OUT, src + ".o")
        cmd = [b._nvcc()] + [f for f in b.NVCC_FLAGS if f not in ("-Xptxas", "-v")] + [*(() if "-DNOTRACE" in extra else ("-DGB_GEMM_TRACE",)), *extra, "-I", b.INCLUDE, "-c", os.path.join(b.CSRC, src), "-o", obj]
        subprocess.run(cmd, check=True)
        objs.append(obj)
    lib = os.path.join(OUT, name)
    subprocess.run([b._nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs, check=True)
    for o in objs:
        os.remove(o)
    print(lib)


def main():
    import ctypes as C
    import numpy as np
    import torch
    from grappa_b200 import _lib, ops
    M, N, K = (int(x) for x in sys.argv[1:4])
    ta, tb = (int(sys.argv[4]), int(sys.argv[5])) if len(sys.argv) > 5 else (0, 0)
    prec = os.environ.get("GRAPPA_B200_PREC", "bf16x3")
    ops.set_matmul_precision(prec)
    dev = torch.device("cuda")
    a = torch.randn((K, M) if ta else (M, K), device=dev)
    b = torch.randn((K, N) if tb else (N, K), device=dev)
    out = torch.empty(M, N, device=dev)
    for _ in range(3):
        ops.gemm(a, b, trans_a=bool(ta), trans_b=bool(tb), out=out)
    torch.cuda.synchronize()
    n_cta = 148
    buf = torch.zeros(n_cta * 128 * 8, dtype=torch.int64, device=dev)
    lib = _lib.lib()
    lib.grappa_b200_debug_set_gemm_trace.argtypes = [C.c_void_p]
    assert lib.grappa_b200_debug_set_gemm_trace(buf.data_ptr()) == 0
    ops.gemm(a, b, trans_a=bool(ta), trans_b=bool(tb), out=out)
    torch.cuda.synchronize()
    lib.grappa_b200_debug_set_gemm_trace(None)
    t = buf.cpu().numpy().reshape(n_cta, 128, 8)
    names = ["tma_issue", "conv_start", "conv_done", "mma_issue", "mma_commit", "fwd_seen", "epi_start", "epi_end"]
    for cta in (0, 1, 72, 73):
        tt = t[cta].astype(np.int64)
        nz = tt[tt > 0]
        if nz.size == 0:
            continue
        t0 = nz.min()
        rel = np.where(tt > 0, tt - t0, -1)
        print(f"--- {prec} M={M} N={N} K={K} ta={ta} tb={tb}  CTA {cta}: clocks since the CTA's first stamp")
        print("  it " + " ".join(f"{n:>10s}" for n in names[:6]))
        n_it = int((tt[:, 0] > 0).sum())
        for it in range(min(n_it, 80)):
            print(f"{it:4d} " + " ".join(f"{rel[it, e]:10d}" for e in range(6)))
        tiles = int((tt[:, 7] > 0).sum())
        for lt in range(tiles):
            print(f"  tile {lt}: epilogue start {rel[lt, 6]}, end {rel[lt, 7]}")
        if n_it > 8:
            d = np.diff(tt[:n_it, 0])
            print(f"  tma_issue period: median {np.median(d):.0f} clk, mean {d.mean():.0f} over {n_it} k-blocks")


if __name__ == "__main__":
    if sys.argv[1:2] == ["build"]:   # build [name.so -Dflag ...]
        build(tuple(sys.argv[3:]), sys.argv[2]) if len(sys.argv) > 2 else build()
    else:
        main()
