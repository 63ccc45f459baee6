"""Energy+force kernel (K13) micro-benchmark on B200: python tools/energy_bench.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from grappa_b200 import graph as gbg, synthetic
from grappa_b200.energy import Energy

dev = torch.device("cuda")
HBM = 6562.6
FULL = "--full" in sys.argv      # also time the full-contract mode (per-tuple x / energy written)


def build(n_mols, n_confs):
    ge = synthetic.peptide_batch(seed=7, batch_size=8, n_res=4, n_confs=n_confs)
    ge = gbg.batch([ge] * (n_mols // 8)).to(dev)
    gen = torch.Generator().manual_seed(1)
    for l in ("n2", "n3", "n4", "n4_improper"):
        T = ge.num_nodes(l)
        if l in ("n2", "n3"):
            ge.nodes[l].data["k"] = (100 + 300 * torch.rand(T, generator=gen)).to(dev)
            ge.nodes[l].data["eq"] = (1.2 + 0.6 * torch.rand(T, generator=gen)).to(dev)
        else:
            ge.nodes[l].data["k"] = torch.randn(T, 3, generator=gen).to(dev)
    return ge


for n_mols, n_confs in ((1000, 100), (104, 1000), (32, 50)):
    ge = build(n_mols, n_confs)
    na = ge.num_nodes("n1")
    tup = [ge.num_nodes(l) for l in ("n2", "n3", "n4", "n4_improper")]
    alg = n_confs * (24 * na + 4 * n_mols) + 4 * (2 * tup[0] + 3 * tup[1] + 4 * tup[2] + 4 * tup[3]) + 4 * (2 * tup[0] + 2 * tup[1] + 3 * tup[2] + 3 * tup[3])
    ref = None
    alg_full = alg + n_confs * (8 * sum(tup) + 16 * n_mols)
    for full in ((False, True) if FULL else (False,)):
        for variant in ((5, 4) if full else (5, 4, 2, 3, 1)):
            en = Energy(write_tuple_terms=full)
            en.kernel_variant = variant
            with torch.no_grad():
                for _ in range(3):
                    g = en(ge)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                reps = 20
                e0.record()
                for _ in range(reps):
                    g = en(ge)
                e1.record()
                torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            E, F = g.nodes["g"].data["energy"].clone(), g.nodes["n1"].data["gradient"].clone()
            if ref is None:
                ref = (E, F)
            dE = ((E - ref[0]).abs().max() / ref[0].abs().max()).item()
            dF = ((F - ref[1]).abs().max() / ref[1].abs().max()).item()
            by = alg_full if full else alg
            print(f"mols={n_mols} confs={n_confs} {'full' if full else 'lean'} variant={variant}: {ms * 1e3:8.1f} us  "
                  f"{n_mols * n_confs / ms / 1e6:8.1f} M evals/s  {by / ms / 1e6:7.1f} GB/s = {by / ms / 1e6 / HBM * 100:5.1f}% HBM "
                  f"({by / (n_mols * n_confs):.0f} B/eval)   dE={dE:.1e} dF={dF:.1e}", flush=True)
