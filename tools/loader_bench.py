"""Host-side cost of feeding one GPU: PackedDataset.collate (+ index tables / round schedule) per batch, the PrefetchLoader
rate for 1, 2 and 4 worker threads, and the topology -> graph path of a 1,502-atom protein (profiles/r1_summary.md
section 13).  CPU only.

    python tools/loader_bench.py [--molecules 64] [--batch 32] [--confs 50]
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from grappa_b200 import dataset, inference, synthetic  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--molecules", type=int, default=64)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--confs", type=int, default=50)
    ap.add_argument("--reps", type=int, default=50)
    a = ap.parse_args()
    rng = np.random.default_rng(0)
    mols = [synthetic.make_molecule(rng, "peptide", n_confs=a.confs, n_res=4) for _ in range(a.molecules)]
    ds = dataset.PackedDataset.from_graphs(mols)
    idx = list(range(a.batch))
    for build_pack in (False, True):
        t0 = time.perf_counter()
        for _ in range(a.reps):
            ds.collate(idx, conf_strategy=a.confs, build_pack=build_pack)
        print(f"collate of {a.batch} peptides x {a.confs} conformations, index tables {'on ' if build_pack else 'off'}: "
              f"{(time.perf_counter() - t0) / a.reps * 1e3:.2f} ms")
    batches = [rng.permutation(a.molecules)[:a.batch].tolist() for _ in range(2 * a.reps)]
    for w in (1, 2, 4):
        t0 = time.perf_counter()
        n = sum(1 for _ in dataset.PrefetchLoader(ds, batches, conf_strategy=a.confs, pin=False, workers=w, depth=2 * w))
        print(f"PrefetchLoader, {w} worker thread(s): {(time.perf_counter() - t0) / n * 1e3:.2f} ms per batch")
    prot = synthetic.protein(seed=3, n_res=149)
    n = prot.num_nodes("n1")
    src, dst = [t.numpy() for t in prot.edges()]
    bonds = [(int(u), int(v)) for u, v in zip(src, dst) if u < v]
    z = (prot.nodes["n1"].data["atomic_number"].argmax(1) + 1).tolist()
    q = prot.nodes["n1"].data["partial_charge"].tolist()
    imp = prot.nodes["n4_improper"].data["idxs"].numpy()[::3].tolist()
    t0 = time.perf_counter()
    for _ in range(10):
        inference.molecule_graph(list(range(n)), bonds, z, q, imp)
    print(f"molecule_graph (topology -> model input graph), {n} atoms: {(time.perf_counter() - t0) / 10 * 1e3:.2f} ms")


if __name__ == "__main__":
    main()
