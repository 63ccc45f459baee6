"""Shared helpers for the test-suite: fixtures -> graphs, tolerances."""
import os

import numpy as np
import torch

from grappa_b200.graph import MolGraph

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
LEVELS = ("n2", "n3", "n4", "n4_improper")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def graph_from_fixture(z, prefix="in.") -> MolGraph:
    counts = {k[len(prefix) + 6:]: torch.from_numpy(z[k]) for k in z.files if k.startswith(prefix + "count.")}
    num = {nt: int(c.sum()) for nt, c in counts.items()}
    g = MolGraph(num, torch.from_numpy(z[prefix + "src"]), torch.from_numpy(z[prefix + "dst"]), counts)
    for k in z.files:
        if not k.startswith(prefix) or k.startswith(prefix + "count.") or k in (prefix + "src", prefix + "dst"):
            continue
        rest = k[len(prefix):]
        for nt in sorted(num.keys(), key=len, reverse=True):
            if rest.startswith(nt + "."):
                g.nodes[nt].data[rest[len(nt) + 1:]] = torch.from_numpy(z[k])
                break
    return g


def rel_err(a, b):
    """max |a-b| / max |b| -- the 'relative' of BASELINE.json's tolerances (norm-wise, per tensor)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if b.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))


def sampled_gradient_errors(z, grads):
    """{parameter name: max |g - g_ref| / max |g_ref|} over the 16 sampled entries per tensor that
    tests/golden/train_batch_grappa12.npz stores for all 305 parameter tensors (`grads`: name -> array or None)."""
    keys = list(z["meta.grad_norms_keys"])
    sidx, sval, gmax = z["meta.grad_sample_idx"], z["meta.grad_sample_val"], z["meta.grad_maxabs"]
    floor = 1e-6 * float(gmax.max())
    out = {}
    for i, k in enumerate(keys):
        g = grads.get(k)
        got = np.zeros(16) if g is None else np.asarray(g, dtype=np.float64).reshape(-1)[sidx[i]]
        out[k] = float(np.max(np.abs(got - sval[i].astype(np.float64))) / max(gmax[i], floor))
    return out
