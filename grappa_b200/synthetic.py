"""Synthetic molecular graphs and conformations of the BASELINE.json shapes.

Datasets and checkpoints are unavailable offline, so every benchmark / parity input is generated
here from a seed:

  * peptide-like  : capped poly-alanine ACE-(ALA)n-NME built atom by atom (12 + 10 n atoms,
                    11 + 10 n bonds, 2 (n + 1) impropers x 3 orderings) -> n=1: 22 atoms,
                    n=4: 52 atoms, n=149: 1502 atoms
  * small-molecule: random trees with max degree 4 plus 0-2 ring-closing bonds, 3-50 atoms
  * rna-like      : 90-100 atom graphs with fused 5/6 rings on a backbone chain

Tuples (angles, propers, impropers) always come from `grappa_b200.tuples`, which reproduces the
reference's `utils/tuple_indices.py` orderings bit-exactly.  Input features follow the grappa-1.2
`in_feat_name` list (experiments/train-grappa-1.2/grappa_config.yaml:91-96): atomic_number
one-hot(53), partial_charge, ring_encoding(7), degree one-hot(6), charge_model(2).
Coordinates are built NeRF-style from internal coordinates (bond 1.0-1.5 A, angle ~109-120 deg,
staggered dihedrals) so that no bonded term sits on a singularity; conformations add N(0, 0.1^2 A^2)
displacements.  Reference labels are random: energy_ref ~ N(0, 3^2) mean-centred per molecule,
gradient_ref ~ N(0, 10^2).
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np
import torch

from . import graph as _graph
from . import tuples as _tuples

MAX_ELEMENT = 53          # reference constants.py:38
CHARGE_MODEL_AMBER99 = (0.0, 1.0)   # reference constants.py:44 ['am1BCC', 'amber99']


# ------------------------------------------------------------------------------------------------
# topology builders
# ------------------------------------------------------------------------------------------------
def polyalanine_topology(n_res: int):
    """ACE-(ALA)n-NME.  Returns (elements, bonds, improper_candidates)."""
    el: List[int] = []
    bonds: List[Tuple[int, int]] = []
    imps: List[Tuple[int, int, int, int]] = []

    def add(z, parent=None):
        el.append(z)
        i = len(el) - 1
        if parent is not None:
            bonds.append((parent, i))
        return i

    # ACE: CH3-C(=O)-
    ch3 = add(6)
    for _ in range(3):
        add(1, ch3)
    c_prev = add(6, ch3)
    o_prev = add(8, c_prev)
    ca_prev = ch3
    for _ in range(n_res):
        n = add(7, c_prev)
        h = add(1, n)
        ca = add(6, n)
        add(1, ca)
        cb = add(6, ca)
        for _ in range(3):
            add(1, cb)
        c = add(6, ca)
        o = add(8, c)
        # impropers centred on the carbonyl carbon and on the amide nitrogen (central atom at index 2)
        imps.append((ca_prev, n, c_prev, o_prev))
        imps.append((c_prev, ca, n, h))
        c_prev, o_prev, ca_prev = c, o, ca
    # NME: -N(H)-CH3
    n = add(7, c_prev)
    h = add(1, n)
    cm = add(6, n)
    for _ in range(3):
        add(1, cm)
    imps.append((ca_prev, n, c_prev, o_prev))
    imps.append((c_prev, cm, n, h))
    return np.array(el, dtype=np.int64), np.array(bonds, dtype=np.int64), imps


def random_tree_topology(rng: np.random.Generator, n_atoms: int, n_ring_closures: int = 0):
    """Random tree with max degree 4 (+ optional ring-closing bonds between atoms 3-6 bonds apart)."""
    n_atoms = max(int(n_atoms), 3)
    deg = np.zeros(n_atoms, dtype=np.int64)
    bonds = []
    heavy_budget = max(2, int(round(n_atoms * 0.45)))
    for i in range(1, n_atoms):
        # prefer attaching to the heavy-atom backbone (first `heavy_budget` atoms)
        cand = np.nonzero(deg[:min(i, heavy_budget)] < 4)[0]
        if len(cand) == 0:
            cand = np.nonzero(deg[:i] < 4)[0]
        p = int(rng.choice(cand))
        bonds.append((p, i))
        deg[p] += 1
        deg[i] += 1
    adj = _adjacency(n_atoms, bonds)
    for _ in range(n_ring_closures):
        for _try in range(20):
            a = int(rng.integers(0, min(n_atoms, heavy_budget)))
            dist = _bfs_dist(adj, a)
            cand = [b for b in range(min(n_atoms, heavy_budget))
                    if 3 <= dist[b] <= 5 and deg[b] < 4 and deg[a] < 4]
            if cand:
                b = int(rng.choice(cand))
                bonds.append((min(a, b), max(a, b)))
                deg[a] += 1
                deg[b] += 1
                adj[a].append(b)
                adj[b].append(a)
                break
    el = _elements_from_degree(rng, deg)
    imps = _improper_candidates(n_atoms, bonds, el)
    return el, np.array(bonds, dtype=np.int64), imps


def rna_like_topology(rng: np.random.Generator, n_atoms: int = 96):
    """Backbone chain carrying fused 5- and 6-membered rings (nucleotide-like), ~90-100 atoms."""
    bonds = []
    n = 0

    def ring(size, attach=None):
        nonlocal n
        ids = list(range(n, n + size))
        n += size
        for a, b in zip(ids, ids[1:] + ids[:1]):
            bonds.append((min(a, b), max(a, b)))
        if attach is not None:
            bonds.append((attach, ids[0]))
        return ids

    backbone_prev = None
    while n < n_atoms - 24:
        sugar = ring(5, backbone_prev)           # ribose-like
        base6 = ring(6, sugar[2])                # base, six-ring
        # fuse a five-ring onto the six-ring (purine-like): shares the bond base6[2]-base6[3]
        extra = list(range(n, n + 3)); n += 3
        chain = [base6[2]] + extra + [base6[3]]
        for a, b in zip(chain[:-1], chain[1:]):
            bonds.append((min(a, b), max(a, b)))
        p = n; n += 1                             # phosphate-like linker
        bonds.append((sugar[4], p))
        backbone_prev = p
    heavy = n
    deg = np.zeros(n_atoms + 64, dtype=np.int64)
    for a, b in bonds:
        deg[a] += 1; deg[b] += 1
    # saturate with terminal atoms up to n_atoms
    i = 0
    while n < n_atoms and i < heavy:
        if deg[i] < 3:
            bonds.append((i, n)); deg[i] += 1; deg[n] += 1; n += 1
        else:
            i += 1
    deg = deg[:n]
    el = _elements_from_degree(rng, deg)
    imps = _improper_candidates(n, bonds, el)
    return el, np.array(bonds, dtype=np.int64), imps


def _adjacency(n_atoms, bonds):
    adj = [[] for _ in range(n_atoms)]
    for a, b in bonds:
        adj[int(a)].append(int(b)); adj[int(b)].append(int(a))
    return adj


def _bfs_dist(adj, s):
    dist = [-1] * len(adj)
    dist[s] = 0
    q = [s]
    for u in q:
        for v in adj[u]:
            if dist[v] < 0:
                dist[v] = dist[u] + 1
                q.append(v)
    return dist


def _elements_from_degree(rng, deg):
    el = np.empty(len(deg), dtype=np.int64)
    for i, d in enumerate(deg):
        if d >= 4:
            el[i] = 6
        elif d == 3:
            el[i] = rng.choice([6, 6, 7])
        elif d == 2:
            el[i] = rng.choice([6, 7, 8, 16])
        else:
            el[i] = rng.choice([1, 1, 1, 1, 8, 9, 17])
    return el


def _improper_candidates(n_atoms, bonds, el):
    """Planar centres: degree-3 C/N atoms get one improper candidate (central atom at index 2)."""
    adj = _adjacency(n_atoms, bonds)
    out = []
    for c in range(n_atoms):
        if len(adj[c]) == 3 and el[c] in (6, 7):
            a, b, d = sorted(adj[c])
            out.append((a, b, c, d))
    return out


# ------------------------------------------------------------------------------------------------
# features
# ------------------------------------------------------------------------------------------------
def ring_encoding(n_atoms: int, bonds) -> np.ndarray:
    """[in ring, in ring of size 3..8] per atom (reference utils/rdkit_utils.py:7-24, rdkit-free): host C++ through the
    C ABI (40 ms -> 0.1 ms for a 1,500-atom protein, which matters next to a 2.4 ms parametrisation)."""
    from . import tuples
    return tuples.ring_encoding(n_atoms, bonds)


def ring_encoding_py(n_atoms: int, bonds) -> np.ndarray:
    """The same search written out in Python: the readable statement of what the C++ routine computes, and its check
    (tests/test_host.py)."""
    adj = _adjacency(n_atoms, bonds)
    enc = np.zeros((n_atoms, 7), dtype=np.float32)
    for a, b in bonds:
        a, b = int(a), int(b)
        # shortest cycle through bond (a,b): BFS from a to b avoiding the bond itself
        dist = {a: 0}
        parent = {a: -1}
        q = [a]
        for u in q:
            if b in dist:
                break
            for v in adj[u]:
                if (u == a and v == b) or v in dist:
                    continue
                dist[v] = dist[u] + 1
                parent[v] = u
                q.append(v)
                if v == b:
                    break
        if b in dist:                      # the bond closes a cycle of dist[b] + 1 atoms
            enc[a, 0] = enc[b, 0] = 1.0    # in a ring of any size (rdkit IsInRing; macrocycles included)
            size = dist[b] + 1
            if 3 <= size <= 8:
                v = b
                while v != -1:
                    enc[v, size - 2] = 1.0
                    v = parent[v]
    return enc


def atom_features(rng, el, bonds, charge_model=CHARGE_MODEL_AMBER99) -> Dict[str, np.ndarray]:
    n = len(el)
    deg = np.zeros(n, dtype=np.int64)
    for a, b in bonds:
        deg[int(a)] += 1; deg[int(b)] += 1
    onehot = np.zeros((n, MAX_ELEMENT), dtype=np.float32)
    onehot[np.arange(n), el - 1] = 1.0
    degree = np.zeros((n, 6), dtype=np.float32)
    degree[np.arange(n), np.clip(deg, 1, 6) - 1] = 1.0
    q = np.clip(rng.normal(0.0, 0.3, size=n), -1.0, 1.0).astype(np.float32)
    return {
        "atomic_number": onehot,
        "partial_charge": q,
        "ring_encoding": ring_encoding(n, bonds),
        "degree": degree,
        "charge_model": np.tile(np.asarray(charge_model, dtype=np.float32), (n, 1)),
    }


# ------------------------------------------------------------------------------------------------
# coordinates
# ------------------------------------------------------------------------------------------------
def _place(p, gp, ggp, r, theta, phi):
    """NeRF: position at distance r from p, angle theta to gp, dihedral phi w.r.t. ggp."""
    bc = p - gp
    bc /= np.linalg.norm(bc)
    n = np.cross(gp - ggp, bc)
    nn = np.linalg.norm(n)
    if nn < 1e-6:
        n = np.cross(bc, np.array([1.0, 0.0, 0.0]))
        if np.linalg.norm(n) < 1e-6:
            n = np.cross(bc, np.array([0.0, 1.0, 0.0]))
        nn = np.linalg.norm(n)
    n /= nn
    m = np.cross(n, bc)
    d = np.array([-r * np.cos(theta), r * np.sin(theta) * np.cos(phi), r * np.sin(theta) * np.sin(phi)])
    return p + d[0] * bc + d[1] * m + d[2] * n


def embed(rng, n_atoms, bonds) -> np.ndarray:
    """(n_atoms, 3) float64 coordinates from a BFS spanning tree with tetrahedral-ish internals."""
    adj = _adjacency(n_atoms, bonds)
    xyz = np.zeros((n_atoms, 3))
    placed = np.zeros(n_atoms, dtype=bool)
    parent = -np.ones(n_atoms, dtype=np.int64)
    nchild = np.zeros(n_atoms, dtype=np.int64)
    base_phi = rng.uniform(-np.pi, np.pi, size=n_atoms)
    placed[0] = True
    order = [0]
    for u in order:
        for v in adj[u]:
            if placed[v]:
                continue
            parent[v] = u
            r = rng.uniform(1.0, 1.5)
            theta = np.deg2rad(rng.uniform(105.0, 121.0))
            k = nchild[u]
            nchild[u] += 1
            gp = parent[u]
            if gp < 0:
                # root's children: spread on a cone around +x, first child along +x
                if k == 0:
                    xyz[v] = xyz[u] + np.array([r, 0.0, 0.0])
                else:
                    first = adj[u][0] if placed[adj[u][0]] else order[1]
                    ref = xyz[first] + np.array([0.3, 1.0, 0.2])
                    xyz[v] = _place(xyz[u], xyz[first], ref, r, theta, base_phi[u] + 2.0944 * k)
            else:
                ggp = parent[gp]
                ggp_pos = xyz[ggp] if ggp >= 0 else xyz[gp] + np.array([0.2, 0.9, 0.4])
                xyz[v] = _place(xyz[u], xyz[gp], ggp_pos, r, theta, base_phi[u] + 2.0944 * k)
            placed[v] = True
            order.append(v)
    return xyz


def conformations(rng, xyz0: np.ndarray, n_confs: int, sigma: float = 0.1) -> np.ndarray:
    """(n_atoms, n_confs, 3) float32, atom-major as reference data/MolData.py:193."""
    n = xyz0.shape[0]
    out = xyz0[:, None, :] + rng.normal(0.0, sigma, size=(n, n_confs, 3))
    return out.astype(np.float32)


# ------------------------------------------------------------------------------------------------
# molecules -> graphs
# ------------------------------------------------------------------------------------------------
def make_molecule(rng, kind: str = "peptide", n_confs: int = 50, n_res: int = 4, n_atoms: int | None = None,
                  labels: bool = True) -> _graph.MolGraph:
    if kind == "peptide":
        el, bonds, imp_cand = polyalanine_topology(n_res)
    elif kind == "small":
        n = int(n_atoms if n_atoms is not None else rng.integers(3, 51))
        el, bonds, imp_cand = random_tree_topology(rng, n, int(rng.integers(0, 3)) if n >= 8 else 0)
    elif kind == "rna":
        el, bonds, imp_cand = rna_like_topology(rng, int(n_atoms if n_atoms is not None else rng.integers(90, 101)))
    else:
        raise ValueError(f"unknown molecule kind {kind!r}")
    n = len(el)
    tup = _tuples.build_tuples(n, bonds, imp_cand)
    feats = atom_features(rng, el, bonds)
    xyz = conformations(rng, embed(rng, n, bonds), n_confs) if n_confs > 0 else None
    g = _graph.graph_from_molecule(n, tup["bonds"], tup["angles"], tup["propers"], tup["impropers"], feats, xyz)
    if labels and n_confs > 0:
        e = rng.normal(0.0, 3.0, size=(1, n_confs)).astype(np.float32)
        e -= e.mean(axis=1, keepdims=True)
        g.nodes["g"].data["energy_ref"] = torch.from_numpy(e)
        g.nodes["n1"].data["gradient_ref"] = torch.from_numpy(
            rng.normal(0.0, 10.0, size=(n, n_confs, 3)).astype(np.float32))
    return g


def peptide_batch(seed: int = 0, batch_size: int = 32, n_res: int = 4, n_confs: int = 50) -> _graph.MolGraph:
    """BASELINE config 2: 32 x ACE-(ALA)4-NME (52 atoms), 50 conformations each."""
    rng = np.random.default_rng(seed)
    return _graph.batch([make_molecule(rng, "peptide", n_confs=n_confs, n_res=n_res) for _ in range(batch_size)])


def dipeptide(seed: int = 0, n_confs: int = 50) -> _graph.MolGraph:
    """BASELINE config 1: capped dipeptide ACE-ALA-NME, 22 atoms, 50 conformations."""
    return make_molecule(np.random.default_rng(seed), "peptide", n_confs=n_confs, n_res=1)


def protein(seed: int = 0, n_res: int = 149) -> _graph.MolGraph:
    """BASELINE config 3: ~1,500-atom protein graph (no conformations)."""
    return make_molecule(np.random.default_rng(seed), "peptide", n_confs=0, n_res=n_res, labels=False)


def espaloma_mix_batch(seed: int = 0, batch_size: int = 32, n_confs: int = 32) -> _graph.MolGraph:
    """BASELINE config 5: ~85 % small molecules / 8 % peptides / 7 % RNA, C=32.

    Every molecule carries >= 1 improper-free safe layout: molecules without planar centres simply
    have zero impropers (our loss handles that; the reference's would NaN, training/loss.py:130-132).
    """
    rng = np.random.default_rng(seed)
    mols = []
    for _ in range(batch_size):
        u = rng.uniform()
        if u < 0.85:
            mols.append(make_molecule(rng, "small", n_confs=n_confs))
        elif u < 0.93:
            mols.append(make_molecule(rng, "peptide", n_confs=n_confs, n_res=int(rng.integers(1, 5))))
        else:
            mols.append(make_molecule(rng, "rna", n_confs=n_confs))
    return _graph.batch(mols)


def deterministic_state_dict(reference_sd: Dict[str, torch.Tensor], seed: int = 0) -> Dict[str, torch.Tensor]:
    """Weights as a pure function of (key, shape, seed) so that the reference model and ours can be
    given IDENTICAL parameters without shipping a 163 MB checkpoint.

    Floating tensors that are trainable-shaped get U(-1/sqrt(fan_in), 1/sqrt(fan_in)) (torch's
    nn.Linear default scale); LayerNorm weights 1 + 0.1 U(-1,1), LayerNorm / Linear biases
    0.1 U(-1,1) scaled.  Buffers (statistics, permutations, positional encodings, n_periodicity)
    are passed through unchanged.  Aliased keys (gnn.blocks.i == gnn.conv_blocks.i for the first n_conv blocks, then
    gnn.att_blocks.(i - n_conv); reference graph_attention.py:129) hash to the same values because the alias is
    normalised first.
    """
    import hashlib
    import re
    out = {}
    n_conv = len({k.split(".")[2] for k in reference_sd if k.startswith("gnn.conv_blocks.")})

    def canonical(key):
        m = re.match(r"gnn\.blocks\.(\d+)\.(.*)", key)
        if not m:
            return key
        i = int(m.group(1))
        return f"gnn.conv_blocks.{i}.{m.group(2)}" if i < n_conv else f"gnn.att_blocks.{i - n_conv}.{m.group(2)}"
    buffer_tags = ("positional_encoding", "permutation", "n_periodicity", "k_mean", "k_std", "to_k.", "to_eq.")
    for key, ref in reference_sd.items():
        if any(t in key for t in buffer_tags) or not torch.is_floating_point(ref):
            out[key] = ref.clone()
            continue
        canon = canonical(key)
        h = int.from_bytes(hashlib.sha256(f"{seed}:{canon}".encode()).digest()[:8], "little") % (2 ** 63)
        gen = torch.Generator().manual_seed(h)
        u = torch.rand(ref.shape, generator=gen, dtype=torch.float32) * 2.0 - 1.0
        is_norm = "norm" in key
        if ref.dim() >= 2:
            out[key] = u / float(ref.shape[1]) ** 0.5
        elif is_norm and key.endswith("weight"):
            out[key] = 1.0 + 0.1 * u
        else:
            out[key] = 0.1 * u
    return out
