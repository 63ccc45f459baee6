"""Packed, kernel-ready view of a batched molecular graph.

The kernels want int32 indices and flat segment tables rather than a graph object:

  atom_off[B+1]            first atom of every molecule
  idx[l]  (T_l, L_l) int32 tuple atom indices per level (n2, n3, n4, n4_improper), already batched
  tup_off[l][B+1]          first tuple of every molecule per level
  edge CSR by destination  indptr[N+1], src[E]  -- the bonded graph is symmetric, so the CSR of
                           in-edges doubles as the CSR of out-edges; `rev[e]` is the position of the
                           reverse edge, which makes the attention backward a pure gather
  inv CSR per level        atom -> (tuple, slot) incidence lists, so the backward of the tuple gather
                           is a deterministic segmented sum instead of atomics

A pack is built once per batch on the HOST (numpy, in the data loader / collate step) and moved to
the device with the graph; it is cached on the graph object.  Index range is validated here, on the
host, instead of by the device->host syncs of reference models/grappa.py:122-128.
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np
import torch

LEVELS = ("n2", "n3", "n4", "n4_improper")
TUPLE_LEN = (2, 3, 4, 4)


ENERGY_SCHED_GROUPS = 8   # tuples per round = warps per CTA of energy_rounds_kernel


def conflict_free_rounds(idx: np.ndarray, tup_off: np.ndarray, n_mols: int, L: int, groups: int):
    """(round_off[n_mols+1], sched[n_rounds, groups]) -- see grappa_b200_conflict_free_rounds."""
    import ctypes as C
    from . import _lib
    lib = _lib.lib()
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    tup_off = np.ascontiguousarray(tup_off, dtype=np.int32)
    ro = np.zeros(n_mols + 1, dtype=np.int32)
    ip = idx.ctypes.data_as(C.c_void_p) if idx.size else None
    # a molecule never needs more rounds than it has tuples, so one pass into a buffer of n_tuples rounds is enough
    # (this runs on the data-loader path once per level and batch)
    cap = int(tup_off[n_mols]) if n_mols else 0
    buf = np.empty((cap, groups), dtype=np.int32)
    n = lib.grappa_b200_conflict_free_rounds(ip, tup_off.ctypes.data_as(C.c_void_p), n_mols, L, groups,
                                             ro.ctypes.data_as(C.c_void_p), buf.ctypes.data_as(C.c_void_p) if cap else None,
                                             cap)
    if n < 0:
        _lib.check(int(n), "conflict_free_rounds")
    sc = buf[:int(n)]
    return ro, sc


def _offsets(counts) -> np.ndarray:
    out = np.zeros(len(counts) + 1, dtype=np.int32)
    np.cumsum(np.asarray(counts, dtype=np.int64), out=out[1:])
    return out


class PackedBatch:
    """Device-resident int32 index tables of one batch (see module docstring)."""

    # order of the tables inside the flat buffer (each starts on a 16-byte boundary)
    TABLE_ORDER = (("atom_off",) + tuple(f"{n}{l}" for l in range(4) for n in ("idx", "tup_off", "inv_ptr", "inv_ent"))
                   + tuple(f"{n}{l}" for l in range(4) for n in ("round_off", "sched")) + ("indptr", "esrc", "erev"))

    @classmethod
    def from_device(cls, flat: torch.Tensor, offsets: Dict[str, int], sizes: Dict[str, int], meta: dict) -> "PackedBatch":
        """A pack whose tables were assembled ON THE DEVICE (dataset.DeviceDataset.collate) in the flat layout above.
        `meta`: n_atoms, n_mols, n_edges, n_tuples, max_atoms_per_mol, max_degree, max_tuples_per_mol, max_rounds_per_mol."""
        p = cls.__new__(cls)
        p.device = flat.device
        for k, v in meta.items():
            setattr(p, k, v)
        p.sched_groups = ENERGY_SCHED_GROUPS
        p.host = None
        p._names = list(cls.TABLE_ORDER)
        p._sizes = [int(sizes[k]) for k in p._names]
        p._offsets = [int(offsets[k]) for k in p._names]
        p._stage = None
        p._flat = flat
        p.bytes = flat.numel() * 4
        shapes = {}
        for l, L in enumerate(TUPLE_LEN):
            shapes[f"idx{l}"] = (-1, L)
            shapes[f"sched{l}"] = (-1, ENERGY_SCHED_GROUPS)
        p._dev = {k: flat[o:o + s].view(shapes.get(k, (-1,))) for k, s, o in zip(p._names, p._sizes, p._offsets)}
        return p

    def __init__(self, g, device=None):
        n_atoms = g.num_nodes("n1")
        dev = device if device is not None else g.nodes["n1"].data[next(iter(g.nodes["n1"].data))].device
        self.device = torch.device(dev)
        self.n_atoms = n_atoms
        atom_counts = np.asarray(g.batch_num_nodes("n1").cpu().numpy(), dtype=np.int64)
        self.n_mols = len(atom_counts)
        self.max_atoms_per_mol = int(atom_counts.max()) if len(atom_counts) else 0
        host: Dict[str, np.ndarray] = {"atom_off": _offsets(atom_counts)}
        self.n_tuples: List[int] = []
        for l, (lvl, L) in enumerate(zip(LEVELS, TUPLE_LEN)):
            if lvl in g.ntypes:
                idx_t = g.nodes[lvl].data["idxs"]
                if idx_t.dtype not in (torch.int64, torch.int32):
                    raise IndexError(f"g.nodes['{lvl}'].data['idxs'] has the wrong datatype. It should be a long "
                                     f"but is {idx_t.dtype}")
                idx = idx_t.detach().cpu().numpy().reshape(-1, L)
                counts = g.batch_num_nodes(lvl).cpu().numpy()
            else:
                idx = np.zeros((0, L), dtype=np.int64)
                counts = np.zeros(self.n_mols, dtype=np.int64)
            if idx.size and (idx.max() >= n_atoms or idx.min() < 0):
                # reference models/grappa.py:128 (its '<=' tolerates idx == n_atoms, which would then fail in the gather)
                raise AssertionError(f"Encountered idxs up to {idx.max()} at the level g.nodes[{lvl}].data[\"idxs\"], "
                                     f"but there are only {n_atoms} atom-level-nodes in the graph")
            self.n_tuples.append(int(idx.shape[0]))
            host[f"idx{l}"] = np.ascontiguousarray(idx.astype(np.int32))
            host[f"tup_off{l}"] = _offsets(counts)
            # inverse incidence: for every atom the flat (tuple*L + slot) entries that reference it
            flat = idx.reshape(-1)
            order = np.argsort(flat, kind="stable").astype(np.int32)
            host[f"inv_ptr{l}"] = _offsets(np.bincount(flat, minlength=n_atoms)) if n_atoms else np.zeros(1, np.int32)
            host[f"inv_ent{l}"] = order
        # conflict-free rounds for the energy kernel (host C++, include/grappa_b200.h)
        self.sched_groups = ENERGY_SCHED_GROUPS
        self.max_tuples_per_mol, self.max_rounds_per_mol = [], []
        for l, L in enumerate(TUPLE_LEN):
            ro, sc = conflict_free_rounds(host[f"idx{l}"], host[f"tup_off{l}"], self.n_mols, L, self.sched_groups)
            host[f"round_off{l}"], host[f"sched{l}"] = ro, sc
            # largest molecule per level: sizes the energy kernel's shared-memory copy of one molecule's records
            self.max_tuples_per_mol.append(int(np.diff(host[f"tup_off{l}"]).max()) if self.n_mols else 0)
            self.max_rounds_per_mol.append(int(np.diff(ro).max()) if self.n_mols else 0)
        # bonded graph CSR by destination
        src, dst = g.edges(etype="n1_edge")
        src = src.detach().cpu().numpy().astype(np.int64)
        dst = dst.detach().cpu().numpy().astype(np.int64)
        self.n_edges = len(src)
        order = np.lexsort((src, dst))               # sort by dst, then src
        s_src, s_dst = src[order], dst[order]
        host["indptr"] = _offsets(np.bincount(s_dst, minlength=n_atoms)) if n_atoms else np.zeros(1, np.int32)
        host["esrc"] = s_src.astype(np.int32)
        # reverse edge position: edge (u -> v) at position e; reverse (v -> u) found by key lookup
        key = s_dst * max(n_atoms, 1) + s_src        # sorted ascending by construction
        rkey = s_src * max(n_atoms, 1) + s_dst
        pos = np.searchsorted(key, rkey)
        if len(key) and (pos.max() >= len(key) or not np.array_equal(key[np.minimum(pos, len(key) - 1)], rkey)):
            raise ValueError("bonded edges must be symmetric (both directions of every bond; reference "
                             "data/Molecule.py:465-472)")
        host["erev"] = pos.astype(np.int32)
        deg = np.diff(host["indptr"])
        if n_atoms and deg.min() == 0:
            raise ValueError("every atom must be part of a bond (zero in-degree atom; reference data/Molecule.py:470)")
        self.max_degree = int(deg.max()) if n_atoms else 0
        self.host = host
        self._names = list(host.keys())
        assert tuple(self._names) == self.TABLE_ORDER
        self._sizes = [host[k].size for k in self._names]
        # every table starts on a 16-byte boundary of the staging buffer (the energy kernel fetches a torsion's four atom
        # indices with one 16-byte load)
        self._offsets, total = [], 0
        for s in self._sizes:
            self._offsets.append(total)
            total += (int(s) + 3) // 4 * 4
        # one (pinned) staging buffer so that the whole pack moves with a single H2D copy
        stage = torch.zeros(total, dtype=torch.int32)
        for k, s, off in zip(self._names, self._sizes, self._offsets):
            stage[off:off + s] = torch.from_numpy(host[k].reshape(-1))
        if torch.cuda.is_available():
            try:
                stage = stage.pin_memory()
            except RuntimeError:  # pragma: no cover
                pass
        self._stage = stage
        self.bytes = total * 4
        self._dev: Dict[str, torch.Tensor] = {}
        self._materialise(self.device)

    def _materialise(self, device):
        self.device = torch.device(device)
        flat = self._stage.to(self.device, non_blocking=True)
        self._flat = flat
        self._dev = {}
        for k, s, off in zip(self._names, self._sizes, self._offsets):
            self._dev[k] = flat[off:off + s].view(self.host[k].shape)

    def to(self, device) -> "PackedBatch":
        """Copy of the pack on another device (one H2D transfer of the staging buffer)."""
        import copy as _copy
        p = _copy.copy(self)
        if self._stage is None:                 # assembled on the device: move the flat buffer itself
            if torch.device(device) == self._flat.device:
                return self
            flat = self._flat.to(device)
            return PackedBatch.from_device(flat, dict(zip(self._names, self._offsets)), dict(zip(self._names, self._sizes)),
                                           dict(n_atoms=self.n_atoms, n_mols=self.n_mols, n_edges=self.n_edges, n_tuples=self.n_tuples,
                                                max_atoms_per_mol=self.max_atoms_per_mol, max_degree=self.max_degree,
                                                max_tuples_per_mol=self.max_tuples_per_mol, max_rounds_per_mol=self.max_rounds_per_mol))
        p._materialise(device)
        return p

    def signature(self):
        """Everything the host bakes into kernel arguments / grids: two packs with equal signatures can share a
        captured CUDA graph (training.Trainer)."""
        return (self.n_atoms, self.n_mols, self.n_edges, tuple(self.n_tuples), self.max_atoms_per_mol, self.max_degree,
                tuple(self._sizes), tuple(self.max_tuples_per_mol), tuple(self.max_rounds_per_mol))

    def copy_from(self, other: "PackedBatch"):
        """Overwrite the device tables in place with another pack of the same signature (one async H2D copy)."""
        if other.signature() != self.signature():
            raise ValueError("PackedBatch.copy_from: signatures differ")
        src = other._stage if (other._flat.device != self._flat.device and other._stage is not None) else other._flat
        self._flat.copy_(src, non_blocking=True)
        self.host = other.host

    def __getitem__(self, k) -> torch.Tensor:
        return self._dev[k]

    def ptr(self, k) -> int:
        t = self._dev[k]
        return t.data_ptr() if t.numel() else 0


def get_pack(g) -> PackedBatch:
    """Cached PackedBatch of a graph (built on first use)."""
    dev = None
    for d in (g.nodes["n1"].data,):
        for v in d.values():
            dev = v.device
            break
    p = getattr(g, "_pack_cache", None)
    if p is None or p.device != dev:
        p = PackedBatch(g, device=dev)
        try:
            g._pack_cache = p
        except Exception:  # pragma: no cover (foreign graph objects that forbid attributes)
            pass
    return p
