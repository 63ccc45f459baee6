"""Packed dataset, collate and prefetching loader for the training hot path.

What this replaces in the reference (host side, around `model(g)`):

    data/Dataset.py:125             dgl.save_graphs storage: one pickled DGL graph per molecule
    utils/dgl_utils.py:132-171      set_number_confs: random sub-sampling / padding ('is_dummy') of conformations
    utils/dgl_utils.py:11-60        batch: deep copies + per-type concatenation + index offsets
    data/GraphDataLoader.py:23-73   collate_fn: conf_strategy -> n_confs, set_number_confs per graph, batch

At > 4,000 molecules/s per GPU the collate has ~6 ms per 32-molecule batch; per-graph Python objects and deep copies
do not fit in that.  Here a dataset is a handful of flat arrays (concatenated per node type, plus offsets) that can be
memory-mapped, a batch is assembled with a few slice-concatenations, the index tables of `pack.PackedBatch` are built in
the same call, and `PrefetchLoader` runs collate + pinning on a worker thread so the step only sees the H2D copy.

Semantics kept from the reference (checked against `tests/golden/ragged_confs.npz`, produced by the reference's own
`set_number_confs` / `batch` / `MolwiseLoss`):
  * n_confs of a batch: int -> min(int, max over the batch); 'min' | 'max' | 'all' | 'mean' (GraphDataLoader.py:52-66);
  * a molecule with MORE conformations keeps a random subset, one with FEWER repeats its last conformation and the
    padding is flagged in g.nodes['g'].data['is_dummy'] (B, n_confs) -- the loss ignores flagged conformations;
  * per-type concatenation in batch order, `idxs` and edges shifted by the atom offset, per-molecule counts kept.
The random subset is drawn from a numpy Generator (the reference uses torch.randperm on the global RNG): same
distribution, different stream.
"""
from __future__ import annotations

import json
import os
from typing import Dict, Iterable, Iterator, List, Optional, Sequence, Union

import numpy as np
import torch

from .graph import NTYPES, MolGraph
from .pack import ENERGY_SCHED_GROUPS as ENERGY_GROUPS

_CONF_FIELDS_G = ("energy",)        # substring match, as the reference does ('energy' in feat)
_CONF_FIELDS_N1 = ("gradient",)


def _n_confs_of(g: MolGraph) -> int:
    return int(g.nodes["n1"].data["xyz"].shape[1]) if "xyz" in g.nodes["n1"].data else 0


def batch_n_confs(confs: Sequence[int], conf_strategy: Union[str, int]) -> int:
    """Number of conformations of a batch (reference data/GraphDataLoader.py:52-66)."""
    if isinstance(conf_strategy, (int, np.integer)):
        return int(min(int(conf_strategy), max(confs)))
    if conf_strategy == "min":
        return int(min(confs))
    if conf_strategy in ("max", "all"):
        return int(max(confs))
    if conf_strategy == "mean":
        return int(np.mean(confs))
    raise ValueError(f"Unknown conf_strategy: {conf_strategy}")


def conformation_indices(present: int, wanted: int, rng: Optional[np.random.Generator]) -> np.ndarray:
    """Which stored conformations fill the `wanted` slots (reference utils/dgl_utils.py:146-160)."""
    if present == wanted:
        return np.arange(present)
    if present > wanted:
        rng = np.random.default_rng() if rng is None else rng
        return rng.permutation(present)[:wanted]
    return np.concatenate((np.arange(present), np.full(wanted - present, present - 1, dtype=np.int64)))


def set_number_confs(g: MolGraph, num_confs: int, rng: Optional[np.random.Generator] = None) -> MolGraph:
    """Shallow copy of a single-molecule graph with exactly `num_confs` conformations and `is_dummy` (1, num_confs)."""
    present = _n_confs_of(g)
    if present == 0:
        return g
    idx = torch.from_numpy(conformation_indices(present, num_confs, rng))
    out = MolGraph({nt: g.num_nodes(nt) for nt in g.ntypes}, g._src, g._dst,
                   {nt: g.batch_num_nodes(nt) for nt in g.ntypes})
    for nt in g.ntypes:
        for k, v in g.nodes[nt].data.items():
            if nt == "n1" and (k == "xyz" or any(f in k for f in _CONF_FIELDS_N1)):
                v = v[:, idx]
            elif nt == "g" and any(f in k for f in _CONF_FIELDS_G):
                v = v[:, idx]
                if torch.isnan(v).any():
                    raise RuntimeError(f"Found nan in {k} after setting number of conformations to {num_confs}")
            out.nodes[nt].data[k] = v
    dummy = torch.zeros((1, num_confs), dtype=torch.float32)
    if present < num_confs:
        dummy[0, present:] = 1.0
    out.nodes["g"].data["is_dummy"] = dummy
    return out


def collate(graphs: Sequence[MolGraph], conf_strategy: Union[str, int] = "mean",
            rng: Optional[np.random.Generator] = None, build_pack: bool = True) -> MolGraph:
    """List of single-molecule graphs -> one batched graph (reference collate_fn), index tables attached."""
    from .graph import batch
    from .pack import get_pack
    confs = [_n_confs_of(g) for g in graphs]
    if any(confs):
        n = batch_n_confs(confs, conf_strategy)
        graphs = [set_number_confs(g, n, rng) for g in graphs]
    bg = batch(graphs)
    if "is_dummy" in bg.nodes["g"].data:
        bg.nodes["g"].data["n_valid"] = (bg.nodes["g"].data["is_dummy"] == 0).sum(dim=1).to(torch.int32)
    if build_pack:
        get_pack(bg)
    return bg


# --------------------------------------------------------------------------------------------------
# flat storage
# --------------------------------------------------------------------------------------------------
def _gather_rows(off: np.ndarray, sel: np.ndarray):
    """(rows, counts): the concatenation of range(off[i], off[i+1]) for i in sel, and the length of each range."""
    starts = off[sel]
    counts = off[sel + 1] - starts
    total = int(counts.sum())
    rows = np.arange(total, dtype=np.int64) + np.repeat(starts - (np.cumsum(counts) - counts), counts)
    return rows, counts


class PackedDataset:
    """Many molecules as flat arrays: per node type and field one concatenated array plus element offsets.

    Conformation-dependent fields (xyz, *energy*, *gradient*) have a molecule-dependent second dimension, so they are
    stored flattened with their own offsets; everything else is concatenated along dim 0.  `save` writes one `.npy`
    per array and an `index.json`; `load(..., mmap=True)` maps them read-only, so a training job touches only the pages
    of the molecules it samples.
    """

    def __init__(self, arrays: Dict[str, np.ndarray], meta: dict):
        self.arrays, self.meta = arrays, meta
        self.n = int(meta["n_molecules"])
        # base-class views (np.memmap -> ndarray, no copy) and Python-int offsets of the ragged fields for collate()
        self._nd = {k: np.asarray(v) for k, v in arrays.items()}
        self._conf_fields = {tuple(f.split(".", 1)) for f in meta["conf_fields"]}
        self._foff = {f: self._nd[f"foff.{f[0]}.{f[1]}"].tolist() for f in self._conf_fields}

    def __len__(self) -> int:
        return self.n

    # ---- construction ------------------------------------------------------------------------
    @classmethod
    def from_graphs(cls, graphs: Sequence[MolGraph], dsnames: Optional[Sequence[str]] = None) -> "PackedDataset":
        n = len(graphs)
        if n == 0:
            raise ValueError("empty dataset")
        arrays: Dict[str, np.ndarray] = {}
        counts = {nt: np.array([g.num_nodes(nt) for g in graphs], dtype=np.int64) for nt in NTYPES}
        for nt in NTYPES:
            arrays[f"off.{nt}"] = np.concatenate(([0], np.cumsum(counts[nt]))).astype(np.int64)
        arrays["confs"] = np.array([_n_confs_of(g) for g in graphs], dtype=np.int64)
        ecounts = np.array([g.num_edges() for g in graphs], dtype=np.int64)
        arrays["off.edges"] = np.concatenate(([0], np.cumsum(ecounts))).astype(np.int64)
        arrays["edges.src"] = np.concatenate([g._src.numpy() for g in graphs]).astype(np.int32)
        arrays["edges.dst"] = np.concatenate([g._dst.numpy() for g in graphs]).astype(np.int32)
        fields = {nt: sorted(graphs[0].nodes[nt].data.keys()) for nt in NTYPES}
        conf_fields = []
        for nt in NTYPES:
            for k in fields[nt]:
                parts = [g.nodes[nt].data[k].numpy() for g in graphs]
                if cls._is_conf_field(nt, k):
                    conf_fields.append(f"{nt}.{k}")
                    arrays[f"data.{nt}.{k}"] = np.concatenate([p.reshape(-1) for p in parts])
                    arrays[f"foff.{nt}.{k}"] = np.concatenate(([0], np.cumsum([p.size for p in parts]))).astype(np.int64)
                else:
                    arrays[f"data.{nt}.{k}"] = np.concatenate(parts, axis=0)
        meta = {"n_molecules": n, "fields": fields, "conf_fields": conf_fields,
                "dsnames": list(dsnames) if dsnames is not None else [""] * n}
        return cls(arrays, meta)

    @staticmethod
    def _is_conf_field(nt: str, k: str) -> bool:
        return (nt == "n1" and (k == "xyz" or any(f in k for f in _CONF_FIELDS_N1))) or \
               (nt == "g" and any(f in k for f in _CONF_FIELDS_G))

    def save(self, path: str) -> None:
        os.makedirs(path, exist_ok=True)
        for name, a in self.arrays.items():
            np.save(os.path.join(path, name + ".npy"), a)
        with open(os.path.join(path, "index.json"), "w") as f:
            json.dump({**self.meta, "arrays": sorted(self.arrays)}, f)

    @classmethod
    def load(cls, path: str, mmap: bool = True) -> "PackedDataset":
        with open(os.path.join(path, "index.json")) as f:
            meta = json.load(f)
        arrays = {name: np.load(os.path.join(path, name + ".npy"), mmap_mode="r" if mmap else None)
                  for name in meta.pop("arrays")}
        return cls(arrays, meta)

    # ---- access ------------------------------------------------------------------------------
    def molecule(self, i: int) -> MolGraph:
        """Molecule i as a single-molecule MolGraph (views into the flat arrays where possible)."""
        return self.collate([i], conf_strategy="max", build_pack=False, keep_is_dummy=False)

    def dsname(self, i: int) -> str:
        return self.meta["dsnames"][i]

    def collate(self, indices: Sequence[int], conf_strategy: Union[str, int] = "mean",
                rng: Optional[np.random.Generator] = None, build_pack: bool = True, keep_is_dummy: bool = True) -> MolGraph:
        """Batched graph of the molecules `indices` (same result as `collate([molecule(i) ...])`, without building
        per-molecule objects): slices of the flat arrays are concatenated per field, `idxs` / edges get the atom offset of
        the batch, conformations are sub-sampled / padded per molecule."""
        from .pack import get_pack
        A = self._nd
        idx = [int(i) for i in indices]
        if idx and (min(idx) < 0 or max(idx) >= self.n):
            raise IndexError(f"molecule index out of range [0, {self.n}): {[i for i in idx if not 0 <= i < self.n]}")
        ia = np.asarray(idx, dtype=np.int64)
        # per node type: the flat row numbers of the selected molecules, in batch order (one fancy-index gather per field
        # instead of a Python loop over molecules: the loader threads spend their time in numpy, not in the interpreter)
        rows, counts = {}, {}
        for nt in NTYPES:
            rows[nt], counts[nt] = _gather_rows(A[f"off.{nt}"], ia)
        atom_shift = np.cumsum(counts["n1"]) - counts["n1"]
        confs = A["confs"][ia].tolist()
        n_confs = batch_n_confs(confs, conf_strategy) if any(confs) else 0
        csel = [conformation_indices(c, n_confs, rng) if n_confs else None for c in confs]
        identity = all(c == n_confs for c in confs)          # every stored conformation, in order: plain copies
        erows, ecounts = _gather_rows(A["off.edges"], ia)
        eshift = np.repeat(atom_shift, ecounts).astype(np.int32)
        src = A["edges.src"][erows] + eshift
        dst = A["edges.dst"][erows] + eshift
        g = MolGraph({nt: int(counts[nt].sum()) for nt in NTYPES}, torch.from_numpy(src), torch.from_numpy(dst),
                     {nt: torch.from_numpy(counts[nt].copy()) for nt in NTYPES})
        conf_fields = self._conf_fields
        for nt in NTYPES:
            cnt = counts[nt].tolist()
            n_rows = int(counts[nt].sum())
            for k in self.meta["fields"][nt]:
                data = A[f"data.{nt}.{k}"]
                if (nt, k) in conf_fields:
                    fo = self._foff[(nt, k)]
                    tail = (3,) if nt == "n1" else ()
                    if identity:
                        arr = np.concatenate([data[fo[i]:fo[i + 1]] for i in idx]).reshape((n_rows, n_confs) + tail)
                    else:
                        arr = np.concatenate([data[fo[i]:fo[i + 1]].reshape((cnt[j], confs[j]) + tail)[:, csel[j]]
                                              for j, i in enumerate(idx)], axis=0)
                else:
                    arr = data[rows[nt]]
                    if k == "idxs":
                        arr += np.repeat(atom_shift, counts[nt]).reshape((-1,) + (1,) * (arr.ndim - 1))
                g.nodes[nt].data[k] = torch.from_numpy(arr)
        if n_confs and keep_is_dummy:
            dummy = np.zeros((len(idx), n_confs), dtype=np.float32)
            for j, c in enumerate(confs):
                dummy[j, min(c, n_confs):] = 1.0
            g.nodes["g"].data["is_dummy"] = torch.from_numpy(dummy)
            g.nodes["g"].data["n_valid"] = torch.from_numpy((dummy == 0).sum(axis=1).astype(np.int32))
        if build_pack:
            get_pack(g)
        return g


# --------------------------------------------------------------------------------------------------
# loader
# --------------------------------------------------------------------------------------------------
class PrefetchLoader:
    """Iterates batched, pinned host graphs built on a worker thread, `depth` batches ahead of the consumer.

    `batches` yields index lists (a sampler); every rank of a data-parallel job passes its own shard
    (`shard_indices(len(ds), rank, world)`).  The consumer hands each graph to `Trainer.step(host_graph)`, whose H2D copy
    into the static device buffers is the only per-step host work left on the critical path.
    """

    def __init__(self, dataset: PackedDataset, batches: Iterable[Sequence[int]], conf_strategy: Union[str, int] = "mean",
                 seed: int = 0, depth: int = 3, pin: bool = True, param_weight: Optional[float] = None,
                 param_weights_by_dataset: Optional[Dict[str, float]] = None, workers: int = 1):
        """`param_weight` / `param_weights_by_dataset`: MolwiseLoss's classical-parameter weights (reference
        training/loss.py:72-76) resolved per molecule here and shipped as g.nodes['g'].data['param_weight'] (B,), an
        ordinary input tensor (a captured step sees each batch's weights; dataset names never reach the device).

        `workers` threads build batches concurrently.  One is normally enough: a 32-peptide batch with its index tables
        takes ~3 ms on one thread (vectorised gathers + the C++ round scheduler) against a ~6 ms device step, and what is
        left is interpreter time, which more threads do not speed up.  Batches are delivered in sampler order and batch k
        draws its conformation sub-sample from its own stream `default_rng([seed, k])`, so the data seen do not depend on
        `workers` or on thread timing."""
        self.dataset, self.batches, self.conf_strategy = dataset, batches, conf_strategy
        self.seed, self.depth, self.pin = seed, max(1, depth), pin and torch.cuda.is_available()
        self.param_weight, self.param_weights_by_dataset = param_weight, dict(param_weights_by_dataset or {})
        self.workers = max(1, int(workers))

    def _build(self, k: int, idx: Sequence[int]) -> MolGraph:
        g = self.dataset.collate(idx, self.conf_strategy, np.random.default_rng([self.seed, k]))
        if self.param_weight is not None:
            w = [self.param_weights_by_dataset.get(self.dataset.dsname(i), self.param_weight) for i in idx]
            g.nodes["g"].data["param_weight"] = torch.tensor(w, dtype=torch.float32)
        if self.pin:
            g = g.pin_memory()
        g.dsnames = [self.dataset.dsname(i) for i in idx]
        return g

    def __iter__(self) -> Iterator[MolGraph]:
        from collections import deque
        from concurrent.futures import ThreadPoolExecutor
        pool = ThreadPoolExecutor(max_workers=self.workers, thread_name_prefix="grappa-loader")
        pending: "deque" = deque()
        it = enumerate(self.batches)
        try:
            while True:
                while len(pending) < self.depth:
                    nxt = next(it, None)
                    if nxt is None:
                        break
                    pending.append(pool.submit(self._build, nxt[0], list(nxt[1])))
                if not pending:
                    return
                yield pending.popleft().result()        # sampler order; a worker's exception surfaces here
        finally:
            for f in pending:
                f.cancel()
            pool.shutdown(wait=True, cancel_futures=True)


def shard_indices(n: int, rank: int, world: int) -> List[int]:
    """Molecule i -> rank i mod world (independent molecules, no collective)."""
    return list(range(rank, n, world))


def shard_conformations(g: MolGraph, rank: int, world: int) -> MolGraph:
    """This rank's slice of the conformation axis of a (batched) graph -- the sharding of the energy / force sweep when
    there are few molecules and many conformations: conformations [C*rank/world, C*(rank+1)/world)
    of xyz (N, C, 3), of every `*gradient*` field and of every (B, C) `*energy*` field, repacked contiguously once
    (xyz is atom-major, so a conformation slice of the original is a strided view).  Topology, indices and parameters
    are shared with `g`; concatenating the shards' results along the conformation axis gives the unsharded result, so
    the path needs no collective."""
    C = _n_confs_of(g)
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside [0, {world})")
    lo, hi = C * rank // world, C * (rank + 1) // world
    out = MolGraph({nt: g.num_nodes(nt) for nt in g.ntypes}, g._src, g._dst,
                   {nt: g.batch_num_nodes(nt) for nt in g.ntypes})
    for nt in g.ntypes:
        for k, v in g.nodes[nt].data.items():
            if PackedDataset._is_conf_field(nt, k) or (nt == "g" and k == "is_dummy"):
                v = v[:, lo:hi].contiguous()
            out.nodes[nt].data[k] = v
    out._pack_cache = g._pack_cache            # same topology: the index tables are shared
    return out


def batch_sampler(indices: Sequence[int], batch_size: int, rng: Optional[np.random.Generator] = None,
                  drop_last: bool = True) -> Iterator[List[int]]:
    """Shuffled fixed-size batches of `indices` (one epoch)."""
    order = np.array(indices)
    if rng is not None:
        order = rng.permutation(order)
    for s in range(0, len(order), batch_size):
        b = order[s:s + batch_size].tolist()
        if len(b) < batch_size and drop_last:
            return
        yield b


# --------------------------------------------------------------------------------------------------
# device-resident dataset: batches assembled on the GPU
# --------------------------------------------------------------------------------------------------
class DeviceDataset:
    """A `PackedDataset` resident in HBM whose batches are assembled ON THE DEVICE by one kernel launch.

    Replaces, per batch, the reference's host-side collate (data/GraphDataLoader.py:23-73), `set_number_confs`
    (utils/dgl_utils.py:132-171: conformation sub-sampling / padding) and `batch` (utils/dgl_utils.py:11-60: per-type
    concatenation, `idxs += atom offset`) AND the host construction of `pack.PackedBatch`'s index tables: every one of
    those tables (inverse incidence CSR, bonded-graph CSR, reverse-edge table, conflict-free schedule of the energy
    kernel) is a concatenation of per-molecule tables shifted by a per-molecule offset, so they are computed ONCE per
    molecule when the dataset is uploaded and merely gathered + shifted per batch (`grappa_b200_collate`).

    Per batch the host does O(B) work: a few cumulative sums over the per-molecule counts, the conformation selection
    (B x n_confs integers, same rules as the host collate) and the job descriptors -- one pinned buffer of a few KB, one
    H2D copy, one kernel launch.  The result is the same batched graph (on the device) with its `PackedBatch`, field for
    field and bit for bit what `PackedDataset.collate(...).to(device)` gives (tests/test_dataset_gpu.py).
    """

    _TABLES = ("idx", "inv_ptr", "inv_ent", "sched")      # per level

    def __init__(self, ds: PackedDataset, device="cuda"):
        from . import _lib
        from .pack import LEVELS, TUPLE_LEN, PackedBatch
        self.ds, self.device = ds, torch.device(device)
        if self.device.type != "cuda":
            raise _lib.GrappaB200Error("DeviceDataset needs a CUDA device (the host collate is PackedDataset.collate)")
        A = ds._nd
        n = ds.n
        self.n = n
        self.counts = {nt: np.diff(A[f"off.{nt}"]).astype(np.int64) for nt in NTYPES}
        self.confs = A["confs"].astype(np.int64)
        self.ecounts = np.diff(A["off.edges"]).astype(np.int64)
        dev = self.device
        up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        # ---- node fields and edges, as stored
        self.field = {}          # (nt, k) -> (device tensor, row_words, kind, torch dtype, trailing shape)
        self.off_dev = {nt: up(A[f"off.{nt}"].astype(np.int64)) for nt in NTYPES}
        self.eoff_dev = up(A["off.edges"].astype(np.int64))
        self.confs_dev = up(self.confs.astype(np.int32))
        self.foff_dev = {}
        for nt in NTYPES:
            for k in ds.meta["fields"][nt]:
                arr = A[f"data.{nt}.{k}"]
                t = torch.from_numpy(np.ascontiguousarray(arr))
                if (nt, k) in ds._conf_fields:
                    tail = 3 if nt == "n1" else 1
                    self.field[(nt, k)] = (t.to(dev), tail, 3, t.dtype, (3,) if nt == "n1" else ())
                    self.foff_dev[(nt, k)] = up(A[f"foff.{nt}.{k}"].astype(np.int64))
                else:
                    trailing = tuple(arr.shape[1:])
                    words = int(np.prod(trailing, dtype=np.int64)) * t.element_size() // 4
                    if t.element_size() not in (4, 8):
                        raise ValueError(f"field {nt}.{k}: element size {t.element_size()} is not supported on the device path")
                    kind = 2 if (k == "idxs" and t.dtype == torch.int64) else (1 if k == "idxs" else 0)
                    self.field[(nt, k)] = (t.to(dev), max(words, 1), kind, t.dtype, trailing)
        self.esrc_dev, self.edst_dev = up(A["edges.src"].astype(np.int32)), up(A["edges.dst"].astype(np.int32))
        # ---- per-molecule index tables, built once with the host code of pack.PackedBatch
        tabs: Dict[str, list] = {}
        self.rounds = {l: np.zeros(n, dtype=np.int64) for l in range(4)}
        self.max_degree = np.zeros(n, dtype=np.int64)
        for i in range(n):
            p = PackedBatch(ds.molecule(i), device="cpu")
            h = p.host
            na = int(self.counts["n1"][i])
            for l in range(4):
                tabs.setdefault(f"idx{l}", []).append(h[f"idx{l}"].reshape(-1))
                tabs.setdefault(f"inv_ptr{l}", []).append(h[f"inv_ptr{l}"][:na])
                tabs.setdefault(f"inv_ent{l}", []).append(h[f"inv_ent{l}"].reshape(-1))
                tabs.setdefault(f"sched{l}", []).append(h[f"sched{l}"].reshape(-1))
                self.rounds[l][i] = h[f"sched{l}"].shape[0]
            tabs.setdefault("indptr", []).append(h["indptr"][:na])
            tabs.setdefault("esrc", []).append(h["esrc"])
            tabs.setdefault("erev", []).append(h["erev"])
            self.max_degree[i] = p.max_degree
        self.table, self.table_off = {}, {}
        for name, parts in tabs.items():
            sizes = np.array([len(x) for x in parts], dtype=np.int64)
            self.table[name] = up(np.concatenate(parts).astype(np.int32) if len(parts) else np.zeros(0, np.int32))
            self.table_off[name] = np.concatenate(([0], np.cumsum(sizes)))
        self._table_rows = {}      # name -> (row_words, per-molecule row offsets on the device)
        for l, L in enumerate(TUPLE_LEN):
            self._table_rows[f"idx{l}"] = (L, self.off_dev[LEVELS[l]])
            self._table_rows[f"inv_ptr{l}"] = (1, self.off_dev["n1"])
            self._table_rows[f"inv_ent{l}"] = (1, up(self.table_off[f"inv_ent{l}"]))
            self._table_rows[f"sched{l}"] = (ENERGY_GROUPS, up(np.concatenate(([0], np.cumsum(self.rounds[l])))))
        self._table_rows["indptr"] = (1, self.off_dev["n1"])
        self._table_rows["esrc"] = (1, self.eoff_dev)
        self._table_rows["erev"] = (1, self.eoff_dev)
        self._names = None

    def __len__(self):
        return self.n

    # ---------------------------------------------------------------------------------------------
    def collate(self, indices: Sequence[int], conf_strategy: Union[str, int] = "mean",
                rng: Optional[np.random.Generator] = None) -> MolGraph:
        """Batched graph of the molecules `indices` on the device, with its PackedBatch attached (see class docstring)."""
        import ctypes as C
        from . import _lib
        from ._lib_ops import COLLATE_MAX_JOBS, CollateArgs
        from .pack import LEVELS, TUPLE_LEN, PackedBatch
        ia = np.asarray([int(i) for i in indices], dtype=np.int64)
        if len(ia) and (ia.min() < 0 or ia.max() >= self.n):
            raise IndexError(f"molecule index out of range [0, {self.n})")
        B = len(ia)
        dev = self.device
        cnt = {nt: self.counts[nt][ia] for nt in NTYPES}
        off = {nt: np.concatenate(([0], np.cumsum(cnt[nt]))).astype(np.int32) for nt in NTYPES}
        ecnt = self.ecounts[ia]
        eoff = np.concatenate(([0], np.cumsum(ecnt))).astype(np.int32)
        confs = self.confs[ia].tolist()
        n_confs = batch_n_confs(confs, conf_strategy) if any(confs) else 0
        csel = np.stack([conformation_indices(c, n_confs, rng) for c in confs]).astype(np.int32) if n_confs else np.zeros((B, 0), np.int32)
        N = int(off["n1"][-1])
        E = int(eoff[-1])
        # ---- small per-batch tables computed on the host (O(B)): they travel in the upload buffer
        small: Dict[str, np.ndarray] = {"mol": ia.astype(np.int32), "csel": csel.reshape(-1)}
        for nt in NTYPES:
            small[f"off.{nt}"] = off[nt]
        small["eoff"] = eoff
        roff = {}
        for l, L in enumerate(TUPLE_LEN):
            roff[l] = np.concatenate(([0], np.cumsum(self.rounds[l][ia]))).astype(np.int32)
            small[f"round_off{l}"] = roff[l]
            small[f"entoff{l}"] = (off[LEVELS[l]].astype(np.int64) * L).astype(np.int32)      # (tuple, slot) entry offsets
            small[f"inv_end{l}"] = np.array([int(off[LEVELS[l]][-1]) * L], dtype=np.int32)
        small["ind_end"] = np.array([E], dtype=np.int32)
        if n_confs:
            dummy = np.zeros((B, n_confs), dtype=np.float32)
            for j, c in enumerate(confs):
                dummy[j, min(c, n_confs):] = 1.0
            small["is_dummy"] = dummy.view(np.int32).reshape(-1)
            small["n_valid"] = (dummy == 0).sum(axis=1).astype(np.int32)
        # ---- output buffers: one int32 pack buffer in PackedBatch's layout, one buffer for the graph fields
        pack_sizes = {"atom_off": B + 1}
        for l, L in enumerate(TUPLE_LEN):
            T = int(off[LEVELS[l]][-1])
            pack_sizes.update({f"idx{l}": T * L, f"tup_off{l}": B + 1, f"inv_ptr{l}": N + 1, f"inv_ent{l}": T * L})
        for l in range(4):
            pack_sizes.update({f"round_off{l}": B + 1, f"sched{l}": int(roff[l][-1]) * ENERGY_GROUPS})
        pack_sizes.update({"indptr": N + 1, "esrc": E, "erev": E})
        names = PackedBatch.TABLE_ORDER
        p_off, total = {}, 0
        for k in names:
            p_off[k] = total
            total += (pack_sizes[k] + 3) // 4 * 4
        pack_flat = torch.zeros(total, dtype=torch.int32, device=dev)
        # graph fields
        f_off, f_total, f_meta = {}, 0, {}
        for (nt, k), (src, words, kind, dtype, trailing) in self.field.items():
            rows = int(off[nt][-1])
            w = words * n_confs if kind == 3 else words
            f_off[(nt, k)] = f_total
            f_meta[(nt, k)] = (rows, w)
            f_total += (rows * w + 3) // 4 * 4
        f_off["src"], f_off["dst"] = f_total, f_total + (E + 3) // 4 * 4
        f_total += 2 * ((E + 3) // 4 * 4)
        field_flat = torch.empty(f_total, dtype=torch.int32, device=dev)
        # ---- upload buffer: [small tables][CollateArgs]
        s_off, s_total = {}, 0
        for k, v in small.items():
            s_off[k] = s_total
            s_total += (v.size + 3) // 4 * 4
        args_words = (C.sizeof(CollateArgs) + 3) // 4
        s_total = (s_total + 1) // 2 * 2                       # 8-byte alignment of the argument block
        up_host = torch.zeros(s_total + args_words, dtype=torch.int32).pin_memory()
        up_np = up_host.numpy()
        for k, v in small.items():
            up_np[s_off[k]:s_off[k] + v.size] = v.reshape(-1)
        up_dev = torch.empty_like(up_host, device=dev)
        base = up_dev.data_ptr()
        sp = lambda k: base + 4 * s_off[k]
        a = CollateArgs()
        a.B, a.mol = B, sp("mol")
        jobs = []

        def job(src, dst, src_off, dst_off, rows, words, kind, add=0, n_out=0, csel_p=0):
            jobs.append((src, dst, src_off, dst_off, add, rows, words, kind, n_out, csel_p))
        pp = lambda k: pack_flat.data_ptr() + 4 * p_off[k]
        fp = lambda k: field_flat.data_ptr() + 4 * f_off[k]
        # host-computed tables -> pack buffer (plain copies)
        job(sp("off.n1"), pp("atom_off"), 0, 0, B + 1, 1, 4)
        for l in range(4):
            job(sp(f"off.{LEVELS[l]}"), pp(f"tup_off{l}"), 0, 0, B + 1, 1, 4)
            job(sp(f"round_off{l}"), pp(f"round_off{l}"), 0, 0, B + 1, 1, 4)
            job(sp(f"inv_end{l}"), pp(f"inv_ptr{l}") + 4 * N, 0, 0, 1, 1, 4)
        job(sp("ind_end"), pp("indptr") + 4 * N, 0, 0, 1, 1, 4)
        # per-molecule tables, shifted
        for l, L in enumerate(TUPLE_LEN):
            lvl = LEVELS[l]
            T = int(off[lvl][-1])
            nr = int(roff[l][-1])
            job(self.table[f"idx{l}"].data_ptr(), pp(f"idx{l}"), self._table_rows[f"idx{l}"][1].data_ptr(), sp(f"off.{lvl}"), T, L, 1, sp("off.n1"))
            job(self.table[f"inv_ptr{l}"].data_ptr(), pp(f"inv_ptr{l}"), self._table_rows[f"inv_ptr{l}"][1].data_ptr(), sp("off.n1"), N, 1, 1, sp(f"entoff{l}"))
            # inverse entries: one row per (tuple, slot); their values are flat (tuple * L + slot) positions
            job(self.table[f"inv_ent{l}"].data_ptr(), pp(f"inv_ent{l}"), self._table_rows[f"inv_ent{l}"][1].data_ptr(), sp(f"entoff{l}"), T * L, 1, 1, sp(f"entoff{l}"))
            job(self.table[f"sched{l}"].data_ptr(), pp(f"sched{l}"), self._table_rows[f"sched{l}"][1].data_ptr(), sp(f"round_off{l}"), nr, ENERGY_GROUPS, 1, sp(f"off.{lvl}"))
        job(self.table["indptr"].data_ptr(), pp("indptr"), self._table_rows["indptr"][1].data_ptr(), sp("off.n1"), N, 1, 1, sp("eoff"))
        job(self.table["esrc"].data_ptr(), pp("esrc"), self._table_rows["esrc"][1].data_ptr(), sp("eoff"), E, 1, 1, sp("off.n1"))
        job(self.table["erev"].data_ptr(), pp("erev"), self._table_rows["erev"][1].data_ptr(), sp("eoff"), E, 1, 1, sp("eoff"))
        # graph edges and node fields
        job(self.esrc_dev.data_ptr(), fp("src"), self.eoff_dev.data_ptr(), sp("eoff"), E, 1, 1, sp("off.n1"))
        job(self.edst_dev.data_ptr(), fp("dst"), self.eoff_dev.data_ptr(), sp("eoff"), E, 1, 1, sp("off.n1"))
        for (nt, k), (src, words, kind, dtype, trailing) in self.field.items():
            rows = int(off[nt][-1])
            if kind == 3:
                job(src.data_ptr(), fp((nt, k)), self.foff_dev[(nt, k)].data_ptr(), sp(f"off.{nt}"), rows, words, 3, 0, n_confs, sp("csel"))
            else:
                job(src.data_ptr(), fp((nt, k)), self.off_dev[nt].data_ptr(), sp(f"off.{nt}"), rows, words, kind, sp("off.n1") if kind in (1, 2) else 0)
        return self._finish(a, jobs, up_host, up_dev, s_total, pack_flat, p_off, pack_sizes, field_flat, f_off, f_meta, off, eoff,
                            cnt, ia, n_confs, small, s_off, B, N, E, roff)

    def _finish(self, a, jobs, up_host, up_dev, s_total, pack_flat, p_off, pack_sizes, field_flat, f_off, f_meta, off, eoff, cnt, ia,
                n_confs, small, s_off, B, N, E, roff):
        import ctypes as C
        from . import _lib
        from ._lib_ops import COLLATE_MAX_JOBS, CollateArgs
        from .pack import LEVELS, TUPLE_LEN, PackedBatch
        dev = self.device
        # the inverse-entry job needs per-molecule OUTPUT offsets = tuple offsets * L: they are the `add.tupL` tables with the total appended
        fixed = []
        for (src, dst, src_off, dst_off, add, rows, words, kind, n_out, csel_p) in jobs:
            fixed.append([src, dst, src_off, dst_off, add, rows, words, kind, n_out, csel_p])
        if len(fixed) > COLLATE_MAX_JOBS:
            raise _lib.GrappaB200Error(f"DeviceDataset.collate: {len(fixed)} jobs > {COLLATE_MAX_JOBS} (too many node fields)")
        max_words = 1
        a.n_jobs = len(fixed)
        for i, (src, dst, src_off, dst_off, add, rows, words, kind, n_out, csel_p) in enumerate(fixed):
            j = a.job[i]
            j.src, j.dst, j.src_off, j.dst_off, j.add = src, dst, src_off or None, dst_off or None, add or None
            j.confs = self.confs_dev.data_ptr() if kind == 3 else None
            j.csel = csel_p or None
            j.row_words, j.kind, j.n_confs_out, j.n_rows = words, kind, n_out, rows
            max_words = max(max_words, rows * (words * n_out if kind == 3 else words))
        raw = (C.c_int32 * ((C.sizeof(CollateArgs) + 3) // 4)).from_buffer_copy(bytes(a).ljust(((C.sizeof(CollateArgs) + 3) // 4) * 4, b"\0"))
        up_host.numpy()[s_total:s_total + len(raw)] = np.frombuffer(raw, dtype=np.int32)
        up_dev.copy_(up_host, non_blocking=True)
        _lib.check(_lib.lib().grappa_b200_collate(up_dev.data_ptr() + 4 * s_total, a.n_jobs, max_words,
                                                  torch.cuda.current_stream().cuda_stream), "collate")
        # ---- views: graph
        src = field_flat[f_off["src"]:f_off["src"] + E]
        dst = field_flat[f_off["dst"]:f_off["dst"] + E]
        g = MolGraph({nt: int(off[nt][-1]) for nt in NTYPES}, src, dst, {nt: torch.from_numpy(cnt[nt].copy()) for nt in NTYPES})
        for (nt, k), (srct, words, kind, dtype, trailing) in self.field.items():
            rows, w = f_meta[(nt, k)]
            flat = field_flat[f_off[(nt, k)]:f_off[(nt, k)] + rows * w]
            shape = (rows, n_confs) + trailing if kind == 3 else (rows,) + trailing
            g.nodes[nt].data[k] = flat.view(dtype).view(shape)
        if n_confs:
            sv = lambda k, n_: up_dev[s_off[k]:s_off[k] + n_]
            g.nodes["g"].data["is_dummy"] = sv("is_dummy", B * n_confs).view(torch.float32).view(B, n_confs)
            g.nodes["g"].data["n_valid"] = sv("n_valid", B)
        g._collate_keepalive = (up_host, up_dev)         # the pinned source of the asynchronous upload must outlive it
        # ---- views: pack
        meta = dict(n_atoms=N, n_mols=B, n_edges=E, n_tuples=[int(off[LEVELS[l]][-1]) for l in range(4)],
                    max_atoms_per_mol=int(cnt["n1"].max()) if B else 0, max_degree=int(self.max_degree[ia].max()) if B else 0,
                    max_tuples_per_mol=[int(cnt[LEVELS[l]].max()) if B else 0 for l in range(4)],
                    max_rounds_per_mol=[int(self.rounds[l][ia].max()) if B else 0 for l in range(4)])
        g._pack_cache = PackedBatch.from_device(pack_flat, p_off, pack_sizes, meta)
        return g
