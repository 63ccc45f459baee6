"""Batched molecular graph container with the field protocol the hot path reads and writes.

The reference passes a `dgl.DGLHeteroGraph` between modules and communicates only through named
tensors on it (`g.nodes[ntype].data[name]`, reference src/grappa/models/grappa.py:111-132,
graph_attention.py:157-181, interaction_parameters.py:160-178,263-264, energy.py:99-145).  DGL has
no sm_100 build, so the B200 path does not depend on it: modules in this package accept ANY object
exposing the small duck-typed surface below -- a real DGL heterograph satisfies it, and so does
`MolGraph`, the light container defined here:

    g.ntypes                               list of node types ('g','n1','n2','n3','n4','n4_improper')
    g.num_nodes(ntype)                     int
    g.nodes[ntype].data                    dict name -> tensor (first dim == num_nodes(ntype))
    g.batch_num_nodes(ntype)               int64 tensor, per-molecule counts
    g.edges(etype='n1_edge')               (src, dst) int tensors, both bond directions
    g.to(device)                           shallow copy with tensors moved

Batching semantics follow reference src/grappa/utils/dgl_utils.py:11-60 (`batch`: per-type
concatenation in graph order, `idxs += atom offset`) and :63-82 (`unbatch`).
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch

LEVELS = ("n2", "n3", "n4", "n4_improper")
NTYPES = ("g", "n1") + LEVELS
TUPLE_LEN = {"n2": 2, "n3": 3, "n4": 4, "n4_improper": 4}


class _NodeView:
    __slots__ = ("data",)

    def __init__(self, store):
        self.data = store


class _Nodes:
    __slots__ = ("_g",)

    def __init__(self, g):
        self._g = g

    def __getitem__(self, ntype):
        return _NodeView(self._g._ndata[ntype])


class MolGraph:
    """Hetero graph of one or several molecules (atoms 'n1', tuples 'n2..n4_improper', global 'g')."""

    def __init__(self, num_nodes: Dict[str, int], edges_src: torch.Tensor, edges_dst: torch.Tensor,
                 batch_num_nodes: Dict[str, torch.Tensor] | None = None):
        self._num_nodes = dict(num_nodes)
        for nt in NTYPES:
            self._num_nodes.setdefault(nt, 0 if nt != "g" else 1)
        self._src = edges_src
        self._dst = edges_dst
        self._ndata: Dict[str, Dict[str, torch.Tensor]] = {nt: {} for nt in self._num_nodes}
        if batch_num_nodes is None:
            batch_num_nodes = {nt: torch.tensor([n], dtype=torch.int64) for nt, n in self._num_nodes.items()}
        self._batch_num_nodes = batch_num_nodes
        self._pack_cache = None  # device-side packed batch (see pack.py); dropped on .to()

    # ---- DGL-compatible surface --------------------------------------------------------------
    @property
    def ntypes(self) -> List[str]:
        return list(self._num_nodes.keys())

    @property
    def nodes(self):
        return _Nodes(self)

    def num_nodes(self, ntype: str) -> int:
        return self._num_nodes[ntype]

    def batch_num_nodes(self, ntype: str) -> torch.Tensor:
        return self._batch_num_nodes[ntype]

    def edges(self, etype: str = "n1_edge"):
        if etype not in ("n1_edge", ("n1", "n1_edge", "n1")):
            raise KeyError(etype)
        return self._src, self._dst

    def num_edges(self, etype: str = "n1_edge") -> int:
        return int(self._src.shape[0])

    @property
    def batch_size(self) -> int:
        return int(self._batch_num_nodes["g"].shape[0])

    @property
    def device(self):
        return self._src.device

    def to(self, device, non_blocking: bool = False):
        g = MolGraph.__new__(MolGraph)
        g._num_nodes = dict(self._num_nodes)
        g._src = self._src.to(device, non_blocking=non_blocking)
        g._dst = self._dst.to(device, non_blocking=non_blocking)
        g._ndata = {nt: {k: v.to(device, non_blocking=non_blocking) for k, v in d.items()}
                    for nt, d in self._ndata.items()}
        g._batch_num_nodes = dict(self._batch_num_nodes)  # counts stay on the host
        # a host-built pack (index tables, see pack.py) travels with the graph: one extra H2D copy
        g._pack_cache = self._pack_cache.to(device) if self._pack_cache is not None else None
        return g

    def cpu(self):
        return self.to("cpu")

    def pin_memory(self):
        g = MolGraph.__new__(MolGraph)
        g._num_nodes = dict(self._num_nodes)
        g._src = self._src.pin_memory()
        g._dst = self._dst.pin_memory()
        g._ndata = {nt: {k: v.pin_memory() for k, v in d.items()} for nt, d in self._ndata.items()}
        g._batch_num_nodes = dict(self._batch_num_nodes)
        g._pack_cache = self._pack_cache
        return g

    def host_bytes(self) -> int:
        """Bytes a host->device transfer of this graph moves (for bench.py's e2e accounting)."""
        n = self._src.numel() * self._src.element_size() + self._dst.numel() * self._dst.element_size()
        for d in self._ndata.values():
            for v in d.values():
                n += v.numel() * v.element_size()
        if self._pack_cache is not None:
            n += self._pack_cache.bytes
        return n


def graph_from_molecule(n_atoms: int, bonds, angles, propers, impropers, feats: Dict[str, torch.Tensor],
                        xyz: torch.Tensor | None = None) -> MolGraph:
    """One molecule -> MolGraph with the reference layout (data/Molecule.py:429-520, MolData.py:155-200).

    bonds (n_bonds, 2) ... impropers (n_imp, 4) are integer arrays of atom indices; both directions
    of every bond become 'n1_edge' edges ordered [all (b0->b1), then all (b1->b0)] as at
    Molecule.py:465-472.  `feats` are the per-atom input features; `xyz` is (n_atoms, n_confs, 3).
    """
    def as_idx(a, L):
        t = torch.as_tensor(a, dtype=torch.int64)
        return t.reshape(-1, L)

    b = as_idx(bonds, 2)
    src = torch.cat((b[:, 0], b[:, 1])).to(torch.int32)
    dst = torch.cat((b[:, 1], b[:, 0])).to(torch.int32)
    tup = {"n2": b, "n3": as_idx(angles, 3), "n4": as_idx(propers, 4), "n4_improper": as_idx(impropers, 4)}
    if len(torch.unique(b)) != n_atoms:
        raise AssertionError("Every atom must be part of a bond (reference data/Molecule.py:470)")
    g = MolGraph({"g": 1, "n1": n_atoms, **{k: int(v.shape[0]) for k, v in tup.items()}}, src, dst)
    for k, v in tup.items():
        g.nodes[k].data["idxs"] = v
    for k, v in feats.items():
        g.nodes["n1"].data[k] = torch.as_tensor(v, dtype=torch.float32)
    if xyz is not None:
        g.nodes["n1"].data["xyz"] = torch.as_tensor(xyz, dtype=torch.float32)
    return g


def as_molgraph(g) -> MolGraph:
    """Any graph object with the duck-typed surface of the module docstring (e.g. a `dgl.DGLHeteroGraph` built by the
    reference's Molecule.to_dgl, data/Molecule.py:429-520) -> MolGraph sharing its tensors.  A MolGraph is returned as is."""
    if isinstance(g, MolGraph):
        return g
    src, dst = g.edges(etype="n1_edge")
    ntypes = [nt for nt in NTYPES if nt in g.ntypes]
    out = MolGraph({nt: int(g.num_nodes(nt)) for nt in ntypes}, src, dst,
                   {nt: torch.as_tensor(g.batch_num_nodes(nt)).detach().cpu().to(torch.int64) for nt in ntypes})
    for nt in ntypes:
        for k, v in g.nodes[nt].data.items():
            out.nodes[nt].data[k] = v
    return out


def batch(graphs: Sequence[MolGraph]) -> MolGraph:
    """Concatenate molecules; tuple indices and edges are shifted by the atom offset.

    Mirrors reference utils/dgl_utils.py:11-60 (without the deep copies: inputs are not mutated).
    """
    if len(graphs) == 0:
        raise ValueError("cannot batch an empty list of graphs")
    confs = {g.nodes["n1"].data["xyz"].shape[1] for g in graphs if "xyz" in g.nodes["n1"].data}
    if len(confs) > 1:
        raise ValueError(f"All graphs must have the same number of conformations but found {sorted(confs)}")
    offs = [0]
    for g in graphs[:-1]:
        offs.append(offs[-1] + g.num_nodes("n1"))
    num_nodes = {nt: sum(g.num_nodes(nt) for g in graphs) for nt in graphs[0].ntypes}
    src = torch.cat([g._src + o for g, o in zip(graphs, offs)])
    dst = torch.cat([g._dst + o for g, o in zip(graphs, offs)])
    bnn = {nt: torch.cat([g.batch_num_nodes(nt) for g in graphs]) for nt in graphs[0].ntypes}
    out = MolGraph(num_nodes, src, dst, bnn)
    for nt in graphs[0].ntypes:
        for k in graphs[0].nodes[nt].data.keys():
            parts = [g.nodes[nt].data[k] for g in graphs]
            if k == "idxs":
                parts = [p + o for p, o in zip(parts, offs)]
            out.nodes[nt].data[k] = torch.cat(parts, dim=0)
    return out


def unbatch(bg: MolGraph) -> List[MolGraph]:
    """Inverse of `batch` (reference utils/dgl_utils.py:63-82, without dummy-conformation removal)."""
    nb = bg.batch_size
    counts = {nt: bg.batch_num_nodes(nt).tolist() for nt in bg.ntypes}
    starts = {nt: 0 for nt in bg.ntypes}
    estart = 0
    out = []
    src, dst = bg.edges()
    # edges of molecule i are those whose source atom falls into its atom range
    for i in range(nb):
        a0, na = starts["n1"], counts["n1"][i]
        m = (src >= a0) & (src < a0 + na)
        g = MolGraph({nt: counts[nt][i] for nt in bg.ntypes}, src[m] - a0, dst[m] - a0)
        for nt in bg.ntypes:
            s, n = starts[nt], counts[nt][i]
            for k, v in bg.nodes[nt].data.items():
                piece = v[s:s + n]
                if k == "idxs":
                    piece = piece - a0
                g.nodes[nt].data[k] = piece
        for nt in bg.ntypes:
            starts[nt] += counts[nt][i]
        out.append(g)
    return out
