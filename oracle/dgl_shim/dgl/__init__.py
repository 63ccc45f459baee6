"""Minimal pure-torch stand-in for the `dgl` package (TEST INFRASTRUCTURE ONLY).

DGL is not installable in this image (no wheel, no network; the reference leaves the DGL
version unpinned: /root/reference/README.md:101, installation_openmm.sh:39-63).  This shim
restates the handful of DGL entry points the reference's hot path touches so that the
reference's own `grappa.models` can be imported UNMODIFIED in the build container to
generate the golden fixtures under tests/golden/.  It is never imported by the product
package `grappa_b200`.

Restated pieces (published DGL semantics, each a few lines of torch):
  * heterograph container: `.nodes[ntype].data`, `.ntypes`, `.num_nodes`, `.batch_num_nodes`,
    `.edges(etype=)`, `.to(device)`, `.node_type_subgraph`
  * `dgl.batch` / `dgl.unbatch`: per-type concatenation with edge-offset shifting
  * `dgl.to_homogeneous`: the n1_edge (src, dst) lists as a homogeneous graph
  * `dgl.readout_nodes(op='sum')`: segment-sum by `batch_num_nodes(ntype)`
  * `dgl.nn.pytorch.conv.DotGatConv`: bias-free `fc`, u_dot_v / sqrt(d), edge softmax over the
    in-edges of each destination node, u_mul_e + sum aggregation.
Call sites in the reference: models/graph_attention.py:170,226,249,283; models/energy.py:69;
utils/graph_utils.py:147,164; utils/dgl_utils.py:60,69; data/Molecule.py:497.
"""
from __future__ import annotations

import copy as _copy
from typing import Dict, List, Tuple

import torch

from . import nn  # noqa: F401  (dgl.nn.pytorch.conv.DotGatConv)
from . import heterograph as _heterograph_module  # noqa: F401  (sub-module first; the function `heterograph` below re-binds the name)


class _NodeView:
    def __init__(self, store):
        self.data = store


class _NodesAccessor:
    def __init__(self, g):
        self._g = g

    def __getitem__(self, ntype):
        return _NodeView(self._g._ndata[ntype])


class _NDataByKey:
    """g.ndata[feat] -> {ntype: tensor}  (only what the reference comments use)."""

    def __init__(self, g):
        self._g = g

    def __getitem__(self, feat):
        return {nt: d[feat] for nt, d in self._g._ndata.items() if feat in d}


class DGLGraph:
    """Dict-of-node-types container with one edge list per canonical edge type."""

    def __init__(self, edges: Dict[Tuple[str, str, str], Tuple[torch.Tensor, torch.Tensor]],
                 num_nodes: Dict[str, int] | None = None):
        self._edges = {k: (v[0].clone(), v[1].clone()) for k, v in edges.items()}
        self._ntypes: List[str] = []
        for (s, _, d) in self._edges:
            for t in (s, d):
                if t not in self._ntypes:
                    self._ntypes.append(t)
        self._ntypes = sorted(self._ntypes)
        self._num_nodes = {}
        for nt in self._ntypes:
            n = 0
            for (s, _, d), (u, v) in self._edges.items():
                if s == nt and len(u):
                    n = max(n, int(u.max()) + 1)
                if d == nt and len(v):
                    n = max(n, int(v.max()) + 1)
            if num_nodes is not None and nt in num_nodes:
                n = num_nodes[nt]
            self._num_nodes[nt] = n
        self._ndata: Dict[str, Dict[str, torch.Tensor]] = {nt: {} for nt in self._ntypes}
        self._batch_num_nodes = {nt: torch.tensor([self._num_nodes[nt]]) for nt in self._ntypes}
        self._batch_num_edges = {k: torch.tensor([len(v[0])]) for k, v in self._edges.items()}

    # -- structure -------------------------------------------------------------------------
    @property
    def ntypes(self):
        return list(self._ntypes)

    @property
    def canonical_etypes(self):
        return list(self._edges.keys())

    @property
    def nodes(self):
        return _NodesAccessor(self)

    @property
    def ndata(self):
        return _NDataByKey(self)

    def num_nodes(self, ntype=None):
        if ntype is None:
            return sum(self._num_nodes.values())
        return self._num_nodes[ntype]

    number_of_nodes = num_nodes

    def _etype_key(self, etype):
        for k in self._edges:
            if k == etype or k[1] == etype:
                return k
        raise KeyError(etype)

    def num_edges(self, etype=None):
        if etype is None:
            return sum(len(v[0]) for v in self._edges.values())
        return len(self._edges[self._etype_key(etype)][0])

    def edges(self, etype=None):
        if etype is None:
            assert len(self._edges) == 1
            return next(iter(self._edges.values()))
        return self._edges[self._etype_key(etype)]

    def batch_num_nodes(self, ntype=None):
        if ntype is None:
            assert len(self._ntypes) == 1
            ntype = self._ntypes[0]
        return self._batch_num_nodes[ntype]

    @property
    def batch_size(self):
        return len(next(iter(self._batch_num_nodes.values())))

    @property
    def device(self):
        for d in self._ndata.values():
            for v in d.values():
                return v.device
        return torch.device("cpu")

    def node_type_subgraph(self, ntypes):
        edges = {k: v for k, v in self._edges.items() if k[0] in ntypes and k[2] in ntypes}
        sub = DGLGraph(edges, {nt: self._num_nodes[nt] for nt in ntypes})
        for nt in ntypes:
            sub._ndata[nt] = dict(self._ndata[nt])
            sub._batch_num_nodes[nt] = self._batch_num_nodes[nt]
        return sub

    def to(self, device):
        g = _copy.copy(self)
        g._edges = {k: (u.to(device), v.to(device)) for k, (u, v) in self._edges.items()}
        g._ndata = {nt: {f: t.to(device) for f, t in d.items()} for nt, d in self._ndata.items()}
        g._batch_num_nodes = {k: v.to(device) for k, v in self._batch_num_nodes.items()}
        return g

    def cpu(self):
        return self.to("cpu")

    def __deepcopy__(self, memo):
        g = DGLGraph.__new__(DGLGraph)
        g._edges = {k: (u.clone(), v.clone()) for k, (u, v) in self._edges.items()}
        g._ntypes = list(self._ntypes)
        g._num_nodes = dict(self._num_nodes)
        g._ndata = {nt: {f: (t.detach().clone() if not t.requires_grad else t.clone())
                         for f, t in d.items()} for nt, d in self._ndata.items()}
        g._batch_num_nodes = {k: v.clone() for k, v in self._batch_num_nodes.items()}
        g._batch_num_edges = {k: v.clone() for k, v in self._batch_num_edges.items()}
        return g


DGLHeteroGraph = DGLGraph


def heterograph(data_dict, num_nodes_dict=None):
    edges = {}
    for k, (u, v) in data_dict.items():
        edges[k] = (torch.as_tensor(u), torch.as_tensor(v))
    return DGLGraph(edges, num_nodes_dict)


def node_type_subgraph(g, ntypes):
    return g.node_type_subgraph(ntypes)


def to_homogeneous(g):
    assert len(g.ntypes) == 1, "shim supports single-node-type graphs only"
    nt = g.ntypes[0]
    keys = [k for k in g.canonical_etypes]
    assert len(keys) == 1
    u, v = g._edges[keys[0]]
    h = DGLGraph({("_N", "_E", "_N"): (u, v)}, {"_N": g.num_nodes(nt)})
    h._batch_num_nodes["_N"] = g._batch_num_nodes[nt]
    return h


def batch(graphs):
    ntypes = graphs[0].ntypes
    offsets = {nt: 0 for nt in ntypes}
    edges = {k: ([], []) for k in graphs[0].canonical_etypes}
    for g in graphs:
        for k, (u, v) in g._edges.items():
            edges[k][0].append(u + offsets[k[0]])
            edges[k][1].append(v + offsets[k[2]])
        for nt in ntypes:
            offsets[nt] += g.num_nodes(nt)
    out = DGLGraph({k: (torch.cat(us), torch.cat(vs)) for k, (us, vs) in edges.items()}, offsets)
    for nt in ntypes:
        feats = graphs[0]._ndata[nt].keys()
        for f in feats:
            out._ndata[nt][f] = torch.cat([g._ndata[nt][f] for g in graphs], dim=0)
        out._batch_num_nodes[nt] = torch.cat([g._batch_num_nodes[nt] for g in graphs])
    return out


def unbatch(bg):
    nb = bg.batch_size
    graphs = []
    starts = {nt: 0 for nt in bg.ntypes}
    for i in range(nb):
        counts = {nt: int(bg._batch_num_nodes[nt][i]) for nt in bg.ntypes}
        edges = {}
        for k, (u, v) in bg._edges.items():
            m = (u >= starts[k[0]]) & (u < starts[k[0]] + counts[k[0]])
            edges[k] = (u[m] - starts[k[0]], v[m] - starts[k[2]])
        g = DGLGraph(edges, counts)
        for nt in bg.ntypes:
            for f, t in bg._ndata[nt].items():
                g._ndata[nt][f] = t[starts[nt]:starts[nt] + counts[nt]]
            starts[nt] += counts[nt]
        graphs.append(g)
    return graphs


def readout_nodes(g, feat, weight=None, *, op="sum", ntype=None):
    assert op == "sum" and weight is None
    x = g.nodes[ntype].data[feat]
    counts = g.batch_num_nodes(ntype).to(x.device).long()
    seg = torch.repeat_interleave(torch.arange(len(counts), device=x.device), counts)
    out = torch.zeros((len(counts),) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    return out.index_add(0, seg, x)


def save_graphs(*a, **k):  # pragma: no cover
    raise NotImplementedError("dgl shim: storage is out of scope")


def load_graphs(*a, **k):  # pragma: no cover
    raise NotImplementedError("dgl shim: storage is out of scope")
