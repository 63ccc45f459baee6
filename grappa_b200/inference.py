"""Batched inference front-end.

Host-side mirror of what sits directly on top of `model(g)` when the reference parametrises a system:

    grappa.py:14-57 (Grappa.predict)        molecule -> graph -> model (eval, no_grad) -> Parameters
    data/Parameters.py:63-140 (from_dgl)    per-interaction atom ids, k / eq, torsion |k| + phase (0 or pi)
    data/Molecule.py:286-311                ring / degree features (rdkit in the reference; graph search here)
    utils/dgl_utils.py:210-245              disconnected-graph check

The reference parametrises ONE molecule per call.  `Grappa.predict_many` batches many systems into as few model calls
as fit an atom budget (molecules are independent: SURVEY.md 8e, no communication) and splits the written parameters per
molecule again; `shard` distributes the list over data-parallel ranks.  OpenMM / GROMACS writers consume `Parameters`
exactly as they consume the reference's object (same field names and conventions).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import graph as _graph
from . import synthetic as _syn
from . import tuples as _tuples
from .models import CHARGE_MODELS, MAX_ELEMENT
from .graph import LEVELS, MolGraph


@dataclass
class Parameters:
    """Field names and conventions of reference data/Parameters.py:18-61: torsion amplitudes are non-negative, the sign
    lives in the phase (0 or pi); atoms are external atom ids, tuples are rows of ids."""
    atoms: np.ndarray
    bonds: np.ndarray
    bond_k: np.ndarray
    bond_eq: np.ndarray
    angles: np.ndarray
    angle_k: np.ndarray
    angle_eq: np.ndarray
    propers: np.ndarray
    proper_ks: np.ndarray
    proper_phases: np.ndarray
    impropers: Optional[np.ndarray] = None
    improper_ks: Optional[np.ndarray] = None
    improper_phases: Optional[np.ndarray] = None

    @classmethod
    def from_graph(cls, g, suffix: str = "", check_eq_values: bool = True) -> "Parameters":
        """Parameters.from_dgl (data/Parameters.py:63-140) for one molecule's graph (host or device)."""
        def np_(t):
            return t.detach().cpu().numpy()
        n1 = g.nodes["n1"].data
        atom_ids = np_(n1["ids"]) if "ids" in n1 else np.arange(g.num_nodes("n1"), dtype=np.int64)
        bonds = atom_ids[np_(g.nodes["n2"].data["idxs"])]
        bond_k, bond_eq = np_(g.nodes["n2"].data["k" + suffix]), np_(g.nodes["n2"].data["eq" + suffix])
        angles = atom_ids[np_(g.nodes["n3"].data["idxs"])]
        angle_k, angle_eq = np_(g.nodes["n3"].data["k" + suffix]), np_(g.nodes["n3"].data["eq" + suffix])
        if check_eq_values:
            MAX_ANGLE, MAX_BOND_LENGTH = 45, 0.5
            if np.any(angle_eq < np.pi / 180 * MAX_ANGLE):
                n_small = int(np.sum(angle_eq < np.pi / 180 * MAX_ANGLE))
                raise RuntimeError(f"{n_small} angles are smaller than {MAX_ANGLE} degrees. This can lead to numerical "
                                   f"instabilities in the model.\\nThe smallest angle is {np.min(angle_eq) * 180 / np.pi} "
                                   f"degrees at atom ids {angles[np.argmin(angle_eq)]}.")
            if np.any(bond_eq < MAX_BOND_LENGTH):
                n_small = int(np.sum(bond_eq < MAX_BOND_LENGTH))
                raise RuntimeError(f"{n_small} bond eq lengths are smaller than {MAX_BOND_LENGTH} Angstrom. This can lead to "
                                   f"numerical instabilities in the model.\\nThe smallest bond eq length is {np.min(bond_eq)} "
                                   f"Angstrom at atom ids {bonds[np.argmin(bond_eq)]}.")
        pk = np_(g.nodes["n4"].data["k" + suffix])
        proper_phases = np.where(pk >= 0., np.zeros_like(pk), np.zeros_like(pk) + np.pi)     # '>=' for propers ...
        ik = np_(g.nodes["n4_improper"].data["k" + suffix])
        improper_phases = np.where(ik > 0, np.zeros_like(ik), np.zeros_like(ik) + np.pi)       # ... '>' for impropers (:118-126)
        return cls(atoms=atom_ids, bonds=bonds, bond_k=bond_k, bond_eq=bond_eq, angles=angles, angle_k=angle_k,
                   angle_eq=angle_eq, propers=atom_ids[np_(g.nodes["n4"].data["idxs"])], proper_ks=np.abs(pk),
                   proper_phases=proper_phases, impropers=atom_ids[np_(g.nodes["n4_improper"].data["idxs"])],
                   improper_ks=np.abs(ik), improper_phases=improper_phases)


def molecule_graph(atoms: Sequence[int], bonds: Sequence[Sequence[int]], atomic_numbers: Sequence[int],
                   partial_charges: Sequence[float], impropers: Sequence[Sequence[int]] = (),
                   charge_model: str = "amber99", xyz: Optional[np.ndarray] = None) -> MolGraph:
    """Topology -> model input graph, as `Molecule.to_dgl` builds it (data/Molecule.py:429-520): atoms / bonds /
    improper candidates are given by external atom id; tuples come from the bit-exact tuple builder; ring membership and
    degree features are computed from the bond graph (no rdkit)."""
    atoms = np.asarray(atoms, dtype=np.int64)
    idx_of = {int(a): i for i, a in enumerate(atoms.tolist())}
    if len(idx_of) != len(atoms):
        raise ValueError("atom ids must be unique")
    b = np.array([[idx_of[int(x)] for x in bond] for bond in bonds], dtype=np.int64).reshape(-1, 2)
    imp = np.array([[idx_of[int(x)] for x in t] for t in impropers], dtype=np.int64).reshape(-1, 4)
    n = len(atoms)
    tup = _tuples.build_tuples(n, b, imp)
    z = np.asarray(atomic_numbers, dtype=np.int64)
    if z.min() < 1 or z.max() > MAX_ELEMENT:
        raise ValueError(f"atomic numbers must be in 1..{MAX_ELEMENT}")
    onehot = np.zeros((n, MAX_ELEMENT), dtype=np.float32)
    onehot[np.arange(n), z - 1] = 1.0
    deg = np.zeros(n, dtype=np.int64)
    np.add.at(deg, b.reshape(-1), 1)
    degree = np.zeros((n, 6), dtype=np.float32)
    ok = (deg >= 1) & (deg <= 6)     # reference utils/rdkit_utils.py:55-67: degree outside 1..6 -> all-zero row
    degree[np.arange(n)[ok], deg[ok] - 1] = 1.0
    if charge_model not in CHARGE_MODELS:
        raise ValueError(f"charge_model must be one of {CHARGE_MODELS}")
    cm = np.zeros((n, len(CHARGE_MODELS)), dtype=np.float32)
    cm[:, CHARGE_MODELS.index(charge_model)] = 1.0
    feats = {"atomic_number": onehot, "partial_charge": np.asarray(partial_charges, dtype=np.float32),
             "ring_encoding": _syn.ring_encoding(n, b), "degree": degree, "charge_model": cm}
    g = _graph.graph_from_molecule(n, tup["bonds"], tup["angles"], tup["propers"], tup["impropers"], feats, xyz)
    g.nodes["n1"].data["ids"] = torch.from_numpy(atoms.copy())
    return g


def connected_components(g) -> int:
    """Number of connected components of the bond graph (union-find on the host)."""
    n = g.num_nodes("n1")
    parent = list(range(n))

    def find(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x

    src, dst = g.edges()
    for a, b in zip(src.tolist(), dst.tolist()):
        ra, rb = find(a), find(b)
        if ra != rb:
            parent[ra] = rb
    return len({find(i) for i in range(n)})


class _Captured:
    __slots__ = ("graph", "g_static", "pack", "keys")


class Grappa:
    """Model wrapper (reference grappa.py:14-57): eval-mode, no-grad parametrisation of molecules.

    `use_cuda_graph`: the forward pass of a batch is ~230 kernel launches whose host-side issue (5.6 ms for a 1,502-atom
    protein) dwarfs the ~0.5 ms of device work, so the second time a batch SHAPE is seen (atoms, edges, tuple counts,
    feature widths) its forward is captured into a CUDA graph; later batches of that shape copy their inputs and index
    tables into the graph's static buffers and replay it with one launch."""

    def __init__(self, model, device: str = "cuda", use_cuda_graph: bool = True, max_graphs: int = 8):
        self.model = model.to(device).eval()
        self.device = torch.device(device)
        self.field_of_view = getattr(model, "field_of_view", None)
        self.use_cuda_graph = use_cuda_graph and self.device.type == "cuda"
        self.max_graphs = max_graphs
        self._captured: Dict[tuple, _Captured] = {}
        self._seen: Dict[tuple, int] = {}
        self._stream = torch.cuda.Stream(device=self.device) if self.use_cuda_graph else None

    # ---- one model call ------------------------------------------------------------------------
    def _input_keys(self, g):
        names = set(self.model.gnn.in_feat_name) | {"partial_charge"}
        return [("n1", k) for k in g.nodes["n1"].data.keys() if k in names]

    def _forward(self, bg: MolGraph) -> MolGraph:
        """Batched host graph -> graph (device) carrying the written parameters."""
        from .pack import get_pack
        if not self.use_cuda_graph:
            with torch.no_grad():
                return self.model(bg.to(self.device))
        pack = get_pack(bg)
        keys = self._input_keys(bg)
        sig = (pack.signature(), tuple((k, tuple(bg.nodes[nt].data[k].shape)) for nt, k in keys))
        cap = self._captured.get(sig)
        caller = torch.cuda.current_stream(self.device)
        self._stream.wait_stream(caller)
        with torch.cuda.stream(self._stream), torch.no_grad():
            if cap is None:
                n = self._seen.get(sig, 0)
                self._seen[sig] = n + 1
                if n == 0 or len(self._captured) >= self.max_graphs:
                    out = self.model(bg.to(self.device))
                    caller.wait_stream(self._stream)
                    return out
                cap = _Captured()
                cap.keys = keys
                gs = bg.to(self.device)
                cap.pack = get_pack(gs)
                cap.g_static = gs
                torch.cuda.synchronize(self.device)
                cap.graph = torch.cuda.CUDAGraph()
                from . import ops as _ops
                _ops.drop_workspaces()           # cached scratch of eager launches must not be baked into the graph
                with torch.cuda.graph(cap.graph, stream=self._stream, capture_error_mode="thread_local"):
                    self.model(gs)
                _ops.drop_workspaces()           # scratch allocated while capturing belongs to the graph's pool
                self._captured[sig] = cap
            else:
                for nt, k in cap.keys:
                    cap.g_static.nodes[nt].data[k].copy_(bg.nodes[nt].data[k], non_blocking=True)
                cap.pack.copy_from(pack)
            cap.graph.replay()
            # the static graph keeps the tuple indices of the batch it was captured with: hand back this batch's
            out = bg.to(self.device)
            for lvl in LEVELS:
                for k, v in cap.g_static.nodes[lvl].data.items():
                    if k != "idxs":
                        out.nodes[lvl].data[k] = v.clone()
            out.nodes["n1"].data["h"] = cap.g_static.nodes["n1"].data["h"].clone()
        caller.wait_stream(self._stream)
        return out

    def predict(self, molecule: MolGraph, check_eq_values: bool = True) -> Parameters:
        return self.predict_many([molecule], check_eq_values=check_eq_values)[0]

    def predict_many(self, molecules: Sequence[MolGraph], max_atoms_per_batch: int = 200_000,
                     check_eq_values: bool = True, allow_disconnected: bool = False) -> List[Parameters]:
        """Parametrise many systems: molecules are packed greedily into batches of <= `max_atoms_per_batch` atoms (a
        1,500-atom protein is ~0.4 ms of tensor math; launch count, not FLOPs, is what batching saves), each batch is
        one model call, results come back in input order."""
        molecules = [_graph.as_molgraph(m) for m in molecules]      # DGL-shaped graphs (Molecule.to_dgl) are accepted as they are
        if not allow_disconnected:
            for i, m in enumerate(molecules):
                if connected_components(m) > 1:
                    raise ValueError(f"molecule {i} is not connected (e.g. it contains solvent): parametrise the bonded "
                                     f"components separately (reference utils/dgl_utils.py:210-245)")
        out: List[Optional[Parameters]] = [None] * len(molecules)
        start = 0
        while start < len(molecules):
            end, atoms = start, 0
            while end < len(molecules) and (end == start or atoms + molecules[end].num_nodes("n1") <= max_atoms_per_batch):
                atoms += molecules[end].num_nodes("n1")
                end += 1
            bg = self._forward(_graph.batch(list(molecules[start:end])))
            for j, g in enumerate(_graph.unbatch(bg.cpu())):
                out[start + j] = Parameters.from_graph(g, check_eq_values=check_eq_values)
            start = end
        return out  # type: ignore[return-value]


def shard(items: Sequence, rank: int, world: int) -> List:
    """Item i -> rank i mod world (inference shards molecules, no collective)."""
    return [items[i] for i in range(rank, len(items), world)]
