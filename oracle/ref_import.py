"""Import the UNMODIFIED reference `grappa.models` in the build container (TEST INFRASTRUCTURE).

Only usable where /root/reference is mounted (this container); never on the GPU box and never
from the product package.  Used by tests/golden/make_golden.py to generate fixtures and by
`-m "not gpu"` tests (skipped when the reference is absent) to pin oracle/grappa_oracle.py.

`grappa/__init__.py` drags in openmm/rdkit wrappers (reference src/grappa/__init__.py:1 ->
grappa.py:8 -> data/Molecule.py:9), so a namespace module `grappa` is pre-registered whose
__path__ points at the reference sources; sub-modules then import normally.  DGL is replaced by
oracle/dgl_shim (see its docstring).
"""
import importlib
import os
import sys
import types
import warnings

REFERENCE_SRC = os.environ.get("GRAPPA_REFERENCE_SRC", "/root/reference/src")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dgl_shim")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_SRC, "grappa", "models"))


def import_reference():
    """Returns a namespace with the reference modules needed on the hot path."""
    if not reference_available():
        raise RuntimeError(f"reference sources not found under {REFERENCE_SRC}")
    if _SHIM not in sys.path:
        sys.path.insert(0, _SHIM)
    if "grappa" not in sys.modules:
        pkg = types.ModuleType("grappa")
        pkg.__path__ = [os.path.join(REFERENCE_SRC, "grappa")]
        sys.modules["grappa"] = pkg
        for sub in ("models", "utils", "training", "data"):
            m = types.ModuleType(f"grappa.{sub}")
            m.__path__ = [os.path.join(REFERENCE_SRC, "grappa", sub)]
            sys.modules[f"grappa.{sub}"] = m
            setattr(pkg, sub, m)
    ns = types.SimpleNamespace()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ns.constants = importlib.import_module("grappa.constants")
        sys.modules["grappa"].constants = ns.constants
        ns.graph_utils = importlib.import_module("grappa.utils.graph_utils")
        ns.dgl_utils = importlib.import_module("grappa.utils.dgl_utils")
        ns.tuple_indices = importlib.import_module("grappa.utils.tuple_indices")
        sys.modules["grappa.utils"].graph_utils = ns.graph_utils
        sys.modules["grappa.utils"].dgl_utils = ns.dgl_utils
        ns.energy = importlib.import_module("grappa.models.energy")
        ns.internal_coordinates = importlib.import_module("grappa.models.internal_coordinates")
        ns.grappa_model = importlib.import_module("grappa.models.grappa")
        ns.deploy = importlib.import_module("grappa.models.deploy")
        ns.graph_attention = importlib.import_module("grappa.models.graph_attention")
        ns.interaction_parameters = importlib.import_module("grappa.models.interaction_parameters")
        ns.loss = importlib.import_module("grappa.training.loss")
    import dgl
    ns.dgl = dgl
    return ns


def to_reference_graph(ns, g):
    """MolGraph (grappa_b200.graph) -> shim DGL heterograph with the layout of data/Molecule.py:429-520."""
    import torch
    src, dst = g.edges(etype="n1_edge")
    data = {("n1", "n1_edge", "n1"): (src.long(), dst.long())}
    num = {"n1": g.num_nodes("n1")}
    for nt in ("n2", "n3", "n4", "n4_improper", "g"):
        n = g.num_nodes(nt)
        data[(nt, f"{nt}_edge", nt)] = (torch.arange(n), torch.arange(n))
        num[nt] = n
    dg = ns.dgl.heterograph(data, num)
    for nt in g.ntypes:
        for k, v in g.nodes[nt].data.items():
            dg.nodes[nt].data[k] = v.clone()
        dg._batch_num_nodes[nt] = g.batch_num_nodes(nt).clone()
    return dg


class no_dihedral_noise:
    """Context manager: torch.randn_like -> zeros, neutralising internal_coordinates.py:194-196."""

    def __enter__(self):
        import torch
        self._orig = torch.randn_like
        torch.randn_like = lambda x, *a, **k: torch.zeros_like(x)
        return self

    def __exit__(self, *exc):
        import torch
        torch.randn_like = self._orig
        return False
