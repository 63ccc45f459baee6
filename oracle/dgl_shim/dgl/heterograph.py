"""`import dgl.heterograph` (reference data/Molecule.py:9) -- the container class under its DGL module path.

Like in DGL itself, the package attribute `dgl.heterograph` stays the FUNCTION: the package imports this sub-module
first and binds the function of the same name afterwards."""


def __getattr__(name):          # resolved lazily: the package body is still executing when this module is imported
    import dgl
    if name in ("DGLGraph", "DGLHeteroGraph"):
        return dgl.DGLGraph
    raise AttributeError(name)
