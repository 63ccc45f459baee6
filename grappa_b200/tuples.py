"""Tuple index construction through the C ABI (host code in csrc/tuples.cpp).

API mirror of reference src/grappa/utils/tuple_indices.py: `get_idx_tuples(bonds)` (:7-63) and
`get_torsions(torsion_ids, bonds)` (:144-216); orderings are bit-exact, results are int64 numpy
arrays instead of lists of tuples.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Sequence

import numpy as np

from . import _lib

IMPROPER_CENTRAL_IDX = 2  # reference constants.py:36


def _as_bonds(bonds) -> np.ndarray:
    b = np.ascontiguousarray(np.asarray(bonds, dtype=np.int64).reshape(-1, 2))
    return b


def get_idx_tuples(bonds) -> Dict[str, np.ndarray]:
    """{'bonds': (n,2) sorted per row, 'angles': (n,3) with a<c, 'propers': (n,4) with p0<p3}."""
    lib = _lib.lib()
    b = _as_bonds(bonds)
    na, npr = C.c_int64(0), C.c_int64(0)
    _lib.check(lib.grappa_b200_tuples_count(b.ctypes.data, len(b), C.byref(na), C.byref(npr)), "tuples_count")
    bs = np.empty_like(b)
    ang = np.empty((na.value, 3), dtype=np.int64)
    pro = np.empty((npr.value, 4), dtype=np.int64)
    _lib.check(lib.grappa_b200_tuples_build(b.ctypes.data, len(b), bs.ctypes.data, ang.ctypes.data, pro.ctypes.data),
               "tuples_build")
    return {"bonds": bs, "angles": ang, "propers": pro}


def get_torsions(torsion_ids, bonds, central_atom_position: int = IMPROPER_CENTRAL_IDX):
    """(propers, impropers): impropers hold 3 cyclic orderings per centre, central atom at index 2."""
    lib = _lib.lib()
    b = _as_bonds(bonds)
    t = np.ascontiguousarray(np.asarray(torsion_ids, dtype=np.int64).reshape(-1, 4))
    pro = np.empty((len(t), 4), dtype=np.int64)
    imp = np.empty((3 * len(t), 4), dtype=np.int64)
    npr, nim = C.c_int64(0), C.c_int64(0)
    _lib.check(lib.grappa_b200_torsions_classify(b.ctypes.data, len(b), t.ctypes.data, len(t), central_atom_position,
                                                 pro.ctypes.data, C.byref(npr), imp.ctypes.data, C.byref(nim)),
               "torsions_classify")
    return pro[:npr.value].copy(), imp[:nim.value].copy()


def ring_encoding(n_atoms: int, bonds) -> np.ndarray:
    """(n_atoms, 7) float32: [in a ring, in a ring of 3, 4, ..., 8 atoms] per atom -- the feature the reference reads
    from rdkit (utils/rdkit_utils.py:7-24), computed from the bond list alone (csrc/tuples.cpp)."""
    b = _as_bonds(bonds)
    enc = np.empty((int(n_atoms), 7), dtype=np.float32)
    _lib.check(_lib.lib().grappa_b200_ring_encoding(int(n_atoms), b.ctypes.data if len(b) else None, len(b),
                                                    enc.ctypes.data if n_atoms else None), "ring_encoding")
    return enc


def build_tuples(n_atoms: int, bonds, improper_candidates: Sequence = ()) -> Dict[str, np.ndarray]:
    """All four tuple levels of one molecule from its bond list (+ candidate improper centres)."""
    d = get_idx_tuples(bonds)
    if len(improper_candidates):
        _, imp = get_torsions(improper_candidates, bonds)
    else:
        imp = np.empty((0, 4), dtype=np.int64)
    d["impropers"] = imp
    return d
