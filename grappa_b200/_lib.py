"""ctypes binding of libgrappa_b200.so -- the only way the Python host code reaches the kernels.

The product path has NO fallback: if the shared object is missing or a kernel entry point returns an
error, a `GrappaB200Error` is raised.  Struct layouts mirror include/grappa_b200.h field by field.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GRAPPA_B200_LIB") or os.path.join(HERE, "libgrappa_b200.so")   # override: instrumented debug builds (tools/gemm_trace.py)


class GrappaB200Error(RuntimeError):
    pass


c_f32p = C.c_void_p   # device pointers travel as integers (tensor.data_ptr())
c_i32p = C.c_void_p


class EnergyArgs(C.Structure):
    _fields_ = [
        ("xyz", c_f32p),
        ("n_atoms", C.c_int32), ("n_confs", C.c_int32), ("n_mols", C.c_int32),
        ("max_atoms_per_mol", C.c_int32),
        ("atom_off", c_i32p),
        ("idx", c_i32p * 4),
        ("tup_off", c_i32p * 4),
        ("n_tuples", C.c_int32 * 4),
        ("k", c_f32p * 4),
        ("eq", c_f32p * 2),
        ("n_per", C.c_int32 * 2),
        ("level_mask", C.c_int32),
        ("offset_torsion", C.c_int32),
        ("energy", c_f32p),
        ("term_energy", c_f32p * 4),
        ("grad", c_f32p),
        ("x", c_f32p * 4),
        ("tuple_energy", c_f32p * 4),
        ("sched", c_i32p * 4),
        ("round_off", c_i32p * 4),
        ("sched_groups", C.c_int32),
        ("max_tuples_per_mol", C.c_int32 * 4),
        ("max_rounds_per_mol", C.c_int32 * 4),
    ]


class EnergyBwdArgs(C.Structure):
    _fields_ = [
        ("fwd", EnergyArgs),
        ("g_energy", c_f32p),
        ("g_grad", c_f32p),
        ("dk", c_f32p * 4),
        ("deq", c_f32p * 2),
        ("workspace", c_f32p),
        ("workspace_bytes", C.c_int64),
    ]


_lib = None


def _declare(lib):
    i64p = C.POINTER(C.c_int64)
    lib.grappa_b200_last_error.restype = C.c_char_p
    lib.grappa_b200_abi_version.restype = C.c_int
    lib.grappa_b200_sm_count.restype = C.c_int
    lib.grappa_b200_launch_count.restype = C.c_int64
    lib.grappa_b200_tuples_count.argtypes = [C.c_void_p, C.c_int64, i64p, i64p]
    lib.grappa_b200_tuples_build.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.grappa_b200_torsions_classify.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int,
                                                  C.c_void_p, i64p, C.c_void_p, i64p]
    lib.grappa_b200_conflict_free_rounds.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                                     C.c_void_p, C.c_int64]
    lib.grappa_b200_conflict_free_rounds.restype = C.c_int64
    lib.grappa_b200_ring_encoding.argtypes = [C.c_int64, C.c_void_p, C.c_int64, C.c_void_p]
    lib.grappa_b200_ring_encoding.restype = C.c_int
    lib.grappa_b200_energy_fwd.argtypes = [C.POINTER(EnergyArgs), C.c_int, C.c_void_p]
    lib.grappa_b200_energy_bwd.argtypes = [C.POINTER(EnergyBwdArgs), C.c_void_p]
    lib.grappa_b200_energy_bwd_workspace.argtypes = [C.POINTER(EnergyArgs)]
    lib.grappa_b200_energy_bwd_workspace.restype = C.c_int64
    from . import _lib_ops
    _lib_ops.declare(lib)


def lib():
    """The loaded shared library (raises GrappaB200Error if it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GrappaB200Error(
                f"{LIB_PATH} not found: the CUDA extension is not built. Run `python -m grappa_b200.build` "
                "(or __graft_entry__.build()). grappa_b200 has no CPU / PyTorch fallback.")
        try:
            l = C.CDLL(LIB_PATH)
        except OSError as e:  # pragma: no cover
            raise GrappaB200Error(f"cannot load {LIB_PATH}: {e}") from e
        _declare(l)
        if l.grappa_b200_abi_version() != 3:
            raise GrappaB200Error("libgrappa_b200.so ABI version mismatch; rebuild")
        _lib = l
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().grappa_b200_last_error().decode("utf-8", "replace")
        raise GrappaB200Error(f"{what} failed (code {rc}): {msg}")


def current_stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise GrappaB200Error(
                "grappa_b200 kernels need CUDA tensors (got a CPU tensor); there is no CPU fallback. "
                "Move the graph and the module to a B200 device first.")


def launch_count() -> int:
    """Kernels launched by libgrappa_b200.so so far in this process."""
    return int(lib().grappa_b200_launch_count())
