"""argtypes for the dense / sparse operator entry points (filled in as kernels are added)."""
import ctypes as C


def declare(lib):
    pass
