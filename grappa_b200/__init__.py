"""grappa_b200: B200-native (sm_100a) implementation of Grappa's model hot path.

GNN atom embedding -> permutation-symmetric tuple heads -> differentiable MM energy/forces, as
drop-in torch.nn.Module replacements for `grappa.models` backed by hand-written CUDA kernels behind
a C ABI (include/grappa_b200.h).  There is no CPU fallback: the modules raise if the shared object
is missing or the tensors are not on a CUDA device.
"""
from ._lib import GrappaB200Error  # noqa: F401

__version__ = "0.1.0"
