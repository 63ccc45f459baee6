"""Gradient exchange over NVLink peer memory (csrc/peer_allreduce.cu) -- host side.

One process per GPU: every rank allocates its flat gradient buffer (plus a block of barrier flags) with
`grappa_b200_ipc_alloc`, the CUDA IPC handles are exchanged through `torch.distributed` (plumbing only), every rank maps
its peers' buffers, and `PeerGradients.allreduce(start, count)` sums one span across the ranks with ONE kernel on the
current stream.  Replaces the NCCL all-reduce the reference gets from pytorch_lightning's DDP strategy
(reference src/grappa/training/trainrun.py:166-176) on single-node runs; `training.Trainer` falls back to
`dist.all_reduce` when the ranks are not on one node or `GRAPPA_B200_PEER_ALLREDUCE=0`.
"""
from __future__ import annotations

import ctypes as C
import os

import torch
import torch.distributed as dist

from . import _lib
from ._lib_ops import MAX_PEERS, PEER_MAX_CTAS, IpcHandle, PeerAllreduceArgs

FLAG_WORDS = MAX_PEERS * PEER_MAX_CTAS


class _DeviceArray:
    """`__cuda_array_interface__` view of raw device memory, so that torch can wrap it without copying."""

    def __init__(self, ptr: int, n_floats: int):
        self.__cuda_array_interface__ = {"shape": (n_floats,), "typestr": "<f4", "data": (ptr, False), "version": 3, "strides": None}


class PeerUnavailable(RuntimeError):
    """Raised on ALL ranks alike when any rank could not set up the peer mappings (the caller falls back to NCCL)."""


def _agree(ok: bool, device: torch.device, what: str):
    """Collective: every rank learns whether every rank succeeded; raises PeerUnavailable everywhere if one did not."""
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if int(flag.item()) == 0:
        raise PeerUnavailable(f"peer-memory setup failed on at least one rank ({what})")


class PeerGradients:
    """Flat fp32 gradient buffer of `n_floats` (rounded up to 16 bytes) in IPC-shareable device memory + the peer mappings."""

    def __init__(self, n_floats: int, device: torch.device, ctas: int | None = None):
        assert dist.is_initialized() and device.type == "cuda"
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        if self.world > MAX_PEERS:
            raise _lib.GrappaB200Error(f"peer all-reduce supports up to {MAX_PEERS} ranks on one node, got {self.world}")
        self.device = device
        self.n = (n_floats + 3) // 4 * 4
        self.ctas = int(os.environ.get("GRAPPA_B200_PEER_CTAS", ctas or 32))
        lib = _lib.lib()
        with torch.cuda.device(device):
            # every step that can fail locally is followed by an agreement, so that either all ranks use the peer kernel or
            # all of them fall back (a rank that silently left the protocol would hang the others in the first barrier)
            ptr, handle = C.c_void_p(), IpcHandle()
            nbytes = self.n * 4 + FLAG_WORDS * 4
            rc = lib.grappa_b200_ipc_alloc(nbytes, C.byref(ptr), C.byref(handle))
            _agree(rc == 0, device, "cudaMalloc / cudaIpcGetMemHandle")
            self._base = ptr.value
            # handles travel as plain bytes through the process group (host plumbing only)
            mine = (self.rank, int(device.index if device.index is not None else torch.cuda.current_device()), bytes(handle.bytes))
            everyone = [None] * self.world
            dist.all_gather_object(everyone, mine)
            self._peer_ptrs = [None] * self.world
            ok = True
            for r, _dev, hb in everyone:
                if r == self.rank:
                    self._peer_ptrs[r] = self._base
                    continue
                h = IpcHandle()
                C.memmove(h.bytes, hb, 64)
                p = C.c_void_p()
                ok = ok and lib.grappa_b200_ipc_open(C.byref(h), C.byref(p)) == 0
                self._peer_ptrs[r] = p.value
            _agree(ok, device, "cudaIpcOpenMemHandle: " + _lib.lib().grappa_b200_last_error().decode("utf-8", "replace"))
            self.grad = torch.as_tensor(_DeviceArray(self._base, self.n), device=device)
            self.epoch = torch.zeros(PEER_MAX_CTAS + 1, dtype=torch.int32, device=device)
            torch.cuda.synchronize(device)
        dist.barrier()     # every rank has zeroed its flags and mapped its peers before the first kernel
        self._args = PeerAllreduceArgs()
        for r in range(self.world):
            self._args.data[r] = self._peer_ptrs[r]
            self._args.flags[r] = self._peer_ptrs[r] + self.n * 4
        self._args.rank, self._args.world = self.rank, self.world
        self._args.epoch = self.epoch.data_ptr()
        self._args.ctas = self.ctas
        self._args.split = int(os.environ.get("GRAPPA_B200_PEER_SPLIT", "1"))

    def allreduce(self, start: int, count: int):
        """Sum grad[start:start+count] over all ranks, in place on every rank, on the current CUDA stream."""
        if count <= 0:
            return
        self._args.start, self._args.count = int(start), int(count)
        _lib.check(_lib.lib().grappa_b200_peer_allreduce(C.byref(self._args), torch.cuda.current_stream().cuda_stream), "peer_allreduce")

    def timed_out(self) -> bool:
        """True if a barrier wait inside any all-reduce kernel gave up (a peer never arrived)."""
        return bool(self.epoch[PEER_MAX_CTAS].item())

    def close(self):
        lib = _lib.lib()
        for r, p in enumerate(self._peer_ptrs):
            if p is not None and r != self.rank:
                lib.grappa_b200_ipc_close(p)
        self._peer_ptrs = [None] * self.world
