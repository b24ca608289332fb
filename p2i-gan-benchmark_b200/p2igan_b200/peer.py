"""NVLink peer-memory gradient exchange: flat gradient buffers that every rank of the node maps through
CUDA IPC, and ``PeerAllReduce`` -- one ``p2i_peer_allreduce`` kernel per rank and exchange instead of an NCCL call, so a
data-parallel training step is a single CUDA graph.  ``torch.distributed`` is only used once, to swap the IPC handles."""
from __future__ import annotations

import ctypes
import os
from typing import List

import torch
import torch.distributed as dist

from ._lib import LIB, stream


class PeerBuffer:
    """A cudaMalloc'd device buffer exported to the other ranks (not from PyTorch's caching allocator: IPC handles name
    whole allocations).  ``tensor`` is a float32 / int32 view PyTorch can use like any other tensor."""

    def __init__(self, nbytes: int, dtype=torch.float32):
        self.nbytes = int(nbytes)
        p = ctypes.c_void_p()
        LIB.call("p2i_peer_alloc", ctypes.byref(p), ctypes.c_longlong(self.nbytes))
        self.ptr = p.value
        self.dtype = dtype
        n = self.nbytes // 4
        typestr = "<f4" if dtype == torch.float32 else "<i4"
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (self.ptr, False), "version": 2}
        self.tensor = torch.as_tensor(self, device=torch.device("cuda", torch.cuda.current_device()))
        self._imported: List[int] = []

    def handle(self) -> bytes:
        h = ctypes.create_string_buffer(64)
        LIB.call("p2i_peer_export", ctypes.c_void_p(self.ptr), h)
        return h.raw

    def close(self):
        for q in self._imported:
            LIB.call("p2i_peer_close", ctypes.c_void_p(q))
        self._imported = []
        if self.ptr:
            LIB.call("p2i_peer_free", ctypes.c_void_p(self.ptr))
            self.ptr = 0


def _exchange(buf: PeerBuffer, group):
    """-> (pointers of every rank's buffer as mapped in this process (own entry = local pointer), error or None).
    Never raises before all collectives of the exchange are done: a local failure is reported in the second value so
    that the caller can take a COLLECTIVE decision (a rank that bailed out early would leave the others hanging)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    err = None
    try:
        mine = buf.handle()
    except Exception as e:  # noqa: BLE001
        mine, err = None, e
    handles = [None] * world
    dist.all_gather_object(handles, mine, group=group)
    ptrs = []
    for r in range(world):
        if r == rank:
            ptrs.append(buf.ptr)
            continue
        try:
            if handles[r] is None:
                raise RuntimeError(f"rank {r} could not export its buffer")
            if os.environ.get("P2I_PEER_FAIL") == str(rank):          # test hook: simulate an IPC failure on one rank
                raise RuntimeError("simulated CUDA IPC failure (P2I_PEER_FAIL)")
            q = ctypes.c_void_p()
            h = ctypes.create_string_buffer(handles[r], 64)
            LIB.call("p2i_peer_import", h, ctypes.byref(q))
            buf._imported.append(q.value)
            ptrs.append(q.value)
        except Exception as e:  # noqa: BLE001
            err = err or e
            ptrs.append(0)
    return ptrs, err


class PeerAllReduce:
    """In-place sum of ``buffer.tensor[:n]`` over the ranks of ``group`` (one node, NVLink / NVSwitch P2P)."""

    def __init__(self, n_elems: int, group=None):
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("PeerAllReduce needs an initialised torch.distributed process group (handle exchange)")
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > 8:
            raise ValueError("PeerAllReduce spans one node (<= 8 ranks)")
        self.n = (int(n_elems) + 3) // 4 * 4
        self.buffer = PeerBuffer(self.n * 4, torch.float32)
        self.flags = PeerBuffer(int(LIB.load().p2i_peer_flags_bytes()), torch.int32)
        self.state = torch.zeros(2, dtype=torch.int32, device=self.buffer.tensor.device)     # [epoch, err]
        (bufs, e1), (flags, e2) = _exchange(self.buffer, group), _exchange(self.flags, group)
        # collective verdict: either every rank mapped every buffer, or every rank raises here
        ok = torch.tensor([0 if (e1 or e2) else 1], dtype=torch.int32, device=self.buffer.tensor.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok) == 0:
            self.close()
            raise RuntimeError(f"CUDA IPC peer mapping failed on at least one rank (local error: {e1 or e2!r})")
        self._bufs = (ctypes.c_void_p * self.world)(*bufs)
        self._flags = (ctypes.c_void_p * self.world)(*flags)

    @property
    def tensor(self) -> torch.Tensor:
        return self.buffer.tensor

    def all_reduce(self) -> None:
        LIB.call("p2i_peer_allreduce", self._bufs, self._flags, self.rank, self.world, ctypes.c_longlong(self.n),
                 ctypes.c_void_p(self.state.data_ptr()), ctypes.c_void_p(self.state.data_ptr() + 4), stream())

    def all_reduce_range(self, offset: int, n: int, blocks: int = 0) -> None:
        """In-place sum of elements [offset, offset + n) only (multiples of 4), with ``blocks`` CTAs (0 = one per SM)."""
        LIB.call("p2i_peer_allreduce_range", self._bufs, self._flags, self.rank, self.world, ctypes.c_longlong(int(offset)),
                 ctypes.c_longlong(int(n)), int(blocks), ctypes.c_void_p(self.state.data_ptr()),
                 ctypes.c_void_p(self.state.data_ptr() + 4), stream())

    def check(self) -> None:
        """Host-side check of the time-out word (synchronises; call outside the hot loop)."""
        if int(self.state[1]) != 0:
            raise RuntimeError("p2i_peer_allreduce: a peer did not reach the exchange within the time-out")

    def close(self):
        self.buffer.close()
        self.flags.close()
