"""Peer-mapped gradient exchange buffers for the fused reduce -> all-reduce -> Adam kernel (igmc_reduce_update).

One allocation per rank (double-buffered flat gradient + arrival flags), mapped by every rank of the node through
CUDA IPC.  ``torch.distributed`` (NCCL or gloo) is used ONCE, to pass the 64-byte handles around; after that the
ranks talk through NVLink loads / stores inside the update kernel.  SURVEY.md §8(e): the reference has no data
parallelism; world 1 (no peers) is the reference's own step.
"""
import ctypes as C

import torch

from . import _lib

try:
    import torch.distributed as dist
except Exception:  # pragma: no cover
    dist = None


class Exchange(object):
    FLAG_BYTES = 64

    def __init__(self, param_count, device, rank=0, world=1):
        if world > _lib.MAX_RANKS:
            raise NotImplementedError("the peer exchange covers one node (<= %d ranks)" % _lib.MAX_RANKS)
        lib = _lib.load()
        self.lib, self.rank, self.world = lib, int(rank), int(world)
        self.stride = (int(param_count) + 31) & ~31
        self.bytes = 2 * self.stride * 4 + self.FLAG_BYTES
        self.device = torch.device(device)
        self.state = torch.zeros(4, dtype=torch.int64, device=self.device)   # step counter, tickets [3 x int32 + pad]
        self.peers = {}
        self.local = None
        with torch.cuda.device(self.device):
            ptr, handle = C.c_void_p(), C.create_string_buffer(64)
            _lib.check(lib.igmc_comm_alloc(self.bytes, C.byref(ptr), handle), "igmc_comm_alloc")
            self.local = ptr.value
            bases = [None] * self.world
            bases[self.rank] = self.local
            if self.world > 1:
                handles = [None] * self.world
                dist.all_gather_object(handles, bytes(handle.raw))
                for r in range(self.world):
                    if r == self.rank:
                        continue
                    pp = C.c_void_p()
                    _lib.check(lib.igmc_comm_open(handles[r], C.byref(pp)), "igmc_comm_open(rank %d)" % r)
                    self.peers[r] = pp.value
                    bases[r] = pp.value
                dist.barrier()   # every rank has mapped every buffer before anyone publishes into it
        c = _lib.Comm()
        for r in range(self.world):
            c.grad[r] = bases[r]
            c.flag[r] = bases[r] + 2 * self.stride * 4
        c.state = self.state.data_ptr()
        c.world, c.rank, c.stride = self.world, self.rank, self.stride
        self.c = c

    def close(self):
        """unmap the peers and free the local allocation (after a barrier: nobody may still be reading it)."""
        if self.local is None:
            return
        torch.cuda.synchronize(self.device)
        if self.world > 1 and dist is not None and dist.is_initialized():
            dist.barrier()
        with torch.cuda.device(self.device):
            for p in self.peers.values():
                self.lib.igmc_comm_close(C.c_void_p(p))
            self.lib.igmc_comm_free(C.c_void_p(self.local))
        self.peers, self.local = {}, None
