"""Data-parallel training step: each rank runs its own tape on a batch shard, the Param gradients
are summed with one NCCL allreduce over a flat bucket, then every rank applies the same axpy!.
(BASELINE.json north_star; SURVEY.md 8e.  The reference is single-process.)

torch.distributed is used only as the launcher-side rendezvous (to hand rank 0's ncclUniqueId to
the other ranks); the collective itself is ag_allreduce_sum on the library stream.
"""
import ctypes as C

import numpy as np

from . import darray as D
from ._lib import lib, check
from .darray import DArray
from .sparse import Sparse


class DataParallel:
    def __init__(self, rank=0, world=1, broadcast_bytes=None):
        """broadcast_bytes(b: bytes|None) -> bytes: returns rank 0's payload on every rank."""
        self.rank, self.world = rank, world
        if world > 1:
            idbuf = (C.c_uint8 * 128)()
            if rank == 0:
                check(lib.ag_comm_unique_id(idbuf))
            payload = broadcast_bytes(bytes(idbuf) if rank == 0 else None)
            idbuf = (C.c_uint8 * 128).from_buffer_copy(payload)
            check(lib.ag_comm_init(world, rank, idbuf))
        self._bucket = None

    @staticmethod
    def shard_columns(ncols, rank, world):
        """Contiguous column range [lo, hi) of this rank (column-major => one contiguous block)."""
        base, rem = divmod(ncols, world)
        lo = rank * base + min(rank, rem)
        return lo, lo + base + (1 if rank < rem else 0)

    def allreduce_grads(self, grads):
        """Sum a list of dense gradients (DArray / Sparse / None) across ranks.  Returns dense DArrays
        that view one flat bucket (cat1d order = params order, src/cat.jl:150-182)."""
        dense = [g.full() if isinstance(g, Sparse) else g for g in grads]
        if self.world == 1:
            return dense
        ref = next(g for g in dense if g is not None)
        sizes = [0 if g is None else g.size for g in dense]
        tot = sum(sizes)
        if self._bucket is None or self._bucket.size != tot or self._bucket.dtype != ref.dtype:
            self._bucket = DArray.empty((tot,), ref.dtype)
        b, off, views = self._bucket, 0, []
        for g, n in zip(dense, sizes):
            if g is None:
                views.append(None)
                continue
            g = D.materialize(g)
            D.slab_copy(g, n, 0, b, tot, off, 1, n, 1)
            views.append(DArray(b.buf, b.ptr + off * b.dtype.itemsize, g.shape, b.dtype))
            off += n
        check(lib.ag_allreduce_sum(b.dt, D.vptr(b), tot))
        return views

    def close(self):
        check(lib.ag_comm_destroy())
