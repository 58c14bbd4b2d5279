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
    def __init__(self, rank=0, world=1, broadcast_bytes=None, allgather_bytes=None, p2p_bytes=8 << 20):
        """broadcast_bytes(b: bytes|None) -> bytes: returns rank 0's payload on every rank.
        allgather_bytes(b: bytes) -> list[bytes] (optional): enables the one-shot peer-memory allreduce
        (ag_p2p_*) for buckets up to p2p_bytes; larger buckets and missing peer access use NCCL."""
        self.rank, self.world = rank, world
        self.p2p_cap = 0
        self._allgather = allgather_bytes
        self._layouts_ok = set()
        if world > 1:
            idbuf = (C.c_uint8 * 128)()
            if rank == 0:
                check(lib.ag_comm_unique_id(idbuf))
            payload = broadcast_bytes(bytes(idbuf) if rank == 0 else None)
            idbuf = (C.c_uint8 * 128).from_buffer_copy(payload)
            check(lib.ag_comm_init(world, rank, idbuf))
            if allgather_bytes is not None and p2p_bytes > 0:
                # every rank must agree on the collective: gather a success flag after each stage and fall back
                # to NCCL everywhere if CUDA IPC / peer access is unavailable on any rank
                h = (C.c_uint8 * 64)()
                ok = lib.ag_p2p_export(int(p2p_bytes), world, h) == 0
                got = allgather_bytes((b"\x01" if ok else b"\x00") + bytes(h))
                if all(g[:1] == b"\x01" for g in got):
                    hb = (C.c_uint8 * (64 * world)).from_buffer_copy(b"".join(g[1:] for g in got))
                    ok = lib.ag_p2p_connect(world, rank, hb) == 0
                    if all(g == b"\x01" for g in allgather_bytes(b"\x01" if ok else b"\x00")):
                        self.p2p_cap = int(p2p_bytes)
                if not self.p2p_cap:
                    lib.ag_p2p_destroy()
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
        from . import lazy
        lazy.flush()        # the bucket is reused in place from step to step
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
            D.slab_copy(g, n, 0, b, tot, off, 1, n, 1, fresh=True)
            views.append(DArray(b.buf, b.ptr + off * b.dtype.itemsize, g.shape, b.dtype))
            off += n
        if self.p2p_cap and b.dtype == np.float32 and tot * 4 <= min(self.p2p_cap, 1 << 20):
            check(lib.ag_p2p_allreduce_sum(b.dt, D.vptr(b), tot))
        else:
            check(lib.ag_allreduce_sum(b.dt, D.vptr(b), tot))
        return views

    def update(self, alpha, grads, params, gclip=None):
        """The data-parallel SGD step: params[k] += alpha * (sum over ranks of grads[k]), i.e. allreduce followed
        by `axpy!(alpha, g, w)` for every Param.  One launch when the peer-memory collective is available (the
        update is fused behind it); one multi-tensor axpy launch on a single rank.  Returns the reduced grads.

        gclip (examples/charlm.jl:116-118,148-152): the step is scaled by min(1, gclip / gnorm) with
        gnorm = sqrt(sum over params of sum(abs2, g)) of the reduced gradients -- computed and applied on the device
        (ag_multi_sumabs2, ag_multi_axpy_scaled), so a clipped step has no host round trip and replays as a graph."""
        from .core import value
        from . import lazy
        if gclip is not None and gclip > 0:
            reduced = self.update(0.0, grads, params) if self.world > 1 else \
                [g.full() if isinstance(g, Sparse) else g for g in grads]
            return self._clipped_update(alpha, reduced, params, float(gclip))
        dense = [g.full() if isinstance(g, Sparse) else g for g in grads]
        if self.world > 1:
            # a Param without a gradient on THIS rank (`nothing`, src/core.jl:215-216) still takes part in the sum: the
            # segment layout of the collective must not depend on the data of one rank
            dense = [D.zeros(value(p).shape, value(p).dtype) if g is None else g for g, p in zip(dense, params)]
        lazy.flush()        # the collective and the update work in place
        live = [(D.materialize(g), value(p)) for g, p in zip(dense, params) if g is not None]
        if not live:
            return dense
        self._check_layout([(g.size, str(g.dtype)) for g, _ in live])
        dt = live[0][0].dtype
        nbytes = sum(-(-g.size // 4) * 4 for g, _ in live) * dt.itemsize
        fits = len(live) <= 16 and all(g.dtype == dt and w.dtype == dt for g, w in live)
        # the peer-memory kernel wins only in its flag-in-data form (Float32, <= 1 MB); beyond that NCCL is faster
        p2p_ok = self.p2p_cap and dt == np.float32 and nbytes <= min(self.p2p_cap, 1 << 20)
        if fits and (self.world == 1 or p2p_ok):
            n = len(live)
            src = (C.c_void_p * n)(*[g.ptr for g, _ in live])
            dst = (C.c_void_p * n)(*[w.ptr for _, w in live])
            sizes = (C.c_int64 * n)(*[g.size for g, _ in live])
            if self.world == 1:
                check(lib.ag_multi_axpy(live[0][0].dt, n, float(alpha), src, dst, sizes))
            else:
                check(lib.ag_p2p_allreduce_multi(live[0][0].dt, n, src, dst, sizes, float(alpha)))
            return dense
        if fits:
            # NCCL path with multi-tensor pack / update: 3 launches (pack, allreduce, axpy!) instead of 2k+1
            n = len(live)
            sizes_l = [g.size for g, _ in live]
            tot = sum(sizes_l)
            if self._bucket is None or self._bucket.size != tot or self._bucket.dtype != dt:
                self._bucket = DArray.empty((tot,), dt)
            b, off, views = self._bucket, 0, []
            for g, _ in live:
                views.append(DArray(b.buf, b.ptr + off * dt.itemsize, g.shape, dt))
                off += g.size
            src = (C.c_void_p * n)(*[g.ptr for g, _ in live])
            bkt = (C.c_void_p * n)(*[v.ptr for v in views])
            dst = (C.c_void_p * n)(*[w.ptr for _, w in live])
            sizes = (C.c_int64 * n)(*sizes_l)
            check(lib.ag_multi_copy(b.dt, n, src, bkt, sizes))
            check(lib.ag_allreduce_sum(b.dt, D.vptr(b), tot))
            check(lib.ag_multi_axpy(b.dt, n, float(alpha), bkt, dst, sizes))
            it = iter(views)
            return [None if g is None else next(it) for g in dense]
        reduced = self.allreduce_grads(dense)
        for g, p in zip(reduced, params):
            if g is not None:
                D.axpy_(alpha, g, value(p))
        return reduced

    def _check_layout(self, layout):
        """Every rank must hand the collective the same segments (count, sizes, eltype) and so make the same choice
        between the peer-memory kernel and NCCL: a mismatch would otherwise end in a hang inside the kernel.  The
        layouts are compared once per distinct layout (one small host-side allgather), not per step."""
        if self.world == 1 or self._allgather is None:
            return
        import hashlib
        h = hashlib.sha1(repr(layout).encode()).digest()
        if h in self._layouts_ok:
            return
        got = self._allgather(h)
        if any(x != h for x in got):
            raise D.L.DeviceError("DataParallel.update: ranks disagree on the gradient layout (this rank: %d segments %s); "
                                  "every rank must pass the same Params in the same order" % (len(layout), layout[:4]))
        self._layouts_ok.add(h)

    def _clipped_update(self, alpha, dense, params, gclip):
        from .core import value
        from . import lazy
        lazy.flush()
        live = [(D.materialize(g), value(p)) for g, p in zip(dense, params) if g is not None]
        if not live:
            return dense
        dt = live[0][0].dtype
        if len(live) > 16 or any(g.dtype != dt or w.dtype != dt for g, w in live):
            raise D.L.UnsupportedOnDevice("clipped update: at most 16 tensors of one eltype")
        n = len(live)
        src = (C.c_void_p * n)(*[g.ptr for g, _ in live])
        dst = (C.c_void_p * n)(*[w.ptr for _, w in live])
        sizes = (C.c_int64 * n)(*[g.size for g, _ in live])
        nrm2 = DArray.empty((), dt)
        check(lib.ag_multi_sumabs2(nrm2.dt, n, src, sizes, D.vptr(nrm2)))
        # scale = min(1, gclip / sqrt(nrm2)) as a (deferred, fused) scalar chain; gnorm = 0 gives min(1, Inf) = 1
        scale = D.binary_fwd("min", D.binary_fwd("div", gclip, D.unary_fwd("sqrt", nrm2)), 1.0)
        self.last_gnorm2 = nrm2
        check(lib.ag_multi_axpy_scaled(nrm2.dt, n, float(alpha), D.vptr(scale), src, dst, sizes))
        return dense

    def close(self):
        check(lib.ag_comm_destroy())
