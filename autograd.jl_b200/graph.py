"""CUDA-graph replay of one recorded step, events and pinned staging buffers.

The reference dispatches every primitive and every gradient rule from the host on each step
(src/core.jl:64-79,156-169).  On device the same op sequence is enqueued once under stream capture
and replayed with one launch: the tape logic (host) runs at capture time only.
"""
import ctypes as C

import numpy as np

from ._lib import lib, check
from . import lazy as _lazy


class StepGraph:
    """with StepGraph() as g: <enqueue one step>;  then g.launch() replays it."""

    def __init__(self):
        self.handle = None
        self.keep = []      # objects whose device buffers the graph reads/writes

    def __enter__(self):
        _lazy.flush()       # nothing enqueued before the capture may run inside it
        check(lib.ag_graph_begin())
        return self

    def __exit__(self, et, ev, tb):
        h = C.c_uint64()
        if et is None:
            _lazy.flush()   # everything the step defined belongs to the captured step
        st = lib.ag_graph_end(C.byref(h))
        if et is None:
            check(st)
            self.handle = h.value
        return False

    def launch(self):
        _lazy.flush()
        check(lib.ag_graph_launch(C.c_uint64(self.handle)))

    def destroy(self):
        if self.handle is not None:
            check(lib.ag_graph_destroy(C.c_uint64(self.handle)))
            self.handle = None
        self.keep = []


class Event:
    def __init__(self):
        h = C.c_uint64()
        check(lib.ag_event_create(C.byref(h)))
        self.h = h.value

    def record(self):
        _lazy.flush()       # the event marks program order
        check(lib.ag_event_record(C.c_uint64(self.h)))
        return self

    def elapsed_ms_since(self, start):
        ms = C.c_float()
        check(lib.ag_event_elapsed_ms(C.c_uint64(start.h), C.c_uint64(self.h), C.byref(ms)))
        return ms.value

    def __del__(self):
        try:
            lib.ag_event_destroy(C.c_uint64(self.h))
        except Exception:
            pass


class _Pinned:
    def __init__(self, nbytes):
        p = C.c_void_p()
        check(lib.ag_host_alloc(int(nbytes), C.byref(p)))
        self.ptr = p.value

    def __del__(self):
        try:
            lib.ag_host_free(C.c_void_p(self.ptr))
        except Exception:
            pass


def pinned_empty(shape, dtype):
    """Fortran-order numpy array backed by pinned host memory (for async h2d/d2h)."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape, dtype=np.int64)) if len(shape) else 1
    owner = _Pinned(max(n, 1) * dtype.itemsize)
    buf = (C.c_char * (max(n, 1) * dtype.itemsize)).from_address(owner.ptr)
    a = np.frombuffer(buf, dtype=dtype, count=n).reshape(shape, order="F")
    a_owner = a.view(_PinnedArray)
    a_owner._owner = owner
    return a_owner


class _PinnedArray(np.ndarray):
    _owner = None

    def __array_finalize__(self, obj):
        if obj is not None and getattr(obj, "_owner", None) is not None:
            self._owner = obj._owner

    @property
    def host_ptr(self):
        return self.ctypes.data
