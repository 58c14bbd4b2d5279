"""Host handle of the C-side tape (csrc/tape.cu, ag_tape_*): the fast path for tapes made only of device primitives.

The reference records and sweeps its tape on the host (src/core.jl:64-129,153-170); here one foreign call records an
operation (and runs its forward kernel) and ONE call runs the whole backward sweep.  `CTape` mirrors the primitive names
of rules.py, so a model function written against a primitive namespace `m` (workloads.py) runs unchanged:

    t = CTape(); w = [t.param(p) for p in params]; loss = workloads.mnist_loss(t, w, t.const(x), t.const(y))
    t.backward(loss); g = t.grad(w[0])
"""
import ctypes as C

import numpy as np

from . import darray as D
from ._lib import lib, check

T_MATMUL, T_ADD, T_SUB, T_MUL, T_DIV_SCALAR, T_MUL_SCALAR, T_UNARY, T_SUM, T_SUMABS2, T_SUM_DIM = range(10)


class TVal:
    """A value of a C tape: (tape, id); shape and dtype are cached on the host."""
    __slots__ = ("tape", "id", "shape", "dtype")

    def __init__(self, tape, id_, shape, dtype):
        self.tape, self.id, self.shape, self.dtype = tape, id_, tuple(shape), np.dtype(dtype)


class CTape:
    def __init__(self):
        h = C.c_uint64()
        check(lib.ag_tape_new(C.byref(h)))
        self.h = h.value
        self._keep = []          # leaf arrays must outlive the tape

    def close(self):
        if self.h is not None:
            check(lib.ag_tape_free(C.c_uint64(self.h)))
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- leaves
    def _leaf(self, x, tracked):
        x = getattr(x, "value", x)            # a Param of the host tape is accepted as is
        x = D.materialize(x)
        self._keep.append(x)
        out = C.c_int32()
        check(lib.ag_tape_leaf(C.c_uint64(self.h), x.dt, D.vptr(x), x.ndim, D._i64arr(x.shape or (1,)), 1 if tracked else 0,
                               C.byref(out)))
        return TVal(self, out.value, x.shape, x.dtype)

    def param(self, x):
        return self._leaf(x, True)

    def const(self, x):
        return self._leaf(x, False)

    # ---- recording
    def _rec(self, op, arg, ins, scalar=0.0):
        ids = (C.c_int32 * len(ins))(*[v.id for v in ins])
        out = C.c_int32()
        check(lib.ag_tape_record(C.c_uint64(self.h), op, int(arg), len(ins), ids, float(scalar), C.byref(out)))
        p, nd, shp = C.c_void_p(), C.c_int32(), (C.c_int64 * 4)()
        check(lib.ag_tape_value(C.c_uint64(self.h), out.value, C.byref(p), C.byref(nd), shp))
        return TVal(self, out.value, tuple(shp[:nd.value]), ins[0].dtype)

    def _v(self, x):
        return x if isinstance(x, TVal) else self.const(x)

    def matmul(self, a, b): return self._rec(T_MATMUL, 0, [self._v(a), self._v(b)])
    def add(self, a, b): return self._rec(T_ADD, 0, [self._v(a), self._v(b)])
    def sub(self, a, b): return self._rec(T_SUB, 0, [self._v(a), self._v(b)])

    def mul(self, a, b):
        if D.isnum(a):
            return self._rec(T_MUL_SCALAR, 0, [self._v(b)], a)
        if D.isnum(b):
            return self._rec(T_MUL_SCALAR, 0, [self._v(a)], b)
        return self._rec(T_MUL, 0, [self._v(a), self._v(b)])

    def div(self, a, b):
        if not D.isnum(b):
            raise D.L.UnsupportedOnDevice("CTape: ./ by an array")
        return self._rec(T_DIV_SCALAR, 0, [self._v(a)], b)

    def _unary(self, name, x): return self._rec(T_UNARY, D.U[name], [self._v(x)])
    def neg(self, x): return self._unary("neg", x)
    def exp(self, x): return self._unary("exp", x)
    def log(self, x): return self._unary("log", x)
    def tanh(self, x): return self._unary("tanh", x)
    def sigm(self, x): return self._unary("sigm", x)
    def relu(self, x): return self._unary("relu", x)
    def abs2(self, x): return self._unary("abs2", x)
    def sumabs2(self, x): return self._rec(T_SUMABS2, 0, [self._v(x)])

    def sum_(self, x, dims=None):
        x = self._v(x)
        if dims is None:
            return self._rec(T_SUM, 0, [x])
        dims = (dims,) if isinstance(dims, (int, np.integer)) else tuple(dims)
        for d in dims:
            x = self._rec(T_SUM_DIM, int(d), [x])
        return x

    # ---- sweep
    def backward(self, result):
        check(lib.ag_tape_backward(C.c_uint64(self.h), result.id))

    def _view(self, ptr, v):
        return D.DArray(self, ptr, v.shape, v.dtype)       # `buf` = the tape: the view keeps it alive

    def grad(self, v):
        p = C.c_void_p()
        check(lib.ag_tape_grad(C.c_uint64(self.h), v.id, C.byref(p)))
        return None if not p.value else self._view(p.value, v)

    def value(self, v):
        p = C.c_void_p()
        check(lib.ag_tape_value(C.c_uint64(self.h), v.id, C.byref(p), None, None))
        return self._view(p.value, v)
