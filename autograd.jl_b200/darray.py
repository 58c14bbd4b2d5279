"""DArray -- the device array type that sits inside Param/Result, and the raw (unboxed) array
operations the tape runtime dispatches to.

In the reference this layer is Julia Base broadcast / mapreduce / LinearAlgebra on `Array`
(SURVEY.md layer L0; src/core.jl:66 calls `f(novalue...)` on the unboxed values).  Here every
operation is one call into libagb200.so.  Layout is COLUMN-MAJOR with Julia's broadcasting rule
(align leading dimensions, pad trailing 1s).  A 0-dimensional DArray is a device scalar: `sum(x)`
stays on the device so a whole step can be enqueued (and graph-captured) without a host sync.
"""
import ctypes as C
import math

import numpy as np

from . import _lib as L
from ._lib import lib, check

_DT = {np.dtype(np.float32): L.F32, np.dtype(np.float64): L.F64}

# op ids (include/agb200.h)
U = dict(identity=0, neg=1, abs=2, abs2=3, exp=4, log=5, tanh=6, sigm=7, sqrt=8, sin=9, cos=10, tan=11,
         sinh=12, cosh=13, asin=14, acos=15, atan=16, asinh=17, acosh=18, atanh=19, exp2=20, exp10=21,
         expm1=22, log2=23, log10=24, log1p=25, cbrt=26, sign=27, relu=28, one=29,
         erf=30, erfc=31, erfcx=32, erfinv=33, erfcinv=34, gamma=35, lgamma=36, digamma=37, trigamma=38,
         besselj0=39, besselj1=40, bessely0=41, bessely1=42)
B = dict(add=0, sub=1, mul=2, div=3, max=4, min=5, pow=6, eq=7, lt=8, le=9, gt=10, ge=11, atan2=12, hypot=13)
R = dict(id=0, abs2=1, abs=2)

_initialised = [False]


def init(device=0):
    check(lib.ag_init(int(device)))
    _initialised[0] = True


def _require_init():
    if not _initialised[0]:
        init(0)


def sync():
    _lazy.flush()
    check(lib.ag_sync())


def _sync_stream():
    """Wait for the stream only (a pageable host source is about to be released); pending deferred operations
    cannot depend on a buffer that is being created, so they stay pending."""
    check(lib.ag_sync())


def launch_count():
    n = C.c_int64()
    check(lib.ag_launch_count(C.byref(n)))
    return n.value


def device_info():
    a = (C.c_int64 * 8)()
    check(lib.ag_device_info(a))
    k = ["device", "sms", "cc_major", "cc_minor", "total_bytes", "free_bytes", "pool_bytes", "live_bytes"]
    return dict(zip(k, list(a)))


class _Buffer:
    """Owns one pool block; freed (stream-ordered, no sync) when the last view dies."""
    __slots__ = ("ptr", "nbytes")

    def __init__(self, nbytes):
        _require_init()
        p = C.c_void_p()
        check(lib.ag_alloc(int(nbytes), C.byref(p)))
        self.ptr = p.value or 0
        self.nbytes = nbytes

    def __del__(self):
        try:
            if self.ptr:
                lib.ag_free(C.c_void_p(self.ptr))
                self.ptr = 0
        except Exception:
            pass


def isnum(x):
    return isinstance(x, (int, float, np.floating, np.integer)) and not isinstance(x, (bool, np.bool_))


def _prod(s):
    n = 1
    for v in s:
        n *= int(v)
    return n


class DArray:
    """Dense column-major device array (Float32/Float64).  `tr=True` marks a lazy adjoint of a
    2-d array (src/linearalgebra.jl:16): storage is that of the parent, `shape` is the logical
    (transposed) shape, and only matmul / materialize understand it."""
    __slots__ = ("buf", "ptr", "shape", "dtype", "tr", "__weakref__")
    __array_priority__ = 2000

    def __init__(self, buf, ptr, shape, dtype, tr=False):
        self.buf, self.ptr, self.shape, self.dtype, self.tr = buf, ptr, tuple(int(s) for s in shape), np.dtype(dtype), tr

    # ---- construction / transfer
    @staticmethod
    def empty(shape, dtype):
        dtype = np.dtype(dtype)
        if dtype not in _DT:
            raise L.UnsupportedOnDevice("DArray eltype %s (Float32/Float64 only)" % dtype)
        shape = tuple(int(s) for s in shape)
        buf = _Buffer(max(_prod(shape), 1) * dtype.itemsize)
        return DArray(buf, buf.ptr, shape, dtype)

    @staticmethod
    def from_numpy(a, dtype=None):
        a = np.asarray(a, dtype=dtype)
        if a.dtype not in _DT:
            a = a.astype(np.float64)
        out = DArray.empty(a.shape, a.dtype)
        if a.size:
            flat = np.ascontiguousarray(a.reshape(-1, order="F"))
            check(lib.ag_copy_h2d(C.c_void_p(out.ptr), flat.ctypes.data_as(C.c_void_p), flat.nbytes))
            _sync_stream()   # `flat` is pageable and about to be released
        return out

    def copy_from_host(self, host_ptr, nbytes):
        """Async h2d from a pinned staging buffer (ag_host_alloc)."""
        _lazy.flush()               # pending operations must read the old contents first
        check(lib.ag_copy_h2d(C.c_void_p(self.ptr), C.c_void_p(host_ptr), int(nbytes)))

    def to_numpy(self):
        src = materialize(self)
        out = np.empty(src.size, dtype=src.dtype)
        if src.size:
            check(lib.ag_copy_d2h(out.ctypes.data_as(C.c_void_p), C.c_void_p(src.ptr), out.nbytes))
        if src.shape == ():
            return out.reshape(())
        return out.reshape(src.shape, order="F")

    def item(self):
        return self.to_numpy().reshape(-1)[0]

    def __float__(self):
        if self.size != 1:
            raise TypeError("only size-1 DArrays convert to a Number")
        return float(self.item())

    # ---- metadata
    @property
    def size(self):
        return _prod(self.shape)

    @property
    def ndim(self):
        return len(self.shape)

    @property
    def nbytes(self):
        return self.size * self.dtype.itemsize

    @property
    def dt(self):
        return _DT[self.dtype]

    def __repr__(self):
        return "DArray{%s}%s%s@0x%x" % (self.dtype.name, self.shape, "'" if self.tr else "", self.ptr)

    def reshape(self, shape):
        shape = tuple(int(s) for s in (shape if isinstance(shape, (tuple, list)) else (shape,)))
        if shape == self.shape and not self.tr:
            return self                   # same dims: the same data (a pending value stays pending)
        src = materialize(self)
        if _prod(shape) != src.size:
            raise L.DimensionMismatch("reshape %s -> %s" % (src.shape, shape))
        return DArray(src.buf, src.ptr, shape, src.dtype)

    # raw arithmetic sugar (the unboxed array also supports + - .* ./ like a Julia Array)
    def __add__(self, o): return _ops().add(self, o)
    def __radd__(self, o): return _ops().add(o, self)
    def __sub__(self, o): return _ops().sub(self, o)
    def __rsub__(self, o): return _ops().sub(o, self)
    def __mul__(self, o): return _ops().mul(self, o)
    def __rmul__(self, o): return _ops().mul(o, self)
    def __truediv__(self, o): return _ops().div(self, o)
    def __rtruediv__(self, o): return _ops().div(o, self)
    def __pow__(self, o): return _ops().pow_(self, o)
    def __neg__(self): return _ops().neg(self)
    def __matmul__(self, o): return _ops().matmul(self, o)
    def __rmatmul__(self, o): return _ops().matmul(o, self)
    def __getitem__(self, i): return _ops().getindex(self, *(i if isinstance(i, tuple) else (i,)))

    @property
    def T(self): return _ops().adjoint(self)


def _ops():
    from . import rules
    return rules


def vptr(x):
    return C.c_void_p(x.ptr if x is not None else None)


def zeros(shape, dtype):
    out = DArray.empty(shape, dtype)
    if out.size:
        check(lib.ag_fill(vptr(out), out.size, out.dt, 0.0))
    return out


def full_like_shape(shape, dtype, v):
    """fill(v, shape): read-only by convention (results that are updated in place start from `zeros`)."""
    if _lazy.enabled[0] and np.dtype(dtype) in _DT:
        return _lazy.fill(shape, dtype, v)
    return _full_like_shape_eager(shape, dtype, v)


def _full_like_shape_eager(shape, dtype, v):
    out = DArray.empty(shape, dtype)
    if out.size:
        check(lib.ag_fill(vptr(out), out.size, out.dt, float(v)))
    return out


def materialize(x):
    """Resolve a lazy adjoint into dense storage (permutedims)."""
    if not x.tr:
        return x
    cols, rows = x.shape           # logical shape of the adjoint; storage is (rows x cols)^T = (cols', rows')
    out = DArray.empty(x.shape, x.dtype)
    # storage holds parent P of shape (rows_p, cols_p) = (x.shape[1], x.shape[0])
    check(lib.ag_transpose(x.dt, vptr(x), x.shape[1], x.shape[0], vptr(out)))
    return out


def copy(x):
    x = materialize(x)
    out = DArray.empty(x.shape, x.dtype)
    check(lib.ag_copy_d2d(vptr(out), vptr(x), x.nbytes))
    return out


# ---- elementwise ---------------------------------------------------------------------------------

def unary_fwd(op, x):
    if _lazy.enabled[0]:
        return _lazy.unary_fwd(op, x)
    return _unary_fwd_eager(op, x)


def _unary_fwd_eager(op, x):
    x = materialize(x)
    y = DArray.empty(x.shape, x.dtype)
    check(lib.ag_unary_fwd(U[op], x.dt, vptr(x), vptr(y), x.size))
    return y


def unary_bwd(op, dy, x, y, shape=None, dtype=None):
    """dx = dy .* g(x, y).  dy: DArray of the full shape, 0-dim DArray, or host Number."""
    if not _lazy.enabled[0]:
        return _unary_bwd_eager(op, dy, x, y, shape, dtype)
    ref = x if x is not None else y
    if ref is None:
        ref = dy
    shape = tuple(ref.shape if shape is None else shape)
    dtype = ref.dtype if dtype is None else dtype
    n = _prod(shape)

    def fit(v, what):
        if v is None or isnum(v) or v.shape == shape:
            return v
        if v.size != n:
            raise L.DimensionMismatch("unary_bwd: %s %s vs %s" % (what, v.shape, shape))
        return materialize(v).reshape(shape)
    if not isnum(dy) and not (dy.size == 1 and n != 1):
        dy = fit(dy, "dy")
    return _lazy.unary_bwd(op, dy, fit(x, "x"), fit(y, "y"), shape, dtype)


def _unary_bwd_eager(op, dy, x, y, shape=None, dtype=None):
    ref = x if x is not None else y
    if ref is None:
        ref = dy
    shape = ref.shape if shape is None else shape
    dtype = ref.dtype if dtype is None else dtype
    dx = DArray.empty(shape, dtype)
    n = dx.size
    if isnum(dy):
        mode, dyp, dys = 2, None, float(dy)
    elif dy.size == 1 and n != 1:
        mode, dyp, dys = 1, dy, 0.0
    else:
        if dy.size != n:
            raise L.DimensionMismatch("unary_bwd: dy %s vs %s" % (dy.shape, shape))
        mode, dyp, dys = 0, materialize(dy), 0.0
    check(lib.ag_unary_bwd(U[op], _DT[np.dtype(dtype)], vptr(dyp), mode, dys,
                           vptr(materialize(x)) if x is not None else None,
                           vptr(materialize(y)) if y is not None else None, vptr(dx), n))
    return dx


def shape_of(x):
    return () if isnum(x) else x.shape


def bcast_shape(*shapes):
    nd = max((len(s) for s in shapes), default=0)
    out = [1] * nd
    for s in shapes:
        for d, e in enumerate(s):
            if e != 1:
                if out[d] != 1 and out[d] != e:
                    raise L.DimensionMismatch("arrays could not be broadcast to a common size: %s" % (shapes,))
                out[d] = e
    # an extent-0 dimension broadcast against 1 stays 0
    for s in shapes:
        for d, e in enumerate(s):
            if e == 0:
                out[d] = 0
    return tuple(out)


def _strides_for(shape, out_shape):
    """Column-major element strides of an operand of `shape` viewed in `out_shape` (0 = broadcast)."""
    st, acc = [], 1
    for d in range(len(out_shape)):
        e = shape[d] if d < len(shape) else 1
        st.append(0 if (e == 1 and out_shape[d] != 1) else acc)
        acc *= e
    return st


def collapse(out_shape, strides_list):
    """Drop extent-1 dims and merge adjacent dims that every operand traverses contiguously."""
    dims = [d for d in range(len(out_shape)) if out_shape[d] != 1]
    if not dims:
        return [1], [[0] for _ in strides_list]
    shp = [out_shape[dims[0]]]
    sts = [[s[dims[0]]] for s in strides_list]
    for d in dims[1:]:
        e = out_shape[d]
        if all(s[d] == st[-1] * shp[-1] for s, st in zip(strides_list, sts)):
            shp[-1] *= e
        else:
            shp.append(e)
            for s, st in zip(strides_list, sts):
                st.append(s[d])
    return shp, sts


def _i64arr(v):
    return (C.c_int64 * len(v))(*[int(x) for x in v])


def _common_dtype(*xs):
    dts = {x.dtype for x in xs if isinstance(x, DArray)}
    if len(dts) != 1:
        raise L.UnsupportedOnDevice("mixed or missing device eltypes %s" % (dts,))
    return dts.pop()


def binary_fwd(op, a, b):
    """op.(a, b) with Julia broadcasting; a, b: DArray or host Number (at least one DArray)."""
    if _lazy.enabled[0]:
        dtype = _common_dtype(a, b)
        out_shape = bcast_shape(shape_of(a), shape_of(b))
        if len(out_shape) <= 4 and _prod(out_shape):
            return _lazy.binary_fwd(op, a, b, out_shape, dtype)
    return _binary_fwd_eager(op, a, b)


def _binary_fwd_eager(op, a, b):
    a = materialize(a) if isinstance(a, DArray) else a
    b = materialize(b) if isinstance(b, DArray) else b
    dtype = _common_dtype(a, b)
    out_shape = bcast_shape(shape_of(a), shape_of(b))
    y = DArray.empty(out_shape, dtype)
    if y.size == 0:
        return y
    sl = [_strides_for(shape_of(x), out_shape) if isinstance(x, DArray) else [0] * len(out_shape) for x in (a, b)]
    oshape = list(out_shape) if out_shape else [1]
    if not out_shape:
        sl = [[0], [0]]
    shp, sts = collapse(oshape, sl + [_strides_for(out_shape, out_shape) if out_shape else [0]])
    if len(shp) > 4:
        raise L.UnsupportedOnDevice("broadcast needs %d collapsed dims (max 4)" % len(shp))
    check(lib.ag_binary_fwd(B[op], _DT[dtype],
                            vptr(a) if isinstance(a, DArray) else None, 0.0 if isinstance(a, DArray) else float(a), _i64arr(sts[0]),
                            vptr(b) if isinstance(b, DArray) else None, 0.0 if isinstance(b, DArray) else float(b), _i64arr(sts[1]),
                            vptr(y), len(shp), _i64arr(shp)))
    return y


def binary_bwd(op, argno, dy, x1, x2, y):
    """dy .* d(op)/d(x_argno), full broadcast shape (before unbroadcast)."""
    if _lazy.enabled[0]:
        vs = [v for v in (dy, x1, x2, y) if v is not None]
        dtype = _common_dtype(*vs)
        out_shape = bcast_shape(*[shape_of(v) for v in vs])
        if len(out_shape) <= 4 and _prod(out_shape):
            form = _lazy._bb_form(op, argno)
            named = {"dy": dy, "x1": x1, "x2": x2, "y": y}
            if form[0] != "binc" and any(named[k] is None for k in form[2:]):
                raise L.DimensionMismatch("binary_bwd: %s needs %s" % (op, "/".join(form[2:])))
            return _lazy.binary_bwd(op, argno, dy, x1, x2, y, out_shape, dtype)
    return _binary_bwd_eager(op, argno, dy, x1, x2, y)


def _binary_bwd_eager(op, argno, dy, x1, x2, y):
    arrs = [materialize(v) if isinstance(v, DArray) else v for v in (dy, x1, x2, y)]
    dy, x1, x2, y = arrs
    dtype = _common_dtype(*[v for v in arrs if v is not None])
    out_shape = bcast_shape(*[shape_of(v) for v in arrs if v is not None])
    dx = DArray.empty(out_shape, dtype)
    if dx.size == 0:
        return dx
    if isnum(dy):
        dy = full_like_shape((), dtype, dy)
    oshape = list(out_shape) if out_shape else [1]

    def st(v):
        if isinstance(v, DArray):
            return _strides_for(v.shape, out_shape) if out_shape else [0]
        return [0] * len(oshape)
    sl = [st(dy), st(x1), st(x2), st(y), (_strides_for(out_shape, out_shape) if out_shape else [0])]
    shp, sts = collapse(oshape, sl)
    if len(shp) > 4:
        raise L.UnsupportedOnDevice("broadcast needs %d collapsed dims (max 4)" % len(shp))

    def pv(v):
        return vptr(v) if isinstance(v, DArray) else None

    def sv(v):
        return float(v) if isnum(v) else 0.0
    check(lib.ag_binary_bwd(B[op], argno, _DT[dtype], pv(dy), _i64arr(sts[0]),
                            pv(x1), sv(x1), _i64arr(sts[1]), pv(x2), sv(x2), _i64arr(sts[2]),
                            pv(y), _i64arr(sts[3]), vptr(dx), len(shp), _i64arr(shp)))
    return dx


# ---- reductions -------------------------------------------------------------------------------------

def sum_all(x, mapname="id"):
    if x.size and _lazy.is_pending(x):
        r = _lazy.reduce(x, 1, x.size, 1, mapname)
        if r is not None:
            return r.reshape(())
    x = materialize(x)
    out = DArray.empty((), x.dtype)
    if x.size == 0:
        check(lib.ag_fill(vptr(out), 1, out.dt, 0.0))
        return out
    check(lib.ag_sum_all(R[mapname], x.dt, vptr(x), x.size, vptr(out)))
    return out


def sum_dims(x, dims, mapname="id"):
    """sum(map, x; dims) keeping reduced dims with extent 1 (Julia).  dims: iterable of 0-based dims."""
    x = materialize(x)
    dims = sorted({int(d) for d in dims if int(d) < x.ndim})
    shape = list(x.shape)
    out_shape = [1 if d in dims else e for d, e in enumerate(shape)]
    if not dims or all(shape[d] == 1 for d in dims):
        if mapname == "id":
            return x.reshape(out_shape) if dims else x
        return unary_fwd(mapname, x).reshape(out_shape)
    # maximal runs of consecutive reduced dims -> one ag_sum_dim each
    runs, i = [], 0
    while i < len(dims):
        j = i
        while j + 1 < len(dims) and dims[j + 1] == dims[j] + 1:
            j += 1
        runs.append((dims[i], dims[j]))
        i = j + 1
    if len(runs) == 1 and _lazy.is_pending(x):
        lo, hi = runs[0]
        r = _lazy.reduce(x, _prod(shape[:lo]), _prod(shape[lo:hi + 1]), _prod(shape[hi + 1:]), mapname)
        if r is not None:
            if _lazy.is_pending(r) and r.shape == tuple(out_shape):
                return r                  # a column sum that stays inside the chain (lazy.reduce)
            return r.reshape(out_shape)
    if len(runs) == 1 and x.size:
        lo, hi = runs[0]
        if _lazy.is_pending(x):
            x.ptr                         # the chain below the sum is evaluated now; the sum itself waits
        r = _lazy.reduce_deferred(x, _prod(shape[:lo]), _prod(shape[lo:hi + 1]), _prod(shape[hi + 1:]), mapname, out_shape)
        if r is not None:
            return r                      # evaluated later, batched with the other reductions of its layout
    cur, cur_shape, m = x, shape, mapname
    for lo, hi in runs:
        inner = _prod(cur_shape[:lo])
        red = _prod(cur_shape[lo:hi + 1])
        outer = _prod(cur_shape[hi + 1:])
        new_shape = [1 if lo <= d <= hi else e for d, e in enumerate(cur_shape)]
        out = DArray.empty(new_shape, x.dtype)
        if out.size:
            check(lib.ag_sum_dim(R[m], x.dt, vptr(cur), inner, red, outer, vptr(out)))
        cur, cur_shape, m = out, new_shape, "id"
    return cur


def reduce_dims(x, dims, op):
    """maximum / minimum / prod over dims (keep-dims, Julia): op in {'max','min','prod'}."""
    x = materialize(x)
    code = {"max": 0, "min": 1, "prod": 2}[op]
    dims = sorted({int(d) for d in dims if int(d) < x.ndim}) if dims is not None else list(range(x.ndim))
    cur, cur_shape = x, list(x.shape)
    if not dims:
        return x
    runs, i = [], 0
    while i < len(dims):
        j = i
        while j + 1 < len(dims) and dims[j + 1] == dims[j] + 1:
            j += 1
        runs.append((dims[i], dims[j]))
        i = j + 1
    for lo, hi in runs:
        inner, red, outer = _prod(cur_shape[:lo]), _prod(cur_shape[lo:hi + 1]), _prod(cur_shape[hi + 1:])
        new_shape = [1 if lo <= d <= hi else e for d, e in enumerate(cur_shape)]
        out = DArray.empty(new_shape, x.dtype)
        if out.size:
            check(lib.ag_reduce_dim(code, x.dt, vptr(cur), inner, red, outer, vptr(out)))
        cur, cur_shape = out, new_shape
    return cur


def permutedims(x, perm):
    x = materialize(x)
    perm = [int(p) for p in perm]
    if sorted(perm) != list(range(x.ndim)):
        raise L.DimensionMismatch("permutedims: %s is not a permutation of %d dims" % (perm, x.ndim))
    out = DArray.empty([x.shape[p] for p in perm], x.dtype)
    if out.size:
        check(lib.ag_permute(x.dt, vptr(x), x.ndim, _i64arr(x.shape), (C.c_int32 * x.ndim)(*perm), vptr(out)))
    return out


def bcast_to(a, shape):
    """Broadcast-copy `a` to `shape` (dy .* one.(x) with a keep-dims dy, src/base.jl:245)."""
    shape = tuple(int(e) for e in shape)
    if bcast_shape(a.shape, shape) != shape:
        raise L.DimensionMismatch("cannot broadcast %s to %s" % (a.shape, shape))
    if _lazy.enabled[0] and len(shape) <= 4 and _prod(shape):
        return _lazy.bcast_to(a, shape)
    return _bcast_to_eager(a, shape)


def _bcast_to_eager(a, shape):
    # through the binary map: a .* 1 evaluated on the target shape
    a = materialize(a)
    out = DArray.empty(shape, a.dtype)
    if out.size == 0:
        return out
    st_a = _strides_for(a.shape, shape)
    shp, sts = collapse(list(shape), [st_a, _strides_for(shape, shape)])
    if len(shp) > 4:
        raise L.UnsupportedOnDevice("broadcast needs %d collapsed dims (max 4)" % len(shp))
    check(lib.ag_binary_fwd(B["mul"], a.dt, vptr(a), 0.0, _i64arr(sts[0]), None, 1.0, _i64arr([0] * len(shp)),
                            vptr(out), len(shp), _i64arr(shp)))
    return out


def add(a, b):
    if a.shape != b.shape:
        raise L.DimensionMismatch("addto!: %s vs %s" % (a.shape, b.shape))
    if a.dtype != b.dtype:
        raise L.UnsupportedOnDevice("addto!: eltypes %s, %s" % (a.dtype, b.dtype))
    if _lazy.enabled[0] and a.size:
        return _lazy.binary_fwd("add", a, b, a.shape, a.dtype)
    a, b = materialize(a), materialize(b)
    out = DArray.empty(a.shape, a.dtype)
    check(lib.ag_add(a.dt, vptr(a), vptr(b), vptr(out), a.size))
    return out


def axpy_(alpha, x, y):
    """y .+= alpha .* x in place (dense)."""
    x = materialize(x)
    if x.size != y.size:
        raise L.DimensionMismatch("axpy!: %s vs %s" % (x.shape, y.shape))
    if y.tr:
        raise L.UnsupportedOnDevice("axpy! into a lazy adjoint")
    _lazy.flush()
    check(lib.ag_axpy(y.dt, float(alpha), vptr(x), vptr(y), y.size))
    return y


def scale_(alpha, x):
    _lazy.flush()
    check(lib.ag_scale(x.dt, float(alpha), vptr(x), x.size))
    return x


# ---- matmul ---------------------------------------------------------------------------------------

def _as_matrix(x, left):
    """(storage DArray, trans flag, rows, cols, ld) of a matmul operand."""
    if x.ndim == 2:
        if x.tr:
            return x, 1, x.shape[0], x.shape[1], x.shape[1]   # storage is (cols x rows)
        return x, 0, x.shape[0], x.shape[1], max(x.shape[0], 1)
    if x.ndim == 1:
        return x, 0, x.shape[0], 1, max(x.shape[0], 1)
    raise L.DimensionMismatch("matmul operand with %d dims" % x.ndim)


def matmul(a, b):
    """a*b for matrices/vectors (src/base.jl:26 forward and both of its gradient products)."""
    dtype = _common_dtype(a, b)
    A, ta, m, ka, lda = _as_matrix(a, True)
    Bm, tb, kb, n, ldb = _as_matrix(b, False)
    if ka != kb:
        raise L.DimensionMismatch("matmul: A has dimensions (%d,%d) but B has dimensions (%d,%d)" % (m, ka, kb, n))
    out_shape = (m,) if (b.ndim == 1) else (m, n)
    if a.ndim == 1 and b.ndim == 2:
        out_shape = (m, n)
    if _lazy.enabled[0] and len(out_shape) == 2 and m and n and ka:
        A.ptr, Bm.ptr                       # operands must exist (a pending operand is evaluated now)
        return _lazy.matmul(A, ta, Bm, tb, m, n, ka, lda, ldb, out_shape, dtype)
    return _matmul_eager(A, Bm, (ta, tb, m, n, ka, lda, ldb), out_shape, dtype)


def _matmul_eager(A, Bm, op, out_shape, dtype):
    ta, tb, m, n, ka, lda, ldb = op
    c = DArray.empty(out_shape, dtype)
    if c.size == 0:
        return c
    if ka == 0:
        check(lib.ag_fill(vptr(c), c.size, c.dt, 0.0))
        return c
    check(lib.ag_gemm(_DT[np.dtype(dtype)], ta, tb, m, n, ka, 1.0, vptr(A), lda, vptr(Bm), ldb, 0.0, vptr(c), max(m, 1)))
    return c


def _gemm_epilogue_fused(A, Bm, op, dtype):
    """Would ag_gemm_epilogue run this product as one launch?"""
    ta, tb, m, n, ka, lda, ldb = op
    flag = C.c_int32(0)
    check(lib.ag_gemm_epilogue_query(_DT[np.dtype(dtype)], ta, tb, m, n, ka, vptr(A), lda, vptr(Bm), ldb, C.byref(flag)))
    return bool(flag.value)


def _gemm_group(jobs):
    """Independent products [(A, B, op, out_shape, dtype, bias or None, act, want_z)] handed over together
    (ag_gemm_group); returns [(z, a)] per job: z = op(A)*op(B) (.+ bias), a = act.(z) or None; z is None when an
    activation was requested without want_z."""
    if not jobs:
        return []
    dtype = np.dtype(jobs[0][4])
    if len(jobs) == 1 or any(np.dtype(j[4]) != dtype for j in jobs):
        out = []
        for A, Bm, op, shape, dt, bias, act, want_z in jobs:
            if bias is None and act == "identity":
                out.append((_matmul_eager(A, Bm, op, shape, dt), None))
            else:
                out.append(_gemm_epilogue(A, Bm, op, shape, dt, bias, act, want_z))
        return out
    descs = (L.GemmDesc * len(jobs))()
    res = []
    for d, (A, Bm, op, shape, dt, bias, act, want_z) in zip(descs, jobs):
        ta, tb, m, n, ka, lda, ldb = op
        c = DArray.empty(shape, dt)
        c2 = DArray.empty(shape, dt) if (want_z and act != "identity") else None
        d.transA, d.transB, d.M, d.N, d.K = ta, tb, m, n, ka
        d.A, d.lda, d.B, d.ldb = A.ptr, lda, Bm.ptr, ldb
        d.bias = bias.ptr if bias is not None else None
        d.act, d.C, d.C2, d.ldc = U[act], c.ptr, (c2.ptr if c2 is not None else None), max(m, 1)
        res.append((c, None) if act == "identity" else ((c, c2) if want_z else (None, c)))
    check(lib.ag_gemm_group(_DT[dtype], len(jobs), descs))
    return res


def _gemm_epilogue(A, Bm, op, out_shape, dtype, bias, act, want_z):
    """(z, a) = (op(A)*op(B) .+ bias, act.(z)) by one ag_gemm_epilogue call; z is None unless want_z (then the
    activation, if any, goes to a second array)."""
    ta, tb, m, n, ka, lda, ldb = op
    c = DArray.empty(out_shape, dtype)
    c2 = DArray.empty(out_shape, dtype) if (want_z and act != "identity") else None
    check(lib.ag_gemm_epilogue(_DT[np.dtype(dtype)], ta, tb, m, n, ka, vptr(A), lda, vptr(Bm), ldb, vptr(bias), U[act],
                               vptr(c), vptr(c2), max(m, 1)))
    if act == "identity":
        return c, None
    return (c, c2) if want_z else (None, c)


def gemm_engine():
    buf = C.create_string_buffer(32)
    check(lib.ag_gemm_last_engine(buf, 32))
    return buf.value.decode()


def set_gemm_engine(e):
    check(lib.ag_gemm_set_engine({"auto": 0, "ffma": 1, "tcgen05": 2}[e]))


def adjoint(x):
    if x.ndim == 1:
        return DArray(x.buf, x.ptr, (1, x.shape[0]), x.dtype)
    if x.ndim != 2:
        raise L.DimensionMismatch("adjoint of a %d-d array" % x.ndim)
    if x.shape[0] == 1 or x.shape[1] == 1:
        src = materialize(x)
        return DArray(src.buf, src.ptr, (x.shape[1], x.shape[0]), x.dtype)
    return DArray(x.buf, x.ptr, (x.shape[1], x.shape[0]), x.dtype, tr=not x.tr)


# ---- indexing ---------------------------------------------------------------------------------------

class IArray:
    """Device-resident Int64 index vector (0-based).  Keeping the indices of a batch on the device (and
    refreshing them in place with `update`) lets `W[:, ids]` and its Sparse gradient run without any host
    synchronisation, so a whole recurrent step can be captured and replayed as one CUDA graph.  The host copy is
    kept for the bookkeeping the tape does on the host (lengths, bounds)."""
    __slots__ = ("host", "buf", "lo", "hi")

    def __init__(self, host):
        self.host = np.ascontiguousarray(host, dtype=np.int64).reshape(-1)
        self.buf = _Buffer(max(self.host.nbytes, 8))
        self._upload()

    def _upload(self):
        self.lo = int(self.host.min()) if self.host.size else 0
        self.hi = int(self.host.max()) if self.host.size else -1
        if self.host.size:
            check(lib.ag_copy_h2d(C.c_void_p(self.buf.ptr), self.host.ctypes.data_as(C.c_void_p), self.host.nbytes))
            sync()

    def update(self, host):
        host = np.ascontiguousarray(host, dtype=np.int64).reshape(-1)
        if host.size != self.host.size:
            raise L.DimensionMismatch("IArray.update: length %d != %d" % (host.size, self.host.size))
        self.host = host
        self._upload()

    @property
    def size(self):
        return self.host.size

    @property
    def ptr(self):
        return self.buf.ptr

    def __len__(self):
        return self.host.size


def index_to_device(idx):
    """int64 numpy index vector -> device buffer (returned as a raw _Buffer)."""
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    buf = _Buffer(max(idx.nbytes, 8))
    if idx.size:
        check(lib.ag_copy_h2d(C.c_void_p(buf.ptr), idx.ctypes.data_as(C.c_void_p), idx.nbytes))
        _sync_stream()
    return buf


def gather(x, flat_idx, out_shape):
    x = materialize(x)
    out = DArray.empty(out_shape, x.dtype)
    if out.size:
        ib = index_to_device(flat_idx)
        check(lib.ag_gather(x.dt, vptr(x), C.c_void_p(ib.ptr), vptr(out), out.size))
    return out


def gather_cols(x, cols):
    x = materialize(x)
    if not isinstance(cols, IArray):
        cols = np.ascontiguousarray(cols, dtype=np.int64)
    out = DArray.empty((x.shape[0], cols.size), x.dtype)
    if out.size:
        ib = cols if isinstance(cols, IArray) else index_to_device(cols)
        check(lib.ag_gather_cols(x.dt, vptr(x), x.shape[0], C.c_void_p(ib.ptr), vptr(out), cols.size))
    return out


def cat_into(out, srcs, inner, outer):
    """out (freshly allocated) = cat of srcs = [(array, extent along the concatenated dimension)] in one launch."""
    n = len(srcs)
    ptrs = (C.c_void_p * n)(*[x.ptr if l else None for x, l in srcs])
    lens = (C.c_int64 * n)(*[l for _, l in srcs])
    check(lib.ag_cat(out.dt, n, ptrs, lens, vptr(out), inner, outer))
    return out


def multi_copy_flat(dt, segs):
    """Contiguous copies [(src pointer, dst pointer, elements)] of one eltype code, 16 per launch (ag_multi_copy): the
    flat concatenations of cat1d / Sparse blocks / index vectors (src/cat.jl:150-182, src/sparse.jl:57) as a handful
    of launches instead of one per block (config 4: 100 value blocks + 100 index blocks per step)."""
    segs = [s for s in segs if s[2]]
    for i in range(0, len(segs), 16):
        part = segs[i:i + 16]
        n = len(part)
        if n == 1:
            check(lib.ag_slab_copy(dt, C.c_void_p(part[0][0]), part[0][2], 0, C.c_void_p(part[0][1]), part[0][2], 0, 1,
                                   part[0][2], 1))
            continue
        src = (C.c_void_p * n)(*[p_[0] for p_ in part])
        dst = (C.c_void_p * n)(*[p_[1] for p_ in part])
        sizes = (C.c_int64 * n)(*[p_[2] for p_ in part])
        check(lib.ag_multi_copy(dt, n, src, dst, sizes))


def concat_index(blocks):
    """Concatenate IArray blocks on the device (raw 8-byte copies); returns (buffer, n, lo, hi)."""
    n = sum(b.size for b in blocks)
    buf = _Buffer(max(n * 8, 8))
    off, segs = 0, []
    for b in blocks:
        segs.append((b.ptr, buf.ptr + off * 8, b.size))
        off += b.size
    multi_copy_flat(L.F64, segs)
    return buf, n, min(b.lo for b in blocks), max(b.hi for b in blocks)


def scatter_add_dev_(dest, idx_buf, n, lo, hi, vals, block=1):
    """scatter_add_ with destinations already on the device (bounds known from the host bookkeeping)."""
    if n == 0:
        return dest
    vals = materialize(vals)
    if vals.size != n * block:
        raise L.DimensionMismatch("scatter-add: %d values for %d destinations" % (vals.size, n * block))
    if lo < 0 or (hi + 1) * block > dest.size:
        raise IndexError("BoundsError")
    _lazy.flush()
    check(lib.ag_scatter_add(dest.dt, vptr(dest), dest.size, C.c_void_p(idx_buf.ptr), vptr(vals), n, block))
    return dest


def scatter_add_(dest, flat_idx, vals, block=1):
    """dest[flat_idx[j]*block .. +block) += vals[j*block .. +block), repeated indices summed."""
    flat_idx = np.ascontiguousarray(flat_idx, dtype=np.int64)
    n = flat_idx.size
    if n == 0:
        return dest
    vals = materialize(vals)
    if vals.size != n * block:
        raise L.DimensionMismatch("scatter-add: %d values for %d destinations" % (vals.size, n * block))
    if flat_idx.min() < 0 or (flat_idx.max() + 1) * block > dest.size:
        raise IndexError("BoundsError")
    ib = index_to_device(flat_idx)
    _lazy.flush()
    check(lib.ag_scatter_add(dest.dt, vptr(dest), dest.size, C.c_void_p(ib.ptr), vptr(vals), n, block))
    return dest


def slab_copy(src, src_red, src_off, dst, dst_red, dst_off, inner, length, outer, fresh=False):
    """Raw strided copy into `dst`.  fresh=True: dst was allocated by the caller for this purpose and nothing pending
    can read it; otherwise dst is an existing array and everything pending is evaluated first (a pending operation
    reads its leaves later than program order says -- lazy.py safety rule 2)."""
    if not fresh:
        _lazy.flush()
    check(lib.ag_slab_copy(src.dt, vptr(src), src_red, src_off, vptr(dst), dst_red, dst_off, inner, length, outer))
    return dst


# ---- host scalar math (Numbers never touch the device) ----------------------------------------------
HOST_UNARY = dict(
    identity=lambda x: x, neg=lambda x: -x, abs=abs, abs2=lambda x: x * x, exp=math.exp, log=math.log,
    tanh=math.tanh, sigm=lambda x: 1.0 / (1.0 + math.exp(-x)), sqrt=math.sqrt, sin=math.sin, cos=math.cos,
    tan=math.tan, sinh=math.sinh, cosh=math.cosh, asin=math.asin, acos=math.acos, atan=math.atan,
    asinh=math.asinh, acosh=math.acosh, atanh=math.atanh, exp2=lambda x: 2.0 ** x, exp10=lambda x: 10.0 ** x,
    expm1=math.expm1, log2=math.log2, log10=math.log10, log1p=math.log1p,
    cbrt=lambda x: math.copysign(abs(x) ** (1.0 / 3.0), x), sign=lambda x: (x > 0) - (x < 0),
    relu=lambda x: x if x > 0 else 0 * x, one=lambda x: type(x)(1))


SP = dict(besselj=0, bessely=1, polygamma=2, invdigamma=3, dawson=4, erfi=5, airyai=6, airyaiprime=7, airybi=8,
          airybiprime=9, besseli=10, besselk=11, floor=12, ceil=13, round=14, trunc=15,
          exponent=16, significand=17, rem2pi=18)


def astype(x, dtype):
    """Float32 <-> Float64 copy of a device array (ag_cast); the same array when the eltype already matches."""
    x = materialize(x)
    dtype = np.dtype(dtype)
    if np.dtype(x.dtype) == dtype:
        return x
    out = DArray.empty(x.shape, dtype.type)
    if x.size:
        check(lib.ag_cast(x.dt, out.dt, vptr(x), vptr(out), x.size))
    return out


def argext(x, want_max):
    """0-based column-major position of the first maximum / minimum of x (ag_argext); read back to the host."""
    x = materialize(x)
    if x.size == 0:
        raise ValueError("ArgumentError: collection must be non-empty")
    out = _Buffer(8)
    check(lib.ag_argext(x.dt, vptr(x), x.size, 1 if want_max else 0, C.c_void_p(out.ptr)))
    host = np.zeros(1, dtype=np.int64)
    check(lib.ag_copy_d2h(host.ctypes.data_as(C.c_void_p), C.c_void_p(out.ptr), 8))
    sync()
    return int(host[0])


def inv_det(x, want_inv=True):
    """(inv(x) or None, (det, log|det|, sign)) of a square matrix in one launch (ag_inv_det); the three scalars are
    read back to the host."""
    x = materialize(x)
    if len(x.shape) != 2 or x.shape[0] != x.shape[1]:
        raise L.DimensionMismatch("matrix is not square: dimensions are %r" % (tuple(x.shape),))
    n = x.shape[0]
    inv = DArray.empty(x.shape, x.dtype) if want_inv else None
    info = DArray.empty((3,), x.dtype)
    check(lib.ag_inv_det(x.dt, vptr(x), n, vptr(inv) if want_inv else None, vptr(info)))
    d = info.to_numpy()
    return inv, (float(d[0]), float(d[1]), float(d[2]))


def special_param(name, m, x):
    """besselj(m, x) / bessely(m, x) of integer order, polygamma(m, x), invdigamma(x), dawson / erfi / the four Airy
    functions (m ignored) (ag_special_param): one pass."""
    x = materialize(x)
    out = DArray.empty(x.shape, x.dtype)
    if x.size:
        check(lib.ag_special_param(SP[name], x.dt, int(m), vptr(x), vptr(out), x.size))
    return out


def _special(name):
    """Host Numbers through scipy.special (the SpecialFunctions.jl counterpart); only when actually called."""
    def f(x):
        import scipy.special as sp
        table = dict(erf=sp.erf, erfc=sp.erfc, erfcx=sp.erfcx, erfinv=sp.erfinv, erfcinv=sp.erfcinv, gamma=sp.gamma,
                     lgamma=sp.gammaln, digamma=sp.digamma, trigamma=lambda v: sp.polygamma(1, v), besselj0=sp.j0,
                     besselj1=sp.j1, bessely0=sp.y0, bessely1=sp.y1)
        return float(table[name](x))
    return f


for _n in ("erf", "erfc", "erfcx", "erfinv", "erfcinv", "gamma", "lgamma", "digamma", "trigamma", "besselj0", "besselj1",
           "bessely0", "bessely1"):
    HOST_UNARY[_n] = _special(_n)
HOST_BINARY = dict(
    add=lambda a, b: a + b, sub=lambda a, b: a - b, mul=lambda a, b: a * b, div=lambda a, b: a / b,
    max=max, min=min, pow=lambda a, b: a ** b, eq=lambda a, b: float(a == b), lt=lambda a, b: float(a < b),
    le=lambda a, b: float(a <= b), gt=lambda a, b: float(a > b), ge=lambda a, b: float(a >= b),
    atan2=math.atan2, hypot=math.hypot)


from . import lazy as _lazy  # noqa: E402  (deferred elementwise evaluation; needs the definitions above)
