"""agref -- CPU restatement (numpy) of the AutoGrad.jl array hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import it.  The product package
(autograd.jl_b200/) never imports anything under oracle/.

What it restates (all citations are file:line under /root/reference):
  * Tape runtime ............ src/core.jl:3-41 (types), :60-79 (forw), :110-129 (Result/record!),
                              :134-171 (differentiate + backward sweep), :187-197 (findresult),
                              :201-232 (Param/@diff/value/grad/gradloss), :235-237 (gcnode)
  * unbroadcast ............. src/unbroadcast.jl:9-26
  * addto!/addtoindex! ...... src/addto.jl:14-31,41-52,91-136
  * Sparse .................. src/sparse.jl:15-83
  * getindex/ungetindex ..... src/getindex.jl:32-33,61-79,91-95
  * cat/uncat/cat1d ......... src/cat.jl:27-31,46-92,122-123,150-182
  * rules ................... src/base.jl:20-49,73-74,202-206,217-250,338-341; src/math.jl:9-74;
                              src/statistics.jl:20-35; src/linearalgebra.jl:16,87
  * gradcheck ............... test/gradcheck.jl:35-133

Pinning: the reference is Julia and Julia is not installed in this image, so the reference
itself cannot be executed here.  The oracle is pinned against every exact known-answer vector
the reference's own tests/docs hold for this path in
tests/test_oracle_known_answers.py; bit-level parity of floating-point results with Julia is
unpinned (Julia RNG inputs are not reproducible outside Julia) and all float comparisons are
tolerance based.

Conventions: arrays are numpy arrays interpreted COLUMN-MAJOR (Julia layout); broadcasting aligns
LEADING dimensions and pads trailing 1s (Julia), not trailing dimensions (numpy).  Indices given
to getindex are 0-based (host language is Python) but keep Julia's semantics otherwise:
one index per dimension is an outer-product (orthogonal) selection, a single index on an N-d
array is a linear column-major index, repeated indices accumulate.
"""
import math
import numpy as np

# ----------------------------------------------------------------------------------------------
# Types (src/core.jl:3-41)
# ----------------------------------------------------------------------------------------------


class Value:
    __array_priority__ = 1000  # numpy scalars/arrays defer to our operators

    # operator sugar: Julia's dotted (broadcasting) forms; `@` is Julia's matrix `*`
    def __add__(self, o): return add(self, o)
    def __radd__(self, o): return add(o, self)
    def __sub__(self, o): return sub(self, o)
    def __rsub__(self, o): return sub(o, self)
    def __mul__(self, o): return mul(self, o)
    def __rmul__(self, o): return mul(o, self)
    def __truediv__(self, o): return div(self, o)
    def __rtruediv__(self, o): return div(o, self)
    def __pow__(self, o): return pow_(self, o)
    def __matmul__(self, o): return matmul(self, o)
    def __rmatmul__(self, o): return matmul(o, self)
    def __neg__(self): return neg(self)
    def __getitem__(self, i): return getindex(self, *(i if isinstance(i, tuple) else (i,)))

    @property
    def T(self): return adjoint(self)


class Tracked(Value):
    pass


class Param(Tracked):
    def __init__(self, value, opt=None):
        if isinstance(value, Value):
            raise TypeError("Param cannot take %s as arg." % type(value).__name__)
        self.value = value
        self.opt = opt


class Result(Tracked):
    def __init__(self, value, func, args, kwargs):
        self.value = value
        self.func = func
        self.args = args
        self.kwargs = kwargs


class Node:
    def __init__(self, v):
        self.Value = v
        self.parents = []
        self.children = []
        self.outgrad = None


class Tape:
    def __init__(self):
        self.dict = {}   # id(Tracked) -> Node   (IdDict)
        self.list = []


_tapes = []


def recording():
    return len(_tapes) > 0


def value(x):
    if isinstance(x, Value):
        return x.value
    if isinstance(x, Tape):
        return None if not x.list else x.list[-1].Value.value
    return x


def _isnum(x):
    return isinstance(x, (int, float, np.floating, np.integer)) and not isinstance(x, bool)


# ----------------------------------------------------------------------------------------------
# Recording (src/core.jl:64-129).  Python has no dot-broadcast syntax, so the Bcasted box is not
# needed: every elementwise primitive here *is* the dotted form.
# ----------------------------------------------------------------------------------------------


def forw(f, *args, **kwargs):
    novalue = tuple(value(a) for a in args)
    v = f.fwd(*novalue, **kwargs)
    if recording() and any(isinstance(a, Tracked) for a in args):
        v = _record(v, f, args, kwargs)
    return v


def _record_node(tape, t):
    n = tape.dict.get(id(t))
    if n is None:
        n = Node(t)
        tape.dict[id(t)] = n
        tape.list.append(n)
    return n


def _record(v, f, args, kwargs):
    if isinstance(v, Tracked):   # Issue #106
        return v
    result = Result(v, f, args, kwargs)
    for tape in _tapes:
        node = Node(result)
        node.parents = [None] * len(args)
        for i, a in enumerate(args):
            if isinstance(a, Tracked):
                p = _record_node(tape, a)
                node.parents[i] = p
                p.children.append(node)
        tape.dict[id(result)] = node
        tape.list.append(node)
    return result


gcnode = [lambda n, t: None]


def set_gc_function(f):
    gcnode[0] = f


def differentiate(f, *x, **o):
    """src/core.jl:134-171."""
    tape = Tape()
    _tapes.append(tape)
    try:
        result = f(*x, **o)
        if isinstance(result, Param):
            result = identity(result)
    except BaseException:
        _tapes.pop()
        raise
    tape1 = _tapes.pop()
    assert tape1 is tape, "Tape stack error"
    if not isinstance(result, Result):
        return result
    assert _isnum(result.value) or (isinstance(result.value, np.ndarray) and result.value.ndim == 0), \
        "Only scalar valued functions supported."
    resultnode = _findresult(tape, result)
    resultnode.outgrad = type(result.value)(1) if not isinstance(result.value, np.ndarray) \
        else np.ones((), result.value.dtype)[()]
    for ti in range(len(tape.list) - 1, -1, -1):
        n = tape.list[ti]
        if n.outgrad is None:
            continue
        r = n.Value
        for i in range(len(n.parents)):
            p = n.parents[i]
            if p is None:
                continue
            if isinstance(n.outgrad, Sparse):
                n.outgrad = full(n.outgrad)
            g = back(r.func, i, n.outgrad, r, *r.args, **r.kwargs)
            p.outgrad = addto(p.outgrad, g)
        if not _tapes and isinstance(r, Result) and n is not resultnode:
            gcnode[0](n, tape)
    return tape


def _findresult(tape, result):
    if tape.list[-1].Value is result:
        return tape.list[-1]
    for i in range(len(tape.list) - 2, -1, -1):
        if tape.list[i].Value is result:
            del tape.list[i + 1:]
            break
    assert tape.list[-1].Value is result, "result not found on tape"
    return tape.list[-1]


def diff(thunk):
    """`@diff expr` -> differentiate(()->expr)  (src/core.jl:205)."""
    return differentiate(thunk)


def grad(a, b=0, loss=False):
    """grad(tape, x) (src/core.jl:215-216) or grad(fun, argnum) (src/core.jl:219-230; argnum 0-based)."""
    if isinstance(a, Tape):
        if not isinstance(b, Tracked):
            return None
        n = a.dict.get(id(b))
        return None if n is None else n.outgrad
    if not callable(a):
        return None
    fun, argnum = a, b

    def gradfun(*args, **kwargs):
        args = list(args)
        if not isinstance(args[argnum], Tracked):
            args[argnum] = Param(args[argnum])
        result = differentiate(fun, *args, **kwargs)
        xgrad = full(grad(result, args[argnum])) if isinstance(result, Tape) else None
        return (xgrad, value(result)) if loss else xgrad
    return gradfun


def gradloss(f, argnum=0):
    return grad(f, argnum, True)


def params(t):
    """params(::Tape) (src/params.jl:6) and a shallow recursive walk for containers (src/params.jl:3)."""
    if isinstance(t, Tape):
        return [n.Value for n in t.list if isinstance(n.Value, Param)]
    out, seen = [], set()

    def walk(x):
        if isinstance(x, Param):
            if id(x) not in seen:
                seen.add(id(x))
                out.append(x)
        elif isinstance(x, (list, tuple)):
            for y in x:
                walk(y)
        elif isinstance(x, dict):
            for y in x.values():
                walk(y)
        elif hasattr(x, "__dict__") and not isinstance(x, (type, np.ndarray)):
            for y in vars(x).values():
                walk(y)
    walk(t)
    return out


# ----------------------------------------------------------------------------------------------
# Primitive definition (src/macros.jl:88-194 surface)
# ----------------------------------------------------------------------------------------------


class Primitive:
    def __init__(self, name, fwd, backs):
        self.name, self.fwd, self.backs = name, fwd, backs

    def __call__(self, *args, **kwargs):
        if any(isinstance(a, Value) for a in args):
            return forw(self, *args, **kwargs)
        return self.fwd(*args, **kwargs)

    def __repr__(self):
        return self.name


def primitive(name, fwd, *backs):
    return Primitive(name, fwd, backs)


def zerograd(name, fwd):
    """@zerograd: unbox, do not record (src/macros.jl:158-170)."""
    def z(*args, **kwargs):
        return fwd(*(value(a) for a in args), **kwargs)
    z.__name__ = name
    return z


def back(f, i, dy, y, *args, **kwargs):
    """back(f, Arg{i+1}, dy, y, args...)  (src/core.jl:165,174)."""
    rule = f.backs[i] if i < len(f.backs) else None
    if rule is None:
        if getattr(f, "only_first", False) or i < len(f.backs):
            return None
        raise TypeError("MethodError: no back method for %s Arg{%d}" % (f.name, i + 1))
    return rule(dy, y, *args, **kwargs)


# ----------------------------------------------------------------------------------------------
# Julia broadcasting on numpy arrays (leading-dimension alignment)
# ----------------------------------------------------------------------------------------------


def size(x, d=None):
    v = value(x)
    s = () if _isnum(v) else tuple(v.shape)
    if d is None:
        return s
    return s[d] if d < len(s) else 1


def ndims(x):
    return len(size(x))


def length(x):
    return int(np.prod(size(x), dtype=np.int64))


def _pad(a, nd):
    if _isnum(a) or a.ndim == nd:
        return a
    return a.reshape(a.shape + (1,) * (nd - a.ndim), order="F")


def _bcast(fn, *xs):
    nd = max((x.ndim for x in xs if isinstance(x, np.ndarray)), default=0)
    dts = [x.dtype for x in xs if isinstance(x, np.ndarray)]
    if dts:   # Julia: Float32 array with an Int/weak scalar stays Float32
        dt = np.result_type(*dts)
        xs = [x if isinstance(x, np.ndarray) else dt.type(x) for x in xs]
    r = fn(*[_pad(x, nd) for x in xs])
    if isinstance(r, np.ndarray) and r.ndim > 0:
        r = np.asfortranarray(r)
    return r


def _like(dy, x):
    """dtype helper for rules producing masks."""
    return x.dtype.type if isinstance(x, np.ndarray) else type(x)


# ----------------------------------------------------------------------------------------------
# unbroadcast (src/unbroadcast.jl:9-26)
# ----------------------------------------------------------------------------------------------


def unbroadcast(x, dx):
    if size(x) == size(dx):
        return dx
    if _isnum(value(x)):
        return sum_(dx)
    if _isnum(value(dx)):
        xv = value(x)
        return fill_like(xv, dx)
    d = []
    for i in range(ndims(dx)):
        if size(x, i) == size(dx, i) and size(dx, i) > 1:
            continue
        if size(x, i) != 1:
            raise ValueError("DimensionMismatch")
        d.append(i)
    return reshape(sum_(dx, dims=tuple(d)), size(x))


def _fill_like_fwd(xv, v):
    return np.full(xv.shape, v, dtype=xv.dtype, order="F")


fill_like = primitive("fill_like", _fill_like_fwd, None, lambda dy, y, x, v: sum_(dy))

# ----------------------------------------------------------------------------------------------
# Sparse (src/sparse.jl:15-83) and addto! (src/addto.jl)
# ----------------------------------------------------------------------------------------------


class Sparse:
    def __init__(self, container, values, indices):
        self.container, self.values, self.indices = container, list(values), list(indices)

    @property
    def shape(self):
        return self.container.shape

    def __add__(self, o):
        if isinstance(o, Sparse):    # src/sparse.jl:79-82
            assert matches(self.container, o.container)
            return Sparse(self.container, self.values + o.values, self.indices + o.indices)
        return _addto_array(np.array(o, order="F", copy=True), self)   # :72-73

    __radd__ = __add__

    def __neg__(self):
        return self * -1

    def __mul__(self, n):           # :65-66
        return Sparse(self.container, [v * n for v in self.values], self.indices)

    __rmul__ = __mul__

    def __truediv__(self, n):
        return Sparse(self.container, [v / n for v in self.values], self.indices)

    def __rsub__(self, a):          # a - s  (:74)
        return _addto_array(np.array(a, order="F", copy=True), -self)

    def __sub__(self, a):           # s - a  (:75)
        return _addto_array(-np.asarray(a), self)


def matches(a, b):
    if _isnum(a) and _isnum(b):
        return True
    if isinstance(a, np.ndarray) and isinstance(b, np.ndarray):
        return a.shape == b.shape
    if isinstance(a, tuple) and isinstance(b, tuple):
        return len(a) == len(b)
    if isinstance(a, dict) and isinstance(b, dict):
        return True
    return False


def zeroslike(a):
    return np.zeros(a.shape, a.dtype, order="F")


def full(b):
    if isinstance(b, Sparse):
        return _addto_array(zeroslike(b.container), b)
    return b


def lmul(a, x):
    return Sparse(x.container, [a * v for v in x.values], x.indices)


def sparse_norm(x):
    """src/sparse.jl:53-54 (ignores repeated indices)."""
    return math.sqrt(sum(float(np.sum(np.square(np.asarray(v, dtype=np.float64)))) for v in x.values))


def axpy(a, x, y):
    """axpy!(a, x, y): y += a*x in place; Sparse x per src/sparse.jl:50; Param y via value."""
    yv = value(y)
    if isinstance(x, Sparse):
        _addto_array(yv, x * a, inplace=True)
    else:
        yv += yv.dtype.type(a) * x
    return y


def to_indices(shape, idx):
    """Normalise a Julia-style index tuple (0-based) to per-dimension position arrays.

    Returns (pos, outshape): pos[d] is an int64 array of positions along the d-th *indexed*
    dimension, outshape the shape of getindex's result (scalar indices drop their dimension,
    an index array contributes its own shape).  A single index on an N-d array is linear
    (column-major).  Mirrors Base.to_indices as used at src/addto.jl:95.
    """
    if not isinstance(idx, tuple):
        idx = (idx,)
    # a full-array Bool mask or a single index => linear indexing over the flattened array
    if len(idx) == 1 and len(shape) != 1:
        dimsizes = [int(np.prod(shape, dtype=np.int64))]
    else:
        # CartesianIndex (a tuple inside the tuple) splices
        flat = []
        for i in idx:
            if isinstance(i, tuple):
                flat.extend(i)
            else:
                flat.append(i)
        idx = tuple(flat)
        dimsizes = list(shape) + [1] * (len(idx) - len(shape))
        if len(idx) < len(shape):
            # trailing dims must be merged into the last index (Julia partial linear indexing)
            dimsizes = list(shape[:len(idx) - 1]) + [int(np.prod(shape[len(idx) - 1:], dtype=np.int64))]
    pos, outshape = [], []
    for i, n in zip(idx, dimsizes):
        if isinstance(i, slice):
            p = np.arange(n, dtype=np.int64)[i]
            outshape.append(len(p))
        elif isinstance(i, range):
            p = np.asarray(list(i), dtype=np.int64)
            outshape.append(len(p))
        elif isinstance(i, (int, np.integer)) and not isinstance(i, (bool, np.bool_)):
            p = np.asarray([int(i)], dtype=np.int64)
        else:
            a = np.asarray(i)
            if a.dtype == np.bool_:
                p = np.flatnonzero(a.reshape(-1, order="F")).astype(np.int64)
                outshape.append(len(p))
            else:
                p = a.astype(np.int64).reshape(-1, order="F")
                outshape.extend(a.shape)
        if p.size and (p.min() < 0 or p.max() >= n):
            raise IndexError("BoundsError")
        pos.append(p)
    return pos, tuple(outshape), dimsizes


def flat_dest(shape, idx):
    """Column-major flat (0-based) destinations of x[idx...] in iteration order (first index
    fastest) -- the bit-exact index bookkeeping of src/addto.jl:107-129."""
    pos, outshape, dimsizes = to_indices(shape, idx)
    return _dest_colmajor(pos, dimsizes), outshape


def _dest_colmajor(pos, dimsizes):
    dest = np.zeros(1, dtype=np.int64)
    stride = 1
    for p, n in zip(pos, dimsizes):
        # new[j*len(dest)+i] = dest[i] + p[j]*stride  (earlier dims fastest)
        dest = (dest[None, :] + (p * stride)[:, None]).reshape(-1)
        stride *= n
    return dest


def addtoindex(A, x, idx):
    """addtoindex!(A, x, I...) with repeated indices summed sequentially, column-major iteration
    (src/addto.jl:107-129)."""
    dest, outshape = flat_dest(A.shape, idx)
    if _isnum(x) or (isinstance(x, np.ndarray) and x.ndim == 0):
        xv = np.full(dest.shape, x, dtype=A.dtype)
    else:
        xv = np.asarray(x, dtype=A.dtype).reshape(-1, order="F")
        if xv.size != dest.size:
            raise ValueError("DimensionMismatch")
    flat = A.reshape(-1, order="F")
    if not np.shares_memory(flat, A):
        raise RuntimeError("container must be column-major contiguous")
    np.add.at(flat, dest, xv)   # unbuffered, sequential in index order == the reference's loop
    return A


def _addto_array(a, b, inplace=True):
    assert matches(a, b.container)
    if recording():
        a = np.array(a, order="F", copy=True)
    for idx, val in zip(b.indices, b.values):
        addtoindex(a, val, idx)
    return a


def _addto_fwd(a, b):
    """Unboxed addto! (src/addto.jl:20-31,41-52,91-98)."""
    if a is None:
        return b
    if b is None:
        return a
    if isinstance(a, Sparse) and isinstance(b, Sparse):
        assert matches(a.container, b.container)
        a.values.extend(b.values)
        a.indices.extend(b.indices)
        return a
    if isinstance(a, Sparse):
        a, b = b, a
    if isinstance(b, Sparse):
        if _isnum(a):
            for idx, val in zip(b.indices, b.values):
                a = a + val
            return a
        return _addto_array(a, b)
    if isinstance(a, np.ndarray) or isinstance(b, np.ndarray):
        if isinstance(a, np.ndarray) and isinstance(b, np.ndarray) and a.shape != b.shape:
            raise ValueError("DimensionMismatch")
        return a + b    # out of place (src/addto.jl:23)
    return a + b


def _addto_call(a, b):
    if isinstance(a, Value) or isinstance(b, Value):
        return forw(addto, a, b)
    return _addto_fwd(a, b)


class _AddTo(Primitive):
    def __call__(self, a, b):
        return _addto_call(a, b)


addto = _AddTo("addto!", _addto_fwd, (lambda dy, y, a, b: dy, lambda dy, y, a, b: dy))

# ----------------------------------------------------------------------------------------------
# Elementwise rules (src/base.jl:20-49,73-74; src/math.jl:9-74)
# ----------------------------------------------------------------------------------------------


def _np(fn):
    return lambda *xs: _bcast(fn, *xs)


def _eqmask(y, x):
    """`y .== x` as a zero-gradient mask in the operand dtype."""
    yv, xv = value(y), value(x)
    r = _bcast(lambda a, b: (a == b), yv, xv)
    dt = yv.dtype if isinstance(yv, np.ndarray) else (xv.dtype if isinstance(xv, np.ndarray) else np.float64)
    return r.astype(dt) if isinstance(r, np.ndarray) else dt.type(r) if hasattr(dt, "type") else float(r)


sign = zerograd("sign", _np(np.sign))
one = zerograd("one", lambda x: np.ones_like(x) if isinstance(x, np.ndarray) else type(x)(1))

add = primitive("+", _np(lambda a, b: a + b),
                lambda dy, y, x1, x2: unbroadcast(x1, dy),
                lambda dy, y, x1, x2: unbroadcast(x2, dy))
sub = primitive("-", _np(lambda a, b: a - b),
                lambda dy, y, x1, x2: unbroadcast(x1, dy),
                lambda dy, y, x1, x2: unbroadcast(x2, neg(dy)))
neg = primitive("neg", _np(lambda a: -a), lambda dy, y, x: neg(dy))
mul = primitive(".*", _np(lambda a, b: a * b),
                lambda dy, y, x1, x2: unbroadcast(x1, mul(dy, x2)),
                lambda dy, y, x1, x2: unbroadcast(x2, mul(x1, dy)))
div = primitive("./", _np(lambda a, b: a / b),
                lambda dy, y, x1, x2: unbroadcast(x1, div(dy, x2)),
                lambda dy, y, x1, x2: unbroadcast(x2, div(mul(neg(dy), x1), abs2(x2))))


def dxndx(x1, x2, dy):
    """src/base.jl:49."""
    if _isnum(value(x2)):
        p = value(x2)
        if p == 0:
            return mul(dy, 0)
        if p == 1:
            return dy
        if p == 2:
            return mul(mul(2, x1), dy)
    return mul(mul(dy, x2), pow_(x1, sub(x2, 1)))


pow_ = primitive(".^", _np(lambda a, b: a ** b),
                 lambda dy, y, x1, x2: unbroadcast(x1, dxndx(x1, x2, dy)),
                 lambda dy, y, x1, x2: unbroadcast(x2, mul(mul(dy, y), log(x1))))

abs_ = primitive("abs", _np(np.abs), lambda dy, y, x: mul(dy, sign(x)))
abs2 = primitive("abs2", _np(lambda a: a * a), lambda dy, y, x: mul(dy, mul(2, x)))
exp = primitive("exp", _np(np.exp), lambda dy, y, x: mul(dy, y))
log = primitive("log", _np(np.log), lambda dy, y, x: mul(dy, div(1, x)))
tanh = primitive("tanh", _np(np.tanh), lambda dy, y, x: mul(dy, sub(1, abs2(y))))
sin = primitive("sin", _np(np.sin), lambda dy, y, x: mul(dy, cos(x)))
cos = primitive("cos", _np(np.cos), lambda dy, y, x: mul(dy, neg(sin(x))))
sqrt = primitive("sqrt", _np(np.sqrt), lambda dy, y, x: mul(dy, div(1, mul(2, y))))
# user-defined primitive of examples/charlm.jl:46-47
sigm = primitive("sigm", _np(lambda a: 1 / (1 + np.exp(-a))), lambda dy, y, x: mul(mul(dy, y), sub(1, y)))
max_ = primitive("max", _np(np.maximum),
                 lambda dy, y, x1, x2: unbroadcast(x1, mul(dy, _eqmask(y, x1))),
                 lambda dy, y, x1, x2: unbroadcast(x2, mul(dy, _eqmask(y, x2))))
min_ = primitive("min", _np(np.minimum),
                 lambda dy, y, x1, x2: unbroadcast(x1, mul(dy, _eqmask(y, x1))),
                 lambda dy, y, x1, x2: unbroadcast(x2, mul(dy, _eqmask(y, x2))))
identity = primitive("identity", lambda x: x, lambda dy, y, x: dy)
copy = primitive("copy", lambda x: np.array(x, order="F", copy=True) if isinstance(x, np.ndarray) else x,
                 lambda dy, y, x: dy)


# ---- the rest of src/math.jl (degree / reciprocal / pi-scaled trigonometry, sinc / cosc, clamp, ldexp, log(b, x)):
# forward by numpy's own functions, gradient expressions as the reference writes them (line numbers of src/math.jl)
_D, _PI = 180.0 / np.pi, np.pi
tan = primitive("tan", _np(np.tan), lambda dy, y, x: mul(dy, add(1, abs2(y))))                                                  # :72
acosd = primitive("acosd", _np(lambda x: np.degrees(np.arccos(x))), lambda dy, y, x: mul(dy, div(-_D, sqrt(sub(1, abs2(x))))))   # :10
acot = primitive("acot", _np(lambda x: np.arctan(1 / x)), lambda dy, y, x: mul(dy, div(-1, add(1, abs2(x)))))                   # :12
acotd = primitive("acotd", _np(lambda x: np.degrees(np.arctan(1 / x))), lambda dy, y, x: mul(dy, div(-_D, add(1, abs2(x)))))    # :13
acoth = primitive("acoth", _np(lambda x: np.arctanh(1 / x)), lambda dy, y, x: mul(dy, div(1, sub(1, abs2(x)))))                 # :14
acsc = primitive("acsc", _np(lambda x: np.arcsin(1 / x)),                                                                       # :15
                 lambda dy, y, x: mul(dy, div(-1, sqrt(mul(mul(abs2(x), sub(x, 1)), add(x, 1))))))
acscd = primitive("acscd", _np(lambda x: np.degrees(np.arcsin(1 / x))),                                                         # :16
                  lambda dy, y, x: mul(dy, div(-_D, sqrt(mul(mul(abs2(x), sub(x, 1)), add(x, 1))))))
acsch = primitive("acsch", _np(lambda x: np.arcsinh(1 / x)), lambda dy, y, x: mul(dy, div(-1, sqrt(add(abs2(abs2(x)), abs2(x))))))   # :17
asec = primitive("asec", _np(lambda x: np.arccos(1 / x)), lambda dy, y, x: mul(dy, div(1, sqrt(sub(abs2(abs2(x)), abs2(x))))))       # :18
asecd = primitive("asecd", _np(lambda x: np.degrees(np.arccos(1 / x))),                                                         # :19
                  lambda dy, y, x: mul(dy, div(_D, sqrt(sub(abs2(abs2(x)), abs2(x))))))
asech = primitive("asech", _np(lambda x: np.arccosh(1 / x)), lambda dy, y, x: mul(dy, div(-1, sqrt(sub(abs2(x), abs2(abs2(x)))))))   # :20
asind = primitive("asind", _np(lambda x: np.degrees(np.arcsin(x))), lambda dy, y, x: mul(dy, div(_D, sqrt(sub(1, abs2(x))))))   # :22
atand = primitive("atand", _np(lambda x: np.degrees(np.arctan(x))), lambda dy, y, x: mul(dy, div(_D, add(1, abs2(x)))))         # :25
cosd = primitive("cosd", _np(lambda x: np.cos(np.radians(x))), lambda dy, y, x: mul(dy, div(mul(neg(sind(x)), _PI), 180)))      # :32
cospi = primitive("cospi", _np(lambda x: np.cos(np.pi * x)), lambda dy, y, x: mul(dy, mul(neg(sinpi(x)), _PI)))                 # :34
cot = primitive("cot", _np(lambda x: 1 / np.tan(x)), lambda dy, y, x: mul(dy, neg(abs2(csc(x)))))                               # :35
cotd = primitive("cotd", _np(lambda x: 1 / np.tan(np.radians(x))), lambda dy, y, x: mul(dy, div(mul(neg(abs2(cscd(x))), _PI), 180)))  # :36
coth = primitive("coth", _np(lambda x: 1 / np.tanh(x)), lambda dy, y, x: mul(dy, neg(abs2(csch(x)))))                           # :37
csc = primitive("csc", _np(lambda x: 1 / np.sin(x)), lambda dy, y, x: mul(dy, mul(mul(-1, y), cot(x))))                         # :38
cscd = primitive("cscd", _np(lambda x: 1 / np.sin(np.radians(x))),                                                              # :39
                 lambda dy, y, x: mul(dy, div(mul(mul(mul(-1, y), cotd(x)), _PI), 180)))
csch = primitive("csch", _np(lambda x: 1 / np.sinh(x)), lambda dy, y, x: mul(dy, mul(mul(-1, y), coth(x))))                     # :40
deg2rad = primitive("deg2rad", _np(np.radians), lambda dy, y, x: mul(dy, _PI / 180))                                            # :41
rad2deg = primitive("rad2deg", _np(np.degrees), lambda dy, y, x: mul(dy, _D))                                                   # :58
sec = primitive("sec", _np(lambda x: 1 / np.cos(x)), lambda dy, y, x: mul(dy, mul(y, tan(x))))                                  # :60
secd = primitive("secd", _np(lambda x: 1 / np.cos(np.radians(x))), lambda dy, y, x: mul(dy, div(mul(mul(y, tand(x)), _PI), 180)))    # :61
sech = primitive("sech", _np(lambda x: 1 / np.cosh(x)), lambda dy, y, x: mul(dy, mul(mul(-1, y), tanh(x))))                     # :62
sind = primitive("sind", _np(lambda x: np.sin(np.radians(x))), lambda dy, y, x: mul(dy, div(mul(cosd(x), _PI), 180)))           # :67
sinpi = primitive("sinpi", _np(lambda x: np.sin(np.pi * x)), lambda dy, y, x: mul(dy, mul(cospi(x), _PI)))                      # :69
tand = primitive("tand", _np(lambda x: np.tan(np.radians(x))), lambda dy, y, x: mul(dy, div(mul(add(1, abs2(y)), _PI), 180)))   # :73


def _cosc_np(x):
    x = np.asarray(x)
    safe = np.where(x == 0, 1.0, x)
    return np.where(x == 0, 0.0, np.cos(np.pi * safe) / safe - np.sin(np.pi * safe) / (np.pi * safe * safe)).astype(x.dtype)


sinc = primitive("sinc", _np(np.sinc), lambda dy, y, x: mul(dy, cosc(x)))                                                       # :65
cosc = primitive("cosc", _np(_cosc_np), lambda dy, y, x: mul(dy, sub(div(mul(-2, y), x), mul(sinc(x), _PI ** 2))))             # :31


def _lemask(a, b):
    av, bv = value(a), value(b)
    r = _bcast(lambda p, q: (p <= q), av, bv)
    dt = av.dtype if isinstance(av, np.ndarray) else (bv.dtype if isinstance(bv, np.ndarray) else np.float64)
    return r.astype(dt) if isinstance(r, np.ndarray) else float(r)


clamp = primitive("clamp", lambda x, lo, hi: np.clip(x, lo, hi),                                                                # :28
                  lambda dy, y, x, lo, hi: unbroadcast(x, mul(dy, mul(_lemask(lo, x), _lemask(x, hi)))))
ldexp = primitive("ldexp", lambda x, n: np.ldexp(x, n), lambda dy, y, x, n: mul(dy, 2.0 ** n))                                  # :47
logb = primitive("log", _np(lambda b, x: np.log(x) / np.log(b)),                                                                # :48 log(b, x)
                 lambda dy, y, x1, x2: unbroadcast(x1, mul(dy, div(neg(log(x2)), mul(x1, abs2(log(x1)))))),
                 lambda dy, y, x1, x2: unbroadcast(x2, mul(dy, div(1, mul(x2, log(x1))))))
exponent = lambda x: (np.frexp(value(x))[1] - 1).astype(np.asarray(value(x)).dtype)                                           # :44 @zerograd
significand = primitive("significand", lambda x: (np.frexp(x)[0] * 2).astype(np.asarray(x).dtype),                              # :64
                        lambda dy, y, x: mul(dy, np.exp2(-exponent(x))))
mod2pi = primitive("mod2pi", lambda x: np.remainder(x, 2 * np.pi), lambda dy, y, x: dy)                                         # :56
_REM = {"nearest": lambda x: x - 2 * np.pi * np.rint(x / (2 * np.pi)), "tozero": lambda x: np.fmod(x, 2 * np.pi),
        "down": lambda x: np.remainder(x, 2 * np.pi), "up": lambda x: -np.remainder(-x, 2 * np.pi)}
rem2pi = primitive("rem2pi", lambda x, r: _REM[r](x), lambda dy, y, x, r: dy)                                                   # :59
# @zerograd rounding functions (src/base.jl:79,100,145,151)
floor = lambda x: np.floor(value(x))
ceil = lambda x: np.ceil(value(x))
round_ = lambda x: np.rint(value(x))
trunc = lambda x: np.trunc(value(x))


# special functions (src/specialfunctions.jl:20-21,29-30,34-42,63-67) through scipy.special, the counterpart of
# SpecialFunctions.jl; gradient expressions exactly as the reference writes them
def _sp(name, *pre):
    def f(a):
        import scipy.special as sp
        return getattr(sp, name)(*pre, a)
    return f


_2SPI, _SPI2 = 2.0 / np.sqrt(np.pi), np.sqrt(np.pi) / 2.0
erf = primitive("erf", _np(_sp("erf")), lambda dy, y, x: mul(dy, mul(exp(neg(abs2(x))), _2SPI)))
erfc = primitive("erfc", _np(_sp("erfc")), lambda dy, y, x: mul(dy, mul(neg(exp(neg(abs2(x)))), _2SPI)))
erfcx = primitive("erfcx", _np(_sp("erfcx")), lambda dy, y, x: mul(dy, sub(mul(mul(2, y), x), _2SPI)))
erfinv = primitive("erfinv", _np(_sp("erfinv")), lambda dy, y, x: mul(dy, mul(exp(abs2(y)), _SPI2)))
erfcinv = primitive("erfcinv", _np(_sp("erfcinv")), lambda dy, y, x: mul(dy, mul(neg(exp(abs2(y))), _SPI2)))
polygamma2 = lambda x: _np(_sp("polygamma", 2))(value(x))                     # forward only
trigamma = primitive("trigamma", _np(_sp("polygamma", 1)), lambda dy, y, x: mul(dy, polygamma2(x)))
digamma = primitive("digamma", _np(_sp("digamma")), lambda dy, y, x: mul(dy, trigamma(x)))
gamma = primitive("gamma", _np(_sp("gamma")), lambda dy, y, x: mul(dy, mul(y, digamma(x))))
lgamma = primitive("lgamma", _np(_sp("gammaln")), lambda dy, y, x: mul(dy, digamma(x)))
loggamma = lgamma
besselj1 = primitive("besselj1", _np(_sp("j1")),
                     lambda dy, y, x: mul(dy, div(sub(besselj0(x), _np(_sp("jn", 2))(value(x))), 2)))
besselj0 = primitive("besselj0", _np(_sp("j0")), lambda dy, y, x: mul(dy, neg(besselj1(x))))
bessely1 = primitive("bessely1", _np(_sp("y1")),
                     lambda dy, y, x: mul(dy, div(sub(bessely0(x), _np(_sp("yn", 2))(value(x))), 2)))
bessely0 = primitive("bessely0", _np(_sp("y0")), lambda dy, y, x: mul(dy, neg(bessely1(x))))


# functions with a (non-differentiated) order: src/specialfunctions.jl:20,25,43,59
def _sp2(name):
    def f(a, x):
        import scipy.special as sp
        return np.asarray(getattr(sp, name)(a, x), dtype=np.asarray(x).dtype) if isinstance(x, np.ndarray) else float(getattr(sp, name)(a, x))
    return f


besselj = primitive("besselj", _sp2("jv"), None,
                    lambda dy, y, a, x: mul(dy, div(sub(besselj(a - 1, x), besselj(a + 1, x)), 2)))
bessely = primitive("bessely", _sp2("yv"), None,
                    lambda dy, y, a, x: mul(dy, div(sub(bessely(a - 1, x), bessely(a + 1, x)), 2)))
besseli = primitive("besseli", _sp2("iv"), None,                                       # src/specialfunctions.jl:13
                    lambda dy, y, a, x: mul(dy, div(add(besseli(a - 1, x), besseli(a + 1, x)), 2)))
besselk = primitive("besselk", _sp2("kv"), None,                                       # :24
                    lambda dy, y, a, x: mul(dy, div(add(besselk(a - 1, x), besselk(a + 1, x)), -2)))
polygamma = primitive("polygamma", _sp2("polygamma"), None,
                      lambda dy, y, a, x: unbroadcast(x, mul(dy, polygamma(a + 1, x))))


def _invdigamma_np(x):
    import scipy.special as sp
    xv = np.asarray(x, dtype=np.float64)
    t = np.where(xv >= -2.22, np.exp(xv) + 0.5, -1.0 / (xv + 0.5772156649015329))
    for _ in range(12):
        t = t - (sp.digamma(t) - xv) / sp.polygamma(1, t)
    return t.astype(np.asarray(x).dtype) if isinstance(x, np.ndarray) else float(t)


invdigamma = primitive("invdigamma", _invdigamma_np, lambda dy, y, x: mul(dy, div(1, trigamma(y))))


# src/specialfunctions.jl:5-8 (Airy functions), :33 (dawson), :39 (erfi); gradient expressions as the reference writes them
def _airy(k):
    def f(a):
        import scipy.special as sp
        r = sp.airy(np.asarray(a, dtype=np.float64))[k]
        return r.astype(a.dtype) if isinstance(a, np.ndarray) else float(r)
    return f


airyai = primitive("airyai", _airy(0), lambda dy, y, x: mul(dy, airyaiprime(x)))
airyaiprime = primitive("airyaiprime", _airy(1), lambda dy, y, x: mul(dy, mul(x, airyai(x))))
airybi = primitive("airybi", _airy(2), lambda dy, y, x: mul(dy, airybiprime(x)))
airybiprime = primitive("airybiprime", _airy(3), lambda dy, y, x: mul(dy, mul(x, airybi(x))))
dawson = primitive("dawson", _np(_sp("dawsn")), lambda dy, y, x: mul(dy, add(mul(mul(-2, y), x), 1)))
erfi = primitive("erfi", _np(_sp("erfi")), lambda dy, y, x: mul(dy, mul(exp(abs2(x)), _2SPI)))


# src/linearalgebra.jl:19-20 (diag), :86 (tr), :90,92 (tril / triu), :176-186 (_kron), rules as the reference writes them
def _diagm_np(v, k=0):
    return np.asfortranarray(np.diag(np.asarray(v).reshape(-1), k))


diag = primitive("diag", lambda x, k=0: np.diag(x, k).copy(), lambda dy, y, x, k=0: _diagm_np(value(dy), k))
diagm = primitive("diagm", _diagm_np, lambda dy, y, v, k=0: reshape(diag(dy, k), np.shape(value(v))))
tr = primitive("tr", lambda x: np.trace(x), lambda dy, y, x: mul(dy, np.asfortranarray(np.eye(*value(x).shape, dtype=value(x).dtype))))
tril = primitive("tril", lambda x: np.asfortranarray(np.tril(x)),
                 lambda dy, y, x: mul(dy, np.asfortranarray(np.tril(np.ones_like(value(x))))))
triu = primitive("triu", lambda x: np.asfortranarray(np.triu(x)),
                 lambda dy, y, x: mul(dy, np.asfortranarray(np.triu(np.ones_like(value(x))))))


# src/linearalgebra.jl:18 (det), :28 (inv), :58 (logdet) through numpy.linalg (LAPACK, like the reference)
inv = primitive("inv", lambda x: np.asfortranarray(np.linalg.inv(x)),
                lambda dy, y, x: neg(matmul(matmul(adjoint(y), dy), adjoint(y))))
det = primitive("det", lambda x: np.linalg.det(x), lambda dy, y, x: mul(mul(dy, y), adjoint(inv(x))))
logdet = primitive("logdet", lambda x: np.linalg.slogdet(x)[1], lambda dy, y, x: mul(dy, adjoint(inv(x))))


def kron(a, b):
    (m1, n1), (m2, n2) = np.shape(value(a)), np.shape(value(b))
    return reshape(mul(reshape(a, (1, m1, 1, n1)), reshape(b, (m2, 1, n2, 1))), (m1 * m2, n1 * n2))


def _normback(x, p, dy, y):
    """src/linearalgebra.jl:222-238, branch for branch (raw arrays)."""
    if x.size == 0:
        return np.empty_like(x)
    if p == 2:
        return (dy / y) * x
    if p == 1:
        return dy * np.sign(x)
    if p == np.inf or p == -np.inf:
        return dy * np.sign(x) * (np.abs(x) == y)
    if p == 0:
        return np.zeros_like(x)
    return dy * y ** (1 - p) * np.abs(x) ** (p - 1) * np.sign(x)


def _norm_fwd(x, p=2):
    v = np.abs(np.asarray(x, dtype=np.float64)).reshape(-1)
    if p == np.inf:
        r = v.max()
    elif p == -np.inf:
        r = v.min()
    elif p == 0:
        r = float(np.count_nonzero(v))
    elif p == 1:
        r = v.sum()
    else:
        r = (v ** p).sum() ** (1.0 / p)
    return np.asarray(x).dtype.type(r)


norm = primitive("norm", _norm_fwd, lambda dy, y, x, p=2: _normback(value(x), p, value(dy), value(y)))


def relu(x):
    """`max.(0, x)` (examples/mnist.jl:32)."""
    return max_(0, x)


# scalar math used by the reference's scalar known-answer tests (host Numbers)
# ----------------------------------------------------------------------------------------------
# Reductions (src/base.jl:245-247; src/statistics.jl:20-22)
# ----------------------------------------------------------------------------------------------


def _dims(dims):
    if dims is None:
        return None
    return (dims,) if isinstance(dims, (int, np.integer)) else tuple(dims)


def _sum_fwd(x, dims=None, f=None):
    if _isnum(x):
        return x if f is None else f(x)
    xx = x if f is None else f(x)
    if dims is None:
        return xx.sum(dtype=x.dtype)[()]
    d = tuple(i for i in _dims(dims) if i < x.ndim)
    return np.asfortranarray(xx.sum(axis=d, keepdims=True, dtype=x.dtype))


sum_ = primitive("sum", lambda x, dims=None: _sum_fwd(x, dims),
                 lambda dy, y, x, dims=None: mul(dy, one(x)))
sumabs2 = primitive("sum(abs2,·)", lambda x, dims=None: _sum_fwd(x, dims, lambda a: a * a),
                    lambda dy, y, x, dims=None: mul(dy, mul(2, x)))
sumabs = primitive("sum(abs,·)", lambda x, dims=None: _sum_fwd(x, dims, np.abs),
                   lambda dy, y, x, dims=None: mul(dy, sign(x)))


def _mean_fwd(x, dims=None, f=None):
    s = _sum_fwd(x, dims, f)
    return s / (length(x) // length(s))


mean = primitive("mean", lambda x, dims=None: _mean_fwd(x, dims),
                 lambda dy, y, x, dims=None: div(mul(dy, one(x)), length(x) // length(dy)))
meanabs2 = primitive("mean(abs2,·)", lambda x, dims=None: _mean_fwd(x, dims, lambda a: a * a),
                     lambda dy, y, x, dims=None: div(mul(dy, mul(2, x)), length(x) // length(dy)))


meanabs = primitive("mean(abs,·)", lambda x, dims=None: _mean_fwd(x, dims, np.abs),                                           # src/statistics.jl:9
                    lambda dy, y, x, dims=None: div(mul(dy, sign(x)), length(x) // length(dy)))


def _extremum(npf):
    def fwd(x, dims=None):
        if dims is None:
            return npf(x)
        return np.asfortranarray(npf(x, axis=tuple(_dims(dims)), keepdims=True))
    return fwd


_absx = lambda npf: (lambda x, dims=None: _extremum(npf)(np.abs(x), dims))
maxabs = primitive("maximum(abs,·)", _absx(np.max), lambda dy, y, x, dims=None: mul(mul(dy, _eqmask(y, np.abs(value(x)))), sign(x)))   # src/base.jl:203
minabs = primitive("minimum(abs,·)", _absx(np.min), lambda dy, y, x, dims=None: mul(mul(dy, _eqmask(y, np.abs(value(x)))), sign(x)))   # :206
rem = primitive("rem", _np(np.fmod), lambda dy, y, x1, x2: unbroadcast(x1, dy),                                                     # :144
                lambda dy, y, x1, x2: unbroadcast(x2, mul(neg(dy), np.trunc(value(x1) / value(x2)))))
ldiv = primitive("\\", _np(lambda a, b: b / a),                                                                                    # :39-42
                 lambda dy, y, x1, x2: unbroadcast(x1, div(mul(neg(dy), x2), abs2(x1))),
                 lambda dy, y, x1, x2: unbroadcast(x2, div(dy, x1)))
maximum = primitive("maximum", _extremum(np.max), lambda dy, y, x, dims=None: mul(dy, _eqmask(y, x)))
minimum = primitive("minimum", _extremum(np.min), lambda dy, y, x, dims=None: mul(dy, _eqmask(y, x)))
prod = primitive("prod", lambda x, dims=None: (np.prod(x) if dims is None else
                                                np.asfortranarray(np.prod(x, axis=tuple(_dims(dims)), keepdims=True))),
                 lambda dy, y, x, dims=None: mul(dy, div(y, x)))

# ----------------------------------------------------------------------------------------------
# Linear algebra (src/base.jl:26; src/linearalgebra.jl:16,87)
# ----------------------------------------------------------------------------------------------


def _matmul_fwd(a, b):
    if _isnum(a) or _isnum(b):
        return _bcast(lambda p, q: p * q, a, b)
    r = a @ b
    return np.asfortranarray(r) if isinstance(r, np.ndarray) and r.ndim else r


def _matmul_back1(dy, y, x1, x2):
    if _isnum(value(x1)) or _isnum(value(x2)):
        return unbroadcast(x1, mul(dy, x2))
    return matmul(dy, adjoint(x2))


def _matmul_back2(dy, y, x1, x2):
    if _isnum(value(x1)) or _isnum(value(x2)):
        return unbroadcast(x2, mul(x1, dy))
    return matmul(adjoint(x1), dy)


matmul = primitive("*", _matmul_fwd, _matmul_back1, _matmul_back2)


def _adjoint_fwd(x):
    if _isnum(x):
        return x
    if x.ndim == 1:
        return x.reshape(1, -1)
    return x.T


adjoint = primitive("adjoint", _adjoint_fwd, lambda dy, y, x: reshape(adjoint(dy), size(x)) if ndims(x) == 1 else adjoint(dy))
transpose = adjoint


def matpow(x, p):
    """`A^p` for a matrix: gradient errors by design (src/base.jl:44)."""
    if isinstance(x, Value) and ndims(x) >= 1:
        raise RuntimeError("A^B grad undefined")
    return np.linalg.matrix_power(x, p) if isinstance(x, np.ndarray) else x ** p


# ----------------------------------------------------------------------------------------------
# Shape rules (src/base.jl:217-224,250)
# ----------------------------------------------------------------------------------------------


def _reshape_fwd(x, shape):
    shape = tuple(int(s) for s in (shape if isinstance(shape, (tuple, list)) else (shape,)))
    return x.reshape(shape, order="F")


reshape = primitive("reshape", _reshape_fwd, lambda dy, y, x, shape: reshape(dy, size(x)))
vec = primitive("vec", lambda x: x.reshape(-1, order="F"), lambda dy, y, x: reshape(dy, size(x)))


def _invperm(d):
    r = [0] * len(d)
    for i, p in enumerate(d):
        r[p] = i
    return tuple(r)


permutedims = primitive("permutedims", lambda x, d: np.asfortranarray(np.transpose(x, d)),
                        lambda dy, y, x, d: permutedims(dy, _invperm(d)))

# ----------------------------------------------------------------------------------------------
# getindex / ungetindex (src/getindex.jl:32-33,61-79,91-95)
# ----------------------------------------------------------------------------------------------


def _getindex_fwd(x, *idx):
    if _isnum(x):
        return x
    dest, outshape = flat_dest(x.shape, idx)
    r = x.reshape(-1, order="F")[dest]
    if outshape == ():
        return r[0]
    return r.reshape(outshape, order="F")


def _ungetindex_fwd(x, dxi, idx):
    if _isnum(x):
        return dxi
    if recording():
        return addtoindex(zeroslike(x), dxi, idx)
    return Sparse(x, [dxi], [idx])


def ungetindex(x, dxi, idx):
    """src/getindex.jl:61-79,91-92."""
    x = value(x)
    if isinstance(dxi, Value):
        if _isnum(x):
            return dxi
        return forw(addto, zeroslike(x), forw(_ungetindex_p, x, dxi, idx))
    return _ungetindex_fwd(x, dxi, idx)


_ungetindex_p = primitive("ungetindex", lambda x, dxi, idx: addtoindex(zeroslike(x), dxi, idx),
                          None, lambda ddx, dx, x, dxi, idx: getindex(ddx, *idx), None)
getindex = primitive("getindex", _getindex_fwd, lambda dxi, xi, x, *idx: ungetindex(x, dxi, idx))
getindex.only_first = True

# ----------------------------------------------------------------------------------------------
# cat / uncat (src/cat.jl:27-31,46-92,122-123) and cat1d (src/cat.jl:150-182)
# ----------------------------------------------------------------------------------------------


def _as_nd(a, nd):
    if _isnum(a):
        a = np.asarray([a])
    return _pad(a, nd)


def _cat_fwd(*xs, dims):
    nd = max(max((0 if _isnum(x) else x.ndim) for x in xs), dims + 1)
    return np.asfortranarray(np.concatenate([_as_nd(x, nd) for x in xs], axis=dims))


def _uncat_idx(y1, argn, dims, xs):
    idx = []
    for d in range(ndims(y1)):
        if d == dims:
            pos = 0
            for j in range(argn):
                pos += size(xs[j], d)
            idx.append(slice(pos, pos + size(xs[argn], d)))
        else:
            idx.append(slice(0, size(y1, d)))
    return tuple(idx)


def _uncat_fwd(y1, argn, dims, *xs):
    x1 = y1[_uncat_idx(y1, argn, dims, xs)]
    if _isnum(value(xs[argn])):
        assert x1.size == 1, "uncat mismatch"
        return x1.reshape(-1)[0]
    return np.asfortranarray(x1).reshape(size(xs[argn]), order="F")


def _uncat1_fwd(x2, y1, argn, dims, *xs):
    y2 = np.zeros_like(y1, order="F")
    sl = _uncat_idx(y1, argn, dims, xs)
    y2[sl] = np.asarray(x2).reshape(y2[sl].shape, order="F")
    return y2


uncat = primitive("uncat", _uncat_fwd, lambda x2, x1, y1, argn, dims, *xs: uncat1(x2, y1, argn, dims, *xs))
uncat1 = primitive("uncat1", _uncat1_fwd, lambda y3, y2, x2, y1, argn, dims, *xs: uncat(y3, argn, dims, *xs))
uncat.only_first = True
uncat1.only_first = True


class _Cat(Primitive):
    def __call__(self, *args, dims):
        if any(isinstance(a, Value) for a in args):
            return forw(self, *args, dims=dims)
        return self.fwd(*args, dims=dims)


cat = _Cat("cat", _cat_fwd, ())


def _cat_back(f, i, dy, y, *xs, dims):
    return uncat(dy, i, dims, *xs)


def vcat(*xs):
    return cat(*xs, dims=0)


def hcat(*xs):
    return cat(*xs, dims=1)


def _cat1d_fwd(*xs):
    return np.concatenate([np.asarray(x).reshape(-1, order="F") for x in xs])


class _Cat1d(Primitive):
    pass


cat1d = _Cat1d("cat1d", _cat1d_fwd, ())


def _cat1d_back(f, i, y1, y, *xs):
    off2 = 0
    for j in range(i + 1):
        off2 += length(xs[j])
    off1 = off2 - length(xs[i])
    return reshape(getindex(y1, slice(off1, off2)), size(xs[i]))


_back_generic = back


def back(f, i, dy, y, *args, **kwargs):   # noqa: F811 -- dispatch for the variadic primitives
    if f is cat:
        return _cat_back(f, i, dy, y, *args, **kwargs)
    if f is cat1d:
        return _cat1d_back(f, i, dy, y, *args)
    return _back_generic(f, i, dy, y, *args, **kwargs)


# ----------------------------------------------------------------------------------------------
# Scalar (host Number) math: the reference's scalar known-answer tests (test/core.jl:35-49,
# test/highorder.jl:6-34) exercise the same rules on Julia Numbers.
# ----------------------------------------------------------------------------------------------

# ----------------------------------------------------------------------------------------------
# gradcheck (test/gradcheck.jl:35-104)
# ----------------------------------------------------------------------------------------------


def gcsum(f, *x, **kw):
    y = f(*x, **kw)
    v = value(y)
    if _isnum(v) or (isinstance(v, np.ndarray) and v.ndim == 0):
        return y
    if v.size == 0:
        return 0
    return sum_(y)


def gradcheck(f, *x, kw=None, args=None, nsample=10, rtol=0.05, atol=0.01, delta=1e-4, rng=None, verbose=0):
    kw = kw or {}
    rng = rng or np.random.default_rng(0)
    args = list(range(len(x))) if args is None else list(args)
    xrec = list(x)
    for i in args:
        xrec[i] = Param(xrec[i])
    result = diff(lambda: gcsum(f, *xrec, **kw))
    f0 = float(value(result))
    xptr = list(x)
    ok = True
    for i in args:
        g = full(grad(result, xrec[i]))
        xi = xptr[i]
        if _isnum(xi):
            d = type(xi)(delta) if not isinstance(xi, int) else delta
            x1 = xi + d if xi >= 0 else xi - d
            xptr[i] = x1
            f1 = float(value(gcsum(f, *xptr, **kw)))
            xptr[i] = xi
            nd = (f1 - f0) / float(x1 - xi)
            ad = 0.0 if g is None else float(g)
            good = bool(np.isclose(nd, ad, rtol=rtol, atol=atol))
            if verbose and not good:
                print("gradcheck fail", i, nd, ad)
            ok &= good
        else:
            n = xi.size
            idxs = range(n) if n <= nsample else rng.integers(0, n, nsample)
            flat = xi.reshape(-1, order="F")
            gflat = None if g is None else np.asarray(g).reshape(-1, order="F")
            for j in idxs:
                old = flat[j]
                d = xi.dtype.type(delta)
                flat[j] = old + d if old >= 0 else old - d
                f1 = float(value(gcsum(f, *xptr, **kw)))
                step = float(flat[j]) - float(old)
                flat[j] = old
                nd = (f1 - f0) / step
                ad = 0.0 if gflat is None else float(gflat[j])
                good = bool(np.isclose(nd, ad, rtol=rtol, atol=atol))
                if verbose and not good:
                    print("gradcheck fail", i, j, nd, ad)
                ok &= good
    return ok


# ---- input formats (examples/mnist.jl:77-88) ------------------------------------------------------------------
def xshape(a, nfeat=784, dtype=np.float32):
    """examples/mnist.jl:80  xshape(a) = reshape(a ./ 255f0, 784, div(length(a),784))  (column-major)."""
    a = np.asarray(a, dtype=np.uint8).reshape(-1, order="F")
    return np.asfortranarray((a.astype(dtype) / dtype(255)).reshape((nfeat, a.size // nfeat), order="F"))


def yshape(a, nclass=10, dtype=np.float32):
    """examples/mnist.jl:81  yshape(a) = (a[a.==0] = 10; full(sparse(Vector{Int}(a), 1:length(a), 1f0)))."""
    a = np.asarray(a, dtype=np.int64).reshape(-1).copy()
    a[a == 0] = 10
    y = np.zeros((nclass, a.size), dtype=dtype, order="F")
    y[a - 1, np.arange(a.size)] = 1
    return y


def charlm_loaddata(text, batch_size, char_limit=0):
    """examples/charlm.jl:166-203 restated: (data, chars, char_to_index (1-based), index_to_char); data[i] is the
    (vocab, batch_size) Float32 one-hot matrix with d[char_to_index[chars[i + nbatch*(j-1)]], j] = 1 (1-based i, j)."""
    chars, char_to_index = [], {}
    for c in text:
        char_to_index.setdefault(c, 1 + len(char_to_index))
        chars.append(c)
        if char_limit > 0 and len(chars) >= char_limit:
            break
    nbatch = len(chars) // batch_size
    data = []
    for i in range(1, nbatch + 1):
        d = np.zeros((len(char_to_index), batch_size), np.float32, order="F")
        for j in range(1, batch_size + 1):
            d[char_to_index[chars[(i + nbatch * (j - 1)) - 1]] - 1, j - 1] = 1
        data.append(d)
    index_to_char = [None] * len(char_to_index)
    for c, i in char_to_index.items():
        index_to_char[i - 1] = c
    return data, chars, char_to_index, index_to_char
