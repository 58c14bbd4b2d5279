"""Primitive catalogue: forward methods and gradient rules on device arrays.

Mirrors the reference's operator API (file:line under /root/reference):
  @primitive / @zerograd surface .......... src/macros.jl:88-194
  + - .* ./ .^ abs abs2 ..................... src/base.jl:20-49,73-74
  exp log tanh ... max min ................... src/math.jl:9-74
  sum / sum(abs2,.) / sum(abs,.) / mean ..... src/base.jl:245-247, src/statistics.jl:20-22
  * (matmul) and adjoint ..................... src/base.jl:26, src/linearalgebra.jl:16
  reshape / vec / copy / identity ............ src/base.jl:217-224,250,338,341
  unbroadcast ................................ src/unbroadcast.jl:9-26
  addto! ..................................... src/addto.jl:14-31,41-52,91-98
  getindex / ungetindex ...................... src/getindex.jl:32-33,61-79,91-95
  cat / uncat / cat1d ........................ src/cat.jl:27-31,46-92,122-123,150-182

Every rule has two forms.  When no outer tape is recording, the rule calls ONE fused kernel
(`dx = dy .* g(x,y)`).  Under an outer tape (higher-order gradients) the rule is the
reference's own expression over primitives, so that it records itself and stays differentiable.
"""
import math

import numpy as np

from . import darray as D
from ._lib import DimensionMismatch, UnsupportedOnDevice
from .core import Result, Tracked, Value, forw, recording, value
from .darray import DArray, IArray, isnum
from .sparse import Sparse, flat_dest

# ---------------------------------------------------------------------------------------------------
# definition helpers (src/macros.jl:88-194)
# ---------------------------------------------------------------------------------------------------


class Primitive:
    """`@primitive f(args...),dy,y  rule1 rule2 ...`: a callable that records itself on the active
    tapes when any argument is boxed."""

    def __init__(self, name, fwd, backs=(), only_first=False, y_unused=False, args_unused=False):
        self.name, self.fwd, self.backs, self.only_first = name, fwd, tuple(backs), only_first
        # which values the gradient rules read (core._mark_soft): y_unused = no rule reads the output;
        # args_unused = True (no rule reads any argument's value) or a tuple of per-argument flags
        self.y_unused, self.args_unused = y_unused, args_unused

    def __call__(self, *args, **kwargs):
        for a in args:
            if isinstance(a, Value):
                return forw(self, *args, **kwargs)
        return self.fwd(*args, **kwargs)

    def __repr__(self):
        return self.name


def primitive(name, fwd, *backs, **kw):
    return Primitive(name, fwd, backs, **kw)


def zerograd(name, fwd):
    """`@zerograd f(x)`: unbox, compute, never record (src/macros.jl:158-170)."""
    def z(*args, **kwargs):
        return fwd(*(value(a) for a in args), **kwargs)
    z.__name__ = name
    return z


def back(f, i, dy, y, *args, **kwargs):
    """back(f, Arg{i+1}, dy, y, args...; kwargs...)  (src/core.jl:165,174)."""
    vb = getattr(f, "vback", None)
    if vb is not None:
        return vb(i, dy, y, *args, **kwargs)
    rule = f.backs[i] if i < len(f.backs) else None
    if rule is None:
        if f.only_first or i < len(f.backs):
            return None
        raise TypeError("MethodError: no method matching back(%s, Arg{%d}, ...)" % (f.name, i + 1))
    return rule(dy, y, *args, **kwargs)


# ---------------------------------------------------------------------------------------------------
# size helpers (all @zerograd in the reference, src/base.jl:221-232)
# ---------------------------------------------------------------------------------------------------

def size(x, d=None):
    v = value(x)
    if isinstance(v, Sparse):
        v = v.container
    s = () if isnum(v) else tuple(v.shape)
    if d is None:
        return s
    return s[d] if d < len(s) else 1


def ndims(x):
    return len(size(x))


def length(x):
    n = 1
    for e in size(x):
        n *= e
    return n


def _isnumv(x):
    return isnum(value(x))


def _fusable(*xs):
    """Fused gradient kernels are legal only when nothing is being recorded."""
    return not recording()


# ---------------------------------------------------------------------------------------------------
# unbroadcast (src/unbroadcast.jl:9-26)
# ---------------------------------------------------------------------------------------------------

def unbroadcast(x, dx):
    sx, sdx = size(x), size(dx)
    if sx == sdx:
        return dx
    if _isnumv(x):
        return sum_(dx)
    if _isnumv(dx):
        return fill_like(value(x), dx)
    d = []
    for i in range(len(sdx)):
        if size(x, i) == sdx[i] and sdx[i] > 1:
            continue
        if size(x, i) != 1:
            raise DimensionMismatch("unbroadcast: %s vs %s" % (sx, sdx))
        d.append(i)
    return reshape(sum_(dx, dims=tuple(d)), sx)


def _fill_like_fwd(xv, v):
    if isinstance(v, DArray):      # device scalar
        return D.unary_bwd("identity", v, None, None, shape=xv.shape, dtype=xv.dtype)
    return D.full_like_shape(xv.shape, xv.dtype, v)


fill_like = primitive("fill!", _fill_like_fwd, None, lambda dy, y, x, v: sum_(dy))

# ---------------------------------------------------------------------------------------------------
# addto! (src/addto.jl:14-31,41-52,91-98)
# ---------------------------------------------------------------------------------------------------


def matches(a, b):
    if isnum(a) and isnum(b):
        return True
    if isinstance(a, DArray) and isinstance(b, DArray):
        return a.shape == b.shape
    if isinstance(a, DArray) and a.shape == () and isnum(b) or isinstance(b, DArray) and b.shape == () and isnum(a):
        return True
    return False


def _addto_fwd(a, b):
    if a is None:
        return b
    if b is None:
        return a
    if isinstance(a, Sparse) and isinstance(b, Sparse):
        assert a.container.shape == b.container.shape, "addto!: Sparse containers differ"
        a.values.extend(b.values)          # mutates the first argument (src/addto.jl:41-46)
        a.indices.extend(b.indices)
        return a
    if isinstance(a, Sparse):
        a, b = b, a
    if isinstance(b, Sparse):
        if isnum(a):
            for v in b.values:
                a = a + v
            return a
        assert a.shape == b.container.shape, "addto!: accumulator does not match the Sparse container"
        return b.addto_dense(a, inplace=not recording())   # copy first under an outer tape (:93)
    if isinstance(a, DArray) and isinstance(b, DArray):
        if a.tr and b.tr and a.shape == b.shape and a.ndim == 2:
            # two lazy adjoints (the gradients that reach an `adjoint` node's parent under an outer tape): a' + b' =
            # (a + b)', one pass over the stored arrays instead of two materialised transposes and a sum
            sa = DArray(a.buf, a.ptr, (a.shape[1], a.shape[0]), a.dtype)
            sb = DArray(b.buf, b.ptr, (b.shape[1], b.shape[0]), b.dtype)
            return D.adjoint(D.add(sa, sb))
        return D.add(a, b)                 # out of place: rules may alias dy (src/addto.jl:23)
    if isinstance(a, DArray) or isinstance(b, DArray):
        arr, num = (a, b) if isinstance(a, DArray) else (b, a)
        if arr.shape != ():
            raise DimensionMismatch("addto!: array %s vs Number" % (arr.shape,))
        return D.binary_fwd("add", arr, num)
    return a + b


class _AddTo(Primitive):
    def __call__(self, a, b):
        if isinstance(a, Value) or isinstance(b, Value):
            return forw(self, a, b)
        return _addto_fwd(a, b)


addto = _AddTo("addto!", _addto_fwd, (lambda dy, y, a, b: dy, lambda dy, y, a, b: dy))


def full(b):
    """full(::Sparse) densifies (src/sparse.jl:57); everything else passes through."""
    if isinstance(b, Sparse):
        return b.full()
    return b


# ---------------------------------------------------------------------------------------------------
# elementwise primitives
# ---------------------------------------------------------------------------------------------------

def _bin_fwd(op):
    def fwd(a, b):
        if isnum(a) and isnum(b):
            return D.HOST_BINARY[op](a, b)
        return D.binary_fwd(op, a, b)
    return fwd


_HOST_OK = {("mul", 1): 2, ("mul", 2): 1, ("div", 1): 2}   # (op, argno) -> the only operand the rule reads


def _bin_back(op, argno, composite):
    """Gradient of a broadcasting binary op w.r.t. argument `argno`: fused kernel + unbroadcast."""
    def rule(dy, y, x1, x2):
        xi = x1 if argno == 1 else x2
        other = _HOST_OK.get((op, argno))
        if other is not None and isnum(value(dy)) and isnum(value(x1 if other == 1 else x2)):
            return unbroadcast(xi, composite(dy, y, x1, x2))     # pure host-scalar arithmetic
        if _fusable() and (isinstance(value(x1), DArray) or isinstance(value(x2), DArray)
                           or isinstance(value(dy), DArray)):
            g = D.binary_bwd(op, argno, value(dy), value(x1), value(x2), value(y))
            return unbroadcast(xi, g)
        return unbroadcast(xi, composite(dy, y, x1, x2))
    return rule


def _mask(a, b):
    """`a .== b` as a non-differentiable 0/1 array (== is @zerograd, src/base.jl:57)."""
    a, b = value(a), value(b)
    if isnum(a) and isnum(b):
        return float(a == b)
    return D.binary_fwd("eq", a, b)


def _passthru(argno, negate=False):
    def rule(dy, y, x1, x2):
        xi = x1 if argno == 1 else x2
        if not negate:
            return unbroadcast(xi, dy)
        if _fusable() and isinstance(value(dy), DArray) and size(xi) != size(dy) and not _isnumv(xi):
            return neg(unbroadcast(xi, dy))      # -(sum dy) == sum(-dy) bit for bit, one pass less
        return unbroadcast(xi, neg(dy))
    return rule


add = primitive("+", _bin_fwd("add"), _passthru(1), _passthru(2), y_unused=True, args_unused=True)
sub = primitive("-", _bin_fwd("sub"), _passthru(1), _passthru(2, negate=True), y_unused=True, args_unused=True)
mul = primitive(".*", _bin_fwd("mul"),
                _bin_back("mul", 1, lambda dy, y, x1, x2: mul(dy, x2)),
                _bin_back("mul", 2, lambda dy, y, x1, x2: mul(x1, dy)))
div = primitive("./", _bin_fwd("div"),
                _bin_back("div", 1, lambda dy, y, x1, x2: div(dy, x2)),
                _bin_back("div", 2, lambda dy, y, x1, x2: div(mul(neg(dy), x1), abs2(x2))))


def dxndx(x1, x2, dy):
    """src/base.jl:49."""
    if _isnumv(x2):
        p = value(x2)
        if p == 0:
            return mul(dy, 0)
        if p == 1:
            return dy
        if p == 2:
            return mul(mul(2, x1), dy)
    return mul(mul(dy, x2), pow_(x1, sub(x2, 1)))


def _pow_back1(dy, y, x1, x2):
    if _isnumv(x2) and value(x2) in (0, 1, 2):
        return unbroadcast(x1, dxndx(x1, x2, dy))
    return _bin_back("pow", 1, lambda dy, y, x1, x2: dxndx(x1, x2, dy))(dy, y, x1, x2)


pow_ = primitive(".^", _bin_fwd("pow"), _pow_back1,
                 _bin_back("pow", 2, lambda dy, y, x1, x2: mul(mul(dy, y), log(x1))))
max_ = primitive("max", _bin_fwd("max"),
                 _bin_back("max", 1, lambda dy, y, x1, x2: mul(dy, _mask(y, x1))),
                 _bin_back("max", 2, lambda dy, y, x1, x2: mul(dy, _mask(y, x2))))
min_ = primitive("min", _bin_fwd("min"),
                 _bin_back("min", 1, lambda dy, y, x1, x2: mul(dy, _mask(y, x1))),
                 _bin_back("min", 2, lambda dy, y, x1, x2: mul(dy, _mask(y, x2))))
atan2 = primitive("atan", _bin_fwd("atan2"),
                  _bin_back("atan2", 1, lambda dy, y, x1, x2: mul(dy, div(x2, add(abs2(x1), abs2(x2))))),
                  _bin_back("atan2", 2, lambda dy, y, x1, x2: mul(dy, div(neg(x1), add(abs2(x1), abs2(x2))))))
hypot = primitive("hypot", _bin_fwd("hypot"),
                  _bin_back("hypot", 1, lambda dy, y, x1, x2: mul(dy, div(x1, y))),
                  _bin_back("hypot", 2, lambda dy, y, x1, x2: mul(dy, div(x2, y))))

# comparison masks are @zerograd (src/base.jl:52-60)
eq = zerograd("==", _bin_fwd("eq"))
lt = zerograd("<", _bin_fwd("lt"))
le = zerograd("<=", _bin_fwd("le"))
gt = zerograd(">", _bin_fwd("gt"))
ge = zerograd(">=", _bin_fwd("ge"))


def relu(x):
    """`max.(0, x)` (examples/mnist.jl:32)."""
    return max_(0, x)


def _un_fwd(op):
    def fwd(x):
        if isnum(x):
            return D.HOST_UNARY[op](x)
        return D.unary_fwd(op, x)
    return fwd


def _un_back(op, composite, needs_x, needs_y):
    def rule(dy, y, x):
        if _fusable() and isinstance(value(x), DArray):
            dyv = value(dy)
            xv, yv = value(x), value(y)
            if isnum(dyv) and not needs_x and not needs_y and xv.size == 1:
                return composite(dy, y, x)                       # scalar chain stays on the host
            if isnum(dyv) or dyv.size == xv.size or dyv.shape == ():
                return D.unary_bwd(op, dyv, xv if needs_x else None, yv if needs_y else None,
                                   shape=xv.shape, dtype=xv.dtype)
        return composite(dy, y, x)
    return rule


def _unary(name, op, composite, needs_x, needs_y):
    return primitive(name, _un_fwd(op), _un_back(op, composite, needs_x, needs_y),
                     y_unused=not needs_y, args_unused=(not needs_x,))


_LN2, _LN10 = math.log(2.0), math.log(10.0)

sign = zerograd("sign", _un_fwd("sign"))
one = zerograd("one", _un_fwd("one"))
identity = primitive("identity", lambda x: x, lambda dy, y, x: dy)
neg = _unary("-", "neg", lambda dy, y, x: neg(dy), False, False)
abs_ = _unary("abs", "abs", lambda dy, y, x: mul(dy, sign(x)), True, False)
abs2 = _unary("abs2", "abs2", lambda dy, y, x: mul(dy, mul(2, x)), True, False)
exp = _unary("exp", "exp", lambda dy, y, x: mul(dy, y), False, True)
log = _unary("log", "log", lambda dy, y, x: mul(dy, div(1, x)), True, False)
tanh = _unary("tanh", "tanh", lambda dy, y, x: mul(dy, sub(1, abs2(y))), False, True)
sigm = _unary("sigm", "sigm", lambda dy, y, x: mul(mul(dy, y), sub(1, y)), False, True)
sqrt = _unary("sqrt", "sqrt", lambda dy, y, x: mul(dy, div(1, mul(2, y))), False, True)
sin = _unary("sin", "sin", lambda dy, y, x: mul(dy, cos(x)), True, False)
cos = _unary("cos", "cos", lambda dy, y, x: mul(dy, neg(sin(x))), True, False)
tan = _unary("tan", "tan", lambda dy, y, x: mul(dy, add(1, abs2(y))), False, True)
sinh = _unary("sinh", "sinh", lambda dy, y, x: mul(dy, cosh(x)), True, False)
cosh = _unary("cosh", "cosh", lambda dy, y, x: mul(dy, sinh(x)), True, False)
asin = _unary("asin", "asin", lambda dy, y, x: mul(dy, div(1, sqrt(sub(1, abs2(x))))), True, False)
acos = _unary("acos", "acos", lambda dy, y, x: mul(dy, div(-1, sqrt(sub(1, abs2(x))))), True, False)
atan = _unary("atan", "atan", lambda dy, y, x: mul(dy, div(1, add(1, abs2(x)))), True, False)
asinh = _unary("asinh", "asinh", lambda dy, y, x: mul(dy, div(1, sqrt(add(1, abs2(x))))), True, False)
acosh = _unary("acosh", "acosh", lambda dy, y, x: mul(dy, div(1, sqrt(sub(abs2(x), 1)))), True, False)
atanh = _unary("atanh", "atanh", lambda dy, y, x: mul(dy, div(1, sub(1, abs2(x)))), True, False)
exp2 = _unary("exp2", "exp2", lambda dy, y, x: mul(dy, mul(y, _LN2)), False, True)
exp10 = _unary("exp10", "exp10", lambda dy, y, x: mul(dy, mul(y, _LN10)), False, True)
expm1 = _unary("expm1", "expm1", lambda dy, y, x: mul(dy, add(1, y)), False, True)
log2 = _unary("log2", "log2", lambda dy, y, x: mul(dy, div(1, mul(_LN2, x))), True, False)
log10 = _unary("log10", "log10", lambda dy, y, x: mul(dy, div(1, mul(_LN10, x))), True, False)
log1p = _unary("log1p", "log1p", lambda dy, y, x: mul(dy, div(1, add(1, x))), True, False)
cbrt = _unary("cbrt", "cbrt", lambda dy, y, x: mul(dy, div(1, mul(3, abs2(y)))), False, True)


# ---- the rest of src/math.jl: degree / reciprocal / pi-scaled trigonometric functions, sinc / cosc, clamp, ldexp,
# log(b, x).  The forward is composed from the table's ops on the unboxed arrays (deferred evaluation runs the chain as
# ONE fused launch, lazy.py); the gradient is the reference's own broadcast expression, line for line.
_D, _R, _PI = 180.0 / math.pi, math.pi / 180.0, math.pi


def _inv(x):
    return div(1, x)


def _comp(name, fwd, rule):
    return primitive(name, fwd, rule)


acosd = _comp("acosd", lambda x: mul(acos(x), _D), lambda dy, y, x: mul(dy, div(-_D, sqrt(sub(1, abs2(x))))))     # math.jl:10
acot = _comp("acot", lambda x: atan(_inv(x)), lambda dy, y, x: mul(dy, div(-1, add(1, abs2(x)))))                  # :12
acotd = _comp("acotd", lambda x: mul(atan(_inv(x)), _D), lambda dy, y, x: mul(dy, div(-_D, add(1, abs2(x)))))      # :13
acoth = _comp("acoth", lambda x: atanh(_inv(x)), lambda dy, y, x: mul(dy, div(1, sub(1, abs2(x)))))                # :14
acsc = _comp("acsc", lambda x: asin(_inv(x)),                                                                      # :15
             lambda dy, y, x: mul(dy, div(-1, sqrt(mul(mul(abs2(x), sub(x, 1)), add(x, 1))))))
acscd = _comp("acscd", lambda x: mul(asin(_inv(x)), _D),                                                           # :16
              lambda dy, y, x: mul(dy, div(-_D, sqrt(mul(mul(abs2(x), sub(x, 1)), add(x, 1))))))
acsch = _comp("acsch", lambda x: asinh(_inv(x)), lambda dy, y, x: mul(dy, div(-1, sqrt(add(abs2(abs2(x)), abs2(x))))))   # :17
asec = _comp("asec", lambda x: acos(_inv(x)), lambda dy, y, x: mul(dy, div(1, sqrt(sub(abs2(abs2(x)), abs2(x))))))       # :18
asecd = _comp("asecd", lambda x: mul(acos(_inv(x)), _D),                                                           # :19
              lambda dy, y, x: mul(dy, div(_D, sqrt(sub(abs2(abs2(x)), abs2(x))))))
asech = _comp("asech", lambda x: acosh(_inv(x)), lambda dy, y, x: mul(dy, div(-1, sqrt(sub(abs2(x), abs2(abs2(x)))))))   # :20
asind = _comp("asind", lambda x: mul(asin(x), _D), lambda dy, y, x: mul(dy, div(_D, sqrt(sub(1, abs2(x))))))       # :22
atand = _comp("atand", lambda x: mul(atan(x), _D), lambda dy, y, x: mul(dy, div(_D, add(1, abs2(x)))))             # :25
cosd = _comp("cosd", lambda x: cos(mul(x, _R)), lambda dy, y, x: mul(dy, div(mul(neg(sind(x)), _PI), 180)))        # :32
cospi = _comp("cospi", lambda x: cos(mul(x, _PI)), lambda dy, y, x: mul(dy, mul(neg(sinpi(x)), _PI)))              # :34
cot = _comp("cot", lambda x: _inv(tan(x)), lambda dy, y, x: mul(dy, neg(abs2(csc(x)))))                            # :35
cotd = _comp("cotd", lambda x: _inv(tan(mul(x, _R))), lambda dy, y, x: mul(dy, div(mul(neg(abs2(cscd(x))), _PI), 180)))  # :36
coth = _comp("coth", lambda x: _inv(tanh(x)), lambda dy, y, x: mul(dy, neg(abs2(csch(x)))))                        # :37
csc = _comp("csc", lambda x: _inv(sin(x)), lambda dy, y, x: mul(dy, mul(mul(-1, y), cot(x))))                      # :38
cscd = _comp("cscd", lambda x: _inv(sin(mul(x, _R))),                                                              # :39
             lambda dy, y, x: mul(dy, div(mul(mul(mul(-1, y), cotd(x)), _PI), 180)))
csch = _comp("csch", lambda x: _inv(sinh(x)), lambda dy, y, x: mul(dy, mul(mul(-1, y), coth(x))))                  # :40
deg2rad = _comp("deg2rad", lambda x: mul(x, _R), lambda dy, y, x: mul(dy, _R))                                     # :41
rad2deg = _comp("rad2deg", lambda x: mul(x, _D), lambda dy, y, x: mul(dy, _D))                                     # :58
sec = _comp("sec", lambda x: _inv(cos(x)), lambda dy, y, x: mul(dy, mul(y, tan(x))))                               # :60
secd = _comp("secd", lambda x: _inv(cos(mul(x, _R))), lambda dy, y, x: mul(dy, div(mul(mul(y, tand(x)), _PI), 180)))     # :61
sech = _comp("sech", lambda x: _inv(cosh(x)), lambda dy, y, x: mul(dy, mul(mul(-1, y), tanh(x))))                  # :62
sind = _comp("sind", lambda x: sin(mul(x, _R)), lambda dy, y, x: mul(dy, div(mul(cosd(x), _PI), 180)))             # :67
sinpi = _comp("sinpi", lambda x: sin(mul(x, _PI)), lambda dy, y, x: mul(dy, mul(cospi(x), _PI)))                   # :69
tand = _comp("tand", lambda x: tan(mul(x, _R)), lambda dy, y, x: mul(dy, div(mul(add(1, abs2(y)), _PI), 180)))     # :73


def _sinc_fwd(x):
    e = eq(x, 0)                      # sinc(0) = 1 without a branch: sin(pi x) / (pi x + [x == 0]) + [x == 0]
    return add(div(sin(mul(x, _PI)), add(mul(x, _PI), e)), e)


def _cosc_fwd(x):
    e = eq(x, 0)                      # cosc(0) = 0: (cos(pi x) - sinc(x)) / (x + [x == 0])
    return div(sub(cos(mul(x, _PI)), _sinc_fwd(x)), add(x, e))


sinc = _comp("sinc", _sinc_fwd, lambda dy, y, x: mul(dy, cosc(x)))                                                 # :65
cosc = _comp("cosc", _cosc_fwd, lambda dy, y, x: mul(dy, sub(div(mul(-2, y), x), mul(sinc(x), _PI ** 2))))         # :31
clamp = primitive("clamp", lambda x, lo, hi: min_(max_(x, lo), hi),                                                # :28
                  lambda dy, y, x, lo, hi: unbroadcast(x, mul(dy, mul(le(lo, x), le(x, hi)))))
ldexp = primitive("ldexp", lambda x, n: mul(x, 2.0 ** n), lambda dy, y, x, n: mul(dy, 2.0 ** n))                   # :47
logb = primitive("log", lambda x1, x2: div(log(x2), log(x1)),                                                      # :48 log(b, x)
                 lambda dy, y, x1, x2: unbroadcast(x1, mul(dy, div(neg(log(x2)), mul(x1, abs2(log(x1)))))),
                 lambda dy, y, x1, x2: unbroadcast(x2, mul(dy, div(1, mul(x2, log(x1))))))


def _round_fwd(name):
    def fwd(x):
        if isinstance(x, DArray):
            return D.special_param(name, 0, x)
        return float({"floor": math.floor, "ceil": math.ceil, "round": round, "trunc": math.trunc}[name](x))
    return fwd


_RMODES = {"nearest": 0, "tozero": 1, "down": 2, "up": 3}


def _sp_m_fwd(name, m=0):
    def fwd(x, *rest):
        mode = _RMODES[rest[0]] if rest else m
        if isinstance(x, DArray):
            return D.special_param(name, mode, x)
        if name == "exponent":
            return float(math.frexp(x)[1] - 1)
        if name == "significand":
            return x if x == 0 else math.frexp(x)[0] * 2.0
        p2 = 2 * math.pi
        return (math.remainder(x, p2) if mode == 0 else math.fmod(x, p2) if mode == 1 else
                x - p2 * math.floor(x / p2) if mode == 2 else x - p2 * math.ceil(x / p2))
    return fwd


exponent = zerograd("exponent", _sp_m_fwd("exponent"))                                                             # :44
significand = primitive("significand", _sp_m_fwd("significand"),                                                   # :64
                        lambda dy, y, x: mul(dy, exp2(neg(exponent(x)))))
mod2pi = primitive("mod2pi", _sp_m_fwd("rem2pi", 2), lambda dy, y, x: dy)                                          # :56
rem2pi = primitive("rem2pi", _sp_m_fwd("rem2pi"), lambda dy, y, x, r: dy)                                          # :59  r: "nearest" | "tozero" | "down" | "up"

# @zerograd floor / ceil / round / trunc (src/base.jl:79,100,145,151): forward only, their own pass (round: ties to even, like Julia)
floor = zerograd("floor", _round_fwd("floor"))
ceil = zerograd("ceil", _round_fwd("ceil"))
round_ = zerograd("round", _round_fwd("round"))
trunc = zerograd("trunc", _round_fwd("trunc"))


# special functions (src/specialfunctions.jl:20-21,29-30,34-42,63-67): the ones CUDA's math library provides, plus
# digamma / trigamma; composites are the reference's own gradient expressions (used under an outer tape)
_2SPI, _SPI2 = 2.0 / math.sqrt(math.pi), math.sqrt(math.pi) / 2.0
polygamma2 = zerograd("polygamma(2,·)", lambda x: D.unary_bwd("trigamma", 1.0, x, None))     # psi''(x), forward only
erf = _unary("erf", "erf", lambda dy, y, x: mul(dy, mul(exp(neg(abs2(x))), _2SPI)), True, False)
erfc = _unary("erfc", "erfc", lambda dy, y, x: mul(dy, mul(neg(exp(neg(abs2(x)))), _2SPI)), True, False)
erfcx = _unary("erfcx", "erfcx", lambda dy, y, x: mul(dy, sub(mul(mul(2, y), x), _2SPI)), True, True)
erfinv = _unary("erfinv", "erfinv", lambda dy, y, x: mul(dy, mul(exp(abs2(y)), _SPI2)), False, True)
erfcinv = _unary("erfcinv", "erfcinv", lambda dy, y, x: mul(dy, mul(neg(exp(abs2(y))), _SPI2)), False, True)
gamma = _unary("gamma", "gamma", lambda dy, y, x: mul(dy, mul(y, digamma(x))), True, True)
lgamma = _unary("lgamma", "lgamma", lambda dy, y, x: mul(dy, digamma(x)), True, False)
loggamma = lgamma
digamma = _unary("digamma", "digamma", lambda dy, y, x: mul(dy, trigamma(x)), True, False)
trigamma = _unary("trigamma", "trigamma", lambda dy, y, x: mul(dy, polygamma2(x)), True, False)
besselj0 = _unary("besselj0", "besselj0", lambda dy, y, x: mul(dy, neg(besselj1(x))), True, False)
bessely0 = _unary("bessely0", "bessely0", lambda dy, y, x: mul(dy, neg(bessely1(x))), True, False)
# order-1 Bessel gradients need order 2 (no forward primitive for it): the fused kernel only, no higher order
def _no_order2(name):
    def composite(dy, y, x):
        raise UnsupportedOnDevice("%s: only the first derivative is available on device (its rule needs order 2)" % name)
    return composite


besselj1 = _unary("besselj1", "besselj1", _no_order2("besselj1"), True, False)
bessely1 = _unary("bessely1", "bessely1", _no_order2("bessely1"), True, False)


# functions with an integer parameter (src/specialfunctions.jl:20,25,43,59): their own pass on the device
# (ag_special_param); the order is not differentiated (`nothing` in the reference's rules)
def _int_order(name, a):
    if isinstance(a, bool) or int(a) != a:
        raise UnsupportedOnDevice("%s: only integer orders are available on device (got %r)" % (name, a))
    return int(a)


def _sp_fwd(name):
    def fwd(a, x):
        if isinstance(x, DArray):
            return D.special_param(name, _int_order(name, a), x)
        import scipy.special as sp
        return float({"besselj": sp.jv, "bessely": sp.yv, "polygamma": sp.polygamma, "besseli": sp.iv, "besselk": sp.kv}[name](a, x))
    return fwd


besselj = primitive("besselj", _sp_fwd("besselj"), None,
                    lambda dy, y, a, x: mul(dy, div(sub(besselj(a - 1, x), besselj(a + 1, x)), 2)))
bessely = primitive("bessely", _sp_fwd("bessely"), None,
                    lambda dy, y, a, x: mul(dy, div(sub(bessely(a - 1, x), bessely(a + 1, x)), 2)))
# src/specialfunctions.jl:13,24
besseli = primitive("besseli", _sp_fwd("besseli"), None,
                    lambda dy, y, a, x: mul(dy, div(add(besseli(a - 1, x), besseli(a + 1, x)), 2)))
besselk = primitive("besselk", _sp_fwd("besselk"), None,
                    lambda dy, y, a, x: mul(dy, div(add(besselk(a - 1, x), besselk(a + 1, x)), -2)))
polygamma = primitive("polygamma", _sp_fwd("polygamma"), None,
                      lambda dy, y, a, x: unbroadcast(x, mul(dy, polygamma(a + 1, x))))


def _invdigamma_fwd(x):
    if isinstance(x, DArray):
        return D.special_param("invdigamma", 0, x)
    import scipy.optimize as so
    import scipy.special as sp
    return float(so.newton(lambda t: sp.digamma(t) - x, math.exp(x) + 0.5 if x >= -2.22 else -1.0 / (x + 0.5772156649015329)))


invdigamma = primitive("invdigamma", _invdigamma_fwd, lambda dy, y, x: mul(dy, div(1, trigamma(y))))


# src/specialfunctions.jl:5-8 (Airy functions), :33 (dawson), :39 (erfi): not in CUDA's math library, so they run as
# their own pass (ag_special_param, parameter ignored) instead of the unary table; rules as the reference writes them
def _sp1_fwd(name):
    def fwd(x):
        if isinstance(x, DArray):
            return D.special_param(name, 0, x)
        import scipy.special as sp
        host = dict(dawson=sp.dawsn, erfi=sp.erfi, airyai=lambda v: sp.airy(v)[0], airyaiprime=lambda v: sp.airy(v)[1],
                    airybi=lambda v: sp.airy(v)[2], airybiprime=lambda v: sp.airy(v)[3])
        return float(host[name](x))
    return fwd


airyai = primitive("airyai", _sp1_fwd("airyai"), lambda dy, y, x: mul(dy, airyaiprime(x)))
airyaiprime = primitive("airyaiprime", _sp1_fwd("airyaiprime"), lambda dy, y, x: mul(dy, mul(x, airyai(x))))
airybi = primitive("airybi", _sp1_fwd("airybi"), lambda dy, y, x: mul(dy, airybiprime(x)))
airybiprime = primitive("airybiprime", _sp1_fwd("airybiprime"), lambda dy, y, x: mul(dy, mul(x, airybi(x))))
dawson = primitive("dawson", _sp1_fwd("dawson"), lambda dy, y, x: mul(dy, add(mul(mul(-2, y), x), 1)))
erfi = primitive("erfi", _sp1_fwd("erfi"), lambda dy, y, x: mul(dy, mul(exp(abs2(x)), _2SPI)))


def _copy_fwd(x):
    return D.copy(x) if isinstance(x, DArray) else x


copy = primitive("copy", _copy_fwd, lambda dy, y, x: dy)

# ---------------------------------------------------------------------------------------------------
# reductions (src/base.jl:245-247; src/statistics.jl:20-22)
# ---------------------------------------------------------------------------------------------------


def _dims(dims):
    if dims is None:
        return None
    return (int(dims),) if isinstance(dims, (int, np.integer)) else tuple(int(d) for d in dims)


def _sum_fwd_map(mapname):
    def fwd(x, dims=None):
        if isnum(x):
            return D.HOST_UNARY[mapname](x) if mapname != "id" else x
        if dims is None:
            return D.sum_all(x, mapname)
        return D.sum_dims(x, _dims(dims), mapname)
    return fwd


def _expand_to(dy, x):
    """dy .* one.(x): broadcast dy (Number, device scalar or keep-dims array) to x's shape."""
    dyv, xv = value(dy), value(x)
    if isnum(xv):
        return dyv
    if isnum(dyv):
        return D.full_like_shape(xv.shape, xv.dtype, dyv)
    if dyv.shape == xv.shape:
        return dyv
    if dyv.size == 1:
        return D.unary_bwd("identity", dyv.reshape(()), None, None, shape=xv.shape, dtype=xv.dtype)
    return _bcast_to(dyv, xv.shape)


def _bcast_to(a, shape):
    return D.bcast_to(a, shape)


def _sum_back(dy, y, x, dims=None):
    if _fusable():
        return _expand_to(dy, x)
    return mul(dy, one(x))


def _sumabs2_back(dy, y, x, dims=None):
    if _fusable() and isinstance(value(x), DArray):
        dyv, xv = value(dy), value(x)
        if isnum(dyv) or dyv.size == 1:
            return D.unary_bwd("abs2", dyv if isnum(dyv) else dyv.reshape(()), xv, None)
        return D.binary_fwd("mul", dyv, D.binary_fwd("mul", 2.0, xv))
    return mul(dy, mul(2, x))


def _sumabs_back(dy, y, x, dims=None):
    if _fusable() and isinstance(value(x), DArray):
        dyv, xv = value(dy), value(x)
        if isnum(dyv) or dyv.size == 1:
            return D.unary_bwd("abs", dyv if isnum(dyv) else dyv.reshape(()), xv, None)
    return mul(dy, sign(x))


sum_ = primitive("sum", _sum_fwd_map("id"), _sum_back)
sumabs2 = primitive("sum(abs2,·)", _sum_fwd_map("abs2"), _sumabs2_back)
sumabs = primitive("sum(abs,·)", _sum_fwd_map("abs"), _sumabs_back)


def _extremum(op, hostf):
    def fwd(x, dims=None):
        if isnum(x):
            return x
        r = D.reduce_dims(x, _dims(dims), op)
        return r.reshape(()) if dims is None else r
    return fwd


def _extremum_back(dy, y, x, dims=None):
    """maximum/minimum: dy.*(y.==x)  (src/base.jl:202,205); y, dy broadcast against x."""
    if _fusable() and isinstance(value(x), DArray):
        xv, yv, dyv = value(x), value(y), value(dy)
        if isinstance(yv, DArray) and yv.ndim < xv.ndim:
            yv = yv.reshape((1,) * xv.ndim)
        if isinstance(dyv, DArray) and dyv.ndim < xv.ndim:
            dyv = dyv.reshape((1,) * xv.ndim)
        return D.binary_bwd("max", 2, dyv, None, xv, yv)
    return mul(dy, _mask(y, x))


# maximum(abs, x; dims), minimum(abs, x; dims): dy.*(y.==abs.(x)).*sign.(x)  (src/base.jl:203,206)
def _absext(op):
    def fwd(x, dims=None):
        if isnum(x):
            return abs(x)
        r = D.reduce_dims(abs_.fwd(x), _dims(dims), op)
        return r.reshape(()) if dims is None else r
    return fwd


def _absext_back(dy, y, x, dims=None):
    return mul(mul(dy, _mask(y, abs_(value(x)))), sign(x))


maxabs = primitive("maximum(abs)", _absext("max"), _absext_back)
minabs = primitive("minimum(abs)", _absext("min"), _absext_back)
maximum = primitive("maximum", _extremum("max", max), _extremum_back)
minimum = primitive("minimum", _extremum("min", min), _extremum_back)
prod = primitive("prod", _extremum("prod", None), lambda dy, y, x, dims=None: mul(dy, div(y, x)))


def _invperm(d):
    r = [0] * len(d)
    for i, p_ in enumerate(d):
        r[p_] = i
    return tuple(r)


permutedims = primitive("permutedims", lambda x, d: D.permutedims(x, d),
                        lambda dy, y, x, d: permutedims(dy, _invperm(d)))


def _mean_fwd(x, dims=None):
    s = _sum_fwd_map("id")(x, dims)
    return div.fwd(s, length(x) // max(length(s), 1))


mean = primitive("mean", _mean_fwd,
                 lambda dy, y, x, dims=None: div(_sum_back(dy, y, x, dims), length(x) // length(dy)))

def _meanmap_fwd(which):
    def fwd(x, dims=None):
        s = _sum_fwd_map(which)(x, dims)
        return div.fwd(s, length(x) // max(length(s), 1))
    return fwd


# mean(abs, x; dims), mean(abs2, x; dims)  (src/statistics.jl:9-10)
meanabs = primitive("mean(abs)", _meanmap_fwd("abs"),
                    lambda dy, y, x, dims=None: div(mul(dy, sign(x)), length(x) // length(dy)))
meanabs2 = primitive("mean(abs2)", _meanmap_fwd("abs2"),
                     lambda dy, y, x, dims=None: div(mul(dy, mul(2, x)), length(x) // length(dy)))


def varm(x, m, dims=None, corrected=True):
    """varm(x, m; corrected, dims) (src/statistics.jl:27)."""
    return var(x, dims=dims, corrected=corrected, mean_=m)


def stdm(x, m, dims=None, corrected=True):
    return sqrt(varm(x, m, dims=dims, corrected=corrected))


def var(x, dims=None, corrected=True, mean_=None):
    """var as a composite over primitives (src/statistics.jl:24-35): sum(abs2, x .- m) / (n - corrected)."""
    m = mean(x, dims=dims) if mean_ is None else mean_
    n = length(x) if dims is None else length(x) // length(m)
    return div(sumabs2(sub(x, m), dims=dims), n - (1 if corrected else 0))


def std(x, dims=None, corrected=True, mean_=None):
    return sqrt(var(x, dims=dims, corrected=corrected, mean_=mean_))


def norm(x, p=2):
    """norm(x), norm(x, p) of the vectorised array (src/linearalgebra.jl:64-65) as composites whose derivatives are the
    branches of `normback` (:222-238): p = 2 sqrt(sum(abs2, x)); 1 sum(abs, x); Inf / -Inf maximum / minimum of abs.(x)
    (ties share the gradient, like `dy*sign.(x).*(abs.(x).==y)`); 0 the number of nonzeros (zero gradient); otherwise
    sum(abs.(x).^p)^(1/p)."""
    if p == 2:
        return sqrt(sumabs2(x))
    if p == 1:
        return sumabs(x)
    if p == math.inf:
        return maximum(abs_(x))
    if p == -math.inf:
        return minimum(abs_(x))
    if p == 0:
        xv = value(x)
        return D.binary_fwd("sub", float(length(xv)), D.sum_all(D.binary_fwd("eq", xv, 0.0)))
    return pow_(sum_(pow_(abs_(x), p)), 1.0 / p)


def dot(a, b):
    """dot(x1,x2) = sum(x1 .* x2)  (src/linearalgebra.jl:23)."""
    return sum_(mul(a, b))


# ---------------------------------------------------------------------------------------------------
# linear algebra (src/base.jl:26; src/linearalgebra.jl:16,87)
# ---------------------------------------------------------------------------------------------------


def _matmul_fwd(a, b):
    if isnum(a) or isnum(b) or (isinstance(a, DArray) and a.shape == ()) or (isinstance(b, DArray) and b.shape == ()):
        return _bin_fwd("mul")(a, b)
    return D.matmul(a, b)


def _scalarish(x):
    v = value(x)
    return isnum(v) or (isinstance(v, DArray) and v.shape == ())


def _matmul_back1(dy, y, x1, x2):
    if _scalarish(x1) or _scalarish(x2):
        return unbroadcast(x1, mul(dy, x2))
    return matmul(dy, adjoint(x2))


def _matmul_back2(dy, y, x1, x2):
    if _scalarish(x1) or _scalarish(x2):
        return unbroadcast(x2, mul(x1, dy))
    return matmul(adjoint(x1), dy)


matmul = primitive("*", _matmul_fwd, _matmul_back1, _matmul_back2, y_unused=True)


def _adjoint_fwd(x):
    return x if isnum(x) else D.adjoint(x)


adjoint = primitive("adjoint", _adjoint_fwd,
                    lambda dy, y, x: reshape(adjoint(dy), size(x)) if ndims(x) == 1 else adjoint(dy))
transpose = adjoint


def matpow(x, p):
    """`A^p` on a matrix has no gradient by design (src/base.jl:44)."""
    if isinstance(x, Value) and ndims(x) >= 1:
        raise RuntimeError("A^B grad undefined")
    raise UnsupportedOnDevice("matrix power on device")


# ---------------------------------------------------------------------------------------------------
# shape rules (src/base.jl:217-224,250)
# ---------------------------------------------------------------------------------------------------

def _reshape_fwd(x, shape):
    shape = tuple(int(s) for s in (shape if isinstance(shape, (tuple, list)) else (shape,)))
    if isnum(x):
        return x
    return x.reshape(shape)


reshape = primitive("reshape", _reshape_fwd, lambda dy, y, x, shape: reshape(dy, size(x)))
vec = primitive("vec", lambda x: x.reshape((x.size,)), lambda dy, y, x: reshape(dy, size(x)))

# ---------------------------------------------------------------------------------------------------
# getindex / ungetindex (src/getindex.jl:32-33,61-79,91-95)
# ---------------------------------------------------------------------------------------------------


def _col_index(x, idx):
    """(Colon, Vector{Int}) on a matrix -> the column ids, else None (embedding lookup fast path)."""
    if x.ndim == 2 and len(idx) == 2 and isinstance(idx[0], slice) and idx[0] == slice(None):
        j = idx[1]
        if isinstance(j, IArray):
            return j
        if isinstance(j, (list, np.ndarray)) and np.asarray(j).dtype != np.bool_ and np.asarray(j).ndim == 1:
            return np.asarray(j, dtype=np.int64)
    return None


def _getindex_fwd(x, *idx):
    if isnum(x):
        return x
    cols = _col_index(x, idx)
    if cols is not None:
        lo, hi = (cols.lo, cols.hi) if isinstance(cols, IArray) else ((cols.min(), cols.max()) if cols.size else (0, -1))
        if cols.size and (lo < 0 or hi >= x.shape[1]):
            raise IndexError("BoundsError")
        return D.gather_cols(x, cols)
    idx = tuple(i.host if isinstance(i, IArray) else i for i in idx)
    dest, outshape = flat_dest(x.shape, idx)
    return D.gather(x, dest, outshape)


def _ungetindex_dense(x, dxi, idx):
    return Sparse(x, [dxi], [idx]).full()


def ungetindex(x, dxi, idx):
    """src/getindex.jl:61-79 (Array branch), :91-92 (Value unboxing)."""
    x = value(x)
    if isinstance(dxi, Value):
        if isnum(x):
            return dxi
        return forw(addto, D.zeros(x.shape, x.dtype), forw(_ungetindex_p, x, dxi, idx))
    if isnum(x):
        return dxi
    if recording():
        return _ungetindex_dense(x, dxi, idx)
    return Sparse(x, [dxi], [idx])


_ungetindex_p = primitive("ungetindex", _ungetindex_dense, None,
                          lambda ddx, dx, x, dxi, idx: getindex(ddx, *idx), None)
getindex = primitive("getindex", _getindex_fwd, lambda dxi, xi, x, *idx: ungetindex(x, dxi, idx),
                     only_first=True)

# ---------------------------------------------------------------------------------------------------
# the @zerograd helpers of src/base.jl:52-62,75-160 that touch array VALUES: predicates, counts, integer division.
# Each unboxes, computes on the device with the table's ops, and is never recorded.
# ---------------------------------------------------------------------------------------------------

def _z(name, f):
    return zerograd(name, f)


ne = _z("!=", lambda a, b: sub(1, eq(a, b)))
isnan = _z("isnan", lambda x: sub(1, eq(x, x)))
isinf = _z("isinf", lambda x: eq(abs_(x), math.inf))
isfinite = _z("isfinite", lambda x: lt(abs_(x), math.inf))
isinteger = _z("isinteger", lambda x: mul(isfinite(x), eq(x, trunc(x))))
signbit = _z("signbit", lambda x: lt(atan2(x, -1.0), 0))        # atan2(-0.0, -1) = -pi: the sign bit of zeros too


def _count(m):
    v = sum_(m)
    return int(round(float(v.to_numpy()) if isinstance(v, DArray) else float(v)))


argmax = _z("argmax", lambda x: D.argext(x, True) if isinstance(x, DArray) else 0)      # 0-based linear (column-major) index
argmin = _z("argmin", lambda x: D.argext(x, False) if isinstance(x, DArray) else 0)
count = _z("count", _count)                                     # count(mask): number of non-zero mask entries
any_ = _z("any", lambda m: _count(m) > 0)
all_ = _z("all", lambda m: _count(m) == length(m))
isequal = _z("isequal", lambda a, b: size(a) == size(b) and _count(add(eq(a, b), mul(isnan(a), isnan(b)))) == length(a))
isapprox = _z("isapprox", lambda a, b, rtol=None, atol=0.0: float(value(norm(sub(a, b)))) <= max(
    atol, (math.sqrt(float(np.finfo(value(a).dtype if isinstance(value(a), DArray) else np.float64).eps)) if rtol is None else rtol)
    * max(float(value(norm(a))), float(value(norm(b))))))
div_ = _z("div", lambda a, b: trunc(div(a, b)))                  # div(x, y) = trunc(x / y)  (@zerograd, src/base.jl:88)
# rem is differentiable (src/base.jl:144): rem(x1, x2) = x1 - x2 * div(x1, x2);  dx1 = dy, dx2 = -dy .* div.(x1, x2)
rem = primitive("rem", lambda a, b: sub(a, mul(b, trunc(div(a, b)))),
                lambda dy, y, x1, x2: unbroadcast(x1, dy),
                lambda dy, y, x1, x2: unbroadcast(x2, mul(neg(dy), div_(x1, x2))))
# x1 .\ x2 = x2 ./ x1 (src/base.jl:39-42)
ldiv = primitive("\\", lambda x1, x2: div(x2, x1),
                 lambda dy, y, x1, x2: unbroadcast(x1, div(mul(neg(dy), x2), abs2(x1))),
                 lambda dy, y, x1, x2: unbroadcast(x2, div(dy, x1)))
unsafe_trunc = trunc



def _dtype_of(t):
    t = value(t)
    return np.dtype(t.dtype if isinstance(t, DArray) else (t if isinstance(t, (type, np.dtype)) else np.float64))


def _oftype_fwd(t, x):
    return D.astype(x, _dtype_of(t)) if isinstance(x, DArray) else x


# oftype(t, x): x converted to t's eltype; the gradient is converted back to x's (src/base.jl:365).  `t` is an array,
# an eltype (np.float32 / np.float64) or a Value of either.
oftype = primitive("oftype", _oftype_fwd, None, lambda dy, y, t, x: oftype(x, dy))
big = primitive("big", lambda x: _oftype_fwd(np.float64, x), lambda dy, y, x: oftype(x, dy))      # src/base.jl:76: widest eltype here is Float64
widemul = mul                                                    # src/base.jl:156: the rules of .*
float_ = identity                                                # src/base.jl:98: dy passes through
eltype = _z("eltype", lambda x: x.dtype if isinstance(x, DArray) else type(x))
eps = _z("eps", lambda x: float(np.finfo(x.dtype).eps) if isinstance(x, DArray) else float(np.spacing(x)))
typemax = _z("typemax", lambda x: math.inf)
typemin = _z("typemin", lambda x: -math.inf)
sizeof = _z("sizeof", lambda x: x.size * np.dtype(x.dtype).itemsize if isinstance(x, DArray) else 8)
isempty = _z("isempty", lambda x: length(x) == 0)
axes = _z("axes", lambda x: tuple(range(n) for n in size(x)))
eachindex = _z("eachindex", lambda x: range(length(x)))
lastindex = _z("lastindex", lambda x: length(x) - 1)            # 0-based like every index on the Python side
strides = _z("strides", lambda x: tuple(int(np.prod(size(x)[:d])) for d in range(ndims(x))))   # column-major
stride = _z("stride", lambda x, d: int(np.prod(size(x)[:d])))
similar = _z("similar", lambda x: DArray.empty(x.shape, x.dtype))
zero = _z("zero", lambda x: D.zeros(x.shape, x.dtype) if isinstance(x, DArray) else type(x)(0))
ones = _z("ones", lambda x: D.full_like_shape(x.shape, x.dtype, 1.0) if isinstance(x, DArray) else D.full_like_shape(tuple(x), np.float64, 1.0))


# ---------------------------------------------------------------------------------------------------
# diag / diagm / tr / tril / triu / kron (src/linearalgebra.jl:19-20,46-48,86,90,92,176-186)
# ---------------------------------------------------------------------------------------------------

def _diag_positions(shape, k):
    """0-based column-major positions of the k-th diagonal of an (m, n) matrix."""
    m, n = shape
    cnt = max(0, min(m, n - k) if k >= 0 else min(m + k, n))
    j = np.arange(cnt, dtype=np.int64)
    return (j + (j + k) * m) if k >= 0 else ((j - k) + j * m)


def diag(x, k=0):
    """`diag(x)`, `diag(x, k)`: the diagonal read through `getindex` with linear positions, so its gradient is
    getindex's own (a Sparse block scattered into zeros = the reference's `diagm(k => dy)`, src/linearalgebra.jl:19-20,
    without its square-matrices-only restriction)."""
    shp = size(x)
    if len(shp) != 2:
        raise DimensionMismatch("diag: a matrix is required")
    return getindex(x, _diag_positions(shp, int(k)))


def _diagm_fwd(v, k=0):
    n = v.size + abs(int(k))
    return _ungetindex_dense(D.zeros((n, n), v.dtype), reshape(v, (v.size,)), (_diag_positions((n, n), int(k)),))


diagm = primitive("diagm", _diagm_fwd, lambda dy, y, v, k=0: reshape(diag(dy, k), size(v)))   # :22 (commented there)


def tr(x):
    """`tr(x)` = sum of the diagonal; gradient dy * I (src/linearalgebra.jl:86) through sum and getindex."""
    return sum_(diag(x))


# det / inv / logdet / logabsdet (src/linearalgebra.jl:18,28,57-58): one-launch elimination (ag_inv_det); the rules are
# the reference's expressions over inv, adjoint and the matrix product
def _inv_fwd(x):
    if not isinstance(x, DArray):
        return 1.0 / x
    return D.inv_det(x)[0]


def _det_fwd(x):
    return D.inv_det(x, want_inv=False)[1][0] if isinstance(x, DArray) else x


def _logdet_fwd(x):
    if not isinstance(x, DArray):
        return math.log(x)
    _, (_, lad, sign) = D.inv_det(x, want_inv=False)
    if sign < 0:
        raise ValueError("DomainError: logdet of a matrix with negative determinant")
    return lad


def _yt_dy_yt(dy, y):
    yt = adjoint(y)
    return neg(matmul(matmul(yt, dy), yt))


inv = primitive("inv", _inv_fwd, lambda dy, y, x: _yt_dy_yt(dy, y))                                # :28  (yt=y'; -yt*dy*yt)
det = primitive("det", _det_fwd, lambda dy, y, x: mul(mul(dy, y), adjoint(inv(x))))                # :18  dy*y*inv(x)'
logdet = primitive("logdet", _logdet_fwd, lambda dy, y, x: mul(dy, adjoint(inv(x))))               # :58  dy*inv(x)'


def logabsdet(x):
    """`logabsdet(x)` -> (log|det|, sign); the gradient flows through the first element (src/linearalgebra.jl:57:
    dy[1]*inv(x)'), the sign is a constant."""
    v = value(x)
    sign = D.inv_det(v, want_inv=False)[1][2] if isinstance(v, DArray) else (1.0 if v >= 0 else -1.0)
    return _logabsdet1(x), sign


def _lad_fwd(x):
    return D.inv_det(x, want_inv=False)[1][1] if isinstance(x, DArray) else math.log(abs(x))


_logabsdet1 = primitive("logabsdet", _lad_fwd, lambda dy, y, x: mul(dy, adjoint(inv(x))))


_TRI_MASKS = {}


def _tri_mask(x, lower):
    """`tril(fill!(similar(x), 1))` / `triu(...)` of the rules (src/linearalgebra.jl:90,92): uploaded once per shape."""
    v = value(x)
    key = (tuple(v.shape), np.dtype(v.dtype).str, lower)
    m = _TRI_MASKS.get(key)
    if m is None:
        ones_ = np.ones(v.shape, dtype=v.dtype)
        m = _TRI_MASKS[key] = DArray.from_numpy(np.asfortranarray(np.tril(ones_) if lower else np.triu(ones_)))
    return m


tril = primitive("tril", lambda x: mul(x, _tri_mask(x, True)), lambda dy, y, x: mul(dy, _tri_mask(x, True)))
triu = primitive("triu", lambda x: mul(x, _tri_mask(x, False)), lambda dy, y, x: mul(dy, _tri_mask(x, False)))


def kron(a, b):
    """`_kron` (src/linearalgebra.jl:176-186): reshape, one broadcast product, reshape -- all recorded primitives."""
    if ndims(a) != 2 or ndims(b) != 2:
        raise DimensionMismatch("Only support matrices for the time being")
    (m1, n1), (m2, n2) = size(a), size(b)
    return reshape(mul(reshape(a, (1, m1, 1, n1)), reshape(b, (m2, 1, n2, 1))), (m1 * m2, n1 * n2))


# view(x, i...) (src/getindex.jl:36): same rule as getindex.  Tracked buffers are never overwritten while recording
# (src/getindex.jl:12-28), so a copy of the selection is indistinguishable from an aliasing view on this path.
view = primitive("view", _getindex_fwd, lambda dxi, xi, x, *idx: ungetindex(x, dxi, idx), only_first=True)

# ---------------------------------------------------------------------------------------------------
# cat / uncat (src/cat.jl:27-31,46-92,122-123) and cat1d (src/cat.jl:150-182)
# ---------------------------------------------------------------------------------------------------


def _shape_nd(x, nd):
    s = size(x)
    if s == () and isnum(value(x)):
        s = (1,)
    return tuple(s) + (1,) * (nd - len(s))


def _cat_fwd(*xs, dims):
    arrs = [x for x in xs if isinstance(x, DArray)]
    if not arrs:
        raise UnsupportedOnDevice("cat of host Numbers only")
    dtype = arrs[0].dtype
    nd = max(max(ndims(x) for x in xs), dims + 1)
    shapes = [_shape_nd(x, nd) for x in xs]
    for s in shapes:
        for d in range(nd):
            if d != dims and s[d] != shapes[0][d]:
                raise DimensionMismatch("cat: mismatch in dimension %d" % (d + 1))
    tot = sum(s[dims] for s in shapes)
    out_shape = tuple(tot if d == dims else shapes[0][d] for d in range(nd))
    out = D.DArray.empty(out_shape, dtype)
    inner = D._prod(out_shape[:dims])
    outer = D._prod(out_shape[dims + 1:])
    srcs = []
    for x, s in zip(xs, shapes):
        if isnum(x):
            x = D.full_like_shape((1,), dtype, x)
        srcs.append((D.materialize(x), s[dims]))
    if out.size:
        for i in range(0, len(srcs), 8):                   # ag_cat: up to 8 sources per launch
            part = srcs[i:i + 8]
            if i == 0 and len(srcs) <= 8:
                D.cat_into(out, part, inner, outer)
            else:                                          # more than 8 sources: slab by slab
                off = sum(l for _, l in srcs[:i])
                for x, l in part:
                    if l:
                        D.slab_copy(x, l, 0, out, tot, off, inner, l, outer, fresh=True)
                    off += l
    return out


def _uncat_range(y1, argn, dims, xs):
    lo = 0
    for j in range(argn):
        lo += size(xs[j], dims) if ndims(xs[j]) > dims or not _isnumv(xs[j]) else 1
    ln = size(xs[argn], dims) if not _isnumv(xs[argn]) else 1
    return lo, ln


def _uncat_fwd(y1, argn, dims, *xs):
    lo, ln = _uncat_range(y1, argn, dims, xs)
    shp = y1.shape + (1,) * (dims + 1 - y1.ndim)
    inner, tot, outer = D._prod(shp[:dims]), shp[dims], D._prod(shp[dims + 1:])
    target = size(xs[argn]) if not _isnumv(xs[argn]) else (1,)
    out = D.DArray.empty(target, y1.dtype)
    if out.size:
        D.slab_copy(D.materialize(y1), tot, lo, out, ln, 0, inner, ln, outer, fresh=True)
    return out


def _uncat1_fwd(x2, y1, argn, dims, *xs):
    lo, ln = _uncat_range(y1, argn, dims, xs)
    shp = y1.shape + (1,) * (dims + 1 - y1.ndim)
    inner, tot, outer = D._prod(shp[:dims]), shp[dims], D._prod(shp[dims + 1:])
    out = D.zeros(y1.shape, y1.dtype)
    if isinstance(x2, DArray) and x2.size:
        D.slab_copy(D.materialize(x2), ln, 0, out, tot, lo, inner, ln, outer, fresh=True)
    return out


uncat = primitive("uncat", _uncat_fwd,
                  lambda x2, x1, y1, argn, dims, *xs: uncat1(x2, y1, argn, dims, *xs), only_first=True)
uncat1 = primitive("uncat1", _uncat1_fwd,
                   lambda y3, y2, x2, y1, argn, dims, *xs: uncat(y3, argn, dims, *xs), only_first=True)


class _Cat(Primitive):
    def __call__(self, *args, dims):
        for a in args:
            if isinstance(a, Value):
                return forw(self, *args, dims=dims)
        return self.fwd(*args, dims=dims)

    @staticmethod
    def vback(i, dy, y, *xs, dims):
        return uncat(dy, i, dims, *xs)


cat = _Cat("cat", _cat_fwd)


def vcat(*xs):
    return cat(*xs, dims=0)


def hcat(*xs):
    return cat(*xs, dims=1)


def _cat1d_fwd(*xs):
    arrs = [D.materialize(x) if isinstance(x, DArray) else x for x in xs]
    dtype = next(a.dtype for a in arrs if isinstance(a, DArray))
    tot = sum(length(a) for a in arrs)
    out = D.DArray.empty((tot,), dtype)
    off, segs, isz = 0, [], np.dtype(dtype).itemsize
    for a in arrs:
        if isnum(a):
            a = D.full_like_shape((1,), dtype, a)
        segs.append((a.ptr, out.ptr + off * isz, a.size))
        off += a.size
    D.multi_copy_flat(out.dt, segs)
    return out


class _Cat1d(Primitive):
    @staticmethod
    def vback(i, y1, y, *xs):
        off2 = 0
        for j in range(i + 1):
            off2 += length(xs[j])
        off1 = off2 - length(xs[i])
        return reshape(getindex(y1, slice(off1, off2)), size(xs[i]))


cat1d = _Cat1d("cat1d", _cat1d_fwd)

# ---------------------------------------------------------------------------------------------------
# update (README.md:70-77; src/sparse.jl:50-54; src/core.jl:55)
# ---------------------------------------------------------------------------------------------------


def axpy(a, x, y):
    """axpy!(a, x, y): y .+= a .* x in place; y may be a Param; x may be Sparse (src/sparse.jl:50)."""
    yv = value(y)
    if x is None:
        return y
    if isinstance(x, Sparse):
        (x * a).addto_dense(yv, inplace=True)
    else:
        D.axpy_(a, x, yv)
    return y
