"""User-defined elementwise primitives compiled to ONE fused kernel.

Reference: `@primitive f(x),dy,y (expr)` (src/macros.jl:88-105) lets a user attach an arbitrary broadcast
expression as the gradient of a new primitive, e.g. examples/charlm.jl:46-47

    sigm(x) = 1 ./ (1 .+ exp.(-x));   @primitive sigm(x),dy,y  (dy .* y .* (1 .- y))

Here the same definition is written once against the symbol namespace `E`:

    sigm = primitive_expr("sigm", lambda x: 1 / (1 + E.exp(-x)), lambda dy, y, x: dy * y * (1 - y))

`primitive_expr` traces the lambdas with symbolic operands and compiles them to postfix bytecode for
ag_ewise_vm (one pass over memory, no temporaries).  Under an outer tape (higher-order gradients) or for
operands that need real broadcasting, the very same lambdas are evaluated with the ordinary primitives, so the
gradient expression records itself exactly like the reference's does.
"""
import ctypes as C

import numpy as np

from . import darray as D
from ._lib import lib, check, UnsupportedOnDevice
from .core import Value, recording, value
from .darray import DArray, isnum

_BIN = {"add": 2, "sub": 3, "mul": 4, "div": 5, "max": 7, "min": 8, "pow": 9, "eq": 10, "lt": 11, "le": 12, "gt": 13,
        "ge": 14}
_NEG, _UNARY0 = 6, 15


class Expr:
    """Symbolic elementwise expression node."""
    __slots__ = ("op", "args", "val")
    __array_priority__ = 4000
    __array_ufunc__ = None

    def __init__(self, op, args=(), val=None):
        self.op, self.args, self.val = op, args, val

    @staticmethod
    def wrap(x):
        if isinstance(x, Expr):
            return x
        if isnum(x):
            return Expr("const", val=float(x))
        raise TypeError("cannot use %r inside a primitive_expr expression" % (type(x).__name__,))

    def _b(self, op, o, swap=False):
        a, b = (Expr.wrap(o), self) if swap else (self, Expr.wrap(o))
        return Expr(op, (a, b))

    def __add__(self, o): return self._b("add", o)
    def __radd__(self, o): return self._b("add", o, True)
    def __sub__(self, o): return self._b("sub", o)
    def __rsub__(self, o): return self._b("sub", o, True)
    def __mul__(self, o): return self._b("mul", o)
    def __rmul__(self, o): return self._b("mul", o, True)
    def __truediv__(self, o): return self._b("div", o)
    def __rtruediv__(self, o): return self._b("div", o, True)
    def __pow__(self, o): return self._b("pow", o)
    def __rpow__(self, o): return self._b("pow", o, True)
    def __neg__(self): return Expr("neg", (self,))


def compile_expr(e):
    """Expr -> (list of int32 codes, list of constants), postfix."""
    code, consts = [], []

    def emit(n):
        if n.op == "in":
            code.append(0 | (n.val << 8))
        elif n.op == "const":
            if n.val not in consts:
                consts.append(n.val)
            code.append(1 | (consts.index(n.val) << 8))
        elif n.op == "neg":
            emit(n.args[0])
            code.append(_NEG)
        elif n.op in _BIN:
            emit(n.args[0])
            emit(n.args[1])
            code.append(_BIN[n.op])
        else:
            emit(n.args[0])
            code.append(_UNARY0 + D.U[n.op])
    emit(e)
    if len(code) > 64 or len(consts) > 16:
        raise UnsupportedOnDevice("expression too long for the fused evaluator (%d instructions)" % len(code))
    return code, consts


class _Namespace:
    """Functions usable inside primitive_expr lambdas: symbolic on Expr operands, the ordinary primitives otherwise."""

    def __getattr__(self, name):
        from . import rules
        uname = {"abs": "abs"}.get(name, name)
        if uname in D.U:
            real = getattr(rules, {"abs": "abs_"}.get(name, name), None)

            def f(x):
                if isinstance(x, Expr):
                    return Expr(uname, (x,))
                if real is None:
                    raise UnsupportedOnDevice("no primitive named %s" % name)
                return real(x)
            return f
        if name in ("max", "min", "pow", "eq", "lt", "le", "gt", "ge"):
            real = getattr(rules, {"max": "max_", "min": "min_", "pow": "pow_"}.get(name, name))

            def g(a, b):
                if isinstance(a, Expr) or isinstance(b, Expr):
                    return Expr(name, (Expr.wrap(a), Expr.wrap(b)))
                return real(a, b)
            return g
        raise AttributeError(name)


E = _Namespace()


class _Program:
    def __init__(self, fn, nin):
        self.fn, self.nin = fn, nin
        e = Expr.wrap(fn(*[Expr("in", val=k) for k in range(nin)]))
        self.code, self.consts = compile_expr(e)
        self.c_code = (C.c_int32 * len(self.code))(*self.code)
        self.c_consts = (C.c_double * max(len(self.consts), 1))(*(self.consts or [0.0]))

    def run(self, operands):
        """operands: DArrays of one common shape, 0-dim DArrays, or Numbers.  Returns None if not applicable."""
        arrs = [o for o in operands if isinstance(o, DArray) and o.shape != ()]
        if not arrs:
            return None
        shape, dtype = arrs[0].shape, arrs[0].dtype
        if any(a.shape != shape or a.dtype != dtype for a in arrs):
            return None                      # needs real broadcasting: the caller falls back to the composite
        n = len(operands)
        ptrs, modes, imms = (C.c_void_p * n)(), (C.c_int32 * n)(), (C.c_double * n)()
        keep = []
        for k, o in enumerate(operands):
            if isinstance(o, DArray):
                o = D.materialize(o)
                keep.append(o)
                ptrs[k], modes[k], imms[k] = o.ptr, (0 if o.shape != () else 1), 0.0
            elif isnum(o):
                ptrs[k], modes[k], imms[k] = None, 2, float(o)
            else:
                return None
        out = DArray.empty(shape, dtype)
        check(lib.ag_ewise_vm(out.dt, self.c_code, len(self.code), self.c_consts, len(self.consts), n, ptrs, modes,
                              imms, D.vptr(out), out.size))
        return out


def primitive_expr(name, fwd, *grads):
    """Define an elementwise primitive from expressions: `fwd(x1,...,xn)` and one gradient expression
    `grad_i(dy, y, x1,...,xn)` per argument (None = no gradient).  n + 2 <= 4 operands."""
    from .rules import Primitive, unbroadcast
    import inspect
    nx = len(inspect.signature(fwd).parameters)
    if nx + 2 > 4:
        raise UnsupportedOnDevice("primitive_expr supports at most 2 arguments (dy, y, x1, x2)")
    fprog = _Program(fwd, nx)
    gprogs = [None if g is None else _Program(g, nx + 2) for g in grads]

    def fwd_raw(*xs):
        if all(isnum(x) for x in xs):
            return _host_eval(fprog, xs)
        r = fprog.run(list(xs))
        if r is None:
            return fwd(*xs)                   # broadcasting operands: composite over the ordinary primitives
        return r

    def make_back(i):
        if i >= len(grads) or grads[i] is None:
            return None

        def rule(dy, y, *xs):
            if not recording():
                ops = [value(dy), value(y)] + [value(x) for x in xs]
                if isnum(ops[0]) and isinstance(ops[1], DArray):
                    pass                       # scalar dy broadcasts as an immediate
                r = gprogs[i].run(ops)
                if r is not None:
                    return unbroadcast(xs[i], r)
            return unbroadcast(xs[i], grads[i](dy, y, *xs))
        return rule
    return Primitive(name, fwd_raw, tuple(make_back(i) for i in range(nx)))


def _host_eval(prog, xs):
    """Evaluate a compiled program on host Numbers (scalar Params never touch the device)."""
    st = []
    for c in prog.code:
        op, arg = c & 0xff, c >> 8
        if op == 0:
            st.append(float(xs[arg]))
        elif op == 1:
            st.append(prog.consts[arg])
        elif op == _NEG:
            st.append(-st.pop())
        elif op >= _UNARY0:
            uname = [k for k, v in D.U.items() if v == op - _UNARY0][0]
            st.append(D.HOST_UNARY[uname](st.pop()))
        else:
            b, a = st.pop(), st.pop()
            bname = [k for k, v in _BIN.items() if v == op][0]
            st.append(D.HOST_BINARY[bname](a, b))
    return st[0]
