"""Tape runtime: the host-side mirror of AutoGrad.jl's src/core.jl for device arrays.

Same surface and semantics as the reference (file:line under /root/reference):
  Param / Result / Node / Tape ......... src/core.jl:3-41
  recording(), the tape stack .......... src/core.jl:60-61
  forw (unbox, call, record) ........... src/core.jl:64-79
  Result()/record! ...................... src/core.jl:110-129
  differentiate (+ backward sweep) ...... src/core.jl:134-171
  findresult ............................ src/core.jl:187-197
  @diff / value / grad / gradloss ....... src/core.jl:201-232
  gcnode hook ........................... src/core.jl:235-237
The host language is Python here because no Julia toolchain exists in this image;
julia/AutoGradB200.jl holds the equivalent Julia binding.  The arrays inside Param/Result are
DArray; every array operation the sweep triggers is a kernel in libagb200.so.
"""
import os
import time
import weakref

from .darray import DArray, isnum

# `@timer` (src/AutoGrad.jl:7-12): with AUTOGRAD_TIMER set, every forward call, every `back` call and every
# accumulation is labelled with its tape position exactly like the reference (ftimer / btimer / stimer,
# src/core.jl:177-182): "[>ti]f", "[ti]f[i]", "[ti]addto![i]".  Each label opens an NVTX range around the kernels it
# launches (ag_range_push / ag_range_pop) and accumulates host wall time (after a stream sync, like the
# reference's `@gs`) and launch counts, reported by timer_report().
TIMER = [bool(os.environ.get("AUTOGRAD_TIMER"))]
_timings = {}


def set_timer(on):
    prev = TIMER[0]
    TIMER[0] = bool(on)
    return prev


class _timed:
    __slots__ = ("label", "t0", "l0")

    def __init__(self, label):
        self.label = label

    def __enter__(self):
        from . import darray as D
        D.check(D.lib.ag_range_push(self.label.encode()))
        self.l0 = D.launch_count()
        self.t0 = time.perf_counter()

    def __exit__(self, *exc):
        from . import darray as D
        D.sync()                                     # `@gs`
        dt = time.perf_counter() - self.t0
        D.check(D.lib.ag_range_pop())
        e = _timings.setdefault(self.label, [0, 0.0, 0])
        e[0] += 1
        e[1] += dt
        e[2] += D.launch_count() - self.l0
        return False


def timer_report(reset=True, top=40):
    """[(label, calls, seconds, launches)] sorted by time (AutoGrad.to of the reference)."""
    rows = sorted(((k, v[0], v[1], v[2]) for k, v in _timings.items()), key=lambda r: -r[2])[:top]
    if reset:
        _timings.clear()
    return rows


class Value:
    __array_priority__ = 3000

    def __add__(self, o): return _r().add(self, o)
    def __radd__(self, o): return _r().add(o, self)
    def __sub__(self, o): return _r().sub(self, o)
    def __rsub__(self, o): return _r().sub(o, self)
    def __mul__(self, o): return _r().mul(self, o)
    def __rmul__(self, o): return _r().mul(o, self)
    def __truediv__(self, o): return _r().div(self, o)
    def __rtruediv__(self, o): return _r().div(o, self)
    def __pow__(self, o): return _r().pow_(self, o)
    def __matmul__(self, o): return _r().matmul(self, o)
    def __rmatmul__(self, o): return _r().matmul(o, self)
    def __neg__(self): return _r().neg(self)
    def __getitem__(self, i): return _r().getindex(self, *(i if isinstance(i, tuple) else (i,)))

    # iteration in terms of getindex, which does the tracking (src/iterate.jl:6-17): `for e in v`, `a, b = v`
    def __len__(self): return _r().length(self)

    def __iter__(self):
        if isinstance(self.value, (int, float)):      # iterate(::Value{<:Number}) yields the Value itself once
            yield self
            return
        for i in range(_r().length(self)):
            yield self[i]

    @property
    def T(self): return _r().adjoint(self)


def _r():
    from . import rules
    return rules


class Tracked(Value):
    pass


class Param(Tracked):
    """Param(x): mark x as differentiable (src/core.jl:7-11,201)."""

    def __init__(self, value, opt=None):
        if isinstance(value, Value):
            raise TypeError("Param cannot take %s as arg." % type(value).__name__)
        self.value = value
        self.opt = opt


class Result(Tracked):
    __slots__ = ("value", "func", "args", "kwargs", "_soft")

    def __init__(self, value, func, args, kwargs):
        self.value, self.func, self.args, self.kwargs = value, func, args, kwargs
        self._soft = False


class Node:
    """src/core.jl:29-34.  `children` holds WEAK references (see `node_children`): parents <-> children would be a
    reference cycle, and a dropped Tape must release its device arrays at once (reference counting), not whenever
    the cyclic collector runs -- pending deferred operations of a dead tape would otherwise still be evaluated."""
    __slots__ = ("Value", "parents", "children", "outgrad", "__weakref__")

    def __init__(self, v):
        self.Value, self.parents, self.children, self.outgrad = v, [], [], None


def node_children(n):
    """The live children of a node (src/core.jl:32)."""
    return [c for c in (r() for r in n.children) if c is not None]


class Tape:
    def __init__(self):
        self.dict = {}    # id(Tracked) -> Node  (IdDict{Tracked,Node})
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


def forw(f, *args, **kwargs):
    """src/core.jl:64-79: strip the boxes, run the primitive on raw values, record if needed."""
    raw = tuple(a.value if isinstance(a, Value) else a for a in args)
    if TIMER[0]:
        with _timed(("[>%d]%s" % (1 + len(_tapes[-1].list), f.name)) if _tapes else f.name):      # ftimer
            v = f.fwd(*raw, **kwargs)
    else:
        v = f.fwd(*raw, **kwargs)
    if _tapes:
        for a in args:
            if isinstance(a, Tracked):
                return _record(v, f, args, kwargs)
    return v


def _node_of(tape, t):
    n = tape.dict.get(id(t))
    if n is None:
        n = Node(t)
        tape.dict[id(t)] = n
        tape.list.append(n)
    return n


def _record(v, f, args, kwargs):
    if isinstance(v, Tracked):       # reference issue #106
        return v
    result = Result(v, f, args, kwargs)
    _mark_soft(result, v, f, args)
    for tape in _tapes:
        node = Node(result)
        node.parents = [None] * len(args)
        for i, a in enumerate(args):
            if isinstance(a, Tracked):
                p = _node_of(tape, a)
                node.parents[i] = p
                p.children.append(weakref.ref(node))
        tape.dict[id(result)] = node
        tape.list.append(node)
    return result


def _mark_soft(result, v, f, args):
    """Deferred evaluation (lazy.py) stores every pending value that something still refers to.  A Result refers to
    its value, but many values on a tape are never read again: the gradient rules of `*` read its arguments, not its
    output (src/base.jl:26); those of `+` / `-` read neither (src/base.jl:21,23: only sizes, for unbroadcast).  A Result
    of such a primitive holds a pending value SOFTLY (it is not stored just because the tape points at it; if somebody
    reads it after all, it is computed then), and a consumer whose rule does read an argument turns that argument's
    hold into a hard one.  This is what lets `w*x .+ b` run as one launch without writing `w*x`."""
    if getattr(v, "kind", None) is not None and getattr(f, "y_unused", False):
        v.soft += 1
        result._soft = True
    unused = getattr(f, "args_unused", False)
    if unused is True:
        return
    for i, a in enumerate(args):
        if isinstance(a, Result) and a._soft and not (unused and i < len(unused) and unused[i]):
            a._soft = False
            if getattr(a.value, "kind", None) is not None:
                a.value.soft -= 1


gcnode = [lambda node, tape: None]


def set_gc_function(f):
    gcnode[0] = f


def release_node(node, tape):
    """A gcnode implementation for device arrays: drop a consumed Result's value and outgrad so
    their pool blocks are reusable by the rest of the sweep (stream-ordered, no sync)."""
    node.outgrad = None
    if isinstance(node.Value, Result):
        node.Value.value = None


def _is_scalar(v):
    return isnum(v) or (isinstance(v, DArray) and v.shape == ())


def differentiate(f, *x, **o):
    """src/core.jl:134-171."""
    from .rules import addto, back, full, identity
    from .sparse import Sparse
    tape = Tape()
    _tapes.append(tape)
    try:
        result = f(*x, **o)
        if isinstance(result, Param):
            result = identity(result)
    except BaseException:
        _tapes.pop()
        raise
    top = _tapes.pop()
    assert top is tape, "Tape stack error"
    if not isinstance(result, Result):
        return result
    assert _is_scalar(result.value), "Only scalar valued functions supported."
    resultnode = _findresult(tape, result)
    resultnode.outgrad = 1.0     # one(result.value); narrowed to the eltype on device
    lst = tape.list
    for ti in range(len(lst) - 1, -1, -1):
        n = lst[ti]
        if n.outgrad is None:
            continue
        r = n.Value
        for i, p in enumerate(n.parents):
            if p is None:
                continue
            if isinstance(n.outgrad, Sparse):
                n.outgrad = full(n.outgrad)
            if TIMER[0]:
                with _timed("[%d]%s[%d]" % (ti + 1, r.func.name, i + 1)):                             # btimer
                    g = back(r.func, i, n.outgrad, r, *r.args, **r.kwargs)
                with _timed("[%d]addto![%d]" % (ti + 1, i + 1)):                                      # stimer
                    p.outgrad = addto(p.outgrad, g)
                continue
            g = back(r.func, i, n.outgrad, r, *r.args, **r.kwargs)
            p.outgrad = addto(p.outgrad, g)
        if not _tapes and isinstance(r, Result) and n is not resultnode:
            gcnode[0](n, tape)
    if not _tapes:
        # latest first: a softly held sum that refers to a softly held product releases it
        soft = [n.Value.value for n in reversed(lst) if isinstance(n.Value, Result) and n.Value._soft
                and getattr(n.Value.value, "kind", None) is not None]
        n = r = p = None
        if soft:
            from . import lazy
            lazy.discard_soft(soft)
    return tape


def _findresult(tape, result):
    lst = tape.list
    if lst[-1].Value is result:
        return lst[-1]
    for i in range(len(lst) - 2, -1, -1):
        if lst[i].Value is result:
            del lst[i + 1:]
            break
    assert lst[-1].Value is result, "result not found on tape"
    return lst[-1]


def diff(thunk):
    """`@diff expr` (src/core.jl:205)."""
    return differentiate(thunk)


def grad(a, b=0, loss=False):
    """grad(tape, x) (src/core.jl:215-216) / grad(fun, argnum) (src/core.jl:219-230, argnum 0-based)."""
    from .rules import full
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
    """params(::Tape) (src/params.jl:6); containers are walked recursively (src/params.jl:3,8-75)."""
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
        elif hasattr(x, "__dict__") and not isinstance(x, (type, DArray)):
            for y in vars(x).values():
                walk(y)
    walk(t)
    return out
