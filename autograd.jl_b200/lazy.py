"""Deferred elementwise evaluation: horizontal fusion of the tape's elementwise chains.

The reference evaluates every dotted call on a tracked value as its own Base broadcast pass
(src/core.jl:68-70 materialises each `Broadcasted`), every gradient rule of src/math.jl:9-74 and
src/base.jl:20-49,73-74 as another pass, and the dense `addto!` (src/addto.jl:23) as one more.  The tape
semantics do not need that: a value only has to exist when a non-elementwise consumer (matmul, a reduction,
a gather, a host read) or an in-place update asks for it.

Here the raw elementwise operations of darray.py return a `LazyArray`: a DArray whose storage does not exist
yet and which remembers the operation that defines it.  Touching `.ptr` / `.buf` (which every kernel wrapper
does) forces it: the whole pending expression DAG below it is compiled to the bytecode of ag_ewise_fused and
evaluated by ONE kernel.  Leaves are read once; the root and every intermediate that something else still
refers to (a Result on the tape, a Node's outgrad, a user variable -- found by reference count) is written
once; temporaries of the gradient expressions are never stored.  Every instruction of the fused kernel
evaluates the same expression as the single-rule kernel, so results are bit-identical to unfused execution
(AGB_FUSION=0 or `set_fusion(False)` runs the unfused path).

Safety rules: (1) shapes and eltypes are checked when the node is created, so DimensionMismatch surfaces at
the call like in the reference; (2) anything that mutates device memory in place (axpy!, lmul!, scatter-add,
host copies into an existing array, collectives, graph capture/launch, sync) first evaluates everything
pending (`flush`), because a pending node reads its leaves later than the program order says; (3) a DAG whose
operand layouts the fused kernel cannot express is evaluated node by node with the ordinary kernels.
"""
import ctypes as C
import itertools
import os
import sys
import weakref

import numpy as np

from . import darray as D
from ._lib import check, DimensionMismatch, UnsupportedOnDevice

_RAW_PTR = D.DArray.__dict__["ptr"]
_RAW_BUF = D.DArray.__dict__["buf"]

MAX_NODES = 24          # pending nodes per DAG before the largest argument is evaluated early
BIG = 1 << 20           # elements from which a chain WITHOUT a specialised kernel is not handed to the interpreter
MAXCODE, MAXCONST, MAXIN, MAXOUT, SLOTS = 96, 16, 12, 12, 8
DYN_MAXIN, DYN_MAXOUT = 8, 6     # what the general interpreter takes; larger programs need a specialised kernel
LD_IN, LD_CONST, LD_TMP, ST_TMP, UNARY, BIN, BIN_IN, BIN_CONST, UGRAD, TERN, STORE, REDUCE = range(12)
# column programs (ag_ewise_colprog): segments over the elements (rows x cols) and over the columns (1 x cols)
CP_PHASE, CP_CRED, CP_LD_COL, CP_ST_COL, CP_CFIN = range(12, 17)
COL_SLOTS = 8           # per-column values a column program can hold
COL_MAXROWS = 1024      # longest column whose sum stays inside a chain
M_FULL, M_DEVSCALAR, M_IMM, M_COLVEC, M_ROWVEC = range(5)

enabled = [os.environ.get("AGB_FUSION", "1") != "0"]
reduce_fusion = [os.environ.get("AGB_FUSE_REDUCE", "1") != "0"]
# Column sums inside a chain (softmax normaliser and its gradient as ONE ag_ewise_colprog launch).  Off by default:
# measured on the B200 (round 2) the MNIST step goes from 19 to 16 launches but from 185.6 to 196-201 us, and config 4
# from 3624 to 3324 launches at the same 14.6 ms -- a column program gives a lane one row of one column (10 of 16
# lanes busy at m = 10, two dependent memory round trips per warp), where the separate chain kernels stream 128-bit
# vectors with every load in flight.  AGB_FUSE_COLUMNS=1 (or set_col_fusion(True)) turns it on.
col_fusion = [os.environ.get("AGB_FUSE_COLUMNS", "0") == "1"]
batch_reductions = [os.environ.get("AGB_BATCH_REDUCTIONS", "1") != "0"]    # defer stand-alone reductions and batch them
static_only = [os.environ.get("AGB_FUSE_BIG_INTERPRETED", "0") == "0"]   # see BIG   # sums of a pending chain in the chain's own launch
stats = {"fused_launches": 0, "fused_nodes": 0, "fallback_nodes": 0, "splits": 0}
trace = None            # set to a list to record (n elements, full-array inputs, outputs, program) of every fused launch
_counter = itertools.count()
_pending = weakref.WeakValueDictionary()


def set_fusion(on):
    """Enable / disable deferred evaluation; returns the previous setting.  Disabling evaluates what is pending."""
    prev = enabled[0]
    if not on:
        flush()
    enabled[0] = bool(on)
    return prev


def set_col_fusion(on):
    """Enable / disable column programs (see col_fusion); returns the previous setting.  Evaluates what is pending."""
    prev = col_fusion[0]
    flush()
    col_fusion[0] = bool(on)
    return prev


class LazyArray(D.DArray):
    """A DArray defined by a pending elementwise operation.  kind is None once the storage exists."""
    __slots__ = ("kind", "op", "args", "nnodes", "seq", "soft")

    def __init__(self, shape, dtype, kind, op, args, nnodes):
        _RAW_BUF.__set__(self, None)
        _RAW_PTR.__set__(self, 0)
        self.shape, self.dtype, self.tr = tuple(int(s) for s in shape), np.dtype(dtype), False
        self.kind, self.op, self.args, self.nnodes = kind, op, args, nnodes
        self.soft = 0           # holders that do not need the value stored (see core._record): a tape Result whose
                                # value no gradient rule reads, e.g. the product under `w*x .+ b`
        self.seq = next(_counter)
        _pending[self.seq] = self

    @property
    def ptr(self):
        if self.kind is not None:
            force(self)
        return _RAW_PTR.__get__(self, D.DArray)

    @property
    def buf(self):
        if self.kind is not None:
            force(self)
        return _RAW_BUF.__get__(self, D.DArray)

    def _set_storage(self, arr):
        _RAW_BUF.__set__(self, _RAW_BUF.__get__(arr, D.DArray))
        _RAW_PTR.__set__(self, _RAW_PTR.__get__(arr, D.DArray))
        self.kind = self.op = self.args = None
        self.nnodes = 0
        _pending.pop(self.seq, None)

    def __repr__(self):
        if self.kind is not None:
            return "LazyArray{%s}%s<%s %s>" % (self.dtype.name, self.shape, self.kind, self.op)
        return D.DArray.__repr__(self)


def kernel_stats():
    """(launches that ran an ahead-of-time specialised program, launches on the general interpreter)."""
    a = (C.c_int64 * 2)()
    check(D.lib.ag_ewise_fused_stats(a))
    return int(a[0]), int(a[1])


def set_static(on):
    """Use (default) / ignore the ahead-of-time specialised kernels; what the host remembers about them is dropped."""
    check(D.lib.ag_ewise_fused_set_static(1 if on else 0))
    _static_cache.clear()


def is_pending(x):
    return isinstance(x, LazyArray) and x.kind is not None


def pending_count():
    return sum(1 for v in list(_pending.values()) if v.kind is not None)


# ---- node construction ----------------------------------------------------------------------------------

def _norm(v):
    """Operand of a node: host Number -> float, lazy adjoint -> dense, everything else as is."""
    if v is None:
        return None
    if D.isnum(v):
        return float(v)
    if isinstance(v, D.DArray):
        if v.tr:
            return D.materialize(v)
        return v
    raise UnsupportedOnDevice("elementwise operand of type %s" % type(v).__name__)


def _make(kind, op, args, shape, dtype):
    shape = tuple(int(s) for s in shape)
    args = tuple(_norm(a) for a in args)
    dts = {a.dtype for a in args if isinstance(a, D.DArray)}
    dts.add(np.dtype(dtype))
    if len(dts) != 1:
        raise UnsupportedOnDevice("mixed or missing device eltypes %s" % (dts,))
    nn = 1
    for a in args:
        if is_pending(a):
            if a.shape != shape and not _col_operand(a, shape):
                force(a)                      # a broadcast operand must be concrete (it becomes a leaf)
            else:
                nn += a.nnodes
    while nn > MAX_NODES:
        big = max((a for a in args if is_pending(a)), key=lambda a: a.nnodes)
        nn -= big.nnodes
        force(big)
    return LazyArray(shape, dtype, kind, op, args, nn)


def _col_operand(a, shape):
    """A pending (1, n) value used by an operation over (m, n): it stays pending, and the whole DAG runs as a column
    program (ag_ewise_colprog) in which the value is held per column."""
    return (col_fusion[0] and len(shape) == 2 and a.shape == (1, shape[1]) and 2 <= shape[0] <= COL_MAXROWS
            and shape[1] >= 2 and shape[0] * shape[1] < BIG and a.kind != "mm")


def unary_fwd(op, x):
    return _make("u", op, (x,), x.shape, x.dtype)


def unary_bwd(op, dy, x, y, shape, dtype):
    """dx = dy .* g(x, y); dy a Number, a size-1 device array or an array of `shape`."""
    return _make("ug", op, (dy, x, y), shape, dtype)


def binary_fwd(op, a, b, out_shape, dtype):
    return _make("b", op, (a, b), out_shape, dtype)


def binary_bwd(op, argno, dy, x1, x2, y, out_shape, dtype):
    return _make("bb", (op, argno), (dy, x1, x2, y), out_shape, dtype)


def fill(shape, dtype, v):
    return _make("fill", None, (float(v),), shape, dtype)


def bcast_to(a, shape):
    return _make("id", None, (a,), shape, a.dtype)


def matmul(A, ta, B, tb, m, n, k, lda, ldb, out_shape, dtype):
    """A pending product op(A)*op(B) (src/base.jl:26 forward).  Deferred so that the `.+ b` and the activation
    that follow it in every model of the reference (examples/mnist.jl:32, examples/charlm.jl:74-77) can run in the
    product's epilogue (ag_gemm_epilogue) instead of as separate passes over the output; a product nobody builds
    on is evaluated by ag_gemm when it is touched.  A and B are concrete storage (the caller forced them)."""
    return LazyArray(out_shape, dtype, "mm", (int(ta), int(tb), int(m), int(n), int(k), int(lda), int(ldb)), (A, B), 1)


# ---- evaluation -------------------------------------------------------------------------------------------

def reduce(x, inner, red, outer, mapname):
    """sum(map, x) over the middle extent of x viewed as inner x red x outer, evaluated in the launch that evaluates
    the pending chain x (src/unbroadcast.jl:17-24 and src/base.jl:245-247 without a second pass over x).  Returns a
    flat DArray of inner*outer values, or None when the sum is not fused (the caller reduces the stored array)."""
    if not (enabled[0] and reduce_fusion[0] and is_pending(x)) or red < 2:
        return None
    if col_fusion[0] and inner == 1 and outer >= 2 and red <= COL_MAXROWS and mapname == "id" and x.kind != "mm" \
            and x.shape == (red, outer) and red * outer < BIG:
        # sum(x, dims=1) of a pending (m, n) chain, m small: the (1, n) result stays pending so that what follows
        # (log, the broadcast back over the rows, the gradient of both) runs in the SAME launch as the chain
        # (examples/mnist.jl:40 `ypred .- log.(sum(exp.(ypred),dims=1))`, src/unbroadcast.jl:17-24)
        nn = 1 + x.nnodes
        if nn > MAX_NODES:
            force(x)
            if x.kind is None:
                return None
        return LazyArray((1, outer), x.dtype, "cr", mapname, (x,), 1 + x.nnodes)
    if x.kind == "mm" or x.size != inner * red * outer:
        return None
    if col_fusion[0] and _mixed_order(_topo(x)):
        if not (inner == 1 and outer == 1):
            return None                     # other layouts: the caller evaluates x (as a column program) and sums it
    else:
        f32 = x.dtype == np.float32
        E, N = (2048, 4) if f32 else (1024, 2)
        if x.size % N:
            return None
        if not ((inner == 1 and (outer == 1 or red <= E)) or (outer == 1 and inner * 16 <= E)):
            return None
    order = _topo(x)
    while _resolve_products(order, x):
        if x.kind is None:
            return None
        order = _topo(x)
    out = D.DArray.empty((inner * outer,), x.dtype)
    try:
        if col_fusion[0] and _mixed_order(order):
            _run_colprog(order, x, (mapname, out))
        else:
            _run_fused(order, x, (inner, red, outer, mapname, out))
    except (_TooBig, _NotFusable, _NotStatic, UnsupportedOnDevice):
        stats["unfused_reductions"] = stats.get("unfused_reductions", 0) + 1
        return None
    return out


def reduce_deferred(x, inner, red, outer, mapname, out_shape):
    """sum(map, x; dims) of a CONCRETE x as a pending value (kind "rs"): it is evaluated when something touches it, and
    then together with every other pending reduction of the same layout in ONE launch (ag_sum_dim_batch) -- the bias
    gradients `unbroadcast(b, dy)` of the four LSTM gates over many timesteps wait in the accumulation chain of
    `addto!` until that chain is evaluated.  Returns None when reductions are not deferred."""
    if not (enabled[0] and reduce_fusion[0] and batch_reductions[0]) or is_pending(x) or x.tr:
        return None
    return LazyArray(out_shape, x.dtype, "rs", (int(inner), int(red), int(outer), mapname), (x,), 1)


def _force_reductions(root):
    """Evaluate the pending reduction `root` and every other pending reduction of its layout (latest first)."""
    batch = [root]
    for k in sorted(_pending.keys(), reverse=True):
        n = _pending.get(k)
        if n is not None and n is not root and n.kind == "rs" and n.op == root.op and n.dtype == root.dtype:
            batch.append(n)
        n = None
    inner, red, outer, mapname = root.op
    outs = [D.DArray.empty(n.shape, n.dtype) for n in batch]
    cnt = len(batch)
    xs = (C.c_void_p * cnt)(*[n.args[0].ptr for n in batch])
    os_ = (C.c_void_p * cnt)(*[o.ptr for o in outs])
    check(D.lib.ag_sum_dim_batch(D.R[mapname], D._DT[np.dtype(root.dtype)], cnt, xs, inner, red, outer, os_))
    stats["batched_reductions"] = stats.get("batched_reductions", 0) + cnt
    stats["reduction_batches"] = stats.get("reduction_batches", 0) + 1
    for n, o in zip(batch, outs):
        n._set_storage(o)


class _TooBig(Exception):
    """The DAG needs more inputs / outputs / instructions / slots than one launch has: evaluate arguments first."""


class _NotStatic(Exception):
    """A large chain has no specialised kernel: evaluate its arguments as chains of their own, single operations
    with the single kernels (the interpreter is issue-bound; it pays off only where launches dominate)."""


_static_cache = {}


def _is_static(code):
    key = tuple(code)
    r = _static_cache.get(key)
    if r is None:
        flag = C.c_int32(0)
        c_code = (C.c_int32 * len(code))(*code)
        check(D.lib.ag_ewise_fused_query(c_code, len(code), C.byref(flag)))
        r = _static_cache[key] = bool(flag.value)
        if len(_static_cache) > 4096:
            _static_cache.clear()
    return r


class _NotFusable(Exception):
    """Some operand layout is not a full array, row vector, column vector or scalar of the iteration space."""


def _visit(n, order, seen):
    seen.add(id(n))
    for a in n.args:
        if is_pending(a) and id(a) not in seen:
            _visit(a, order, seen)
    order.append(n)


def _topo(root):
    """Pending nodes below `root`, arguments first.  (Module-level recursion: a self-referencing closure would be a
    reference cycle keeping `order` -- and so every node of the DAG -- alive until the cyclic collector runs.)"""
    order = []
    _visit(root, order, set())
    return order


def flush():
    """Evaluate everything pending (latest first, so that whole chains go out as one kernel)."""
    if not _pending:
        return
    for k in sorted(_pending.keys(), reverse=True):
        n = _pending.get(k)
        if n is not None and n.kind is not None:
            force(n)
        n = None


def discard_soft(nodes):
    """End of a backward sweep (core.differentiate): pending values that only the finished tape's Results hold, and
    that no gradient rule reads (soft holds), will never be needed by the sweep.  They are dropped instead of being
    computed by the next flush (an in-place `axpy!` on a Param flushes, and would otherwise run the product under
    `w*x .+ b` once more just to store a value nobody reads).  Reading such a value afterwards raises: its inputs may
    have been updated in place since, so recomputing it could silently give a different array than the reference's
    stored one.  Keep a reference to `value(x)` during the forward pass to have it stored."""
    for i in range(len(nodes)):
        if nodes[i].kind is not None and nodes[i].kind != "lost" and nodes[i].soft > 0 and _external(nodes[i], 0) <= 0:
            v = nodes[i]
            v.kind, v.op, v.args, v.nnodes = "lost", None, (), 0
            _pending.pop(v.seq, None)
            v = None
            stats["discarded"] = stats.get("discarded", 0) + 1


def force(root):
    """Evaluate `root` (and every pending node below it that something else refers to)."""
    if root.kind == "lost":
        raise RuntimeError("this intermediate value was not stored: no gradient rule reads it and nothing else held it "
                           "when the backward sweep finished (lazy.discard_soft); hold value(x) to keep it")
    while root.kind is not None:
        if root.kind == "mm":
            root._set_storage(D._matmul_eager(root.args[0], root.args[1], root.op, root.shape, root.dtype))
            return
        if root.kind == "rs":
            _force_reductions(root)
            return
        order = _topo(root)
        if _resolve_products(order, root):
            order = None
            continue                        # some nodes got their storage from a fused product: walk the rest again
        rs = None
        for i in range(len(order)):
            if order[i].kind == "rs":
                rs = order[i]
                break
        if rs is not None:                  # deferred reductions below the root: evaluate them (batched), walk again
            order = None
            _force_reductions(rs)
            rs = None
            continue
        action, ext = None, None
        if col_fusion[0] and _mixed_order(order):
            ok = False
            try:
                _run_colprog(order, root)
                ok = True
            except (_TooBig, _NotFusable, UnsupportedOnDevice):
                pass
            if ok:
                return
            stats["col_fallbacks"] = stats.get("col_fallbacks", 0) + 1
            _split_columns(order)           # evaluate a column sum / column operand on its own, then walk again
            order = None
            continue
        try:
            _run_fused(order, root)
            return
        except _TooBig as why:
            action, ext = "split:" + (why.args[0] if why.args else "?"), (why.args[1] if len(why.args) > 1 else set())
        except _NotStatic:
            action = "unspecialised"
        except _NotFusable:
            action = "eager"
        # The handlers only take notes: while an exception is being handled its traceback keeps the frames of
        # _run_fused alive, and with them lists that refer to every node of the DAG -- every node would look
        # externally referenced to the nested force() calls below and be written out.
        if action == "eager":
            _run_eager(order)
            return
        if action == "unspecialised":
            stats["big_unspecialised"] = stats.get("big_unspecialised", 0) + 1
            order = None
            sub = [a for a in root.args if is_pending(a)]
            if not sub:
                _run_eager([root])
                return
            for a in sub:
                force(a)
            a = sub = None
            continue
        stats["splits"] += 1
        key = "split_" + action[6:]
        stats[key] = stats.get(key, 0) + 1
        sub = _pick_split(order, root, ext)
        order = ext = None
        if sub is None:                     # a single operation always fits; defensive
            _run_eager([root])
            return
        force(sub)
        sub = None


def _mixed_order(order):
    """Does the DAG hold a column sum, or an operation whose pending operand has another shape (column space)?"""
    for n in order:
        if n.kind == "cr":
            return True
        for a in n.args:
            if is_pending(a) and a.shape != n.shape:
                return True
    return False


def _split_columns(order):
    """Fallback of a DAG that does not fit a column program: its first column sum is evaluated as before (fused into
    the chain below it when possible, else by ag_sum_dim); without column sums, the pending column-space operands."""
    for i in range(len(order)):
        if order[i].kind == "cr":
            c = order[i]
            x = c.args[0]
            red, outer = x.shape
            out = None
            if is_pending(x):
                sub = _topo(x)
                if not _mixed_order(sub) and not any(n.kind == "mm" for n in sub):
                    f32 = x.dtype == np.float32
                    E, N = (2048, 4) if f32 else (1024, 2)
                    if x.size % N == 0 and red <= E:
                        out = D.DArray.empty((outer,), x.dtype)
                        try:
                            _run_fused(sub, x, (1, red, outer, "id", out))
                        except (_TooBig, _NotFusable, _NotStatic, UnsupportedOnDevice):
                            out = None
                sub = None
            if out is None:
                out = D.DArray.empty((outer,), x.dtype)
                check(D.lib.ag_sum_dim(D.R["id"], D._DT[np.dtype(x.dtype)], C.c_void_p(x.ptr), 1, red, outer,
                                       C.c_void_p(out.ptr)))
            c._set_storage(D.DArray(out.buf, out.ptr, c.shape, out.dtype))
            return
    for i in range(len(order)):
        for a in order[i].args:
            if is_pending(a) and a.shape != order[i].shape:
                force(a)
                return
    raise RuntimeError("lazy: a mixed DAG without column operands")


def _pick_split(order, root, ext):
    """The largest sub-DAG (by pending nodes) that fits one launch: its root is evaluated first, then `root` is
    retried with that node as a leaf.  Sizes are tree estimates (a shared node counts once per use)."""
    leaves, nout, nn, best = {}, {}, {}, None
    for n in order:
        lv, no, k = set(), 0, 1
        for a in n.args:
            if is_pending(a):
                lv |= leaves[id(a)]
                no += nout[id(a)]
                k += nn[id(a)]
            elif isinstance(a, D.DArray):
                lv.add(id(a))
        leaves[id(n)], nn[id(n)] = lv, k
        nout[id(n)] = no + (1 if id(n) in ext else 0)
        if n is root:
            continue
        outs_if_root = no + 1                   # the sub-root is written whether or not something refers to it
        if len(lv) <= DYN_MAXIN and outs_if_root <= DYN_MAXOUT and 4 * k <= MAXCODE and k <= SLOTS * 2:
            if best is None or k > nn[id(best)]:
                best = n
    return best


def _run_eager(order):
    """Node-by-node evaluation with the ordinary kernels (general <= 4-d strided broadcasting)."""
    for n in order:
        if n.kind is None:
            continue
        k, a = n.kind, n.args
        if k == "u":
            r = D._unary_fwd_eager(n.op, a[0])
        elif k == "ug":
            r = D._unary_bwd_eager(n.op, a[0], a[1], a[2], shape=n.shape, dtype=n.dtype)
        elif k == "b":
            r = D._binary_fwd_eager(n.op, a[0], a[1])
        elif k == "bb":
            r = D._binary_bwd_eager(n.op[0], n.op[1], a[0], a[1], a[2], a[3])
        elif k == "fill":
            r = D._full_like_shape_eager(n.shape, n.dtype, a[0])
        elif k == "mm":
            r = D._matmul_eager(a[0], a[1], n.op, n.shape, n.dtype)
        elif k == "rs":
            _force_reductions(n)
            stats["fallback_nodes"] += 1
            continue
        elif k == "cr":
            r = D.DArray.empty(n.shape, n.dtype)
            check(D.lib.ag_sum_dim(D.R["id"], D._DT[np.dtype(n.dtype)], C.c_void_p(a[0].ptr), 1, a[0].shape[0], a[0].shape[1],
                                   C.c_void_p(r.ptr)))
        else:
            r = D._bcast_to_eager(a[0], n.shape)
        stats["fallback_nodes"] += 1
        n._set_storage(r)


# (op, argno) of a broadcasting binary rule -> ("bin", binary op, other operand) or ("tern", map id, p, q)
# with operands named dy/x1/x2/y: exactly the table of ag_binary_bwd (csrc/ewise.cu).
def _bb_form(op, argno):
    xi = "x1" if argno == 1 else "x2"
    if op == "add":
        return ("binc", "mul", 1.0)
    if op == "sub":
        return ("binc", "mul", 1.0 if argno == 1 else -1.0)
    if op == "mul":
        return ("bin", "mul", "x2" if argno == 1 else "x1")
    if op == "div":
        return ("bin", "div", "x2") if argno == 1 else ("tern", 0, "x1", "x2")
    if op in ("max", "min"):
        return ("tern", 1, "y", xi)
    if op == "pow":
        return ("tern", 2, "x1", "x2") if argno == 1 else ("tern", 3, "y", "x1")
    if op == "atan2":
        return ("tern", 4 if argno == 1 else 5, "x1", "x2")
    if op == "hypot":
        return ("tern", 6, xi, "y")
    raise UnsupportedOnDevice("binary op %s has no gradient (zerograd)" % op)


_REF_BASE = 0


def _external(n, uses):
    """Number of references to the pending node n from outside its DAG that need the value stored: everything that
    holds n except the DAG's own bookkeeping (the `order` list, the call's own references), the uses inside the DAG
    and the soft holders.  Always called as `_external(order[i], uses)`."""
    return sys.getrefcount(n) - _REF_BASE - uses - n.soft


def _calibrate_refs():
    """The bookkeeping references seen by `_external(lst[i], 0)` for an object held by nothing but the list: measured
    on a throwaway object instead of hard-coding CPython's calling convention (the count differs between interpreter
    versions and builds; 3.14 borrowed references, tracers).  Returns None if the measurement is not stable."""
    class _Probe:
        __slots__ = ("soft",)

        def __init__(self):
            self.soft = 0
    seen = set()
    for _ in range(3):
        lst = [_Probe(), _Probe()]
        for i in range(len(lst)):
            seen.add(_external(lst[i], 0))
    return seen.pop() if len(seen) == 1 else None


_REF_BASE = _calibrate_refs()
_REFS_OK = _REF_BASE is not None
if not _REFS_OK:          # cannot tell internal from external references: store every anchor (correct, just more writes)
    _REF_BASE = -(1 << 30)

_ACTS = {("b", "max"): "relu", ("u", "tanh"): "tanh", ("u", "sigm"): "sigm"}


def _resolve_products(order, root):
    """Evaluate the pending products (kind "mm") below `root`.  A product whose only use is `M .+ bias` with a column
    vector bias -- optionally followed by max.(0, .), tanh.(.) or sigm.(.) as the only use of the sum -- runs as ONE
    launch with the sum and the activation in its epilogue (ag_gemm_epilogue); only the values something outside
    still needs are written (the product itself and, when no gradient rule reads it, the pre-activation sum are
    never stored).  Any other product is evaluated by ag_gemm.  Returns True if anything was evaluated."""
    if not any(n.kind == "mm" for n in order):
        return False
    uses = {}
    for n in order:
        for a in n.args:
            if is_pending(a):
                uses[id(a)] = uses.get(id(a), 0) + 1
    n = a = None
    # outside references first, while nothing but `order` and the nodes' own argument tuples holds the nodes
    ext = [order[i] is root or _external(order[i], uses.get(id(order[i]), 0)) > 0 for i in range(len(order))]
    index = {id(order[i]): i for i in range(len(order))}
    users = {}
    for n in order:
        for a in n.args:
            if is_pending(a):
                users.setdefault(id(a), []).append(n)
    n = a = None
    jobs = []           # (product node, bias, act, z node or None, activation node or None, z_needed)
    for i in range(len(order)):
        m = order[i]
        if m.kind != "mm":
            continue
        us = users.get(id(m), [])
        z = us[0] if len(us) == 1 and uses.get(id(m), 0) == 1 else None
        bias = None
        if z is not None and z.kind == "b" and z.op == "add" and z.shape == m.shape and not ext[i]:
            other = z.args[1] if z.args[0] is m else z.args[0]
            if isinstance(other, D.DArray) and not is_pending(other) and not other.tr and other.dtype == m.dtype \
                    and other.shape in ((m.shape[0],), (m.shape[0], 1)) \
                    and D._gemm_epilogue_fused(m.args[0], m.args[1], m.op, m.dtype):
                bias = other
        if bias is None:
            jobs.append((m, None, "identity", None, None, False))
            stats["products_plain"] = stats.get("products_plain", 0) + 1
            continue
        z_needed = ext[index[id(z)]]
        zu = users.get(id(z), [])
        act, anode = "identity", None
        if len(zu) == 1 and uses.get(id(z), 0) == 1 and zu[0].shape == z.shape:
            c = zu[0]
            if c.kind == "u" and ("u", c.op) in _ACTS:
                act, anode = _ACTS[("u", c.op)], c
            elif c.kind == "b" and c.op == "max" and ((c.args[0] == 0.0 and c.args[1] is z) if isinstance(c.args[0], float)
                                                      else (isinstance(c.args[1], float) and c.args[1] == 0.0 and c.args[0] is z)):
                act, anode = "relu", c
        jobs.append((m, bias, act, z, anode, z_needed or anode is None))
        stats["products_fused"] = stats.get("products_fused", 0) + 1
    # All products of the DAG are independent of each other (operands exist): hand them over together, so that the
    # library runs those of one class as ONE launch (ag_gemm_group) -- the four gate products of an LSTM step, their
    # four dx products, the dW products of several timesteps.
    results = D._gemm_group([(m.args[0], m.args[1], m.op, m.shape, m.dtype, bias, act, want_z)
                             for m, bias, act, z, anode, want_z in jobs])
    if len(jobs) > 1:
        stats["product_groups"] = stats.get("product_groups", 0) + 1
    for (m, bias, act, z, anode, want_z), (zs, as_) in zip(jobs, results):
        if z is None:
            m._set_storage(zs)
        elif anode is not None:
            anode._set_storage(as_)
            if want_z:
                z._set_storage(zs)
        else:
            z._set_storage(zs)
        # m (and z when nothing needs it) stay pending: they are only held softly and would be recomputed if read
    return True


class _Codegen:
    """Expression DAG -> bytecode of ag_ewise_fused.  (A class rather than nested functions: mutually recursive
    closures are reference cycles, and everything they capture -- pending nodes included -- would only die when
    the cyclic collector runs.)"""

    def __init__(self, uses, outputs, ext_ids):
        self.uses, self.outputs, self.ext_ids = uses, outputs, ext_ids
        self.code, self.consts, self.leaves, self.leaf_idx = [], [], [], {}
        self.depth = self.maxdepth = 0
        self.tmp_slot, self.remaining, self.free_slots = {}, {}, list(range(SLOTS - 1, -1, -1))

    def emit(self, op, a=0, b=0):
        self.code.append(op | (a << 8) | (b << 16))

    def const_idx(self, v):
        for i, c in enumerate(self.consts):
            if c == v and np.signbit(c) == np.signbit(v):
                return i
        if len(self.consts) >= MAXCONST:
            raise _TooBig("constants", self.ext_ids)
        self.consts.append(v)
        return len(self.consts) - 1

    def leaf(self, x):
        i = self.leaf_idx.get(id(x))
        if i is None:
            if len(self.leaves) >= MAXIN:
                raise _TooBig("inputs", self.ext_ids)
            i = self.leaf_idx[id(x)] = len(self.leaves)
            self.leaves.append(x)
        return i

    def pushed(self, push):
        if push:
            self.depth += 1
            self.maxdepth = max(self.maxdepth, self.depth)

    @staticmethod
    def simple(x):
        return isinstance(x, float) or (isinstance(x, D.DArray) and not is_pending(x))

    def gen(self, x, push):
        if x is None:
            x = 0.0
        if isinstance(x, float):
            self.emit(LD_CONST, self.const_idx(x), 1 if push else 0)
            self.pushed(push)
        elif is_pending(x):
            if id(x) in self.tmp_slot:
                self.emit(LD_TMP, self.tmp_slot[id(x)], 1 if push else 0)
                self.pushed(push)
                self.remaining[id(x)] -= 1
                if self.remaining[id(x)] == 0:
                    self.free_slots.append(self.tmp_slot[id(x)])
            else:
                self.gen_node(x, push)
        else:
            self.emit(LD_IN, self.leaf(x), 1 if push else 0)
            self.pushed(push)

    def gen_bin(self, bop, a, b, push):
        bid = D.B[bop]
        if self.simple(b):
            self.gen(a, push)
            if isinstance(b, float):
                self.emit(BIN_CONST, bid, self.const_idx(b))
            else:
                self.emit(BIN_IN, bid, self.leaf(b))
        elif self.simple(a):
            self.gen(b, push)
            if isinstance(a, float):
                self.emit(BIN_CONST, bid, self.const_idx(a) | 0x80)
            else:
                self.emit(BIN_IN, bid, self.leaf(a) | 0x80)
        else:
            self.gen(a, push)
            self.gen(b, True)
            self.emit(BIN, bid)
            self.depth -= 1

    def gen_node(self, n, push):
        k, a = n.kind, n.args
        if k == "u":
            self.gen(a[0], push)
            self.emit(UNARY, D.U[n.op])
        elif k == "b":
            self.gen_bin(n.op, a[0], a[1], push)
        elif k == "ug":
            self.gen(a[0], push)
            self.gen(a[1], True)
            self.gen(a[2], True)
            self.emit(UGRAD, D.U[n.op])
            self.depth -= 2
        elif k == "bb":
            named = {"dy": a[0], "x1": a[1], "x2": a[2], "y": a[3]}
            form = _bb_form(*n.op)
            if form[0] == "binc":
                self.gen_bin(form[1], named["dy"], form[2], push)
            elif form[0] == "bin":
                self.gen_bin(form[1], named["dy"], named[form[2]], push)
            else:
                self.gen(named["dy"], push)
                self.gen(named[form[2]], True)
                self.gen(named[form[3]], True)
                self.emit(TERN, form[1])
                self.depth -= 2
        else:                                   # "fill" / "id": a constant or a broadcast operand as is
            self.gen(a[0], push)

    def run(self, anchors, reduce_map=None):
        code = self.code
        for n in anchors:
            self.depth = 0
            self.gen_node(n, False)
            if n in self.outputs:
                self.emit(STORE, self.outputs.index(n))
            u = self.uses.get(id(n), 0)
            if u > 0:
                if not self.free_slots:
                    raise _TooBig("slots", self.ext_ids)
                s = max(self.free_slots)
                self.free_slots.remove(s)
                self.tmp_slot[id(n)], self.remaining[id(n)] = s, u
                self.emit(ST_TMP, s)
        if reduce_map is not None:              # the root (last anchor) is on top: sum it, or abs2 / abs of it
            if reduce_map != "id":
                self.emit(UNARY, D.U[reduce_map])
            self.emit(REDUCE)
        # peephole: "ST_TMP s; LD_TMP s (overwrite)" with no later load of s -- the value is already on top
        k = 0
        while k + 1 < len(code):
            a_, b_ = code[k], code[k + 1]
            if (a_ & 0xff) == ST_TMP and b_ == (LD_TMP | (a_ & 0xff00)):
                slot = (a_ >> 8) & 0xff
                later = False
                for c in code[k + 2:]:
                    if (c & 0xff) == ST_TMP and ((c >> 8) & 0xff) == slot:
                        break
                    if (c & 0xff) == LD_TMP and ((c >> 8) & 0xff) == slot:
                        later = True
                        break
                if not later:
                    del code[k:k + 2]
                    continue
            k += 1
        used_slots = {(c >> 8) & 0xff for c in code if (c & 0xff) in (ST_TMP, LD_TMP)}
        lowest_tmp = min(used_slots) if used_slots else SLOTS
        if len(code) > MAXCODE or self.maxdepth > lowest_tmp:
            raise _TooBig("code" if len(code) > MAXCODE else "slots", self.ext_ids)


def _run_fused(order, root, reduce=None):
    S = root.shape
    n_elem = D._prod(S)
    dtype = root.dtype
    # ---- which nodes must exist afterwards / are used more than once
    uses = {}
    for n in order:
        for a in n.args:
            if is_pending(a):
                uses[id(a)] = uses.get(id(a), 0) + 1
    n = a = None
    anchors, outputs = [], []
    for idx in range(len(order)):
        n = order[idx]
        u = uses.get(id(n), 0)
        # references from outside the DAG that need the value (tape Results that read it, outgrads, user variables)
        n = None
        ext = order[idx] is root or _external(order[idx], u) > 0
        n = order[idx]
        if ext or u > 1:
            anchors.append(n)
            if ext:
                outputs.append(n)
    n = None
    ext_ids = {id(o) for o in outputs}
    if len(outputs) > MAXOUT:
        raise _TooBig("outputs", ext_ids)
    if n_elem == 0:
        for o in outputs:
            o._set_storage(D.DArray.empty(S, dtype))
        return

    cg = _Codegen(uses, outputs, ext_ids)
    cg.run(anchors, None if reduce is None else reduce[3])
    code, consts, leaves = cg.code, cg.consts, cg.leaves
    cg = None
    if (len(leaves) > DYN_MAXIN or len(outputs) > DYN_MAXOUT) and \
            (n_elem % (4 if dtype == np.float32 else 2) or not _is_static(code)):
        raise _TooBig("inputs" if len(leaves) > DYN_MAXIN else "outputs", ext_ids)
    if reduce is None and n_elem >= BIG and static_only[0] and len(order) >= 1:
        if n_elem % (4 if dtype == np.float32 else 2) or not _is_static(code):
            raise _NotStatic()

    # ---- operand layouts: full / scalar / row vector / column vector of a collapsed rows x cols space
    modes, imms, bc = [], [], []
    for x in leaves:
        if x.shape == S:
            modes.append(M_FULL)
        elif x.size == 1:
            modes.append(M_DEVSCALAR)
        else:
            modes.append(None)
            bc.append(x)
    rows, cols = n_elem, 1
    if bc:
        sl = [D._strides_for(x.shape, S) for x in bc] + [D._strides_for(S, S)]
        shp, sts = D.collapse(list(S), sl)
        if len(shp) != 2:
            raise _NotFusable()
        rows, cols = shp
        it = iter(sts)
        for i, m in enumerate(modes):
            if m is None:
                st = next(it)
                if st == [1, 0]:
                    modes[i] = M_COLVEC
                elif st == [0, 1]:
                    modes[i] = M_ROWVEC
                else:
                    raise _NotFusable()

    outs = [D.DArray.empty(S, dtype) for _ in outputs]
    nin, nout = len(leaves), len(outs)
    c_code = (C.c_int32 * len(code))(*code)
    c_consts = (C.c_double * max(len(consts), 1))(*(consts or [0.0]))
    c_in = (C.c_void_p * max(nin, 1))(*[x.ptr for x in leaves])
    c_mode = (C.c_int32 * max(nin, 1))(*modes)
    c_imm = (C.c_double * max(nin, 1))()
    c_out = (C.c_void_p * max(nout, 1))(*[o.ptr for o in outs])
    if reduce is None:
        check(D.lib.ag_ewise_fused(D._DT[dtype], c_code, len(code), c_consts, len(consts), nin, c_in, c_mode, c_imm,
                                   nout, c_out, rows, cols))
    else:
        check(D.lib.ag_ewise_fused_reduce(D._DT[dtype], c_code, len(code), c_consts, len(consts), nin, c_in, c_mode,
                                          c_imm, nout, c_out, rows, cols, reduce[0], reduce[1], reduce[2],
                                          C.c_void_p(reduce[4].ptr)))
        stats["fused_reductions"] = stats.get("fused_reductions", 0) + 1
    stats["fused_launches"] += 1
    stats["fused_nodes"] += len(order)
    if trace is not None:
        trace.append((n_elem, sum(1 for m in modes if m == M_FULL), nout, tuple(code)))
    for o, arr in zip(outputs, outs):
        o._set_storage(arr)


# ---- column programs ------------------------------------------------------------------------------------------

class _ColCodegen(_Codegen):
    """Code of one segment of a column program: pending column-space values that have a slot are loaded from it,
    every other pending node is generated in place (element-space nodes used by several segments are recomputed)."""

    def __init__(self, outputs, ext_ids, cslot):
        _Codegen.__init__(self, {}, outputs, ext_ids)
        self.cslot = cslot

    def gen(self, x, push):
        if is_pending(x):
            if id(x) in self.cslot:
                self.emit(CP_LD_COL, self.cslot[id(x)], 1 if push else 0)
                self.pushed(push)
            else:
                self.gen_node(x, push)
        else:
            _Codegen.gen(self, x, push)


def _run_colprog(order, root, reduce=None):
    """Evaluate a DAG over an (m, n) element space and its (1, n) column space -- column sums (`cr` nodes, the softmax
    normaliser and the unbroadcast of its gradient), operations on them, and their broadcast back over the rows -- by
    ONE launch of ag_ewise_colprog.  The program is a sequence of segments: a column segment runs once per column, an
    element segment for every row of the column; a column sum is accumulated (CP_CRED) by the element segment that
    computes its operand and finished (CP_CFIN) by the next column segment.  reduce = (map, out): additionally the sum
    over everything of map(root) goes to `out`."""
    dtype = root.dtype
    SE = None
    for n in order:
        if n.kind == "cr":
            SE = n.args[0].shape
            break
        for a in n.args:
            if is_pending(a) and a.shape != n.shape:
                SE = n.shape
                break
        if SE is not None:
            break
    n = a = None
    if SE is None or len(SE) != 2 or SE[0] < 2 or SE[1] < 2:
        raise _NotFusable()
    m, ncols = SE
    SC = (1, ncols)
    for n in order:
        if n.shape != SE and n.shape != SC:
            raise _NotFusable()
        if n.kind == "cr" and (n.args[0].shape != SE or n.shape != SC):
            raise _NotFusable()
        if n.kind == "mm":
            raise _NotFusable()
    n = None
    uses = {}
    for n in order:
        for a in n.args:
            if is_pending(a):
                uses[id(a)] = uses.get(id(a), 0) + 1
    n = a = None
    ext = []
    for idx in range(len(order)):
        u = uses.get(id(order[idx]), 0)
        ext.append(order[idx] is root or _external(order[idx], u) > 0)
    outputs = [order[i] for i in range(len(order)) if ext[i]]
    ext_ids = {id(o) for o in outputs}
    if len(outputs) > MAXOUT:
        raise _TooBig("outputs", ext_ids)
    # levels: a column sum lives one level above its operand
    lev, used_by_e, feeds = {}, set(), {}
    for n in order:
        lv = 0
        for a in n.args:
            if is_pending(a):
                lv = max(lv, lev[id(a)])
                if a.shape == SC and n.shape == SE:
                    used_by_e.add(id(a))
        if n.kind == "cr":
            lv += 1
            feeds.setdefault(id(n.args[0]), []).append(n)
        lev[id(n)] = lv
    n = a = None
    cslot = {}
    for i in range(len(order)):
        n = order[i]
        if n.shape == SC and (n.kind == "cr" or ext[i] or id(n) in used_by_e or uses.get(id(n), 0) > 1):
            if len(cslot) >= COL_SLOTS:
                raise _TooBig("colslots", ext_ids)
            cslot[id(n)] = len(cslot)
    n = None
    cg = _ColCodegen(outputs, ext_ids, cslot)
    rmap = None if reduce is None else reduce[0]
    for lv in range(max(lev.values()) + 1):
        start = len(cg.code)
        cg.emit(CP_PHASE, 1)
        for i in range(len(order)):
            n = order[i]
            if n.shape != SC or lev[id(n)] != lv or id(n) not in cslot:
                continue
            cg.depth = 0
            if n.kind == "cr":
                cg.emit(CP_CFIN, cslot[id(n)])
            else:
                cg.gen_node(n, False)
                cg.emit(CP_ST_COL, cslot[id(n)])
            if ext[i]:
                cg.emit(STORE, outputs.index(n))
            if n is root and reduce is not None:
                if rmap != "id":
                    cg.emit(UNARY, D.U[rmap])
                cg.emit(REDUCE)
        if len(cg.code) == start + 1:
            cg.code.pop()
        start = len(cg.code)
        cg.emit(CP_PHASE, 0)
        for i in range(len(order)):
            n = order[i]
            if n.shape != SE or lev[id(n)] != lv:
                continue
            is_root = n is root and reduce is not None
            if not (ext[i] or id(n) in feeds or is_root):
                continue
            cg.depth = 0
            cg.gen_node(n, False)
            if ext[i]:
                cg.emit(STORE, outputs.index(n))
            for c in feeds.get(id(n), ()):
                cg.emit(CP_CRED, cslot[id(c)])
            if is_root:
                if rmap != "id":
                    cg.emit(UNARY, D.U[rmap])
                cg.emit(REDUCE)
        if len(cg.code) == start + 1:
            cg.code.pop()
    n = c = None
    code, consts, leaves = cg.code, cg.consts, cg.leaves
    maxdepth = cg.maxdepth
    cg = None
    if len(code) > MAXCODE or maxdepth > SLOTS:
        raise _TooBig("code" if len(code) > MAXCODE else "slots", ext_ids)
    if len(leaves) > DYN_MAXIN or len(outputs) > DYN_MAXOUT:
        raise _TooBig("inputs" if len(leaves) > DYN_MAXIN else "outputs", ext_ids)
    modes = []
    for x in leaves:
        if x.shape == SE:
            modes.append(M_FULL)
        elif x.size == 1:
            modes.append(M_DEVSCALAR)
        elif x.shape == SC:
            modes.append(M_ROWVEC)
        elif x.shape in ((m, 1), (m,)):
            modes.append(M_COLVEC)
        else:
            raise _NotFusable()
    outs = [D.DArray.empty(o.shape, dtype) for o in outputs]
    nin, nout = len(leaves), len(outs)
    c_code = (C.c_int32 * len(code))(*code)
    c_consts = (C.c_double * max(len(consts), 1))(*(consts or [0.0]))
    c_in = (C.c_void_p * max(nin, 1))(*[x.ptr for x in leaves])
    c_mode = (C.c_int32 * max(nin, 1))(*modes)
    c_imm = (C.c_double * max(nin, 1))()
    c_out = (C.c_void_p * max(nout, 1))(*[o.ptr for o in outs])
    check(D.lib.ag_ewise_colprog(D._DT[dtype], c_code, len(code), c_consts, len(consts), nin, c_in, c_mode, c_imm,
                                 nout, c_out, m, ncols, C.c_void_p(reduce[1].ptr) if reduce is not None else None))
    stats["col_launches"] = stats.get("col_launches", 0) + 1
    stats["fused_launches"] += 1
    stats["fused_nodes"] += len(order)
    if reduce is not None:
        stats["fused_reductions"] = stats.get("fused_reductions", 0) + 1
    if trace is not None:
        trace.append((m * ncols, sum(1 for md in modes if md == M_FULL), nout, tuple(code)))
    for o, arr in zip(outputs, outs):
        o._set_storage(arr)
