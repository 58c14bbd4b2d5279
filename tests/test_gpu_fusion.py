"""GPU tests of the fused elementwise evaluator (csrc/fuse.cu, ag_ewise_fused) and of deferred evaluation
(autograd.jl_b200/lazy.py): a fused chain must be BIT-IDENTICAL to the same chain run as single kernels
(every instruction evaluates the expression of the single-rule kernel), for every op of the unary table, its
gradient, the binary maps and their gradient maps, with row / column / scalar operands, odd sizes, Float32
and Float64; whole tapes of the BASELINE configs must give identical parameters with and without fusion."""
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture()
def L(ag):
    """Deferred evaluation on, sums NOT fused into their producers: a fused sum adds in another order than the
    stand-alone reduction kernels, so only the elementwise part can be compared bit for bit."""
    lz = sys.modules["agb200.lazy"]
    prev, prev_r = lz.set_fusion(True), lz.reduce_fusion[0]
    lz.reduce_fusion[0] = False
    yield lz
    lz.flush()
    lz.reduce_fusion[0] = prev_r
    lz.set_fusion(prev)


@pytest.fixture()
def LR(ag):
    lz = sys.modules["agb200.lazy"]
    prev, prev_r = lz.set_fusion(True), lz.reduce_fusion[0]
    lz.reduce_fusion[0] = True
    yield lz
    lz.flush()
    lz.reduce_fusion[0] = prev_r
    lz.set_fusion(prev)


def both_modes(L, fn):
    """fn() -> list of DArrays; returns (fused results, unfused results) as numpy."""
    out = []
    for on in (True, False):
        prev = L.set_fusion(on)
        try:
            out.append([r.to_numpy() for r in fn()])
        finally:
            L.set_fusion(prev)
    return out


def same(a, b):
    return a.shape == b.shape and np.array_equal(a, b, equal_nan=True)


UNARY = ["identity", "neg", "abs", "abs2", "exp", "log", "tanh", "sigm", "sqrt", "sin", "cos", "tan", "sinh", "cosh",
         "asin", "acos", "atan", "asinh", "acosh", "atanh", "exp2", "exp10", "expm1", "log2", "log10", "log1p",
         "cbrt", "sign", "relu", "one"]


def _domain(op, rng, shape, dtype):
    x = rng.random(shape) * 0.9 + 0.05                  # (0.05, 0.95)
    if op == "acosh":
        x = x + 1.0
    elif op not in ("log", "sqrt", "log2", "log10", "asin", "acos", "atanh"):
        x = (x - 0.5) * 4.0
    return np.asfortranarray(x.astype(dtype))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("shape", [(64, 96), (10, 333), (7,)])
def test_every_unary_op_and_gradient_fused_equals_single_kernels(ag, L, dtype, shape):
    D = sys.modules["agb200.darray"]
    rng = np.random.default_rng(11)
    for op in UNARY:
        x = ag.DArray.from_numpy(_domain(op, rng, shape, dtype))
        dy = ag.DArray.from_numpy(np.asfortranarray(rng.standard_normal(shape).astype(dtype)))
        bias = ag.DArray.from_numpy(rng.standard_normal((shape[0],) + (1,) * (len(shape) - 1)).astype(dtype) * 0.01)

        def fn():
            y = D.unary_fwd(op, x)
            g = D.unary_bwd(op, dy, x, y)                          # needs y (kept) and x
            z = D.binary_fwd("add", D.binary_fwd("mul", g, 0.5), bias)   # chain on top, column-vector operand
            g2 = D.unary_bwd(op, 1.25, x, y)                       # host-scalar dy
            return [y, g, z, g2]
        f, u = both_modes(L, fn)
        for a, b, name in zip(f, u, ("y", "g", "z", "g2")):
            assert same(a, b), (op, name, dtype, shape)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_binary_maps_and_gradient_maps_with_broadcast_operands(ag, L, dtype):
    D = sys.modules["agb200.darray"]
    rng = np.random.default_rng(12)
    R, Cc = 10, 1030
    mk = lambda s: ag.DArray.from_numpy(np.asfortranarray((rng.random(s) + 0.5).astype(dtype)))
    a, b, dy = mk((R, Cc)), mk((R, Cc)), mk((R, Cc))
    row, col, sc = mk((1, Cc)), mk((R, 1)), mk((1, 1))
    for op in ("add", "sub", "mul", "div", "max", "min", "pow", "eq", "lt", "le", "gt", "ge", "atan2", "hypot"):
        def fn():
            r = [D.binary_fwd(op, a, b), D.binary_fwd(op, a, row), D.binary_fwd(op, col, a), D.binary_fwd(op, a, 0.75),
                 D.binary_fwd(op, 0.75, a), D.binary_fwd(op, a, sc), D.binary_fwd(op, row, col)]
            r.append(D.unary_fwd("exp", D.binary_fwd(op, D.binary_fwd("mul", a, row), D.binary_fwd("add", b, col))))
            return r
        f, u = both_modes(L, fn)
        for i, (x, y) in enumerate(zip(f, u)):
            assert same(x, y), (op, i, dtype)
    for op in ("add", "sub", "mul", "div", "max", "min", "pow", "atan2", "hypot"):
        for argno in (1, 2):
            def fn():
                y = D.binary_fwd(op, a, row)
                g = D.binary_bwd(op, argno, dy, a, row, y)
                h = D.binary_bwd(op, argno, dy, a, 0.75, D.binary_fwd(op, a, 0.75))
                return [g, h, D.binary_fwd("add", g, h)]
            f, u = both_modes(L, fn)
            for i, (x, y) in enumerate(zip(f, u)):
                assert same(x, y), (op, argno, i, dtype)


def test_softmax_gradient_chain_is_one_launch_and_exact(ag, L):
    """The MNIST backward chain expand -> dy.*e -> addto! (src/base.jl:245, src/math.jl:42, src/addto.jl:23)."""
    D = sys.modules["agb200.darray"]
    rng = np.random.default_rng(13)
    B = 4099
    e = ag.DArray.from_numpy(np.asfortranarray(rng.random((10, B), dtype=np.float32)))
    dn = ag.DArray.from_numpy(np.asfortranarray(rng.standard_normal((10, B)).astype(np.float32)))
    ds = ag.DArray.from_numpy(rng.standard_normal((1, B)).astype(np.float32))

    def fn():
        return [D.add(dn, D.unary_bwd("exp", D.bcast_to(ds, (10, B)), None, e))]
    n0 = ag.launch_count()
    r = fn()[0]
    got = r.to_numpy()
    assert ag.launch_count() - n0 == 1
    f, u = both_modes(L, fn)
    assert same(f[0], u[0]) and same(f[0], got)
    ref = dn.to_numpy().astype(np.float64) + ds.to_numpy().astype(np.float64) * e.to_numpy().astype(np.float64)
    assert np.max(np.abs(got - ref)) <= 1e-6 * np.max(np.abs(ref))


def test_large_and_misaligned_shapes(ag, L):
    D = sys.modules["agb200.darray"]
    rng = np.random.default_rng(14)
    for shape in [(1 << 21,), (1023, 1025), (3, 5, 7), (2, 3, 4, 5)]:
        x = ag.DArray.from_numpy(np.asfortranarray(rng.standard_normal(shape).astype(np.float32)))
        v = ag.DArray.from_numpy(rng.standard_normal((shape[0],) + (1,) * (len(shape) - 1)).astype(np.float32))

        def fn():
            t = D.unary_fwd("tanh", D.binary_fwd("add", D.binary_fwd("mul", 0.1, x), v))
            return [t, D.unary_bwd("tanh", D.unary_fwd("abs2", t), None, t)]
        f, u = both_modes(L, fn)
        for a, b in zip(f, u):
            assert same(a, b), shape


def _step_all(ag, W, cfg, fused, L):
    prev = L.set_fusion(fused)
    try:
        rng = np.random.default_rng(21)
        dev = ag.DArray.from_numpy
        if cfg == "mnist":
            w, x, y = W.mnist_init(rng, 4096)
            xd, yd = dev(x), dev(y)
            loss = lambda p: W.mnist_loss(ag, p, xd, yd)
        elif cfg == "sweep":
            x, brow, bcol = W.sweep_init(rng, 1024, 512)
            xd = dev(x)
            w = [brow, bcol]
            loss = lambda p: W.sweep_loss(ag, xd, p[0], p[1])
        elif cfg == "rnn":
            H, V, B, T = 32, 20, 64, 4
            w = W.rnn_init(rng, H, V, np.float32)
            xs, y = [], np.zeros((V, B), np.float32)
            for _ in range(T):
                oh = np.zeros((V, B), np.float32)
                oh[rng.integers(0, V, B), np.arange(B)] = 1
                xs.append(dev(np.asfortranarray(oh)))
            y[rng.integers(0, V, B), np.arange(B)] = 1
            yd, hd = dev(np.asfortranarray(y)), dev(np.zeros((H, B), np.float32))
            loss = lambda p: W.rnn_loss(ag, p, xs, yd, hd)
        else:
            H, E, V, B, T = 64, 32, 40, 48, 5
            wd = W.lstm_init(rng, H, E, V, np.float32)
            keys = list(wd)
            w = [wd[k] for k in keys]
            ids = [rng.integers(0, V, B) for _ in range(T)]
            tg = []
            for _ in range(T):
                oh = np.zeros((V, B), np.float32)
                oh[rng.integers(0, V, B), np.arange(B)] = 1
                tg.append(dev(np.asfortranarray(oh)))
            h0, c0 = dev(np.zeros((H, B), np.float32)), dev(np.zeros((H, B), np.float32))
            loss = lambda p: W.lstm_loss(ag, dict(zip(keys, p)), ids, tg, h0, c0)
        p = [ag.Param(dev(a)) for a in w]
        n0 = ag.launch_count()
        losses = []
        for _ in range(2):
            t = ag.diff(lambda: loss(p))
            losses.append(ag.value(t))
            for q in p:
                ag.axpy(-0.05, ag.grad(t, q), q)
        res = [float(v) for v in losses] + [q.value.to_numpy() for q in p]
        return res, ag.launch_count() - n0
    finally:
        L.set_fusion(prev)


@pytest.mark.parametrize("cfg", ["mnist", "sweep", "rnn", "lstm"])
def test_whole_steps_identical_with_and_without_fusion(ag, L, cfg):
    W = ag.workloads
    rf, nf = _step_all(ag, W, cfg, True, L)
    ru, nu = _step_all(ag, W, cfg, False, L)
    for a, b in zip(rf, ru):
        assert np.array_equal(np.asarray(a), np.asarray(b))
    assert nf < nu, (cfg, nf, nu)
    print("%s: %d launches fused, %d unfused" % (cfg, nf, nu))


@pytest.mark.parametrize("cfg", ["mnist", "sweep", "rnn", "lstm"])
def test_specialised_and_interpreted_kernels_agree(ag, L, cfg):
    """The ahead-of-time instantiated programs (fuse_static.inc) and the general interpreter give the same bits."""
    W = ag.workloads
    try:
        L.set_static(False)
        ri, _ = _step_all(ag, W, cfg, True, L)
    finally:
        L.set_static(True)
    rs, _ = _step_all(ag, W, cfg, True, L)
    for a, b in zip(ri, rs):
        assert np.array_equal(np.asarray(a), np.asarray(b))


@pytest.mark.parametrize("cfg", ["mnist", "sweep", "rnn", "lstm"])
def test_every_fused_launch_specialised_equals_interpreted(ag, L, cfg):
    """Each ag_ewise_fused call of a whole step is executed on BOTH evaluators and the outputs compared bit for
    bit (this is what catches e.g. a mul of one instruction contracted with the add of the next into an FMA)."""
    import ctypes as C
    D = sys.modules["agb200.darray"]
    lib, real = D.lib, D.lib.ag_ewise_fused
    bad, seen = {}, [0]

    def d2h(ptr, n, dt):
        out = np.empty(n, dt)
        ag._lib.check(lib.ag_copy_d2h(out.ctypes.data_as(C.c_void_p), C.c_void_p(ptr), out.nbytes))
        return out

    class Shim:
        def __getattr__(self, k):
            return getattr(lib, k)

        def ag_ewise_fused(self, dtype, code, ncode, consts, nconsts, nin, ins, modes, imms, nout, outs, rows, cols):
            dt, n, res = (np.float32 if dtype == 0 else np.float64), rows * cols, []
            if nin > 8 or nout > 6:          # only a specialised kernel takes these; nothing to compare with
                return real(dtype, code, ncode, consts, nconsts, nin, ins, modes, imms, nout, outs, rows, cols)
            for static in (1, 0):
                lib.ag_ewise_fused_set_static(static)
                st = real(dtype, code, ncode, consts, nconsts, nin, ins, modes, imms, nout, outs, rows, cols)
                if st:
                    lib.ag_ewise_fused_set_static(1)
                    return st
                res.append([d2h(outs[k], n, dt) for k in range(nout)])
            lib.ag_ewise_fused_set_static(1)
            seen[0] += 1
            for k in range(nout):
                if not np.array_equal(res[0][k], res[1][k], equal_nan=True):
                    bad.setdefault(tuple(hex(int(code[j])) for j in range(ncode)), []).append(k)
            return real(dtype, code, ncode, consts, nconsts, nin, ins, modes, imms, nout, outs, rows, cols)
    D.lib = Shim()
    try:
        _step_all(ag, ag.workloads, cfg, True, L)
    finally:
        D.lib = lib
        lib.ag_ewise_fused_set_static(1)
    assert seen[0] > 0 and not bad, bad


def test_graph_capture_of_a_fused_step(ag, L):
    W = ag.workloads
    w, x, y = W.mnist_init(np.random.default_rng(9), 2048)
    xd, yd = ag.DArray.from_numpy(x), ag.DArray.from_numpy(y)

    def step(params):
        t = ag.diff(lambda: W.mnist_loss(ag, params, xd, yd))
        for p in params:
            ag.axpy(-0.1, ag.grad(t, p), p)
        return ag.value(t)
    pe = [ag.Param(ag.DArray.from_numpy(a)) for a in w]
    pg = [ag.Param(ag.DArray.from_numpy(a)) for a in w]
    prev = L.set_fusion(False)
    for _ in range(3):
        le = step(pe)                          # unfused eager reference
    L.set_fusion(True)
    step(pg)
    with ag.StepGraph() as g:
        lg = step(pg)
    assert L.pending_count() == 0, "everything the step defined must be inside the captured graph"
    g.launch()
    g.launch()
    ag.sync()
    L.set_fusion(prev)
    for a, b in zip(pe, pg):
        assert np.array_equal(a.value.to_numpy(), b.value.to_numpy())
    assert float(le) == float(lg)
    g.destroy()


# ---- sums fused into the launch of the chain they reduce (ag_ewise_fused_reduce) ---------------------------------

@pytest.mark.parametrize("dtype,tol", [(np.float32, 1e-6), (np.float64, 1e-14)])
@pytest.mark.parametrize("shape", [(10, 4099), (64, 5000), (128, 300), (1024, 96), (7, 33), (2048, 5), (100, 64), (12,)])
def test_fused_sums_against_float64(ag, LR, dtype, tol, shape):
    """Whole-array sums, column sums (dims=1) and row sums (dims=2) of a pending chain, for the run lengths /
    row counts of the configs (10, 64, 128, 1024) and awkward ones; reference = numpy in float64."""
    D = sys.modules["agb200.darray"]
    rng = np.random.default_rng(31)
    x = np.asfortranarray(rng.standard_normal(shape).astype(dtype))
    b = rng.standard_normal((shape[0],) + (1,) * (len(shape) - 1)).astype(dtype)
    xd, bd = ag.DArray.from_numpy(x), ag.DArray.from_numpy(b)
    chain = lambda: D.unary_fwd("tanh", D.binary_fwd("add", D.binary_fwd("mul", xd, 0.5), bd))
    ref = np.tanh((x * dtype(0.5) + b).astype(dtype)).astype(np.float64)     # per-op rounding of the chain
    f0 = LR.stats.get("fused_reductions", 0)
    cases = [(None, "id", ref.sum()), (None, "abs2", (ref * ref).sum())]
    if len(shape) == 2:
        cases += [(0, "id", ref.sum(axis=0, keepdims=True)), (1, "id", ref.sum(axis=1, keepdims=True)),
                  (1, "abs2", (ref * ref).sum(axis=1, keepdims=True))]
    for dims, mp, want in cases:
        y = chain()
        f = {"id": ag.sum_, "abs2": ag.sumabs2}[mp]
        got = f(y) if dims is None else f(y, dims=dims)
        got = got.to_numpy().astype(np.float64)
        scale = np.abs(ref).sum() if dims is None else np.abs(ref).sum(axis=dims, keepdims=True)
        assert got.shape == np.asarray(want).shape
        assert np.all(np.abs(got - want) <= tol * np.maximum(scale, 1e-30)), (dims, mp, shape)
        # the chain's own value is still the chain's value (stored by the same launch)
        assert np.allclose(y.to_numpy(), ref, rtol=2e-6 if dtype == np.float32 else 1e-13, atol=0)
    if len(shape) == 2 and shape[0] * shape[1] % 4 == 0:
        assert LR.stats.get("fused_reductions", 0) > f0, "no sum was fused for %s" % (shape,)


@pytest.mark.parametrize("cfg", ["mnist", "sweep", "rnn", "lstm"])
def test_whole_steps_with_fused_sums_within_tolerance(ag, LR, cfg):
    W = ag.workloads
    f0 = LR.stats.get("fused_reductions", 0)
    rf, nf = _step_all(ag, W, cfg, True, LR)
    assert LR.stats.get("fused_reductions", 0) > f0
    ru, nu = _step_all(ag, W, cfg, False, LR)
    for a, b in zip(rf, ru):
        a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
        assert np.max(np.abs(a - b)) <= 1e-5 * max(np.max(np.abs(b)), 1e-30)
    assert nf < nu
    print("%s: %d launches with fused sums, %d unfused" % (cfg, nf, nu))


def test_fused_sum_is_deterministic(ag, LR):
    D = sys.modules["agb200.darray"]
    x = ag.DArray.from_numpy(np.asfortranarray(np.random.default_rng(5).standard_normal((64, 8192)).astype(np.float32)))
    outs = []
    for _ in range(3):
        y = D.unary_fwd("exp", D.binary_fwd("mul", x, 0.1))
        outs.append((float(ag.sum_(y)), ag.sum_(D.unary_fwd("exp", D.binary_fwd("mul", x, 0.1)), dims=1).to_numpy()))
    assert outs[0][0] == outs[1][0] == outs[2][0]
    assert np.array_equal(outs[0][1], outs[1][1]) and np.array_equal(outs[0][1], outs[2][1])


@pytest.mark.parametrize("cfg", ["mnist", "rnn", "lstm"])
def test_column_programs_on_the_device(ag, L, cfg):
    """Column fusion on (ag_ewise_colprog: softmax normaliser, what follows it and the unbroadcast of its gradient in one
    launch each): the whole step agrees with the default path to rounding of the sums (double accumulation on both
    sides), needs fewer launches, and the specialised and the interpreted column kernels give the same bits."""
    W = ag.workloads
    r0, n0 = _step_all(ag, W, cfg, True, L)
    prev = L.set_col_fusion(True)
    try:
        c0 = L.stats.get("col_launches", 0)
        r1, n1 = _step_all(ag, W, cfg, True, L)
        assert L.stats.get("col_launches", 0) > c0
        try:
            L.set_static(False)
            r2, _ = _step_all(ag, W, cfg, True, L)
        finally:
            L.set_static(True)
    finally:
        L.set_col_fusion(prev)
    for a, b in zip(r0, r1):
        a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
        assert np.max(np.abs(a - b)) <= 1e-5 * max(np.max(np.abs(a)), 1e-30)
    for a, b in zip(r1, r2):
        assert np.array_equal(np.asarray(a), np.asarray(b))
    assert n1 < n0, (cfg, n1, n0)
