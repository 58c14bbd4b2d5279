"""CPU tests of the HOST side of the product (tape runtime, rule dispatch, stride collapsing, index
bookkeeping) against the oracle, with the C ABI replaced by tests/fake_abi.py.  The real kernels are
checked by the -m gpu tests."""
import numpy as np
import pytest


def to_dev(ag, a):
    return ag.DArray.from_numpy(a)


def relerr(got, ref):
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    return np.max(np.abs(got - ref)) / max(np.max(np.abs(ref)), 1e-300)


def run_both(hostlib, oracle, loss, params_np, consts_np, dtype):
    """loss(m, params, *consts) on oracle (float64) and on the device path; returns (loss, grads) x2."""
    ag = hostlib
    po = [oracle.Param(np.asfortranarray(p.astype(np.float64))) for p in params_np]
    co = [np.asfortranarray(c.astype(np.float64)) if isinstance(c, np.ndarray) and c.dtype.kind == "f" else c
          for c in consts_np]
    to = oracle.diff(lambda: loss(oracle, po, *co))
    go = [oracle.full(oracle.grad(to, p)) for p in po]
    pd = [ag.Param(to_dev(ag, p.astype(dtype))) for p in params_np]
    cd = [to_dev(ag, c.astype(dtype)) if isinstance(c, np.ndarray) and c.dtype.kind == "f" else c
          for c in consts_np]
    td = ag.diff(lambda: loss(ag, pd, *cd))
    gd = [ag.full(ag.grad(td, p)) for p in pd]
    gd = [None if g is None else (g if isinstance(g, float) else g.to_numpy()) for g in gd]
    return float(oracle.value(to)), go, float(ag.value(td)), gd


def test_readme_known_answer(hostlib):
    ag = hostlib
    x = ag.Param(to_dev(ag, np.array([1.0, 2.0, 3.0])))
    t = ag.diff(lambda: ag.sumabs2(x))
    assert float(ag.value(t)) == 14.0
    assert np.array_equal(ag.grad(t, x).to_numpy(), [2.0, 4.0, 6.0])
    g = ag.grad(lambda v: ag.sum_(ag.pow_(v, 2)))(to_dev(ag, np.array([[1.0, 2.0], [3.0, 4.0]])))
    assert np.array_equal(g.to_numpy(), [[2.0, 4.0], [6.0, 8.0]])


def test_tape_order_matches_readme(hostlib):
    """README.md:116-125 pins the node order of sum(abs2, w*x .+ b - y)."""
    ag = hostlib
    rng = np.random.default_rng(0)
    w = ag.Param(to_dev(ag, rng.standard_normal((1, 3))))
    b = ag.Param(to_dev(ag, np.zeros((1, 1))))
    x, y = to_dev(ag, rng.standard_normal((3, 4))), to_dev(ag, rng.standard_normal((1, 4)))
    t = ag.diff(lambda: ag.sumabs2(ag.sub(ag.add(ag.matmul(w, x), b), y)))
    names = [("P" if isinstance(n.Value, ag.Param) else n.Value.func.name) for n in t.list]
    assert names == ["P", "*", "P", "+", "-", "sum(abs2,·)"]


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-12), (np.float32, 1e-5)])
def test_housing(hostlib, oracle, dtype, tol):
    W = hostlib.workloads
    w, x, y = W.housing_init(np.random.default_rng(1))
    lo, go, ld, gd = run_both(hostlib, oracle, W.housing_loss, w, [x, y], dtype)
    assert abs(lo - ld) <= tol * 10 * max(abs(lo), 1)
    for a, b in zip(gd, go):
        assert relerr(a, b) <= tol * 10


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-12), (np.float32, 1e-5)])
def test_mnist_mlp(hostlib, oracle, dtype, tol):
    W = hostlib.workloads
    w, x, y = W.mnist_init(np.random.default_rng(2), 100)
    lo, go, ld, gd = run_both(hostlib, oracle, W.mnist_loss, w, [x, y], dtype)
    assert abs(lo - ld) <= tol * 10 * max(abs(lo), 1)
    for a, b in zip(gd, go):
        assert a.shape == b.shape
        assert relerr(a, b) <= tol * 10


def test_sweep(hostlib, oracle):
    W = hostlib.workloads
    x, brow, bcol = W.sweep_init(np.random.default_rng(3), 64, 48)

    def loss(m, p):
        return W.sweep_loss(m, p[0], p[1], p[2])
    lo, go, ld, gd = run_both(hostlib, oracle, loss, [x, brow, bcol], [], np.float64)
    assert abs(lo - ld) <= 1e-11 * abs(lo)
    for a, b in zip(gd, go):
        assert a.shape == b.shape and relerr(a, b) <= 1e-11


def test_lstm_small(hostlib, oracle):
    W = hostlib.workloads
    rng = np.random.default_rng(4)
    H, E, V, B, T = 16, 8, 12, 6, 3
    w = W.lstm_init(rng, H, E, V, np.float64)
    keys = list(w)
    ids = [rng.integers(0, V, B) for _ in range(T)]
    ids[0][1] = ids[0][0]                       # repeated index inside one step
    tg = []
    for t in range(T):
        oh = np.zeros((V, B))
        oh[rng.integers(0, V, B), np.arange(B)] = 1
        tg.append(np.asfortranarray(oh))
    h0, c0 = np.zeros((H, B)), np.zeros((H, B))

    def loss(m, p, h0, c0, *tg):
        return W.lstm_loss(m, dict(zip(keys, p)), ids, list(tg), h0, c0)
    lo, go, ld, gd = run_both(hostlib, oracle, loss, [w[k] for k in keys], [h0, c0] + tg, np.float64)
    assert abs(lo - ld) <= 1e-11 * abs(lo)
    for k, a, b in zip(keys, gd, go):
        assert a.shape == b.shape, k
        assert relerr(a, b) <= 1e-11, k


def test_sparse_known_answer(hostlib):
    """test/sparse.jl:7-16."""
    ag = hostlib
    z = to_dev(ag, np.zeros((3, 4)))
    one2 = to_dev(ag, np.array([1.0, 1.0]))
    a = ag.Sparse(z, [one2, one2, 1.0, 1.0], [([0, 1],), (range(2, 4),), (1, 1), (0,)])
    fa = a.full().to_numpy()
    exp = np.zeros(12)
    exp[:5] = [2, 1, 1, 1, 1]
    assert np.array_equal(fa.reshape(-1, order="F"), exp)
    assert np.array_equal((a + a).full().to_numpy(), 2 * fa)
    assert isinstance(ag.addto(a, a), ag.Sparse)


def test_flat_dest_matches_oracle(hostlib, oracle):
    """Index bookkeeping must be bit-exact: every index kind of test/getindex.jl:8-26."""
    ag = hostlib
    shape = (3, 4)
    mask = np.zeros(shape, bool)
    mask[1, 2] = mask[0, 0] = True
    kinds = [(slice(None),), (0,), (range(0, 2),), (range(0, 3, 2),), ([0, 1],), ([1, 1],), ([],),
             (mask,), (slice(None), 1), (1, slice(None)), ([0, 2], [1, 1, 3]), (np.array([[0, 1], [2, 2]]),),
             ((1, 2),), (slice(None), np.array([True, False, True, False]))]
    for idx in kinds:
        d1, s1 = ag.flat_dest(shape, idx)
        d2, s2 = oracle.flat_dest(shape, idx)
        assert s1 == s2 and d1.dtype == np.int64 and np.array_equal(d1, d2), idx


def test_getindex_grads(hostlib, oracle):
    ag = hostlib
    rng = np.random.default_rng(5)
    x = rng.standard_normal((3, 4))
    for idx in [(slice(None), [1, 1, 3]), ([1, 1],), (range(0, 2), 2), (slice(None), 1), (5,)]:
        go = oracle.grad(lambda v: oracle.sumabs2(oracle.getindex(v, *idx)))(np.asfortranarray(x.copy()))
        gd = ag.grad(lambda v: ag.sumabs2(ag.getindex(v, *idx)))(to_dev(ag, x))
        assert relerr(gd.to_numpy(), go) <= 1e-14, idx


def test_highorder_tanh(hostlib):
    """test/highorder.jl:33-34 on a 1-element device array."""
    ag = hostlib
    x = to_dev(ag, np.array([1.0]))
    g1 = ag.grad(lambda v: ag.sum_(ag.tanh(v)))
    g2 = ag.grad(lambda v: ag.sum_(g1(v)))
    t = np.tanh(1.0)
    assert abs(g1(x).to_numpy()[0] - (1 - t * t)) < 1e-15
    assert abs(g2(x).to_numpy()[0] - (-2 * t * (1 - t * t))) < 1e-15


def test_neuralnet_hessian_vs_oracle(hostlib, oracle):
    """test/neuralnet.jl + test/highorder.jl:36-42: grad of (sum of squared grad) through relu MLP."""
    rng = np.random.default_rng(6)
    w1, w2 = rng.standard_normal((5, 4)), rng.standard_normal((2, 5))
    x, y = rng.standard_normal((4, 7)), rng.standard_normal((2, 7))

    def make(m, dev):
        X, Y, W2 = dev(x), dev(y), dev(w2)

        def loss(w):
            return m.sumabs2(m.sub(m.matmul(W2, m.tanh(m.matmul(w, X))), Y))
        inner = m.grad(loss)
        return m.grad(lambda w: m.sumabs2(inner(w)))
    go = make(oracle, lambda a: np.asfortranarray(a.copy()))(np.asfortranarray(w1.copy()))
    gd = make(hostlib, lambda a: to_dev(hostlib, a))(to_dev(hostlib, w1))
    assert relerr(gd.to_numpy(), go) <= 1e-12


def test_unbroadcast_shapes(hostlib, oracle):
    """test/base.jl:46-58 shapes: (2,3,5) with (1,3,5), (2,3,5,4) with (2,3,1,4), vectors."""
    rng = np.random.default_rng(7)
    for sa, sb in [((2, 3, 5), (1, 3, 5)), ((2, 3, 5, 4), (2, 3, 1, 4)), ((4, 6), (4,)), ((4, 6), (1, 6)),
                   ((4, 6), (1, 1)), ((3,), (3, 5))]:
        a, b = rng.standard_normal(sa), rng.standard_normal(sb)
        for name in ("add", "sub", "mul", "div", "max_", "min_"):
            def loss(m, p):
                return m.sumabs2(getattr(m, name)(p[0], p[1]))
            lo, go, ld, gd = run_both(hostlib, oracle, loss, [a, b], [], np.float64)
            assert abs(lo - ld) <= 1e-12 * max(abs(lo), 1), (name, sa, sb)
            for u, v in zip(gd, go):
                assert np.shape(u) == np.shape(v), (name, sa, sb)
                assert relerr(u, v) <= 1e-12, (name, sa, sb)


def test_gradcheck_unary(hostlib):
    ag = hostlib
    rng = np.random.default_rng(8)
    x = rng.uniform(0.2, 0.8, (3, 2))
    for name in ["exp", "log", "tanh", "sigm", "sqrt", "sin", "cos", "tan", "sinh", "cosh", "asin", "acos",
                 "atan", "asinh", "atanh", "exp2", "exp10", "expm1", "log2", "log10", "log1p", "cbrt", "abs_",
                 "abs2", "neg"]:
        assert ag.gradcheck(getattr(ag, name), to_dev(ag, x)), name
    assert ag.gradcheck(ag.acosh, to_dev(ag, x + 1.5))


def test_cat(hostlib, oracle):
    rng = np.random.default_rng(9)
    a, b = rng.standard_normal((2, 3)), rng.standard_normal((4, 3))

    def loss(m, p):
        return m.sumabs2(m.mul(m.vcat(p[0], p[1]), 2.0))
    lo, go, ld, gd = run_both(hostlib, oracle, loss, [a, b], [], np.float64)
    assert abs(lo - ld) <= 1e-12 * abs(lo)
    for u, v in zip(gd, go):
        assert relerr(u, v) <= 1e-13


def test_reductions_and_permute(hostlib, oracle):
    """maximum / minimum / prod / permutedims / mean / var (src/base.jl:202-221, src/statistics.jl:20-35)."""
    rng = np.random.default_rng(10)
    x = rng.uniform(0.5, 1.5, (3, 4, 5))
    for name, kw in [("maximum", {}), ("maximum", {"dims": 1}), ("minimum", {"dims": (0, 2)}), ("prod", {"dims": 2}),
                     ("mean", {"dims": 0}), ("sumabs", {"dims": 1})]:
        def loss(m, p):
            return m.sumabs2(getattr(m, name)(p[0], **kw))
        lo, go, ld, gd = run_both(hostlib, oracle, loss, [x], [], np.float64)
        assert abs(lo - ld) <= 1e-12 * abs(lo), (name, kw)
        assert relerr(gd[0], go[0]) <= 1e-12, (name, kw)

    def loss(m, p):
        return m.sumabs2(m.mul(m.permutedims(p[0], (2, 0, 1)), 2.0))
    lo, go, ld, gd = run_both(hostlib, oracle, loss, [x], [], np.float64)
    assert abs(lo - ld) <= 1e-12 * abs(lo) and relerr(gd[0], go[0]) <= 1e-12
    ag = hostlib
    xd = ag.DArray.from_numpy(x)
    assert abs(float(ag.var(xd)) - x.var(ddof=1)) < 1e-12
    assert relerr(ag.var(xd, dims=1).to_numpy(), x.var(axis=1, ddof=1, keepdims=True)) < 1e-12
    assert abs(float(ag.norm(xd)) - np.linalg.norm(x)) < 1e-12
    assert ag.gradcheck(lambda z: ag.std(z, dims=0), xd)


def test_primitive_expr(hostlib, oracle):
    """User-defined primitives (src/macros.jl:88-105; sigm of examples/charlm.jl:46-47) compiled to the fused
    evaluator: value, first and second derivative against the built-in primitive / closed forms."""
    ag = hostlib
    E = ag.E
    sigm2 = ag.primitive_expr("sigm2", lambda x: 1 / (1 + E.exp(-x)), lambda dy, y, x: dy * y * (1 - y))
    swish = ag.primitive_expr("swish", lambda x: x * E.sigm(x),
                              lambda dy, y, x: dy * (E.sigm(x) + x * E.sigm(x) * (1 - E.sigm(x))))
    scaled = ag.primitive_expr("scaled", lambda a, b: a * E.tanh(b), lambda dy, y, a, b: dy * E.tanh(b),
                               lambda dy, y, a, b: dy * a * (1 - E.tanh(b) ** 2))
    rng = np.random.default_rng(11)
    x = rng.standard_normal((5, 7))
    xd = ag.DArray.from_numpy(x)
    assert relerr(sigm2(xd).to_numpy(), 1 / (1 + np.exp(-x))) < 1e-15
    assert relerr(ag.grad(lambda v: ag.sum_(sigm2(v)))(xd).to_numpy(), ag.grad(lambda v: ag.sum_(ag.sigm(v)))(xd).to_numpy()) < 1e-14
    assert ag.gradcheck(swish, xd)
    assert ag.gradcheck(scaled, xd, ag.DArray.from_numpy(rng.standard_normal((5, 7))))
    assert ag.gradcheck(scaled, xd, ag.DArray.from_numpy(rng.standard_normal((5, 1))))      # broadcasting operand
    # second order: the gradient expression records itself under the outer tape
    g1 = ag.grad(lambda v: ag.sum_(sigm2(v)))
    g2 = ag.grad(lambda v: ag.sum_(g1(v)))(xd).to_numpy()
    s = 1 / (1 + np.exp(-x))
    assert relerr(g2, s * (1 - s) * (1 - 2 * s)) < 1e-13
    # host scalars never touch the device
    assert abs(sigm2(0.0) - 0.5) < 1e-15
    kernels = ag.launch_count()
    sigm2(xd)
    assert ag.launch_count() - kernels == 1                # one fused pass


def test_gradcheck_every_primitive_host_logic(hostlib):
    """The same catalogue as the GPU test, through the host rules on the numpy stand-in (CPU)."""
    from gradcheck_catalogue import check_catalogue
    check_catalogue(hostlib)

def _clip_case(ag, dtype, gclip):
    rng = np.random.default_rng(41)
    shapes = [(6, 5), (6, 1), (3, 6), (3,)]
    w = [rng.standard_normal(s).astype(dtype) for s in shapes]
    g = [(rng.standard_normal(s) * 3).astype(dtype) for s in shapes]
    params = [ag.Param(ag.DArray.from_numpy(a)) for a in w]
    grads = [ag.DArray.from_numpy(a) for a in g]
    dp = ag.DataParallel(0, 1)
    dp.update(-0.1, grads, params, gclip=gclip)
    gn = np.sqrt(sum(float((a.astype(np.float64) ** 2).sum()) for a in g))
    scale = min(1.0, gclip / gn) if gclip else 1.0
    return [p.value.to_numpy() for p in params], [a.astype(np.float64) - 0.1 * scale * b.astype(np.float64) for a, b in zip(w, g)], gn


def test_clipped_update_host_logic(hostlib):
    """examples/charlm.jl:116-118,148-152: the step is scaled by min(1, gclip/gnorm); gnorm over ALL parameters."""
    for gclip in (5.0, 1e6, None):
        got, want, gn = _clip_case(hostlib, np.float64, gclip)
        assert gn > 5.0
        for a, b in zip(got, want):
            assert np.allclose(a, b, rtol=1e-12), gclip


def test_timer_labels_follow_the_reference(hostlib):
    """AUTOGRAD_TIMER labels (src/core.jl:165-166,177-182): "[>ti]f" forward, "[ti]f[i]" back, "[ti]addto![i]"."""
    ag = hostlib
    w = ag.Param(ag.DArray.from_numpy(np.ones((2, 3))))
    x = ag.DArray.from_numpy(np.ones((3, 4)))
    prev = ag.set_timer(True)
    try:
        ag.timer_report()
        ag.diff(lambda: ag.sumabs2(ag.matmul(w, x)))
        rows = {r[0]: r for r in ag.timer_report()}
    finally:
        ag.set_timer(prev)
    # tape: 1 P(w), 2 *(w,x), 3 sum(abs2, .)   (README.md:116-125 numbering, 1-based)
    # ftimer prints 1 + length(tape.list) at call time: the Param's node is only created when `*` is recorded
    assert "[>1]*" in rows and "[>3]sum(abs2,·)" in rows
    assert "[3]sum(abs2,·)[1]" in rows and "[3]addto![1]" in rows
    assert "[2]*[1]" in rows and "[2]addto![1]" in rows
    assert all(r[1] == 1 for r in rows.values())


def test_math_rest_host_rules_against_the_oracle(hostlib, oracle):
    """The remainder of src/math.jl (degree / reciprocal / pi-scaled trigonometry, sinc / cosc, clamp, ldexp,
    log(b, x)): values and gradients of the product's composed forwards + reference rules against the oracle's
    (numpy's own functions + the same rules), first and second order."""
    from gradcheck_catalogue import MATH_REST, MATH_REST_DOMAINS
    ag = hostlib
    rng = np.random.default_rng(41)
    for name in MATH_REST:
        lo, hi = MATH_REST_DOMAINS.get(name, (1.2, 2.0))
        x = np.asfortranarray(rng.uniform(lo, hi, (5, 4)))
        xd = ag.DArray.from_numpy(x)
        fo, fd = getattr(oracle, name), getattr(ag, name)
        assert np.allclose(fd(xd).to_numpy(), fo(x), rtol=1e-13, atol=1e-15), name
        go = oracle.grad(lambda v: oracle.sum_(fo(v)))(x)
        gd = ag.grad(lambda v: ag.sum_(fd(v)))(xd).to_numpy()
        assert np.allclose(gd, go, rtol=1e-12, atol=1e-14), name
        assert abs(fd(1.3 if lo > 1 else 0.4) - float(fo(np.array([1.3 if lo > 1 else 0.4]))[0])) < 1e-13, name
    x = np.asfortranarray(rng.uniform(1.0, 2.0, (5, 4)))
    xd = ag.DArray.from_numpy(x)
    for fo, fd in ((lambda v: oracle.clamp(v, 1.3, 1.7), lambda v: ag.clamp(v, 1.3, 1.7)),
                   (lambda v: oracle.ldexp(v, -2), lambda v: ag.ldexp(v, -2)),
                   (lambda v: oracle.logb(v, 3.0), lambda v: ag.logb(v, 3.0)),
                   (lambda v: oracle.logb(1.7, v), lambda v: ag.logb(1.7, v))):
        assert np.allclose(fd(xd).to_numpy(), fo(x), rtol=1e-14)
        assert np.allclose(ag.grad(lambda v: ag.sum_(fd(v)))(xd).to_numpy(), oracle.grad(lambda v: oracle.sum_(fo(v)))(x),
                           rtol=1e-12)
    # second order through the rules (they are compositions of recorded primitives): d2/dx2 sec = sec (tan^2 + sec^2)
    g2 = ag.grad(lambda v: ag.sum_(ag.grad(lambda u: ag.sum_(ag.sec(u)))(v)))(xd).to_numpy()
    assert np.allclose(g2, (np.tan(x) ** 2 + 1 / np.cos(x) ** 2) / np.cos(x), rtol=1e-11)
    # exponent (@zerograd) / significand / mod2pi / rem2pi (src/math.jl:44,56,59,64)
    xs = np.asfortranarray(np.concatenate([rng.uniform(-40.0, 40.0, 37), [0.75, 1.0, 2.0, -3.999, 1e-3]]))
    xsd = ag.DArray.from_numpy(xs)
    assert np.array_equal(ag.exponent(xsd).to_numpy(), oracle.exponent(xs))
    assert np.array_equal(ag.significand(xsd).to_numpy(), oracle.significand(xs))
    assert np.allclose(ag.mod2pi(xsd).to_numpy(), oracle.mod2pi(xs), atol=1e-13)
    for r in ("nearest", "tozero", "down", "up"):
        assert np.allclose(ag.rem2pi(xsd, r).to_numpy(), oracle.rem2pi(xs, r), atol=1e-13), r
        assert abs(ag.rem2pi(7.5, r) - float(oracle.rem2pi(np.array([7.5]), r)[0])) < 1e-14
    for fo, fd in ((oracle.significand, ag.significand), (oracle.mod2pi, ag.mod2pi),
                   (lambda v: oracle.rem2pi(v, "up"), lambda v: ag.rem2pi(v, "up"))):
        go = oracle.grad(lambda v: oracle.sum_(oracle.mul(fo(v), v)))(xs)
        gd = ag.grad(lambda v: ag.sum_(ag.mul(fd(v), v)))(xsd).to_numpy()
        assert np.allclose(gd, go, rtol=1e-12, atol=1e-13)
    # sinc(0) = 1, cosc(0) = 0; rounding functions are @zerograd (ties to even for round)
    z = ag.DArray.from_numpy(np.array([0.0, 0.5, -1.5, 2.5, -0.2]))
    assert np.array_equal(ag.sinc(z).to_numpy()[:1], [1.0]) and np.array_equal(ag.cosc(z).to_numpy()[:1], [0.0])
    assert np.array_equal(ag.round_(z).to_numpy(), [0.0, 0.0, -2.0, 2.0, -0.0])
    assert np.array_equal(ag.floor(z).to_numpy(), np.floor(z.to_numpy())) and np.array_equal(ag.ceil(z).to_numpy(), np.ceil(z.to_numpy()))
    assert np.array_equal(ag.trunc(z).to_numpy(), np.trunc(z.to_numpy()))
    assert ag.grad(lambda v: ag.sum_(ag.mul(v, ag.floor(v))))(z).to_numpy().tolist() == np.floor(z.to_numpy()).tolist()


def test_parametrised_special_functions_host_rules(hostlib, oracle):
    """besselj / bessely / polygamma / invdigamma: the host rules on the numpy stand-in against the oracle's."""
    ag = hostlib
    rng = np.random.default_rng(23)
    x = rng.uniform(0.5, 9.0, (6, 5))
    xd = ag.DArray.from_numpy(np.asfortranarray(x))
    for name, m in (("besselj", 2), ("bessely", 1), ("polygamma", 1), ("polygamma", 3), ("besseli", 2), ("besselk", 1), ("besselk", 0)):
        go = oracle.grad(lambda v: oracle.sum_(getattr(oracle, name)(m, v)))(np.asfortranarray(x))
        gd = ag.grad(lambda v: ag.sum_(getattr(ag, name)(m, v)))(xd).to_numpy()
        assert np.allclose(gd, go, rtol=1e-10, atol=1e-12), (name, m)
    xs = rng.uniform(-3.0, 2.5, (6, 5))
    xsd = ag.DArray.from_numpy(np.asfortranarray(xs))
    for name in ("airyai", "airyaiprime", "airybi", "airybiprime", "dawson", "erfi"):
        go = oracle.grad(lambda v: oracle.sum_(getattr(oracle, name)(v)))(np.asfortranarray(xs))
        gd = ag.grad(lambda v: ag.sum_(getattr(ag, name)(v)))(xsd).to_numpy()
        assert np.allclose(gd, go, rtol=1e-12, atol=1e-14), name
        assert getattr(ag, name)(0.7) == getattr(oracle, name)(0.7)          # host Numbers
    y = rng.uniform(-2.0, 3.0, (4, 3))
    go = oracle.grad(lambda v: oracle.sum_(oracle.invdigamma(v)))(np.asfortranarray(y))
    gd = ag.grad(lambda v: ag.sum_(ag.invdigamma(v)))(ag.DArray.from_numpy(np.asfortranarray(y))).to_numpy()
    assert np.allclose(gd, go, rtol=1e-8, atol=1e-10)


def test_linear_algebra_helpers_host_rules(hostlib, oracle):
    """diag / diagm / tr / tril / triu / kron (src/linearalgebra.jl:19-20,86,90,92,176-186): values against numpy,
    gradients against the oracle's (which carry the reference's rule expressions)."""
    rng = np.random.default_rng(47)
    sq = np.asfortranarray(rng.standard_normal((6, 6)).astype(np.float64))
    rect = np.asfortranarray(rng.standard_normal((4, 7)).astype(np.float64))
    sqd, rectd = hostlib.DArray.from_numpy(sq), hostlib.DArray.from_numpy(rect)
    for k in (0, 1, -2, 5):
        assert np.array_equal(hostlib.diag(sqd, k).to_numpy(), np.diag(sq, k)), k
        assert np.array_equal(hostlib.diag(rectd, k).to_numpy(), np.diag(rect, k)), k
    v = rng.standard_normal(5).astype(np.float64)
    vd = hostlib.DArray.from_numpy(v)
    for k in (0, 2, -1):
        assert np.array_equal(hostlib.diagm(vd, k).to_numpy(), np.diag(v, k)), k
    assert np.array_equal(hostlib.tril(rectd).to_numpy(), np.tril(rect)) and np.array_equal(hostlib.triu(rectd).to_numpy(), np.triu(rect))
    assert abs(float(hostlib.value(hostlib.tr(sqd))) - np.trace(sq.astype(np.float64))) <= 1e-12
    a, b = np.asfortranarray(rng.standard_normal((3, 4)).astype(np.float64)), np.asfortranarray(rng.standard_normal((5, 2)).astype(np.float64))
    ad, bd = hostlib.DArray.from_numpy(a), hostlib.DArray.from_numpy(b)
    assert np.array_equal(hostlib.kron(ad, bd).to_numpy(), np.kron(a, b))
    w = np.asfortranarray(rng.standard_normal((6, 6)))
    wd = hostlib.DArray.from_numpy(w.astype(np.float64))
    sq64 = np.asfortranarray(sq.astype(np.float64))
    for fo, fd in ((lambda z: oracle.diag(z, 1), lambda z: hostlib.diag(z, 1)), (oracle.tr, hostlib.tr), (oracle.tril, hostlib.tril),
                   (oracle.triu, hostlib.triu)):
        def lo(z):
            r = fo(z)
            return oracle.sum_(oracle.mul(r, w[:np.shape(oracle.value(r))[0]] if np.ndim(oracle.value(r)) == 2 else (w[0, :np.size(oracle.value(r))] if np.ndim(oracle.value(r)) else 1.0)))
        def ld(z):
            r = fd(z)
            shp = hostlib.size(r)
            wt = wd if len(shp) == 2 else (hostlib.DArray.from_numpy(w[0, :shp[0]].astype(np.float64)) if len(shp) else 1.0)
            return hostlib.sum_(hostlib.mul(r, wt))
        go = oracle.grad(lo)(sq64)
        gd = hostlib.grad(ld)(sqd)
        gd = gd.full().to_numpy() if hasattr(gd, "full") else gd.to_numpy()
        assert np.max(np.abs(gd - go)) <= 1e-12 * 10, fd
    a64, b64 = np.asfortranarray(a.astype(np.float64)), np.asfortranarray(b.astype(np.float64))
    wk = np.asfortranarray(rng.standard_normal((15, 8)))
    wkd = hostlib.DArray.from_numpy(wk.astype(np.float64))
    go = oracle.grad(lambda p: oracle.sum_(oracle.mul(oracle.kron(p, b64), wk)))(a64)
    gd = hostlib.grad(lambda p: hostlib.sum_(hostlib.mul(hostlib.kron(p, bd), wkd)))(ad).to_numpy()
    assert np.max(np.abs(gd - go)) <= 1e-12 * 20
    go = oracle.grad(lambda q: oracle.sum_(oracle.mul(oracle.kron(a64, q), wk)))(b64)
    gd = hostlib.grad(lambda q: hostlib.sum_(hostlib.mul(hostlib.kron(ad, q), wkd)))(bd).to_numpy()
    assert np.max(np.abs(gd - go)) <= 1e-12 * 20


def test_inv_det_logdet_host_rules(hostlib, oracle):
    ag = hostlib
    """inv / det / logdet / logabsdet (src/linearalgebra.jl:18,28,57-58; ag_inv_det: one-launch Gauss-Jordan elimination
    with partial pivoting): values against LAPACK (numpy.linalg), gradients against the oracle's rules, a matrix that
    needs row exchanges, a singular one."""
    rng = np.random.default_rng(53)
    for n in (1, 2, 5, 33, 130):
        A = np.asfortranarray((rng.standard_normal((n, n)) + (0.0 if n < 33 else 0.2) * np.eye(n)).astype(np.float64))
        Ad, A64 = ag.DArray.from_numpy(A), A.astype(np.float64)
        scale = np.linalg.cond(A64)
        got = ag.inv(Ad).to_numpy()
        assert got.dtype == np.float64 and np.max(np.abs(got - np.linalg.inv(A64))) <= 1e-12 * scale * max(np.max(np.abs(np.linalg.inv(A64))), 1.0), n
        sign, lad = np.linalg.slogdet(A64)
        assert abs(np.linalg.det(A64)) > 1e30 or abs(ag.det(Ad) - np.linalg.det(A64)) <= 1e-12 * scale * abs(np.linalg.det(A64)), n
        l, s = ag.logabsdet(Ad)
        assert abs(l - lad) <= 1e-12 * scale * max(abs(lad), 1.0) and s == sign, n
    P = np.asfortranarray(np.array([[0.0, 2.0, 1.0], [1.0, 0.0, 3.0], [4.0, 1.0, 0.0]], dtype=np.float64))      # zero leading pivots
    assert np.allclose(ag.inv(ag.DArray.from_numpy(P)).to_numpy(), np.linalg.inv(P.astype(np.float64)), atol=1e-12 * 10)
    assert abs(ag.det(ag.DArray.from_numpy(P)) - 25.0) <= 1e-12 * 100
    S = np.asfortranarray(np.array([[1.0, 2.0], [2.0, 4.0]], dtype=np.float64))
    assert ag.det(ag.DArray.from_numpy(S)) == 0.0
    spd = rng.standard_normal((6, 6))
    spd = np.asfortranarray((spd @ spd.T + 6 * np.eye(6)).astype(np.float64))
    spdd, spd64 = ag.DArray.from_numpy(spd), np.asfortranarray(spd.astype(np.float64))
    assert abs(ag.logdet(spdd) - np.linalg.slogdet(spd64)[1]) <= 1e-12 * 100
    w = np.asfortranarray(rng.standard_normal((6, 6)))
    wd = ag.DArray.from_numpy(w.astype(np.float64))
    go = oracle.grad(lambda z: oracle.sum_(oracle.mul(oracle.inv(z), w)))(spd64)
    gd = ag.grad(lambda z: ag.sum_(ag.mul(ag.inv(z), wd)))(spdd).to_numpy()
    assert np.max(np.abs(gd - go)) <= 1e-12 * 1e3 * np.max(np.abs(go))
    for fo, fd in ((oracle.det, ag.det), (oracle.logdet, ag.logdet)):
        go = oracle.grad(fo)(spd64)
        gd = ag.grad(fd)(spdd).to_numpy()
        assert np.max(np.abs(gd - go)) <= 1e-12 * 1e3 * np.max(np.abs(go)), fd
    gd = ag.grad(lambda z: ag.logabsdet(z)[0])(spdd).to_numpy()
    assert np.max(np.abs(gd - np.linalg.inv(spd64).T)) <= 1e-12 * 1e3
    with pytest.raises(ag.DimensionMismatch):
        ag.inv(ag.DArray.from_numpy(np.asfortranarray(np.ones((2, 3), dtype=np.float64))))


def test_zerograd_value_helpers_host(hostlib):
    ag = hostlib
    """The value-touching @zerograd helpers of src/base.jl (predicates, counts, integer division, array queries): device
    results against numpy; a Value argument is unboxed and nothing is recorded."""
    x = np.asfortranarray(np.array([[1.5, -2.0, np.nan, 0.0], [np.inf, -np.inf, -0.0, 7.0], [3.25, -3.0, 2.0, -7.5]], dtype=np.float64))
    xd = ag.DArray.from_numpy(x)
    with np.errstate(invalid="ignore"):
        assert np.array_equal(ag.isnan(xd).to_numpy(), np.isnan(x).astype(np.float64)) and np.array_equal(ag.isinf(xd).to_numpy(), np.isinf(x).astype(np.float64))
        assert np.array_equal(ag.isfinite(xd).to_numpy(), np.isfinite(x).astype(np.float64))
        assert np.array_equal(ag.signbit(xd).to_numpy()[~np.isnan(x)], np.signbit(x)[~np.isnan(x)].astype(np.float64))
        assert np.array_equal(ag.isinteger(xd).to_numpy(), (np.isfinite(x) & (x == np.trunc(x))).astype(np.float64))
    assert ag.count(ag.isnan(xd)) == 1 and ag.any_(ag.isinf(xd)) and not ag.all_(ag.isfinite(xd)) and ag.all_(ag.eq(xd, xd)) is False
    a = np.asfortranarray(np.array([[7.5, -7.5, 9.0], [2.0, 3.0, -4.0]], dtype=np.float64))
    b = np.asfortranarray(np.array([[2.0, 2.0, -4.0], [0.5, 5.0, 3.0]], dtype=np.float64))
    ad, bd = ag.DArray.from_numpy(a), ag.DArray.from_numpy(b)
    assert np.array_equal(ag.div_(ad, bd).to_numpy(), np.trunc(a / b)) and np.array_equal(ag.rem(ad, bd).to_numpy(), np.fmod(a, b))
    # rem is differentiable (src/base.jl:144): dx1 = dy, dx2 = -dy .* div.(x1, x2); float / widemul pass gradients
    g1 = ag.grad(lambda v: ag.sum_(ag.rem(v, bd)))(ad).to_numpy()
    g2 = ag.grad(lambda v: ag.sum_(ag.rem(ad, v)))(bd).to_numpy()
    assert np.array_equal(g1, np.ones_like(a)) and np.array_equal(g2, -np.trunc(a / b))
    assert np.array_equal(ag.grad(lambda v: ag.sum_(ag.widemul(ag.float_(v), bd)))(ad).to_numpy(), b)
    assert np.array_equal(ag.ldiv(ad, bd).to_numpy(), b / a)
    assert np.array_equal(ag.ne(ad, bd).to_numpy(), (a != b).astype(np.float64)) and np.array_equal(ag.widemul(ad, bd).to_numpy(), a * b)
    assert ag.isequal(xd, ag.DArray.from_numpy(x.copy(order="F"))) and not ag.isequal(ad, bd) and not ag.isequal(ad, xd)
    assert ag.isapprox(ad, ag.DArray.from_numpy(np.asfortranarray(a * (1 + 1e-9 if np.float64 == np.float64 else 1 + 1e-5)))) and not ag.isapprox(ad, bd)
    assert ag.eltype(ad) == np.float64 and ag.eps(ad) == float(np.finfo(np.float64).eps) and ag.sizeof(ad) == a.nbytes and not ag.isempty(ad)
    assert ag.strides(ad) == (1, 2) and ag.stride(ad, 1) == 2 and ag.lastindex(ad) == 5 and ag.axes(ad) == (range(2), range(3))
    assert ag.similar(ad).shape == (2, 3) and not np.any(ag.zero(ad).to_numpy()) and np.all(ag.ones(ad).to_numpy() == 1)
    assert ag.typemax(ad) == np.inf and ag.typemin(ad) == -np.inf and ag.float_(ad) is ad
    big = np.asfortranarray(np.random.default_rng(59).standard_normal((257, 1031)).astype(np.float64))
    big[5, 7] = big[200, 900] = 9.0
    big[100, 3] = big[11, 1000] = -9.0
    bigd = ag.DArray.from_numpy(big)
    flat = big.reshape(-1, order="F")
    assert ag.argmax(bigd) == int(np.argmax(flat)) == 5 + 7 * 257 and ag.argmin(bigd) == int(np.argmin(flat)) == 100 + 3 * 257
    assert ag.argmax(xd) == ag.argmin(xd) == 2 * 3 + 0                   # the NaN at (0, 2) wins both, like Julia
    assert ag.argmax(ag.Param(bigd)) == 5 + 7 * 257
    # oftype / big (src/base.jl:76,365): eltype conversion on the device, the gradient comes back in x's eltype
    other = np.float32 if np.float64 == np.float64 else np.float64
    odd = np.asfortranarray(np.random.default_rng(67).standard_normal((7, 11)).astype(np.float64))          # 77 elements: tail path
    oddd = ag.DArray.from_numpy(odd)
    conv = ag.oftype(other, oddd)
    assert conv.dtype == other and np.array_equal(conv.to_numpy(), odd.astype(other)) and ag.oftype(np.float64, oddd) is oddd
    assert ag.big(oddd).dtype == np.float64 and np.array_equal(ag.big(oddd).to_numpy(), odd.astype(np.float64))
    g = ag.grad(lambda v: ag.sum_(ag.abs2(ag.oftype(other, v))))(oddd)
    assert g.dtype == np.float64 and np.allclose(g.to_numpy(), 2 * odd, rtol=1e-6)
    # unboxing: inside a tape the helpers see the Param's value and contribute no gradient
    g = ag.grad(lambda v: ag.sum_(ag.mul(v, ag.isfinite(v))))(ad).to_numpy()
    assert np.array_equal(g, np.ones_like(a))


def test_iterating_a_tracked_array_goes_through_getindex(hostlib, oracle):
    """src/iterate.jl:6-17: `for e in v` and `a, b, c = v` on a tracked array are getindex calls, so the elements are
    tracked and the gradient comes back through ungetindex."""
    ag = hostlib
    x = np.asfortranarray(np.array([1.5, -2.0, 0.5]))
    xd = ag.DArray.from_numpy(x)

    def f(v):
        a, b, c = v
        t = ag.mul(a, b)
        for e in v:
            t = ag.add(t, ag.mul(e, c))
        return t
    g = ag.grad(f)(xd)
    g = g.full().to_numpy() if hasattr(g, "full") else g.to_numpy()
    # f = a b + (a + b + c) c
    assert np.allclose(g, [x[1] + x[2], x[0] + x[2], x[0] + x[1] + 2 * x[2]])
    p = ag.Param(2.5)
    assert [e is p for e in p] == [True]


def test_mean_of_abs_and_varm_host_rules(hostlib, oracle):
    """mean(abs, x; dims), mean(abs2, x; dims), varm / stdm (src/statistics.jl:9-10,26-35) against the oracle / numpy."""
    ag = hostlib
    rng = np.random.default_rng(61)
    x = np.asfortranarray(rng.standard_normal((5, 7)))
    xd = ag.DArray.from_numpy(x)
    for dims in (None, 0, 1):
        for fo, fd in ((oracle.meanabs, ag.meanabs), (oracle.meanabs2, ag.meanabs2)):
            assert np.allclose(float(ag.value(fd(xd, dims=dims))) if dims is None else fd(xd, dims=dims).to_numpy(), fo(x, dims=dims))
            w = rng.standard_normal(np.shape(fo(x, dims=dims))) if dims is not None else 1.7
            wd = ag.DArray.from_numpy(np.asfortranarray(w)) if dims is not None else w
            go = oracle.grad(lambda v: oracle.sum_(oracle.mul(fo(v, dims=dims), w)))(x)
            gd = ag.grad(lambda v: ag.sum_(ag.mul(fd(v, dims=dims), wd)))(xd).to_numpy()
            assert np.allclose(gd, go, rtol=1e-12, atol=1e-14), (dims, fd)
    m = x.mean(axis=0, keepdims=True) + 0.1
    md = ag.DArray.from_numpy(np.asfortranarray(m))
    ref = ((x - m) ** 2).sum(axis=0, keepdims=True) / (5 - 1)
    assert np.allclose(ag.varm(xd, md, dims=0).to_numpy(), ref) and np.allclose(ag.stdm(xd, md, dims=0).to_numpy(), np.sqrt(ref))
    g = ag.grad(lambda v: ag.sum_(ag.varm(v, md, dims=0)))(xd).to_numpy()
    assert np.allclose(g, 2 * (x - m) / 4)
