"""GPU parity tests: every kernel family of libagb200.so, called through the C ABI (ctypes binding),
against the numpy oracle on the same seeded inputs.  Tolerances (BASELINE.json north_star): Float32
1e-5, Float64 1e-12, norm-wise against the float64 oracle; index bookkeeping bit-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = {np.float32: 1e-5, np.float64: 1e-12}


def relerr(got, ref):
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    if ref.size == 0:
        return 0.0
    return np.max(np.abs(got - ref)) / max(np.max(np.abs(ref)), 1e-300)


UNARY = ["neg", "abs_", "abs2", "exp", "log", "tanh", "sigm", "sqrt", "sin", "cos", "tan", "sinh", "cosh",
         "asin", "acos", "atan", "asinh", "acosh", "atanh", "exp2", "exp10", "expm1", "log2", "log10", "log1p",
         "cbrt"]
NP_F = dict(neg=lambda x: -x, abs_=np.abs, abs2=lambda x: x * x, exp=np.exp, log=np.log, tanh=np.tanh,
            sigm=lambda x: 1 / (1 + np.exp(-x)), sqrt=np.sqrt, sin=np.sin, cos=np.cos, tan=np.tan, sinh=np.sinh,
            cosh=np.cosh, asin=np.arcsin, acos=np.arccos, atan=np.arctan, asinh=np.arcsinh, acosh=np.arccosh,
            atanh=np.arctanh, exp2=np.exp2, exp10=lambda x: 10.0 ** x, expm1=np.expm1, log2=np.log2,
            log10=np.log10, log1p=np.log1p, cbrt=np.cbrt)
NP_G = dict(neg=lambda x, y: -1 + 0 * x, abs_=lambda x, y: np.sign(x), abs2=lambda x, y: 2 * x, exp=lambda x, y: y,
            log=lambda x, y: 1 / x, tanh=lambda x, y: 1 - y * y, sigm=lambda x, y: y * (1 - y),
            sqrt=lambda x, y: 1 / (2 * y), sin=lambda x, y: np.cos(x), cos=lambda x, y: -np.sin(x),
            tan=lambda x, y: 1 + y * y, sinh=lambda x, y: np.cosh(x), cosh=lambda x, y: np.sinh(x),
            asin=lambda x, y: 1 / np.sqrt(1 - x * x), acos=lambda x, y: -1 / np.sqrt(1 - x * x),
            atan=lambda x, y: 1 / (1 + x * x), asinh=lambda x, y: 1 / np.sqrt(1 + x * x),
            acosh=lambda x, y: 1 / np.sqrt(x * x - 1), atanh=lambda x, y: 1 / (1 - x * x),
            exp2=lambda x, y: y * np.log(2), exp10=lambda x, y: y * np.log(10), expm1=lambda x, y: 1 + y,
            log2=lambda x, y: 1 / (np.log(2) * x), log10=lambda x, y: 1 / (np.log(10) * x),
            log1p=lambda x, y: 1 / (1 + x), cbrt=lambda x, y: 1 / (3 * y * y))


SPECIAL = {"erf": (-1.5, 1.5), "erfc": (-1.5, 1.5), "erfcx": (-0.5, 3.0), "erfinv": (-0.9, 0.9), "erfcinv": (0.1, 1.9),
           "gamma": (0.3, 4.0), "lgamma": (0.3, 6.0), "digamma": (0.3, 20.0), "trigamma": (0.3, 20.0),
           "besselj0": (0.2, 9.0), "besselj1": (0.2, 9.0), "bessely0": (0.3, 9.0), "bessely1": (0.3, 9.0)}


@pytest.mark.parametrize("dtype,tol", [(np.float32, 2e-5), (np.float64, 1e-11)])
def test_special_functions_against_the_oracle(ag, oracle, dtype, tol):
    """src/specialfunctions.jl on device (CUDA math library + digamma / trigamma / polygamma(2,.) series) against the
    oracle's scipy.special restatement: values and the gradients of the reference's rules, also for negative
    non-integer arguments of the gamma family (reflection)."""
    rng = np.random.default_rng(77)
    for name, (lo, hi) in SPECIAL.items():
        xs = [rng.uniform(lo, hi, 4099)]
        if name in ("gamma", "lgamma", "digamma", "trigamma"):
            xs.append(-rng.uniform(0.05, 0.95, 4099) - rng.integers(0, 3, 4099))       # negative, not an integer
        for x in xs:
            x = x.astype(dtype)
            x64 = x.astype(np.float64)
            fo = getattr(oracle, name)
            yref = fo(x64)
            gref = oracle.grad(lambda v: oracle.sum_(fo(v)))(x64)
            xd = ag.DArray.from_numpy(x)
            y = getattr(ag, name)(xd).to_numpy().astype(np.float64)
            g = ag.grad(lambda v: ag.sum_(getattr(ag, name)(v)))(xd).to_numpy().astype(np.float64)
            # element-wise, relative to the size of the value and of the terms that cancel near zeros
            yscale = np.maximum(np.abs(yref), 1.0) if name != "lgamma" else np.maximum(np.abs(yref), np.abs(np.log(np.abs(x64))) + 1)
            assert np.max(np.abs(y - yref) / yscale) <= tol * (30 if x[0] < 0 else 4), (name, "value", x[0] < 0)
            gscale = np.maximum(np.abs(gref), 1.0 + np.abs(yref))
            assert np.max(np.abs(g - gref) / gscale) <= tol * (60 if x[0] < 0 else 8), (name, "gradient", x[0] < 0)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("n", [1, 3, 1023, 4096, 100003])
def test_unary_fwd_bwd(ag, dtype, n):
    rng = np.random.default_rng(n)
    x = rng.uniform(0.15, 0.85, n).astype(dtype)
    dy = rng.standard_normal(n).astype(dtype)
    for name in UNARY:
        xx = x + 1.5 if name == "acosh" else x
        x64, dy64 = xx.astype(np.float64), dy.astype(np.float64)
        f = getattr(ag, name)
        xd = ag.DArray.from_numpy(xx)
        y = f(xd)
        yref = NP_F[name](x64)
        assert relerr(y.to_numpy(), yref) <= TOL[dtype], (name, "fwd")
        # gradient through the tape: d/dx sum(dy .* f(x)) = dy .* g(x,y)
        p = ag.Param(xd)
        dyd = ag.DArray.from_numpy(dy)
        t = ag.diff(lambda: ag.sum_(ag.mul(dyd, f(p))))
        g = ag.grad(t, p).to_numpy()
        gref = dy64 * NP_G[name](x64, yref)
        assert relerr(g, gref) <= 4 * TOL[dtype], (name, "bwd")


def test_unary_misaligned_view(ag):
    """A reshape/slice view that is not 16-byte aligned must take the scalar path and stay correct."""
    rng = np.random.default_rng(0)
    x = rng.standard_normal(1001).astype(np.float32)
    xd = ag.DArray.from_numpy(x)
    v = ag.DArray(xd.buf, xd.ptr + 4, (1000,), xd.dtype)
    assert relerr(ag.tanh(v).to_numpy(), np.tanh(x[1:].astype(np.float64))) <= 1e-6
    assert relerr(ag.add(v, v).to_numpy(), 2 * x[1:].astype(np.float64)) == 0
    assert relerr(ag.sum_(v).to_numpy(), x[1:].astype(np.float64).sum()) <= 1e-6


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_relu_ties(ag, dtype):
    """max.(0,x): relu'(0) = 1 (src/math.jl:54)."""
    x = np.array([-1.0, -0.0, 0.0, 2.0, 5.0, -3.0, 0.0, 1.0], dtype)
    p = ag.Param(ag.DArray.from_numpy(x))
    t = ag.diff(lambda: ag.sum_(ag.relu(p)))
    assert np.array_equal(ag.grad(t, p).to_numpy(), np.array([0, 1, 1, 1, 1, 0, 1, 1], dtype))
    a = ag.Param(ag.DArray.from_numpy(np.array([1.0, 2.0, 3.0], dtype)))
    b = ag.Param(ag.DArray.from_numpy(np.array([1.0, 5.0, 0.0], dtype)))
    t = ag.diff(lambda: ag.sum_(ag.max_(a, b)))
    assert np.array_equal(ag.grad(t, a).to_numpy(), [1, 0, 1])
    assert np.array_equal(ag.grad(t, b).to_numpy(), [1, 1, 0])


SHAPES = [((2, 3, 5), (1, 3, 5)), ((2, 3, 5, 4), (2, 3, 1, 4)), ((64, 100), (64,)), ((10, 100), (1, 100)),
          ((7, 9), (1, 1)), ((3,), (3, 5)), ((128, 257), (128, 257)), ((4, 1, 6), (1, 5, 1)), ((1000,), (1000,)),
          ((5, 3, 2, 4, 3), (5, 3, 1, 4, 3))]


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("sa,sb", SHAPES)
def test_binary_broadcast_fwd_bwd(ag, oracle, dtype, sa, sb):
    rng = np.random.default_rng(len(sa) * 100 + len(sb))
    a = rng.uniform(0.5, 2.0, sa).astype(dtype)
    b = rng.uniform(0.5, 2.0, sb).astype(dtype)
    for name in ("add", "sub", "mul", "div", "max_", "min_", "pow_", "atan2", "hypot"):
        if not hasattr(oracle, name):
            continue
        po = [oracle.Param(np.asfortranarray(a.astype(np.float64))), oracle.Param(np.asfortranarray(b.astype(np.float64)))]
        to = oracle.diff(lambda: oracle.sumabs2(getattr(oracle, name)(po[0], po[1])))
        pd = [ag.Param(ag.DArray.from_numpy(a)), ag.Param(ag.DArray.from_numpy(b))]
        td = ag.diff(lambda: ag.sumabs2(getattr(ag, name)(pd[0], pd[1])))
        lo, ld = float(oracle.value(to)), float(ag.value(td))
        assert abs(lo - ld) <= 4 * TOL[dtype] * abs(lo), (name, sa, sb)
        for q, r in zip(pd, po):
            assert relerr(ag.grad(td, q).to_numpy(), oracle.grad(to, r)) <= 4 * TOL[dtype], (name, sa, sb)


def test_binary_scalar_operands(ag):
    x = np.linspace(-2, 2, 37).astype(np.float32)
    xd = ag.DArray.from_numpy(x)
    assert relerr(ag.mul(0.1, xd).to_numpy(), np.float32(0.1) * x) <= 1e-7
    assert relerr(ag.sub(1, xd).to_numpy(), 1 - x) == 0
    assert relerr(ag.div(xd, 4).to_numpy(), x / 4) == 0
    assert np.array_equal(ag.relu(xd).to_numpy(), np.maximum(x, 0))
    s = ag.sum_(xd)                                    # device scalar
    assert s.shape == ()
    assert relerr(ag.mul(xd, s).to_numpy(), x * x.sum(dtype=np.float64)) <= 1e-5


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("n", [1, 5, 4096, 4097, 1 << 20, (1 << 22) + 3])
def test_sum_all(ag, dtype, n):
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n).astype(dtype)
    xd = ag.DArray.from_numpy(x)
    x64 = x.astype(np.float64)
    for f, ref in ((ag.sum_, x64.sum()), (ag.sumabs2, (x64 * x64).sum()), (ag.sumabs, np.abs(x64).sum())):
        got = float(f(xd))
        scale = np.abs(x64).sum() if f is ag.sum_ else abs(ref)
        assert abs(got - ref) <= TOL[dtype] * max(scale, 1e-30), (f, n)


RED = [((10, 1000), 0), ((10, 1000), 1), ((64, 5000), 1), ((64, 5000), 0), ((1, 777), 1), ((300, 50), 1),
       ((1000, 33), 1), ((1000, 33), 0), ((100, 7, 13), 1), ((2, 3, 5), (0, 2)), ((2, 3, 5), (0, 1)),
       ((6, 4, 5, 3), (1, 2)), ((5000, 3), 0), ((3, 100000), 1), ((100000, 3), 0), ((17, 1), 1),
       # tiled strided form (inner > 256, short reduced extent): the LSTM bias gradients
       ((512, 256), 1), ((640, 256), 1), ((513, 17), 1), ((257, 64, 3), 1), ((300, 16), 1), ((4099, 1000), 1)]


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("shape,dims", RED)
def test_sum_dims(ag, dtype, shape, dims):
    rng = np.random.default_rng(sum(shape))
    x = rng.standard_normal(shape).astype(dtype)
    xd = ag.DArray.from_numpy(x)
    ax = dims if isinstance(dims, tuple) else (dims,)
    ref = x.astype(np.float64).sum(axis=ax, keepdims=True)
    got = ag.sum_(xd, dims=dims).to_numpy()
    absref = np.abs(x.astype(np.float64)).sum(axis=ax, keepdims=True)
    assert got.shape == ref.shape
    assert np.max(np.abs(got - ref) / absref) <= TOL[dtype]
    got2 = ag.sumabs2(xd, dims=dims).to_numpy()
    assert relerr(got2, (x.astype(np.float64) ** 2).sum(axis=ax, keepdims=True)) <= TOL[dtype]


GEMM = [(64, 784, 1000, 0, 0), (64, 1000, 784, 0, 1), (64, 10, 1000, 1, 0), (10, 64, 3000, 0, 1),
        (1, 13, 506, 0, 0), (1, 506, 13, 0, 1), (130, 70, 33, 0, 0), (130, 70, 33, 1, 1), (5, 1, 7, 0, 0),
        (512, 640, 256, 0, 0), (512, 256, 640, 0, 1), (640, 512, 256, 1, 0), (128, 128, 4096, 0, 1),
        (10, 5000, 64, 0, 1), (64, 4100, 64, 0, 1), (10, 64, 4096, 0, 0), (64, 10, 5000, 1, 0), (33, 17, 777, 0, 0),
        (64, 784, 4096, 0, 0), (64, 8192, 784, 0, 1), (132, 72, 2000, 1, 1),
        # streaming skinny forms: S1a (few rows, K % 4 == 0), S1b (K <= 16, many rows), S2 (huge contraction)
        (7, 20, 1000, 0, 0), (16, 64, 515, 1, 0), (12, 16, 65536, 0, 0), (40, 3, 777, 0, 0), (33, 16, 600, 0, 0),
        (64, 10, 65536, 1, 0), (3, 2051, 12, 0, 1), (16, 70001, 64, 0, 1), (10, 65536, 64, 0, 1),
        # warp-per-output form (tiny outputs): housing's dw = dy*x', dot products, all four transposes
        (3, 100, 5, 1, 0), (7, 64, 9, 1, 1), (1, 2000, 1, 0, 0), (16, 40, 64, 0, 0), (13, 506, 1, 0, 1), (2, 33, 3, 0, 1)]


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("M,K,N,ta,tb", GEMM)
def test_gemm(ag, dtype, M, K, N, ta, tb):
    """op(A) (MxK) * op(B) (KxN); exercises NN forward, NT (dy*x') and TN (w'*dy) of src/base.jl:26."""
    rng = np.random.default_rng(M * 7 + N)
    a = rng.standard_normal((K, M) if ta else (M, K)).astype(dtype)
    b = rng.standard_normal((N, K) if tb else (K, N)).astype(dtype)
    ad, bd = ag.DArray.from_numpy(a), ag.DArray.from_numpy(b)
    A = ag.adjoint(ad) if ta else ad
    Bm = ag.adjoint(bd) if tb else bd
    c = ag.matmul(A, Bm).to_numpy()
    a64, b64 = a.astype(np.float64), b.astype(np.float64)
    ref = (a64.T if ta else a64) @ (b64.T if tb else b64)
    # norm-wise against the magnitude of the accumulated products
    mag = (np.abs(a64.T if ta else a64) @ np.abs(b64.T if tb else b64)).max()
    assert c.shape == ref.shape
    assert np.max(np.abs(c - ref)) <= TOL[dtype] * mag, ag.gemm_engine()


def test_gemm_vector_and_errors(ag):
    rng = np.random.default_rng(1)
    a, v = rng.standard_normal((5, 7)), rng.standard_normal(7)
    r = ag.matmul(ag.DArray.from_numpy(a), ag.DArray.from_numpy(v))
    assert r.shape == (5,) and relerr(r.to_numpy(), a @ v) <= 1e-14
    with pytest.raises(ag.DimensionMismatch):
        ag.matmul(ag.DArray.from_numpy(a), ag.DArray.from_numpy(a))
    with pytest.raises(ag.DimensionMismatch):
        ag.add(ag.DArray.from_numpy(a), ag.DArray.from_numpy(v))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_scatter_add_and_gather(ag, oracle, dtype):
    rng = np.random.default_rng(3)
    E, V, Bn, T = 32, 50, 64, 5
    W = rng.standard_normal((E, V)).astype(dtype)
    Wd = ag.DArray.from_numpy(W)
    ids = [rng.integers(0, V, Bn) for _ in range(T)]
    ids[0][:8] = 7                                     # heavy repeats
    vals = [rng.standard_normal((E, Bn)).astype(dtype) for _ in range(T)]
    # gather
    assert np.array_equal(ag.getindex(Wd, slice(None), ids[0]).to_numpy(), W[:, ids[0]])
    # Sparse with T blocks -> one batched scatter-add (column-block path)
    sp = ag.Sparse(Wd, [ag.DArray.from_numpy(v) for v in vals], [(slice(None), i) for i in ids])
    ref = oracle.full(oracle.Sparse(np.asfortranarray(W.astype(np.float64)),
                                    [np.asfortranarray(v.astype(np.float64)) for v in vals],
                                    [(slice(None), i) for i in ids]))
    got = sp.full().to_numpy()
    assert relerr(got, ref) <= TOL[dtype]
    # determinism: bit-identical on repeat
    assert np.array_equal(got, sp.full().to_numpy())
    # untouched destinations stay exactly zero; touched set is bit-exact
    assert np.array_equal(got != 0, ref != 0)
    # generic (element-granular) path with mixed index kinds
    x = rng.standard_normal((6, 7)).astype(dtype)
    for idx in [([1, 1, 3], [0, 2]), ([5, 5, 5],), (range(0, 4, 2), slice(None)), (3, 4)]:
        go = oracle.grad(lambda v: oracle.sumabs2(oracle.getindex(v, *idx)))(np.asfortranarray(x.astype(np.float64)))
        gd = ag.grad(lambda v: ag.sumabs2(ag.getindex(v, *idx)))(ag.DArray.from_numpy(x))
        assert relerr(gd.to_numpy(), go) <= TOL[dtype], idx


def test_scatter_out_of_range_is_an_error(ag):
    z = ag.zeros((4, 4), np.float32)
    with pytest.raises(IndexError):
        ag.Sparse(z, [ag.DArray.from_numpy(np.ones((4, 1), np.float32))], [(slice(None), [9])]).full()


def test_cat_transpose_copy(ag):
    rng = np.random.default_rng(4)
    a, b = rng.standard_normal((128, 33)).astype(np.float32), rng.standard_normal((512, 33)).astype(np.float32)
    ad, bd = ag.DArray.from_numpy(a), ag.DArray.from_numpy(b)
    assert np.array_equal(ag.vcat(ad, bd).to_numpy(), np.vstack([a, b]))
    assert np.array_equal(ag.hcat(ad, ad).to_numpy(), np.hstack([a, a]))
    assert np.array_equal(ag.adjoint(bd).to_numpy(), b.T)
    assert np.array_equal(ag.copy(ad).to_numpy(), a)
    assert np.array_equal(ag.cat1d(ad, bd).to_numpy(), np.concatenate([a.reshape(-1, order="F"), b.reshape(-1, order="F")]))


def test_axpy_and_fill(ag):
    x = np.arange(1001, dtype=np.float32)
    y = np.ones(1001, dtype=np.float32)
    yd = ag.DArray.from_numpy(y)
    ag.axpy(-0.5, ag.DArray.from_numpy(x), yd)
    assert np.array_equal(yd.to_numpy(), y - np.float32(0.5) * x)
    assert np.array_equal(ag.zeros((3, 5), np.float64).to_numpy(), np.zeros((3, 5)))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_extrema_prod_permute_stats(ag, oracle, dtype):
    """maximum / minimum / prod / permutedims / mean / var / std / norm on device against the oracle."""
    rng = np.random.default_rng(13)
    x = rng.uniform(0.5, 1.5, (6, 7, 33)).astype(dtype)
    tol = TOL[dtype]
    for name, kw in [("maximum", {}), ("maximum", {"dims": 0}), ("maximum", {"dims": 1}), ("minimum", {"dims": (0, 2)}),
                     ("minimum", {"dims": 2}), ("prod", {"dims": 0}), ("prod", {"dims": 1}), ("mean", {"dims": 2})]:
        po = oracle.Param(np.asfortranarray(x.astype(np.float64)))
        to = oracle.diff(lambda: oracle.sumabs2(getattr(oracle, name)(po, **kw)))
        pd = ag.Param(ag.DArray.from_numpy(x))
        td = ag.diff(lambda: ag.sumabs2(getattr(ag, name)(pd, **kw)))
        lo, ld = float(oracle.value(to)), float(ag.value(td))
        assert abs(lo - ld) <= 8 * tol * abs(lo), (name, kw)
        assert relerr(ag.grad(td, pd).to_numpy(), oracle.grad(to, po)) <= 8 * tol, (name, kw)
    xd = ag.DArray.from_numpy(x)
    for perm in [(1, 0, 2), (2, 0, 1), (2, 1, 0)]:
        assert np.array_equal(ag.permutedims(xd, perm).to_numpy(), np.transpose(x, perm))
    x64 = x.astype(np.float64)
    assert abs(float(ag.var(xd)) - x64.var(ddof=1)) <= 20 * tol * x64.var(ddof=1)
    assert relerr(ag.std(xd, dims=1).to_numpy(), x64.std(axis=1, ddof=1, keepdims=True)) <= 20 * tol
    assert abs(float(ag.norm(xd)) - np.linalg.norm(x64)) <= tol * np.linalg.norm(x64)
    assert ag.gradcheck(lambda z: ag.permutedims(z, (1, 0)), ag.DArray.from_numpy(x64[:, :, 0]))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("n", [1, 5, 4099, 1 << 20])
def test_fused_expression_evaluator(ag, dtype, n):
    """ag_ewise_vm: user-defined `@primitive f(x),dy,y (expr)` (src/macros.jl:88-105; sigm of
    examples/charlm.jl:46-47) as ONE fused pass; value, gradient and second derivative against closed forms."""
    E = ag.E
    sigm2 = ag.primitive_expr("sigm2", lambda x: 1 / (1 + E.exp(-x)), lambda dy, y, x: dy * y * (1 - y))
    mix = ag.primitive_expr("mix", lambda a, b: E.max(a, 0.25) * E.tanh(b) - a ** 2 / (1 + E.abs2(b)),
                            lambda dy, y, a, b: dy * (E.gt(a, 0.25) * E.tanh(b) - 2 * a / (1 + E.abs2(b))),
                            lambda dy, y, a, b: dy * (E.max(a, 0.25) * (1 - E.tanh(b) ** 2)
                                                      + 2 * b * a ** 2 / (1 + E.abs2(b)) ** 2))
    rng = np.random.default_rng(n)
    x, b = rng.standard_normal(n).astype(dtype), rng.standard_normal(n).astype(dtype)
    x64, b64 = x.astype(np.float64), b.astype(np.float64)
    xd, bd = ag.DArray.from_numpy(x), ag.DArray.from_numpy(b)
    s = 1 / (1 + np.exp(-x64))
    k0 = ag.launch_count()
    y = sigm2(xd)
    assert ag.launch_count() - k0 == 1
    assert relerr(y.to_numpy(), s) < TOL[dtype]
    assert relerr(ag.grad(lambda v: ag.sum_(sigm2(v)))(xd).to_numpy(), s * (1 - s)) < TOL[dtype]
    g1 = ag.grad(lambda v: ag.sum_(sigm2(v)))
    assert relerr(ag.grad(lambda v: ag.sum_(g1(v)))(xd).to_numpy(), s * (1 - s) * (1 - 2 * s)) < 4 * TOL[dtype]
    ref = np.maximum(x64, 0.25) * np.tanh(b64) - x64 ** 2 / (1 + b64 * b64)
    assert relerr(mix(xd, bd).to_numpy(), ref) < TOL[dtype]
    ga, gb = ag.grad(lambda u: ag.sum_(mix(u, bd)))(xd), ag.grad(lambda v: ag.sum_(mix(xd, v)))(bd)
    assert relerr(ga.to_numpy(), (x64 > 0.25) * np.tanh(b64) - 2 * x64 / (1 + b64 * b64)) < TOL[dtype]
    assert relerr(gb.to_numpy(), np.maximum(x64, 0.25) * (1 - np.tanh(b64) ** 2)
                  + 2 * b64 * x64 ** 2 / (1 + b64 * b64) ** 2) < TOL[dtype]
    # scalar operand as an immediate and a broadcasting operand (composite path) agree with the fused path
    assert relerr(mix(xd, 0.5).to_numpy(), np.maximum(x64, 0.25) * np.tanh(0.5) - x64 ** 2 / 1.25) < TOL[dtype]


def test_fused_expression_evaluator_rejects_bad_programs(ag):
    import ctypes as C
    lib = ag._lib.lib
    x = ag.DArray.from_numpy(np.ones(8, np.float32))
    out = ag.DArray.empty((8,), np.float32)
    ptrs, modes, imms = (C.c_void_p * 1)(x.ptr), (C.c_int32 * 1)(0), (C.c_double * 1)(0.0)
    consts = (C.c_double * 1)(0.0)
    for prog in ([2], [0, 0], [0 | (3 << 8)], [1 | (5 << 8)], [0, 15 + 200]):   # underflow, 2 results, bad input/const/op
        code = (C.c_int32 * len(prog))(*prog)
        st = lib.ag_ewise_vm(out.dt, code, len(prog), consts, 1, 1, ptrs, modes, imms, out.ptr, 8)
        assert st != 0, prog


@pytest.mark.parametrize("dtype,tol", [(np.float32, 2e-5), (np.float64, 2e-11)])
def test_special_functions_with_an_integer_parameter(ag, dtype, tol):
    """besselj(m, x), bessely(m, x) of integer order, polygamma(m, x), invdigamma(x) (ag_special_param) against
    scipy.special, and their gradients (src/specialfunctions.jl:20,25,43,59) against the closed forms."""
    import scipy.special as sp
    rng = np.random.default_rng(17)
    x = rng.uniform(0.3, 25.0, 4001).astype(dtype)
    xd = ag.DArray.from_numpy(x)
    x64 = x.astype(np.float64)

    def close(got, ref, t=tol):
        return np.max(np.abs(got.astype(np.float64) - ref)) <= t * max(np.max(np.abs(ref)), 1.0)
    for m in (-1, 0, 1, 2, 5):
        assert close(ag.besselj(m, xd).to_numpy(), sp.jv(m, x64)), ("besselj", m)
    for m in (0, 1, 3):
        assert close(ag.bessely(m, xd).to_numpy(), sp.yv(m, x64), tol * 20), ("bessely", m)
    for m in (0, 1, 2, 3, 6):
        ref = sp.polygamma(m, x64)
        got = ag.polygamma(m, xd).to_numpy().astype(np.float64)
        floor = 1.0 if m <= 1 else 1e-30                  # digamma crosses zero: absolute there, relative elsewhere
        assert np.max(np.abs(got - ref) / np.maximum(np.abs(ref), floor)) <= (5e-6 if dtype == np.float32 else 1e-10), ("polygamma", m)
    y = rng.uniform(-3.0, 4.0, 2001).astype(dtype)
    inv = ag.invdigamma(ag.DArray.from_numpy(y)).to_numpy().astype(np.float64)
    assert np.max(np.abs(sp.digamma(inv) - y.astype(np.float64))) <= (2e-5 if dtype == np.float32 else 1e-11)
    # gradients
    g = ag.grad(lambda v: ag.sum_(ag.besselj(2, v)))(xd).to_numpy()
    assert close(g, (sp.jv(1, x64) - sp.jv(3, x64)) / 2)
    g = ag.grad(lambda v: ag.sum_(ag.polygamma(1, v)))(xd).to_numpy().astype(np.float64)
    ref = sp.polygamma(2, x64)
    assert np.max(np.abs(g - ref) / np.maximum(np.abs(ref), 1e-30)) <= (5e-6 if dtype == np.float32 else 1e-10)
    yd = ag.DArray.from_numpy(y)
    g = ag.grad(lambda v: ag.sum_(ag.invdigamma(v)))(yd).to_numpy().astype(np.float64)
    assert np.max(np.abs(g * sp.polygamma(1, inv) - 1.0)) <= (1e-4 if dtype == np.float32 else 1e-9)
    with pytest.raises(ag.UnsupportedOnDevice):
        ag.besselj(0.5, xd)


@pytest.mark.parametrize("dtype,tol", [(np.float32, 2e-5), (np.float64, 1e-12)])
def test_math_rest_on_device(ag, oracle, dtype, tol):
    """The remainder of src/math.jl on the device (composed forwards, reference rules): values and gradients against
    the float64 oracle; sinc(0) / cosc(0); clamp, ldexp, log(b, x); the @zerograd rounding functions bit-exact."""
    from gradcheck_catalogue import MATH_REST, MATH_REST_DOMAINS
    rng = np.random.default_rng(43)
    for name in MATH_REST:
        lo, hi = MATH_REST_DOMAINS.get(name, (1.2, 2.0))
        x = np.asfortranarray(rng.uniform(lo, hi, (33, 17)).astype(dtype))
        xd, x64 = ag.DArray.from_numpy(x), np.asfortranarray(x.astype(np.float64))
        fo, fd = getattr(oracle, name), getattr(ag, name)
        ref = fo(x64)
        got = fd(xd).to_numpy()
        assert got.dtype == dtype and np.max(np.abs(got - ref)) <= tol * max(np.max(np.abs(ref)), 1.0), name
        go = oracle.grad(lambda v: oracle.sum_(fo(v)))(x64)
        gd = ag.grad(lambda v: ag.sum_(fd(v)))(xd).to_numpy()
        assert np.max(np.abs(gd - go)) <= 5 * tol * max(np.max(np.abs(go)), 1.0), name
    x = np.asfortranarray(rng.uniform(1.0, 2.0, (33, 17)).astype(dtype))
    xd, x64 = ag.DArray.from_numpy(x), np.asfortranarray(x.astype(np.float64))
    lo_, hi_ = float(dtype(1.3)), float(dtype(1.7))
    for fo, fd in ((lambda v: oracle.clamp(v, lo_, hi_), lambda v: ag.clamp(v, lo_, hi_)),
                   (lambda v: oracle.ldexp(v, -2), lambda v: ag.ldexp(v, -2)),
                   (lambda v: oracle.logb(v, 3.0), lambda v: ag.logb(v, 3.0)),
                   (lambda v: oracle.logb(1.7, v), lambda v: ag.logb(1.7, v))):
        ref, go = fo(x64), oracle.grad(lambda v: oracle.sum_(fo(v)))(x64)
        assert np.max(np.abs(fd(xd).to_numpy() - ref)) <= tol * max(np.max(np.abs(ref)), 1.0)
        assert np.max(np.abs(ag.grad(lambda v: ag.sum_(fd(v)))(xd).to_numpy() - go)) <= 5 * tol * max(np.max(np.abs(go)), 1.0)
    xs = np.concatenate([rng.uniform(-40.0, 40.0, 500), [0.75, 1.0, 2.0, -3.999, 1e-3]]).astype(dtype)
    xsd, xs64 = ag.DArray.from_numpy(xs), xs.astype(np.float64)
    assert np.array_equal(ag.exponent(xsd).to_numpy(), oracle.exponent(xs)) and np.array_equal(ag.significand(xsd).to_numpy(), oracle.significand(xs))
    assert np.max(np.abs(ag.mod2pi(xsd).to_numpy() - oracle.mod2pi(xs64))) <= 10 * tol
    for r in ("nearest", "tozero", "down", "up"):
        assert np.max(np.abs(ag.rem2pi(xsd, r).to_numpy() - oracle.rem2pi(xs64, r))) <= 10 * tol, r
    gd = ag.grad(lambda v: ag.sum_(ag.mul(ag.significand(v), v)))(xsd).to_numpy()
    go = oracle.grad(lambda v: oracle.sum_(oracle.mul(oracle.significand(v), v)))(xs64)
    assert np.max(np.abs(gd - go)) <= 10 * tol * np.max(np.abs(go))
    z = np.array([0.0, 0.5, -1.5, 2.5, -0.2, 1e-3, 7.49, -7.51], dtype=dtype)
    zd = ag.DArray.from_numpy(z)
    assert ag.sinc(zd).to_numpy()[0] == 1.0 and ag.cosc(zd).to_numpy()[0] == 0.0
    assert np.allclose(ag.sinc(zd).to_numpy(), np.sinc(z.astype(np.float64)), atol=tol)
    for f, r in ((ag.floor, np.floor), (ag.ceil, np.ceil), (ag.round_, np.rint), (ag.trunc, np.trunc)):
        assert np.array_equal(f(zd).to_numpy(), r(z))


@pytest.mark.parametrize("dtype,tol", [(np.float32, 3e-6), (np.float64, 2e-12)])
def test_modified_bessel_functions_of_integer_order(ag, dtype, tol):
    """besseli(m, x), besselk(m, x) for integer m (src/specialfunctions.jl:13,24; ag_special_param ops 10, 11) against
    scipy.special, with the reference's rules (I(m-1) + I(m+1))/2 and -(K(m-1) + K(m+1))/2."""
    import scipy.special as sp
    rng = np.random.default_rng(31)
    x = np.concatenate([rng.uniform(0.01, 60.0, 4000), np.logspace(-5, 0, 200), [7.8, 7.9]]).astype(dtype)
    xd, x64 = ag.DArray.from_numpy(x), x.astype(np.float64)

    def rel(got, ref):
        ok = np.isfinite(ref) & (np.abs(ref) > 1e-30) & (np.abs(ref) < 1e30)
        return np.max(np.abs(got.astype(np.float64)[ok] - ref[ok]) / np.abs(ref[ok]))
    for m in (0, 1, 2, 5, -3, 12):
        assert rel(ag.besseli(m, xd).to_numpy(), sp.iv(m, x64)) <= tol, ("besseli", m)
        assert rel(ag.besselk(m, xd).to_numpy(), sp.kv(m, x64)) <= tol, ("besselk", m)
    xn = -x[:500]
    assert rel(ag.besseli(3, ag.DArray.from_numpy(xn)).to_numpy(), sp.iv(3, xn.astype(np.float64))) <= tol
    bad = ag.besselk(1, ag.DArray.from_numpy(np.array([-1.0, 0.0], dtype=dtype))).to_numpy()
    assert np.isnan(bad[0]) and np.isinf(bad[1])
    xg = rng.uniform(0.3, 6.0, 1500).astype(dtype)
    xgd, xg64 = ag.DArray.from_numpy(xg), xg.astype(np.float64)
    g = ag.grad(lambda v: ag.sum_(ag.besseli(2, v)))(xgd).to_numpy()
    assert rel(g, (sp.iv(1, xg64) + sp.iv(3, xg64)) / 2) <= 10 * tol
    g = ag.grad(lambda v: ag.sum_(ag.besselk(0, v)))(xgd).to_numpy()
    assert rel(g, -sp.kv(1, xg64)) <= 10 * tol
    with pytest.raises(ag.UnsupportedOnDevice):
        ag.besselk(0.5, xd)


@pytest.mark.parametrize("dtype,tol", [(np.float32, 3e-6), (np.float64, 2e-12)])
def test_airy_dawson_erfi_on_device(ag, dtype, tol):
    """airyai / airyaiprime / airybi / airybiprime, dawson, erfi (src/specialfunctions.jl:5-8,33,39; ag_special_param,
    functions CUDA's math library does not have) against scipy.special over every branch of the device code (series,
    quadrature on both sides, asymptotic expansion), the Wronskian Ai Bi' - Ai' Bi = 1/pi, and the reference's rules."""
    import scipy.special as sp
    rng = np.random.default_rng(29)
    x = np.concatenate([rng.uniform(-40.0, 12.0, 6000), np.linspace(-2.5, 2.5, 501), [0.0, -2.0, 2.0, 8.5, 8.6],
                        rng.uniform(12.0, 25.0, 200)]).astype(dtype)
    xd = ag.DArray.from_numpy(x)
    x64 = x.astype(np.float64)
    ref = sp.airy(x64)
    got = [getattr(ag, n)(xd).to_numpy().astype(np.float64) for n in ("airyai", "airyaiprime", "airybi", "airybiprime")]
    z = np.maximum(np.abs(x64), 1.0)
    env = [np.where(x64 < 0, z ** -0.25, np.abs(ref[0])), np.where(x64 < 0, z ** 0.25, np.abs(ref[1])),
           np.where(x64 < 0, z ** -0.25, np.abs(ref[2])), np.where(x64 < 0, z ** 0.25, np.abs(ref[3]))]
    for k in range(4):                       # relative to the function for x >= 0, to the oscillation envelope for x < 0
        err = np.abs(got[k] - ref[k]) / np.maximum(env[k], 1e-300 if dtype == np.float64 else 1e-38)
        ok = np.isfinite(ref[k]) & (np.abs(ref[k]) > (0 if dtype == np.float64 else 1e-30))
        assert np.max(err[ok]) <= tol, (k, float(np.max(err[ok])), float(x64[ok][np.argmax(err[ok])]))
    if dtype == np.float64:
        w = got[0] * got[3] - got[1] * got[2]
        near = np.abs(x64) < 6
        assert np.max(np.abs(w[near] * np.pi - 1.0)) <= 1e-10
    xs = np.concatenate([rng.uniform(-30.0, 30.0, 4000), rng.uniform(-0.3, 0.3, 300), [0.0, 0.2, -0.2, 150.0]]).astype(dtype)
    xsd, xs64 = ag.DArray.from_numpy(xs), xs.astype(np.float64)
    d = ag.dawson(xsd).to_numpy().astype(np.float64)
    assert np.max(np.abs(d - sp.dawsn(xs64)) / np.maximum(np.abs(sp.dawsn(xs64)), 1e-30)) <= tol
    xe = rng.uniform(-5.0, 5.0, 3000).astype(dtype)
    e = ag.erfi(ag.DArray.from_numpy(xe)).to_numpy().astype(np.float64)
    refe = sp.erfi(xe.astype(np.float64))
    assert np.max(np.abs(e - refe) / np.maximum(np.abs(refe), 1e-30)) <= tol
    # rules: (Ai)' = Ai', (Ai')' = x Ai, (Bi')' = x Bi, dawson' = 1 - 2 x dawson, erfi' = 2/sqrt(pi) exp(x^2)
    xg = rng.uniform(-6.0, 4.0, 2000).astype(dtype)
    xgd, xg64 = ag.DArray.from_numpy(xg), xg.astype(np.float64)
    ra = sp.airy(xg64)
    gtol = 2e-5 if dtype == np.float32 else 1e-11

    def close(g, r):
        return np.max(np.abs(g.astype(np.float64) - r)) <= gtol * max(np.max(np.abs(r)), 1.0)
    assert close(ag.grad(lambda v: ag.sum_(ag.airyai(v)))(xgd).to_numpy(), ra[1])
    assert close(ag.grad(lambda v: ag.sum_(ag.airyaiprime(v)))(xgd).to_numpy(), xg64 * ra[0])
    assert close(ag.grad(lambda v: ag.sum_(ag.airybi(v)))(xgd).to_numpy(), ra[3])
    assert close(ag.grad(lambda v: ag.sum_(ag.airybiprime(v)))(xgd).to_numpy(), xg64 * ra[2])
    assert close(ag.grad(lambda v: ag.sum_(ag.dawson(v)))(xgd).to_numpy(), 1 - 2 * xg64 * sp.dawsn(xg64))
    xe64 = xe.astype(np.float64)
    assert close(ag.grad(lambda v: ag.sum_(ag.erfi(v)))(ag.DArray.from_numpy(xe)).to_numpy(),
                 2 / np.sqrt(np.pi) * np.exp(xe64 * xe64))


@pytest.mark.parametrize("dtype,tol", [(np.float32, 1e-5), (np.float64, 1e-12)])
def test_linear_algebra_helpers_on_device(ag, oracle, dtype, tol):
    """diag / diagm / tr / tril / triu / kron (src/linearalgebra.jl:19-20,86,90,92,176-186): values against numpy,
    gradients against the oracle's (which carry the reference's rule expressions)."""
    rng = np.random.default_rng(47)
    sq = np.asfortranarray(rng.standard_normal((6, 6)).astype(dtype))
    rect = np.asfortranarray(rng.standard_normal((4, 7)).astype(dtype))
    sqd, rectd = ag.DArray.from_numpy(sq), ag.DArray.from_numpy(rect)
    for k in (0, 1, -2, 5):
        assert np.array_equal(ag.diag(sqd, k).to_numpy(), np.diag(sq, k)), k
        assert np.array_equal(ag.diag(rectd, k).to_numpy(), np.diag(rect, k)), k
    v = rng.standard_normal(5).astype(dtype)
    vd = ag.DArray.from_numpy(v)
    for k in (0, 2, -1):
        assert np.array_equal(ag.diagm(vd, k).to_numpy(), np.diag(v, k)), k
    assert np.array_equal(ag.tril(rectd).to_numpy(), np.tril(rect)) and np.array_equal(ag.triu(rectd).to_numpy(), np.triu(rect))
    assert abs(float(ag.value(ag.tr(sqd))) - np.trace(sq.astype(np.float64))) <= tol
    a, b = np.asfortranarray(rng.standard_normal((3, 4)).astype(dtype)), np.asfortranarray(rng.standard_normal((5, 2)).astype(dtype))
    ad, bd = ag.DArray.from_numpy(a), ag.DArray.from_numpy(b)
    assert np.array_equal(ag.kron(ad, bd).to_numpy(), np.kron(a, b))
    w = np.asfortranarray(rng.standard_normal((6, 6)))
    wd = ag.DArray.from_numpy(w.astype(dtype))
    sq64 = np.asfortranarray(sq.astype(np.float64))
    for fo, fd in ((lambda z: oracle.diag(z, 1), lambda z: ag.diag(z, 1)), (oracle.tr, ag.tr), (oracle.tril, ag.tril),
                   (oracle.triu, ag.triu)):
        def lo(z):
            r = fo(z)
            return oracle.sum_(oracle.mul(r, w[:np.shape(oracle.value(r))[0]] if np.ndim(oracle.value(r)) == 2 else (w[0, :np.size(oracle.value(r))] if np.ndim(oracle.value(r)) else 1.0)))
        def ld(z):
            r = fd(z)
            shp = ag.size(r)
            wt = wd if len(shp) == 2 else (ag.DArray.from_numpy(w[0, :shp[0]].astype(dtype)) if len(shp) else 1.0)
            return ag.sum_(ag.mul(r, wt))
        go = oracle.grad(lo)(sq64)
        gd = ag.grad(ld)(sqd)
        gd = gd.full().to_numpy() if hasattr(gd, "full") else gd.to_numpy()
        assert np.max(np.abs(gd - go)) <= tol * 10, fd
    a64, b64 = np.asfortranarray(a.astype(np.float64)), np.asfortranarray(b.astype(np.float64))
    wk = np.asfortranarray(rng.standard_normal((15, 8)))
    wkd = ag.DArray.from_numpy(wk.astype(dtype))
    go = oracle.grad(lambda p: oracle.sum_(oracle.mul(oracle.kron(p, b64), wk)))(a64)
    gd = ag.grad(lambda p: ag.sum_(ag.mul(ag.kron(p, bd), wkd)))(ad).to_numpy()
    assert np.max(np.abs(gd - go)) <= tol * 20
    go = oracle.grad(lambda q: oracle.sum_(oracle.mul(oracle.kron(a64, q), wk)))(b64)
    gd = ag.grad(lambda q: ag.sum_(ag.mul(ag.kron(ad, q), wkd)))(bd).to_numpy()
    assert np.max(np.abs(gd - go)) <= tol * 20


@pytest.mark.parametrize("dtype,tol", [(np.float32, 2e-6), (np.float64, 1e-13)])
def test_inv_det_logdet_on_device(ag, oracle, dtype, tol):
    """inv / det / logdet / logabsdet (src/linearalgebra.jl:18,28,57-58; ag_inv_det: one-launch Gauss-Jordan elimination
    with partial pivoting): values against LAPACK (numpy.linalg), gradients against the oracle's rules, a matrix that
    needs row exchanges, a singular one."""
    rng = np.random.default_rng(53)
    for n in (1, 2, 5, 33, 130):
        A = np.asfortranarray((rng.standard_normal((n, n)) + (0.0 if n < 33 else 0.2) * np.eye(n)).astype(dtype))
        Ad, A64 = ag.DArray.from_numpy(A), A.astype(np.float64)
        scale = np.linalg.cond(A64)
        got = ag.inv(Ad).to_numpy()
        assert got.dtype == dtype and np.max(np.abs(got - np.linalg.inv(A64))) <= tol * scale * max(np.max(np.abs(np.linalg.inv(A64))), 1.0), n
        sign, lad = np.linalg.slogdet(A64)
        assert abs(np.linalg.det(A64)) > 1e30 or abs(ag.det(Ad) - np.linalg.det(A64)) <= tol * scale * abs(np.linalg.det(A64)), n
        l, s = ag.logabsdet(Ad)
        assert abs(l - lad) <= tol * scale * max(abs(lad), 1.0) and s == sign, n
    P = np.asfortranarray(np.array([[0.0, 2.0, 1.0], [1.0, 0.0, 3.0], [4.0, 1.0, 0.0]], dtype=dtype))      # zero leading pivots
    assert np.allclose(ag.inv(ag.DArray.from_numpy(P)).to_numpy(), np.linalg.inv(P.astype(np.float64)), atol=tol * 10)
    assert abs(ag.det(ag.DArray.from_numpy(P)) - 25.0) <= tol * 100
    S = np.asfortranarray(np.array([[1.0, 2.0], [2.0, 4.0]], dtype=dtype))
    assert ag.det(ag.DArray.from_numpy(S)) == 0.0
    spd = rng.standard_normal((6, 6))
    spd = np.asfortranarray((spd @ spd.T + 6 * np.eye(6)).astype(dtype))
    spdd, spd64 = ag.DArray.from_numpy(spd), np.asfortranarray(spd.astype(np.float64))
    assert abs(ag.logdet(spdd) - np.linalg.slogdet(spd64)[1]) <= tol * 100
    w = np.asfortranarray(rng.standard_normal((6, 6)))
    wd = ag.DArray.from_numpy(w.astype(dtype))
    go = oracle.grad(lambda z: oracle.sum_(oracle.mul(oracle.inv(z), w)))(spd64)
    gd = ag.grad(lambda z: ag.sum_(ag.mul(ag.inv(z), wd)))(spdd).to_numpy()
    assert np.max(np.abs(gd - go)) <= tol * 1e3 * np.max(np.abs(go))
    for fo, fd in ((oracle.det, ag.det), (oracle.logdet, ag.logdet)):
        go = oracle.grad(fo)(spd64)
        gd = ag.grad(fd)(spdd).to_numpy()
        assert np.max(np.abs(gd - go)) <= tol * 1e3 * np.max(np.abs(go)), fd
    gd = ag.grad(lambda z: ag.logabsdet(z)[0])(spdd).to_numpy()
    assert np.max(np.abs(gd - np.linalg.inv(spd64).T)) <= tol * 1e3
    with pytest.raises(ag.DimensionMismatch):
        ag.inv(ag.DArray.from_numpy(np.asfortranarray(np.ones((2, 3), dtype=dtype))))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_zerograd_value_helpers_on_device(ag, dtype):
    """The value-touching @zerograd helpers of src/base.jl (predicates, counts, integer division, array queries): device
    results against numpy; a Value argument is unboxed and nothing is recorded."""
    x = np.asfortranarray(np.array([[1.5, -2.0, np.nan, 0.0], [np.inf, -np.inf, -0.0, 7.0], [3.25, -3.0, 2.0, -7.5]], dtype=dtype))
    xd = ag.DArray.from_numpy(x)
    with np.errstate(invalid="ignore"):
        assert np.array_equal(ag.isnan(xd).to_numpy(), np.isnan(x).astype(dtype)) and np.array_equal(ag.isinf(xd).to_numpy(), np.isinf(x).astype(dtype))
        assert np.array_equal(ag.isfinite(xd).to_numpy(), np.isfinite(x).astype(dtype))
        assert np.array_equal(ag.signbit(xd).to_numpy()[~np.isnan(x)], np.signbit(x)[~np.isnan(x)].astype(dtype))
        assert np.array_equal(ag.isinteger(xd).to_numpy(), (np.isfinite(x) & (x == np.trunc(x))).astype(dtype))
    assert ag.count(ag.isnan(xd)) == 1 and ag.any_(ag.isinf(xd)) and not ag.all_(ag.isfinite(xd)) and ag.all_(ag.eq(xd, xd)) is False
    a = np.asfortranarray(np.array([[7.5, -7.5, 9.0], [2.0, 3.0, -4.0]], dtype=dtype))
    b = np.asfortranarray(np.array([[2.0, 2.0, -4.0], [0.5, 5.0, 3.0]], dtype=dtype))
    ad, bd = ag.DArray.from_numpy(a), ag.DArray.from_numpy(b)
    assert np.array_equal(ag.div_(ad, bd).to_numpy(), np.trunc(a / b)) and np.array_equal(ag.rem(ad, bd).to_numpy(), np.fmod(a, b))
    # rem is differentiable (src/base.jl:144): dx1 = dy, dx2 = -dy .* div.(x1, x2); float / widemul pass gradients
    g1 = ag.grad(lambda v: ag.sum_(ag.rem(v, bd)))(ad).to_numpy()
    g2 = ag.grad(lambda v: ag.sum_(ag.rem(ad, v)))(bd).to_numpy()
    assert np.array_equal(g1, np.ones_like(a)) and np.array_equal(g2, -np.trunc(a / b))
    assert np.array_equal(ag.grad(lambda v: ag.sum_(ag.widemul(ag.float_(v), bd)))(ad).to_numpy(), b)
    assert np.array_equal(ag.ldiv(ad, bd).to_numpy(), b / a)
    assert np.array_equal(ag.ne(ad, bd).to_numpy(), (a != b).astype(dtype)) and np.array_equal(ag.widemul(ad, bd).to_numpy(), a * b)
    assert ag.isequal(xd, ag.DArray.from_numpy(x.copy(order="F"))) and not ag.isequal(ad, bd) and not ag.isequal(ad, xd)
    assert ag.isapprox(ad, ag.DArray.from_numpy(np.asfortranarray(a * (1 + 1e-9 if dtype == np.float64 else 1 + 1e-5)))) and not ag.isapprox(ad, bd)
    assert ag.eltype(ad) == dtype and ag.eps(ad) == float(np.finfo(dtype).eps) and ag.sizeof(ad) == a.nbytes and not ag.isempty(ad)
    assert ag.strides(ad) == (1, 2) and ag.stride(ad, 1) == 2 and ag.lastindex(ad) == 5 and ag.axes(ad) == (range(2), range(3))
    assert ag.similar(ad).shape == (2, 3) and not np.any(ag.zero(ad).to_numpy()) and np.all(ag.ones(ad).to_numpy() == 1)
    assert ag.typemax(ad) == np.inf and ag.typemin(ad) == -np.inf and ag.float_(ad) is ad
    big = np.asfortranarray(np.random.default_rng(59).standard_normal((257, 1031)).astype(dtype))
    big[5, 7] = big[200, 900] = 9.0
    big[100, 3] = big[11, 1000] = -9.0
    bigd = ag.DArray.from_numpy(big)
    flat = big.reshape(-1, order="F")
    assert ag.argmax(bigd) == int(np.argmax(flat)) == 5 + 7 * 257 and ag.argmin(bigd) == int(np.argmin(flat)) == 100 + 3 * 257
    assert ag.argmax(xd) == ag.argmin(xd) == 2 * 3 + 0                   # the NaN at (0, 2) wins both, like Julia
    assert ag.argmax(ag.Param(bigd)) == 5 + 7 * 257
    # oftype / big (src/base.jl:76,365): eltype conversion on the device, the gradient comes back in x's eltype
    other = np.float32 if dtype == np.float64 else np.float64
    odd = np.asfortranarray(np.random.default_rng(67).standard_normal((7, 11)).astype(dtype))          # 77 elements: tail path
    oddd = ag.DArray.from_numpy(odd)
    conv = ag.oftype(other, oddd)
    assert conv.dtype == other and np.array_equal(conv.to_numpy(), odd.astype(other)) and ag.oftype(dtype, oddd) is oddd
    assert ag.big(oddd).dtype == np.float64 and np.array_equal(ag.big(oddd).to_numpy(), odd.astype(np.float64))
    g = ag.grad(lambda v: ag.sum_(ag.abs2(ag.oftype(other, v))))(oddd)
    assert g.dtype == dtype and np.allclose(g.to_numpy(), 2 * odd, rtol=1e-6)
    # unboxing: inside a tape the helpers see the Param's value and contribute no gradient
    g = ag.grad(lambda v: ag.sum_(ag.mul(v, ag.isfinite(v))))(ad).to_numpy()
    assert np.array_equal(g, np.ones_like(a))


def test_inv_det_limits_and_tracked_iteration(ag):
    """ag_inv_det refuses n > 1024 loudly (no fallback); iterating a tracked device array goes through getindex
    (src/iterate.jl:6-17) and the gradient comes back through the Sparse scatter."""
    big = ag.DArray.from_numpy(np.asfortranarray(np.eye(1025, dtype=np.float32)))
    with pytest.raises(ag.UnsupportedOnDevice):
        ag.inv(big)
    ok = ag.DArray.from_numpy(np.asfortranarray(2 * np.eye(1024, dtype=np.float32)))
    assert np.array_equal(ag.inv(ok).to_numpy(), 0.5 * np.eye(1024, dtype=np.float32))
    x = np.asfortranarray(np.array([1.5, -2.0, 0.5], dtype=np.float32))
    xd = ag.DArray.from_numpy(x)

    def f(v):
        a, b, c = v
        t = ag.mul(a, b)
        for e in v:
            t = ag.add(t, ag.mul(e, c))
        return t
    g = ag.grad(f)(xd)
    g = g.full().to_numpy() if hasattr(g, "full") else g.to_numpy()
    assert np.allclose(g, [x[1] + x[2], x[0] + x[2], x[0] + x[1] + 2 * x[2]])
