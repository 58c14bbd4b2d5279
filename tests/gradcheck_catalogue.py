"""Finite-difference check of EVERY primitive of the catalogue (shared by the GPU parity test and the CPU host-logic
test; the reference's instrument is test/gradcheck.jl:35-133, its per-primitive lists test/math.jl:21-87, test/base.jl)."""
import numpy as np


MATH_REST = ("acosd acot acotd acoth acsc acscd acsch asec asecd asech asind atand cosd cospi cot cotd coth csc cscd csch "
             "deg2rad rad2deg sec secd sech sind sinpi tand sinc cosc").split()
MATH_REST_DOMAINS = {"acosd": (-0.8, 0.8), "asind": (-0.8, 0.8), "asech": (0.2, 0.8)}

def check_catalogue(ag):
    rng = np.random.default_rng(18)
    dev = ag.DArray.from_numpy
    doms = {"acosh": (1.2, 2.0), "log": (0.2, 2.0), "sqrt": (0.2, 2.0), "log2": (0.2, 2.0), "log10": (0.2, 2.0),
            "log1p": (-0.5, 2.0), "asin": (-0.8, 0.8), "acos": (-0.8, 0.8), "atanh": (-0.8, 0.8), "cbrt": (0.2, 2.0),
            "tan": (-1.0, 1.0), "abs_": (0.2, 2.0), "relu": (0.2, 2.0)}
    unary = ["neg", "abs_", "abs2", "exp", "log", "tanh", "sigm", "sqrt", "sin", "cos", "tan", "sinh", "cosh", "asin",
             "acos", "atan", "asinh", "acosh", "atanh", "exp2", "exp10", "expm1", "log2", "log10", "log1p", "cbrt",
             "relu", "identity", "copy"]
    for name in unary:
        lo, hi = doms.get(name, (-1.0, 1.0))
        x = dev(rng.uniform(lo, hi, (3, 4)))
        assert ag.gradcheck(getattr(ag, name), x), name
    # src/specialfunctions.jl (erf family, gamma family, Bessel functions of order 0 and 1)
    for name, (lo, hi) in {"erf": (-1, 1), "erfc": (-1, 1), "erfcx": (0.1, 1.5), "erfinv": (-0.7, 0.7),
                           "erfcinv": (0.3, 1.5), "gamma": (1.2, 3.0), "lgamma": (1.2, 3.0), "digamma": (1.2, 3.0),
                           "trigamma": (1.2, 3.0), "besselj0": (0.5, 3.0), "besselj1": (0.5, 3.0),
                           "bessely0": (0.5, 3.0), "bessely1": (0.5, 3.0)}.items():
        assert ag.gradcheck(getattr(ag, name), dev(rng.uniform(lo, hi, (3, 4)))), name
    # the rest of src/math.jl: degree / reciprocal / pi-scaled trigonometry, sinc / cosc, clamp, ldexp, log(b, x)
    for name in MATH_REST:
        lo, hi = MATH_REST_DOMAINS.get(name, (1.2, 2.0))
        assert ag.gradcheck(getattr(ag, name), dev(rng.uniform(lo, hi, (3, 4)))), name
    away = dev(np.array([[1.25, 1.45, 1.62], [1.9, 1.31, 1.05]]))            # not within the step of a clamp bound
    assert ag.gradcheck(lambda z: ag.clamp(z, 1.4, 1.7), away)
    assert ag.gradcheck(lambda z: ag.ldexp(z, 3), away)
    assert ag.gradcheck(ag.logb, away, dev(np.array([[2.5, 3.0, 1.5], [2.2, 4.0, 1.8]])))
    assert ag.gradcheck(lambda z: ag.logb(z, 3.0), away) and ag.gradcheck(lambda z: ag.logb(2.0, z), away)
    # src/specialfunctions.jl:5-8,33,39: Airy functions, dawson, erfi (their own pass on the device)
    for name, (lo, hi) in {"airyai": (-3.0, 3.0), "airyaiprime": (-3.0, 3.0), "airybi": (-3.0, 2.5),
                           "airybiprime": (-3.0, 2.5), "dawson": (-2.0, 2.0), "erfi": (-1.2, 1.2)}.items():
        assert ag.gradcheck(getattr(ag, name), dev(rng.uniform(lo, hi, (3, 4)))), name
    a, b = dev(rng.uniform(0.5, 1.5, (3, 4))), dev(rng.uniform(0.5, 1.5, (3, 4)))
    row, col = dev(rng.uniform(0.5, 1.5, (1, 4))), dev(rng.uniform(0.5, 1.5, (3, 1)))
    for name in ["add", "sub", "mul", "div", "pow_", "atan2", "hypot"]:
        f = getattr(ag, name)
        for q in (b, row, col):
            assert ag.gradcheck(f, a, q), (name, q.shape)
        assert ag.gradcheck(lambda z: f(z, 1.7), a) and ag.gradcheck(lambda z: f(1.7, z), a), name
    far = dev(np.asarray(a.to_numpy() + np.where(rng.random((3, 4)) > 0.5, 0.4, -0.4)))      # no ties
    for name in ["max_", "min_"]:
        assert ag.gradcheck(getattr(ag, name), a, far), name
    for f in (ag.sum_, ag.sumabs2, ag.sumabs, ag.mean):
        assert ag.gradcheck(f, a), f
        for dims in (0, 1):
            assert ag.gradcheck(lambda z: f(z, dims=dims), a), (f, dims)
    distinct = dev(rng.permutation(12).reshape(3, 4) / 7.0 + 0.3)
    for f in (ag.maximum, ag.minimum, ag.prod):
        assert ag.gradcheck(f, distinct) and ag.gradcheck(lambda z: f(z, dims=0), distinct), f
    for f in (ag.var, ag.std, ag.norm):
        assert ag.gradcheck(f, a), f
    spread = dev(rng.permutation(12).reshape(3, 4) / 7.0 - 0.8)          # distinct |values|, both signs
    for f in (ag.maxabs, ag.minabs):                                     # src/base.jl:203,206
        assert ag.gradcheck(f, spread) and ag.gradcheck(lambda z: f(z, dims=0), spread) and ag.gradcheck(lambda z: f(z, dims=1), spread), f
    num, den = dev(rng.uniform(3.1, 8.9, (3, 4))), dev(rng.uniform(1.05, 1.95, (3, 4)))
    assert ag.gradcheck(ag.ldiv, num, den) and ag.gradcheck(lambda z: ag.ldiv(z, 2.0), num) and ag.gradcheck(ag.widemul, num, den)
    assert ag.gradcheck(ag.float_, num)
    for f in (ag.meanabs, ag.meanabs2):                                  # src/statistics.jl:9-10
        assert ag.gradcheck(f, a) and ag.gradcheck(lambda z: f(z, dims=0), a) and ag.gradcheck(lambda z: f(z, dims=1), a), f
    assert ag.gradcheck(lambda z: ag.varm(z, 0.9, dims=0), a) and ag.gradcheck(lambda z: ag.stdm(z, 0.9), a)
    assert ag.gradcheck(ag.dot, a, b)
    # src/linearalgebra.jl:19-20,86,90,92,176-186
    sq = dev(rng.standard_normal((4, 4)))
    assert ag.gradcheck(ag.diag, sq) and ag.gradcheck(lambda z: ag.diag(z, 1), sq) and ag.gradcheck(lambda z: ag.diag(z, -2), sq)
    assert ag.gradcheck(ag.diag, a) and ag.gradcheck(ag.tr, sq) and ag.gradcheck(ag.tril, sq) and ag.gradcheck(ag.triu, a)
    assert ag.gradcheck(ag.diagm, dev(rng.standard_normal(3))) and ag.gradcheck(lambda z: ag.diagm(z, -1), dev(rng.standard_normal(3)))
    assert ag.gradcheck(ag.kron, dev(rng.standard_normal((2, 3))), dev(rng.standard_normal((4, 5))))
    wellc = dev(rng.standard_normal((4, 4)) * 0.3 + 2 * np.eye(4))
    assert ag.gradcheck(ag.inv, wellc) and ag.gradcheck(ag.det, wellc) and ag.gradcheck(ag.logdet, wellc)
    m1, m2 = dev(rng.standard_normal((4, 3))), dev(rng.standard_normal((3, 5)))
    assert ag.gradcheck(ag.matmul, m1, m2)
    assert ag.gradcheck(lambda p, q: ag.matmul(ag.adjoint(p), q), dev(rng.standard_normal((3, 4))), m2)
    assert ag.gradcheck(lambda p, q: ag.matmul(p, ag.adjoint(q)), m1, dev(rng.standard_normal((5, 3))))
    assert ag.gradcheck(lambda z: ag.reshape(z, (2, 6)), a) and ag.gradcheck(ag.vec, a) and ag.gradcheck(ag.adjoint, a)
    assert ag.gradcheck(lambda z: ag.permutedims(z, (1, 0)), a)
    assert ag.gradcheck(lambda p, q: ag.vcat(p, q), a, b) and ag.gradcheck(lambda p, q: ag.hcat(p, q), a, b)
    assert ag.gradcheck(lambda p, q: ag.cat1d(p, q), a, row)
    for idx in [(slice(None), [1, 1, 2]), ([0, 2],), (1, 2), (slice(0, 2), slice(1, 3))]:
        assert ag.gradcheck(lambda z: ag.getindex(z, *idx), a), idx
        assert ag.gradcheck(lambda z: ag.view(z, *idx), a), ("view", idx)             # src/getindex.jl:36
