"""Pins oracle/agref.py against every exact known-answer vector the reference holds for the hot
path.  Citations: file:line under /root/reference."""
import math

import numpy as np
import pytest

from oracle import agref as ag
from oracle.agref import Param, diff, grad, gradloss, value, Sparse, full


def F(a):
    return np.asfortranarray(np.array(a, dtype=np.float64))


def test_readme_sum_abs2():
    # README.md:37-44 / src/AutoGrad.jl:17-19
    x = Param(F([1, 2, 3]))
    y = diff(lambda: ag.sumabs2(x))
    assert value(y) == 14
    assert np.array_equal(grad(y, x), [2, 4, 6])


def test_gradloss():
    # src/AutoGrad.jl:38-42
    g, l = gradloss(lambda x: ag.sumabs2(x))(F([[1, 2, 3]]))
    assert l == 14 and np.array_equal(g, [[2, 4, 6]])


def test_power_rules_exact():
    # test/base.jl:94-110
    assert grad(lambda x: x ** 2)(1) == 2
    assert np.array_equal(grad(lambda x: ag.sum_(x ** 2))(F([1, 2, 3])), [2, 4, 6])
    assert np.array_equal(grad(lambda x: ag.sum_(x ** 2.0))(F([1, 2, 3])), [2., 4., 6.])
    r = F([[1, 2], [3, 4]])
    assert np.array_equal(grad(lambda x: ag.sum_(x ** 2))(r), [[2, 4], [6, 8]])
    assert np.array_equal(grad(lambda x: ag.sum_(x ** 2.0))(r), [[2, 4], [6, 8]])
    with pytest.raises(RuntimeError):
        grad(lambda x: ag.sum_(ag.matpow(x, 2)))(r)
    assert grad(lambda x: ag.sum_(x ** 3.1))(r).dtype == np.float64


def test_core_known_answers():
    # test/core.jl:35
    x = Param(1)
    assert grad(diff(lambda: ag.sin(ag.sqrt(x))), x) == math.cos(1) / 2
    # :40 (Issue 101.1) identity on a Param
    x = Param(1.0)
    assert grad(diff(lambda: x), x) == 1
    # :44 (Issue 101.2)
    x = Param(F([1., 2.]))
    assert np.array_equal(grad(diff(lambda: ag.sum_(1 * x)), x), [1.0, 1.0])
    # :48-49 (Issue 103)
    assert grad(lambda x: ag.exp(x))(1.) == math.exp(1.)
    assert np.array_equal(grad(lambda x: ag.sum_(ag.exp(x)))(F([1.])), [math.exp(1.)])
    # :65-68 result not last on tape
    x = Param(np.asfortranarray(np.random.default_rng(0).random((2, 3))))

    def f(x):
        x1 = ag.sum_(x)
        x2 = 2 * x  # noqa: F841 recorded after the result
        return x1
    assert np.array_equal(grad(diff(lambda: f(x)), x), np.ones((2, 3)))


def test_highorder_sin_cycle():
    # test/highorder.jl:6-14 -- exact equality through order 9
    g = ag.sin
    expect = [math.cos(1), -math.sin(1), -math.cos(1), math.sin(1)]
    for k in range(9):
        g = grad(g)
        assert g(1) == expect[k % 4], k


def test_highorder_exp_tanh():
    # test/highorder.jl:28-34
    assert math.exp(1) == grad(ag.exp)(1) == grad(grad(ag.exp))(1)
    f = lambda x: ag.exp(x * x)  # noqa: E731
    assert np.isclose(f(1), grad(f)(1) / 2) and np.isclose(f(1), grad(grad(f))(1) / 6)
    t = math.tanh(1)
    assert grad(ag.tanh)(1) == 1 - t ** 2
    assert np.isclose(grad(grad(ag.tanh))(1), -2 * t * (1 - t ** 2), rtol=1e-15)
    # PR #75 (test/highorder.jl:19)
    assert grad(lambda x: x * grad(lambda y: x + y)(1 * x))(5.0) == 1


def test_getindex_size_known_answer():
    # test/getindex.jl:51-58
    g = grad(lambda x: ag.size(x, 0) * ag.sumabs2(x))(np.ones((3, 3), order="F"))
    assert np.array_equal(g, np.full((3, 3), 6.0))


def test_sparse_known_answers():
    # test/sparse.jl:7-16 (indices shifted to 0-based)
    a = Sparse(np.zeros((3, 4), order="F"), [F([1., 1.]), F([1., 1.]), 1., 1.],
               [([0, 1],), (range(2, 4),), (1, 1), (0,)])
    fa = full(a)
    lin = fa.reshape(-1, order="F")
    assert list(lin[:5]) == [2, 1, 1, 1, 1] and not lin[5:].any()
    b = a + a
    assert isinstance(b, Sparse)
    assert np.array_equal(full(b), full(a) + full(a))
    ag.addto(a, Sparse(a.container, list(a.values), list(a.indices)))
    assert isinstance(a, Sparse)
    assert np.array_equal(full(a), full(b))
    # lmul! stays Sparse (test/sparse.jl:18-23)
    w = Param(np.asfortranarray(np.random.default_rng(1).standard_normal((2, 2))))

    def foo(w):
        s = 0.0
        for i in range(4):
            s = s + w[i]
        return s
    J = diff(lambda: foo(w))
    assert isinstance(ag.lmul(0.5, grad(J, w)), Sparse)
    assert np.array_equal(full(grad(J, w)), np.ones((2, 2)))


def test_tape_node_order_readme():
    # README.md:116-125 pins the node order for sum(abs2, w*x .+ b - y)
    rng = np.random.default_rng(0)
    w, b = Param(F(rng.standard_normal((2, 3)))), Param(F(rng.standard_normal(2)))
    x, y = F(rng.standard_normal((3, 4))), F(rng.standard_normal((2, 4)))
    J = diff(lambda: ag.sumabs2((w @ x + b) - y))
    names = [("P" if isinstance(n.Value, Param) else n.Value.func.name) for n in J.list]
    assert names == ["P", "*", "P", "+", "-", "sum(abs2,·)"]
    assert J.list[0].Value is w and J.list[2].Value is b


def test_relu_tie_passes_gradient():
    # src/math.jl:54: dy.*(y.==xi) -> derivative 1 at 0
    g = grad(lambda x: ag.sum_(ag.relu(x)))(F([-1., 0., 2.]))
    assert np.array_equal(g, [0., 1., 1.])
    g = grad(lambda x: ag.sum_(ag.abs_(x)))(F([-1., 0., 2.]))
    assert np.array_equal(g, [-1., 0., 1.])


def test_getindex_kinds_gradcheck():
    # test/getindex.jl:8-26 index kinds, repeated indices summed
    rng = np.random.default_rng(3)
    a = F(rng.standard_normal((3, 4)))
    for idx in [(slice(None),), (0,), (range(0, 2),), (range(0, 3, 2),), ([0, 1],), ([1, 1],), ([],),
                (np.array([True, False, True] * 4),), (0, 1), (slice(None), [1, 1, 3]),
                ([0, 2], range(1, 3)), (np.array([[0, 1], [2, 3]]),)]:
        assert ag.gradcheck(lambda x: ag.getindex(x, *idx), a), idx
    g = grad(lambda x: ag.sum_(ag.getindex(x, [1, 1])))(a)
    assert g.reshape(-1, order="F")[1] == 2.0


def test_unbroadcast_shapes():
    # test/base.jl:46-58 shapes (2,3,5)/(1,3,5) and (2,3,5,4)/(2,3,1,4)
    rng = np.random.default_rng(4)
    for s1, s2 in [((2, 3, 5), (1, 3, 5)), ((2, 3, 5, 4), (2, 3, 1, 4)), ((64, 7), (64,)), ((5, 7), (1, 7)), ((5, 7), ())]:
        a = F(rng.standard_normal(s1))
        b = F(rng.standard_normal(s2)) if s2 else float(rng.standard_normal())
        for op in (ag.add, ag.sub, ag.mul, ag.div):
            assert ag.gradcheck(lambda p, q: op(p, q), a, b), (s1, s2, op)
            assert ag.gradcheck(lambda p, q: op(q, p), a, b), (s1, s2, op)
    x = Param(F(rng.standard_normal(64)))
    J = diff(lambda: ag.sum_(F(rng.standard_normal((64, 5))) + x))
    assert grad(J, x).shape == (64,)
    with pytest.raises(ValueError):
        ag.unbroadcast(np.zeros((3, 2)), np.zeros((3, 4)))


def test_every_rule_gradcheck():
    # test/math.jl:21-87, test/base.jl:28-89, test/statistics.jl:9-21, test/cat.jl, test/linearalgebra.jl:57
    rng = np.random.default_rng(5)
    x = F(rng.standard_normal((3, 4)))
    xp = np.abs(x) + 0.5
    for f, a in [(ag.exp, x), (ag.log, xp), (ag.tanh, x), (ag.sin, x), (ag.cos, x), (ag.sqrt, xp), (ag.sigm, x),
                 (ag.abs_, x), (ag.abs2, x), (ag.neg, x), (ag.relu, x), (ag.vec, x), (ag.copy, x), (ag.adjoint, x)]:
        assert ag.gradcheck(f, a), f
    for f in (ag.sum_, ag.sumabs2, ag.sumabs, ag.mean, ag.meanabs2, ag.maximum, ag.minimum, ag.prod):
        assert ag.gradcheck(f, x), f
        for d in (0, 1, (0, 1)):
            assert ag.gradcheck(lambda a: f(a, dims=d), x), (f, d)
    w = F(rng.standard_normal((2, 3)))
    assert ag.gradcheck(ag.matmul, w, x)
    assert ag.gradcheck(lambda a, b: ag.matmul(a, ag.adjoint(b)), x, x.copy(order="F"))
    assert ag.gradcheck(ag.max_, x, F(rng.standard_normal((3, 4))))
    assert ag.gradcheck(ag.min_, x, F(rng.standard_normal((3, 1))))
    assert ag.gradcheck(lambda a: a ** 3, xp) and ag.gradcheck(lambda a, b: a ** b, xp, xp.copy(order="F"))
    assert ag.gradcheck(ag.vcat, x, x.copy(order="F")) and ag.gradcheck(ag.hcat, x, F(rng.standard_normal((3, 2))))
    assert ag.gradcheck(lambda a, b: ag.cat1d(a, b), x, w)
    assert ag.gradcheck(lambda a: ag.reshape(a, (4, 3)), x)
    assert ag.gradcheck(lambda a: ag.permutedims(a, (1, 0)), x)


def test_neuralnet_first_and_second_order():
    # test/neuralnet.jl:4-18 (2-layer relu MLP); test/highorder.jl:36-42 (hessian shapes)
    rng = np.random.default_rng(6)
    w1, b1, w2, b2 = (F(rng.standard_normal(s)) for s in [(4, 3), (4,), (2, 4), (2,)])
    x, y = F(rng.standard_normal((3, 5))), F(rng.standard_normal((2, 5)))

    def loss(w1, b1, w2, b2):
        h = ag.relu(ag.add(ag.matmul(w1, x), b1))
        return ag.sumabs2(ag.sub(ag.add(ag.matmul(w2, h), b2), y))
    assert ag.gradcheck(loss, w1, b1, w2, b2)

    def nll(p):
        p = ag.exp(p)
        p = ag.div(p, ag.sum_(p, dims=0))
        p = ag.getindex(p, 0, slice(None))
        return ag.mean(ag.neg(ag.log(p)))
    w, b, x = F(rng.standard_normal((2, 3))), F(rng.standard_normal(2)), F(rng.standard_normal((3, 4)))
    hess = lambda f: grad(lambda a: ag.sum_(grad(f)(a)))  # noqa: E731
    assert hess(lambda w_: nll(ag.add(ag.matmul(w_, x), b)))(w).shape == w.shape
    assert hess(lambda b_: nll(ag.add(ag.matmul(w, x), b_)))(b).shape == b.shape
    # second-order gradient checked numerically
    assert ag.gradcheck(lambda w_: ag.sum_(grad(lambda a: nll(ag.add(ag.matmul(a, x), b)))(w_)), w)


def test_softmax_xent_identity():
    # analytic identity (SURVEY 8c "additional"): d/dp of -sum(y.*(p .- log(sum(exp(p),dims=1))))/B = (softmax(p)-y)/B
    rng = np.random.default_rng(7)
    p = F(rng.standard_normal((10, 6)))
    y = np.zeros((10, 6), order="F")
    y[rng.integers(0, 10, 6), np.arange(6)] = 1

    def loss(p):
        return ag.div(ag.neg(ag.sum_(ag.mul(y, ag.sub(p, ag.log(ag.sum_(ag.exp(p), dims=0)))))), 6)
    g = grad(loss)(p)
    sm = np.exp(p) / np.exp(p).sum(0, keepdims=True)
    assert np.allclose(g, (sm - y) / 6, rtol=1e-13, atol=1e-15)


def test_special_function_known_answers(oracle):
    """src/specialfunctions.jl rules on closed-form points: erf'(0) = 2/sqrt(pi), gamma(5) = 24 with
    gamma'(1) = -euler_gamma = digamma(1), trigamma(1) = pi^2/6 = digamma'(1), lgamma'(x) = digamma(x),
    besselj0'(x) = -besselj1(x), erfinv(erf(x)) = x."""
    import math
    g = oracle.grad
    eg = 0.5772156649015329
    assert abs(g(lambda x: oracle.sum_(oracle.erf(x)))(np.array([0.0]))[0] - 2 / math.sqrt(math.pi)) < 1e-15
    assert abs(g(lambda x: oracle.sum_(oracle.erfc(x)))(np.array([0.0]))[0] + 2 / math.sqrt(math.pi)) < 1e-15
    assert abs(oracle.gamma(np.array([5.0]))[0] - 24.0) < 1e-12
    assert abs(g(lambda x: oracle.sum_(oracle.gamma(x)))(np.array([1.0]))[0] + eg) < 1e-12
    assert abs(oracle.digamma(np.array([1.0]))[0] + eg) < 1e-14
    assert abs(g(lambda x: oracle.sum_(oracle.digamma(x)))(np.array([1.0]))[0] - math.pi ** 2 / 6) < 1e-12
    assert abs(g(lambda x: oracle.sum_(oracle.lgamma(x)))(np.array([2.5]))[0] - oracle.digamma(np.array([2.5]))[0]) < 1e-14
    x = np.array([0.3, 1.7, 4.0])
    assert np.allclose(g(lambda v: oracle.sum_(oracle.besselj0(v)))(x), -oracle.besselj1(x), rtol=1e-14)
    assert np.allclose(oracle.erfinv(oracle.erf(np.array([0.4]))), [0.4], rtol=1e-13)
    for name, pt in [("erf", 0.3), ("erfc", 0.3), ("erfcx", 0.7), ("erfinv", 0.4), ("erfcinv", 0.6), ("gamma", 2.3),
                     ("lgamma", 2.3), ("digamma", 1.9), ("trigamma", 1.9), ("besselj0", 1.1), ("besselj1", 1.1),
                     ("bessely0", 1.1), ("bessely1", 1.1)]:
        assert oracle.gradcheck(getattr(oracle, name), np.asfortranarray([[pt, pt + 0.2], [pt + 0.35, pt + 0.1]])), name


def test_airy_dawson_erfi_known_answers(oracle):
    """src/specialfunctions.jl:5-8,33,39.  Exact values: Ai(0) = 3^(-2/3)/Gamma(2/3), Ai'(0) = -3^(-1/3)/Gamma(1/3),
    Bi(0) = sqrt(3) Ai(0), Bi'(0) = -sqrt(3) Ai'(0); the Wronskian Ai Bi' - Ai' Bi = 1/pi at any x; the rules
    (Ai)'' = x Ai; dawson'(x) = 1 - 2 x dawson(x) with dawson(0) = 0; erfi'(0) = 2/sqrt(pi)."""
    g = oracle.grad
    z = np.array([0.0])
    ai0, aip0 = 3 ** (-2 / 3) / math.gamma(2 / 3), -(3 ** (-1 / 3)) / math.gamma(1 / 3)
    assert abs(oracle.airyai(z)[0] - ai0) < 1e-15 and abs(oracle.airyaiprime(z)[0] - aip0) < 1e-15
    assert abs(oracle.airybi(z)[0] - math.sqrt(3) * ai0) < 1e-15 and abs(oracle.airybiprime(z)[0] + math.sqrt(3) * aip0) < 1e-15
    x = np.array([-4.3, -1.0, 0.4, 2.2])
    w = oracle.airyai(x) * oracle.airybiprime(x) - oracle.airyaiprime(x) * oracle.airybi(x)
    assert np.allclose(w, 1 / math.pi, rtol=1e-12)
    assert np.allclose(g(lambda v: oracle.sum_(oracle.airyai(v)))(x), oracle.airyaiprime(x), rtol=1e-14)
    assert np.allclose(g(lambda v: oracle.sum_(oracle.airyaiprime(v)))(x), x * oracle.airyai(x), rtol=1e-14)
    assert np.allclose(g(lambda v: oracle.sum_(oracle.airybiprime(v)))(x), x * oracle.airybi(x), rtol=1e-14)
    assert oracle.dawson(z)[0] == 0.0 and g(lambda v: oracle.sum_(oracle.dawson(v)))(z)[0] == 1.0
    assert np.allclose(g(lambda v: oracle.sum_(oracle.dawson(v)))(x), 1 - 2 * x * oracle.dawson(x), rtol=1e-14)
    assert abs(g(lambda v: oracle.sum_(oracle.erfi(v)))(z)[0] - 2 / math.sqrt(math.pi)) < 1e-15
    for name, pt in [("airyai", -1.3), ("airyaiprime", 0.4), ("airybi", -2.1), ("airybiprime", 0.6), ("dawson", 0.8),
                     ("erfi", 0.5)]:
        assert oracle.gradcheck(getattr(oracle, name), np.asfortranarray([[pt, pt + 0.2], [pt + 0.35, pt + 0.1]])), name


def test_norm_known_answers(oracle):
    """src/linearalgebra.jl:64-65,222-238 on a 3-4-5 triangle: norm = 5, gradient x/5; p = 1: 7, sign(x); p = Inf: 4,
    the gradient goes to the largest entry; p = 0 counts nonzeros and has a zero gradient."""
    x = np.array([3.0, -4.0, 0.0])
    g = lambda p: oracle.grad(lambda v: oracle.norm(v, p))(x)
    assert oracle.norm(x) == 5.0 and np.allclose(g(2), [0.6, -0.8, 0.0])
    assert oracle.norm(x, 1) == 7.0 and np.array_equal(g(1), [1.0, -1.0, 0.0])
    assert oracle.norm(x, np.inf) == 4.0 and np.array_equal(g(np.inf), [0.0, -1.0, 0.0])
    assert oracle.norm(x, 0) == 2.0 and not np.any(g(0))
    assert abs(oracle.norm(x, 3) - (27 + 64) ** (1 / 3)) < 1e-14
