"""GPU parity of whole tapes (forward + backward sweep + update) for the BASELINE.json configs against
the oracle, plus size-independent properties at the full bench size."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def relerr(got, ref):
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    return np.max(np.abs(got - ref)) / max(np.max(np.abs(ref)), 1e-300)


def both(ag, oracle, loss, params_np, consts_np, dtype):
    po = [oracle.Param(np.asfortranarray(p.astype(np.float64))) for p in params_np]
    co = [np.asfortranarray(c.astype(np.float64)) if isinstance(c, np.ndarray) and c.dtype.kind == "f" else c
          for c in consts_np]
    to = oracle.diff(lambda: loss(oracle, po, *co))
    go = [oracle.full(oracle.grad(to, p)) for p in po]
    pd = [ag.Param(ag.DArray.from_numpy(p.astype(dtype))) for p in params_np]
    cd = [ag.DArray.from_numpy(c.astype(dtype)) if isinstance(c, np.ndarray) and c.dtype.kind == "f" else c
          for c in consts_np]
    td = ag.diff(lambda: loss(ag, pd, *cd))
    gd = [ag.full(ag.grad(td, p)).to_numpy() for p in pd]
    return float(oracle.value(to)), go, float(ag.value(td)), gd


def test_readme_known_answers(ag):
    x = ag.Param(ag.DArray.from_numpy(np.array([1.0, 2.0, 3.0])))
    t = ag.diff(lambda: ag.sumabs2(x))
    assert float(ag.value(t)) == 14.0
    assert np.array_equal(ag.grad(t, x).to_numpy(), [2.0, 4.0, 6.0])


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-12), (np.float32, 1e-5)])
def test_housing(ag, oracle, dtype, tol):
    W = ag.workloads
    w, x, y = W.housing_init(np.random.default_rng(1))
    lo, go, ld, gd = both(ag, oracle, W.housing_loss, w, [x, y], dtype)
    assert abs(lo - ld) <= tol * abs(lo)
    for a, b in zip(gd, go):
        assert relerr(a, b) <= tol


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-12), (np.float32, 1e-5)])
@pytest.mark.parametrize("batch", [100, 4096])
def test_mnist_mlp(ag, oracle, dtype, tol, batch):
    W = ag.workloads
    w, x, y = W.mnist_init(np.random.default_rng(2), batch)
    lo, go, ld, gd = both(ag, oracle, W.mnist_loss, w, [x, y], dtype)
    assert abs(lo - ld) <= tol * abs(lo)
    for a, b in zip(gd, go):
        assert relerr(a, b) <= tol


def test_mnist_sgd_trajectory(ag, oracle):
    """5 steps of @diff + grad + axpy! stay within tolerance of the oracle's trajectory."""
    W = ag.workloads
    w, x, y = W.mnist_init(np.random.default_rng(5), 512)
    po = [oracle.Param(np.asfortranarray(p.astype(np.float64))) for p in w]
    pd = [ag.Param(ag.DArray.from_numpy(p)) for p in w]
    xo, yo = x.astype(np.float64), y.astype(np.float64)
    xd, yd = ag.DArray.from_numpy(x), ag.DArray.from_numpy(y)
    for _ in range(5):
        to = oracle.diff(lambda: W.mnist_loss(oracle, po, xo, yo))
        for p in po:
            oracle.axpy(-0.1, oracle.grad(to, p), p)
        td = ag.diff(lambda: W.mnist_loss(ag, pd, xd, yd))
        for p in pd:
            ag.axpy(-0.1, ag.grad(td, p), p)
    assert abs(float(ag.value(td)) - float(oracle.value(to))) <= 1e-5 * abs(float(oracle.value(to)))
    for a, b in zip(pd, po):
        assert relerr(a.value.to_numpy(), b.value) <= 2e-5


@pytest.mark.parametrize("rows,cols", [(1024, 1024), (64, 4096), (4096, 64)])
def test_sweep(ag, oracle, rows, cols):
    W = ag.workloads
    x, brow, bcol = W.sweep_init(np.random.default_rng(3), rows, cols)
    lo, go, ld, gd = both(ag, oracle, lambda m, p: W.sweep_loss(m, p[0], p[1], p[2]), [x, brow, bcol], [], np.float32)
    assert abs(lo - ld) <= 1e-5 * abs(lo)
    for a, b in zip(gd, go):
        assert relerr(a, b) <= 1e-5


def test_lstm(ag, oracle):
    W = ag.workloads
    rng = np.random.default_rng(4)
    H, E, V, B, T = 64, 32, 40, 48, 4
    w = W.lstm_init(rng, H, E, V, np.float32)
    keys = list(w)
    ids = [rng.integers(0, V, B) for _ in range(T)]
    tg = []
    for t in range(T):
        oh = np.zeros((V, B), np.float32)
        oh[rng.integers(0, V, B), np.arange(B)] = 1
        tg.append(np.asfortranarray(oh))
    h0, c0 = np.zeros((H, B), np.float32), np.zeros((H, B), np.float32)

    def loss(m, p, h0, c0, *tg):
        return W.lstm_loss(m, dict(zip(keys, p)), ids, list(tg), h0, c0)
    lo, go, ld, gd = both(ag, oracle, loss, [w[k] for k in keys], [h0, c0] + tg, np.float32)
    assert abs(lo - ld) <= 1e-5 * abs(lo)
    for k, a, b in zip(keys, gd, go):
        assert relerr(a, b) <= 2e-5, k


def test_rnn_grad_of_grad(ag, oracle):
    """Config 5 pattern (test/highorder.jl:36-42): d/dW_hh of sum(abs2, d loss / d W_hh)."""
    W = ag.workloads
    rng = np.random.default_rng(6)
    H, V, B, T = 16, 10, 32, 3
    w = W.rnn_init(rng, H, V, np.float64)
    xs = []
    for t in range(T):
        oh = np.zeros((V, B))
        oh[rng.integers(0, V, B), np.arange(B)] = 1
        xs.append(np.asfortranarray(oh))
    y = np.zeros((V, B))
    y[rng.integers(0, V, B), np.arange(B)] = 1
    h0 = np.zeros((H, B))

    def make(m, dev):
        wd = [dev(a) for a in w]
        xd, yd, hd = [dev(a) for a in xs], dev(y), dev(h0)

        def loss(whh):
            return W.rnn_loss(m, [wd[0], whh, wd[2], wd[3], wd[4]], xd, yd, hd)
        inner = m.grad(loss)
        return m.grad(lambda whh: m.sumabs2(inner(whh))), wd[1]
    fo, wo = make(oracle, lambda a: np.asfortranarray(a.copy()))
    fd, wdv = make(ag, ag.DArray.from_numpy)
    assert relerr(fd(wdv).to_numpy(), fo(wo)) <= 1e-10


def test_highorder_tanh(ag):
    x = ag.DArray.from_numpy(np.array([1.0]))
    g1 = ag.grad(lambda v: ag.sum_(ag.tanh(v)))
    g2 = ag.grad(lambda v: ag.sum_(g1(v)))
    t = np.tanh(1.0)
    assert abs(g1(x).to_numpy()[0] - (1 - t * t)) < 1e-15
    assert abs(g2(x).to_numpy()[0] - (-2 * t * (1 - t * t))) < 1e-14


def test_gradcheck_primitives(ag):
    """test/gradcheck.jl on device for the hot-path primitives (reference tolerances)."""
    rng = np.random.default_rng(8)
    x = ag.DArray.from_numpy(rng.uniform(0.2, 0.8, (3, 2)))
    for name in ["exp", "log", "tanh", "sigm", "sqrt", "sin", "cos", "abs2", "abs_", "neg", "relu"]:
        assert ag.gradcheck(getattr(ag, name), x), name
    a = ag.DArray.from_numpy(rng.standard_normal((4, 3)))
    b = ag.DArray.from_numpy(rng.standard_normal((3, 5)))
    assert ag.gradcheck(ag.matmul, a, b)
    v = ag.DArray.from_numpy(rng.uniform(0.5, 1.5, (4, 1)))
    for name in ["add", "sub", "mul", "div", "max_", "min_"]:
        assert ag.gradcheck(getattr(ag, name), a, v), name
    assert ag.gradcheck(lambda z: ag.sum_(z, dims=1), a)
    assert ag.gradcheck(lambda z: ag.getindex(z, slice(None), [1, 1, 2]), a)
    assert ag.gradcheck(lambda p, q: ag.vcat(p, q), a, ag.DArray.from_numpy(rng.standard_normal((2, 3))))


def test_gradcheck_every_primitive(ag):
    """north_star: gradients "checked with test/gradcheck.jl on every primitive" -- the whole catalogue of rules.py
    (src/math.jl:9-74, src/base.jl:20-49,73-74,202-250, src/statistics.jl:20-35, src/linearalgebra.jl:16,23,64-65,
    src/cat.jl, src/getindex.jl), each on a domain where it is differentiable, Float64, reference tolerances."""
    from gradcheck_catalogue import check_catalogue
    check_catalogue(ag)


def test_norm_p_against_normback(ag, oracle):
    """norm(x, p) for the branches of normback (src/linearalgebra.jl:222-238): device composites against the
    oracle's branch-for-branch restatement, values and gradients."""
    rng = np.random.default_rng(23)
    x = np.asfortranarray(rng.standard_normal((7, 9)))
    x[2, 3] = 0.0
    xd = ag.DArray.from_numpy(x)
    for p in (2, 1, np.inf, -np.inf, 0, 3, 2.5, 0.5):
        xx, xxd = (x, xd) if p != -np.inf else (x + np.sign(x) + (x == 0), ag.DArray.from_numpy(x + np.sign(x) + (x == 0)))
        vo, vd = float(oracle.norm(xx, p)), float(ag.norm(xxd, p))
        assert abs(vo - vd) <= 1e-12 * max(abs(vo), 1), p
        if p == 0.5:
            xx, xxd = np.abs(x) + 0.1, ag.DArray.from_numpy(np.abs(x) + 0.1)       # |x|^(p-1) at 0
        go = oracle.grad(lambda v: oracle.norm(v, p))(xx)
        gd = ag.grad(lambda v: ag.norm(v, p))(xxd)
        gd = np.zeros_like(xx) if gd is None else gd.to_numpy()
        assert relerr(gd, go) <= 1e-11 if np.any(go) else not np.any(gd), p


def test_graph_replay_matches_eager(ag):
    """CUDA-graph replay of a whole step gives bit-identical parameters to the eager step."""
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
    for _ in range(3):
        le = step(pe)
    step(pg)                                   # eager warm-up populates pool and workspace
    with ag.StepGraph() as g:
        lg = step(pg)
    g.launch()                                 # the capture itself did not execute: 2nd and 3rd step
    g.launch()
    ag.sync()
    for a, b in zip(pe, pg):
        assert np.array_equal(a.value.to_numpy(), b.value.to_numpy())
    assert float(le) == float(lg)
    g.destroy()


def test_lstm_graph_replay_with_device_indices(ag, oracle):
    """Config 4 with device-resident ids (IArray): the whole step (W[:, ids] gathers, Sparse gradient
    accumulation, one batched scatter-add, axpy! with a Sparse gradient) captures into one CUDA graph; replay
    with refreshed ids matches the eager host-driven tape bit for bit and the oracle within tolerance."""
    W = ag.workloads
    rng = np.random.default_rng(21)
    H, E, V, B, T = 32, 16, 20, 24, 3
    wd = W.lstm_init(rng, H, E, V, np.float32)
    keys = list(wd)
    ids_a = [rng.integers(0, V, B) for _ in range(T)]
    ids_b = [rng.integers(0, V, B) for _ in range(T)]
    tg = []
    for t in range(T):
        oh = np.zeros((V, B), np.float32)
        oh[rng.integers(0, V, B), np.arange(B)] = 1
        tg.append(np.asfortranarray(oh))
    tgd = [ag.DArray.from_numpy(a) for a in tg]
    h0 = ag.zeros((H, B), np.float32)

    def make():
        return {k: ag.Param(ag.DArray.from_numpy(wd[k])) for k in keys}

    def step(params, ids):
        t = ag.diff(lambda: W.lstm_loss(ag, params, ids, tgd, h0, h0))
        for k in keys:
            ag.axpy(-0.05, ag.grad(t, params[k]), params[k])
        return t
    # eager reference: two steps with host (numpy) indices
    pe = make()
    step(pe, ids_a)
    step(pe, ids_b)
    # graph: capture with device indices, replay twice with refreshed ids
    pg = make()
    dev_ids = [ag.IArray(i) for i in ids_a]
    warm = make()
    step(warm, dev_ids)                      # eager warm-up (pool, workspace)
    with ag.StepGraph() as g:
        step(pg, dev_ids)
    g.launch()
    for d, i in zip(dev_ids, ids_b):
        d.update(i)
    g.launch()
    ag.sync()
    for k in keys:
        assert np.array_equal(pe[k].value.to_numpy(), pg[k].value.to_numpy()), k
    g.destroy()
    # and against the oracle (one step from the initial weights)
    po = {k: oracle.Param(np.asfortranarray(wd[k].astype(np.float64))) for k in keys}
    to = oracle.diff(lambda: W.lstm_loss(oracle, po, ids_a, [a.astype(np.float64) for a in tg],
                                         np.zeros((H, B)), np.zeros((H, B))))
    p1 = make()
    t1 = ag.diff(lambda: W.lstm_loss(ag, p1, [ag.IArray(i) for i in ids_a], tgd, h0, h0))
    for k in keys:
        assert relerr(ag.full(ag.grad(t1, p1[k])).to_numpy(), oracle.full(oracle.grad(to, po[k]))) <= 2e-5, k


def test_multi_tensor_update_matches_axpy_loop(ag):
    """DataParallel.update on one rank (one multi-tensor launch) == the reference loop of axpy! calls."""
    W = ag.workloads
    w, x, y = W.mnist_init(np.random.default_rng(12), 300)
    xd, yd = ag.DArray.from_numpy(x), ag.DArray.from_numpy(y)
    pa = [ag.Param(ag.DArray.from_numpy(a)) for a in w]
    pb = [ag.Param(ag.DArray.from_numpy(a)) for a in w]
    ta = ag.diff(lambda: W.mnist_loss(ag, pa, xd, yd))
    for p in pa:
        ag.axpy(-0.1, ag.grad(ta, p), p)
    tb = ag.diff(lambda: W.mnist_loss(ag, pb, xd, yd))
    ag.DataParallel(0, 1).update(-0.1, [ag.grad(tb, p) for p in pb], pb)
    for a, b in zip(pa, pb):
        assert np.array_equal(a.value.to_numpy(), b.value.to_numpy())


def test_full_size_properties(ag):
    """BASELINE config 2 at the bench size (B = 65536): properties that need no oracle.
    (a) softmax cross-entropy: column sums of d loss / d logits are 0, so grad(b2) sums to ~0;
    (b) linearity: grads of 2*loss are exactly 2x; (c) grad(b1) == row sums of the relu-masked dy."""
    W = ag.workloads
    B = 65536
    w, x, y = W.mnist_init(np.random.default_rng(10), B)
    pd = [ag.Param(ag.DArray.from_numpy(a)) for a in w]
    xd, yd = ag.DArray.from_numpy(x), ag.DArray.from_numpy(y)
    t1 = ag.diff(lambda: W.mnist_loss(ag, pd, xd, yd))
    g1 = [ag.grad(t1, p).to_numpy() for p in pd]
    assert abs(g1[3].astype(np.float64).sum()) <= 1e-5
    t2 = ag.diff(lambda: ag.mul(2.0, W.mnist_loss(ag, pd, xd, yd)))
    g2 = [ag.grad(t2, p).to_numpy() for p in pd]
    for a, b in zip(g1, g2):
        assert np.array_equal(2 * a, b)
    # against a float64 numpy evaluation of the same formula (cheap: one pass on the host)
    w64 = [a.astype(np.float64) for a in w]
    a2 = w64[0] @ x.astype(np.float64) + w64[1][:, None]
    a3 = np.maximum(a2, 0)
    p = w64[2] @ a3 + w64[3][:, None]
    sm = np.exp(p - p.max(0))
    sm /= sm.sum(0)
    dp = (sm - y) / B
    assert relerr(g1[3], dp.sum(1)) <= 1e-5
    assert relerr(g1[2], dp @ a3.T) <= 1e-5
    # relu'(a2) is discontinuous at 0: take the mask from the device's own Float32 pre-activation so a
    # last-bit difference in a2 cannot flip an element between the two evaluations
    a2_dev = ag.add(ag.matmul(pd[0].value, xd), pd[1].value).to_numpy()
    da2 = (w64[2].T @ dp) * (a2_dev >= 0)
    assert relerr(g1[1], da2.sum(1)) <= 1e-5
    assert relerr(g1[0], da2 @ x.astype(np.float64).T) <= 1e-5


def test_sweep_full_size_properties(ag):
    """BASELINE config 3 at 2^26 elements (256 MB arrays): size-independent properties.
    (a) scaling x by 2 scales sum(abs2, .) by exactly 4 and its gradient by exactly 2 (powers of two are exact);
    (b) the broadcast gradients equal the row / column sums of the full gradient (unbroadcast, src/unbroadcast.jl:17-24);
    (c) deterministic: two runs are bit-identical."""
    rng = np.random.default_rng(31)
    rows, cols = 1024, 65536
    x = rng.standard_normal((rows, cols), dtype=np.float32)
    xd = ag.DArray.from_numpy(x)
    x2 = ag.mul(2.0, xd)
    g1 = ag.grad(lambda v: ag.sumabs2(v))(xd)
    g2 = ag.grad(lambda v: ag.sumabs2(v))(x2)
    assert float(ag.sumabs2(x2)) == 4.0 * float(ag.sumabs2(xd))
    assert np.array_equal(g2.to_numpy(), 2 * g1.to_numpy())
    ref = float(np.sum(x.astype(np.float64) ** 2))
    assert abs(float(ag.sumabs2(xd)) - ref) <= 1e-6 * ref
    brow = ag.Param(ag.DArray.from_numpy((0.1 * rng.standard_normal((1, cols))).astype(np.float32)))
    bcol = ag.Param(ag.DArray.from_numpy((0.1 * rng.standard_normal((rows, 1))).astype(np.float32)))
    px = ag.Param(xd)
    W = ag.workloads
    t = ag.diff(lambda: W.sweep_loss(ag, px, brow, bcol))
    gx = ag.grad(t, px).to_numpy().astype(np.float64)
    gcol = ag.grad(t, bcol).to_numpy()
    grow = ag.grad(t, brow).to_numpy()
    # d/dbrow = column sums of d/d(0.1x + brow) = 10 * column sums of gx
    assert relerr(grow, 10.0 * gx.sum(axis=0, keepdims=True)) <= 1e-5
    t2 = ag.diff(lambda: W.sweep_loss(ag, px, brow, bcol))
    assert np.array_equal(ag.grad(t2, bcol).to_numpy(), gcol)
    assert float(ag.value(t)) == float(ag.value(t2))


def test_sparse_scatter_full_size(ag):
    """Config 4's embedding gradient at full size: 100 blocks of (128 x 256) values, 25 600 column indices into a
    (128 x 128) container.  Touched-column set bit-exact, values within tolerance of a float64 host scatter,
    result bit-reproducible."""
    rng = np.random.default_rng(32)
    E, V, B, T = 128, 128, 256, 100
    Wd = ag.zeros((E, V), np.float32)
    ids = [rng.integers(0, V - 5, B) for _ in range(T)]          # the last 5 columns stay untouched
    vals = [rng.standard_normal((E, B)).astype(np.float32) for _ in range(T)]
    sp = ag.Sparse(Wd, [ag.DArray.from_numpy(v) for v in vals], [(slice(None), ag.IArray(i)) for i in ids])
    got = sp.full().to_numpy()
    ref = np.zeros((E, V))
    for i, v in zip(ids, vals):
        np.add.at(ref, (slice(None), i), v.astype(np.float64))
    assert np.array_equal(got[:, V - 5:], np.zeros((E, 5), np.float32))
    assert np.array_equal(np.any(got != 0, axis=0), np.any(ref != 0, axis=0))
    assert relerr(got, ref) <= 1e-5
    assert np.array_equal(got, sp.full().to_numpy())


@pytest.mark.parametrize("dtype,tol", [(np.float32, 1e-6), (np.float64, 1e-13)])
def test_clipped_update_on_device_and_in_a_graph(ag, dtype, tol):
    """Gradient clipping without a host round trip (ag_multi_sumabs2 + ag_multi_axpy_scaled): against numpy, and
    captured in a CUDA graph whose replay clips with the norm of the CURRENT gradients."""
    from test_host_logic import _clip_case
    for gclip in (5.0, 1e6, None):
        got, want, gn = _clip_case(ag, dtype, gclip)
        for a, b in zip(got, want):
            assert np.max(np.abs(a - b)) <= tol * max(np.max(np.abs(b)), 1), gclip
    rng = np.random.default_rng(42)
    w0 = rng.standard_normal((64, 33)).astype(dtype)
    g1, g2 = (rng.standard_normal((64, 33)) * 2).astype(dtype), (rng.standard_normal((64, 33)) * 0.01).astype(dtype)
    p = ag.Param(ag.DArray.from_numpy(w0.copy()))
    gd = ag.DArray.from_numpy(g1)
    dp = ag.DataParallel(0, 1)
    dp.update(-0.5, [gd], [p], gclip=3.0)                      # eager warm-up: w0 - 0.5*s1*g1
    with ag.StepGraph() as g:
        dp.update(-0.5, [gd], [p], gclip=3.0)
    g2_flat = np.ascontiguousarray(g2.reshape(-1, order="F"))       # must outlive the (asynchronous) copy
    gd.copy_from_host(g2_flat.ctypes.data, g2_flat.nbytes)
    ag.sync()
    g.launch()                                                 # clips with the norm of g2 (no clipping: small)
    ag.sync()
    s1 = min(1.0, 3.0 / np.sqrt(float((g1.astype(np.float64) ** 2).sum())))
    want = w0.astype(np.float64) - 0.5 * s1 * g1 - 0.5 * 1.0 * g2
    assert s1 < 1.0 and np.max(np.abs(p.value.to_numpy() - want)) <= 4 * tol * np.max(np.abs(want))
    g.destroy()
