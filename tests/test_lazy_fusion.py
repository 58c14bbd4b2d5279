"""CPU tests of the deferred-evaluation / horizontal-fusion host logic (autograd.jl_b200/lazy.py) on the numpy
stand-in for the C ABI: the bytecode the host emits must compute exactly what the unfused sequence computes,
values the tape still needs must be written, temporaries must not, and in-place updates must not be reordered
across pending reads.  The CUDA evaluator itself is checked by tests/test_gpu_fusion.py (-m gpu)."""
import numpy as np
import pytest


def dev(ag, a):
    return ag.DArray.from_numpy(np.asarray(a))


@pytest.fixture()
def lz(hostlib):
    import sys
    L = sys.modules["agb200.lazy"]
    prev = L.set_fusion(True)
    yield hostlib, L, sys.modules["agb200.darray"]
    L.flush()
    L.set_fusion(prev)


def run_step(ag, W, cfg, dtype, fused):
    import sys
    L = sys.modules["agb200.lazy"]
    prev = L.set_fusion(fused)
    try:
        rng = np.random.default_rng(7)
        if cfg == "mnist":
            w, x, y = W.mnist_init(rng, 37, dtype)
            loss = lambda p, xd, yd: W.mnist_loss(ag, p, xd, yd)
            consts = [x, y]
        elif cfg == "housing":
            w, x, y = W.housing_init(rng, dtype=dtype)
            loss = lambda p, xd, yd: W.housing_loss(ag, p, xd, yd)
            consts = [x, y]
        else:
            x, brow, bcol = W.sweep_init(rng, 24, 20, dtype)
            w = [brow, bcol]
            loss = lambda p, xd: W.sweep_loss(ag, xd, p[0], p[1])
            consts = [x]
        p = [ag.Param(dev(ag, a)) for a in w]
        c = [dev(ag, a) for a in consts]
        before = ag._lib.lib.launches if hasattr(ag._lib.lib, "launches") else 0
        t = ag.diff(lambda: loss(p, *c))
        g = [ag.full(ag.grad(t, q)).to_numpy() for q in p]
        v = float(ag.value(t))
        for q in p:
            ag.axpy(-0.1, ag.grad(t, q), q)
        after = [ag.value(q).to_numpy() for q in p]
        n = ag._lib.lib.launches - before
        return v, g, after, n
    finally:
        L.set_fusion(prev)


@pytest.mark.parametrize("cfg", ["mnist", "housing", "sweep"])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_fused_step_equals_unfused(hostlib, cfg, dtype):
    ag, W = hostlib, hostlib.workloads
    v0, g0, a0, n0 = run_step(ag, W, cfg, dtype, False)
    v1, g1, a1, n1 = run_step(ag, W, cfg, dtype, True)
    assert v0 == v1
    for x, y in zip(g0 + a0, g1 + a1):
        assert np.array_equal(x, y)
    assert n1 < n0, "fusion must need fewer launches (%d vs %d)" % (n1, n0)


def test_temporaries_are_not_stored_and_kept_values_are(lz):
    ag, L, D = lz
    rng = np.random.default_rng(0)
    x = dev(ag, rng.standard_normal((6, 5)).astype(np.float32))
    b = dev(ag, rng.standard_normal((6,)).astype(np.float32))
    r = dev(ag, rng.standard_normal((1, 5)).astype(np.float32))
    lib = ag._lib.lib
    n0, f0 = lib.launches, L.stats["fused_launches"]
    kept = D.binary_fwd("add", x, b)                        # something else refers to it: must be written
    y = D.unary_fwd("tanh", D.binary_fwd("mul", D.unary_fwd("exp", kept), r))   # exp and mul are temporaries
    assert L.is_pending(y) and L.is_pending(kept) and lib.launches == n0
    got = y.to_numpy()
    assert lib.launches == n0 + 1 and L.stats["fused_launches"] == f0 + 1
    assert not L.is_pending(kept), "a value with an outside reference is an output of the same launch"
    xe = x.to_numpy() + b.to_numpy()[:, None]
    assert np.array_equal(kept.to_numpy(), xe.astype(np.float32))
    ref = np.tanh((np.exp(xe.astype(np.float32)) * r.to_numpy()).astype(np.float32)).astype(np.float32)
    assert np.allclose(got, ref, rtol=1e-6, atol=0)
    assert lib.launches == n0 + 1


def test_shared_subexpression_uses_a_temporary(lz):
    ag, L, D = lz
    x = dev(ag, np.linspace(-1, 1, 12).reshape(3, 4))
    lib = ag._lib.lib
    e = D.unary_fwd("exp", x)
    y = D.binary_fwd("mul", D.binary_fwd("add", e, 1.0), D.binary_fwd("sub", e, 0.5))
    del e
    n0 = lib.launches
    got = y.to_numpy()
    assert lib.launches == n0 + 1
    ex = np.exp(x.to_numpy())
    assert np.array_equal(got, (ex + 1.0) * (ex - 0.5))


def test_inplace_update_does_not_overtake_pending_reads(lz):
    ag, L, D = lz
    w = dev(ag, np.ones((4, 3)))
    g = dev(ag, np.full((4, 3), 2.0))
    y = D.binary_fwd("mul", w, 3.0)          # pending: reads w
    D.axpy_(-1.0, g, w)                      # in place: w becomes -1
    assert np.array_equal(y.to_numpy(), np.full((4, 3), 3.0))
    assert np.array_equal(w.to_numpy(), np.full((4, 3), -1.0))
    z = D.unary_fwd("abs2", w)
    D.scale_(2.0, w)
    assert np.array_equal(z.to_numpy(), np.ones((4, 3)))


def test_errors_surface_at_the_call(lz):
    ag, L, D = lz
    a, b = dev(ag, np.ones((3, 4))), dev(ag, np.ones((2, 4)))
    with pytest.raises(ag.DimensionMismatch):
        D.binary_fwd("add", a, b)
    with pytest.raises(ag.DimensionMismatch):
        D.add(a, dev(ag, np.ones((4, 3))))
    with pytest.raises(ag.UnsupportedOnDevice):
        D.binary_fwd("add", a, dev(ag, np.ones((3, 4), np.float32)))
    with pytest.raises(ag.UnsupportedOnDevice):
        D.binary_bwd("eq", 1, a, a, a, a)


def test_unfusable_layout_falls_back_to_single_kernels(lz):
    ag, L, D = lz
    rng = np.random.default_rng(1)
    x = rng.standard_normal((3, 4, 5))
    m = rng.standard_normal((1, 4, 1))       # broadcast along the first and the last dim: not a row/column vector
    f0 = L.stats["fallback_nodes"]
    y = D.unary_fwd("exp", D.binary_fwd("mul", dev(ag, x), dev(ag, m)))
    assert np.allclose(y.to_numpy(), np.exp(x * m), rtol=1e-14)
    assert L.stats["fallback_nodes"] > f0


def test_many_operands_split_into_several_launches(lz):
    ag, L, D = lz
    rng = np.random.default_rng(2)
    xs = [rng.standard_normal((5, 7)) for _ in range(13)]
    ds = [dev(ag, a) for a in xs]
    acc = ds[0]
    for a in ds[1:]:
        acc = D.binary_fwd("add", acc, a)
    s0 = L.stats["splits"]
    ref = xs[0]
    for a in xs[1:]:
        ref = ref + a
    assert np.array_equal(acc.to_numpy(), ref)
    assert L.stats["splits"] > s0


def test_long_chain_is_cut_by_the_node_cap(lz):
    ag, L, D = lz
    x = dev(ag, np.full((4, 4), 0.5))
    y = x
    for _ in range(100):
        y = D.unary_fwd("sin", y)
    ref = np.full((4, 4), 0.5)
    for _ in range(100):
        ref = np.sin(ref)
    assert np.array_equal(y.to_numpy(), ref)


def test_gradient_maps_and_scalar_operands(lz):
    ag, L, D = lz
    rng = np.random.default_rng(3)
    x1, x2, dy = (rng.random((4, 6)) + 0.5 for _ in range(3))
    d = [dev(ag, v) for v in (x1, x2, dy)]
    y = D.binary_fwd("pow", d[0], d[1])
    cases = {
        ("mul", 1): dy * x2, ("mul", 2): dy * x1, ("div", 1): dy / x2, ("div", 2): -dy * x1 / (x2 * x2),
        ("pow", 1): dy * x2 * x1 ** (x2 - 1), ("pow", 2): dy * (x1 ** x2) * np.log(x1),
        ("sub", 2): -dy, ("add", 1): dy, ("atan2", 1): dy * x2 / (x1 * x1 + x2 * x2),
    }
    for (op, argno), ref in cases.items():
        got = D.binary_bwd(op, argno, d[2], d[0], d[1], y)
        assert L.is_pending(got)
        assert np.allclose(got.to_numpy(), ref, rtol=1e-13), (op, argno)
    mx = D.binary_fwd("max", d[0], 1.0)
    gm = D.binary_bwd("max", 1, d[2], d[0], 1.0, mx)
    assert np.array_equal(gm.to_numpy(), dy * (np.maximum(x1, 1.0) == x1))
    s = ag.sum_(d[0])                                   # device scalar
    z = D.unary_bwd("tanh", s, None, y)                 # scalar dy broadcast over y
    assert np.allclose(z.to_numpy(), x1.sum() * (1 - (x1 ** x2) ** 2), rtol=1e-13)
    z2 = D.unary_bwd("abs2", 0.25, d[0], None)          # host Number dy
    assert np.array_equal(z2.to_numpy(), 0.25 * (2 * x1))
    f = D.full_like_shape((4, 6), np.float64, 1.5)
    assert L.is_pending(f)
    assert np.array_equal(D.binary_fwd("mul", f, d[0]).to_numpy(), 1.5 * x1)
    e = D.bcast_to(dev(ag, x1[:1, :]), (4, 6))
    assert np.array_equal(e.to_numpy(), np.broadcast_to(x1[:1, :], (4, 6)))


def test_flush_evaluates_only_what_is_still_referenced(lz):
    ag, L, D = lz
    x = dev(ag, np.arange(6.0).reshape(2, 3))
    lib = ag._lib.lib
    a = D.unary_fwd("exp", x)
    b = D.unary_fwd("neg", a)
    del a
    n0 = lib.launches
    ag.sync()
    assert lib.launches == n0 + 1 and not L.is_pending(b)
    assert L.pending_count() == 0
    D.unary_fwd("exp", x)                                # result dropped at once: never evaluated
    ag.sync()
    assert lib.launches == n0 + 1


def test_higher_order_through_deferred_rules(lz, oracle):
    ag, L, D = lz
    x = np.array([0.3, -0.2, 0.9])

    def f(m, v):
        return m.sum_(m.tanh(m.mul(v, v)))
    g2 = ag.grad(lambda v: ag.sum_(ag.grad(lambda u: f(ag, u))(v)))(dev(ag, x))
    r2 = oracle.grad(lambda v: oracle.sum_(oracle.grad(lambda u: f(oracle, u))(v)))(x)
    assert np.allclose(g2.to_numpy(), r2, rtol=1e-12)


def test_large_chain_without_specialised_kernel_uses_single_kernels(lz):
    """From lazy.BIG elements on, a chain the library has no ahead-of-time kernel for is evaluated by the single
    kernels (each at the memory roofline) instead of the issue-bound interpreter; small chains still fuse."""
    ag, L, D = lz
    lib = ag._lib.lib
    lib.all_static = False
    L._static_cache.clear()
    try:
        n = L.BIG
        x = dev(ag, np.linspace(-1, 1, n).astype(np.float32).reshape(1024, n // 1024))
        f0, e0, b0 = L.stats["fused_launches"], L.stats["fallback_nodes"], L.stats.get("big_unspecialised", 0)
        y = D.unary_fwd("tanh", D.binary_fwd("mul", D.unary_fwd("exp", x), 0.5))
        got = y.to_numpy()
        assert L.stats["fused_launches"] == f0 and L.stats["fallback_nodes"] == e0 + 3
        assert L.stats.get("big_unspecialised", 0) > b0
        ref = np.tanh((np.exp(x.to_numpy()) * np.float32(0.5)).astype(np.float32)).astype(np.float32)
        assert np.allclose(got, ref, rtol=1e-6)
        small = dev(ag, np.ones((8, 8), np.float32))
        z = D.unary_fwd("tanh", D.binary_fwd("mul", D.unary_fwd("exp", small), 0.5))
        z.to_numpy()
        assert L.stats["fused_launches"] == f0 + 1
    finally:
        lib.all_static = True
        L._static_cache.clear()


def test_specialised_program_list_is_up_to_date(agpkg):
    """csrc/fuse_static.inc must be what tools/gen_fused_programs.py generates from the current host logic
    (a stale list would silently send the configs' chains to the slower general interpreter)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "gen_fused_programs.py"), "--check"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, "run `python tools/gen_fused_programs.py` and rebuild\n" + r.stderr[-2000:]


# ---- round 2: deferred products, soft tape holds, raw writers ---------------------------------------------------

def test_product_bias_activation_is_one_launch_and_only_needed_values_are_stored(lz):
    """`max.(0, w*x .+ b)` (examples/mnist.jl:32): product, sum and relu go out as ONE ag_gemm_epilogue call that writes
    the sum (relu's rule reads it, src/math.jl:54) and the relu, never the product; `sigm.(w*x .+ b)`
    (examples/charlm.jl:74) writes only the activation.  Values are identical to the unfused sequence."""
    ag, L, D = lz
    rng = np.random.default_rng(3)
    w, b, x = (rng.standard_normal(s).astype(np.float32) for s in ((40, 30), (40,), (30, 50)))
    wd, bd, xd = ag.Param(dev(ag, w)), ag.Param(dev(ag, b)), dev(ag, x)
    fake = ag._lib.lib
    calls = []
    orig = fake.ag_gemm_epilogue

    def spy(dtype, ta, tb, M, N, K, A, lda, B, ldb, bias, act, Cp, C2p, ldc):
        calls.append((act, bool(getattr(C2p, "value", C2p))))
        return orig(dtype, ta, tb, M, N, K, A, lda, B, ldb, bias, act, Cp, C2p, ldc)
    fake.ag_gemm_epilogue = spy
    try:
        for fn, act_id, two in ((ag.relu, 28, True), (ag.sigm, 7, False), (ag.tanh, 6, False)):
            calls.clear()
            t = ag.diff(lambda: ag.sum_(fn(ag.add(ag.matmul(wd, xd), bd))))
            gw = ag.grad(t, wd).to_numpy()
            assert calls == [(act_id, two)], (fn, calls)
            prev = L.set_fusion(False)
            t0 = ag.diff(lambda: ag.sum_(fn(ag.add(ag.matmul(wd, xd), bd))))
            L.set_fusion(prev)
            assert float(ag.value(t)) == float(ag.value(t0))
            assert np.array_equal(gw, ag.grad(t0, wd).to_numpy())
    finally:
        fake.ag_gemm_epilogue = orig


def test_soft_held_product_is_recomputed_on_demand_and_dropped_after_the_sweep(lz):
    ag, L, D = lz
    rng = np.random.default_rng(4)
    w, b, x = (rng.standard_normal(s).astype(np.float32) for s in ((40, 30), (40,), (30, 50)))
    wd, bd, xd = ag.Param(dev(ag, w)), ag.Param(dev(ag, b)), dev(ag, x)
    seen = {}

    def f():
        m = ag.matmul(wd, xd)
        seen["m"] = m                               # the tracked product (a Result)
        return ag.sum_(ag.tanh(ag.add(m, bd)))
    t = ag.diff(f)
    assert L.stats.get("discarded", 0) >= 1
    with pytest.raises(RuntimeError):               # dropped at the end of the sweep: nobody held its value
        ag.value(seen["m"]).to_numpy()
    # held by the user during the forward pass -> stored (hard reference), equal to the plain product
    held = {}

    def g():
        m = ag.matmul(wd, xd)
        held["v"] = ag.value(m)
        return ag.sum_(ag.tanh(ag.add(m, bd)))
    ag.diff(g)
    assert np.allclose(held["v"].to_numpy(), w.astype(np.float64) @ x, rtol=1e-5, atol=1e-5)


def test_externally_held_intermediate_is_stored_exactly_once(lz):
    """The reference-count test that decides which values of a chain are written (calibrated at import, lazy._REF_BASE):
    a held intermediate is an output of the one fused launch, an unheld one is not."""
    ag, L, D = lz
    assert L._REFS_OK
    x = dev(ag, np.linspace(0.1, 1.0, 64).astype(np.float32))
    fake = ag._lib.lib
    nouts = []
    orig = fake.ag_ewise_fused

    def spy(dtype, code, ncode, consts, nconsts, nin, inp, in_mode, in_imm, nout, out, rows, cols):
        nouts.append(nout)
        return orig(dtype, code, ncode, consts, nconsts, nin, inp, in_mode, in_imm, nout, out, rows, cols)
    fake.ag_ewise_fused = spy
    try:
        held = D.unary_fwd("exp", x)
        y = D.unary_fwd("tanh", D.binary_fwd("mul", held, 2.0))
        y.ptr
        assert nouts == [2]                          # y and the held exp, not the unheld product
        held.ptr
        assert nouts == [2]                          # already stored: no second launch
        nouts.clear()
        z = D.unary_fwd("tanh", D.binary_fwd("mul", D.unary_fwd("exp", x), 2.0))
        z.ptr
        assert nouts == [1]
    finally:
        fake.ag_ewise_fused = orig


def test_slab_copy_into_an_existing_array_flushes_pending_reads(lz):
    """y = W .+ 1 pending, then a raw copy into W, then reading y must still see the OLD W."""
    ag, L, D = lz
    W = dev(ag, np.ones((4, 2), np.float32))
    src = dev(ag, np.full((4, 2), 7.0, np.float32))
    y = D.binary_fwd("add", W, 1.0)
    D.slab_copy(src, 8, 0, W, 8, 0, 1, 8, 1)
    assert np.array_equal(y.to_numpy(), np.full((4, 2), 2.0, np.float32))
    assert np.array_equal(W.to_numpy(), np.full((4, 2), 7.0, np.float32))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_column_programs_equal_separate_launches(hostlib, dtype):
    """With column fusion on (lazy.set_col_fusion) the softmax normaliser, what follows it and the unbroadcast of its
    gradient run as ag_ewise_colprog launches: same values (sums in double either way), fewer launches."""
    import sys
    ag, W = hostlib, hostlib.workloads
    L = sys.modules["agb200.lazy"]
    v0, g0, a0, n0 = run_step(ag, W, "mnist", dtype, True)
    prev = L.set_col_fusion(True)
    try:
        c0 = L.stats.get("col_launches", 0)
        v1, g1, a1, n1 = run_step(ag, W, "mnist", dtype, True)
        assert L.stats.get("col_launches", 0) > c0
        assert L.stats.get("col_fallbacks", 0) == 0
    finally:
        L.set_col_fusion(prev)
    tol = 2e-6 if dtype == np.float32 else 1e-14
    assert abs(v0 - v1) <= tol * abs(v0)
    for x, y in zip(g0 + a0, g1 + a1):
        assert np.max(np.abs(x - y)) <= tol * max(np.max(np.abs(x)), 1e-30)
    assert n1 < n0


def test_column_program_fallback_when_it_does_not_fit(lz):
    """A DAG with more column values than a column program holds is evaluated piecewise (column sums on their own)."""
    ag, L, _ = lz
    prev = L.set_col_fusion(True)
    try:
        rng = np.random.default_rng(3)
        x = dev(ag, rng.standard_normal((6, 9)))
        acc = None
        for k in range(10):                           # ten column sums feeding one element-space expression
            s = ag.sum_(ag.mul(ag.exp(ag.mul(x, 0.1 * (k + 1))), 1.0), dims=0)
            t = ag.sub(x, s)
            acc = t if acc is None else ag.add(acc, t)
        got = acc.to_numpy()
        xn = x.to_numpy()
        ref = sum(xn - np.exp(xn * (0.1 * (k + 1))).sum(axis=0, keepdims=True) for k in range(10))
        assert np.allclose(got, ref, rtol=1e-12, atol=1e-12)
    finally:
        L.set_col_fusion(prev)


def test_deferred_reductions_keep_program_order(lz):
    """A stand-alone sum(x; dims) is deferred (kind "rs"); mutating x in place afterwards must not change it."""
    ag, L, D = lz
    rng = np.random.default_rng(4)
    x = dev(ag, rng.standard_normal((300, 20)))
    y = dev(ag, rng.standard_normal((300, 20)))
    xn = x.to_numpy().copy()
    s = ag.sum_(x, dims=1)
    assert L.is_pending(s) and s.kind == "rs"
    t = ag.sum_(y, dims=1)
    b0 = L.stats.get("reduction_batches", 0)
    ag.axpy(2.0, y, x)                                  # flushes: both reductions run (one batch) on the old x
    assert L.stats.get("reduction_batches", 0) == b0 + 1
    assert not L.is_pending(s) and not L.is_pending(t)
    assert np.allclose(s.to_numpy(), xn.sum(axis=1, keepdims=True), rtol=1e-12, atol=1e-12)
    prev = L.batch_reductions[0]
    L.batch_reductions[0] = False
    try:
        assert not L.is_pending(ag.sum_(y, dims=1))
    finally:
        L.batch_reductions[0] = prev
