"""Secondary BASELINE.json configs for bench.py (north_star: "throughput is reported on synthetic data of each
config's shape at 1, 2, 4 and 8 GPUs ... as a fraction of the HBM or tensor roofline"): configs[0] housing,
configs[2] broadcast/unbroadcast sweep, configs[3] charlm LSTM, configs[4] RNN grad-of-grad.  Each is the whole
@diff + grad (+ allreduce) + update step of the reference's example, captured once as a CUDA graph and replayed;
weak scaling (every rank runs the per-GPU shape; the LSTM sums its Param gradients over the ranks).
Also dp_check: a numeric check of the N-rank MNIST step against a float64 evaluation of the full batch.
"""
import time

import numpy as np


def _replay_ms(ag, g, reps, barrier, allmax):
    for _ in range(2):
        g.launch()
    barrier()
    l0 = ag.launch_count()
    e0, e1 = ag.Event(), ag.Event()
    e0.record()
    for _ in range(reps):
        g.launch()
    e1.record()
    ms = e1.elapsed_ms_since(e0) / reps
    launches = (ag.launch_count() - l0) // reps
    barrier()
    return allmax(ms), launches


def housing(ag, W, barrier, allmax, world):
    """configs[0]: examples/housing.jl linear regression, 506 x 13 Float64, sum(abs2, w*x .+ b - y)/n, axpy! SGD."""
    w, x, y = W.housing_init(np.random.default_rng(1))
    params = [ag.Param(ag.DArray.from_numpy(a)) for a in w]
    xd, yd = ag.DArray.from_numpy(x), ag.DArray.from_numpy(y)

    def step():
        t = ag.diff(lambda: W.housing_loss(ag, params, xd, yd))
        g = [ag.grad(t, p) for p in params]
        del t
        for p, gi in zip(params, g):
            ag.axpy(-0.1, gi, p)
    for _ in range(2):
        step()
    with ag.StepGraph() as g:
        step()
    ms, nl = _replay_ms(ag, g, 500, barrier, allmax)
    g.destroy()
    return {"workload": "examples/housing.jl 506x13 Float64, @diff + grad + axpy! (every rank runs a replica)",
            "dtype": "f64", "ms_per_step": ms, "steps_per_s": world * 1e3 / ms, "launches_per_step": nl,
            "bound": "launch latency (13 x 506 problem): no roofline fraction", "scaling": "replicas"}


def sweep(ag, W, barrier, allmax, world, hbm_peak, rows=1024, cols=262144):
    """configs[2]: sum(abs2, tanh.(exp.(0.1 .* x .+ brow) .+ bcol)) on rows x cols Float32 per rank."""
    x, brow, bcol = W.sweep_init(np.random.default_rng(3), rows, cols)
    ps = [ag.Param(ag.DArray.from_numpy(a)) for a in (x, brow, bcol)]
    del x
    ag.set_gc_function(ag.release_node)
    try:
        def step():
            t = ag.diff(lambda: W.sweep_loss(ag, ps[0], ps[1], ps[2]))
            return [ag.grad(t, p) for p in ps]
        for _ in range(2):
            keep = step()
        with ag.StepGraph() as g:
            keep = step()
        ms, nl = _replay_ms(ag, g, 5, barrier, allmax)
        g.destroy()
    finally:
        ag.set_gc_function(lambda n, t: None)
    n = rows * cols
    algo = 4.0 * n * 25            # unfused tape: forward 6 passes (12 N s) + backward 13 N s (SURVEY.md 8d per-rule figures)
    return {"workload": "broadcast/unbroadcast sweep, x %dx%d Float32 (2^%d elements) per GPU, @diff + grad" % (
                rows, cols, int(np.log2(n))),
            "dtype": "f32", "ms_per_step": ms, "launches_per_step": nl, "algorithmic_bytes": algo,
            "algorithmic_GBps": world * algo / (ms * 1e-3) / 1e9,
            "roofline_frac_per_gpu": algo / (ms * 1e-3) / 1e9 / hbm_peak,
            "note": "algorithmic bytes are those of the UNFUSED tape; the fused step moves fewer, so the fraction may exceed 1",
            "scaling": "weak"}


def lstm(ag, W, dp, barrier, allmax, world, peaks, H=512, E=128, V=128, B=256, T=100):
    """configs[3]: examples/charlm.jl LSTM, embedding as W[:, ids] (Sparse gradient), gclip = 5 update."""
    rng = np.random.default_rng(4 + dp.rank)
    wd = W.lstm_init(np.random.default_rng(4), H, E, V, np.float32)
    params = {k: ag.Param(ag.DArray.from_numpy(v)) for k, v in wd.items()}
    plist = list(params.values())
    ids = [ag.IArray(rng.integers(0, V, B)) for _ in range(T)]
    tg = []
    for _ in range(T):
        oh = np.zeros((V, B), np.float32)
        oh[rng.integers(0, V, B), np.arange(B)] = 1
        tg.append(ag.DArray.from_numpy(np.asfortranarray(oh)))
    h0 = ag.zeros((H, B), np.float32)
    ag.set_gc_function(ag.release_node)
    try:
        def step():
            t = ag.diff(lambda: W.lstm_loss(ag, params, ids, tg, h0, h0, global_batch=B * world))
            g = [ag.grad(t, p) for p in plist]
            del t
            dp.update(-0.01, g, plist, gclip=5.0)
        step()
        ag.sync()
        t0 = time.perf_counter()
        step()
        ag.sync()
        eager_ms = (time.perf_counter() - t0) * 1e3
        with ag.StepGraph() as g:
            step()
        ms, nl = _replay_ms(ag, g, 5, barrier, allmax)
        g.destroy()
    finally:
        ag.set_gc_function(lambda n, t: None)
    fwd = T * (4 * 2.0 * H * (E + H) * B + 2.0 * V * H * B)
    flops = 3.0 * fwd                                  # backward = two products per forward product
    tf = flops / (ms * 1e-3) / 1e12
    tf32_peak = float(peaks.get("bf16_tflops", 1625.0)) / 2.0
    # HBM: per timestep the tape's arrays (SURVEY.md 8d per-rule figures, unfused): dominated by (H,B) elementwise
    # arrays, ~60 of them per step forward + backward, plus the weight reads of the products
    algo_bytes = T * (60.0 * H * B * 4 + 3 * 4 * 4.0 * H * (E + H))
    return {"workload": "examples/charlm.jl LSTM H=%d E=%d V=%d T=%d, batch %d per GPU, Float32, W[:, ids] embedding "
                        "(Sparse gradient), @diff + grad + allreduce + gclip=5 update" % (H, E, V, T, B),
            "dtype": "f32", "ms_per_step": ms, "steps_per_s": world * 1e3 / ms, "launches_per_step": nl,
            "eager_host_driven_ms": eager_ms, "flops_per_step": flops, "tflops_per_gpu": tf,
            "frac_of_tf32_peak": tf / tf32_peak, "frac_of_3xtf32_peak": tf / (tf32_peak / 3.0),
            "tf32_peak_tflops": tf32_peak, "peak_source": "MEASURED_PEAKS.json bf16_tflops / 2 (nominal tf32 : bf16 ratio)",
            "approx_algorithmic_bytes": algo_bytes,
            "approx_hbm_frac": algo_bytes / (ms * 1e-3) / 1e9 / float(peaks.get("hbm_gbs", 6650.0)),
            "allreduce": ("NCCL (5.6 MB bucket)" if world > 1 else None), "scaling": "weak"}


def rnn_grad_of_grad(ag, W, barrier, allmax, world, peaks, H=256, V=100, T=15, B=4096):
    """configs[4]: examples/rnn_lonely_integer.jl RNN under test/highorder.jl:36-42: d/dW_hh sum(abs2, dloss/dW_hh)."""
    rng = np.random.default_rng(6)
    w = [ag.DArray.from_numpy(a) for a in W.rnn_init(rng, H, V, np.float32)]
    xs = []
    for _ in range(T):
        oh = np.zeros((V, B), np.float32)
        oh[rng.integers(0, V, B), np.arange(B)] = 1
        xs.append(ag.DArray.from_numpy(np.asfortranarray(oh)))
    yy = np.zeros((V, B), np.float32)
    yy[rng.integers(0, V, B), np.arange(B)] = 1
    yd, hh = ag.DArray.from_numpy(np.asfortranarray(yy)), ag.zeros((H, B), np.float32)
    inner = ag.grad(lambda whh: W.rnn_loss(ag, [w[0], whh, w[2], w[3], w[4]], xs, yd, hh))
    outer = ag.grad(lambda whh: ag.sumabs2(inner(whh)))
    keep = {}

    def step():
        keep["g"] = outer(w[1])
    step()
    step()
    ag.sync()
    t0 = time.perf_counter()
    for _ in range(3):
        step()
    ag.sync()
    eager_ms = (time.perf_counter() - t0) / 3 * 1e3
    out = {"workload": "examples/rnn_lonely_integer.jl RNN H=%d V=%d T=%d batch %d Float32, grad of sum(abs2, grad) "
                       "w.r.t. W_hh (every inner rule and addto! recorded on the outer tape)" % (H, V, T, B),
           "dtype": "f32", "eager_host_driven_ms": allmax(eager_ms), "scaling": "replicas"}
    try:
        with ag.StepGraph() as g:
            step()
        ms, nl = _replay_ms(ag, g, 10, barrier, allmax)
        g.destroy()
        # products: forward T*(2HVB + 2HHB) + 2VHB; first backward 2x; second backward ~2x the first two together
        fwd = T * (2.0 * H * V * B + 2.0 * H * H * B) + 2.0 * V * H * B
        flops = fwd * (1 + 2 + 6)
        tf = flops / (ms * 1e-3) / 1e12
        tf32_peak = float(peaks.get("bf16_tflops", 1625.0)) / 2.0
        out.update({"ms_per_step": ms, "steps_per_s": world * 1e3 / ms, "launches_per_step": nl,
                    "approx_flops_per_step": flops, "tflops_per_gpu": tf, "frac_of_tf32_peak": tf / tf32_peak,
                    "frac_of_3xtf32_peak": tf / (tf32_peak / 3.0), "replay": "CUDA graph of the double backward"})
    except Exception as exc:            # noqa: BLE001  (reported, not hidden)
        out["graph_capture_error"] = str(exc)[:200]
    return out


def dp_check(ag, W, dp, rank, world, allgather, batch=256):
    """One untimed CHECKED data-parallel MNIST step at a small batch: every rank's Params after the step must be
    bit-identical across ranks (sha1 of their bytes, gathered over gloo) and within 1e-5 (norm-wise) of a float64
    numpy evaluation of the FULL-batch step (closed-form softmax cross-entropy gradients; independent of oracle/)."""
    import hashlib
    w_np, _, _ = W.mnist_init(np.random.default_rng(1), 1)
    shards = [W.mnist_from_ubyte(*W.mnist_ubyte_batch(np.random.default_rng(2000 + r), batch)) for r in range(world)]
    x_np, y_np = shards[rank]
    params = [ag.Param(ag.DArray.from_numpy(a)) for a in w_np]
    xd, yd = ag.DArray.from_numpy(x_np), ag.DArray.from_numpy(y_np)
    state = {}
    W.mnist_train_step(ag, dp, params, xd, yd, 0.1, state, global_batch=batch * world)
    ag.sync()
    got = [p.value.to_numpy() for p in params]
    digest = hashlib.sha1(b"".join(np.ascontiguousarray(a).tobytes() for a in got)).hexdigest()
    digests = allgather(digest) if world > 1 else [digest]
    # float64 full-batch step
    X = np.concatenate([s[0] for s in shards], axis=1).astype(np.float64)
    Y = np.concatenate([s[1] for s in shards], axis=1).astype(np.float64)
    w1, b1, w2, b2 = [a.astype(np.float64) for a in w_np]
    a2 = w1 @ X + b1[:, None]
    a3 = np.maximum(a2, 0)
    p = w2 @ a3 + b2[:, None]
    sm = np.exp(p - p.max(0))
    sm /= sm.sum(0)
    dpv = (sm - Y) / X.shape[1]
    da2 = (w2.T @ dpv) * (a2 >= 0)
    grads = [da2 @ X.T, da2.sum(1), dpv @ a3.T, dpv.sum(1)]
    want = [w - 0.1 * g for w, g in zip((w1, b1, w2, b2), grads)]
    # the summed gradients (dp.update leaves the reduced gradients in state["grads"]) against float64, norm-wise, and
    # the updated Params against w - lr*g (Float32 rounding of the stored Param: relative to the Param)
    gd = [ag.full(g).to_numpy().astype(np.float64).reshape(r.shape) for g, r in zip(state["grads"], grads)]
    err_g = max(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300) for a, b in zip(gd, grads))
    err_w = max(np.max(np.abs(a.astype(np.float64) - b)) / max(np.max(np.abs(b)), 1e-300) for a, b in zip(got, want))
    same = len(set(digests)) == 1
    return {"ranks": world, "batch_per_rank": batch, "identical_across_ranks": same,
            "max_rel_err": float(err_g), "max_rel_err_params": float(err_w), "tolerance": 1e-5,
            "reference": "float64 numpy evaluation of the full-batch step (closed-form softmax cross-entropy gradients)",
            "ok": bool(same and err_g <= 1e-5 and err_w <= 1e-6),
            "collective": "ag_p2p_allreduce_multi (fused axpy!)" if (world > 1 and dp.p2p_cap) else ("NCCL" if world > 1 else None)}
