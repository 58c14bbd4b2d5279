"""Timing of the secondary BASELINE configs on the GPU box (eager, host-driven tape): config 3 sweep, config 4
charlm LSTM (H=512, T=100, B=256, E=128, V=128), config 5 RNN grad-of-grad (H=256, V=100, T=15, B=4096)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
ag = ge.load_package()
ag.init(0)
W = ag.workloads
rng = np.random.default_rng(0)


def timeit(fn, reps=3):
    fn(); fn(); ag.sync()
    l0 = ag.launch_count()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    ag.sync()
    return (time.perf_counter() - t0) / reps * 1e3, (ag.launch_count() - l0) // reps


# ---- config 3: sweep at 2^24 and 2^28 elements
for rows, cols in [(1024, 16384), (1024, 262144)]:
    x, brow, bcol = W.sweep_init(rng, rows, cols)
    ps = [ag.Param(ag.DArray.from_numpy(a)) for a in (x, brow, bcol)]
    def step():
        t = ag.diff(lambda: W.sweep_loss(ag, ps[0], ps[1], ps[2]))
        return [ag.grad(t, p) for p in ps]
    ms, nl = timeit(step, 5)
    n = rows * cols
    # algorithmic bytes: fwd 6 passes (r+w) + bwd: sumabs2 2N, tanh 3N, add alias+reduce N, exp 3N, add reduce N, mul 3N (s=4)
    bytes_ = 4.0 * n * (12 + 13)
    print("sweep %dx%d: %.3f ms/step, %d launches, ~%.0f GB/s algorithmic" % (rows, cols, ms, nl, bytes_ / ms / 1e6))
    del ps, x

# ---- config 4: charlm LSTM
H, E, V, B, T = 512, 128, 128, 256, 100
wd = W.lstm_init(rng, H, E, V, np.float32)
params = {k: ag.Param(ag.DArray.from_numpy(v)) for k, v in wd.items()}
ids = [rng.integers(0, V, B) for _ in range(T)]
tg = []
for t in range(T):
    oh = np.zeros((V, B), np.float32); oh[rng.integers(0, V, B), np.arange(B)] = 1
    tg.append(ag.DArray.from_numpy(np.asfortranarray(oh)))
h0 = ag.zeros((H, B), np.float32)
def lstm_step():
    t = ag.diff(lambda: W.lstm_loss(ag, params, ids, tg, h0, h0))
    for k, p in params.items():
        ag.axpy(-0.01, ag.grad(t, p), p)
    return t
ms, nl = timeit(lstm_step, 2)
print("charlm LSTM H=%d T=%d B=%d: %.1f ms/step (eager host-driven tape), %d launches/step" % (H, T, B, ms, nl))
ids_host = ids
ids = [ag.IArray(i) for i in ids_host]          # device-resident ids: the step captures into one CUDA graph
lstm_step(); ag.sync()
t0 = time.perf_counter()
with ag.StepGraph() as g:
    lstm_step()
print("  graph capture + instantiate: %.1f ms" % ((time.perf_counter() - t0) * 1e3))
for _ in range(2):
    g.launch()
ag.sync()
e0, e1 = ag.Event(), ag.Event()
e0.record()
for _ in range(5):
    g.launch()
e1.record()
print("  CUDA-graph replay: %.2f ms/step" % (e1.elapsed_ms_since(e0) / 5))
g.destroy()
ids = ids_host
t = lstm_step()
print("  loss %.4f  grad(W_embedding) type %s" % (float(ag.value(t)), type(ag.grad(t, params["W_embedding"])).__name__))

# ---- config 5: RNN grad of grad
Hh, Vv, Tt, Bb = 256, 100, 15, 4096
w = [ag.DArray.from_numpy(a) for a in W.rnn_init(rng, Hh, Vv, np.float32)]
xs = []
for t in range(Tt):
    oh = np.zeros((Vv, Bb), np.float32); oh[rng.integers(0, Vv, Bb), np.arange(Bb)] = 1
    xs.append(ag.DArray.from_numpy(np.asfortranarray(oh)))
yy = np.zeros((Vv, Bb), np.float32); yy[rng.integers(0, Vv, Bb), np.arange(Bb)] = 1
yd = ag.DArray.from_numpy(np.asfortranarray(yy)); hh = ag.zeros((Hh, Bb), np.float32)
inner = ag.grad(lambda whh: W.rnn_loss(ag, [w[0], whh, w[2], w[3], w[4]], xs, yd, hh))
outer = ag.grad(lambda whh: ag.sumabs2(inner(whh)))
ms, nl = timeit(lambda: outer(w[1]), 2)
print("RNN grad-of-grad H=%d T=%d B=%d: %.1f ms (eager), %d launches" % (Hh, Tt, Bb, ms, nl))
