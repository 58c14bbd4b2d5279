"""Config 4 only (charlm LSTM, H=512 E=128 V=128 B=256, T from env, default 100): one step eager + as a CUDA graph.
Run under `ncu --metrics gpu__time_duration.sum` with GRAPH=0 for the per-kernel launch list of one step."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
ag = ge.load_package()
ag.init(0)
W = ag.workloads
rng = np.random.default_rng(0)
H, E, V, B, T = 512, 128, 128, 256, int(os.environ.get("T", 100))
wd = W.lstm_init(rng, H, E, V, np.float32)
params = {k: ag.Param(ag.DArray.from_numpy(v)) for k, v in wd.items()}
ids = [ag.IArray(rng.integers(0, V, B)) for _ in range(T)]
tg = []
for t in range(T):
    oh = np.zeros((V, B), np.float32); oh[rng.integers(0, V, B), np.arange(B)] = 1
    tg.append(ag.DArray.from_numpy(np.asfortranarray(oh)))
h0 = ag.zeros((H, B), np.float32)
if os.environ.get("GC", "1") == "1":
    ag.set_gc_function(ag.release_node)


GCLIP = float(os.environ.get("GCLIP", 5.0))      # examples/charlm.jl:148-152 (0: plain axpy! loop)
dp = ag.DataParallel(0, 1)
plist = list(params.values())


def lstm_step():
    t = ag.diff(lambda: W.lstm_loss(ag, params, ids, tg, h0, h0))
    if GCLIP > 0:
        dp.update(-0.01, [ag.grad(t, p) for p in plist], plist, gclip=GCLIP)
    else:
        for k, p in params.items():
            ag.axpy(-0.01, ag.grad(t, p), p)
    return t


lstm_step(); ag.sync()
l0 = ag.launch_count()
t0 = time.perf_counter()
lstm_step(); ag.sync()
print("eager: %.1f ms, %d launches" % ((time.perf_counter() - t0) * 1e3, ag.launch_count() - l0))
if os.environ.get("GRAPH", "1") == "1":
    with ag.StepGraph() as g:
        lstm_step()
    for _ in range(2):
        g.launch()
    ag.sync()
    e0, e1 = ag.Event(), ag.Event()
    e0.record()
    for _ in range(5):
        g.launch()
    e1.record()
    print("graph replay: %.2f ms/step" % (e1.elapsed_ms_since(e0) / 5))
    g.destroy()
lz = sys.modules["agb200.lazy"]
print("lazy stats", lz.stats)
