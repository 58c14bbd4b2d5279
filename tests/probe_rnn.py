"""Config 5 only (RNN grad-of-grad, H=256 V=100 T=15 B=4096): eager time and launches; run under ncu for per-kernel times."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
ag = ge.load_package()
ag.init(0)
W = ag.workloads
rng = np.random.default_rng(0)
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
reps = int(os.environ.get("REPS", 3))
outer(w[1]); outer(w[1]); ag.sync()
l0 = ag.launch_count()
t0 = time.perf_counter()
for _ in range(reps):
    r = outer(w[1])
t_host = time.perf_counter() - t0
ag.sync()
t_all = time.perf_counter() - t0
lz = sys.modules["agb200.lazy"]
print("RNN grad-of-grad: %.1f ms/iter (host enqueue %.1f ms), %d launches, pool %s, lazy stats %s"
      % (t_all / reps * 1e3, t_host / reps * 1e3, (ag.launch_count() - l0) // reps, ag.device_info(), lz.stats))
