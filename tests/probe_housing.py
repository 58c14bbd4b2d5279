"""Config 1 (examples/housing.jl linear regression, 506x13 Float64): launch-bound on a GPU; steps/s eager and as a graph."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
ag = ge.load_package()
ag.init(0)
W = ag.workloads
w, x, y = W.housing_init(np.random.default_rng(1))
params = [ag.Param(ag.DArray.from_numpy(a)) for a in w]
xd, yd = ag.DArray.from_numpy(x), ag.DArray.from_numpy(y)


def step():
    t = ag.diff(lambda: W.housing_loss(ag, params, xd, yd))
    for p in params:
        ag.axpy(-0.1, ag.grad(t, p), p)
    return ag.value(t)


for _ in range(3):
    step()
ag.sync()
l0 = ag.launch_count()
t0 = time.perf_counter()
for _ in range(200):
    loss = step()
ag.sync()
dt = (time.perf_counter() - t0) / 200
print("housing eager: %.1f us/step (%.0f steps/s), %d launches/step, loss %.6f" % (dt * 1e6, 1 / dt, (ag.launch_count() - l0) // 200, float(loss)))
with ag.StepGraph() as g:
    loss = step()
for _ in range(5):
    g.launch()
e0, e1 = ag.Event(), ag.Event()
e0.record()
for _ in range(1000):
    g.launch()
e1.record()
ms = e1.elapsed_ms_since(e0) / 1000
print("housing graph replay: %.2f us/step (%.0f steps/s), loss %.6f" % (ms * 1e3, 1e3 / ms, float(loss)))
g.destroy()
