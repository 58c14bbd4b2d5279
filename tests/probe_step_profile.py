"""Per-call device time of one eager MNIST step (warm L2), by wrapping every C-ABI call in CUDA events.
Run on the GPU box: python tests/probe_step_profile.py [batch]"""
import ctypes as C
import os
import sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
ag = ge.load_package()
ag.init(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
W = ag.workloads
w, x, y = W.mnist_init(np.random.default_rng(1), B)
params = [ag.Param(ag.DArray.from_numpy(a)) for a in w]
xd, yd = ag.DArray.from_numpy(x), ag.DArray.from_numpy(y)


def step():
    t = ag.diff(lambda: W.mnist_loss(ag, params, xd, yd))
    for p in params:
        ag.axpy(-0.1, ag.grad(t, p), p)


for _ in range(3):
    step()
ag.sync()

real = ag._lib.lib
records = []


class Wrap:
    def __getattr__(self, name):
        f = getattr(real, name)
        if name in ("ag_event_create", "ag_event_record", "ag_event_elapsed_ms", "ag_event_destroy", "ag_alloc", "ag_free",
                    "ag_last_error", "ag_sync", "ag_launch_count"):
            return f

        def g(*a):
            e0, e1 = ag.Event(), ag.Event()
            e0.record()
            r = f(*a)
            e1.record()
            tag = name
            if name in ("ag_unary_fwd", "ag_unary_bwd", "ag_binary_fwd", "ag_binary_bwd", "ag_sum_dim", "ag_sum_all"):
                tag += ":%s" % (a[0],)
            if name == "ag_gemm":
                tag += ":%d%d m%d n%d k%d" % (a[1], a[2], a[3], a[4], a[5])
            if name in ("ag_unary_fwd", "ag_unary_bwd"):
                tag += " n=%d" % a[-1]
            if name == "ag_sum_dim":
                tag += " i%d r%d o%d" % (a[3], a[4], a[5])
            if name in ("ag_binary_fwd", "ag_binary_bwd"):
                shp = a[-1]
                tag += " shape=%s" % ([shp[i] for i in range(a[-2])],)
            records.append((tag, e0, e1))
            return r
        return g


import agb200.darray as D, agb200.graph as G, agb200.comm as Cm
w_ = Wrap()
for m in (ag._lib, D, G, Cm):
    m.lib = w_
records.clear()
step()
ag.sync()
tot = 0.0
for tag, e0, e1 in records:
    ms = e1.elapsed_ms_since(e0)
    tot += ms
    print("%8.1f us  %s" % (ms * 1e3, tag))
print("sum %.1f us over %d calls" % (tot * 1e3, len(records)))
