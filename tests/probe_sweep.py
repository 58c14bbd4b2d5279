"""Config 3 (broadcast / unbroadcast sweep) at one size: per-launch algorithmic traffic of the fused step, to be
set against an `ncu --metrics gpu__time_duration.sum` launch list of the same run (profiles/).
ROWS x COLS from the environment (default 1024 x 262144 = 2^28 elements, 1 GiB per array)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
ag = ge.load_package()
ag.init(0)
L = sys.modules["agb200.lazy"]
W = ag.workloads
rows, cols = int(os.environ.get("ROWS", 1024)), int(os.environ.get("COLS", 262144))
x, brow, bcol = W.sweep_init(np.random.default_rng(0), rows, cols)
ps = [ag.Param(ag.DArray.from_numpy(a)) for a in (x, brow, bcol)]
del x
ag.set_gc_function(ag.release_node)


def step():
    t = ag.diff(lambda: W.sweep_loss(ag, ps[0], ps[1], ps[2]))
    return [ag.grad(t, p) for p in ps]


step(); ag.sync()
L.trace = []
l0 = ag.launch_count()
e0, e1 = ag.Event(), ag.Event()
e0.record()
g = step()
e1.record()
ag.sync()
ms = e1.elapsed_ms_since(e0)
n = rows * cols
print("sweep %dx%d: %.3f ms/step, %d launches, algorithmic (unfused tape) %.0f GB/s" % (rows, cols, ms, ag.launch_count() - l0, 4.0 * n * 25 / ms / 1e6))
tot = 0
for ne, nfull, nout, code in L.trace:
    b = 4 * ne * (nfull + nout)
    tot += b
    print("  fused launch: %10d elements, %d full inputs, %d outputs -> %.1f MB   program %s" % (ne, nfull, nout, b / 1e6, [hex(c) for c in code]))
print("fused launches move %.1f MB in total" % (tot / 1e6))
