"""Timing probe of the two big MNIST products as a function of the batch (run on the GPU box): is the time per
128-column tile the same when the tiles divide evenly over the SMs (B = 148*128*k) as at B = 65536 (3.46 waves)?"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
ag = ge.load_package()
ag.init(0)
rng = np.random.default_rng(0)
w1 = ag.DArray.from_numpy((0.1 * rng.standard_normal((64, 784))).astype(np.float32))
big = ag.zeros((1 << 26,), np.float32)       # 256 MB: L2 flush between timed launches
PEAK = 6550.1


def t(fn, reps=10):
    for _ in range(3):
        fn()
    tot = 0.0
    for _ in range(reps):
        ag.neg(big).ptr
        a, b = ag.Event(), ag.Event()
        a.record()
        fn()
        b.record()
        ag.sync()
        tot += b.elapsed_ms_since(a)
    return tot / reps * 1e3


Bs = [int(b) for b in os.environ.get("BS", "18944,37888,56832,65536,75776,16384,32768,131072").split(",")]
for B in Bs:
    x = ag.DArray.from_numpy(rng.random((784, B), dtype=np.float32))
    dy = ag.DArray.from_numpy(rng.standard_normal((64, B)).astype(np.float32))
    by = 4 * (64 * 784 + 784 * B + 64 * B)
    tf = t(lambda: ag.matmul(w1, x))
    ef = ag.gemm_engine()
    tb = t(lambda: ag.matmul(dy, ag.adjoint(x)))
    eb = ag.gemm_engine()
    print("B=%7d tiles=%6.2f waves=%5.2f | w1*x %6.1f us %5.0f GB/s frac %.3f (%s) | dy*x' %6.1f us %5.0f GB/s frac %.3f (%s)"
          % (B, B / 128, B / 128 / 148, tf, by / tf / 1e3, by / tf / 1e3 / PEAK, ef, tb, by / tb / 1e3, by / tb / 1e3 / PEAK, eb),
          flush=True)
    del x, dy
