"""Timing probe of the three skinny MNIST products (run on the GPU box): y2 = w2*a1, da1 = w2'*dy2, dw2 = dy2*a1'."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
ag = ge.load_package()
ag.init(0)
B = int(os.environ.get("B", 65536))
rng = np.random.default_rng(0)
w2 = ag.DArray.from_numpy((0.1 * rng.standard_normal((10, 64))).astype(np.float32))
a1 = ag.DArray.from_numpy(rng.random((64, B), dtype=np.float32))
dy2 = ag.DArray.from_numpy(rng.standard_normal((10, B)).astype(np.float32))
big = ag.DArray.from_numpy(np.zeros((1 << 26,), np.float32))       # 256 MB: L2 flush between timed launches


def t(fn, reps=10, flush=True):
    for _ in range(3):
        fn()
    tot = 0.0
    for _ in range(reps):
        if flush:
            ag.neg(big)
        a, b = ag.Event(), ag.Event()
        a.record()
        fn()
        b.record()
        ag.sync()
        tot += b.elapsed_ms_since(a)
    return tot / reps * 1e3


for flush in (True, False):
    print("flush=%s  y2=w2*a1 %.1f us   da1=w2'*dy2 %.1f us   dw2=dy2*a1' %.1f us   (%s)" % (
        flush, t(lambda: ag.matmul(w2, a1), flush=flush), t(lambda: ag.matmul(ag.adjoint(w2), dy2), flush=flush),
        t(lambda: ag.matmul(dy2, ag.adjoint(a1)), flush=flush), ag.gemm_engine()))
