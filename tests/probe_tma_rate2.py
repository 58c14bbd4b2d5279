import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
ag = ge.load_package()
ag.init(0)
def t(fn, reps=20):
    for _ in range(3): fn()
    a, b = ag.Event(), ag.Event()
    a.record()
    for _ in range(reps): fn()
    b.record()
    return b.elapsed_ms_since(a) / reps * 1e3
for K in (32, 64, 128, 256, 512):
    B = (200 * 1024 * 1024 // (K * 4)) // 128 * 128
    w1 = ag.DArray.from_numpy(np.zeros((64, K), np.float32))
    x = ag.DArray.empty((K, B), np.float32)
    ag.set_gemm_engine("tcgen05")
    us = t(lambda: ag.matmul(w1, x))
    print("fwd K=%4d B=%8d  %.1f us  %.2f TB/s (x only)  engine %s" % (K, B, us, K * B * 4 / us / 1e6, ag.gemm_engine()))
