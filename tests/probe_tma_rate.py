"""[historic probe: the AGB_TC_MMAS / AGB_TC_NOSPLIT experiment knobs it mentions were removed from gemm_tc.cu after the
TS-mode rewrite; it still runs and times the current kernel]  TMA feed-rate probe: the forward product with different K (row pitch of x) -- run with
AGB_TC_MMAS=1 AGB_TC_NOSPLIT=1 so the kernel is essentially a TMA streaming loop."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
ag = ge.load_package()
ag.init(0)
rng = np.random.default_rng(0)
def t(fn, reps=20):
    for _ in range(3): fn()
    a, b = ag.Event(), ag.Event()
    a.record()
    for _ in range(reps): fn()
    b.record()
    return b.elapsed_ms_since(a) / reps * 1e3
for K, B in [(768, 65536), (784, 65536), (800, 65536), (1024, 49152), (784, 18944), (784, 37888), (784, 151552)]:
    w1 = ag.DArray.from_numpy(np.zeros((64, K), np.float32))
    x = ag.DArray.empty((K, B), np.float32)
    us = t(lambda: ag.matmul(w1, x))
    print("fwd K=%4d B=%6d  tiles=%4d  %.1f us  %.2f TB/s" % (K, B, B // 128, us, K * B * 4 / us / 1e6))
    xt = ag.DArray.empty((64, B), np.float32)
    us = t(lambda: ag.matmul(xt, ag.adjoint(x)))
    print("dW  K=%4d B=%6d               %.1f us  %.2f TB/s" % (K, B, us, K * B * 4 / us / 1e6))
