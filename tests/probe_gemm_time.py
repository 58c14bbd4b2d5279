"""[historic probe: the AGB_TC_MMAS / AGB_TC_NOSPLIT experiment knobs it mentions were removed from gemm_tc.cu after the
TS-mode rewrite; it still runs and times the current kernel]  Timing probe of the two big MNIST products (run on the GPU box)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
ag = ge.load_package()
ag.init(0)
B = 65536
rng = np.random.default_rng(0)
w1 = ag.DArray.from_numpy((0.1 * rng.standard_normal((64, 784))).astype(np.float32))
x = ag.DArray.from_numpy(rng.random((784, B), dtype=np.float32))
dy = ag.DArray.from_numpy(rng.standard_normal((64, B)).astype(np.float32))
def t(fn, reps=20):
    for _ in range(3): fn()
    a, b = ag.Event(), ag.Event()
    a.record()
    for _ in range(reps): fn()
    b.record()
    return b.elapsed_ms_since(a) / reps * 1e3
print("env MMAS=%s NOSPLIT=%s  fwd %.1f us   dW1 %.1f us" % (os.environ.get("AGB_TC_MMAS"), os.environ.get("AGB_TC_NOSPLIT"),
      t(lambda: ag.matmul(w1, x)), t(lambda: ag.matmul(dy, ag.adjoint(x)))))
