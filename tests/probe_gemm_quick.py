"""The two big MNIST products (B = 65536) and a few LSTM-size ones: value check against float64 and warm back-to-back
timing (GPU-bound: each launch takes longer than the host needs to enqueue the next).  AGB200_LIB selects a variant."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
ag = ge.load_package()
ag.init(0)
rng = np.random.default_rng(0)
PEAK = 6550.1


def t(fn, reps=30):
    for _ in range(3):
        fn()
    a, b = ag.Event(), ag.Event()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    return b.elapsed_ms_since(a) / reps * 1e3


def relerr(got, ref):
    return float(np.linalg.norm(got - ref) / np.linalg.norm(ref))


for (M, K, N) in [(64, 784, 4096), (512, 640, 256), (640, 512, 256), (100, 50, 300)]:
    a = rng.standard_normal((M, K)).astype(np.float32)
    b = rng.standard_normal((K, N)).astype(np.float32)
    ad, bd = ag.DArray.from_numpy(np.asfortranarray(a)), ag.DArray.from_numpy(np.asfortranarray(b))
    ref = a.astype(np.float64) @ b.astype(np.float64)
    e1 = relerr(ag.matmul(ad, bd).to_numpy(), ref)
    dy = rng.standard_normal((M, N)).astype(np.float32)
    dyd = ag.DArray.from_numpy(np.asfortranarray(dy))
    e2 = relerr(ag.matmul(dyd, ag.adjoint(bd)).to_numpy(), dy.astype(np.float64) @ b.astype(np.float64).T)
    e3 = relerr(ag.matmul(ag.adjoint(ad), dyd).to_numpy(), a.astype(np.float64).T @ dy.astype(np.float64))
    print("M=%d K=%d N=%d  relerr NN %.2e NT %.2e TN %.2e (%s)" % (M, K, N, e1, e2, e3, ag.gemm_engine()), flush=True)
    assert max(e1, e2, e3) < 1e-5

B = 65536
w1 = ag.DArray.from_numpy((0.1 * rng.standard_normal((64, 784))).astype(np.float32))
x = ag.DArray.from_numpy(rng.random((784, B), dtype=np.float32))
dy = ag.DArray.from_numpy(rng.standard_normal((64, B)).astype(np.float32))
by = 4 * (64 * 784 + 784 * B + 64 * B)
for rep in range(2):
    tf = t(lambda: ag.matmul(w1, x).ptr)
    tb = t(lambda: ag.matmul(dy, ag.adjoint(x)).ptr)
    print("%s  B=%d | w1*x %6.1f us frac %.3f | dy*x' %6.1f us frac %.3f" % (os.environ.get("AGB200_LIB", "default")[-24:], B, tf, by / tf / 1e3 / PEAK,
                                                                          tb, by / tb / 1e3 / PEAK), flush=True)
