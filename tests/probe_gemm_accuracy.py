"""Accuracy + timing probe of the GEMM engines on the MNIST shapes (run on the GPU box)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
ag = ge.load_package()
ag.init(0)
rng = np.random.default_rng(0)
B = 8192
for (M, K, N, ta, tb, name) in [(64, 784, B, 0, 0, "fwd"), (64, B, 784, 0, 1, "dW1")]:
    a = rng.standard_normal((K, M) if ta else (M, K)).astype(np.float32)
    b = rng.random((N, K) if tb else (K, N), dtype=np.float32)
    ad, bd = ag.DArray.from_numpy(a), ag.DArray.from_numpy(b)
    A = ag.adjoint(ad) if ta else ad
    Bm = ag.adjoint(bd) if tb else bd
    ref = (a.T if ta else a).astype(np.float64) @ (b.T if tb else b).astype(np.float64)
    mag = (np.abs(a.T if ta else a).astype(np.float64) @ np.abs(b.T if tb else b).astype(np.float64)).max()
    for eng in ("ffma", "tcgen05"):
        ag.set_gemm_engine(eng)
        c = ag.matmul(A, Bm).to_numpy()
        e = np.abs(c - ref).max()
        print("%s %-8s maxerr=%.3e  /max|ref|=%.3e  /mag=%.3e" % (name, eng, e, e / np.abs(ref).max(), e / mag))
    ag.set_gemm_engine("auto")
