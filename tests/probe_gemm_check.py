"""Correctness of the tcgen05 product on shapes whose tiles are shared between CTAs (stream-K / split-K paths), against float64."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
ag = ge.load_package()
ag.init(0)
rng = np.random.default_rng(0)
for (M, K, N) in [(64, 784, 19072), (64, 784, 65536), (64, 100, 40000), (200, 300, 30000), (30000, 300, 200), (64, 65536, 784), (784, 65536, 64)]:
    a = rng.standard_normal((M, K)).astype(np.float32)
    b = rng.standard_normal((K, N)).astype(np.float32)
    ref = a.astype(np.float64) @ b.astype(np.float64)
    ad, bd = ag.DArray.from_numpy(a), ag.DArray.from_numpy(b)
    ag.set_gemm_engine("tcgen05")
    c1 = ag.matmul(ad, bd).to_numpy()
    c2 = ag.matmul(ad, bd).to_numpy()
    print("%dx%dx%d  max|err|/max|ref| = %.3e  deterministic=%s" % (M, K, N, np.abs(c1 - ref).max() / np.abs(ref).max(), np.array_equal(c1, c2)), flush=True)
