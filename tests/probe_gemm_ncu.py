"""The two big MNIST products alone (B = 65536), three launches each: the target of `ncu --set full -k regex:k_gemm_tc`."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
ag = ge.load_package()
ag.init(0)
rng = np.random.default_rng(0)
B = 65536
w1 = ag.DArray.from_numpy((0.1 * rng.standard_normal((64, 784))).astype(np.float32))
b1 = ag.DArray.from_numpy((0.1 * rng.standard_normal((64, 1))).astype(np.float32))
x = ag.DArray.from_numpy(rng.random((784, B), dtype=np.float32))
dy = ag.DArray.from_numpy(rng.standard_normal((64, B)).astype(np.float32))
for _ in range(3):
    ag.relu(ag.add(ag.matmul(w1, x), b1)).ptr          # the form the step runs: bias + relu in the epilogue
    ag.matmul(dy, ag.adjoint(x)).ptr
ag.sync()
