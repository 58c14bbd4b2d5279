"""Per-role wait cycles of k_gemm_tc (needs a library built with -DAGB_TC_TIMING; AGB200_LIB points at it, AGB_TC_TIMING=1)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
ag = ge.load_package()
ag.init(0)
rng = np.random.default_rng(0)
w1 = ag.DArray.from_numpy((0.1 * rng.standard_normal((64, 784))).astype(np.float32))
for B in [int(b) for b in os.environ.get("BS", "56832,65536").split(",")]:
    x = ag.DArray.from_numpy(rng.random((784, B), dtype=np.float32))
    dy = ag.DArray.from_numpy(rng.standard_normal((64, B)).astype(np.float32))
    for i in range(3):
        sys.stderr.write("B=%d w1*x: " % B); sys.stderr.flush()
        ag.matmul(w1, x); ag.sync()
    for i in range(3):
        sys.stderr.write("B=%d dy*x': " % B); sys.stderr.flush()
        ag.matmul(dy, ag.adjoint(x)); ag.sync()
