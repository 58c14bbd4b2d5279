"""The four layout variants of one small product interleaved (NN, TN, NT, TT, NN, ...) with a few elementwise kernels in
between, as a tape interleaves them: run under ncu --cache-control none and compare with the back-to-back repeats of
probe_small_gemm_layouts.py (instruction-cache effect of switching between 0.5 MB kernel variants)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
ag = ge.load_package()
ag.init(0)
rng = np.random.default_rng(0)
arr = lambda *s: ag.DArray.from_numpy(np.asfortranarray(rng.standard_normal(s).astype(np.float32)))
M, N, K = 512, 256, 128
A, At, B, Bt = arr(M, K), arr(K, M), arr(K, N), arr(N, K)
x = arr(512, 256)
fns = (lambda: ag.matmul(A, B), lambda: ag.matmul(ag.adjoint(At), B), lambda: ag.matmul(A, ag.adjoint(Bt)),
       lambda: ag.matmul(ag.adjoint(At), ag.adjoint(Bt)))
for rep in range(4):
    for fn in fns:
        fn().ptr
        if os.environ.get("FILLER", "1") == "1":
            ag.tanh(ag.add(x, 1.0)).ptr
            ag.sum_(ag.exp(x), dims=0).ptr
            ag.sigm(x).ptr
ag.sync()
