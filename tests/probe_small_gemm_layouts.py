"""One small product (M x N x K = 512 x 256 x 128 and its mirror) in all four operand layouts: run under
`ncu --metrics gpu__time_duration.sum --cache-control none` to compare the kernel variants at equal work."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
ag = ge.load_package()
ag.init(0)
rng = np.random.default_rng(0)
arr = lambda *s: ag.DArray.from_numpy(np.asfortranarray(rng.standard_normal(s).astype(np.float32)))
K = int(os.environ.get("K", 128))
for (M, N) in ((512, 256), (256, 512)):
    A, At, B, Bt = arr(M, K), arr(K, M), arr(K, N), arr(N, K)
    for name, fn in (("NN", lambda: ag.matmul(A, B)), ("TN", lambda: ag.matmul(ag.adjoint(At), B)),
                     ("NT", lambda: ag.matmul(A, ag.adjoint(Bt))), ("TT", lambda: ag.matmul(ag.adjoint(At), ag.adjoint(Bt)))):
        for _ in range(4):
            fn().ptr
        ag.sync()
        print(M, N, K, name, ag.gemm_engine(), flush=True)
