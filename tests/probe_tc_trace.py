"""Per-k-block timeline of block 0 of one small product (experiment build with -DAGB_TC_TRACE, selected by AGB200_LIB):
where the time between k-blocks goes when the TMEM partial is promoted every k-block."""
import os, sys, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
ag = ge.load_package()
ag.init(0)
rng = np.random.default_rng(0)
arr = lambda *s: ag.DArray.from_numpy(np.asfortranarray(rng.standard_normal(s).astype(np.float32)))
lib = ag._lib.lib
lib.ag_gemm_trace.argtypes = [C.POINTER(C.c_uint64)]
for name, a, b in (("gates-like NN 2048x256x640 (no split: 64 tiles, 20 k-blocks)", arr(2048, 640), arr(640, 256)),
                   ("dx-like TN 2560x256x512 (80 tiles, 16 k-blocks)", ag.adjoint(arr(512, 2560)), arr(512, 256))):
    for _ in range(3):
        ag.matmul(a, b).ptr
    buf = (C.c_uint64 * (6 * 256))()
    assert lib.ag_gemm_trace(buf) == 0
    t = np.array(buf[:], dtype=np.int64).reshape(6, 256)
    n = 20 if "gates" in name else 16
    t0 = t[0][0]
    print(name, ag.gemm_engine())
    print("  kb: loads issued | data seen | split done | mma issued | partial seen | acc released   (ns from first issue)")
    for k in range(n):
        print("  %2d: %6d %6d %6d %6d %6d %6d" % ((k,) + tuple(int(t[r][k] - t0) for r in range(6))))
