import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
ag = ge.load_package()
ag.init(0)
rng = np.random.default_rng(0)
M, K = 64, int(os.environ.get("K", 784))
ag.set_gemm_engine("tcgen05")
for N in [int(v) for v in os.environ.get("NS", "56832,65536").split(",")]:
    a = rng.standard_normal((M, K)).astype(np.float32)
    b = rng.standard_normal((K, N)).astype(np.float32)
    ref = a.astype(np.float64) @ b.astype(np.float64)
    ad, bd = ag.DArray.from_numpy(a), ag.DArray.from_numpy(b)
    G, KB = 148, (K + 31) // 32
    tiles = (N + 127) // 128
    U = tiles * KB
    for rep in range(2):
        c = ag.matmul(ad, bd).to_numpy()
        err = np.abs(c - ref) / np.abs(ref).max()
        badcols = np.nonzero((err > 1e-5).any(axis=0))[0]
        print("N=%d rep %d: bad columns %d; tiles touched %d" % (N, rep, badcols.size, np.unique(badcols // 128).size))
        shown = 0
        for t in np.unique(badcols // 128)[:12]:
            cols = badcols[badcols // 128 == t] % 128
            c_lo = ((t * KB + 1) * G - 1) // U
            c_hi = ((t * KB + KB) * G - 1) // U
            u0 = [cc * U // G for cc in (c_lo, c_lo + 1)]
            # which k range is missing? compare with partial sums over k-blocks
            col = t * 128 + cols[0]
            got = c[:, col].astype(np.float64)
            pk = np.stack([a[:, k0:k0 + 32].astype(np.float64) @ b[k0:k0 + 32, col].astype(np.float64) for k0 in range(0, K, 32)])
            diff = ref[:, col] - got
            # least squares: which k-blocks explain the difference (coefficients ~1 = missing, ~-1 = doubled)
            coef, *_ = np.linalg.lstsq(pk.T, diff, rcond=None)
            print("  tile %d (ctas %d..%d, cta unit range starts %s, tile units %d..%d) rows(lanes) %s  kblock coef %s" % (
                t, c_lo, c_hi, u0, t * KB, t * KB + KB - 1, cols.tolist(), np.round(coef, 2).tolist()))
