"""Scatter-add and the small movers at the shapes of config 4 (and a many-destination case): values against numpy
(float64 accumulate), bit-exact repeatability, and warm timing against the algorithmic bytes
n*block*s + 8n + 2*touched*block*s.  Run on the GPU box."""
import os, sys
import ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
ag = ge.load_package()
ag.init(0)
D = sys.modules["agb200.darray"]
L = ag._lib
lib, check = L.lib, L.check
rng = np.random.default_rng(0)
PEAK = 6550.1


def timeit(fn, reps=50):
    for _ in range(5):
        fn()
    a, b = ag.Event(), ag.Event()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    return b.elapsed_ms_since(a) / reps * 1e3


for name, nkeys, n, block, dtype in [("embedding gradient, config 4 (V=128, 100 steps x 256 ids, E=128)", 128, 25600, 128, np.float32),
                                     ("same, Float64", 128, 25600, 128, np.float64),
                                     ("scalar destinations (2^20 keys, 2^20 entries, block 1)", 1 << 20, 1 << 20, 1, np.float32),
                                     ("many destinations (50000 keys, 200000 entries, block 16)", 50000, 200000, 16, np.float32),
                                     ("odd block (1000 keys, 30000 entries, block 7)", 1000, 30000, 7, np.float32)]:
    idx = rng.integers(0, nkeys, n).astype(np.int64)
    vals = rng.standard_normal((n, block)).astype(dtype)
    ref = np.zeros((nkeys, block), np.float64)
    np.add.at(ref, idx, vals.astype(np.float64))
    dv = ag.DArray.from_numpy(vals.reshape(-1))
    ib = D.index_to_device(idx)
    outs = []
    for _ in range(2):
        dest = ag.zeros((nkeys * block,), dtype)
        check(lib.ag_scatter_add(dest.dt, C.c_void_p(dest.ptr), dest.size, C.c_void_p(ib.ptr), C.c_void_p(dv.ptr), n, block))
        outs.append(dest.to_numpy().reshape(nkeys, block))
    err = np.max(np.abs(outs[0] - ref)) / max(np.max(np.abs(ref)), 1e-30)
    same = np.array_equal(outs[0], outs[1])
    dest = ag.zeros((nkeys * block,), dtype)
    l0 = ag.launch_count()
    check(lib.ag_scatter_add(dest.dt, C.c_void_p(dest.ptr), dest.size, C.c_void_p(ib.ptr), C.c_void_p(dv.ptr), n, block))
    nl = ag.launch_count() - l0
    us = timeit(lambda: check(lib.ag_scatter_add(dest.dt, C.c_void_p(dest.ptr), dest.size, C.c_void_p(ib.ptr), C.c_void_p(dv.ptr), n, block)))
    s = np.dtype(dtype).itemsize
    touched = len(np.unique(idx))
    by = n * block * s + 8 * n + 2 * touched * block * s
    print("%-72s rel err %.1e repeatable %s | %d launches %7.1f us  %6.0f GB/s (%.3f of HBM peak; %.2f MB)" %
          (name, err, same, nl, us, by / us / 1e3, by / us / 1e3 / PEAK, by / 1e6), flush=True)
    assert err < (1e-5 if dtype == np.float32 else 1e-12) and same

# gather of columns (getindex forward) and slab copies (cat / uncat) at config 4's shapes
Wm = ag.DArray.from_numpy(np.asfortranarray(rng.standard_normal((128, 128)).astype(np.float32)))
ids = D.IArray(rng.integers(0, 128, 256))
us = timeit(lambda: D.gather_cols(Wm, ids))
print("gather_cols (128 x 256 of a 128 x 128 embedding): %.1f us (launch-bound: %.0f KB)" % (us, 128 * 256 * 4 * 2 / 1e3))
big = ag.DArray.from_numpy(np.asfortranarray(rng.standard_normal((1024, 65536)).astype(np.float32)))
ids2 = D.IArray(rng.integers(0, 65536, 65536))
us = timeit(lambda: D.gather_cols(big, ids2), reps=10)
by = 2 * 1024 * 65536 * 4 + 8 * 65536
print("gather_cols (1024 x 65536 from 1024 x 65536): %.1f us %.0f GB/s (%.3f of HBM peak)" % (us, by / us / 1e3, by / us / 1e3 / PEAK))
