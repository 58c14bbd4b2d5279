"""Warm, event-timed launches of the products of one charlm LSTM timestep (config 4: H=512, E=128, V=128, B=256),
alone and as ag_gemm_group launches.  Run on the GPU box."""
import os, sys
import ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
ag = ge.load_package()
ag.init(0)
L = ag._lib
lib, check = L.lib, L.check
rng = np.random.default_rng(0)
H, EH, B, V = 512, 640, 256, 128


def arr(*s):
    return ag.DArray.from_numpy(np.asfortranarray(rng.standard_normal(s).astype(np.float32)))


def timeit(fn, reps=50):
    for _ in range(5):
        fn()
    a, b = ag.Event(), ag.Event()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    return b.elapsed_ms_since(a) / reps * 1e3


def group(probs):
    descs = (L.GemmDesc * len(probs))()
    keep = []
    for d, (ta, tb, M, N, K, A, lda, Bm, ldb) in zip(descs, probs):
        c = ag.DArray.empty((M, N), np.float32)
        keep.append(c)
        d.transA, d.transB, d.M, d.N, d.K, d.A, d.lda, d.B, d.ldb = ta, tb, M, N, K, A.ptr, lda, Bm.ptr, ldb
        d.bias, d.act, d.C, d.C2, d.ldc = None, 0, c.ptr, None, M
    return lambda: check(lib.ag_gemm_group(0, len(probs), descs)), keep


W = [arr(H, EH) for _ in range(4)]
x = arr(EH, B)
dz = [arr(H, B) for _ in range(4)]
Wp = arr(V, H)
h = arr(H, B)
dyp = arr(V, B)
cases = {
    "gate  W*x        (512x256x640, NN)": (0, 0, H, B, EH, W[0], H, x, EH),
    "dx    W'*dz      (640x256x512, TN)": (1, 0, EH, B, H, W[0], H, dz[0], H),
    "dW    dz*x'      (512x640x256, NT)": (0, 1, H, EH, B, dz[0], H, x, EH),
    "pred  Wp*h       (128x256x512, NN)": (0, 0, V, B, H, Wp, V, h, H),
    "dh    Wp'*dy     (512x256x128, TN)": (1, 0, H, B, V, Wp, V, dyp, V),
    "dWp   dy*h'      (128x512x256, NT)": (0, 1, V, H, B, dyp, V, h, H),
}
ENG = os.environ.get("ENGINE", "auto")
ag.set_gemm_engine(ENG)
print("engine:", ENG)
for name, p in cases.items():
    fn, keep = group([p])
    l0 = ag.launch_count()
    fn()
    nl = ag.launch_count() - l0
    us = timeit(fn)
    eng = ag.gemm_engine()
    print("%-40s alone %6.1f us  (%d launches, %s)" % (name, us, nl, eng))
for name, ps in {
    "4 gates": [(0, 0, H, B, EH, W[g], H, x, EH) for g in range(4)],
    "4 dx": [(1, 0, EH, B, H, W[g], H, dz[g], H) for g in range(4)],
    "4 dW": [(0, 1, H, EH, B, dz[g], H, x, EH) for g in range(4)],
    "8 dW": [(0, 1, H, EH, B, dz[g % 4], H, x, EH) for g in range(8)],
}.items():
    fn, keep = group(ps)
    l0 = ag.launch_count()
    fn()
    nl = ag.launch_count() - l0
    print("%-40s group %6.1f us  (%d launches)" % (name, timeit(fn), nl))
