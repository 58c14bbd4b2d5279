"""Timing of the product with its fused epilogue variants at the MNIST layer-1 shape (GPU box)."""
import os, sys, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
ag = ge.load_package()
ag.init(0)
lib, check = ag._lib.lib, ag._lib.check
rng = np.random.default_rng(0)
B = int(os.environ.get("B", 65536))
w1 = ag.DArray.from_numpy((0.1 * rng.standard_normal((64, 784))).astype(np.float32))
b1 = ag.DArray.from_numpy(rng.standard_normal((64,)).astype(np.float32))
x = ag.DArray.from_numpy(rng.random((784, B), dtype=np.float32))
c = ag.DArray.empty((64, B), np.float32)
c2 = ag.DArray.empty((64, B), np.float32)
vp = lambda a: C.c_void_p(a.ptr if a is not None else None)


def run(bias, act, two):
    check(lib.ag_gemm_epilogue(0, 0, 0, 64, B, 784, vp(w1), 64, vp(x), 784, vp(bias), act, vp(c), vp(c2 if two else None), 64))


def t(fn, reps=30):
    for _ in range(3):
        fn()
    a, b = ag.Event(), ag.Event()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    return b.elapsed_ms_since(a) / reps * 1e3


print("plain ag_gemm            %.1f us" % t(lambda: check(lib.ag_gemm(0, 0, 0, 64, B, 784, 1.0, vp(w1), 64, vp(x), 784, 0.0, vp(c), 64))))
print("epilogue none            %.1f us" % t(lambda: run(None, 0, False)))
print("epilogue bias            %.1f us" % t(lambda: run(b1, 0, False)))
print("epilogue bias+relu (1)   %.1f us" % t(lambda: run(b1, 28, False)))
print("epilogue bias+relu (2)   %.1f us" % t(lambda: run(b1, 28, True)))
print("epilogue bias+tanh (1)   %.1f us" % t(lambda: run(b1, 6, False)))
bp = ag.DArray.from_numpy(np.full((64,), 1000.0, np.float32))
bn = ag.DArray.from_numpy(np.full((64,), -1000.0, np.float32))
print("relu, all outputs > 0    %.1f us" % t(lambda: run(bp, 28, False)))
print("relu, all outputs = 0    %.1f us" % t(lambda: run(bn, 28, False)))
print("bias only, bias -1000    %.1f us" % t(lambda: run(bn, 0, False)))
z = ag.zeros((64, B), np.float32)
print("fill zeros 16.8MB x1     %.1f us" % t(lambda: check(lib.ag_fill(vp(c), 64 * B, 0, 0.0))))
print("fill ones 16.8MB x1      %.1f us" % t(lambda: check(lib.ag_fill(vp(c), 64 * B, 0, 1.0))))
