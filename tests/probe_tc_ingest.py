"""Per-SM fill rate (experiment build with -DAGB_TC_TRACE, selected by AGB200_LIB): every SM streams 16 KB tiles of a
large array by TMA and, per tile, 8 KB of a small L2-resident array not at all / by TMA / by ld.global + st.shared."""
import os, sys, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
ag = ge.load_package()
ag.init(0)
lib = ag._lib.lib
lib.ag_debug_ingest.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_longlong, C.c_void_p, C.POINTER(C.c_double)]
tiles = 96
B = ag.zeros((32 * 512,), np.float32)
for pattern, pname in ((0, "contiguous 16 KB tiles"), (1, "w1*x layout: 128 rows x 128 B at a 3136-byte pitch")):
    rows = 148 * tiles * 128 if pattern == 0 else 148 * (tiles // 24) * 128
    A = ag.zeros(((32 if pattern == 0 else 784) * rows,), np.float32)
    print(pname)
    for mode, name in ((0, "A only (16 KB / tile, DRAM)"), (1, "A + B by TMA (24 KB / tile)"), (2, "A by TMA + B by ld.global (24 KB / tile)")):
        ns = C.c_double()
        st = lib.ag_debug_ingest(mode, tiles, pattern, C.c_void_p(A.ptr), rows, C.c_void_p(B.ptr), C.byref(ns))
        kb = 16 if mode == 0 else 24
        print("  %-42s status %d  %7.1f ns per tile  = %5.1f GB/s per SM of A (%5.1f in total), %.2f TB/s of DRAM over 148 SMs"
              % (name, st, ns.value, 16384 / ns.value, kb * 1024 / ns.value, 148 * 16384 / ns.value / 1e3))
    del A
