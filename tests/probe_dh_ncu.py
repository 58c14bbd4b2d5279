"""Two small LSTM products alone, for `ncu --set full`: dh = Wp'*dy (512x256x128, TN) and dWp = dy*h' (128x512x256, NT)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
ag = ge.load_package()
ag.init(0)
rng = np.random.default_rng(0)
H, B, V = 512, 256, 128
arr = lambda *s: ag.DArray.from_numpy(np.asfortranarray(rng.standard_normal(s).astype(np.float32)))
Wp, h, dy = arr(V, H), arr(H, B), arr(V, B)
which = os.environ.get("WHICH", "dh")
for _ in range(6):
    if which == "dh":
        ag.matmul(ag.adjoint(Wp), dy).ptr
    else:
        ag.matmul(dy, ag.adjoint(h)).ptr
ag.sync()
