"""Where does the Float32 error of the BASELINE-size LSTM come from?  (GPU box.)  Runs the config-4 tape with the
tcgen05 engine and with the FFMA engine against the float64 oracle, and the bare gate product against float64."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import __graft_entry__ as ge
ag = ge.load_package()
ag.init(0)
from oracle import agref
import test_gpu_baseline_sizes as tb

rng = np.random.default_rng(0)
for (M, K, N) in [(512, 640, 256), (2048, 640, 256), (128, 512, 256), (64, 784, 8192)]:
    a = (rng.standard_normal((M, K)) / np.sqrt(K)).astype(np.float32)
    b = rng.standard_normal((K, N)).astype(np.float32)
    ref = a.astype(np.float64) @ b.astype(np.float64)
    ad, bd = ag.DArray.from_numpy(a), ag.DArray.from_numpy(b)
    for eng in ("ffma", "tcgen05"):
        ag.set_gemm_engine(eng)
        try:
            c = ag.matmul(ad, bd).to_numpy()
        except Exception as e:
            print(M, K, N, eng, "unsupported", e)
            continue
        d = c - ref
        print("gemm %dx%dx%d %-8s max|err|/max|ref| = %.3e   shrink (slope of err on ref) = %.3e   rms = %.3e" % (
            M, K, N, eng, np.abs(d).max() / np.abs(ref).max(), (d * ref).sum() / (ref * ref).sum(),
            np.sqrt((d ** 2).mean()) / np.abs(ref).max()))
ag.set_gemm_engine("auto")

T = int(os.environ.get("T", 100))
W, w, ids, tg = tb._lstm_case(T=T)
keys = list(w)
H, B = 512, 256
po = {k: agref.Param(np.asfortranarray(w[k].astype(np.float64))) for k in keys}
to = agref.diff(lambda: W.lstm_loss(agref, po, ids, [a.astype(np.float64) for a in tg], np.zeros((H, B)), np.zeros((H, B))))
g64 = {k: np.asarray(agref.full(agref.grad(to, po[k]))) for k in keys}
tgd = [ag.DArray.from_numpy(a) for a in tg]
h0 = ag.zeros((H, B), np.float32)
for eng in ("auto", "ffma"):
    ag.set_gemm_engine(eng)
    params = {k: ag.Param(ag.DArray.from_numpy(w[k])) for k in keys}
    t = ag.diff(lambda: W.lstm_loss(ag, params, ids, tgd, h0, h0))
    print("LSTM T=%d engine=%s loss relerr %.3e" % (T, eng, abs(float(ag.value(t)) - float(agref.value(to))) / abs(float(agref.value(to)))))
    print("   ", {k: "%.2e" % tb.relerr(ag.full(ag.grad(t, params[k])).to_numpy(), g64[k]) for k in keys})
