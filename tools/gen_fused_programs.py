#!/usr/bin/env python
"""Generate autograd.jl_b200/csrc/fuse_static.inc: the fused elementwise programs that the BASELINE.json
configs produce, to be instantiated ahead of time as specialised kernels (csrc/fuse.cu, k_fused_static).

The host runtime (autograd.jl_b200/lazy.py) compiles each pending elementwise chain of a tape to bytecode.
Which programs occur depends only on the STRUCTURE of the model code (not on sizes or data), so the list is
collected here on the CPU: the five config workloads (autograd.jl_b200/workloads.py) are run at tiny sizes on
the numpy stand-in for the C ABI (tests/fake_abi.py, a build-time use of the test double: nothing is computed
for the product) and every distinct program handed to ag_ewise_fused is recorded.  Programs outside the list
still run -- on the general interpreter kernel.

Usage: python tools/gen_fused_programs.py   (then rebuild: python __graft_entry__.py)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import __graft_entry__ as ge  # noqa: E402
import fake_abi  # noqa: E402

OUT = os.path.join(ROOT, "autograd.jl_b200", "csrc", "fuse_static.inc")
OUT_COL = os.path.join(ROOT, "autograd.jl_b200", "csrc", "fuse_static_col.inc")
MAX_STATIC = 160
MAX_COL_STATIC = 48


def main(out=OUT):
    ag = ge.load_package()
    progs, colprogs = {}, {}

    class Recorder(fake_abi.FakeLib):
        def ag_ewise_fused(self, dtype, code, ncode, consts, nconsts, nin, *rest):
            key = tuple(int(code[k]) & 0xffffffff for k in range(ncode))
            ent = progs.setdefault(key, {"nin": nin, "count": 0, "where": set()})
            ent["count"] += 1
            ent["where"].add(CUR[0])
            return super().ag_ewise_fused(dtype, code, ncode, consts, nconsts, nin, *rest)

        def ag_ewise_fused_reduce(self, dtype, code, ncode, consts, nconsts, nin, *rest):
            st = super().ag_ewise_fused_reduce(dtype, code, ncode, consts, nconsts, nin, *rest)
            if st == 0:                      # chains whose sum is fused into their own launch (instruction 11)
                key = tuple(int(code[k]) & 0xffffffff for k in range(ncode))
                ent = progs.setdefault(key, {"nin": nin, "count": 0, "where": set()})
                ent["count"] += 1
                ent["where"].add(CUR[0])
                # the same chain without the fused sum: what runs when the layout of the sum is not fusable
                # (e.g. row sums of a matrix with more than 128 rows), which depends on sizes, not on structure
                plain = list(key[:-1])
                if len(plain) >= 2 and (plain[-1] & 0xff) == 4 and (plain[-2] & 0xff) == 10:
                    plain.pop()
                if plain and (plain[-1] & 0xff) == 10:
                    ent = progs.setdefault(tuple(plain), {"nin": nin, "count": 0, "where": set()})
                    ent["count"] += 1
                    ent["where"].add(CUR[0])
            return st

        def ag_ewise_colprog(self, dtype, code, ncode, consts, nconsts, nin, *rest):
            key = tuple(int(code[k]) & 0xffffffff for k in range(ncode))
            ent = colprogs.setdefault(key, {"nin": nin, "count": 0, "where": set(), "f32": False})
            ent["count"] += 1
            ent["where"].add(CUR[0])
            ent["f32"] = ent["f32"] or dtype == 0
            return super().ag_ewise_colprog(dtype, code, ncode, consts, nconsts, nin, *rest)

        # Which products take their `.+ b` / activation into the epilogue depends on sizes (the tcgen05 engine takes
        # them); both outcomes are recorded: EPI[0] = None -> the stand-in's size rule, 0 / 1 -> never / always.
        def ag_gemm_epilogue_query(self, dtype, ta, tb, M, N, K, A, lda, B, ldb, out):
            if EPI[0] is None:
                return super().ag_gemm_epilogue_query(dtype, ta, tb, M, N, K, A, lda, B, ldb, out)
            fake_abi._out(out).value = EPI[0]
            return 0

        # stream capture on the stand-in: the step simply runs (only the sequence of programs matters here)
        def ag_graph_begin(self): return 0
        def ag_graph_end(self, ref): ref._obj.value = 1; return 0
        def ag_graph_launch(self, h): return 0
        def ag_graph_destroy(self, h): return 0

    CUR = ["?"]
    EPI = [None]
    fake, restore = fake_abi.install(ag, Recorder())
    L = sys.modules["agb200.lazy"]
    prev = L.set_fusion(True)
    W = ag.workloads
    dev = ag.DArray.from_numpy
    rng = np.random.default_rng(0)

    def sgd(loss, params, steps=2, lr=-0.05, clip=None):
        for _ in range(steps):                  # the same step without reading the loss (tests/probe_*.py, user loops)
            t = ag.diff(lambda: loss(params))
            gs = [ag.grad(t, p) for p in params]
            del t
            for p, g in zip(params, gs):
                ag.axpy(lr, g, p)
            ag.sync()
            t = ag.diff(lambda: loss(params))
            gs = [ag.grad(t, p) for p in params]
            ag.sync()
            del t, gs
        for _ in range(steps):
            t = ag.diff(lambda: loss(params))
            float(ag.value(t))
            gs = [ag.grad(t, p) for p in params]
            if clip is not None:                                   # examples/charlm.jl:116-118,148-149
                gn = ag.sqrt(sum((ag.sumabs2(ag.full(g)) for g in gs[1:]), ag.sumabs2(ag.full(gs[0]))))
                float(gn)
            for p, g in zip(params, gs):
                ag.axpy(lr, g, p)
        ag.sync()

    try:
        # (gc function, epilogue rule, column programs): column sums stay inside a chain only for small arrays
        # (lazy.BIG), so the chains of both forms occur
        for gc, epi, colf in ((None, None, True), (ag.release_node, None, True), (ag.release_node, 1, True), (None, 1, True),
                              (None, None, False), (ag.release_node, None, False), (ag.release_node, 1, False)):
            ag.set_gc_function(gc if gc is not None else (lambda n, t: None))
            EPI[0] = epi
            L.flush()
            L.col_fusion[0] = colf
            for dtype in ((np.float64, np.float32) if epi is None else (np.float32,)):
                CUR[0] = "housing"
                w, x, y = W.housing_init(rng, 24, 5, dtype)
                xd, yd = dev(x), dev(y)
                sgd(lambda p: W.housing_loss(ag, p, xd, yd), [ag.Param(dev(a)) for a in w])
                w2 = [w[0], 0.0]                                   # scalar bias Param (README.md:72)
                CUR[0] = "mnist"
                w, x, y = W.mnist_init(rng, 32, dtype)
                xd, yd = dev(x), dev(y)
                sgd(lambda p: W.mnist_loss(ag, p, xd, yd), [ag.Param(dev(a)) for a in w])
                sgd(lambda p: W.mnist_loss(ag, p, xd, yd, global_batch=64), [ag.Param(dev(a)) for a in w])
                dp, state = ag.DataParallel(0, 1), {}                # bench.py's step (state keeps loss and grads)
                pb = [ag.Param(dev(a)) for a in w]
                for gb in (None, 64):
                    for _ in range(3):
                        W.mnist_train_step(ag, dp, pb, xd, yd, 0.1, state, global_batch=gb)
                    with ag.StepGraph():
                        W.mnist_train_step(ag, dp, pb, xd, yd, 0.1, state, global_batch=gb)
                    ag.sync()
                CUR[0] = "sweep"
                x, brow, bcol = W.sweep_init(rng, 16, 24, dtype)
                sgd(lambda p: W.sweep_loss(ag, p[0], p[1], p[2]), [ag.Param(dev(a)) for a in (x, brow, bcol)])
                xd = dev(x)
                sgd(lambda p: W.sweep_loss(ag, xd, p[0], p[1]), [ag.Param(dev(a)) for a in (brow, bcol)])
                CUR[0] = "charlm"
                H, E, V, B = 16, 8, 12, 8
                wd = W.lstm_init(rng, H, E, V, dtype)
                keys = list(wd)
                for dev_ids, T in ((False, 3), (True, 3), (True, 10)):   # long sequences: gradient-accumulation chains
                    ids = [rng.integers(0, V, B) for _ in range(T)]
                    if dev_ids:
                        ids = [ag.IArray(i) for i in ids]
                    tg = []
                    for _ in range(T):
                        oh = np.zeros((V, B), dtype)
                        oh[rng.integers(0, V, B), np.arange(B)] = 1
                        tg.append(dev(np.asfortranarray(oh)))
                    h0, c0 = dev(np.zeros((H, B), dtype)), dev(np.zeros((H, B), dtype))
                    sgd(lambda p: W.lstm_loss(ag, dict(zip(keys, p)), ids, tg, h0, c0),
                        [ag.Param(dev(wd[k])) for k in keys], clip=5.0)
                CUR[0] = "selftest"                                # the generic chain of tests/test_gpu_fusion.py
                D = sys.modules["agb200.darray"]
                xs_, bs_ = dev(rng.standard_normal((8, 12)).astype(dtype)), dev(rng.standard_normal((8, 1)).astype(dtype))
                for f in (ag.sum_, ag.sumabs2):
                    for dims in (None, 0, 1):
                        y_ = D.unary_fwd("tanh", D.binary_fwd("add", D.binary_fwd("mul", xs_, 0.5), bs_))
                        (f(y_) if dims is None else f(y_, dims=dims)).to_numpy()
                        del y_
                CUR[0] = "zoo"        # close relatives of the configs (test/neuralnet.jl-style nets), same step loops
                xz, yz = dev(rng.standard_normal((12, 16)).astype(dtype)), dev(rng.standard_normal((4, 16)).astype(dtype))
                oh = np.zeros((4, 16), dtype)
                oh[rng.integers(0, 4, 16), np.arange(16)] = 1
                ohz = dev(np.asfortranarray(oh))
                wz = lambda: [ag.Param(dev((0.1 * rng.standard_normal(sh)).astype(dtype)))
                              for sh in ((8, 12), (8, 1), (4, 8), (4, 1))]
                for act in (ag.tanh, ag.sigm, ag.relu):
                    def net(p, act=act):
                        return ag.add(ag.matmul(p[2], act(ag.add(ag.matmul(p[0], xz), p[1]))), p[3])
                    sgd(lambda p: ag.div(ag.sumabs2(ag.sub(net(p), yz)), 16), wz())                 # squared error
                    sgd(lambda p: ag.div(ag.neg(ag.sum_(ag.mul(ohz, ag.sub(net(p), ag.log(ag.sum_(ag.exp(net(p)), dims=0)))))), 16), wz())
                wl = [ag.Param(dev((0.1 * rng.standard_normal((1, 12))).astype(dtype))), ag.Param(dev(np.zeros((1, 1), dtype)))]
                tz = dev((rng.random((1, 16)) > 0.5).astype(dtype))
                def logistic(p):                                   # -mean(t*log(s) + (1-t)*log(1-s))
                    s_ = ag.sigm(ag.add(ag.matmul(p[0], xz), p[1]))
                    return ag.div(ag.neg(ag.sum_(ag.add(ag.mul(tz, ag.log(s_)), ag.mul(ag.sub(1, tz), ag.log(ag.sub(1, s_)))))), 16)
                sgd(logistic, wl)
                CUR[0] = "rnn"
                H, V, B, T = 8, 6, 8, 3
                w = W.rnn_init(rng, H, V, dtype)
                xs = []
                for _ in range(T):
                    oh = np.zeros((V, B), dtype)
                    oh[rng.integers(0, V, B), np.arange(B)] = 1
                    xs.append(dev(np.asfortranarray(oh)))
                y = np.zeros((V, B), dtype)
                y[rng.integers(0, V, B), np.arange(B)] = 1
                yd, hd = dev(np.asfortranarray(y)), dev(np.zeros((H, B), dtype))
                sgd(lambda p: W.rnn_loss(ag, p, xs, yd, hd), [ag.Param(dev(a)) for a in w])
                wdv = [dev(a) for a in w]

                def loss(whh):
                    return W.rnn_loss(ag, [wdv[0], whh, wdv[2], wdv[3], wdv[4]], xs, yd, hd)
                inner = ag.grad(loss)
                ag.grad(lambda whh: ag.sumabs2(inner(whh)))(wdv[1]).to_numpy()
    finally:
        ag.set_gc_function(lambda n, t: None)
        L.flush()
        L.col_fusion[0] = False
        L.set_fusion(prev)
        restore()

    items = sorted(progs.items(), key=lambda kv: (-kv[1]["count"], kv[0]))[:MAX_STATIC]
    lines = ["// fuse_static.inc -- GENERATED by tools/gen_fused_programs.py; do not edit.",
             "// The fused elementwise programs of the BASELINE.json configs (housing, mnist, sweep, charlm, rnn incl.",
             "// grad-of-grad), instantiated ahead of time by csrc/fuse.cu.  Instruction word: op | a << 8 | b << 16.",
             "#define FZ_NSTATIC %d" % len(items),
             "#define FZ_STATIC_LIST(X) " + " ".join("X(%d)" % i for i in range(len(items))),
             "constexpr StaticProg kProgs[%d] = {" % max(len(items), 1)]
    for i, (code, ent) in enumerate(items):
        lines.append("  /* %2d: %s */ {%d, %d, {%s}}," % (i, ",".join(sorted(ent["where"])), len(code), ent["nin"],
                                                         ", ".join("0x%xu" % c for c in code)))
    if not items:
        lines.append("  {0, 0, {0}},")
    lines.append("};")
    text = "\n".join(lines) + "\n"
    citems = sorted(((k, v) for k, v in colprogs.items() if v["f32"]), key=lambda kv: (-kv[1]["count"], kv[0]))[:MAX_COL_STATIC]
    lines = ["// fuse_static_col.inc -- GENERATED by tools/gen_fused_programs.py; do not edit.",
             "// The column programs (ag_ewise_colprog) of the BASELINE.json configs, instantiated ahead of time for Float32.",
             "#define FZ_NCOLSTATIC %d" % len(citems),
             "#define FZ_COLSTATIC_LIST(X) " + " ".join("X(%d)" % i for i in range(len(citems))),
             "constexpr StaticColProg kColProgs[%d] = {" % max(len(citems), 1)]
    for i, (code, ent) in enumerate(citems):
        lines.append("  /* %2d: %s */ {%d, {%s}}," % (i, ",".join(sorted(ent["where"])), len(code),
                                                    ", ".join("0x%xu" % c for c in code)))
    if not citems:
        lines.append("  {0, {0}},")
    lines.append("};")
    ctext = "\n".join(lines) + "\n"
    if out is not None:
        with open(out, "w") as f:
            f.write(text)
        with open(OUT_COL, "w") as f:
            f.write(ctext)
        print("wrote %d programs (of %d distinct) to %s; %d column programs (of %d) to %s"
              % (len(items), len(progs), out, len(citems), len(colprogs), OUT_COL))
    return text + ctext


if __name__ == "__main__":
    if "--check" in sys.argv[1:]:           # exit 1 when the committed list is stale
        with open(OUT) as f, open(OUT_COL) as fc:
            sys.exit(0 if f.read() + fc.read() == main(out=None) else 1)
    main()
