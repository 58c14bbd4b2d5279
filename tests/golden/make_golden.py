"""Generate the committed golden fixtures (tests/golden/*.npz) with the float64 oracle.

The reference is a Julia package and Julia is not installed in this image, so the reference itself cannot
produce these vectors; they are outputs of oracle/agref.py (which is pinned to the reference's own
known-answer vectors by tests/test_oracle_known_answers.py) on fixed seeded inputs.  They freeze the oracle
(regression) and give the GPU tests an input/output set that does not depend on re-running the oracle.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import agref  # noqa: E402
import __graft_entry__ as ge  # noqa: E402  (workload definitions only; no device call)

W = ge.load_package().workloads


def grads(loss, params):
    ps = [agref.Param(np.asfortranarray(p.astype(np.float64))) for p in params]
    t = agref.diff(lambda: loss(ps))
    return float(agref.value(t)), [np.asarray(agref.full(agref.grad(t, p))) for p in ps]


def save(name, **arrays):
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrays)
    print("wrote", name, {k: np.shape(v) for k, v in arrays.items()})


def main():
    # config 1: housing (Float64)
    w, x, y = W.housing_init(np.random.default_rng(1))
    loss, g = grads(lambda p: W.housing_loss(agref, p, x, y), w)
    save("housing", w0=w[0], w1=w[1], x=x, y=y, loss=loss, g0=g[0], g1=g[1])
    # config 2: MNIST MLP at the reference batch (100)
    w, x, y = W.mnist_init(np.random.default_rng(2), 100)
    loss, g = grads(lambda p: W.mnist_loss(agref, p, x.astype(np.float64), y.astype(np.float64)), w)
    save("mnist_b100", w0=w[0], w1=w[1], w2=w[2], w3=w[3], x=x, y=y, loss=loss, g0=g[0], g1=g[1], g2=g[2], g3=g[3])
    # config 3: broadcast / unbroadcast sweep (small instance)
    x, brow, bcol = W.sweep_init(np.random.default_rng(3), 64, 48)
    loss, g = grads(lambda p: W.sweep_loss(agref, p[0], p[1], p[2]), [x, brow, bcol])
    save("sweep_64x48", x=x, brow=brow, bcol=bcol, loss=loss, g0=g[0], g1=g[1], g2=g[2])
    # config 4: LSTM with embedding getindex (Sparse scatter-add), small instance
    rng = np.random.default_rng(4)
    H, E, V, B, T = 16, 8, 12, 6, 3
    wd = W.lstm_init(rng, H, E, V, np.float64)
    keys = list(wd)
    ids = np.stack([rng.integers(0, V, B) for _ in range(T)])
    ids[0, 1] = ids[0, 0]
    tg = np.zeros((T, V, B))
    for t in range(T):
        tg[t, rng.integers(0, V, B), np.arange(B)] = 1
    h0 = np.zeros((H, B))
    loss, g = grads(lambda p: W.lstm_loss(agref, dict(zip(keys, p)), list(ids), [np.asfortranarray(a) for a in tg], h0, h0),
                    [wd[k] for k in keys])
    save("lstm_small", ids=ids, targets=tg, loss=loss, **{"w_" + k: wd[k] for k in keys},
         **{"g_" + k: gi for k, gi in zip(keys, g)})
    # Sparse / addtoindex! bookkeeping: destinations (bit-exact) for every index kind of test/getindex.jl:8-26
    shape = (3, 4)
    mask = np.zeros(shape, bool)
    mask[1, 2] = mask[0, 0] = True
    kinds = {"colon": (slice(None),), "int": (0,), "range": (range(0, 2),), "steprange": (range(0, 3, 2),),
             "vec": ([0, 1],), "repeated": ([1, 1],), "empty": ([],), "mask": (mask,), "colon_int": (slice(None), 1),
             "int_colon": (1, slice(None)), "outer": ([0, 2], [1, 1, 3]), "matrix_index": (np.array([[0, 1], [2, 2]]),)}
    save("flat_dest_3x4", **{k: agref.flat_dest(shape, idx)[0] for k, idx in kinds.items()})
    # src/specialfunctions.jl and norm(x, p): values and gradients on a fixed grid (Float64)
    grid = np.linspace(0.35, 3.6, 27)
    doms = {"erfinv": grid / 4.0, "erfcinv": grid / 2.0, "erf": grid - 1.5, "erfc": grid - 1.5}
    out = {"grid": grid}
    for name in ("erf", "erfc", "erfcx", "erfinv", "erfcinv", "gamma", "lgamma", "digamma", "trigamma", "besselj0",
                 "besselj1", "bessely0", "bessely1"):
        xs = doms.get(name, grid)
        f = getattr(agref, name)
        out["x_" + name], out["y_" + name] = xs, f(xs)
        out["g_" + name] = agref.grad(lambda v: agref.sum_(f(v)))(xs)
    xn = np.asfortranarray(np.random.default_rng(6).standard_normal((5, 6)))
    out["norm_x"] = xn
    for p in (1, 2, 3, 2.5, np.inf):
        tag = str(p).replace(".", "_")
        out["norm_" + tag] = agref.norm(xn, p)
        out["gnorm_" + tag] = agref.grad(lambda v: agref.norm(v, p))(xn)
    save("special_and_norm", **out)
    # the MNIST loader's conversions (examples/mnist.jl:80-81) on a fixed byte batch
    rb = np.random.default_rng(7)
    px, lb = np.asfortranarray(rb.integers(0, 256, (784, 12), dtype=np.uint8)), rb.integers(0, 10, 12, dtype=np.uint8)
    save("mnist_bytes", pixels=px, labels=lb, x=agref.xshape(px), y=agref.yshape(lb))


if __name__ == "__main__":
    main()
