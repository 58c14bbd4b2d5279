"""Independent second opinion for the oracle: the five BASELINE configs written directly in
torch (CPU, float64, torch.autograd) -- NOT through workloads.py and not through any tape of this repo -- against
oracle/agref.py running the shared workload definitions.  torch is not the reference, so this gates nothing by
itself; it guards against a misreading of the reference that the oracle and the product's host logic (which are
close restatements of the same Julia source) would share.  CPU only.
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

TOL = 1e-10


def relerr(got, ref):
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    return np.max(np.abs(got - ref)) / max(np.max(np.abs(ref)), 1e-300)


def T(a, grad=False):
    t = torch.tensor(np.ascontiguousarray(a), dtype=torch.float64)
    t.requires_grad_(grad)
    return t


@pytest.fixture(scope="module")
def W(agpkg):
    return agpkg.workloads


def oracle_grads(oracle, loss, params):
    po = [oracle.Param(np.asfortranarray(np.asarray(p, dtype=np.float64))) for p in params]
    t = oracle.diff(lambda: loss(po))
    return float(oracle.value(t)), [np.asarray(oracle.full(oracle.grad(t, p))) for p in po]


def test_housing_vs_torch(oracle, W):
    w, x, y = W.housing_init(np.random.default_rng(1))
    lo, go = oracle_grads(oracle, lambda p: W.housing_loss(oracle, p, x, y), w)
    tw, tb = T(w[0], True), T(w[1], True)
    lt = ((tw @ T(x) + tb - T(y)) ** 2).sum() / x.shape[1]
    lt.backward()
    assert abs(lo - lt.item()) <= TOL * abs(lo)
    assert relerr(go[0], tw.grad.numpy()) <= TOL and relerr(go[1], tb.grad.numpy()) <= TOL


def test_mnist_vs_torch(oracle, W):
    w, x, y = W.mnist_init(np.random.default_rng(2), 257, np.float64)
    lo, go = oracle_grads(oracle, lambda p: W.mnist_loss(oracle, p, x, y), w)
    tw = [T(a, True) for a in w]
    h = torch.relu(tw[0] @ T(x) + tw[1][:, None])
    logits = tw[2] @ h + tw[3][:, None]
    # cross entropy with integer labels, mean over the batch (examples/mnist.jl:38-42)
    lt = torch.nn.functional.cross_entropy(logits.T, torch.tensor(y.argmax(0)), reduction="mean")
    lt.backward()
    assert abs(lo - lt.item()) <= TOL * abs(lo)
    for a, b in zip(go, tw):
        assert relerr(a, b.grad.numpy()) <= TOL


def test_sweep_vs_torch(oracle, W):
    x, brow, bcol = W.sweep_init(np.random.default_rng(3), 96, 130, np.float64)
    lo, go = oracle_grads(oracle, lambda p: W.sweep_loss(oracle, p[0], p[1], p[2]), [x, brow, bcol])
    tx, tr, tc = T(x, True), T(brow, True), T(bcol, True)
    lt = (torch.tanh(torch.exp(0.1 * tx + tr) + tc) ** 2).sum()
    lt.backward()
    assert abs(lo - lt.item()) <= TOL * abs(lo)
    for a, b in zip(go, (tx, tr, tc)):
        assert relerr(a, b.grad.numpy()) <= TOL


def test_lstm_with_clipped_update_vs_torch(oracle, W):
    """Config 4 (examples/charlm.jl:55-118,148-152) with torch.nn.functional.embedding + a hand-written cell."""
    rng = np.random.default_rng(4)
    H, E, V, B, Tn = 24, 12, 17, 9, 5
    w = W.lstm_init(rng, H, E, V, np.float64)
    keys = list(w)
    ids = [rng.integers(0, V, B) for _ in range(Tn)]
    lab = [rng.integers(0, V, B) for _ in range(Tn)]
    tg = []
    for t in range(Tn):
        oh = np.zeros((V, B))
        oh[lab[t], np.arange(B)] = 1
        tg.append(np.asfortranarray(oh))
    h0 = np.zeros((H, B))
    lo, go = oracle_grads(oracle, lambda p: W.lstm_loss(oracle, dict(zip(keys, p)), ids, tg, h0, h0),
                          [w[k] for k in keys])
    tw = {k: T(w[k], True) for k in keys}
    h = c = torch.zeros(H, B, dtype=torch.float64)
    total = 0.0
    for t in range(Tn):
        emb = torch.nn.functional.embedding(torch.tensor(ids[t]), tw["W_embedding"].T).T     # (E, B)
        xh = torch.cat([emb, h], 0)
        f = torch.sigmoid(tw["W_forget"] @ xh + tw["b_forget"])
        i = torch.sigmoid(tw["W_ingate"] @ xh + tw["b_ingate"])
        o = torch.sigmoid(tw["W_outgate"] @ xh + tw["b_outgate"])
        g = torch.tanh(tw["W_change"] @ xh + tw["b_change"])
        c = c * f + i * g
        h = o * torch.tanh(c)
        logits = tw["W_predict"] @ h + tw["b_predict"]
        total = total + torch.nn.functional.cross_entropy(logits.T, torch.tensor(lab[t]), reduction="sum")
    lt = total / B
    lt.backward()
    assert abs(lo - lt.item()) <= TOL * abs(lo)
    for k, a in zip(keys, go):
        assert relerr(a, tw[k].grad.numpy()) <= TOL, k
    # gclip (examples/charlm.jl:116-118,148-152) against torch's clip_grad_norm_
    gn_o = np.sqrt(sum(float((a ** 2).sum()) for a in go))
    gn_t = float(torch.nn.utils.clip_grad_norm_(list(tw.values()), 0.5 * gn_o))
    assert abs(gn_o - gn_t) <= TOL * gn_o
    scale = min(1.0, 0.5 * gn_o / gn_o)
    for k, a in zip(keys, go):
        assert relerr(scale * a, tw[k].grad.numpy()) <= 1e-6, k       # torch divides by (norm + 1e-6)


def test_rnn_grad_of_grad_vs_torch(oracle, W):
    """Config 5 (test/highorder.jl:36-42 pattern): torch.autograd.grad(create_graph=True) twice."""
    rng = np.random.default_rng(6)
    H, V, B, Tn = 14, 9, 21, 4
    w = W.rnn_init(rng, H, V, np.float64)
    xs, labx = [], []
    for t in range(Tn):
        oh = np.zeros((V, B))
        oh[rng.integers(0, V, B), np.arange(B)] = 1
        xs.append(np.asfortranarray(oh))
    lab = rng.integers(0, V, B)
    y = np.zeros((V, B))
    y[lab, np.arange(B)] = 1
    h0 = np.zeros((H, B))
    inner = oracle.grad(lambda whh: W.rnn_loss(oracle, [w[0], whh, w[2], w[3], w[4]], xs, y, h0))
    outer = oracle.grad(lambda whh: oracle.sumabs2(inner(whh)))
    g1o, g2o = inner(np.asfortranarray(w[1].copy())), outer(np.asfortranarray(w[1].copy()))
    tw = [T(a, i == 1) for i, a in enumerate(w)]
    h = torch.zeros(H, B, dtype=torch.float64)
    for x in xs:
        h = torch.tanh(tw[0] @ T(x) + tw[1] @ h + tw[2])
    logits = tw[3] @ h + tw[4]
    lt = torch.nn.functional.cross_entropy(logits.T, torch.tensor(lab), reduction="mean")
    (g1,) = torch.autograd.grad(lt, tw[1], create_graph=True)
    (g2,) = torch.autograd.grad((g1 ** 2).sum(), tw[1])
    assert relerr(g1o, g1.detach().numpy()) <= TOL
    assert relerr(g2o, g2.numpy()) <= 1e-9
