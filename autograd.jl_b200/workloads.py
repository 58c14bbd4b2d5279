"""The reference's example models restated once, against an abstract primitive namespace `m`
(this package, or any module with the same primitive names), in Julia-1.x form:

  housing   examples/housing.jl:23-27 / README.md:70-77   sum(abs2, w*x .+ b - y)
  mnist     examples/mnist.jl:30-42,65-75                 784-64-10 relu MLP, softmax cross-entropy
  sweep     BASELINE.json configs[2]                      sum(abs2, tanh.(exp.(0.1 .* x .+ brow) .+ bcol))
  charlm    examples/charlm.jl:55-114                     LSTM char LM, embedding as W[:, ids]
  rnn       examples/rnn_lonely_integer.jl:41-79          tanh RNN + softmax loss

Data is synthetic with the shapes of each config; arrays are numpy (column-major meaning) and are
moved to the device by the caller.
"""
import numpy as np


# ---- housing --------------------------------------------------------------------------------------
def housing_init(rng, n=506, d=13, dtype=np.float64):
    w = (0.1 * rng.standard_normal((1, d))).astype(dtype)
    b = np.zeros((1, 1), dtype)
    x = rng.standard_normal((d, n)).astype(dtype)
    y = rng.standard_normal((1, n)).astype(dtype)
    return [w, b], x, y


def housing_loss(m, w, x, y):
    return m.div(m.sumabs2(m.sub(m.add(m.matmul(w[0], x), w[1]), y)), x.shape[1] if hasattr(x, "shape") else 1)


# ---- MNIST MLP --------------------------------------------------------------------------------------
def mnist_init(rng, batch, dtype=np.float32, hidden=64, nin=784, nout=10):
    w = [(0.1 * rng.standard_normal((hidden, nin))).astype(dtype), np.zeros((hidden,), dtype),
         (0.1 * rng.standard_normal((nout, hidden))).astype(dtype), np.zeros((nout,), dtype)]
    x = rng.random((nin, batch), dtype=np.float32).astype(dtype)
    labels = rng.integers(0, nout, batch)
    y = np.zeros((nout, batch), dtype)
    y[labels, np.arange(batch)] = 1
    return w, np.asfortranarray(x), np.asfortranarray(y)


def mnist_ubyte_batch(rng, batch, nin=784):
    """A synthetic minibatch in the dataset's byte format (idx-ubyte, examples/mnist.jl:83-86): pixels 0..255 as
    (nin, batch) uint8 and digit labels 0..9."""
    return (np.asfortranarray(rng.integers(0, 256, (nin, batch), dtype=np.uint8)),
            rng.integers(0, 10, batch, dtype=np.uint8))


def mnist_from_ubyte(pixels, labels, dtype=np.float32, nout=10):
    """The host arrays the reference builds from the bytes (examples/mnist.jl:80-81): x = pixels ./ 255f0 and the
    one-hot y with digit 0 in row 10."""
    x = np.asfortranarray(pixels.astype(dtype) / dtype(255))
    cls = labels.astype(np.int64)
    cls[cls == 0] = 10
    y = np.zeros((nout, labels.size), dtype, order="F")
    y[cls - 1, np.arange(labels.size)] = 1
    return x, y


def mnist_predict(m, w, x):
    a = m.relu(m.add(m.matmul(w[0], x), w[1]))
    return m.add(m.matmul(w[2], a), w[3])


def mnist_loss(m, w, x, y, global_batch=None):
    p = mnist_predict(m, w, x)
    logz = m.log(m.sum_(m.exp(p), dims=0))
    nb = global_batch if global_batch is not None else y.shape[1]
    return m.div(m.neg(m.sum_(m.mul(y, m.sub(p, logz)))), nb)


def mnist_train_step(m, dp, params, x, y, lr, state, global_batch=None):
    """One data-parallel SGD step of the MLP (examples/mnist.jl:52-62 with the update of README.md:73-76):
    @diff + grad for every Param, then allreduce + axpy! through `dp.update`.  `state` keeps the loss and the
    gradients alive exactly as bench.py does (what stays referenced decides which values a fused chain stores)."""
    t = m.diff(lambda: mnist_loss(m, params, x, y, global_batch=global_batch))
    state["loss"] = m.value(t)
    state["grads"] = [m.grad(t, p) for p in params]
    del t
    state["grads"] = dp.update(-lr, state["grads"], params)


# ---- broadcast / unbroadcast sweep ---------------------------------------------------------------
def sweep_init(rng, rows, cols, dtype=np.float32):
    x = np.asfortranarray(rng.standard_normal((rows, cols), dtype=np.float32).astype(dtype))
    brow = (0.1 * rng.standard_normal((1, cols))).astype(dtype)
    bcol = (0.1 * rng.standard_normal((rows, 1))).astype(dtype)
    return x, brow, bcol


def sweep_loss(m, x, brow, bcol):
    return m.sumabs2(m.tanh(m.add(m.exp(m.add(m.mul(0.1, x), brow)), bcol)))


# ---- charlm LSTM ---------------------------------------------------------------------------------------
def _xavier(rng, shape, dtype):
    fan_out, fan_in = shape[0], shape[1]
    s = np.sqrt(2.0 / (fan_in + fan_out))
    return (s * rng.standard_normal(shape)).astype(dtype)


def lstm_init(rng, hidden=512, embed=128, vocab=128, dtype=np.float32):
    """Weights in the order of examples/charlm.jl:55-66 (W,b per gate) plus embedding and predict."""
    w = {}
    for g in ("forget", "ingate", "outgate", "change"):
        w["W_" + g] = _xavier(rng, (hidden, embed + hidden), dtype)
        w["b_" + g] = (np.ones((hidden, 1), dtype) if g == "forget" else np.zeros((hidden, 1), dtype))
    w["W_embedding"] = _xavier(rng, (embed, vocab), dtype)
    w["W_predict"] = _xavier(rng, (vocab, hidden), dtype)
    w["b_predict"] = np.zeros((vocab, 1), dtype)
    return w


def lstm_cell(m, w, inp, hidden, cell):
    x = m.vcat(inp, hidden)
    forget = m.sigm(m.add(m.matmul(w["W_forget"], x), w["b_forget"]))
    ingate = m.sigm(m.add(m.matmul(w["W_ingate"], x), w["b_ingate"]))
    outgate = m.sigm(m.add(m.matmul(w["W_outgate"], x), w["b_outgate"]))
    change = m.tanh(m.add(m.matmul(w["W_change"], x), w["b_change"]))
    cell = m.add(m.mul(cell, forget), m.mul(ingate, change))
    hidden = m.mul(outgate, m.tanh(cell))
    return hidden, cell


def lstm_loss(m, w, ids, targets, h0, c0, global_batch=None):
    """Sum over time of the softmax cross-entropy of the next-char prediction
    (examples/charlm.jl:93-114); ids[t] int vectors (0-based), targets[t] one-hot (V,B)."""
    hidden, cell = h0, c0
    total = 0.0
    nb = global_batch if global_batch is not None else targets[0].shape[1]
    for t in range(len(ids)):
        inp = m.getindex(w["W_embedding"], slice(None), ids[t])
        hidden, cell = lstm_cell(m, w, inp, hidden, cell)
        ypred = m.add(m.matmul(w["W_predict"], hidden), w["b_predict"])
        logz = m.log(m.sum_(m.exp(ypred), dims=0))
        total = m.add(total, m.sum_(m.mul(targets[t], m.sub(ypred, logz))))
    return m.div(m.neg(total), nb)


# ---- RNN (lonely integer) -----------------------------------------------------------------------------
def rnn_init(rng, hidden=256, vocab=100, dtype=np.float32):
    return [(0.1 * rng.standard_normal((hidden, vocab))).astype(dtype),     # W_xh
            (0.1 * rng.standard_normal((hidden, hidden))).astype(dtype),    # W_hh
            np.zeros((hidden, 1), dtype),                                   # b_h
            (0.1 * rng.standard_normal((vocab, hidden))).astype(dtype),     # W_hy
            np.zeros((vocab, 1), dtype)]                                    # b_y


def rnn_loss(m, w, xs, y, h0, global_batch=None):
    h = h0
    for x in xs:
        h = m.tanh(m.add(m.add(m.matmul(w[0], x), m.matmul(w[1], h)), w[2]))
    p = m.add(m.matmul(w[3], h), w[4])
    logz = m.log(m.sum_(m.exp(p), dims=0))
    nb = global_batch if global_batch is not None else y.shape[1]
    return m.div(m.neg(m.sum_(m.mul(y, m.sub(p, logz)))), nb)
