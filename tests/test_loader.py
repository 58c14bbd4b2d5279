"""Input formats either side of the path: the oracle's restatement of examples/mnist.jl:80-81
(xshape / yshape), the host logic of UByteBatch on the numpy stand-in, and -- on the GPU -- bit-exact parity of the
device conversions with the oracle, plus a whole MNIST step fed from bytes against the same step fed from the
reference's Float32 arrays."""
import gzip
import struct

import numpy as np
import pytest


def test_oracle_xshape_yshape_known_answers(oracle):
    # examples/mnist.jl:80-81 by hand: pixels/255 in Float32; label 0 -> class 10, label d -> class d (1-based rows)
    x = oracle.xshape(np.array([0, 255, 51, 128], np.uint8), nfeat=2)
    assert x.dtype == np.float32 and x.shape == (2, 2)
    assert np.array_equal(x, np.array([[0.0, np.float32(51) / np.float32(255)], [1.0, np.float32(128) / np.float32(255)]], np.float32))
    y = oracle.yshape(np.array([0, 1, 9, 3], np.uint8))
    assert y.shape == (10, 4) and y.sum() == 4
    assert [int(np.argmax(y[:, j])) for j in range(4)] == [9, 0, 8, 2]


def test_idx_reader_roundtrip(agpkg, tmp_path):
    rng = np.random.default_rng(0)
    imgs = rng.integers(0, 256, (5, 4, 3), dtype=np.uint8)           # 5 images of 4x3
    labels = rng.integers(0, 10, 5, dtype=np.uint8)
    pi, pl = tmp_path / "img-idx3-ubyte.gz", tmp_path / "lab-idx1-ubyte"
    with gzip.open(pi, "wb") as f:
        f.write(struct.pack(">HBBIII", 0, 8, 3, 5, 4, 3) + imgs.tobytes())
    with open(pl, "wb") as f:
        f.write(struct.pack(">HBBI", 0, 8, 1, 5) + labels.tobytes())
    a = agpkg.read_idx_ubyte(pi)
    assert a.shape == (12, 5) and np.array_equal(a[:, 2], imgs[2].reshape(-1))
    assert np.array_equal(agpkg.read_idx_ubyte(pl), labels)
    with open(pl, "wb") as f:
        f.write(b"\x00\x00\x0d\x01" + b"\x00" * 8)
    with pytest.raises(ValueError):
        agpkg.read_idx_ubyte(pl)


def _batch(rng, nfeat, ncols):
    px = np.asfortranarray(rng.integers(0, 256, (nfeat, ncols), dtype=np.uint8))
    lb = rng.integers(0, 10, ncols, dtype=np.uint8)
    return px, lb


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_ubyte_batch_host_logic(hostlib, oracle, dtype):
    ag = hostlib
    px, lb = _batch(np.random.default_rng(1), 12, 7)
    b = ag.UByteBatch(12, 7, dtype=dtype)
    x, y = b.load(px, lb)
    assert np.array_equal(x.to_numpy(), oracle.xshape(px, 12, dtype))
    assert np.array_equal(y.to_numpy(), oracle.yshape(lb, 10, dtype))
    assert b.h2d_bytes == 12 * 7 + 7
    with pytest.raises(ag.DimensionMismatch):
        b.load(px[:, :3], lb)
    with pytest.raises(TypeError):
        b.load(px.astype(np.float32), lb)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("nfeat,ncols", [(784, 100), (784, 4099), (13, 1), (33, 515)])
def test_device_conversions_match_the_oracle_bit_for_bit(ag, oracle, dtype, nfeat, ncols):
    px, lb = _batch(np.random.default_rng(2), nfeat, ncols)
    b = ag.UByteBatch(nfeat, ncols, dtype=dtype)
    x, y = b.load(px, lb)
    assert np.array_equal(x.to_numpy(), oracle.xshape(px, nfeat, dtype))
    assert np.array_equal(y.to_numpy(), oracle.yshape(lb, 10, dtype))
    px2, lb2 = _batch(np.random.default_rng(3), nfeat, ncols)          # in place: same device arrays
    p0 = x.ptr
    x2, y2 = b.load(px2, lb2)
    assert x2.ptr == p0 and np.array_equal(x2.to_numpy(), oracle.xshape(px2, nfeat, dtype))
    assert np.array_equal(y2.to_numpy(), oracle.yshape(lb2, 10, dtype))


@pytest.mark.gpu
def test_prefetched_batches_arrive_in_order(ag, oracle):
    """prefetch (copy stream) / convert (main stream): batch i+1 crosses PCIe while batch i is in use, and every
    batch still reaches the device arrays intact and in order; same for Float32 arrays through StagedInput."""
    rng = np.random.default_rng(8)
    nfeat, ncols, nb = 784, 2048, 6
    hosts = []
    for _ in range(nb):
        px, lb = ag.pinned_empty((nfeat, ncols), np.uint8), ag.pinned_empty((ncols,), np.uint8)
        p, l = _batch(rng, nfeat, ncols)
        px[...] = p
        lb[...] = l
        hosts.append((px, lb))
    b = ag.UByteBatch(nfeat, ncols)
    b.prefetch(*hosts[0])
    sums = []
    for i in range(nb):
        x, y = b.convert()
        if i + 1 < nb:
            b.prefetch(*hosts[i + 1])                  # overlaps the "step" below
        sums.append((ag.sumabs2(ag.matmul(y, ag.adjoint(x))), x.to_numpy() if i in (0, nb - 1) else None))
    for i, (s_, xs) in enumerate(sums):
        xr, yr = oracle.xshape(hosts[i][0]), oracle.yshape(hosts[i][1])
        ref = float(np.sum((yr.astype(np.float64) @ xr.astype(np.float64).T) ** 2))
        assert abs(float(s_) - ref) <= 1e-5 * ref
        if xs is not None:
            assert np.array_equal(xs, xr)
    xf = [ag.pinned_empty((64, 512), np.float32) for _ in range(3)]
    for k, h in enumerate(xf):
        h[...] = rng.standard_normal((64, 512)).astype(np.float32) + k
    xd = ag.DArray.empty((64, 512), np.float32)
    st = ag.StagedInput(xd)
    st.prefetch(xf[0].host_ptr)
    for k in range(3):
        st.commit()
        if k + 1 < 3:
            st.prefetch(xf[k + 1].host_ptr)
        assert np.array_equal(xd.to_numpy(), np.asarray(xf[k]))


@pytest.mark.gpu
def test_bad_label_raises_the_sticky_flag(ag):
    b = ag.UByteBatch(4, 3)
    b.load(np.zeros((4, 3), np.uint8), np.array([1, 200, 3], np.uint8))
    with pytest.raises(IndexError):
        ag.sync()
    ag.sync()


@pytest.mark.gpu
def test_mnist_step_from_bytes_equals_step_from_float_arrays(ag, oracle):
    W = ag.workloads
    rng = np.random.default_rng(4)
    B = 2048
    w, _, _ = W.mnist_init(rng, 1)
    px, lb = _batch(rng, 784, B)
    xf, yf = oracle.xshape(px), oracle.yshape(lb)

    def run(xd, yd):
        p = [ag.Param(ag.DArray.from_numpy(a)) for a in w]
        t = ag.diff(lambda: W.mnist_loss(ag, p, xd, yd))
        return [float(ag.value(t))] + [ag.grad(t, q).to_numpy() for q in p]
    bt = ag.UByteBatch(784, B)
    r1 = run(*bt.load(px, lb))
    r2 = run(ag.DArray.from_numpy(xf), ag.DArray.from_numpy(yf))
    for a, c in zip(r1, r2):
        assert np.array_equal(np.asarray(a), np.asarray(c))
    # and against the float64 oracle tape on the reference's own arrays
    po = [oracle.Param(np.asfortranarray(a.astype(np.float64))) for a in w]
    to = oracle.diff(lambda: W.mnist_loss(oracle, po, xf.astype(np.float64), yf.astype(np.float64)))
    for q, g in zip(po, r1[1:]):
        ref = oracle.full(oracle.grad(to, q))
        assert np.max(np.abs(g - ref)) <= 1e-5 * np.max(np.abs(ref))


TEXT = "the quick brown fox jumps over the lazy dog; THE QUICK BROWN FOX!\n" * 7


def test_charlm_data_matches_the_reference_pipeline(agpkg, oracle):
    """read_chars / charlm_minibatch_ids against the restated loaddata + minibatch of examples/charlm.jl:166-203:
    same vocabulary order, and the id of stream j in batch i is the row of the 1 in column j of the reference's
    one-hot matrix."""
    ag = agpkg
    data, chars, c2i, i2c = oracle.charlm_loaddata(TEXT, 16)
    ids, c2i0, i2c0 = ag.read_chars(TEXT)
    assert i2c0 == i2c and all(c2i0[c] == c2i[c] - 1 for c in c2i)
    assert "".join(i2c0[k] for k in ids) == TEXT
    b = ag.charlm_minibatch_ids(ids, 16)
    assert b.shape == (len(data), 16)
    for i, d in enumerate(data):
        assert np.array_equal(np.argmax(d, axis=0), b[i]) and d.sum() == 16
    ids2, _, _ = ag.read_chars(TEXT, char_limit=50)
    assert ids2.size == 50 and np.array_equal(ids2, ids[:50])


@pytest.mark.gpu
def test_charlm_batches_on_device(ag, oracle):
    data, chars, c2i, i2c = oracle.charlm_loaddata(TEXT, 16)
    ids, _, i2c0 = ag.read_chars(TEXT)
    cb = ag.CharBatches(ag.charlm_minibatch_ids(ids, 16)[:8], len(i2c0))
    W = ag.DArray.from_numpy(np.asfortranarray(np.random.default_rng(0).standard_normal((5, len(i2c0))).astype(np.float32)))
    for t in range(8):
        assert np.array_equal(cb.onehot(t).to_numpy(), data[t])                      # the reference's matrix, bit for bit
        emb = ag.getindex(W, slice(None), cb.ids[t]).to_numpy()                      # W[:, ids] == W * onehot
        assert np.array_equal(emb, W.to_numpy() @ data[t])
