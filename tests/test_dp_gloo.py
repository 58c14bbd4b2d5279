"""CPU test of the data-parallel path at world size 2 (gloo): each rank runs its own tape on a contiguous
column shard, gradients are bucketed and all-reduced, and must equal the single-process full-batch
gradient of the oracle.  The device ABI is replaced by tests/fake_abi.py; the collective inside
ag_allreduce_sum is delegated to gloo."""
import os
import socket
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    import __graft_entry__ as ge
    import fake_abi
    ag = ge.load_package()

    def allreduce(v):
        t = torch.from_numpy(np.array(v, copy=True))
        dist.all_reduce(t)
        return t.numpy()
    fake_abi.install(ag, fake_abi.FakeLib(allreduce=allreduce))

    def bcast(payload):
        obj = [payload]
        dist.broadcast_object_list(obj, src=0)
        return obj[0]
    def allgather(payload):
        out = [None] * world
        dist.all_gather_object(out, payload)
        return out
    dp = ag.DataParallel(rank, world, bcast, allgather if os.environ.get('AGB_TEST_P2P') else None)
    W = ag.workloads
    B = 96
    w, x, y = W.mnist_init(np.random.default_rng(11), B, np.float64)
    lo, hi = dp.shard_columns(B, rank, world)
    params = [ag.Param(ag.DArray.from_numpy(a)) for a in w]
    xd, yd = ag.DArray.from_numpy(x[:, lo:hi]), ag.DArray.from_numpy(y[:, lo:hi])
    t = ag.diff(lambda: W.mnist_loss(ag, params, xd, yd, global_batch=B))
    grads = dp.update(-0.1, [ag.grad(t, p) for p in params], params)      # allreduce + fused axpy!
    # a second, clipped step on fresh Params (examples/charlm.jl:148-152): allreduce first, then the norm of the SUMMED
    # gradients decides the scale, identically on every rank
    params2 = [ag.Param(ag.DArray.from_numpy(a)) for a in w]
    t2 = ag.diff(lambda: W.mnist_loss(ag, params2, xd, yd, global_batch=B))
    dp.update(-0.1, [ag.grad(t2, p) for p in params2], params2, gclip=0.05)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), *[p.value.to_numpy() for p in params],
             *[g.to_numpy() for g in grads], *[p.value.to_numpy() for p in params2])
    # config 4 (charlm LSTM): the batch dimension of ids[t], targets and the recurrent state is split the same way;
    # the embedding gradient is a Sparse on every rank and is densified before the reduction
    H, E, V, Bl, T = 12, 6, 9, 10, 3
    rl = np.random.default_rng(5)
    wd = W.lstm_init(rl, H, E, V, np.float64)
    keys = list(wd)
    ids = [rl.integers(0, V, Bl) for _ in range(T)]
    tg = []
    for _ in range(T):
        oh = np.zeros((V, Bl))
        oh[rl.integers(0, V, Bl), np.arange(Bl)] = 1
        tg.append(oh)
    lo, hi = dp.shard_columns(Bl, rank, world)
    pl = [ag.Param(ag.DArray.from_numpy(wd[k])) for k in keys]
    ids_r = [ag.IArray(i[lo:hi]) for i in ids]
    tg_r = [ag.DArray.from_numpy(np.asfortranarray(t_[:, lo:hi])) for t_ in tg]
    h0 = ag.zeros((H, hi - lo), np.float64)
    tl = ag.diff(lambda: W.lstm_loss(ag, dict(zip(keys, pl)), ids_r, tg_r, h0, h0, global_batch=Bl))
    gl = dp.update(-0.05, [ag.grad(tl, p) for p in pl], pl)
    np.savez(os.path.join(out_dir, "lstm_rank%d.npz" % rank), *[p.value.to_numpy() for p in pl],
             *[ag.full(g).to_numpy() for g in gl])
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradients_match_full_batch(tmp_path, oracle):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    W = ge.load_package().workloads
    B = 96
    w, x, y = W.mnist_init(np.random.default_rng(11), B, np.float64)
    po = [oracle.Param(np.asfortranarray(a)) for a in w]
    to = oracle.diff(lambda: W.mnist_loss(oracle, po, x, y))
    gref = [oracle.grad(to, p) for p in po]
    r0 = np.load(os.path.join(str(tmp_path), "rank0.npz"))
    r1 = np.load(os.path.join(str(tmp_path), "rank1.npz"))
    for i, g in enumerate(gref):
        got0, got1 = r0["arr_%d" % (4 + i)], r1["arr_%d" % (4 + i)]
        assert np.array_equal(got0, got1)                       # every rank holds the same reduced gradient
        assert np.max(np.abs(got0 - g)) <= 1e-12 * max(np.max(np.abs(g)), 1e-30)
        want = w[i] - 0.1 * g                                   # identical axpy! on every rank
        assert np.max(np.abs(r0["arr_%d" % i] - want)) <= 1e-12
        assert np.array_equal(r0["arr_%d" % i], r1["arr_%d" % i])
    gn = np.sqrt(sum(float((g ** 2).sum()) for g in gref))
    assert gn > 0.05                                            # so the clipped step really is scaled
    for i, g in enumerate(gref):
        want = w[i] - 0.1 * (0.05 / gn) * g
        assert np.max(np.abs(r0["arr_%d" % (8 + i)] - want)) <= 1e-12
        assert np.array_equal(r0["arr_%d" % (8 + i)], r1["arr_%d" % (8 + i)])


def test_two_rank_lstm_matches_full_batch(tmp_path, oracle):
    """Runs after the worker test above in the same spawn (files are written there); the LSTM part: reduced
    gradients (dense and the Sparse embedding gradient) and updated Params equal the oracle's full-batch step."""
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    W = ge.load_package().workloads
    H, E, V, Bl, T = 12, 6, 9, 10, 3
    rl = np.random.default_rng(5)
    wd = W.lstm_init(rl, H, E, V, np.float64)
    keys = list(wd)
    ids = [rl.integers(0, V, Bl) for _ in range(T)]
    tg = []
    for _ in range(T):
        oh = np.zeros((V, Bl))
        oh[rl.integers(0, V, Bl), np.arange(Bl)] = 1
        tg.append(np.asfortranarray(oh))
    po = [oracle.Param(np.asfortranarray(wd[k])) for k in keys]
    h0 = np.zeros((H, Bl))
    to = oracle.diff(lambda: W.lstm_loss(oracle, dict(zip(keys, po)), ids, tg, h0, h0))
    gref = [oracle.full(oracle.grad(to, p)) for p in po]
    r0 = np.load(os.path.join(str(tmp_path), "lstm_rank0.npz"))
    r1 = np.load(os.path.join(str(tmp_path), "lstm_rank1.npz"))
    n = len(keys)
    for i, (k, g) in enumerate(zip(keys, gref)):
        got0, got1 = r0["arr_%d" % (n + i)], r1["arr_%d" % (n + i)]
        assert np.array_equal(got0, got1), k
        assert np.max(np.abs(got0 - g)) <= 1e-11 * max(np.max(np.abs(g)), 1e-30), k
        assert np.max(np.abs(r0["arr_%d" % i] - (wd[k] - 0.05 * g))) <= 1e-11, k
        assert np.array_equal(r0["arr_%d" % i], r1["arr_%d" % i]), k


def test_shard_columns_cover_the_batch(agpkg):
    for n in (1, 7, 64, 65536):
        for world in (1, 2, 4, 8):
            spans = [agpkg.DataParallel.shard_columns(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
