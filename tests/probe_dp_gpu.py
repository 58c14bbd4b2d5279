"""Multi-GPU check (run under torchrun on the GPU box): sharded tapes + peer-memory allreduce vs NCCL vs the
full-batch float64 oracle.  torchrun --nproc-per-node N tests/probe_dp_gpu.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch.distributed as dist
import __graft_entry__ as ge
from oracle import agref

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dist.init_process_group("gloo", rank=rank, world_size=world)
ag = ge.load_package()
ag.init(local)


def bcast(p):
    o = [p]; dist.broadcast_object_list(o, src=0); return o[0]


def allgather(p):
    o = [None] * world; dist.all_gather_object(o, p); return o


dp = ag.DataParallel(rank, world, bcast, allgather)
dist.barrier()
W = ag.workloads
B = 512 * world
w, x, y = W.mnist_init(np.random.default_rng(5), B)
lo, hi = dp.shard_columns(B, rank, world)
params = [ag.Param(ag.DArray.from_numpy(a)) for a in w]
xd, yd = ag.DArray.from_numpy(np.asfortranarray(x[:, lo:hi])), ag.DArray.from_numpy(np.asfortranarray(y[:, lo:hi]))
ok = True
for it in range(5):          # several epochs: exercises the double-buffered flags
    t = ag.diff(lambda: W.mnist_loss(ag, params, xd, yd, global_batch=B))
    local_g = [ag.grad(t, p) for p in params]
    g_p2p = [g.to_numpy() for g in dp.allreduce_grads(local_g)]
    cap, dp.p2p_cap = dp.p2p_cap, 0                       # force NCCL on the same inputs
    g_nccl = [g.to_numpy() for g in dp.allreduce_grads(local_g)]
    dp.p2p_cap = cap
    for a, b in zip(g_p2p, g_nccl):
        ok &= bool(np.max(np.abs(a - b)) <= 1e-6 * max(np.max(np.abs(b)), 1e-30))
po = [agref.Param(np.asfortranarray(a.astype(np.float64))) for a in w]
to = agref.diff(lambda: W.mnist_loss(agref, po, x.astype(np.float64), y.astype(np.float64)))
for a, p in zip(g_p2p, po):
    ref = agref.grad(to, p)
    ok &= bool(np.max(np.abs(a - ref)) <= 1e-5 * np.max(np.abs(ref)))
# fused allreduce + axpy!: params must become w - lr * global gradient on every rank
t = ag.diff(lambda: W.mnist_loss(ag, params, xd, yd, global_batch=B))
dp.update(-0.1, [ag.grad(t, p) for p in params], params)
for p, a, w0 in zip(params, po, w):
    want = w0.astype(np.float64) - 0.1 * agref.grad(to, a)
    err = np.max(np.abs(p.value.to_numpy() - want)) / max(np.max(np.abs(want)), 1e-30)
    gerr = np.max(np.abs(ag.grad(t, p).to_numpy() - agref.grad(to, a))) / np.max(np.abs(agref.grad(to, a)))
    print("rank %d update check shape %s: rel err %.3e  (reduced grad err %.3e) ok_so_far=%s" % (rank, want.shape, err, gerr, ok), flush=True)
    ok &= bool(err <= 1e-5)
# config 4: the LSTM sharded over the batch dimension (Sparse embedding gradient -> dense -> reduction), with the
# clipped update of examples/charlm.jl:148-152, against the oracle's full-batch step
H, E, V, T = 64, 32, 40, 4
Bl = 24 * world
rl = np.random.default_rng(17)
wdl = W.lstm_init(rl, H, E, V, np.float32)
keys = list(wdl)
ids = [rl.integers(0, V, Bl) for _ in range(T)]
tgs = []
for _ in range(T):
    oh = np.zeros((V, Bl), np.float32); oh[rl.integers(0, V, Bl), np.arange(Bl)] = 1
    tgs.append(np.asfortranarray(oh))
lo2, hi2 = dp.shard_columns(Bl, rank, world)
pl = [ag.Param(ag.DArray.from_numpy(wdl[k])) for k in keys]
ids_r = [ag.IArray(i[lo2:hi2]) for i in ids]
tg_r = [ag.DArray.from_numpy(np.asfortranarray(t_[:, lo2:hi2])) for t_ in tgs]
h0 = ag.zeros((H, hi2 - lo2), np.float32)
tl = ag.diff(lambda: W.lstm_loss(ag, dict(zip(keys, pl)), ids_r, tg_r, h0, h0, global_batch=Bl))
dp.update(-0.05, [ag.grad(tl, p) for p in pl], pl, gclip=0.2)
pol = [agref.Param(np.asfortranarray(wdl[k].astype(np.float64))) for k in keys]
h064 = np.zeros((H, Bl))
tol_ = agref.diff(lambda: W.lstm_loss(agref, dict(zip(keys, pol)), ids, [t_.astype(np.float64) for t_ in tgs], h064, h064))
gl = [agref.full(agref.grad(tol_, p)) for p in pol]
gn = np.sqrt(sum(float((g ** 2).sum()) for g in gl))
sc = min(1.0, 0.2 / gn)
worst = 0.0
for k, p, g in zip(keys, pl, gl):
    want = wdl[k].astype(np.float64) - 0.05 * sc * g
    worst = max(worst, float(np.max(np.abs(p.value.to_numpy() - want)) / max(np.max(np.abs(want)), 1e-30)))
print("rank %d LSTM clipped data-parallel update: gnorm %.4f scale %.4f worst rel err %.3e" % (rank, gn, sc, worst), flush=True)
ok &= bool(worst <= 2e-5)
# every rank must hold bit-identical reduced gradients
digest = [float(np.sum(a.astype(np.float64))) for a in g_p2p]
alld = allgather(digest)
ok &= all(d == alld[0] for d in alld)
# latency of the two collectives on the MNIST-size bucket (50 890 floats) and an LSTM-size one (1.4 M floats)
for nfl in (50890, 1400000):
    bkt = ag.zeros((nfl,), np.float32)
    import ctypes as C
    def run(fn, reps=50):
        for _ in range(5): fn()
        ag.sync(); dist.barrier()
        a, b = ag.Event(), ag.Event()
        a.record()
        for _ in range(reps): fn()
        b.record()
        return b.elapsed_ms_since(a) / reps * 1e3
    lib = ag._lib.lib
    t_p2p = run(lambda: ag._lib.check(lib.ag_p2p_allreduce_sum(0, C.c_void_p(bkt.ptr), nfl)))
    t_nccl = run(lambda: ag._lib.check(lib.ag_allreduce_sum(0, C.c_void_p(bkt.ptr), nfl)))
    if rank == 0:
        print("allreduce %d floats: p2p %.1f us, nccl %.1f us" % (nfl, t_p2p, t_nccl), flush=True)
print("rank %d/%d p2p_cap=%d: %s" % (rank, world, dp.p2p_cap, "OK" if ok else "MISMATCH"), flush=True)
dist.barrier()
dp.close()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
