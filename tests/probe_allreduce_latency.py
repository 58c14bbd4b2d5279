"""Latency of the gradient-bucket collectives (run under torchrun): NCCL vs the peer-memory kernel."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch.distributed as dist
import __graft_entry__ as ge
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dist.init_process_group("gloo", rank=rank, world_size=world)
ag = ge.load_package(); ag.init(local)
def bcast(p):
    o = [p]; dist.broadcast_object_list(o, src=0); return o[0]
def allgather(p):
    o = [None] * world; dist.all_gather_object(o, p); return o
dp = ag.DataParallel(rank, world, bcast, allgather)
dist.barrier()
lib = ag._lib.lib
for nfl in (50890, 1400000):
    bkt = ag.zeros((nfl,), np.float32)
    def run(fn, reps=100):
        for _ in range(10): fn()
        ag.sync(); dist.barrier()
        a, b = ag.Event(), ag.Event()
        a.record()
        for _ in range(reps): fn()
        b.record()
        return b.elapsed_ms_since(a) / reps * 1e3
    t_nccl = run(lambda: ag._lib.check(lib.ag_allreduce_sum(0, C.c_void_p(bkt.ptr), nfl)))
    t_p2p = run(lambda: ag._lib.check(lib.ag_p2p_allreduce_sum(0, C.c_void_p(bkt.ptr), nfl)))
    if rank == 0:
        print("world %d, %d floats [NCCL_PROTO=%s NCCL_ALGO=%s]: nccl %.1f us, p2p %.1f us" % (
            world, nfl, os.environ.get("NCCL_PROTO"), os.environ.get("NCCL_ALGO"), t_nccl, t_p2p), flush=True)
dist.barrier(); dp.close(); dist.destroy_process_group()
