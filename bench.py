#!/usr/bin/env python
"""bench.py -- BASELINE.json metric: `@diff`+`grad` steps/sec and backward HBM GB/s on the MNIST-shape MLP
(examples/mnist.jl, 784-64-10, Float32) at batch 65536 per GPU, data-parallel over N B200s.

One "step" = forward recording + backward sweep (+ NCCL allreduce of the Param gradients when N > 1) +
axpy! update of all four Params, on one 65536-column batch shard per rank (weak scaling).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

`--impl reference` times the CPU restatement of the reference (oracle/agref.py, numpy + OpenBLAS on
all host threads; Julia is not installed in this image) on the same config.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

if "reference" in sys.argv[1:]:
    # The CPU arm uses every host core it is allowed to: torch.distributed.run exports OMP_NUM_THREADS=1 to its
    # workers, which silently made the N > 1 reference single-threaded in round 1.  Must happen before numpy
    # (OpenBLAS) is imported.
    _ncores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    for _k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_k] = str(_ncores)

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH = 65536
LR = 0.1
METRIC = "@diff+grad steps/sec (MNIST-shape MLP 784-64-10, Float32, batch 65536 per GPU)"
UNIT = "steps/s"


def workload_name(B):
    """config.workload: the same string in both arms (the driver compares them)."""
    return ("examples/mnist.jl MLP 784-64-10 softmax (configs[1]), batch %d per GPU, Float32, "
            "@diff + grad + allreduce + axpy!; value counts one per-rank step per rank" % B)


def load_workloads():
    """autograd.jl_b200/workloads.py by file path: pure numpy model definitions.  The reference arm must not import
    the product package (that would map libagb200.so into the CPU arm's process)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("agb200_workloads",
                                                  os.path.join(ROOT, "autograd.jl_b200", "workloads.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([int(p.get("num_threads", 1)) for p in threadpool_info()] or [1])
    except Exception:
        return int(os.environ.get("OMP_NUM_THREADS", "0") or 0) or None


def reference_runtime():
    """prefer the real reference when it can run.  It is a Julia package: needs `julia` on PATH or
    under baseline/_ref/.  Neither exists in this image, so the numpy restatement (oracle/agref.py) is timed."""
    import shutil
    j = shutil.which("julia")
    if j is None:
        for cand in (os.path.join(ROOT, "baseline", "_ref", "bin", "julia"), os.path.join(ROOT, "baseline", "_ref", "julia")):
            if os.path.exists(cand):
                j = cand
    return {"julia": j, "baseline_ref": os.path.isdir(os.path.join(ROOT, "baseline", "_ref")),
            "used": "oracle/agref.py (numpy restatement; julia not found)" if j is None else
                    "oracle/agref.py (julia found at %s but AutoGrad.jl's dependencies cannot be resolved offline)" % j}


def algo_bytes(B, s=4):
    """Algorithmic bytes of the unfused tape: (forward, backward)."""
    N1, N2, NX = 64 * B, 10 * B, 784 * B
    return s * (NX + 6 * N1 + 12 * N2 + 4 * B), s * (15 * N2 + 5 * B + 7 * N1 + NX)


class Clocks:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0=None, t1=None, t_load=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for ts, r in self.rows if (t0 is None or ts >= t0 - 0.02) and (t1 is None or ts <= t1 + 0.02)]
        window = "timed region"
        if len(rows) < 3:       # the timed region is only a few ms: fall back to every sample taken under load
            rows = [r for ts, r in self.rows if t_load is None or ts >= t_load]
            window = "pre-load + timed region, all under load (timed region shorter than 3 sampling periods)"
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "window": window, "reasons": sorted(reasons)}


def bind_to_gpu_numa(local_rank, world, dev=None, stride=1):
    """Bind this rank's host threads to the CPUs local to its GPU (NVML's affinity mask = the GPU's NUMA node),
    split evenly among the ranks that share that mask, before any pinned staging buffer is allocated: the pages
    are then first-touched on the GPU's own node and the ranks' host loops do not migrate.  (Round 1: 8 ranks
    pulling 208 MB/step each through unpinned processes collapsed the end-to-end rate to 44 %.)"""
    info = {"bound": False}
    try:
        import pynvml
        pynvml.nvmlInit()
        dev = local_rank if dev is None else dev
        h = pynvml.nvmlDeviceGetHandleByIndex(dev)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [i for i in range(ncpu) if (words[i // 64] >> (i % 64)) & 1]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            peers = []          # local ranks sharing this mask
            for r in range(world):
                try:
                    hw = pynvml.nvmlDeviceGetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(r * stride), (ncpu + 63) // 64)
                    if list(hw) == list(words):
                        peers.append(r)
                except Exception:
                    pass
            k = peers.index(local_rank) if local_rank in peers else 0
            share = max(1, len(allowed) // max(len(peers), 1))
            mine = allowed[k * share:(k + 1) * share] or allowed
            os.sched_setaffinity(0, mine)
            info = {"bound": True, "gpu_local_cpus": len(allowed), "ranks_on_node": len(peers), "cpus": len(mine)}
        pynvml.nvmlShutdown()
    except Exception as exc:        # noqa: BLE001
        info["error"] = str(exc)[:120]
    return info


def cpu_reference_run(steps, warmup, budget_s=20.0):
    """The CPU restatement of the reference on the host cores: same tape, same temporaries."""
    from oracle import agref
    W = load_workloads()                   # workload definitions only (pure numpy; the product package is not imported)
    rng = np.random.default_rng(1)
    w, _, _ = W.mnist_init(rng, 1)
    x, y = W.mnist_from_ubyte(*W.mnist_ubyte_batch(np.random.default_rng(1000), BATCH))   # the arrays of mnist.jl:80-81
    params = [agref.Param(np.asfortranarray(a)) for a in w]

    def step():
        tape = agref.diff(lambda: W.mnist_loss(agref, params, x, y))
        for p in params:
            agref.axpy(-LR, agref.grad(tape, p), p)
        return float(agref.value(tape))
    for _ in range(warmup):
        step()
    times, t_start = [], time.perf_counter()
    for i in range(steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start > budget_s and len(times) >= 3:
            break
    mean = sum(times) / len(times)
    nthr = blas_threads()
    return {"value": 1.0 / mean, "unit": UNIT, "cores": nthr or os.cpu_count(), "kind": "port",
            "host_cpus": os.cpu_count(), "blas_threads": nthr,
            "sample": "%d full steps (forward+backward+axpy!) at batch %d Float32, numpy/OpenBLAS, %s BLAS threads"
                      % (len(times), BATCH, nthr if nthr else "?"), "ms_per_step": mean * 1e3}, len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    W_ = max(args.warmup, 3)               # the same warm-up rule as the device arm
    cb, done = cpu_reference_run(args.steps, W_, budget_s=150.0)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": done, "warmup": W_, "ms_per_step": cb["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(BATCH), "global_batch": BATCH, "parallelism": "cpu",
                       "arm": "CPU restatement of the reference (oracle/agref.py, numpy/OpenBLAS) on rank 0's host cores",
                       "reference_runtime": reference_runtime()},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="agb200")
    ap.add_argument("--batch", type=int, default=BATCH, help=argparse.SUPPRESS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the other BASELINE configs (housing, sweep, charlm, rnn)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    W_ = max(args.warmup, 3)
    K = args.steps
    B = args.batch

    # Which GPU this rank drives.  With more GPUs visible than ranks (the driver's scaling run shows all eight at every
    # N) the ranks are spread evenly over the box (N = 4: GPUs 0, 2, 4, 6) instead of packed onto the first N: the
    # host-to-device paths are shared between neighbouring GPUs on this box (measured: 4 ranks on GPUs 0-3 get 28 GB/s
    # each, 2 ranks 54 GB/s each), and NVSwitch makes every pair equally close for the allreduce.  AGB_SPREAD_DEVICES=0
    # packs them (device = LOCAL_RANK).
    dev, ndev, stride = local, None, 1
    if world > 1 and os.environ.get("AGB_SPREAD_DEVICES", "1") != "0":
        try:
            import torch
            ndev = torch.cuda.device_count()
            if ndev > world and ndev % world == 0:
                stride = ndev // world
                dev = local * stride
        except Exception:      # noqa: BLE001
            pass
    affinity = bind_to_gpu_numa(local, world, dev, stride)     # before any pinned allocation (first touch = this rank's node)
    affinity["device"] = dev
    affinity["visible_devices"] = ndev
    import __graft_entry__ as ge
    ag = ge.load_package()
    ag.init(dev)
    Wl = ag.workloads
    clocks = Clocks(dev) if rank == 0 else None          # started early: nvidia-smi needs about a second before its first sample

    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("gloo", rank=rank, world_size=world)

        def bcast(payload):
            obj = [payload]
            dist.broadcast_object_list(obj, src=0)
            return obj[0]
        def allgather(payload):
            out = [None] * world
            dist.all_gather_object(out, payload)
            return out
        use_p2p = os.environ.get("AGB_ALLREDUCE", "p2p") == "p2p"      # AGB_ALLREDUCE=nccl forces the NCCL path
        # NCCL prints its version banner on the C stdout while the communicator is created: point fd 1 at stderr for
        # that call (and flush the C buffer before switching back), so that stdout carries the one JSON line only
        import ctypes as _ct
        sys.stdout.flush()
        _saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dp = ag.DataParallel(rank, world, bcast, allgather if use_p2p else None)
        finally:
            try:
                _ct.CDLL(None).fflush(None)
            except Exception:      # noqa: BLE001
                pass
            os.dup2(_saved, 1)
            os.close(_saved)
        dist.barrier()
    else:
        dp = ag.DataParallel(0, 1)

    def barrier():
        ag.sync()
        if dist is not None:
            dist.barrier()

    # identical Params on every rank, per-rank batch shard
    w_np, _, _ = Wl.mnist_init(np.random.default_rng(1), 1)
    # the batch in the dataset's byte format and the Float32 arrays the reference builds from it (mnist.jl:80-81)
    px_np, lb_np = Wl.mnist_ubyte_batch(np.random.default_rng(1000 + rank), B)
    x_np, y_np = Wl.mnist_from_ubyte(px_np, lb_np)
    params = [ag.Param(ag.DArray.from_numpy(a)) for a in w_np]
    batch = ag.UByteBatch(x_np.shape[0], B)
    xd, yd = batch.x, batch.y
    xh, yh = ag.pinned_empty(x_np.shape, np.float32), ag.pinned_empty(y_np.shape, np.float32)
    xh[...] = x_np
    yh[...] = y_np
    pxh, lbh = ag.pinned_empty(px_np.shape, np.uint8), ag.pinned_empty(lb_np.shape, np.uint8)
    pxh[...] = px_np
    lbh[...] = lb_np
    loss_h = ag.pinned_empty((1,), np.float32)
    batch.load(pxh, lbh)
    ag.sync()
    if rank == 0:       # the device conversions rebuild the reference's arrays bit for bit
        assert np.array_equal(xd.to_numpy(), x_np) and np.array_equal(yd.to_numpy(), y_np), "ubyte batch mismatch"
    gB = B * world
    state = {}

    def fwd_bwd():
        t = ag.diff(lambda: Wl.mnist_loss(ag, params, xd, yd, global_batch=gB))
        state["loss"] = ag.value(t)
        state["grads"] = [ag.grad(t, p) for p in params]
        return t

    def exchange():
        pass        # folded into dp.update: allreduce (+ fused axpy!) is one launch

    def update():
        state["grads"] = dp.update(-LR, state["grads"], params)

    def eager_step():       # = fwd_bwd + update; shared with tools/gen_fused_programs.py
        Wl.mnist_train_step(ag, dp, params, xd, yd, LR, state, global_batch=gB)

    # eager warm-up populates the pool / workspace, then capture the step
    for _ in range(2):
        eager_step()
    ag.sync()
    graphs, replay, fused_in_step = [], None, None
    try:
        # the whole step (tape forward + backward sweep + NCCL allreduce + axpy!) as ONE CUDA graph
        fz0 = sys.modules["agb200.lazy"].kernel_stats()
        with ag.StepGraph() as g_all:
            eager_step()
        fz1 = sys.modules["agb200.lazy"].kernel_stats()
        fused_in_step = {"specialised": fz1[0] - fz0[0], "interpreted": fz1[1] - fz0[1]}
        graphs, replay = [g_all], "CUDA graph (single, collective captured)" if world > 1 else "CUDA graph"

        def step():
            g_all.launch()
    except Exception as exc:          # NCCL refused capture: graph | eager allreduce | graph
        if world == 1:
            raise
        sys.stderr.write("rank %d: single-graph capture failed (%s); falling back to split graphs\n" % (rank, exc))
        ag.sync()
        with ag.StepGraph() as g_a:
            fwd_bwd()
        pre = state["grads"]
        exchange()                       # eager once: fixes the bucket views
        with ag.StepGraph() as g_b:
            update()
        graphs, replay = [g_a, g_b], "CUDA graphs around an eager NCCL allreduce"

        def step():
            g_a.launch()
            state["grads"] = pre
            exchange()
            g_b.launch()

    for _ in range(W_):
        step()
    barrier()

    # ---- timed region: device-resident inputs (x = 205 MB > 126 MB L2, so no L2 flush needed)
    t_load0 = time.time()
    for _ in range(2000):         # ~0.4 s under load before the timed region: nvidia-smi samples every 20 ms
        step()
    barrier()
    l0 = ag.launch_count()
    e0, e1 = ag.Event(), ag.Event()
    barrier()
    t_wall0 = time.time()
    e0.record()
    for _ in range(K):
        step()
    e1.record()
    ms = e1.elapsed_ms_since(e0)
    barrier()
    t_wall1 = time.time()
    launches = ag.launch_count() - l0
    clk = clocks.stop(t_wall0, t_wall1, t_load0) if clocks is not None else None

    # ---- forward / backward split (two graphs, events between them), N-independent
    #      (recorded on a tape, so the forward writes every value the backward sweep will read)
    from agb200 import core as _core
    _lz = sys.modules["agb200.lazy"]
    with ag.StepGraph() as g_f:
        tape_f = _core.Tape()
        _core._tapes.append(tape_f)
        try:
            loss_t = Wl.mnist_loss(ag, params, xd, yd, global_batch=gB)
        finally:
            _core._tapes.pop()
        # what the full step does at the end of its sweep (core.differentiate): values that only the tape holds
        # softly (the product under `w*x .+ b`, which no rule reads) are dropped, not computed by the closing flush --
        # otherwise the forward graph would run w1*x a second time and the split below would charge it to "forward"
        soft = [n.Value.value for n in reversed(tape_f.list) if isinstance(n.Value, _core.Result) and n.Value._soft
                and getattr(n.Value.value, "kind", None) is not None]
        if soft:
            _lz.discard_soft(soft)
        del soft
    ef0, ef1 = ag.Event(), ag.Event()
    ef0.record()
    for _ in range(K):
        g_f.launch()
    ef1.record()
    fwd_ms = ef1.elapsed_ms_since(ef0) / K
    g_f.destroy()

    # ---- end to end through the public API with HOST buffers: h2d of the batch + step + d2h of the loss
    C_ = __import__("ctypes")

    # Every step copies ITS batch from pinned host memory (h2d), runs, and reads its loss back (d2h + sync).  The copy
    # of step i+1 is issued on the copy stream before step i is launched, so it overlaps the compute (input prefetch).
    staged = ag.StagedInput(xd, yd)

    def e2e_step():
        staged.commit()                                        # batch i: staging -> xd, yd (device to device)
        staged.prefetch(xh.host_ptr, yh.host_ptr)              # batch i+1 starts crossing PCIe now
        step()
        ag._lib.check(ag._lib.lib.ag_copy_d2h(C_.c_void_p(loss_h.host_ptr), C_.c_void_p(state["loss"].ptr), 4))

    def e2e_ubyte_step():       # same arrays rebuilt on the device from the dataset's bytes
        batch.convert()                                        # batch i: staged bytes -> x, y
        batch.prefetch(pxh, lbh)                               # batch i+1
        step()
        ag._lib.check(ag._lib.lib.ag_copy_d2h(C_.c_void_p(loss_h.host_ptr), C_.c_void_p(state["loss"].ptr), 4))

    def time_e2e(fn, first):
        first()                                                # the first batch's copy (every later one is prefetched)
        for _ in range(2):
            fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(Ke):
            fn()
        ag._lib.check(ag._lib.lib.ag_prefetch_wait())          # the copy issued by the last step also completes in here
        ag.sync()
        dt = (time.perf_counter() - t0) / Ke
        barrier()
        return dt
    Ke = max(3, min(K, 20))
    e2e_s = time_e2e(e2e_step, lambda: staged.prefetch(xh.host_ptr, yh.host_ptr))
    e2e_ub_s = time_e2e(e2e_ubyte_step, lambda: batch.prefetch(pxh, lbh))
    del staged

    # ---- dominant kernel, timed alone on the library stream (inputs > L2)
    cand = {}
    a1 = ag.matmul(params[0].value, xd)
    a1.ptr                                   # products are deferred (lazy.py): touching the pointer evaluates them
    dy1 = ag.DArray.empty(a1.shape, np.float32)
    ag._lib.check(ag._lib.lib.ag_fill(C_.c_void_p(dy1.ptr), dy1.size, 0, 0.001))

    def time_op(fn, reps=max(5, min(K, 30))):
        """Average launch duration (CUDA events on the library stream around `reps` back-to-back launches), median of
        three such batches: one batch occasionally catches a hiccup of the box (seen: 73 us instead of 60)."""
        for _ in range(3):
            fn()
        out = []
        for _ in range(3):
            a, b = ag.Event(), ag.Event()
            a.record()
            for _ in range(reps):
                fn()
            b.record()
            out.append(b.elapsed_ms_since(a) / reps)
        return sorted(out)[1]
    s = 4
    cand["gemm_fwd w1*x (64x784 * 784xB, NN)"] = (time_op(lambda: ag.matmul(params[0].value, xd).ptr),
                                                   s * (64 * 784 + 784 * B + 64 * B), 2.0 * 64 * 784 * B)
    eng_f = ag.gemm_engine()
    cand["gemm_dw1 dy*x' (64xB * Bx784, NT)"] = (time_op(lambda: ag.matmul(dy1, ag.adjoint(xd)).ptr),
                                                  s * (64 * B + 784 * B + 64 * 784), 2.0 * 64 * 784 * B)
    eng_b = ag.gemm_engine()
    # the form the step actually runs: max.(0, w1*x .+ b1) as ONE launch (bias and relu in the product's epilogue,
    # writes a2 and a3, never writes w1*x); algorithmic bytes of the three tape nodes it replaces
    def fused_layer1():
        z = ag.add(ag.matmul(params[0].value, xd), params[1].value)       # held, as the tape's Result holds it:
        a = ag.relu(z)                                                    # both a2 (z) and a3 (a) are written
        a.ptr
        return z, a
    t_fused = time_op(fused_layer1)
    fused_fwd = {"ms": t_fused, "engine": ag.gemm_engine(),
                 "algorithmic_bytes": s * (64 * 784 + 784 * B + 64 * B) + s * (3 * 64 * B) + s * (3 * 64 * B)}
    fused_fwd["GBps"] = fused_fwd["algorithmic_bytes"] / (t_fused * 1e-3) / 1e9
    # The launches the step actually runs: the fused layer-1 forward (reads x, w1, b1 once; writes a2 = w1*x .+ b1 and
    # a3 = max.(0, a2) once -- its own algorithmic bytes, not those of the three tape nodes it replaces) and dy*x'
    # with its split-K combine.  The plain w1*x above is kept for comparison with round 1 but no longer runs in the step.
    k_fused = "fused w1*x .+ b1 -> max.(0,.) (64x784 * 784xB, NN; the launch the step runs)"
    cand[k_fused] = (t_fused, s * (64 * 784 + 784 * B + 64 + 2 * 64 * B), 2.0 * 64 * 784 * B)
    in_step = [k_fused, "gemm_dw1 dy*x' (64xB * Bx784, NT)"]
    top = max(in_step, key=lambda k: cand[k][0])
    t_ms, bytes_, flops = cand[top]
    # the two elementwise launches of the step that move > 32 MB, timed alone the same way:
    # bias + relu (reads a1, writes a2 and a3) and the relu gradient with its fused bias-gradient row sums
    Dm = sys.modules["agb200.darray"]
    b1v, N1 = params[1].value, 64 * B

    def ew_fwd():
        a2 = ag.add(a1, b1v)
        a3 = ag.relu(a2)
        a3.ptr
        return a2, a3
    ew_fwd()
    a2k = ag.add(a1, b1v)
    a3k = ag.relu(a2k)
    a3k.ptr

    def ew_bwd():
        g = Dm.binary_bwd("max", 2, dy1, 0.0, a2k, a3k)
        db = ag.sum_(g, dims=1)
        db.ptr
        return g, db
    flush_buf = ag.DArray.empty((1 << 26,), np.float32)       # 256 MB > L2: written between timed launches

    def flush_l2():
        ag._lib.check(ag._lib.lib.ag_fill(C_.c_void_p(flush_buf.ptr), flush_buf.size, 0, 0.0))

    def time_op_flushed(fn, reps=20):
        """These operands (17 MB each) would stay in the 126 MB L2 from launch to launch, and a host-enqueued launch
        between two events would be timed with the host's enqueue latency: replay a graph of [L2 flush, launch] and
        subtract the replay time of a graph of [L2 flush] alone."""
        fn()
        keep = []
        with ag.StepGraph() as g1:
            flush_l2()
            keep.append(fn())
        with ag.StepGraph() as g0:
            flush_l2()
        out = []
        for g in (g1, g0):
            for _ in range(3):
                g.launch()
            ea, eb = ag.Event(), ag.Event()
            ea.record()
            for _ in range(reps):
                g.launch()
            eb.record()
            out.append(eb.elapsed_ms_since(ea) / reps)
            g.destroy()
        return max(out[0] - out[1], 1e-6)
    elementwise = {
        "bias+relu fwd as its own launch (64xB): a1 -> a2, a3 [no longer in the step: runs in the w1*x epilogue]":
            (time_op_flushed(ew_fwd), s * 3 * N1),
        "fused relu gradient + bias-gradient row sums (64xB)": (time_op_flushed(ew_bwd), s * (4 * N1 + 64)),
    }
    del flush_buf
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    traffic = None          # dram__bytes_read+write of this kernel, one `ncu --set full` launch (profiles/)
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
        traffic = tj["kernels"]["dw1" if "dw1" in top else "fwd_fused"]["traffic"]
    except Exception:
        pass
    roofline = {"kernel": top, "engine": eng_b if "dw1" in top else fused_fwd["engine"], "bound": "hbm",
                "achieved": bytes_ / (t_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": bytes_ / (t_ms * 1e-3) / 1e9 / hbm_peak, "traffic": traffic,
                "algorithmic_bytes": bytes_,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650",
                "ms_per_launch": t_ms, "tflops": flops / (t_ms * 1e-3) / 1e12,
                "all": {k: {"ms": v[0], "GBps": v[1] / (v[0] * 1e-3) / 1e9, "algorithmic_bytes": v[1],
                            "frac": v[1] / (v[0] * 1e-3) / 1e9 / hbm_peak, "in_step": k in in_step} for k, v in cand.items()},
                "fused_forward_layer1": dict(fused_fwd, frac=fused_fwd["GBps"] / hbm_peak,
                                             note="the same launch counted in the bytes of the three unfused tape nodes it replaces (SURVEY.md 8d per-node figures)"),
                "elementwise": {k: {"ms": v[0], "GBps": v[1] / (v[0] * 1e-3) / 1e9,
                                    "frac": v[1] / (v[0] * 1e-3) / 1e9 / hbm_peak, "algorithmic_bytes": v[1],
                                    "l2": "flushed (256 MB written) before every launch; graph replay, flush-only replay subtracted"}
                                for k, v in elementwise.items()}}

    # ---- numeric check of the N-rank step, and the other BASELINE configs (bench_configs.py)
    import bench_configs as BC

    def allmax(v):
        if dist is None:
            return v
        import torch
        tt_ = torch.tensor([v], dtype=torch.float64)
        dist.all_reduce(tt_, op=dist.ReduceOp.MAX)
        return float(tt_[0])

    def allgather_obj(o):
        out = [None] * world
        dist.all_gather_object(out, o)
        return out
    dpc = BC.dp_check(ag, Wl, dp, rank, world, allgather_obj if dist is not None else None)
    secondary = {}
    if not args.no_secondary:
        for name, fn in (("housing", lambda: BC.housing(ag, Wl, barrier, allmax, world)),
                         ("sweep", lambda: BC.sweep(ag, Wl, barrier, allmax, world, hbm_peak)),
                         ("charlm", lambda: BC.lstm(ag, Wl, dp, barrier, allmax, world, peaks)),
                         ("rnn_grad_of_grad", lambda: BC.rnn_grad_of_grad(ag, Wl, barrier, allmax, world, peaks))):
            try:
                secondary[name] = fn()
            except Exception as exc:        # noqa: BLE001  (a failing secondary config must not hide the headline)
                secondary[name] = {"error": "%s: %s" % (type(exc).__name__, str(exc)[:300])}
            ag.sync()

    # ---- max over ranks
    ms_all, e2e_all = ms, e2e_s
    if dist is not None:
        import torch
        tt = torch.tensor([ms, e2e_s, fwd_ms, e2e_ub_s], dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_all, e2e_all, fwd_ms, e2e_ub_s = tt.tolist()
    ms_per_step = ms_all / K
    fb, bb = algo_bytes(B)
    bwd_ms = max(ms_per_step - fwd_ms, 1e-9)

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_base, _ = cpu_reference_run(1000, 1, budget_s=15.0)

    if rank == 0:
        line = {
            "metric": METRIC, "value": world * 1e3 / ms_per_step, "unit": UNIT, "n_gpus": world, "steps": K,
            "warmup": W_, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(B),
                       "global_batch": gB, "parallelism": "dp%d" % world, "replay": replay,
                       "allreduce": ("one-shot peer-memory kernel over NVLink, flag-in-data, fused axpy! (ag_p2p_allreduce_multi)" if dp.p2p_cap else
                                     "NCCL") if world > 1 else None,
                       "fusion": ("elementwise chains of the tape deferred and evaluated by one ag_ewise_fused launch each"
                                  if sys.modules["agb200.lazy"].enabled[0] else "off (AGB_FUSION=0)"),
                       "launches_per_step": launches / max(K, 1), "fused_launches_per_step": fused_in_step,
                       "l2": "inputs (x = %.0f MB per rank) larger than the 126 MB L2; no flush" % (x_np.nbytes / 1e6)},
            "forward_ms": fwd_ms, "backward_update_ms": bwd_ms,
            "backward_hbm_gbs": bb / (bwd_ms * 1e-3) / 1e9,
            "step_hbm_gbs": (fb + bb) / (ms_per_step * 1e-3) / 1e9,
            "algorithmic_bytes": {"forward": fb, "backward": bb},
            "clocks": clk,
            "e2e": {"value": world / e2e_ub_s, "unit": UNIT, "h2d_bytes_per_step": int(batch.h2d_bytes),
                    "d2h_bytes_per_step": 4, "ms_per_step": e2e_ub_s * 1e3,
                    "h2d_GBps_per_rank": batch.h2d_bytes / e2e_ub_s / 1e9,
                    "input": "the batch in the reference loader's own format (examples/mnist.jl:77-96: idx-ubyte, 1 B per "
                             "pixel / label) in pinned host memory; x ./ 255f0 and the one-hot y of :80-81 are rebuilt on "
                             "the device, bit-identical to the host arrays (asserted); the copy of batch i+1 overlaps "
                             "step i (second stream); one h2d + one d2h (loss) per step",
                    "host_binding": affinity},
            "e2e_f32": {"value": world / e2e_all, "unit": UNIT, "h2d_bytes_per_step": int(xh.nbytes + yh.nbytes),
                        "d2h_bytes_per_step": 4, "ms_per_step": e2e_all * 1e3,
                        "h2d_GBps_per_rank": (xh.nbytes + yh.nbytes) / e2e_all / 1e9,
                        "input": "the host arrays the reference builds from those bytes: x (784,B) and one-hot y (10,B) "
                                 "Float32, pinned (4x the bytes; PCIe-bound)"},
            "gpu_launches": int(launches),
            "dp_check": dpc,
            "secondary": secondary,
            "roofline": roofline,
            "cpu_baseline": cpu_base,
        }
        print(json.dumps(line))
    for g in graphs:
        g.destroy()
    if dist is not None:
        dist.barrier()
        dp.close()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
