// runtime.cu -- context, stream-ordered caching pool, copies, events, CUDA-graph capture.
// No reference counterpart (the reference is CPU-only Julia); the hooks it serves are
// `@gs` (src/AutoGrad.jl:11) -> ag_sync and gcnode (src/core.jl:168,235-237) -> ag_free.
#include "common.cuh"
#include "gemm_tc.cuh"

#include <nvtx3/nvToolsExt.h>

#include <map>
#include <mutex>
#include <unordered_map>
#include <vector>

namespace agb {

static Context g_ctx;
Context& ctx() { return g_ctx; }

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

ag_status cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  set_error("CUDA error %d (%s) at %s:%d in %s", (int)e, cudaGetErrorString(e), file, line, what);
  return e == cudaErrorMemoryAllocation ? AG_ENOMEM : AG_ECUDA;
}

// ---- pool ---------------------------------------------------------------------------------------
struct Block {
  size_t bytes;
  uint64_t graph;  // 0 = not owned by a graph
};

struct GraphRec {
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  int64_t kernels = 0;
  std::vector<void*> blocks;                           // every block handed out during capture
  std::map<size_t, std::vector<void*>> free_in_capture; // reusable only while capturing
};

static std::mutex g_mu;  // ag_free may be called from a finalizer thread
static std::map<size_t, std::vector<void*>> g_free;
static std::unordered_map<void*, Block> g_live;
static std::unordered_map<void*, size_t> g_all;  // every cudaMalloc'ed block -> size
static size_t g_pool_bytes = 0, g_live_bytes = 0;
static std::unordered_map<uint64_t, GraphRec*> g_graphs;
static uint64_t g_next_graph = 1;
static uint64_t g_capturing_graph = 0;
static int64_t g_capture_launch0 = 0;
static std::vector<void*> g_parked_ws;     // workspaces outgrown while graphs that point into them exist

static size_t round_size(size_t b) {
  if (b == 0) b = 1;
  if (b < (1u << 20)) return (b + 511) & ~size_t(511);
  return (b + (1u << 20) - 1) & ~size_t((1u << 20) - 1);
}

static void trim_locked() {
  for (auto& kv : g_free) {
    for (void* p : kv.second) {
      cudaFree(p);
      g_pool_bytes -= kv.first;
      g_all.erase(p);
    }
  }
  g_free.clear();
}

static ag_status pool_alloc(size_t bytes, void** out) {
  size_t sz = round_size(bytes);
  std::lock_guard<std::mutex> lk(g_mu);
  void* p = nullptr;
  GraphRec* gr = g_capturing_graph ? g_graphs[g_capturing_graph] : nullptr;
  if (gr) {
    auto it = gr->free_in_capture.find(sz);
    if (it != gr->free_in_capture.end() && !it->second.empty()) {
      p = it->second.back();
      it->second.pop_back();
      g_live[p] = Block{sz, g_capturing_graph};
      g_live_bytes += sz;
      *out = p;
      return AG_OK;
    }
  }
  auto it = g_free.find(sz);
  if (it != g_free.end() && !it->second.empty()) {
    p = it->second.back();
    it->second.pop_back();
  } else {
    cudaError_t e = cudaMalloc(&p, sz);
    if (e != cudaSuccess) {
      cudaGetLastError();
      // free cached blocks and retry -- but not while graphs exist: a block the host has freed may still be read by
      // a captured kernel that did not own it (an input of the captured step)
      if (g_graphs.empty()) trim_locked();
      e = cudaMalloc(&p, sz);
      if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("out of device memory allocating %zu bytes (pool holds %zu, live %zu)", sz,
                  g_pool_bytes, g_live_bytes);
        return AG_ENOMEM;
      }
    }
    g_pool_bytes += sz;
    g_all[p] = sz;
  }
  g_live[p] = Block{sz, g_capturing_graph};
  g_live_bytes += sz;
  if (gr) gr->blocks.push_back(p);
  *out = p;
  return AG_OK;
}

static ag_status pool_free(void* p) {
  if (!p) return AG_OK;
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_live.find(p);
  if (it == g_live.end()) {
    set_error("ag_free: pointer %p was not allocated by ag_alloc (or double free)", p);
    return AG_EINVAL;
  }
  Block b = it->second;
  g_live.erase(it);
  g_live_bytes -= b.bytes;
  if (b.graph) {
    auto git = g_graphs.find(b.graph);
    if (git != g_graphs.end()) {
      // graph-owned: reusable inside the same capture, otherwise parked until ag_graph_destroy
      if (g_capturing_graph == b.graph) git->second->free_in_capture[b.bytes].push_back(p);
      return AG_OK;
    }
  }
  g_free[b.bytes].push_back(p);
  return AG_OK;
}

void* workspace(size_t bytes) {
  Context& c = ctx();
  if (bytes <= c.ws_bytes) return c.ws;
  if (c.capturing) {
    set_error("workspace of %zu bytes needed during graph capture (have %zu): run the step once "
              "eagerly before capturing", bytes, c.ws_bytes);
    return nullptr;
  }
  cudaStreamSynchronize(c.stream);
  // Instantiated graphs have the workspace pointer baked into their kernel arguments (split-K partials, reduction
  // partials, sort scratch): while any graph exists the old block is parked, not freed, until the last graph is
  // destroyed -- replaying a graph after an eager call grew the workspace must not touch freed memory.
  if (c.ws) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_graphs.empty()) cudaFree(c.ws);
    else g_parked_ws.push_back(c.ws);
  }
  size_t nb = bytes * 2;
  if (cudaMalloc(&c.ws, nb) != cudaSuccess) {
    cudaGetLastError();
    c.ws = nullptr;
    c.ws_bytes = 0;
    set_error("out of device memory growing the workspace to %zu bytes", nb);
    return nullptr;
  }
  c.ws_bytes = nb;
  return c.ws;
}

static int32_t* g_oob = nullptr;
static bool g_oob_armed = false;
int32_t* oob_flag() {
  if (!g_oob) {
    if (cudaMalloc((void**)&g_oob, 4) != cudaSuccess) {
      cudaGetLastError();
      set_error("out of device memory allocating the index-check flag");
      return nullptr;
    }
    cudaMemset(g_oob, 0, 4);
  }
  g_oob_armed = true;
  return g_oob;
}

// ---- kernels --------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) k_fill(T* __restrict__ dst, int64_t n, T v) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) dst[i] = v;
}

}  // namespace agb

using namespace agb;

extern "C" {

ag_status ag_init(int32_t device) {
  Context& c = ctx();
  if (c.ready) {
    AG_CHECK(c.device == device, AG_ESTATE, "ag_init: already bound to device %d", c.device);
    return AG_OK;
  }
  int ndev = 0;
  AG_CUDA(cudaGetDeviceCount(&ndev));
  AG_CHECK(device >= 0 && device < ndev, AG_EINVAL, "ag_init: device %d of %d", device, ndev);
  AG_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  AG_CUDA(cudaGetDeviceProperties(&prop, device));
  AG_CHECK(prop.major == 10, AG_EUNSUPPORTED,
           "libagb200 is built for sm_100a only; device %d is sm_%d%d", device, prop.major,
           prop.minor);
  c.device = device;
  c.num_sms = prop.multiProcessorCount;
  AG_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
  c.ws_bytes = (size_t)64 << 20;
  AG_CUDA(cudaMalloc(&c.ws, c.ws_bytes));
  AG_CUDA(cudaMemsetAsync(c.ws, 0, c.ws_bytes, c.stream));
  if (!oob_flag()) return AG_ENOMEM;   // allocate the sticky index-check flag outside of any capture
  g_oob_armed = false;
  if (!reduce_scratch_init()) return AG_ENOMEM;
  if (!colprog_init()) return AG_ENOMEM;
  if (!gemm_tc_init()) return AG_ENOMEM;
  c.launches = 0;
  c.capturing = false;
  c.ready = true;
  return AG_OK;
}

ag_status ag_shutdown(void) {
  Context& c = ctx();
  if (!c.ready) return AG_OK;
  cudaStreamSynchronize(c.stream);
  ag_comm_destroy();
  {
    std::lock_guard<std::mutex> lk(g_mu);
    for (auto& kv : g_graphs) {
      if (kv.second->exec) cudaGraphExecDestroy(kv.second->exec);
      if (kv.second->graph) cudaGraphDestroy(kv.second->graph);
      delete kv.second;
    }
    g_graphs.clear();
    for (auto& kv : g_all) cudaFree(kv.first);
    g_all.clear();
    g_free.clear();
    g_live.clear();
    g_pool_bytes = g_live_bytes = 0;
  }
  if (c.ws) cudaFree(c.ws);
  c.ws = nullptr;
  c.ws_bytes = 0;
  cudaStreamDestroy(c.stream);
  c.stream = nullptr;
  c.ready = false;
  return AG_OK;
}

ag_status ag_sync(void) {
  AG_REQUIRE_INIT();
  AG_CUDA(cudaStreamSynchronize(ctx().stream));
  if (!ctx().capturing) {
    const int who = comm_take_timeout();
    if (who) {
      set_error("peer-memory allreduce: the wait for rank %d expired (AGB_P2P_TIMEOUT_MS); the ranks' parameters "
                "are no longer in step", who - 1);
      return AG_ENCCL;
    }
  }
  if (!ctx().capturing) {
    char why[160];
    if (comm_nccl_async_error(why, sizeof(why))) {
      set_error("%s", why);
      return AG_ENCCL;
    }
  }
  if (g_oob_armed && g_oob && !ctx().capturing) {
    int32_t h = 0;
    AG_CUDA(cudaMemcpy(&h, g_oob, 4, cudaMemcpyDeviceToHost));
    g_oob_armed = false;
    if (h) {
      cudaMemset(g_oob, 0, 4);
      set_error("index out of range in an earlier gather/scatter (BoundsError)");
      return AG_EINVAL;
    }
  }
  return AG_OK;
}

ag_status ag_last_error(char* buf, int64_t buflen) {
  if (!buf || buflen <= 0) return AG_EINVAL;
  strncpy(buf, g_err, (size_t)buflen - 1);
  buf[buflen - 1] = 0;
  return AG_OK;
}

ag_status ag_device_info(int64_t* info) {
  AG_REQUIRE_INIT();
  AG_CHECK(info, AG_EINVAL, "ag_device_info: null");
  cudaDeviceProp prop;
  AG_CUDA(cudaGetDeviceProperties(&prop, ctx().device));
  size_t fr = 0, tot = 0;
  AG_CUDA(cudaMemGetInfo(&fr, &tot));
  info[0] = ctx().device;
  info[1] = prop.multiProcessorCount;
  info[2] = prop.major;
  info[3] = prop.minor;
  info[4] = (int64_t)tot;
  info[5] = (int64_t)fr;
  std::lock_guard<std::mutex> lk(g_mu);
  info[6] = (int64_t)g_pool_bytes;
  info[7] = (int64_t)g_live_bytes;
  return AG_OK;
}

ag_status ag_stream(uint64_t* s) {
  AG_REQUIRE_INIT();
  *s = (uint64_t)(uintptr_t)ctx().stream;
  return AG_OK;
}

ag_status ag_launch_count(int64_t* n) {
  AG_REQUIRE_INIT();
  *n = ctx().launches;
  return AG_OK;
}

ag_status ag_alloc(int64_t bytes, void** out) {
  AG_REQUIRE_INIT();
  AG_CHECK(bytes >= 0 && out, AG_EINVAL, "ag_alloc: bytes=%lld", (long long)bytes);
  return pool_alloc((size_t)bytes, out);
}

ag_status ag_free(void* p) {
  if (!ctx().ready) return AG_OK;  // finalizers may run after shutdown
  return pool_free(p);
}

ag_status ag_pool_trim(void) {
  AG_REQUIRE_INIT();
  AG_CUDA(cudaStreamSynchronize(ctx().stream));
  std::lock_guard<std::mutex> lk(g_mu);
  AG_CHECK(g_graphs.empty(), AG_ESTATE, "ag_pool_trim: %zu captured graphs may read cached blocks; destroy them first",
           g_graphs.size());
  trim_locked();
  return AG_OK;
}

ag_status ag_host_alloc(int64_t bytes, void** out) {
  AG_REQUIRE_INIT();
  AG_CHECK(bytes >= 0 && out, AG_EINVAL, "ag_host_alloc: bytes=%lld", (long long)bytes);
  AG_CUDA(cudaMallocHost(out, (size_t)(bytes ? bytes : 1)));
  return AG_OK;
}

ag_status ag_host_free(void* p) {
  if (!ctx().ready || !p) return AG_OK;
  AG_CUDA(cudaFreeHost(p));
  return AG_OK;
}

ag_status ag_copy_h2d(void* dst, const void* src, int64_t bytes) {
  AG_REQUIRE_INIT();
  if (bytes == 0) return AG_OK;
  AG_CHECK(dst && src && bytes > 0, AG_EINVAL, "ag_copy_h2d: bad arguments");
  AG_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyHostToDevice, ctx().stream));
  return AG_OK;
}

// ---- input prefetch on a second stream: the h2d copy of the NEXT batch overlaps the step on the current one ----
static cudaStream_t g_copy_stream = nullptr;
// Two slots (ag_prefetch_slot): with two staging buffers the copy of batch i+1 only waits for the consumer of batch i-1,
// so the copy engine never idles while batch i is being converted.
constexpr int kPrefetchSlots = 2;
static cudaEvent_t g_ev_ready[kPrefetchSlots] = {nullptr, nullptr}, g_ev_consumed[kPrefetchSlots] = {nullptr, nullptr};
static int g_prefetch_slot = 0;

static ag_status prefetch_init() {
  if (g_copy_stream) return AG_OK;
  AG_CUDA(cudaStreamCreateWithFlags(&g_copy_stream, cudaStreamNonBlocking));
  for (int k = 0; k < kPrefetchSlots; ++k) {
    AG_CUDA(cudaEventCreateWithFlags(&g_ev_ready[k], cudaEventDisableTiming));
    AG_CUDA(cudaEventCreateWithFlags(&g_ev_consumed[k], cudaEventDisableTiming));
    AG_CUDA(cudaEventRecord(g_ev_consumed[k], ctx().stream));
    AG_CUDA(cudaEventRecord(g_ev_ready[k], ctx().stream));
  }
  return AG_OK;
}

ag_status ag_prefetch_slot(int32_t slot) {
  AG_REQUIRE_INIT();
  AG_CHECK(slot >= 0 && slot < kPrefetchSlots, AG_EINVAL, "ag_prefetch_slot: slot %d of %d", slot, kPrefetchSlots);
  g_prefetch_slot = slot;
  return AG_OK;
}

ag_status ag_copy_h2d_prefetch(void* dst, const void* src, int64_t bytes) {
  AG_REQUIRE_INIT();
  if (bytes == 0) return AG_OK;
  AG_CHECK(dst && src && bytes > 0, AG_EINVAL, "ag_copy_h2d_prefetch: bad arguments");
  ag_status st = prefetch_init();
  if (st != AG_OK) return st;
  // the last reader of this slot's staging buffer is done
  AG_CUDA(cudaStreamWaitEvent(g_copy_stream, g_ev_consumed[g_prefetch_slot], 0));
  AG_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyHostToDevice, g_copy_stream));
  AG_CUDA(cudaEventRecord(g_ev_ready[g_prefetch_slot], g_copy_stream));
  return AG_OK;
}

ag_status ag_prefetch_wait(void) {
  AG_REQUIRE_INIT();
  ag_status st = prefetch_init();
  if (st != AG_OK) return st;
  AG_CUDA(cudaStreamWaitEvent(ctx().stream, g_ev_ready[g_prefetch_slot], 0));
  return AG_OK;
}

ag_status ag_prefetch_release(void) {
  AG_REQUIRE_INIT();
  ag_status st = prefetch_init();
  if (st != AG_OK) return st;
  AG_CUDA(cudaEventRecord(g_ev_consumed[g_prefetch_slot], ctx().stream));
  return AG_OK;
}

ag_status ag_copy_d2h_async(void* dst, const void* src, int64_t bytes) {
  AG_REQUIRE_INIT();
  if (bytes == 0) return AG_OK;
  AG_CHECK(dst && src && bytes > 0, AG_EINVAL, "ag_copy_d2h: bad arguments");
  AG_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToHost, ctx().stream));
  return AG_OK;
}

ag_status ag_copy_d2h(void* dst, const void* src, int64_t bytes) {
  ag_status s = ag_copy_d2h_async(dst, src, bytes);
  if (s != AG_OK) return s;
  AG_CUDA(cudaStreamSynchronize(ctx().stream));
  return AG_OK;
}

ag_status ag_copy_d2d(void* dst, const void* src, int64_t bytes) {
  AG_REQUIRE_INIT();
  if (bytes == 0) return AG_OK;
  AG_CHECK(dst && src && bytes > 0, AG_EINVAL, "ag_copy_d2d: bad arguments");
  AG_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToDevice, ctx().stream));
  return AG_OK;
}

ag_status ag_fill(void* dst, int64_t n, int32_t dtype, double value) {
  AG_REQUIRE_INIT();
  if (n == 0) return AG_OK;
  AG_CHECK(dst && n > 0, AG_EINVAL, "ag_fill: bad arguments");
  AG_CHECK(dtype == AG_F32 || dtype == AG_F64, AG_EUNSUPPORTED, "ag_fill: dtype %d", dtype);
  Context& c = ctx();
  if (value == 0.0) {
    AG_CUDA(cudaMemsetAsync(dst, 0, (size_t)n * (dtype == AG_F32 ? 4 : 8), c.stream));
    return AG_OK;
  }
  int64_t blocks = cdiv(n, 256);
  int64_t cap = (int64_t)c.num_sms * 16;
  if (blocks > cap) blocks = cap;
  if (dtype == AG_F32)
    k_fill<float><<<(unsigned)blocks, 256, 0, c.stream>>>((float*)dst, n, (float)value);
  else
    k_fill<double><<<(unsigned)blocks, 256, 0, c.stream>>>((double*)dst, n, value);
  AG_LAUNCHED();
  return AG_OK;
}

// ---- profiling ranges: the `@timer` labels of the reference (src/AutoGrad.jl:7-12, src/core.jl:165-166,177-182) as NVTX
// ranges, so that an nsys / ncu timeline carries the tape position of every kernel ("[ti]f[i]", "[ti]addto![i]") --------
ag_status ag_range_push(const char* label) {
  if (label) nvtxRangePushA(label);
  return AG_OK;
}

ag_status ag_range_pop(void) {
  nvtxRangePop();
  return AG_OK;
}

// ---- events -----------------------------------------------------------------------------------
ag_status ag_event_create(uint64_t* ev) {
  AG_REQUIRE_INIT();
  cudaEvent_t e;
  AG_CUDA(cudaEventCreate(&e));
  *ev = (uint64_t)(uintptr_t)e;
  return AG_OK;
}

ag_status ag_event_record(uint64_t ev) {
  AG_REQUIRE_INIT();
  AG_CUDA(cudaEventRecord((cudaEvent_t)(uintptr_t)ev, ctx().stream));
  return AG_OK;
}

ag_status ag_event_elapsed_ms(uint64_t a, uint64_t b, float* ms) {
  AG_REQUIRE_INIT();
  AG_CUDA(cudaEventSynchronize((cudaEvent_t)(uintptr_t)b));
  AG_CUDA(cudaEventElapsedTime(ms, (cudaEvent_t)(uintptr_t)a, (cudaEvent_t)(uintptr_t)b));
  return AG_OK;
}

ag_status ag_event_destroy(uint64_t ev) {
  if (!ctx().ready) return AG_OK;
  AG_CUDA(cudaEventDestroy((cudaEvent_t)(uintptr_t)ev));
  return AG_OK;
}

// ---- graphs -----------------------------------------------------------------------------------
ag_status ag_graph_begin(void) {
  AG_REQUIRE_INIT();
  Context& c = ctx();
  AG_CHECK(!c.capturing, AG_ESTATE, "ag_graph_begin: a capture is already active");
  AG_CUDA(cudaStreamSynchronize(c.stream));
  {
    std::lock_guard<std::mutex> lk(g_mu);
    g_capturing_graph = g_next_graph++;
    g_graphs[g_capturing_graph] = new GraphRec();
  }
  cudaError_t e = cudaStreamBeginCapture(c.stream, cudaStreamCaptureModeRelaxed);
  if (e != cudaSuccess) {
    std::lock_guard<std::mutex> lk(g_mu);
    delete g_graphs[g_capturing_graph];
    g_graphs.erase(g_capturing_graph);
    g_capturing_graph = 0;
    return cuda_fail(e, "cudaStreamBeginCapture", __FILE__, __LINE__);
  }
  c.capturing = true;
  g_capture_launch0 = c.launches;
  return AG_OK;
}

ag_status ag_graph_end(uint64_t* out) {
  AG_REQUIRE_INIT();
  Context& c = ctx();
  AG_CHECK(c.capturing, AG_ESTATE, "ag_graph_end: no capture active");
  uint64_t id = g_capturing_graph;
  GraphRec* gr = g_graphs[id];
  cudaGraph_t g = nullptr;
  cudaError_t e = cudaStreamEndCapture(c.stream, &g);
  c.capturing = false;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    g_capturing_graph = 0;
    gr->free_in_capture.clear();
  }
  if (e != cudaSuccess || !g) {
    cudaGetLastError();
    ag_graph_destroy(id);
    return cuda_fail(e, "cudaStreamEndCapture", __FILE__, __LINE__);
  }
  gr->graph = g;
  gr->kernels = c.launches - g_capture_launch0;
  c.launches = g_capture_launch0;  // captured launches did not execute
  e = cudaGraphInstantiate(&gr->exec, g, 0);
  if (e != cudaSuccess) {
    cudaGetLastError();
    ag_graph_destroy(id);
    return cuda_fail(e, "cudaGraphInstantiate", __FILE__, __LINE__);
  }
  *out = id;
  return AG_OK;
}

ag_status ag_graph_launch(uint64_t id) {
  AG_REQUIRE_INIT();
  auto it = g_graphs.find(id);
  AG_CHECK(it != g_graphs.end() && it->second->exec, AG_EINVAL, "ag_graph_launch: unknown graph");
  AG_CUDA(cudaGraphLaunch(it->second->exec, ctx().stream));
  ctx().launches += it->second->kernels;
  return AG_OK;
}

ag_status ag_graph_destroy(uint64_t id) {
  if (!ctx().ready) return AG_OK;
  cudaStreamSynchronize(ctx().stream);
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_graphs.find(id);
  if (it == g_graphs.end()) return AG_OK;
  GraphRec* gr = it->second;
  if (gr->exec) cudaGraphExecDestroy(gr->exec);
  if (gr->graph) cudaGraphDestroy(gr->graph);
  // blocks the graph owned: those already freed by the host go back to the pool, live ones are
  // released from graph ownership and return to the pool when freed.
  for (void* p : gr->blocks) {
    auto lit = g_live.find(p);
    if (lit != g_live.end()) {
      lit->second.graph = 0;
    } else {
      auto ait = g_all.find(p);
      if (ait != g_all.end()) g_free[ait->second].push_back(p);
    }
  }
  delete gr;
  g_graphs.erase(it);
  if (g_graphs.empty()) {                   // nothing can replay into the outgrown workspaces any more
    for (void* p : g_parked_ws) cudaFree(p);
    g_parked_ws.clear();
  }
  return AG_OK;
}

}  // extern "C"
