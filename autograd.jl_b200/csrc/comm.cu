// comm.cu -- data-parallel exchange of Param gradients: one NCCL sum-allreduce over a flat bucket
// on the library stream (BASELINE.json north_star; no counterpart in the reference, which is
// single-process).  NCCL is resolved with dlopen so the library shares whichever libnccl.so.2 the
// process already holds (torch ships 2.28.9; the system has 2.27.3) and loads without NCCL when
// only one rank is used.
#include "common.cuh"

#include <dlfcn.h>

namespace agb {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclFloat32 = 7, ncclFloat64 = 8, ncclSum = 0 };

struct Nccl {
  void* h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclComm_t comm = nullptr;
  int nranks = 1, rank = 0;
};
static Nccl g_nccl;

static ag_status nccl_load() {
  if (g_nccl.h) return AG_OK;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    g_nccl.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.h) break;
  }
  AG_CHECK(g_nccl.h, AG_ENCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
#define SYM(field, name)                                                  \
  *(void**)(&g_nccl.field) = dlsym(g_nccl.h, name);                       \
  AG_CHECK(g_nccl.field, AG_ENCCL, "libnccl is missing symbol %s", name)
  SYM(GetUniqueId, "ncclGetUniqueId");
  SYM(CommInitRank, "ncclCommInitRank");
  SYM(CommDestroy, "ncclCommDestroy");
  SYM(AllReduce, "ncclAllReduce");
  SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
  return AG_OK;
}

#define AG_NCCL(call)                                                                   \
  do {                                                                                  \
    ncclResult_t _r = (call);                                                           \
    if (_r != 0) {                                                                      \
      set_error("NCCL error %d (%s) in %s", _r, g_nccl.GetErrorString(_r), #call);      \
      return AG_ENCCL;                                                                  \
    }                                                                                   \
  } while (0)

// ---- one-shot peer-memory allreduce over NVLink -------------------------------------------------------
// The gradient bucket of the MNIST/LSTM configs is 0.2-6 MB: an NCCL allreduce of that size is pure latency
// (measured ~45 us at 8 ranks).  Here every rank publishes its bucket in a symmetric buffer that all peers
// have mapped through CUDA IPC, raises a flag in every peer's memory, and then sums all peers' buckets itself
// with plain loads over NVLink/NVSwitch, in rank order (bit-identical result on every rank, deterministic).
//   symmetric region per rank: [flags: 64 x u32][ctrl: epoch, arrive1, arrive2][pad][buf0: cap][buf1: cap]
struct P2P {
  int nranks = 0, rank = 0;
  size_t cap = 0;
  void* sym = nullptr;                 // my region
  void* peer[16] = {nullptr};          // all regions (peer[rank] == sym)
  void** peer_dev = nullptr;           // device copy of peer[]
};
static P2P g_p2p;
constexpr size_t P2P_HDR = 1024;

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

constexpr int P2P_MAXSEG = 16;
struct P2PSeg {
  void* src;        // this rank's gradient (overwritten with the global sum)
  void* dst;        // optional: dst += alpha * sum   (the axpy! update fused behind the collective)
  int64_t n;        // elements
  int64_t off;      // element offset inside the symmetric buffer (16-byte aligned)
};
struct P2PArgs {
  P2PSeg seg[P2P_MAXSEG];
  int nseg;
  double alpha;
};

// One kernel = gather the segments into my symmetric buffer, publish (flag every peer), wait for every peer,
// sum all peers' buffers in rank order, write the sums back to the segments (+ optional fused axpy!).
template <typename T>
__global__ void __launch_bounds__(256)
    k_p2p_allreduce(void* const* __restrict__ peers, int nranks, int rank, size_t cap, const P2PArgs a) {
  constexpr int NV = 16 / sizeof(T);
  struct alignas(16) V { T v[NV]; };
  uint8_t* my = (uint8_t*)peers[rank];
  uint32_t* flags = (uint32_t*)my;                       // flags[r]: last epoch published by rank r (+1)
  uint32_t* ctrl = (uint32_t*)(my + 256);                // [0] epoch, [1] arrive1, [2] arrive2
  __shared__ uint32_t s_epoch;
  if (threadIdx.x == 0) s_epoch = *(volatile uint32_t*)&ctrl[0];
  __syncthreads();
  const uint32_t e = s_epoch;
  const size_t boff = P2P_HDR + (size_t)(e & 1) * cap;
  const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
  // phase 1: publish my segments in my symmetric buffer
  T* mine = (T*)(my + boff);
  for (int k = 0; k < a.nseg; ++k) {
    const T* __restrict__ src = (const T*)a.seg[k].src;
    T* __restrict__ d = mine + a.seg[k].off;
    const int64_t n = a.seg[k].n;
    const int64_t nvec = ((((uintptr_t)src) & 15) == 0) ? n / NV : 0;
    for (int64_t v = t0; v < nvec; v += stride) ((V*)d)[v] = ((const V*)src)[v];
    for (int64_t i = nvec * NV + t0; i < n; i += stride) d[i] = src[i];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();                              // cumulative: orders the whole block's stores (bar.sync above)
    const uint32_t t = atomicAdd(&ctrl[1], 1u);
    if (t == gridDim.x - 1) {                            // whole bucket is visible: raise my flag everywhere
      ctrl[1] = 0;
      __threadfence_system();
      for (int r = 0; r < nranks; ++r) st_release_sys((uint32_t*)peers[r] + rank, e + 1);
    }
    // phase 2: wait until every rank has published epoch e (flags are written into MY memory: local polling)
    for (int r = 0; r < nranks; ++r) {
      uint32_t spin = 0;
      while ((int32_t)(ld_acquire_sys(&flags[r]) - (e + 1)) < 0) {
        if (++spin > (1u << 26)) {
          printf("libagb200 p2p allreduce: rank %d timed out waiting for rank %d (epoch %u)\n", rank, r, e);
          __trap();
        }
      }
    }
  }
  __syncthreads();
  // phase 3: sum all published buffers in rank order; 128-bit loads, all peers' loads in flight together
  const T alpha = (T)a.alpha;
  for (int k = 0; k < a.nseg; ++k) {
    T* __restrict__ src = (T*)a.seg[k].src;
    T* __restrict__ dst = (T*)a.seg[k].dst;
    const int64_t n = a.seg[k].n, off = a.seg[k].off;
    const bool vec = ((((uintptr_t)src) & 15) == 0) && (!dst || (((uintptr_t)dst) & 15) == 0);
    const int64_t nvec = vec ? n / NV : 0;
    for (int64_t v = t0; v < nvec; v += stride) {
      V acc, t[8];
#pragma unroll
      for (int j = 0; j < NV; ++j) acc.v[j] = T(0);
      for (int r0 = 0; r0 < nranks; r0 += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (r0 + u < nranks) t[u] = ((const V*)((const T*)((const uint8_t*)peers[r0 + u] + boff) + off))[v];
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (r0 + u < nranks) {
#pragma unroll
            for (int j = 0; j < NV; ++j) acc.v[j] += t[u].v[j];
          }
      }
      ((V*)src)[v] = acc;
      if (dst) {
        V w = ((V*)dst)[v];
#pragma unroll
        for (int j = 0; j < NV; ++j) w.v[j] = fma(alpha, acc.v[j], w.v[j]);
        ((V*)dst)[v] = w;
      }
    }
    for (int64_t i = nvec * NV + t0; i < n; i += stride) {
      T sum = T(0);
      for (int r = 0; r < nranks; ++r) sum += ((const T*)((const uint8_t*)peers[r] + boff))[off + i];
      src[i] = sum;
      if (dst) dst[i] = fma(alpha, sum, dst[i]);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t t = atomicAdd(&ctrl[2], 1u);
    if (t == gridDim.x - 1) {                            // every block has read the epoch: advance it
      ctrl[2] = 0;
      __threadfence();
      *(volatile uint32_t*)&ctrl[0] = e + 1;
    }
  }
}

// y_k += alpha * x_k for up to 16 tensors in one launch (the SGD loop `for w in params; axpy!(-lr, grad, w)`)
template <typename T, bool COPY>
__global__ void __launch_bounds__(256) k_multi_axpy(const P2PArgs a) {
  const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
  const T alpha = (T)a.alpha;
  for (int k = 0; k < a.nseg; ++k) {
    const T* __restrict__ x = (const T*)a.seg[k].src;
    T* __restrict__ y = (T*)a.seg[k].dst;
    for (int64_t i = t0; i < a.seg[k].n; i += stride) y[i] = COPY ? x[i] : fma(alpha, x[i], y[i]);
  }
}

static ag_status fill_args(P2PArgs& a, int32_t dtype, int32_t nseg, void* const* src, void* const* dst,
                           const int64_t* sizes, double alpha, int64_t* total_out) {
  AG_CHECK(nseg >= 1 && nseg <= P2P_MAXSEG && src && sizes, AG_EINVAL, "multi-tensor call: nseg=%d (max %d)", nseg,
           P2P_MAXSEG);
  const int64_t nv = dtype == AG_F32 ? 4 : 2;
  int64_t off = 0;
  for (int k = 0; k < nseg; ++k) {
    AG_CHECK(sizes[k] >= 0 && (sizes[k] == 0 || src[k]), AG_EINVAL, "multi-tensor call: bad segment %d", k);
    a.seg[k].src = src[k];
    a.seg[k].dst = dst ? dst[k] : nullptr;
    a.seg[k].n = sizes[k];
    a.seg[k].off = off;
    off += cdiv(sizes[k], nv) * nv;       // keep every segment 16-byte aligned inside the symmetric buffer
  }
  a.nseg = nseg;
  a.alpha = alpha;
  *total_out = off;
  return AG_OK;
}

}  // namespace agb

using namespace agb;

extern "C" {

ag_status ag_p2p_export(int64_t cap_bytes, uint8_t* handle64) {
  AG_REQUIRE_INIT();
  AG_CHECK(cap_bytes > 0 && handle64, AG_EINVAL, "ag_p2p_export: bad arguments");
  AG_CHECK(!g_p2p.sym, AG_ESTATE, "ag_p2p_export: already exported");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  g_p2p.cap = ((size_t)cap_bytes + 255) & ~size_t(255);
  const size_t total = P2P_HDR + 2 * g_p2p.cap;
  AG_CUDA(cudaMalloc(&g_p2p.sym, total));
  AG_CUDA(cudaMemset(g_p2p.sym, 0, total));
  AG_CUDA(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  AG_CUDA(cudaIpcGetMemHandle(&h, g_p2p.sym));
  memcpy(handle64, &h, 64);
  return AG_OK;
}

ag_status ag_p2p_connect(int32_t nranks, int32_t rank, const uint8_t* handles) {
  AG_REQUIRE_INIT();
  AG_CHECK(g_p2p.sym, AG_ESTATE, "ag_p2p_connect: call ag_p2p_export first");
  AG_CHECK(nranks >= 1 && nranks <= 16 && rank >= 0 && rank < nranks && handles, AG_EINVAL,
           "ag_p2p_connect: rank %d of %d", rank, nranks);
  g_p2p.nranks = nranks;
  g_p2p.rank = rank;
  for (int r = 0; r < nranks; ++r) {
    if (r == rank) {
      g_p2p.peer[r] = g_p2p.sym;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + 64 * r, 64);
    AG_CUDA(cudaIpcOpenMemHandle(&g_p2p.peer[r], h, cudaIpcMemLazyEnablePeerAccess));
  }
  AG_CUDA(cudaMalloc((void**)&g_p2p.peer_dev, sizeof(void*) * 16));
  AG_CUDA(cudaMemcpy(g_p2p.peer_dev, g_p2p.peer, sizeof(void*) * 16, cudaMemcpyHostToDevice));
  return AG_OK;
}

ag_status ag_p2p_allreduce_multi(int32_t dtype, int32_t nseg, void* const* src, void* const* dst,
                                 const int64_t* sizes, double alpha) {
  AG_REQUIRE_INIT();
  AG_CHECK(dtype == AG_F32 || dtype == AG_F64, AG_EUNSUPPORTED, "ag_p2p_allreduce_multi: dtype %d", dtype);
  AG_CHECK(g_p2p.peer_dev, AG_ESTATE, "ag_p2p_allreduce_multi: call ag_p2p_connect first");
  P2PArgs a;
  int64_t total = 0;
  ag_status st = fill_args(a, dtype, nseg, src, dst, sizes, alpha, &total);
  if (st != AG_OK) return st;
  if (total == 0) return AG_OK;
  const size_t bytes = (size_t)total * (dtype == AG_F32 ? 4 : 8);
  AG_CHECK(bytes <= g_p2p.cap, AG_EINVAL, "ag_p2p_allreduce_multi: %zu bytes exceed the symmetric buffer (%zu)",
           bytes, g_p2p.cap);
  Context& c = ctx();
  int64_t blocks = cdiv(total, 256 * 16);          // latency-bound: few blocks, 4 x 128-bit items per thread
  if (blocks > c.num_sms) blocks = c.num_sms;      // every block polls the flags: co-resident by construction
  if (dtype == AG_F32)
    k_p2p_allreduce<float><<<(unsigned)blocks, 256, 0, c.stream>>>(g_p2p.peer_dev, g_p2p.nranks, g_p2p.rank,
                                                                   g_p2p.cap, a);
  else
    k_p2p_allreduce<double><<<(unsigned)blocks, 256, 0, c.stream>>>(g_p2p.peer_dev, g_p2p.nranks, g_p2p.rank,
                                                                    g_p2p.cap, a);
  AG_LAUNCHED();
  return AG_OK;
}

ag_status ag_p2p_allreduce_sum(int32_t dtype, void* buf, int64_t n) {
  if (n == 0) return AG_OK;
  void* src[1] = {buf};
  int64_t sizes[1] = {n};
  return ag_p2p_allreduce_multi(dtype, 1, src, nullptr, sizes, 0.0);
}

ag_status ag_multi_axpy(int32_t dtype, int32_t nseg, double alpha, void* const* x, void* const* y,
                        const int64_t* sizes) {
  AG_REQUIRE_INIT();
  AG_CHECK(dtype == AG_F32 || dtype == AG_F64, AG_EUNSUPPORTED, "ag_multi_axpy: dtype %d", dtype);
  AG_CHECK(y, AG_EINVAL, "ag_multi_axpy: null y");
  P2PArgs a;
  int64_t total = 0;
  ag_status st = fill_args(a, dtype, nseg, x, y, sizes, alpha, &total);
  if (st != AG_OK) return st;
  if (total == 0) return AG_OK;
  for (int k = 0; k < nseg; ++k) AG_CHECK(sizes[k] == 0 || y[k], AG_EINVAL, "ag_multi_axpy: null y[%d]", k);
  Context& c = ctx();
  int64_t blocks = cdiv(total, 256 * 2);
  if (blocks > 4 * c.num_sms) blocks = 4 * c.num_sms;
  if (dtype == AG_F32) k_multi_axpy<float, false><<<(unsigned)blocks, 256, 0, c.stream>>>(a);
  else k_multi_axpy<double, false><<<(unsigned)blocks, 256, 0, c.stream>>>(a);
  AG_LAUNCHED();
  return AG_OK;
}

ag_status ag_multi_copy(int32_t dtype, int32_t nseg, void* const* src, void* const* dst, const int64_t* sizes) {
  AG_REQUIRE_INIT();
  AG_CHECK(dtype == AG_F32 || dtype == AG_F64, AG_EUNSUPPORTED, "ag_multi_copy: dtype %d", dtype);
  AG_CHECK(dst, AG_EINVAL, "ag_multi_copy: null dst");
  P2PArgs a;
  int64_t total = 0;
  ag_status st = fill_args(a, dtype, nseg, src, dst, sizes, 0.0, &total);
  if (st != AG_OK) return st;
  if (total == 0) return AG_OK;
  for (int k = 0; k < nseg; ++k) AG_CHECK(sizes[k] == 0 || dst[k], AG_EINVAL, "ag_multi_copy: null dst[%d]", k);
  Context& c = ctx();
  int64_t blocks = cdiv(total, 256 * 2);
  if (blocks > 4 * c.num_sms) blocks = 4 * c.num_sms;
  if (dtype == AG_F32) k_multi_axpy<float, true><<<(unsigned)blocks, 256, 0, c.stream>>>(a);
  else k_multi_axpy<double, true><<<(unsigned)blocks, 256, 0, c.stream>>>(a);
  AG_LAUNCHED();
  return AG_OK;
}

ag_status ag_p2p_destroy(void) {
  if (!g_p2p.sym) return AG_OK;
  cudaDeviceSynchronize();
  for (int r = 0; r < g_p2p.nranks; ++r)
    if (r != g_p2p.rank && g_p2p.peer[r]) cudaIpcCloseMemHandle(g_p2p.peer[r]);
  if (g_p2p.peer_dev) cudaFree(g_p2p.peer_dev);
  cudaFree(g_p2p.sym);
  g_p2p = P2P();
  return AG_OK;
}

ag_status ag_comm_unique_id(uint8_t* id128) {
  AG_CHECK(id128, AG_EINVAL, "ag_comm_unique_id: null");
  ag_status s = nccl_load();
  if (s != AG_OK) return s;
  ncclUniqueId id;
  AG_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(id128, id.internal, 128);
  return AG_OK;
}

ag_status ag_comm_init(int32_t nranks, int32_t rank, const uint8_t* id128) {
  AG_REQUIRE_INIT();
  AG_CHECK(nranks >= 1 && rank >= 0 && rank < nranks, AG_EINVAL, "ag_comm_init: rank %d of %d", rank,
           nranks);
  AG_CHECK(!g_nccl.comm, AG_ESTATE, "ag_comm_init: communicator already initialised");
  g_nccl.nranks = nranks;
  g_nccl.rank = rank;
  if (nranks == 1) return AG_OK;
  AG_CHECK(id128, AG_EINVAL, "ag_comm_init: null id");
  ag_status s = nccl_load();
  if (s != AG_OK) return s;
  ncclUniqueId id;
  memcpy(id.internal, id128, 128);
  AG_NCCL(g_nccl.CommInitRank(&g_nccl.comm, nranks, id, rank));
  return AG_OK;
}

ag_status ag_comm_destroy(void) {
  ag_p2p_destroy();
  if (g_nccl.comm) {
    g_nccl.CommDestroy(g_nccl.comm);
    g_nccl.comm = nullptr;
  }
  g_nccl.nranks = 1;
  g_nccl.rank = 0;
  return AG_OK;
}

ag_status ag_allreduce_sum(int32_t dtype, void* buf, int64_t n) {
  AG_REQUIRE_INIT();
  if (n == 0 || g_nccl.nranks == 1) return AG_OK;
  AG_CHECK(g_nccl.comm, AG_ESTATE, "ag_allreduce_sum: call ag_comm_init first");
  AG_CHECK(buf && n > 0, AG_EINVAL, "ag_allreduce_sum: bad arguments");
  AG_CHECK(dtype == AG_F32 || dtype == AG_F64, AG_EUNSUPPORTED, "ag_allreduce_sum: dtype %d", dtype);
  AG_NCCL(g_nccl.AllReduce(buf, buf, (size_t)n, dtype == AG_F32 ? ncclFloat32 : ncclFloat64, ncclSum,
                           g_nccl.comm, ctx().stream));
  ctx().launches += 1;
  return AG_OK;
}

}  // extern "C"
