// comm.cu -- data-parallel exchange of Param gradients: one NCCL sum-allreduce over a flat bucket
// on the library stream (BASELINE.json north_star; no counterpart in the reference, which is
// single-process).  NCCL is resolved with dlopen so the library shares whichever libnccl.so.2 the
// process already holds (torch ships 2.28.9; the system has 2.27.3) and loads without NCCL when
// only one rank is used.
#include "common.cuh"

#include <dlfcn.h>
#include <stdlib.h>

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
  ncclResult_t (*CommGetAsyncError)(ncclComm_t, ncclResult_t*) = nullptr;   // optional (old libraries lack it)
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
  *(void**)(&g_nccl.CommGetAsyncError) = dlsym(g_nccl.h, "ncclCommGetAsyncError");
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
//   symmetric region per rank: [ctrl: epoch, done][flags: 16 x 64 x u32][buf0: cap][buf1: cap]
struct P2P {
  int nranks = 0, rank = 0;
  size_t cap = 0;
  void* sym = nullptr;                 // my region
  void* peer[16] = {nullptr};          // all regions (peer[rank] == sym)
  void** peer_dev = nullptr;           // device copy of peer[]
  size_t ll_cap = 0;                   // payload bytes the LL (flag-in-data) receive slots can take
  size_t ll_off = 0;                   // byte offset of the LL receive area inside the region
  int ll_ranks = 0;                    // slots allocated per parity
};
static P2P g_p2p;
constexpr size_t P2P_HDR = 8192;    // ctrl words (1 KB) + flags[16 ranks][64 blocks]

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Waits on peers are bounded by a wall-clock deadline (AGB_P2P_TIMEOUT_MS, default 10 minutes: a rank that is merely
// late -- data loading, a checkpoint write, first-call compilation -- must not kill the job; NCCL would simply wait).
// A wait that does expire raises a sticky device flag that the next ag_sync reports as AG_ENCCL; the kernel does
// not __trap (a trap poisons the CUDA context of the whole process).
__device__ int32_t g_p2p_timeout_flag = 0;
__device__ __forceinline__ uint64_t gtimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
struct Deadline {
  uint64_t t0, limit_ns;
  uint32_t spin;
  __device__ __forceinline__ Deadline(uint64_t limit) : t0(0), limit_ns(limit), spin(0) {}
  // true when the wait has expired (checked every 4096 polls, so the clock read costs nothing on the fast path)
  __device__ __forceinline__ bool expired(int who) {
    if ((++spin & 4095u) != 0) return false;
    const uint64_t now = gtimer_ns();
    if (t0 == 0) { t0 = now; return false; }
    if (now - t0 < limit_ns) return false;
    atomicExch(&g_p2p_timeout_flag, 1 + who);
    return true;
  }
};

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
  uint64_t timeout_ns; // wall-clock bound of every wait on a peer
  const void* scale;   // optional device scalar multiplied into alpha (gradient clipping); nullptr: none
};

// One kernel = gather the segments into my symmetric buffer, publish, wait for the peers, sum all peers' buffers
// in rank order, write the sums back to the segments (+ optional fused axpy!).
// Latency design: the flat index space is cut into one contiguous slice per CTA, identically on every rank, and
// CTA b only ever touches slice b.  So there is NO grid-wide synchronisation: CTA b copies its slice, fences,
// raises flag[rank][b] in every peer's memory, spins (one thread per peer) on its own flag[r][b] words, and then
// reads slice b of every peer.  Flags live in the reader's memory (local polling); they carry the epoch, buffers are
// double-buffered by epoch parity, so one barrier per call suffices.
constexpr int P2P_MAXBLK = 64;

template <typename T>
__global__ void __launch_bounds__(256)
    k_p2p_allreduce(void* const* __restrict__ peers, int nranks, int rank, size_t cap, const P2PArgs a,
                    int64_t total_vec) {
  constexpr int NV = 16 / sizeof(T);
  struct alignas(16) V { T v[NV]; };
  uint8_t* my = (uint8_t*)peers[rank];
  uint32_t* flags = (uint32_t*)(my + 1024);              // flags[r * P2P_MAXBLK + b]
  uint32_t* ctrl = (uint32_t*)my;                        // [0] epoch, [1] blocks done
  __shared__ uint32_t s_epoch;
  if (threadIdx.x == 0) s_epoch = *(volatile uint32_t*)&ctrl[0];
  __syncthreads();
  const uint32_t e = s_epoch;
  const size_t boff = P2P_HDR + (size_t)(e & 1) * cap;
  // this CTA's slice of the flat (16-byte vector) index space
  const int64_t per = (total_vec + gridDim.x - 1) / gridDim.x;
  const int64_t v0 = (int64_t)blockIdx.x * per, v1 = min(total_vec, v0 + per);
  V* mine = (V*)(my + boff);
  // phase 1: publish my slice (segments are padded to whole vectors inside the symmetric buffer)
  for (int64_t v = v0 + threadIdx.x; v < v1; v += blockDim.x) {
    // locate the segment of vector v
    int k = 0;
    while (k + 1 < a.nseg && (a.seg[k + 1].off / NV) <= v) ++k;
    const int64_t i = v * NV - a.seg[k].off;             // element offset inside the segment
    const T* __restrict__ src = (const T*)a.seg[k].src;
    V t;
    if (i + NV <= a.seg[k].n && ((((uintptr_t)src) & 15) == 0)) {
      t = *reinterpret_cast<const V*>(src + i);
    } else {
#pragma unroll
      for (int j = 0; j < NV; ++j) t.v[j] = (i + j < a.seg[k].n) ? src[i + j] : T(0);
    }
    mine[v] = t;
  }
  __syncthreads();
  if ((int)threadIdx.x < nranks) {
    __threadfence_system();                              // cumulative over the CTA's stores (bar.sync above)
    st_release_sys((uint32_t*)((uint8_t*)peers[threadIdx.x] + 1024) + rank * P2P_MAXBLK + blockIdx.x, e + 1);
    // phase 2: wait for slice b of peer `threadIdx.x`
    const uint32_t* f = flags + threadIdx.x * P2P_MAXBLK + blockIdx.x;
    Deadline dl(a.timeout_ns);
    while ((int32_t)(ld_acquire_sys(f) - (e + 1)) < 0) {
      if (dl.expired((int)threadIdx.x)) break;           // reported by the next ag_sync (AG_ENCCL)
    }
  }
  __syncthreads();
  // phase 3: sum slice b of every peer in rank order; all peers' loads in flight together
  const T alpha = (T)a.alpha;
  for (int64_t v = v0 + threadIdx.x; v < v1; v += blockDim.x) {
    V acc, t[8];
#pragma unroll
    for (int j = 0; j < NV; ++j) acc.v[j] = T(0);
    for (int r0 = 0; r0 < nranks; r0 += 8) {
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (r0 + u < nranks) t[u] = ((const V*)((const uint8_t*)peers[r0 + u] + boff))[v];
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (r0 + u < nranks) {
#pragma unroll
          for (int j = 0; j < NV; ++j) acc.v[j] += t[u].v[j];
        }
    }
    int k = 0;
    while (k + 1 < a.nseg && (a.seg[k + 1].off / NV) <= v) ++k;
    const int64_t i = v * NV - a.seg[k].off;
    T* __restrict__ src = (T*)a.seg[k].src;
    T* __restrict__ dst = (T*)a.seg[k].dst;
    if (i + NV <= a.seg[k].n && ((((uintptr_t)src) & 15) == 0) && (!dst || (((uintptr_t)dst) & 15) == 0)) {
      *reinterpret_cast<V*>(src + i) = acc;
      if (dst) {
        V w = *reinterpret_cast<V*>(dst + i);
#pragma unroll
        for (int j = 0; j < NV; ++j) w.v[j] = fma(alpha, acc.v[j], w.v[j]);
        *reinterpret_cast<V*>(dst + i) = w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < NV; ++j)
        if (i + j < a.seg[k].n) {
          src[i + j] = acc.v[j];
          if (dst) dst[i + j] = fma(alpha, acc.v[j], dst[i + j]);
        }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t t = atomicAdd(&ctrl[1], 1u);
    if (t == gridDim.x - 1) {                            // every CTA has read the epoch: advance it
      ctrl[1] = 0;
      __threadfence();
      *(volatile uint32_t*)&ctrl[0] = e + 1;
    }
  }
}

// ---- LL variant (Float32, small buckets): flag-in-data, push model -------------------------------------
// Every element travels as one 8-byte word {float bits, epoch+1}.  A rank PUSHES its words into slot `rank` of
// every peer's receive area (remote 8-byte stores are single-copy atomic), then sums the nranks slots of its OWN
// receive area in rank order, spinning on each word until its flag shows the current epoch.  No fence, no flag
// round trip, no grid-wide synchronisation: the critical path is one NVLink store latency plus local loads.
// Receive areas are double-buffered by epoch parity (a sender can be at most one epoch ahead of a receiver).
__global__ void __launch_bounds__(256)
    k_p2p_allreduce_ll(void* const* __restrict__ peers, int nranks, int rank, size_t ll_off, size_t slot_words,
                       int ll_ranks, const P2PArgs a, int64_t total) {
  uint8_t* my = (uint8_t*)peers[rank];
  uint32_t* ctrl = (uint32_t*)my;                        // [0] epoch, [1] blocks done
  __shared__ uint32_t s_epoch;
  if (threadIdx.x == 0) s_epoch = *(volatile uint32_t*)&ctrl[0];
  __syncthreads();
  const uint32_t e = s_epoch, flag = e + 1;
  const size_t par_words = (size_t)(e & 1) * ll_ranks * slot_words;
  const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
  // push my elements into slot `rank` of every rank (my own included: uniform summation order)
  for (int64_t i = i0; i < total; i += stride) {
    int k = 0;
    while (k + 1 < a.nseg && a.seg[k + 1].off <= i) ++k;
    const int64_t j = i - a.seg[k].off;
    const float v = j < a.seg[k].n ? ((const float*)a.seg[k].src)[j] : 0.f;
    const uint2 w = make_uint2(__float_as_uint(v), flag);
    for (int r = 0; r < nranks; ++r) {
      uint2* dst = (uint2*)((uint8_t*)peers[r] + ll_off) + par_words + (size_t)rank * slot_words + i;
      asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(dst), "r"(w.x), "r"(w.y) : "memory");
    }
  }
  // sum the slots of my receive area in rank order
  const uint2* recv = (const uint2*)(my + ll_off) + par_words;
  const float alpha = (float)a.alpha;
  for (int64_t i = i0; i < total; i += stride) {
    float sum = 0.f;
    for (int r = 0; r < nranks; ++r) {
      const uint2* src = recv + (size_t)r * slot_words + i;
      uint2 w;
      Deadline dl(a.timeout_ns);
      do {
        asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(w.x), "=r"(w.y) : "l"(src) : "memory");
        if (w.y != flag && dl.expired(r)) break;         // reported by the next ag_sync (AG_ENCCL)
      } while (w.y != flag);
      sum += __uint_as_float(w.x);
    }
    int k = 0;
    while (k + 1 < a.nseg && a.seg[k + 1].off <= i) ++k;
    const int64_t j = i - a.seg[k].off;
    if (j < a.seg[k].n) {
      ((float*)a.seg[k].src)[j] = sum;
      float* dst = (float*)a.seg[k].dst;
      if (dst) dst[j] = fmaf(alpha, sum, dst[j]);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t t = atomicAdd(&ctrl[1], 1u);
    if (t == gridDim.x - 1) {
      ctrl[1] = 0;
      __threadfence();
      *(volatile uint32_t*)&ctrl[0] = e + 1;
    }
  }
}

// y_k += alpha * x_k for up to 16 tensors in one launch (the SGD loop `for w in params; axpy!(-lr, grad, w)`)
template <typename T, bool COPY>
__global__ void __launch_bounds__(256) k_multi_axpy(const P2PArgs a) {
  const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
  T alpha = (T)a.alpha;
  if (!COPY && a.scale) alpha *= *(const T*)a.scale;
  for (int k = 0; k < a.nseg; ++k) {
    const T* __restrict__ x = (const T*)a.seg[k].src;
    T* __restrict__ y = (T*)a.seg[k].dst;
    for (int64_t i = t0; i < a.seg[k].n; i += stride) y[i] = COPY ? x[i] : fma(alpha, x[i], y[i]);
  }
}

// sum over all tensors of sum(abs2, x_k): the gradient norm of examples/charlm.jl:116-118 in one launch.
// Fixed grid, fixed per-thread order, double accumulation: deterministic.  part[blockIdx.x] = the block's sum.
template <typename T>
__global__ void __launch_bounds__(256) k_multi_sumabs2(const P2PArgs a, double* __restrict__ part) {
  __shared__ double sm[8];
  const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
  double acc = 0;
  for (int k = 0; k < a.nseg; ++k) {
    const T* __restrict__ x = (const T*)a.seg[k].src;
    T s0 = 0, s1 = 0;
    int64_t i = t0;
    for (; i + stride < a.seg[k].n; i += 2 * stride) {
      const T u = x[i], v = x[i + stride];
      s0 = fma(u, u, s0);
      s1 = fma(v, v, s1);
    }
    if (i < a.seg[k].n) s0 = fma(x[i], x[i], s0);
    acc += (double)s0 + (double)s1;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double r = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) r += sm[w];
    part[blockIdx.x] = r;
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
  a.scale = nullptr;
  static const uint64_t timeout_ns = []() {
    const char* e = getenv("AGB_P2P_TIMEOUT_MS");
    const double ms = e ? atof(e) : 600000.0;
    return (uint64_t)((ms > 1.0 ? ms : 1.0) * 1.0e6);
  }();
  a.timeout_ns = timeout_ns;
  *total_out = off;
  return AG_OK;
}

// Called by ag_sync: a peer wait that expired since the last check (0 = none, else 1 + the rank waited for).
// Asynchronous NCCL failures (a peer died, a transport error) do not surface from the enqueue call: poll the
// communicator (ncclCommGetAsyncError) at every ag_sync and before every NCCL allreduce of the fallback path, so that a
// broken job raises AG_ENCCL instead of hanging in the next collective.  0 = healthy / no communicator.
int comm_nccl_async_error(char* msg, int msglen) {
  if (!g_nccl.comm || !g_nccl.CommGetAsyncError) return 0;
  ncclResult_t st = 0;
  if (g_nccl.CommGetAsyncError(g_nccl.comm, &st) != 0) st = -1;
  if (st == 0 || st == 7 /* ncclInProgress */) return 0;
  if (msg && msglen > 0)
    snprintf(msg, msglen, "NCCL asynchronous error %d (%s)", st, st > 0 && g_nccl.GetErrorString ? g_nccl.GetErrorString(st) : "?");
  return st;
}

int comm_take_timeout() {
  if (!g_p2p.peer_dev) return 0;
  int32_t h = 0;
  if (cudaMemcpyFromSymbol(&h, g_p2p_timeout_flag, sizeof(h)) != cudaSuccess) { cudaGetLastError(); return 0; }
  if (h) {
    const int32_t z = 0;
    cudaMemcpyToSymbol(g_p2p_timeout_flag, &z, sizeof(z));
  }
  return h;
}

}  // namespace agb

using namespace agb;

extern "C" {

ag_status ag_p2p_export(int64_t cap_bytes, int32_t nranks, uint8_t* handle64) {
  AG_REQUIRE_INIT();
  AG_CHECK(cap_bytes > 0 && handle64 && nranks >= 1 && nranks <= 16, AG_EINVAL, "ag_p2p_export: bad arguments");
  AG_CHECK(!g_p2p.sym, AG_ESTATE, "ag_p2p_export: already exported");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  g_p2p.cap = ((size_t)cap_bytes + 255) & ~size_t(255);
  // LL receive area: 2 parities x nranks slots x (8-byte word per Float32 element), payload <= 1 MB
  g_p2p.ll_cap = g_p2p.cap < ((size_t)1 << 20) ? g_p2p.cap : ((size_t)1 << 20);
  g_p2p.ll_ranks = nranks;
  g_p2p.ll_off = P2P_HDR + 2 * g_p2p.cap;
  const size_t total = g_p2p.ll_off + (size_t)2 * nranks * (g_p2p.ll_cap * 2);
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
  AG_CHECK(nranks == g_p2p.ll_ranks, AG_EINVAL, "ag_p2p_connect: exported for %d ranks, connecting %d",
           g_p2p.ll_ranks, nranks);
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
  static const bool no_ll = getenv("AGB_P2P_NO_LL") != nullptr;      // experiment knob, read once
  if (dtype == AG_F32 && bytes <= g_p2p.ll_cap && !no_ll) {
    int64_t blocks = cdiv(total, 256);
    if (blocks > 2 * c.num_sms) blocks = 2 * c.num_sms;
    k_p2p_allreduce_ll<<<(unsigned)blocks, 256, 0, c.stream>>>(g_p2p.peer_dev, g_p2p.nranks, g_p2p.rank,
                                                               g_p2p.ll_off, g_p2p.ll_cap / 4, g_p2p.ll_ranks, a, total);
    AG_LAUNCHED();
    return AG_OK;
  }
  const int64_t nv = dtype == AG_F32 ? 4 : 2;
  const int64_t total_vec = total / nv;              // segments are padded to whole 16-byte vectors
  int64_t blocks = cdiv(total_vec, 256);             // latency-bound: one vector per thread where possible
  if (blocks > P2P_MAXBLK) blocks = P2P_MAXBLK;
  if (blocks < 1) blocks = 1;
  if (dtype == AG_F32)
    k_p2p_allreduce<float><<<(unsigned)blocks, 256, 0, c.stream>>>(g_p2p.peer_dev, g_p2p.nranks, g_p2p.rank,
                                                                   g_p2p.cap, a, total_vec);
  else
    k_p2p_allreduce<double><<<(unsigned)blocks, 256, 0, c.stream>>>(g_p2p.peer_dev, g_p2p.nranks, g_p2p.rank,
                                                                    g_p2p.cap, a, total_vec);
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

ag_status ag_multi_axpy_scaled(int32_t dtype, int32_t nseg, double alpha, const void* scale_dev, void* const* x,
                               void* const* y, const int64_t* sizes) {
  AG_REQUIRE_INIT();
  AG_CHECK(dtype == AG_F32 || dtype == AG_F64, AG_EUNSUPPORTED, "ag_multi_axpy_scaled: dtype %d", dtype);
  AG_CHECK(y && scale_dev, AG_EINVAL, "ag_multi_axpy_scaled: null y / scale");
  P2PArgs a;
  int64_t total = 0;
  ag_status st = fill_args(a, dtype, nseg, x, y, sizes, alpha, &total);
  if (st != AG_OK) return st;
  if (total == 0) return AG_OK;
  for (int k = 0; k < nseg; ++k) AG_CHECK(sizes[k] == 0 || y[k], AG_EINVAL, "ag_multi_axpy_scaled: null y[%d]", k);
  a.scale = scale_dev;
  Context& c = ctx();
  int64_t blocks = cdiv(total, 256 * 2);
  if (blocks > 4 * c.num_sms) blocks = 4 * c.num_sms;
  if (dtype == AG_F32) k_multi_axpy<float, false><<<(unsigned)blocks, 256, 0, c.stream>>>(a);
  else k_multi_axpy<double, false><<<(unsigned)blocks, 256, 0, c.stream>>>(a);
  AG_LAUNCHED();
  return AG_OK;
}

ag_status ag_multi_sumabs2(int32_t dtype, int32_t nseg, void* const* x, const int64_t* sizes, void* out_scalar) {
  AG_REQUIRE_INIT();
  AG_CHECK(dtype == AG_F32 || dtype == AG_F64, AG_EUNSUPPORTED, "ag_multi_sumabs2: dtype %d", dtype);
  AG_CHECK(out_scalar, AG_EINVAL, "ag_multi_sumabs2: null result");
  P2PArgs a;
  int64_t total = 0;
  ag_status st = fill_args(a, dtype, nseg, x, nullptr, sizes, 0.0, &total);
  if (st != AG_OK) return st;
  Context& c = ctx();
  int64_t blocks = cdiv(total > 0 ? total : 1, 256 * 4);
  if (blocks > 2 * c.num_sms) blocks = 2 * c.num_sms;
  double* part = (double*)workspace(((size_t)blocks + 64) * sizeof(double));
  if (!part) return AG_ENOMEM;
  if (dtype == AG_F32) k_multi_sumabs2<float><<<(unsigned)blocks, 256, 0, c.stream>>>(a, part);
  else k_multi_sumabs2<double><<<(unsigned)blocks, 256, 0, c.stream>>>(a, part);
  AG_LAUNCHED();
  return reduce_double_partials(dtype, part, blocks, 1, out_scalar);
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
  {
    char why[160];
    if (comm_nccl_async_error(why, sizeof(why))) {
      set_error("ag_allreduce_sum: %s", why);
      return AG_ENCCL;
    }
  }
  AG_CHECK(buf && n > 0, AG_EINVAL, "ag_allreduce_sum: bad arguments");
  AG_CHECK(dtype == AG_F32 || dtype == AG_F64, AG_EUNSUPPORTED, "ag_allreduce_sum: dtype %d", dtype);
  AG_NCCL(g_nccl.AllReduce(buf, buf, (size_t)n, dtype == AG_F32 ? ncclFloat32 : ncclFloat64, ncclSum,
                           g_nccl.comm, ctx().stream));
  ctx().launches += 1;
  return AG_OK;
}

}  // extern "C"
