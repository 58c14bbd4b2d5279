// reduce.cu -- sum reductions: sum(x), sum(abs2,x), sum(x;dims) and the reductions inside
// unbroadcast (src/unbroadcast.jl:9-26; rules src/base.jl:245-247).
//
// All reductions are deterministic (fixed tree, no float atomics): phase 1 writes per-block partials
// to the workspace, phase 2 reduces the partials.  Per-thread accumulation is in T with several
// independent accumulators (pairwise-like error growth); block/partial combination is in double.
// HBM-bound: algorithmic bytes N*s + n_out*s.
#include "common.cuh"

namespace agb {

constexpr int kRT = 256;  // threads per reduction block

template <int MAP, typename T>
__device__ __forceinline__ T premap(T v) {
  if (MAP == AG_R_ABS2) return v * v;
  if (MAP == AG_R_ABS) return fabs(v);
  return v;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// sum of `v` over the block (kRT threads); result valid in thread 0
__device__ __forceinline__ double block_sum(double v) {
  __shared__ double sm[kRT / 32];
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = 0;
  if (threadIdx.x < 32) {
    r = threadIdx.x < kRT / 32 ? sm[threadIdx.x] : 0.0;
    r = warp_sum(r);
  }
  __syncthreads();
  return r;
}

// ---- contiguous segments: out[o] = sum_r map(x[r + red*o]) -----------------------------------
// grid (outer, nchunks); chunk c of segment o covers [c*chunk, min(red,(c+1)*chunk))
template <int MAP, typename T, typename TO>
__global__ void __launch_bounds__(kRT)
    k_seg_block(const T* __restrict__ x, int64_t red, int64_t chunk, TO* __restrict__ out) {
  const int64_t o = blockIdx.x, c = blockIdx.y;
  const T* __restrict__ p = x + o * red;
  int64_t lo = c * chunk, hi = lo + chunk;
  if (hi > red) hi = red;
  T a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  // 128-bit loads over the aligned middle of [lo,hi)
  constexpr int N = 16 / sizeof(T);
  int64_t head = lo;
  while (head < hi && (((uintptr_t)(p + head)) & 15)) ++head;
  if (threadIdx.x == 0)
    for (int64_t i = lo; i < head; ++i) a0 += premap<MAP>(p[i]);
  int64_t nvec = (hi - head) / N;
  const T* __restrict__ pv = p + head;
  for (int64_t v = threadIdx.x; v < nvec; v += 2 * kRT) {
    if constexpr (N == 4) {
      float4 q = *reinterpret_cast<const float4*>(pv + v * N);
      a0 += premap<MAP>((T)q.x); a1 += premap<MAP>((T)q.y);
      a2 += premap<MAP>((T)q.z); a3 += premap<MAP>((T)q.w);
      if (v + kRT < nvec) {
        float4 q2 = *reinterpret_cast<const float4*>(pv + (v + kRT) * N);
        a0 += premap<MAP>((T)q2.x); a1 += premap<MAP>((T)q2.y);
        a2 += premap<MAP>((T)q2.z); a3 += premap<MAP>((T)q2.w);
      }
    } else {
      double2 q = *reinterpret_cast<const double2*>(pv + v * N);
      a0 += premap<MAP>((T)q.x); a1 += premap<MAP>((T)q.y);
      if (v + kRT < nvec) {
        double2 q2 = *reinterpret_cast<const double2*>(pv + (v + kRT) * N);
        a2 += premap<MAP>((T)q2.x); a3 += premap<MAP>((T)q2.y);
      }
    }
  }
  if (threadIdx.x == 0)
    for (int64_t i = head + nvec * N; i < hi; ++i) a1 += premap<MAP>(p[i]);
  double s = block_sum(((double)a0 + (double)a1) + ((double)a2 + (double)a3));
  if (threadIdx.x == 0) out[o * gridDim.y + c] = (TO)s;
}

// short segments: one thread per segment (red <= 64)
template <int MAP, typename T, typename TI>
__global__ void __launch_bounds__(kRT)
    k_seg_thread(const TI* __restrict__ x, int64_t red, int64_t outer, T* __restrict__ out) {
  int64_t o = (int64_t)blockIdx.x * kRT + threadIdx.x;
  if (o >= outer) return;
  const TI* __restrict__ p = x + o * red;
  double s = 0;
  for (int64_t r = 0; r < red; ++r) s += (double)premap<MAP>(p[r]);
  out[o] = (T)s;
}

// ---- strided: out[i,o] = sum_r map(x[i + inner*(r + red*o)]), inner <= kRT ---------------------
// The (inner x red) slab of one `o` is contiguous: thread t walks it with a stride S that is a
// multiple of inner, so it always sees the same i and every warp load is fully coalesced.
// grid (outer, nchunks); partial/out layout [o][c][i].
template <int MAP, typename T, typename TI, typename TO>
__global__ void __launch_bounds__(kRT)
    k_strided_small(const TI* __restrict__ x, int64_t inner, int64_t red, int64_t rows_per_chunk,
                    TO* __restrict__ out) {
  __shared__ double sm[kRT];
  const int64_t o = blockIdx.x, c = blockIdx.y;
  const int S = (kRT / (int)inner) * (int)inner;
  const int t = threadIdx.x;
  int64_t r0 = c * rows_per_chunk, r1 = r0 + rows_per_chunk;
  if (r1 > red) r1 = red;
  const TI* __restrict__ p = x + o * inner * red;
  const int64_t lo = r0 * inner, hi = r1 * inner;
  T a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  if (t < S) {
    int64_t q = lo + t;
    for (; q + 3 * (int64_t)S < hi; q += 4 * (int64_t)S) {
      T v0 = (T)p[q], v1 = (T)p[q + S], v2 = (T)p[q + 2 * (int64_t)S], v3 = (T)p[q + 3 * (int64_t)S];
      a0 += premap<MAP>(v0); a1 += premap<MAP>(v1); a2 += premap<MAP>(v2); a3 += premap<MAP>(v3);
    }
    for (; q < hi; q += S) a0 += premap<MAP>((T)p[q]);
  }
  sm[t] = ((double)a0 + (double)a1) + ((double)a2 + (double)a3);
  __syncthreads();
  if (t < inner) {
    double s = 0;
    for (int j = t; j < S; j += (int)inner) s += sm[j];
    out[(o * gridDim.y + c) * inner + t] = (TO)s;
  }
}

// Vector variant of the above for inner | kRT*N (N = elements per 16 bytes): every thread streams 128-bit
// loads with stride kRT*N elements, so its N lanes always belong to the same N consecutive i.
// grid (outer, nchunks); partial/out layout [o][c][i].
template <int MAP, typename T, typename TO>
__global__ void __launch_bounds__(kRT)
    k_strided_vec(const T* __restrict__ x, int64_t inner, int64_t red, int64_t rows_per_chunk,
                  TO* __restrict__ out) {
  constexpr int N = 16 / sizeof(T);
  struct alignas(16) V { T v[N]; };
  __shared__ double sm[kRT * N];
  const int64_t o = blockIdx.x, c = blockIdx.y;
  const int t = threadIdx.x;
  int64_t r0 = c * rows_per_chunk, r1 = r0 + rows_per_chunk;
  if (r1 > red) r1 = red;
  const T* __restrict__ p = x + o * inner * red;
  const int64_t lo = r0 * inner, hi = r1 * inner, S = (int64_t)kRT * N;
  T a0[N], a1[N], a2[N], a3[N];
#pragma unroll
  for (int j = 0; j < N; ++j) a0[j] = a1[j] = a2[j] = a3[j] = T(0);
  int64_t q = lo + (int64_t)t * N;
  for (; q + 3 * S < hi; q += 4 * S) {
    const V v0 = *reinterpret_cast<const V*>(p + q), v1 = *reinterpret_cast<const V*>(p + q + S),
            v2 = *reinterpret_cast<const V*>(p + q + 2 * S), v3 = *reinterpret_cast<const V*>(p + q + 3 * S);
#pragma unroll
    for (int j = 0; j < N; ++j) {
      a0[j] += premap<MAP>(v0.v[j]); a1[j] += premap<MAP>(v1.v[j]);
      a2[j] += premap<MAP>(v2.v[j]); a3[j] += premap<MAP>(v3.v[j]);
    }
  }
  for (; q < hi; q += S) {
    const V v0 = *reinterpret_cast<const V*>(p + q);
#pragma unroll
    for (int j = 0; j < N; ++j) a0[j] += premap<MAP>(v0.v[j]);
  }
#pragma unroll
  for (int j = 0; j < N; ++j) sm[t * N + j] = ((double)a0[j] + (double)a1[j]) + ((double)a2[j] + (double)a3[j]);
  __syncthreads();
  if (t < inner) {
    double s = 0;
    const int groups = (int)(S / inner);       // threads per i-group
    const int gstride = (int)(inner / N);      // thread distance between members of a group
    for (int m = 0; m < groups; ++m) s += sm[((t / N) + m * gstride) * N + (t % N)];
    out[(o * gridDim.y + c) * inner + t] = (TO)s;
  }
}

// inner > kRT: thread per i, loop over the chunk's rows.  grid (outer, nchunks, itiles)
template <int MAP, typename T, typename TI, typename TO>
__global__ void __launch_bounds__(kRT)
    k_strided_large(const TI* __restrict__ x, int64_t inner, int64_t red, int64_t rows_per_chunk,
                    TO* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.z * kRT + threadIdx.x;
  const int64_t c = blockIdx.y, o = blockIdx.x;
  if (i >= inner) return;
  int64_t r0 = c * rows_per_chunk, r1 = r0 + rows_per_chunk;
  if (r1 > red) r1 = red;
  const TI* __restrict__ p = x + o * inner * red + i;
  T a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  int64_t r = r0;
  for (; r + 3 < r1; r += 4) {
    T v0 = (T)p[r * inner], v1 = (T)p[(r + 1) * inner], v2 = (T)p[(r + 2) * inner],
      v3 = (T)p[(r + 3) * inner];
    a0 += premap<MAP>(v0); a1 += premap<MAP>(v1); a2 += premap<MAP>(v2); a3 += premap<MAP>(v3);
  }
  for (; r < r1; ++r) a0 += premap<MAP>((T)p[r * inner]);
  double s = ((double)a0 + (double)a1) + ((double)a2 + (double)a3);
  out[(o * gridDim.y + c) * inner + i] = (TO)s;
}

// inner > kRT and a short reduced extent (LSTM bias gradients: sum(dy, dims=2) of (512,256)): the thread-per-i
// kernel above would run two blocks that each walk all rows.  Here a block owns 32 consecutive i (one 128-byte
// run per row) and its 8 warps take rows w, w+8, ... with 8 loads in flight; fixed-order combine (deterministic).
// grid (outer, 1, ceil(inner/32)), ONE launch.
template <int MAP, int LW, typename T, typename TI, typename TO>
__global__ void __launch_bounds__(256)
    k_strided_tile(const TI* __restrict__ x, int64_t inner, int64_t red, TO* __restrict__ out) {
  constexpr int G = 256 / LW;                    // row groups of the block: group w takes rows w, w+G, ...
  __shared__ double sm[G][LW + 1];
  const int lane = threadIdx.x % LW, w = threadIdx.x / LW;
  const int64_t i = (int64_t)blockIdx.z * LW + lane, o = blockIdx.x;
  double s = 0;
  if (i < inner) {
    const TI* __restrict__ p = x + o * inner * red + i;
    T a[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) a[u] = T(0);
    int64_t r = w;
    for (; r + 7 * G < red; r += 8 * G) {
      T v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = (T)p[(r + G * u) * inner];
#pragma unroll
      for (int u = 0; u < 8; ++u) a[u] += premap<MAP>(v[u]);
    }
    for (; r < red; r += G) a[0] += premap<MAP>((T)p[r * inner]);
    s = (((double)a[0] + (double)a[1]) + ((double)a[2] + (double)a[3])) +
        (((double)a[4] + (double)a[5]) + ((double)a[6] + (double)a[7]));
  }
  sm[w][lane] = s;
  __syncthreads();
  if (w == 0 && i < inner) {
    double r = 0;
#pragma unroll
    for (int q = 0; q < G; ++q) r += sm[q][lane];
    out[o * inner + i] = (TO)r;
  }
}

// The same tile kernel for up to 32 arrays of ONE layout in one launch (blockIdx.y = array): the bias-gradient row sums
// of the four LSTM gates over many timesteps (the host defers them, lazy.py kind "rs").  Same arithmetic, same order:
// bit-identical to separate launches.
constexpr int kBatchMax = 32;
struct BatchPtrs {
  const void* x[kBatchMax];
  void* out[kBatchMax];
};
template <int MAP, int LW, typename T>
__global__ void __launch_bounds__(256)
    k_strided_tile_batch(const __grid_constant__ BatchPtrs b, int64_t inner, int64_t red) {
  constexpr int G = 256 / LW;
  __shared__ double sm[G][LW + 1];
  const T* __restrict__ x = (const T*)b.x[blockIdx.y];
  T* __restrict__ out = (T*)b.out[blockIdx.y];
  const int lane = threadIdx.x % LW, w = threadIdx.x / LW;
  const int64_t i = (int64_t)blockIdx.z * LW + lane, o = blockIdx.x;
  double s = 0;
  if (i < inner) {
    const T* __restrict__ p = x + o * inner * red + i;
    T a[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) a[u] = T(0);
    int64_t r = w;
    for (; r + 7 * G < red; r += 8 * G) {
      T v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = p[(r + G * u) * inner];
#pragma unroll
      for (int u = 0; u < 8; ++u) a[u] += premap<MAP>(v[u]);
    }
    for (; r < red; r += G) a[0] += premap<MAP>(p[r * inner]);
    s = (((double)a[0] + (double)a[1]) + ((double)a[2] + (double)a[3])) +
        (((double)a[4] + (double)a[5]) + ((double)a[6] + (double)a[7]));
  }
  sm[w][lane] = s;
  __syncthreads();
  if (w == 0 && i < inner) {
    double r = 0;
#pragma unroll
    for (int q = 0; q < G; ++q) r += sm[q][lane];
    out[o * inner + i] = (T)r;
  }
}

// phase 2 of the strided reductions: out[i,o] = sum_c part[o][c][i]; 8 outputs x 32 lanes per block: a lane sums
// every 32nd partial (8 consecutive outputs = one 64-byte run per load), fixed-order combine (deterministic).
// grid (ceil(inner/8), outer)
template <typename TO>
__global__ void __launch_bounds__(256)
    k_partials_reduce(const double* __restrict__ part, int nchunks, int64_t inner, TO* __restrict__ out) {
  __shared__ double sm[32][9];
  const int li = threadIdx.x % 8, g = threadIdx.x / 8;
  const int64_t i = (int64_t)blockIdx.x * 8 + li, o = blockIdx.y;
  const double* __restrict__ p = part + o * nchunks * inner;
  double s0 = 0, s1 = 0;
  if (i < inner) {
    int c = g;
    for (; c + 7 * 32 < nchunks; c += 8 * 32) {      // 8 independent loads in flight
      double v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = p[(int64_t)(c + 32 * u) * inner + i];
#pragma unroll
      for (int u = 0; u < 8; u += 2) { s0 += v[u]; s1 += v[u + 1]; }
    }
    for (; c < nchunks; c += 32) s0 += p[(int64_t)c * inner + i];
  }
  sm[g][li] = s0 + s1;
  __syncthreads();
  if (g == 0 && i < inner) {
    double r = 0;
#pragma unroll
    for (int q = 0; q < 32; ++q) r += sm[q][li];
    out[o * inner + i] = (TO)r;
  }
}

// The same combine when there are MANY chunks but only a few outputs (bias gradients: 64 outputs, 2048 chunks): the
// grid above would be 8 blocks.  Here the chunk range is split over gridDim.z blocks; each writes its partial sums to
// a second-level buffer and the LAST block to finish for an output tile (atomic ticket) adds those in a fixed order,
// so the result does not depend on the schedule.  grid (ceil(inner/8), outer, Z).
constexpr int kPRTiles = 1024, kPRZ = 16;
static int32_t* g_pr_tickets = nullptr;     // kPRTiles tickets, zero between launches (the last block resets its own)
static double* g_pr_lvl2 = nullptr;         // kPRTiles * kPRZ * 8 second-level partials

bool reduce_scratch_init() {
  if (g_pr_tickets) return true;
  if (cudaMalloc((void**)&g_pr_tickets, kPRTiles * sizeof(int32_t)) != cudaSuccess ||
      cudaMalloc((void**)&g_pr_lvl2, (size_t)kPRTiles * kPRZ * 8 * sizeof(double)) != cudaSuccess) {
    cudaGetLastError();
    set_error("out of device memory allocating the reduction scratch");
    return false;
  }
  cudaMemset(g_pr_tickets, 0, kPRTiles * sizeof(int32_t));
  return true;
}

template <typename TO>
__global__ void __launch_bounds__(256)
    k_partials_reduce_split(const double* __restrict__ part, int nchunks, int64_t inner, double* __restrict__ lvl2,
                            int32_t* __restrict__ tickets, TO* __restrict__ out) {
  __shared__ double sm[32][9];
  __shared__ int last;
  const int li = threadIdx.x % 8, g = threadIdx.x / 8, Z = gridDim.z, z = blockIdx.z;
  const int64_t i = (int64_t)blockIdx.x * 8 + li, o = blockIdx.y;
  const int tile = (int)(o * gridDim.x + blockIdx.x);
  const int per = (nchunks + Z - 1) / Z, c0 = z * per, c1 = min(nchunks, c0 + per);
  const double* __restrict__ p = part + o * nchunks * inner;
  double s0 = 0, s1 = 0;
  if (i < inner) {
    int c = c0 + g;
    for (; c + 7 * 32 < c1; c += 8 * 32) {           // 8 independent loads in flight
      double v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = p[(int64_t)(c + 32 * u) * inner + i];
#pragma unroll
      for (int u = 0; u < 8; u += 2) { s0 += v[u]; s1 += v[u + 1]; }
    }
    for (; c < c1; c += 32) s0 += p[(int64_t)c * inner + i];
  }
  sm[g][li] = s0 + s1;
  __syncthreads();
  if (g == 0) {
    double r = 0;
#pragma unroll
    for (int q = 0; q < 32; ++q) r += sm[q][li];
    lvl2[((int64_t)tile * Z + z) * 8 + li] = r;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(&tickets[tile], 1) == Z - 1;
  __syncthreads();
  if (last) {
    if (g == 0 && i < inner) {
      double r = 0;
      for (int q = 0; q < Z; ++q) r += __ldcg(&lvl2[((int64_t)tile * Z + q) * 8 + li]);
      out[o * inner + i] = (TO)r;
    }
    if (threadIdx.x == 0) tickets[tile] = 0;
  }
}

// out[i,o] = sum_c part[o][c][i]: picks the split form when the plain grid would leave the GPU idle
template <typename TO>
static void launch_partials_reduce(const double* part, int64_t nchunks, int64_t inner, int64_t outer, TO* out) {
  Context& c = ctx();
  const int64_t tiles = cdiv(inner, 8) * outer;
  if (tiles < 64 && tiles <= kPRTiles && nchunks >= 512 && g_pr_tickets) {
    int Z = (int)(nchunks / 128);
    if (Z > kPRZ) Z = kPRZ;
    k_partials_reduce_split<TO><<<dim3((unsigned)cdiv(inner, 8), (unsigned)outer, (unsigned)Z), 256, 0, c.stream>>>(
        part, (int)nchunks, inner, g_pr_lvl2, g_pr_tickets, out);
  } else {
    k_partials_reduce<TO><<<dim3((unsigned)cdiv(inner, 8), (unsigned)outer), 256, 0, c.stream>>>(part, (int)nchunks,
                                                                                                  inner, out);
  }
}

// Partials are kept in double in the workspace so the second phase loses nothing.
template <int MAP, typename T>
static ag_status sum_dim_impl(const T* x, int64_t inner, int64_t red, int64_t outer, T* out) {
  Context& c = ctx();
  const int64_t target_blocks = (int64_t)c.num_sms * 8;
  if (inner == 1) {
    if (red <= 64) {
      k_seg_thread<MAP, T, T><<<(unsigned)cdiv(outer, kRT), kRT, 0, c.stream>>>(x, red, outer, out);
      AG_LAUNCHED();
      return AG_OK;
    }
    int64_t nchunks = 1;
    if (outer < target_blocks) {
      nchunks = cdiv(target_blocks, outer);
      int64_t maxc = cdiv(red, (int64_t)kRT * 16);  // >= 4096 elements per block
      if (nchunks > maxc) nchunks = maxc;
      if (nchunks > 1024) nchunks = 1024;
      if (nchunks < 1) nchunks = 1;
    }
    int64_t chunk = cdiv(red, nchunks);
    chunk = cdiv(chunk, 4) * 4;
    nchunks = cdiv(red, chunk);
    dim3 grid((unsigned)outer, (unsigned)nchunks);
    if (nchunks == 1) {
      k_seg_block<MAP, T, T><<<grid, kRT, 0, c.stream>>>(x, red, chunk, out);
      AG_LAUNCHED();
      return AG_OK;
    }
    double* part = (double*)workspace((size_t)(nchunks * outer) * sizeof(double));
    if (!part) return AG_ENOMEM;
    k_seg_block<MAP, T, double><<<grid, kRT, 0, c.stream>>>(x, red, chunk, part);
    AG_LAUNCHED();
    if (nchunks <= 64) {
      k_seg_thread<AG_R_ID, T, double><<<(unsigned)cdiv(outer, kRT), kRT, 0, c.stream>>>(
          part, nchunks, outer, out);
    } else {
      k_seg_block<AG_R_ID, double, T><<<dim3((unsigned)outer, 1), kRT, 0, c.stream>>>(
          part, nchunks, nchunks, out);
    }
    AG_LAUNCHED();
    return AG_OK;
  }

  // strided
  constexpr int VN = 16 / sizeof(T);
  const bool vecok = inner <= kRT && inner % VN == 0 && ((int64_t)kRT * VN) % inner == 0 &&
                     (((uintptr_t)x) & 15) == 0 && red >= 64;
  if (vecok) {
    const int64_t rows_per_step = (int64_t)kRT * VN / inner;
    int64_t nchunks = cdiv((int64_t)c.num_sms * 4, outer);
    int64_t maxc = cdiv(red, rows_per_step * 4);
    if (nchunks > maxc) nchunks = maxc;
    if (nchunks < 1) nchunks = 1;
    AG_CHECK(outer <= 65535 || nchunks == 1, AG_EUNSUPPORTED, "ag_sum_dim: outer extent too large");
    int64_t rows = cdiv(red, nchunks);
    nchunks = cdiv(red, rows);
    if (nchunks == 1) {
      k_strided_vec<MAP, T, T><<<dim3((unsigned)outer, 1), kRT, 0, c.stream>>>(x, inner, red, rows, out);
      AG_LAUNCHED();
      return AG_OK;
    }
    double* part = (double*)workspace((size_t)(nchunks * outer * inner) * sizeof(double));
    if (!part) return AG_ENOMEM;
    k_strided_vec<MAP, T, double><<<dim3((unsigned)outer, (unsigned)nchunks), kRT, 0, c.stream>>>(x, inner, red,
                                                                                                 rows, part);
    AG_LAUNCHED();
    launch_partials_reduce<T>(part, nchunks, inner, outer, out);
    AG_LAUNCHED();
    return AG_OK;
  }
  const bool small = inner <= kRT;
  const int64_t itiles = small ? 1 : cdiv(inner, kRT);
  AG_CHECK(itiles <= 65535, AG_EUNSUPPORTED, "ag_sum_dim: inner extent %lld too large",
           (long long)inner);
  const int64_t rows_per_iter = small ? (kRT / inner) : 1;
  int64_t nchunks = 1;
  if (outer * itiles < target_blocks && red > 256) {
    nchunks = cdiv(target_blocks, outer * itiles);
    int64_t maxc = cdiv(red, rows_per_iter * 8);
    if (nchunks > maxc) nchunks = maxc;
    if (nchunks > 1024) nchunks = 1024;
    if (nchunks < 1) nchunks = 1;
  }
  int64_t rows = cdiv(red, nchunks);
  nchunks = cdiv(red, rows);
  if (nchunks == 1) {
    if (small)
      k_strided_small<MAP, T, T, T><<<dim3((unsigned)outer, 1), kRT, 0, c.stream>>>(x, inner, red, rows, out);
    else if (red >= 16 && cdiv(inner, 16) <= 65535) {
      // 16-lane tiles (64-byte runs, twice the blocks) while the grid is small, 32-lane tiles otherwise
      if (outer * cdiv(inner, 32) < 2 * c.num_sms && red >= 32)
        k_strided_tile<MAP, 16, T, T, T><<<dim3((unsigned)outer, 1, (unsigned)cdiv(inner, 16)), 256, 0, c.stream>>>(
            x, inner, red, out);
      else
        k_strided_tile<MAP, 32, T, T, T><<<dim3((unsigned)outer, 1, (unsigned)cdiv(inner, 32)), 256, 0, c.stream>>>(
            x, inner, red, out);
    }
    else
      k_strided_large<MAP, T, T, T><<<dim3((unsigned)outer, 1, (unsigned)itiles), kRT, 0, c.stream>>>(
          x, inner, red, rows, out);
    AG_LAUNCHED();
    return AG_OK;
  }
  double* part = (double*)workspace((size_t)(nchunks * outer * inner) * sizeof(double));
  if (!part) return AG_ENOMEM;
  if (small)
    k_strided_small<MAP, T, T, double><<<dim3((unsigned)outer, (unsigned)nchunks), kRT, 0, c.stream>>>(
        x, inner, red, rows, part);
  else
    k_strided_large<MAP, T, T, double><<<dim3((unsigned)outer, (unsigned)nchunks, (unsigned)itiles), kRT, 0,
                                          c.stream>>>(x, inner, red, rows, part);
  AG_LAUNCHED();
  // phase 2: partials (inner, nchunks, outer) -> out (inner, outer)
  AG_CHECK(outer <= 65535, AG_EUNSUPPORTED, "ag_sum_dim: outer extent %lld too large for a split reduction",
           (long long)outer);
  launch_partials_reduce<T>(part, nchunks, inner, outer, out);
  AG_LAUNCHED();
  return AG_OK;
}

// ---- maximum / minimum / prod over one (merged) dimension --------------------------------------------
// out[i,o] = op_r x[i,r,o].  Rules: src/base.jl:202-206,221 (`dy.*(y.==x)`, `dy.*(y./x)`).  Not on the bench
// path: one warp per output when inner == 1 (contiguous segment), one thread per output otherwise.
template <int OP, typename T>
__device__ __forceinline__ T red_combine(T a, T b) {
  if (OP == 0) return (a != a || b != b) ? (a + b) : (a > b ? a : b);   // max, NaN-propagating like Julia
  if (OP == 1) return (a != a || b != b) ? (a + b) : (a < b ? a : b);   // min
  return a * b;                                                         // prod
}

template <int OP, typename T>
__global__ void __launch_bounds__(256)
    k_reduce_op(const T* __restrict__ x, int64_t inner, int64_t red, int64_t outer, T* __restrict__ out) {
  if (inner == 1) {
    const int lane = threadIdx.x & 31;
    const int64_t o = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (o >= outer) return;
    const T* __restrict__ p = x + o * red;
    T acc = OP == 2 ? T(1) : p[0];                     // identity: 1 for prod; a real element for max/min
    for (int64_t r = lane; r < red; r += 32) acc = red_combine<OP>(acc, p[r]);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc = red_combine<OP>(acc, __shfl_xor_sync(0xffffffffu, acc, off));
    if (lane == 0) out[o] = acc;
  } else {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= inner * outer) return;
    const int64_t i = idx % inner, o = idx / inner;
    const T* __restrict__ p = x + i + inner * red * o;
    T acc = p[0];
    for (int64_t r = 1; r < red; ++r) acc = red_combine<OP>(acc, p[r * inner]);
    out[idx] = acc;
  }
}

template <typename T>
static ag_status sum_dim_map(int32_t map, const void* x, int64_t inner, int64_t red, int64_t outer,
                             void* out) {
  switch (map) {
    case AG_R_ID: return sum_dim_impl<AG_R_ID, T>((const T*)x, inner, red, outer, (T*)out);
    case AG_R_ABS2: return sum_dim_impl<AG_R_ABS2, T>((const T*)x, inner, red, outer, (T*)out);
    case AG_R_ABS: return sum_dim_impl<AG_R_ABS, T>((const T*)x, inner, red, outer, (T*)out);
  }
  set_error("ag_sum: unknown map %d", map);
  return AG_EUNSUPPORTED;
}

// second phase of a sum fused into an elementwise kernel (fuse.cu): out[i] = sum_c part[c*inner + i].
// `part` has room for 64 more doubles behind its nchunks*inner entries (second level of a long whole-array sum).
ag_status reduce_double_partials(int dtype, const double* part, int64_t nchunks, int64_t inner, void* out) {
  Context& c = ctx();
  if (inner == 1) {
    if (nchunks > 8192) {           // one block per 1/64th of the partials, then one block over the 64
      double* lvl2 = const_cast<double*>(part) + nchunks;
      k_seg_block<AG_R_ID, double, double><<<dim3(1, 64), kRT, 0, c.stream>>>(part, nchunks, cdiv(nchunks, 64), lvl2);
      AG_LAUNCHED();
      part = lvl2;
      nchunks = 64;
    }
    if (dtype == AG_F32) k_seg_block<AG_R_ID, double, float><<<dim3(1, 1), kRT, 0, c.stream>>>(part, nchunks, nchunks, (float*)out);
    else k_seg_block<AG_R_ID, double, double><<<dim3(1, 1), kRT, 0, c.stream>>>(part, nchunks, nchunks, (double*)out);
  } else {
    if (dtype == AG_F32) launch_partials_reduce<float>(part, nchunks, inner, 1, (float*)out);
    else launch_partials_reduce<double>(part, nchunks, inner, 1, (double*)out);
  }
  AG_LAUNCHED();
  return AG_OK;
}

}  // namespace agb

using namespace agb;

extern "C" {

ag_status ag_sum_dim(int32_t map, int32_t dtype, const void* x, int64_t inner, int64_t red,
                     int64_t outer, void* out) {
  AG_REQUIRE_INIT();
  AG_CHECK(dtype == AG_F32 || dtype == AG_F64, AG_EUNSUPPORTED, "ag_sum_dim: dtype %d", dtype);
  AG_CHECK(inner >= 0 && red >= 0 && outer >= 0, AG_EINVAL, "ag_sum_dim: negative extent");
  if (inner * outer == 0) return AG_OK;
  AG_CHECK(out, AG_EINVAL, "ag_sum_dim: null output");
  if (red == 0) return ag_fill(out, inner * outer, dtype, 0.0);
  AG_CHECK(x, AG_EINVAL, "ag_sum_dim: null input");
  return dtype == AG_F32 ? sum_dim_map<float>(map, x, inner, red, outer, out)
                         : sum_dim_map<double>(map, x, inner, red, outer, out);
}

ag_status ag_sum_dim_batch(int32_t map, int32_t dtype, int32_t n, const void* const* xs, int64_t inner, int64_t red,
                           int64_t outer, void* const* outs) {
  AG_REQUIRE_INIT();
  AG_CHECK(dtype == AG_F32 || dtype == AG_F64, AG_EUNSUPPORTED, "ag_sum_dim_batch: dtype %d", dtype);
  AG_CHECK(n >= 0 && (n == 0 || (xs && outs)) && inner >= 0 && red >= 0 && outer >= 0, AG_EINVAL,
           "ag_sum_dim_batch: bad arguments");
  AG_CHECK(map >= 0 && map <= 2, AG_EINVAL, "ag_sum_dim_batch: map %d", map);
  if (n == 0 || inner * outer == 0) return AG_OK;
  for (int k = 0; k < n; ++k) AG_CHECK(outs[k] && (red == 0 || xs[k]), AG_EINVAL, "ag_sum_dim_batch: null array %d", k);
  Context& c = ctx();
  // the layout of the tile kernel (ag_sum_dim's own choice for it): one launch per 32 arrays; anything else one by one
  const bool tile = red >= 32 && inner > kRT && outer * cdiv(inner, (int64_t)32) < 2 * c.num_sms &&
                    cdiv(inner, (int64_t)16) <= 65535 && red <= 256;
  if (!tile || n == 1) {
    for (int k = 0; k < n; ++k) {
      ag_status st = ag_sum_dim(map, dtype, xs[k], inner, red, outer, outs[k]);
      if (st != AG_OK) return st;
    }
    return AG_OK;
  }
  for (int k0 = 0; k0 < n; k0 += kBatchMax) {
    const int m = n - k0 < kBatchMax ? n - k0 : kBatchMax;
    BatchPtrs b;
    for (int k = 0; k < m; ++k) { b.x[k] = xs[k0 + k]; b.out[k] = outs[k0 + k]; }
    for (int k = m; k < kBatchMax; ++k) { b.x[k] = nullptr; b.out[k] = nullptr; }
    const dim3 grid((unsigned)outer, (unsigned)m, (unsigned)cdiv(inner, (int64_t)16));
#define SDB(MAP_)                                                                                              \
  do {                                                                                                         \
    if (dtype == AG_F32) k_strided_tile_batch<MAP_, 16, float><<<grid, 256, 0, c.stream>>>(b, inner, red);     \
    else k_strided_tile_batch<MAP_, 16, double><<<grid, 256, 0, c.stream>>>(b, inner, red);                    \
  } while (0)
    if (map == AG_R_ID) SDB(AG_R_ID); else if (map == AG_R_ABS2) SDB(AG_R_ABS2); else SDB(AG_R_ABS);
#undef SDB
    AG_LAUNCHED();
  }
  return AG_OK;
}

ag_status ag_reduce_dim(int32_t op, int32_t dtype, const void* x, int64_t inner, int64_t red, int64_t outer,
                        void* out) {
  AG_REQUIRE_INIT();
  AG_CHECK(dtype == AG_F32 || dtype == AG_F64, AG_EUNSUPPORTED, "ag_reduce_dim: dtype %d", dtype);
  AG_CHECK(op >= 0 && op <= 2, AG_EUNSUPPORTED, "ag_reduce_dim: op %d", op);
  AG_CHECK(inner >= 0 && red >= 0 && outer >= 0, AG_EINVAL, "ag_reduce_dim: negative extent");
  if (inner * outer == 0) return AG_OK;
  AG_CHECK(red > 0, AG_EINVAL, "ag_reduce_dim: reducing over an empty collection is not allowed (ArgumentError)");
  AG_CHECK(x && out, AG_EINVAL, "ag_reduce_dim: null pointer");
  Context& c = ctx();
  const int64_t threads = inner == 1 ? outer * 32 : inner * outer;
  const unsigned nb = (unsigned)cdiv(threads, 256);
#define RED_LAUNCH(OP_)                                                                                        \
  do {                                                                                                         \
    if (dtype == AG_F32) k_reduce_op<OP_, float><<<nb, 256, 0, c.stream>>>((const float*)x, inner, red, outer, (float*)out); \
    else k_reduce_op<OP_, double><<<nb, 256, 0, c.stream>>>((const double*)x, inner, red, outer, (double*)out);  \
  } while (0)
  if (op == 0) RED_LAUNCH(0); else if (op == 1) RED_LAUNCH(1); else RED_LAUNCH(2);
#undef RED_LAUNCH
  AG_LAUNCHED();
  return AG_OK;
}

ag_status ag_sum_all(int32_t map, int32_t dtype, const void* x, int64_t n, void* out_scalar) {
  return ag_sum_dim(map, dtype, x, 1, n, 1, out_scalar);
}

}  // extern "C"
