// skinny.cu -- CUDA-core matrix products for the "tiny matrix x very wide matrix" shapes of the tape:
//   S1  C[MxN] = op(A)[MxK] * B[KxN],  M,K <= 64, N huge      forward  w2*a3  (10x64 * 64xB)  and the
//                                                              gradient w2'*dy (64x10 * 10xB), src/base.jl:26
//   S2  C[MxN] = A[MxK] * B[NxK]',     M,N <= 64, K huge      gradient dy*a3' (10xB * Bx64)
// These are HBM/L2-bound streaming passes over the wide operand (algorithmic bytes (MK+KN+MN)*s); the
// tiny operand lives in shared memory (S1: one thread per column of the wide operand; S2: one warp per k-range).  A 128-row tensor-core tile would waste >= 84 % of its rows on
// M = 10 and TMA cannot address a 40-byte leading dimension, so they stay on the FFMA/DFMA pipe.
// Reductions over K are deterministic (per-block partials, fixed-order second phase).
#include "common.cuh"

#include <stdlib.h>
#include <type_traits>

namespace agb {

constexpr int SK_T = 256;     // threads
constexpr int SK_COLS = 64;   // columns of the wide operand per block (S1)
constexpr int SK_MAXD = 64;   // max extent of the small dims

// ---- S2 ---------------------------------------------------------------------------------------------
// block b handles k in [b*kc, min(K,(b+1)*kc)); part[b][m + M*n] = sum_k A[m,k]*B[n,k]
template <typename T, int NJ>
__global__ void __launch_bounds__(SK_T)
    k_skinny_nt(int M, int N, int64_t K, int64_t kc, const T* __restrict__ A, int64_t lda,
                const T* __restrict__ B, int64_t ldb, T* __restrict__ part) {
  constexpr int KT = 32;
  __shared__ T At[KT * SK_MAXD];   // At[k*M + m]
  __shared__ T Bt[KT * SK_MAXD];   // Bt[k*N + n]
  const int t = threadIdx.x;
  const int n = t % SK_COLS, mg = t / SK_COLS;
  const int64_t k0 = (int64_t)blockIdx.x * kc;
  const int64_t k1 = min(K, k0 + kc);
  T acc[NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) acc[j] = T(0);
  for (int64_t kk = k0; kk < k1; kk += KT) {
    const int kt = (int)min((int64_t)KT, k1 - kk);
    if (lda == M) {
      const T* __restrict__ src = A + kk * lda;
      for (int i = t; i < kt * M; i += SK_T) At[i] = src[i];
    } else {
      for (int i = t; i < kt * M; i += SK_T) At[i] = A[(i % M) + (kk + i / M) * lda];
    }
    if (ldb == N) {
      const T* __restrict__ src = B + kk * ldb;
      for (int i = t; i < kt * N; i += SK_T) Bt[i] = src[i];
    } else {
      for (int i = t; i < kt * N; i += SK_T) Bt[i] = B[(i % N) + (kk + i / N) * ldb];
    }
    __syncthreads();
    if (n < N) {
      for (int k = 0; k < kt; ++k) {
        const T b = Bt[k * N + n];
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          const int m = mg + 4 * j;
          if (m < M) acc[j] = fma(At[k * M + m], b, acc[j]);
        }
      }
    }
    __syncthreads();
  }
  if (n < N) {
    T* __restrict__ dst = part + (int64_t)blockIdx.x * M * N;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int m = mg + 4 * j;
      if (m < M) dst[m + M * n] = acc[j];
    }
  }
}

// ---- S1 (column-per-thread form): C[:, col] = op(A) * B[:, col] ---------------------------------------
// One thread owns one column of the wide operand: it streams the column's K values straight from
// global memory (each thread walks whole cache lines; the warp's 32 columns are one contiguous run) and keeps
// MB accumulators in registers; op(A) sits in shared memory as As[k][m] and is read with broadcast
// 128-bit loads.  No staging of the wide operand, no block-level synchronisation in the main loop.
template <typename T, int MB>
__global__ void __launch_bounds__(128)
    k_skinny_cols(int transA, int M, int64_t N, int K, T alpha, const T* __restrict__ A, int64_t lda,
                  const T* __restrict__ B, int64_t ldb, T beta, T* __restrict__ C, int64_t ldc) {
  constexpr int NV = 16 / sizeof(T);               // elements per 128-bit access
  struct alignas(16) V { T v[NV]; };
  __shared__ __align__(16) T As[SK_MAXD * MB];     // As[k*MB + m], zero padded to MB rows
  for (int i = threadIdx.x; i < K * MB; i += 128) {
    const int k = i / MB, m = i % MB;
    T v = T(0);
    if (m < M) v = transA ? A[k + (int64_t)m * lda] : A[m + (int64_t)k * lda];
    As[i] = v;
  }
  __syncthreads();
  const int64_t col = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (col >= N) return;
  const T* __restrict__ b = B + col * ldb;
  T acc[MB];
#pragma unroll
  for (int m = 0; m < MB; ++m) acc[m] = T(0);
  // rank-1 update with a 128-bit broadcast read of NV rows of op(A) per instruction
  auto fma_k = [&](int k, T bv) {
#pragma unroll
    for (int m = 0; m < MB; m += NV) {
      const V a4 = *reinterpret_cast<const V*>(&As[k * MB + m]);
#pragma unroll
      for (int j = 0; j < NV; ++j) acc[m + j] = fma(a4.v[j], bv, acc[m + j]);
    }
  };
  // the whole column (K <= 64) is fetched before any arithmetic: one round of memory latency per thread
  const bool vecB = (ldb % NV) == 0 && ((((uintptr_t)B) & 15) == 0) && (K % NV) == 0;
  if (vecB) {
    V q[SK_MAXD / NV];
#pragma unroll
    for (int u = 0; u < SK_MAXD / NV; ++u)
      if (u * NV < K) q[u] = *reinterpret_cast<const V*>(b + u * NV);
#pragma unroll
    for (int u = 0; u < SK_MAXD / NV; ++u)
      if (u * NV < K) {
#pragma unroll
        for (int j = 0; j < NV; ++j) fma_k(u * NV + j, q[u].v[j]);
      }
  } else {
    for (int k0 = 0; k0 < K; k0 += 16) {
      T bv[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) bv[u] = (k0 + u < K) ? b[k0 + u] : T(0);
#pragma unroll
      for (int u = 0; u < 16; ++u)
        if (k0 + u < K) fma_k(k0 + u, bv[u]);
    }
  }
  T* __restrict__ c = C + col * ldc;
  if (beta == T(0)) {
    if (MB == M && (ldc % NV) == 0 && (((uintptr_t)C) & 15) == 0) {
#pragma unroll
      for (int m = 0; m < MB; m += NV) {
        V o;
#pragma unroll
        for (int j = 0; j < NV; ++j) o.v[j] = alpha * acc[m + j];
        *reinterpret_cast<V*>(c + m) = o;
      }
    } else {
#pragma unroll
      for (int m = 0; m < MB; ++m)
        if (m < M) c[m] = alpha * acc[m];
    }
  } else {
#pragma unroll
    for (int m = 0; m < MB; ++m)
      if (m < M) c[m] = alpha * acc[m] + beta * c[m];
  }
}

// ---- S2 (warp form, M <= 16, N <= 64): part[blk][m + M*n] = sum_{k in block range} A[m,k]*B[n,k] ----------
// Each warp walks a contiguous k-range: lanes own n = lane and lane+32 (coalesced row reads of B), the M
// values A[:,k] are warp-uniform (broadcast loads).  Block partials are combined through shared memory.
template <typename T, int MB>
__global__ void __launch_bounds__(128)
    k_skinny_nt_warp(int M, int N, int64_t K, int64_t kc, const T* __restrict__ A, int64_t lda,
                     const T* __restrict__ B, int64_t ldb, T* __restrict__ part) {
  constexpr int NW = 4;
  __shared__ T sm[NW][MB * 64];
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const int64_t k0 = (int64_t)blockIdx.x * kc;
  const int64_t k1 = min(K, k0 + kc);
  T acc0[MB], acc1[MB];
#pragma unroll
  for (int m = 0; m < MB; ++m) acc0[m] = acc1[m] = T(0);
  const bool n0ok = lane < N, n1ok = lane + 32 < N;
  int64_t k = k0 + wrp;
  for (; k + 7 * NW < k1; k += 8 * NW) {      // eight k per trip: all loads issued before the FMAs
    T bb0[8], bb1[8], av[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const T* __restrict__ b = B + (k + u * NW) * ldb;
      bb0[u] = n0ok ? b[lane] : T(0);
      bb1[u] = n1ok ? b[lane + 32] : T(0);
      av[u] = lane < M ? A[(k + u * NW) * lda + lane] : T(0);     // lane m holds A[m,k]
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int m = 0; m < MB; ++m) {
        const T a = __shfl_sync(0xffffffffu, av[u], m);
        acc0[m] = fma(a, bb0[u], acc0[m]);
        acc1[m] = fma(a, bb1[u], acc1[m]);
      }
    }
  }
  for (; k < k1; k += NW) {
    const T* __restrict__ b = B + k * ldb;
    const T b0 = n0ok ? b[lane] : T(0), b1 = n1ok ? b[lane + 32] : T(0);
    const T avl = lane < M ? A[k * lda + lane] : T(0);
#pragma unroll
    for (int m = 0; m < MB; ++m) {
      const T a = __shfl_sync(0xffffffffu, avl, m);
      acc0[m] = fma(a, b0, acc0[m]);
      acc1[m] = fma(a, b1, acc1[m]);
    }
  }
#pragma unroll
  for (int m = 0; m < MB; ++m) {
    sm[wrp][m * 64 + lane] = acc0[m];
    sm[wrp][m * 64 + lane + 32] = acc1[m];
  }
  __syncthreads();
  T* __restrict__ dst = part + (int64_t)blockIdx.x * M * N;
  for (int i = threadIdx.x; i < M * N; i += 128) {
    const int m = i % M, n = i / M;
    T s = T(0);
#pragma unroll
    for (int w = 0; w < NW; ++w) s += sm[w][m * 64 + n];
    dst[i] = s;
  }
}

// ======================================================================================================
// Float32 streaming forms (the hot shapes of the MLP tape).  All wide-operand traffic is coalesced: tiles are
// brought into shared memory with cp.async in the order they lie in HBM, and the per-thread ("one column per
// lane") access pattern is served from a padded shared-memory layout instead of from 32 cache lines per request.
// ======================================================================================================
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  const int sz = valid ? 16 : 0;      // src-size 0: the 16 bytes are zero-filled, nothing is read
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- S1a: C[:, col] = op(A)[MxK] * B[:, col], K % 4 == 0, K <= 64, M <= MB <= 16 (input-dominated) -----------
// One warp per tile of 32 columns.  The tile (32 x K floats, contiguous runs of K) is copied with 16-byte
// cp.async into rows of KP floats, KP/4 odd, so that "lane = column" 128-bit reads are bank-conflict free.
// Arithmetic is packed (FFMA2): accumulator pair m = (sum over even k, sum over odd k); the pair of B values is
// one half of the 128-bit tile read and op(A) is stored as As2[k/2][m] = (A[m,k], A[m,k+1]).
constexpr int S1A_WARPS = 2;
template <int MB>
__global__ void __launch_bounds__(S1A_WARPS * 32)
    k_skinny_s1a(int transA, int M, int64_t N, int K, float alpha, const float* __restrict__ A, int64_t lda,
                 const float* __restrict__ B, int64_t ldb, float* __restrict__ C, int64_t ldc) {
  __shared__ __align__(16) float2 As2[(SK_MAXD / 2) * MB];        // zero padded rows m in [M, MB)
  __shared__ __align__(16) float Ts[S1A_WARPS][32 * (SK_MAXD + 4)];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int kq = K >> 2;
  const int KP = (kq & 1) ? K : K + 4;
  const int64_t tile = (int64_t)blockIdx.x * S1A_WARPS + warp;
  const int64_t c0 = tile * 32;
  float* T = Ts[warp];
  if (c0 < N) {
    const int nchunk = 32 * kq;
    const int dc = 32 / kq, dq = 32 - dc * kq;
    int c = lane / kq, q = lane - c * kq;
    const int ncol = (int)min((int64_t)32, N - c0);
    const float* __restrict__ src = B + c0 * ldb;
    for (int e = lane; e < nchunk; e += 32) {
      const bool ok = c < ncol;
      cp_async16(T + c * KP + 4 * q, ok ? src + (int64_t)c * ldb + 4 * q : B, ok);
      c += dc; q += dq;
      if (q >= kq) { q -= kq; ++c; }
    }
  }
  cp_async_commit();
  for (int i = threadIdx.x; i < (K >> 1) * MB; i += S1A_WARPS * 32) {
    const int kp = i / MB, m = i - kp * MB;
    float2 v = make_float2(0.f, 0.f);
    if (m < M) {
      if (transA) v = *reinterpret_cast<const float2*>(A + 2 * kp + (int64_t)m * lda);   // K % 4 == 0, lda even below
      else v = make_float2(A[m + (int64_t)(2 * kp) * lda], A[m + (int64_t)(2 * kp + 1) * lda]);
    }
    As2[i] = v;
  }
  cp_async_wait<0>();
  __syncthreads();
  const int64_t col = c0 + lane;
  if (col >= N) return;
  float2 acc[MB];
#pragma unroll
  for (int m = 0; m < MB; ++m) acc[m] = make_float2(0.f, 0.f);
  const float4* __restrict__ t4 = reinterpret_cast<const float4*>(T + lane * KP);
#pragma unroll 2
  for (int j = 0; j < kq; ++j) {
    const float4 b4 = t4[j];
    const float2 bp[2] = {make_float2(b4.x, b4.y), make_float2(b4.z, b4.w)};
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const float4* __restrict__ a4 = reinterpret_cast<const float4*>(&As2[(2 * j + u) * MB]);
#pragma unroll
      for (int m = 0; m < MB; m += 2) {
        const float4 a = a4[m >> 1];
        acc[m] = __ffma2_rn(make_float2(a.x, a.y), bp[u], acc[m]);
        acc[m + 1] = __ffma2_rn(make_float2(a.z, a.w), bp[u], acc[m + 1]);
      }
    }
  }
  float* __restrict__ c = C + col * ldc;
  if ((M & 1) == 0 && (ldc & 1) == 0 && (((uintptr_t)C) & 7) == 0) {
#pragma unroll
    for (int m = 0; m < MB; m += 2)
      if (m < M)
        *reinterpret_cast<float2*>(c + m) =
            make_float2(alpha * (acc[m].x + acc[m].y), alpha * (acc[m + 1].x + acc[m + 1].y));
  } else {
#pragma unroll
    for (int m = 0; m < MB; ++m)
      if (m < M) c[m] = alpha * (acc[m].x + acc[m].y);
  }
}

// ---- S1b: C[:, col] = op(A)[MxK] * B[:, col], K <= KB <= 16, M <= 64 (output-dominated) ----------------------
// One warp per tile of 32 columns; lane owns output rows 2*lane and 2*lane+1 and keeps its two rows of op(A) in
// registers (as k-pairs for FFMA2), so every column is written as one contiguous run of M floats (64-bit stores,
// fully coalesced).
constexpr int S1B_WARPS = 2;
template <int KB>
__global__ void __launch_bounds__(S1B_WARPS * 32)
    k_skinny_s1b(int transA, int M, int64_t N, int K, float alpha, const float* __restrict__ A, int64_t lda,
                 const float* __restrict__ B, int64_t ldb, float* __restrict__ C, int64_t ldc) {
  __shared__ __align__(16) float Bs[S1B_WARPS][32 * KB];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t tile = (int64_t)blockIdx.x * S1B_WARPS + warp;
  const int64_t c0 = tile * 32;
  if (c0 >= N) return;
  float* T = Bs[warp];
  const int ncol = (int)min((int64_t)32, N - c0);
  if (KB != K)
    for (int e = lane; e < 32 * KB; e += 32)
      if (e % KB >= K) T[e] = 0.f;                    // zero the k padding (never touched by the copies)
  {
    const int dc = 32 / K, dk = 32 - dc * K;
    int c = lane / K, k = lane - c * K;
    const float* __restrict__ src = B + c0 * ldb;
    for (int e = lane; e < ncol * K; e += 32) {       // contiguous when ldb == K
      cp_async4(T + c * KB + k, src + (int64_t)c * ldb + k);
      c += dc; k += dk;
      if (k >= K) { k -= K; ++c; }
    }
  }
  cp_async_commit();
  const int m0 = 2 * lane;
  float2 a0[KB / 2], a1[KB / 2];
  auto ld_a = [&](int m, int k) -> float {
    if (k >= K || m >= M) return 0.f;
    return alpha * (transA ? A[k + (int64_t)m * lda] : A[m + (int64_t)k * lda]);
  };
#pragma unroll
  for (int k = 0; k < KB; k += 2) {
    a0[k >> 1] = make_float2(ld_a(m0, k), ld_a(m0, k + 1));
    a1[k >> 1] = make_float2(ld_a(m0 + 1, k), ld_a(m0 + 1, k + 1));
  }
  cp_async_wait<0>();
  __syncwarp();
  const bool pair = (ldc & 1) == 0 && (((uintptr_t)C) & 7) == 0 && m0 + 1 < M;
  float* __restrict__ cbase = C + c0 * ldc + m0;
#pragma unroll 4
  for (int c = 0; c < ncol; ++c) {
    const float4* __restrict__ b4 = reinterpret_cast<const float4*>(T + c * KB);
    float2 s0 = make_float2(0.f, 0.f), s1 = make_float2(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < KB; k += 4) {
      const float4 b = b4[k >> 2];
      s0 = __ffma2_rn(a0[k >> 1], make_float2(b.x, b.y), s0);
      s1 = __ffma2_rn(a1[k >> 1], make_float2(b.x, b.y), s1);
      s0 = __ffma2_rn(a0[(k >> 1) + 1], make_float2(b.z, b.w), s0);
      s1 = __ffma2_rn(a1[(k >> 1) + 1], make_float2(b.z, b.w), s1);
    }
    float* __restrict__ dst = cbase + (int64_t)c * ldc;
    const float r0 = s0.x + s0.y, r1 = s1.x + s1.y;
    if (pair) *reinterpret_cast<float2*>(dst) = make_float2(r0, r1);
    else {
      if (m0 < M) dst[0] = r0;
      if (m0 + 1 < M) dst[1] = r1;
    }
  }
}

// ---- S2 (streaming form): part[blk][m + M*n] = sum_{k in block range} A[m,k]*B[n,k], M <= MB <= 16, N <= 64 ----
// Each warp walks a contiguous k-range in chunks of 8 k through a 4-deep cp.async ring.  A 128-bit read covers
// 4 n of one k: lanes 0-15 hold the even k of a pair, lanes 16-31 the odd k; the two halves and the warps of
// the block are combined in a fixed order at the end (deterministic).  A is staged duplicated, (a, a) per
// element, so that one FFMA2 updates two n at once.
constexpr int S2_WARPS = 8;
constexpr int S2_CK = 8;        // k per chunk
constexpr int S2_DEPTH = 4;     // ring depth (chunks in flight per warp)
template <int MB>
__global__ void __launch_bounds__(S2_WARPS * 32)
    k_skinny_s2(int M, int N, int64_t K, int64_t kpw, const float* __restrict__ A, int64_t lda,
                const float* __restrict__ B, int64_t ldb, float* __restrict__ part) {
  constexpr int BW = S2_CK * 64;            // floats of B per chunk (rows padded to 64)
  constexpr int AW = S2_CK * MB * 2;        // floats of A per chunk, duplicated
  constexpr int RING = S2_DEPTH * (BW + AW);
  static_assert(RING >= 16 * MB * 4, "the block combine reuses the ring");
  extern __shared__ __align__(16) float sm[];       // S2_WARPS * RING floats
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int h = lane >> 4, q = lane & 15;
  float* ring = sm + warp * RING;
  const int64_t gw = (int64_t)blockIdx.x * S2_WARPS + warp;
  const int64_t k0 = gw * kpw;
  const int64_t k1 = min(K, k0 + kpw);
  const int nq = N >> 2;
  float2 acc[MB][2];
#pragma unroll
  for (int m = 0; m < MB; ++m) acc[m][0] = acc[m][1] = make_float2(0.f, 0.f);

  // zero the A rows' padding (m in [M, MB)) once; B lanes beyond N/4 are zero-filled by the copies
  if (MB != M)
    for (int e = lane; e < S2_DEPTH * AW; e += 32) {
      const int st = e / AW, r = e - st * AW;
      if ((r >> 1) % MB >= M) ring[st * (BW + AW) + BW + r] = 0.f;
    }
  // this lane's share of the A elements of a chunk: e = lane + 32*t -> (j = e / M, m = e % M), fixed per lane
  const int dj = 32 / M, dm = 32 - dj * M;
  const int j0 = lane / M, m0 = lane - j0 * M;
  auto issue = [&](int64_t kb, int st) {
    float* bs = ring + st * (BW + AW);
    float* as = bs + BW;
    if (kb < k1) {
#pragma unroll
      for (int i = 0; i < S2_CK / 2; ++i) {
        const int64_t k = kb + 2 * i + h;
        const bool ok = k < k1 && q < nq;
        cp_async16(bs + (2 * i + h) * 64 + 4 * q, ok ? B + k * ldb + 4 * q : B, ok);
      }
      const int kn = (int)min((int64_t)S2_CK, k1 - kb);
      int j = j0, m = m0;
      const float* __restrict__ src = A + kb * lda;
      for (int e = lane; e < kn * M; e += 32) {
        const float* g = src + m + (int64_t)j * lda;
        float* d = as + (j * MB + m) * 2;
        cp_async4(d, g);
        cp_async4(d + 1, g);
        j += dj; m += dm;
        if (m >= M) { m -= M; ++j; }
      }
    }
    cp_async_commit();
  };
  if (k0 < k1) {
#pragma unroll
    for (int d = 0; d < S2_DEPTH - 1; ++d) issue(k0 + (int64_t)d * S2_CK, d);
    int st = 0;
    for (int64_t kb = k0; kb < k1; kb += S2_CK) {
      int stn = st + S2_DEPTH - 1;
      if (stn >= S2_DEPTH) stn -= S2_DEPTH;
      issue(kb + (int64_t)(S2_DEPTH - 1) * S2_CK, stn);
      cp_async_wait<S2_DEPTH - 1>();
      __syncwarp();
      const float* bs = ring + st * (BW + AW);
      const float* as = bs + BW;
      const int kn = (int)min((int64_t)S2_CK, k1 - kb);
#pragma unroll
      for (int i = 0; i < S2_CK / 2; ++i) {
        const int j = 2 * i + h;
        if (j < kn) {                     // stale A rows of a short last chunk are never read
          const float4 b = *reinterpret_cast<const float4*>(bs + j * 64 + 4 * q);
          const float2 b01 = make_float2(b.x, b.y), b23 = make_float2(b.z, b.w);
          const float4* __restrict__ a4 = reinterpret_cast<const float4*>(as + j * MB * 2);
#pragma unroll
          for (int m = 0; m < MB; m += 2) {
            const float4 a = a4[m >> 1];            // (a_m, a_m, a_m+1, a_m+1)
            acc[m][0] = __ffma2_rn(make_float2(a.x, a.y), b01, acc[m][0]);
            acc[m][1] = __ffma2_rn(make_float2(a.x, a.y), b23, acc[m][1]);
            acc[m + 1][0] = __ffma2_rn(make_float2(a.z, a.w), b01, acc[m + 1][0]);
            acc[m + 1][1] = __ffma2_rn(make_float2(a.z, a.w), b23, acc[m + 1][1]);
          }
        }
      }
      __syncwarp();                       // the stage is refilled by the next trip's issue
      if (++st == S2_DEPTH) st = 0;
    }
  }
  cp_async_wait<0>();
  // even-k half + odd-k half
#pragma unroll
  for (int m = 0; m < MB; ++m)
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      acc[m][u].x += __shfl_xor_sync(0xffffffffu, acc[m][u].x, 16);
      acc[m][u].y += __shfl_xor_sync(0xffffffffu, acc[m][u].y, 16);
    }
  __syncthreads();                        // every warp is done with its ring: reuse it for the block combine
  float* comb = sm + warp * (16 * MB * 4);
  if (h == 0) {
#pragma unroll
    for (int m = 0; m < MB; ++m)
      *reinterpret_cast<float4*>(comb + m * 64 + 4 * q) = make_float4(acc[m][0].x, acc[m][0].y, acc[m][1].x, acc[m][1].y);
  }
  __syncthreads();
  float* __restrict__ dst = part + (int64_t)blockIdx.x * M * N;
  for (int i = threadIdx.x; i < M * N; i += S2_WARPS * 32) {
    const int m = i % M, n = i / M;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < S2_WARPS; ++w) s += sm[w * (16 * MB * 4) + m * 64 + n];
    dst[i] = s;
  }
}

// C[m + ldc*n] = alpha * sum_p part[p][i] (+ beta*C), i = m + M*n; 8 outputs x 32 lanes per block: a lane sums
// every 32nd partial, then a fixed-order combine (deterministic).
template <typename T>
__global__ void __launch_bounds__(256)
    k_reduce_partials(const T* __restrict__ part, int nparts, int64_t len, T alpha, T beta, T* __restrict__ C,
                      int64_t M, int64_t ldc) {
  __shared__ double sm[32][9];
  const int li = threadIdx.x % 8, g = threadIdx.x / 8;
  const int64_t i = (int64_t)blockIdx.x * 8 + li;
  double s0 = 0, s1 = 0;
  if (i < len) {
    int p = g;
    for (; p + 7 * 32 < nparts; p += 8 * 32) {       // 8 independent loads in flight
      T v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = part[(int64_t)(p + 32 * u) * len + i];
#pragma unroll
      for (int u = 0; u < 8; u += 2) { s0 += (double)v[u]; s1 += (double)v[u + 1]; }
    }
    for (; p < nparts; p += 32) s0 += (double)part[(int64_t)p * len + i];
  }
  sm[g][li] = s0 + s1;
  __syncthreads();
  if (g == 0 && i < len) {
    double r = 0;
#pragma unroll
    for (int q = 0; q < 32; ++q) r += sm[q][li];
    T* dst = C + (i % M) + (i / M) * ldc;
    *dst = beta == T(0) ? (T)(alpha * r) : (T)(alpha * r + (double)beta * (double)(*dst));
  }
}

template <typename T>
ag_status reduce_partials(const T* part, int nparts, int64_t M, int64_t N, double alpha, double beta, T* C,
                          int64_t ldc) {
  Context& c = ctx();
  k_reduce_partials<T><<<(unsigned)cdiv(M * N, 8), 256, 0, c.stream>>>(part, nparts, M * N, (T)alpha, (T)beta, C,
                                                                       M, ldc);
  AG_LAUNCHED();
  return AG_OK;
}
template ag_status reduce_partials<float>(const float*, int, int64_t, int64_t, double, double, float*, int64_t);
template ag_status reduce_partials<double>(const double*, int, int64_t, int64_t, double, double, double*, int64_t);

// Float32 streaming forms; *done stays false when the shape / alignment is not theirs.
static ag_status skinny_f32(int transA, int transB, int64_t M, int64_t N, int64_t K, double alpha, const float* A,
                            int64_t lda, const float* B, int64_t ldb, double beta, float* C, int64_t ldc,
                            bool* done) {
  Context& c = ctx();
  static const bool old_only = getenv("AGB_SKINNY_OLD") != nullptr;   // experiment knob, read once
  if (old_only) return AG_OK;
  const bool b16 = (((uintptr_t)B) & 15) == 0 && (ldb & 3) == 0;
  if (!transB && beta == 0.0 && N >= 512 && N < ((int64_t)1 << 35)) {
    if (M <= 16 && K >= 16 && K <= SK_MAXD && (K & 3) == 0 && b16 &&
        (!transA || ((lda & 1) == 0 && (((uintptr_t)A) & 7) == 0))) {
      const unsigned grid = (unsigned)cdiv(cdiv(N, 32), S1A_WARPS);
#define S1A(MB_) k_skinny_s1a<MB_><<<grid, S1A_WARPS * 32, 0, c.stream>>>(transA, (int)M, N, (int)K, (float)alpha, A, lda, B, ldb, C, ldc)
      if (M <= 4) S1A(4); else if (M <= 8) S1A(8); else if (M <= 12) S1A(12); else S1A(16);
#undef S1A
      AG_LAUNCHED();
      *done = true;
      return AG_OK;
    }
    if (M > 16 && M <= SK_MAXD && K <= 16) {
      const unsigned grid = (unsigned)cdiv(cdiv(N, 32), S1B_WARPS);
#define S1B(KB_) k_skinny_s1b<KB_><<<grid, S1B_WARPS * 32, 0, c.stream>>>(transA, (int)M, N, (int)K, (float)alpha, A, lda, B, ldb, C, ldc)
      if (K <= 4) S1B(4); else if (K <= 8) S1B(8); else if (K <= 12) S1B(12); else S1B(16);
#undef S1B
      AG_LAUNCHED();
      *done = true;
      return AG_OK;
    }
  }
  if (!transA && transB && M <= 16 && N <= SK_MAXD && (N & 3) == 0 && b16 && K >= 2048) {
    const int64_t warps = (int64_t)c.num_sms * S2_WARPS;
    const int64_t kpw = cdiv(cdiv(K, warps), S2_CK) * S2_CK;
    const int64_t nblk = cdiv(K, kpw * S2_WARPS);
    float* part = (float*)workspace((size_t)(nblk * M * N) * sizeof(float));
    if (!part) return AG_ENOMEM;
#define S2(MB_)                                                                                              \
  do {                                                                                                       \
    constexpr int smem = S2_WARPS * S2_DEPTH * (S2_CK * 64 + S2_CK * MB_ * 2) * (int)sizeof(float);              \
    static bool attr_set = false;                                                                            \
    if (!attr_set) {                                                                                         \
      AG_CUDA(cudaFuncSetAttribute(k_skinny_s2<MB_>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));    \
      attr_set = true;                                                                                       \
    }                                                                                                        \
    k_skinny_s2<MB_><<<(unsigned)nblk, S2_WARPS * 32, smem, c.stream>>>((int)M, (int)N, K, kpw, A, lda, B, ldb, part); \
  } while (0)
    if (M <= 4) S2(4); else if (M <= 8) S2(8); else if (M <= 12) S2(12); else S2(16);
#undef S2
    AG_LAUNCHED();
    *done = true;
    return reduce_partials<float>(part, (int)nblk, M, N, alpha, beta, C, ldc);
  }
  return AG_OK;
}

template <typename T>
static ag_status skinny_impl(int transA, int transB, int64_t M, int64_t N, int64_t K, double alpha, const T* A,
                             int64_t lda, const T* B, int64_t ldb, double beta, T* C, int64_t ldc, bool* done) {
  Context& c = ctx();
  *done = false;
  if (std::is_same<T, float>::value) {
    ag_status st = skinny_f32(transA, transB, M, N, K, alpha, (const float*)A, lda, (const float*)B, ldb, beta,
                              (float*)C, ldc, done);
    if (st != AG_OK || *done) return st;
  }
  if (!transB && M <= SK_MAXD && K <= SK_MAXD && N >= 512 && cdiv(N, 128) < ((int64_t)1 << 31)) {
    {
      const unsigned gridc = (unsigned)cdiv(N, 128);
      if (M <= 4) k_skinny_cols<T, 4><<<gridc, 128, 0, c.stream>>>(transA, (int)M, N, (int)K, (T)alpha, A, lda, B, ldb, (T)beta, C, ldc);
      else if (M <= 8) k_skinny_cols<T, 8><<<gridc, 128, 0, c.stream>>>(transA, (int)M, N, (int)K, (T)alpha, A, lda, B, ldb, (T)beta, C, ldc);
      else if (M <= 12) k_skinny_cols<T, 12><<<gridc, 128, 0, c.stream>>>(transA, (int)M, N, (int)K, (T)alpha, A, lda, B, ldb, (T)beta, C, ldc);
      else if (M <= 16) k_skinny_cols<T, 16><<<gridc, 128, 0, c.stream>>>(transA, (int)M, N, (int)K, (T)alpha, A, lda, B, ldb, (T)beta, C, ldc);
      else if (M <= 32) k_skinny_cols<T, 32><<<gridc, 128, 0, c.stream>>>(transA, (int)M, N, (int)K, (T)alpha, A, lda, B, ldb, (T)beta, C, ldc);
      else k_skinny_cols<T, 64><<<gridc, 128, 0, c.stream>>>(transA, (int)M, N, (int)K, (T)alpha, A, lda, B, ldb, (T)beta, C, ldc);
      AG_LAUNCHED();
      *done = true;
      return AG_OK;
    }
  }
  if (!transA && transB && M <= SK_MAXD && N <= SK_MAXD && K >= 2048) {
    int64_t nblk = (M <= 16 ? 8 : 2) * (int64_t)c.num_sms;
    int64_t kc = cdiv(cdiv(K, nblk), 32) * 32;
    nblk = cdiv(K, kc);
    T* part = (T*)workspace((size_t)(nblk * M * N) * sizeof(T));
    if (!part) return AG_ENOMEM;
    if (M <= 16) {
      if (M <= 4) k_skinny_nt_warp<T, 4><<<(unsigned)nblk, 128, 0, c.stream>>>((int)M, (int)N, K, kc, A, lda, B, ldb, part);
      else if (M <= 8) k_skinny_nt_warp<T, 8><<<(unsigned)nblk, 128, 0, c.stream>>>((int)M, (int)N, K, kc, A, lda, B, ldb, part);
      else if (M <= 12) k_skinny_nt_warp<T, 12><<<(unsigned)nblk, 128, 0, c.stream>>>((int)M, (int)N, K, kc, A, lda, B, ldb, part);
      else k_skinny_nt_warp<T, 16><<<(unsigned)nblk, 128, 0, c.stream>>>((int)M, (int)N, K, kc, A, lda, B, ldb, part);
      AG_LAUNCHED();
      *done = true;
      return reduce_partials<T>(part, (int)nblk, M, N, alpha, beta, C, ldc);
    }
#define SK_NT(NJ_) k_skinny_nt<T, NJ_><<<(unsigned)nblk, SK_T, 0, c.stream>>>((int)M, (int)N, K, kc, A, lda, B, ldb, part)
    const int nj = (int)cdiv(M, 4);
    if (nj <= 1) SK_NT(1); else if (nj <= 2) SK_NT(2); else if (nj <= 3) SK_NT(3);
    else if (nj <= 4) SK_NT(4); else if (nj <= 8) SK_NT(8); else if (nj <= 12) SK_NT(12); else SK_NT(16);
#undef SK_NT
    AG_LAUNCHED();
    *done = true;
    return reduce_partials<T>(part, (int)nblk, M, N, alpha, beta, C, ldc);
  }
  return AG_OK;
}

// The shapes skinny_impl takes (kept in step with its two branches): a product with two extents <= 64.
bool gemm_skinny_takes(int dtype, int transA, int transB, int64_t M, int64_t N, int64_t K) {
  (void)dtype;
  if (!transB && M <= SK_MAXD && K <= SK_MAXD && N >= 512) return true;
  if (!transA && transB && M <= SK_MAXD && N <= SK_MAXD && K >= 2048) return true;
  return false;
}

ag_status gemm_skinny(int dtype, int transA, int transB, int64_t M, int64_t N, int64_t K, double alpha,
                      const void* A, int64_t lda, const void* B, int64_t ldb, double beta, void* C, int64_t ldc,
                      bool* done) {
  if (dtype == AG_F32)
    return skinny_impl<float>(transA, transB, M, N, K, alpha, (const float*)A, lda, (const float*)B, ldb, beta,
                              (float*)C, ldc, done);
  return skinny_impl<double>(transA, transB, M, N, K, alpha, (const double*)A, lda, (const double*)B, ldb, beta,
                             (double*)C, ldc, done);
}

}  // namespace agb
