// skinny.cu -- CUDA-core matrix products for the "tiny matrix x very wide matrix" shapes of the tape:
//   S1  C[MxN] = op(A)[MxK] * B[KxN],  M,K <= 64, N huge      forward  w2*a3  (10x64 * 64xB)  and the
//                                                              gradient w2'*dy (64x10 * 10xB), src/base.jl:26
//   S2  C[MxN] = A[MxK] * B[NxK]',     M,N <= 64, K huge      gradient dy*a3' (10xB * Bx64)
// These are HBM/L2-bound streaming passes over the wide operand (algorithmic bytes (MK+KN+MN)*s); the
// tiny operand lives in shared memory (S1: one thread per column of the wide operand; S2: one warp per k-range).  A 128-row tensor-core tile would waste >= 84 % of its rows on
// M = 10 and TMA cannot address a 40-byte leading dimension, so they stay on the FFMA/DFMA pipe.
// Reductions over K are deterministic (per-block partials, fixed-order second phase).
#include "common.cuh"

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

template <typename T>
static ag_status skinny_impl(int transA, int transB, int64_t M, int64_t N, int64_t K, double alpha, const T* A,
                             int64_t lda, const T* B, int64_t ldb, double beta, T* C, int64_t ldc, bool* done) {
  Context& c = ctx();
  *done = false;
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
