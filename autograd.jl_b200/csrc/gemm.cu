// gemm.cu -- C = alpha*op(A)*op(B) + beta*C, column-major BLAS convention.
//
// Reference: forward `*` (src/core.jl:66 -> LinearAlgebra/OpenBLAS) and the rules
// `@primitive1 *(x1,x2),dy (dy*x2') (x1'*dy)` (src/base.jl:26; lazy adjoint src/linearalgebra.jl:16),
// i.e. NN forward, NT for dA and TN for dB.
//
// Engines: "tcgen05" (gemm_tc.cu: TMA + tcgen05.mma kind::tf32, 3xTF32 error compensation, fp32
// TMEM accumulators, split-K) for qualifying Float32 shapes; "ffma"/"dfma" (this file): a
// shared-memory tiled CUDA-core kernel with deterministic split-K for everything else (Float64
// has no tcgen05 kind).
#include "common.cuh"
#include "gemm_tc.cuh"

#include <vector>

namespace agb {

ag_status gemm_skinny(int dtype, int transA, int transB, int64_t M, int64_t N, int64_t K, double alpha,
                      const void* A, int64_t lda, const void* B, int64_t ldb, double beta, void* C, int64_t ldc,
                      bool* done);
template <typename T>
ag_status reduce_partials(const T* part, int nparts, int64_t M, int64_t N, double alpha, double beta, T* C,
                          int64_t ldc);

constexpr int BM = 64, BN = 64, BK = 16, GT = 256, PADM = 4;

// Partial product over k in [k0,k1) of one 64x64 tile.  SPLIT: write raw partial sums to `ws`
// ([split][N][M] column-major tiles of the full MxN), else apply alpha/beta and write C.
template <typename T, bool SPLIT>
__global__ void __launch_bounds__(GT)
    k_gemm_simt(int transA, int transB, int64_t M, int64_t N, int64_t K, T alpha,
                const T* __restrict__ A, int64_t lda, const T* __restrict__ B, int64_t ldb, T beta,
                T* __restrict__ C, int64_t ldc, int64_t kchunk, T* __restrict__ ws) {
  __shared__ T As[BK][BM + PADM];
  __shared__ T Bs[BK][BN + PADM];
  const int t = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * BM, n0 = (int64_t)blockIdx.y * BN;
  const int64_t kbeg = (int64_t)blockIdx.z * kchunk;
  int64_t kend = kbeg + kchunk;
  if (kend > K) kend = K;
  const int tx = t % 16, ty = t / 16;  // thread computes rows tx*4.., cols ty*4..
  T acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = T(0);

  for (int64_t kk = kbeg; kk < kend; kk += BK) {
    // ---- stage A tile (BM x BK) as As[k][m]
    if (!transA) {  // A[m + k*lda]: m contiguous
      const int m = t % BM, kq = t / BM;  // kq in 0..3
#pragma unroll
      for (int j = 0; j < BK / 4; ++j) {
        int k = kq + 4 * j;
        int64_t gm = m0 + m, gk = kk + k;
        As[k][m] = (gm < M && gk < kend) ? A[gm + gk * lda] : T(0);
      }
    } else {  // A[k + m*lda]: k contiguous
      const int k = t % BK, mq = t / BK;  // mq in 0..15
#pragma unroll
      for (int j = 0; j < BM / 16; ++j) {
        int m = mq + 16 * j;
        int64_t gm = m0 + m, gk = kk + k;
        As[k][m] = (gm < M && gk < kend) ? A[gk + gm * lda] : T(0);
      }
    }
    // ---- stage B tile (BK x BN) as Bs[k][n]
    if (!transB) {  // B[k + n*ldb]: k contiguous
      const int k = t % BK, nq = t / BK;
#pragma unroll
      for (int j = 0; j < BN / 16; ++j) {
        int n = nq + 16 * j;
        int64_t gn = n0 + n, gk = kk + k;
        Bs[k][n] = (gn < N && gk < kend) ? B[gk + gn * ldb] : T(0);
      }
    } else {  // B[n + k*ldb]: n contiguous
      const int n = t % BN, kq = t / BN;
#pragma unroll
      for (int j = 0; j < BK / 4; ++j) {
        int k = kq + 4 * j;
        int64_t gn = n0 + n, gk = kk + k;
        Bs[k][n] = (gn < N && gk < kend) ? B[gn + gk * ldb] : T(0);
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      T a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][tx * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[k][ty * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int64_t gn = n0 + ty * 4 + j;
    if (gn >= N) continue;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int64_t gm = m0 + tx * 4 + i;
      if (gm >= M) continue;
      if (SPLIT) {
        ws[((int64_t)blockIdx.z * N + gn) * M + gm] = acc[i][j];
      } else {
        T* c = C + gm + gn * ldc;
        *c = beta == T(0) ? alpha * acc[i][j] : alpha * acc[i][j] + beta * (*c);
      }
    }
  }
}

template <typename T>
static ag_status gemm_simt(int transA, int transB, int64_t M, int64_t N, int64_t K, double alpha,
                           const void* A, int64_t lda, const void* B, int64_t ldb, double beta,
                           void* C, int64_t ldc) {
  Context& c = ctx();
  int64_t tm = cdiv(M, BM), tn = cdiv(N, BN);
  AG_CHECK(tn <= 65535, AG_EUNSUPPORTED, "ag_gemm: N=%lld too large for the CUDA-core engine",
           (long long)N);
  int64_t tiles = tm * tn;
  int64_t splits = 1;
  if (tiles < 2 * c.num_sms && K >= 512) {
    splits = cdiv(2 * (int64_t)c.num_sms, tiles);
    int64_t maxs = cdiv(K, 256);
    if (splits > maxs) splits = maxs;
    if (splits > 512) splits = 512;
  }
  int64_t kchunk = BK;
  if (K > 0) {
    kchunk = cdiv(cdiv(K, splits), BK) * BK;
    splits = cdiv(K, kchunk);
  } else {
    splits = 1;
  }
  dim3 grid((unsigned)tm, (unsigned)tn, (unsigned)splits);
  if (splits == 1) {
    k_gemm_simt<T, false><<<grid, GT, 0, c.stream>>>(transA, transB, M, N, K, (T)alpha, (const T*)A,
                                                     lda, (const T*)B, ldb, (T)beta, (T*)C, ldc,
                                                     kchunk, nullptr);
    AG_LAUNCHED();
    return AG_OK;
  }
  T* ws = (T*)workspace((size_t)(splits * M * N) * sizeof(T));
  if (!ws) return AG_ENOMEM;
  k_gemm_simt<T, true><<<grid, GT, 0, c.stream>>>(transA, transB, M, N, K, (T)alpha, (const T*)A, lda,
                                                  (const T*)B, ldb, (T)beta, (T*)C, ldc, kchunk, ws);
  AG_LAUNCHED();
  return reduce_partials<T>(ws, (int)splits, M, N, alpha, beta, (T*)C, ldc);
}

// ---- tiny outputs with a long contraction (housing.jl: dw = dy*x', 1 x 506 times 506 x 13) -------------------
// The tiled kernel above would spend a whole 64 x 64 tile of (double) FMAs on 13 outputs.  Here one warp owns one
// output element: lanes stride over k, fixed-order shuffle reduction (deterministic).
template <typename T>
__global__ void __launch_bounds__(256)
    k_gemm_warp_per_output(int transA, int transB, int64_t M, int64_t N, int64_t K, T alpha, const T* __restrict__ A,
                           int64_t lda, const T* __restrict__ B, int64_t ldb, T beta, T* __restrict__ C, int64_t ldc) {
  const int lane = threadIdx.x & 31;
  const int64_t o = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (o >= M * N) return;
  const int64_t m = o % M, n = o / M;
  const int64_t sa = transA ? 1 : lda, sb = transB ? ldb : 1;
  const T* __restrict__ a = A + (transA ? m * lda : m);
  const T* __restrict__ b = B + (transB ? n : n * ldb);
  T s0 = 0, s1 = 0;
  int64_t k = lane;
  for (; k + 32 < K; k += 64) {
    s0 = fma(a[k * sa], b[k * sb], s0);
    s1 = fma(a[(k + 32) * sa], b[(k + 32) * sb], s1);
  }
  if (k < K) s0 = fma(a[k * sa], b[k * sb], s0);
  T s = s0 + s1;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
  if (lane == 0) {
    T* c = C + m + n * ldc;
    *c = beta == T(0) ? alpha * s : alpha * s + beta * (*c);
  }
}

bool gemm_skinny_takes(int dtype, int transA, int transB, int64_t M, int64_t N, int64_t K);   // skinny.cu

static int g_engine = 0;
static const char* g_last_engine = "none";

}  // namespace agb

using namespace agb;

extern "C" {

ag_status ag_gemm_set_engine(int32_t engine) {
  AG_CHECK(engine >= 0 && engine <= 2, AG_EINVAL, "ag_gemm_set_engine: %d", engine);
  g_engine = engine;
  return AG_OK;
}

ag_status ag_gemm_last_engine(char* buf, int64_t buflen) {
  if (!buf || buflen <= 0) return AG_EINVAL;
  strncpy(buf, g_last_engine, (size_t)buflen - 1);
  buf[buflen - 1] = 0;
  return AG_OK;
}

ag_status ag_gemm(int32_t dtype, int32_t transA, int32_t transB, int64_t M, int64_t N, int64_t K,
                  double alpha, const void* A, int64_t lda, const void* B, int64_t ldb, double beta,
                  void* C, int64_t ldc) {
  AG_REQUIRE_INIT();
  AG_CHECK(dtype == AG_F32 || dtype == AG_F64, AG_EUNSUPPORTED, "ag_gemm: dtype %d", dtype);
  AG_CHECK(M >= 0 && N >= 0 && K >= 0, AG_EINVAL, "ag_gemm: negative dimension");
  AG_CHECK((transA == 0 || transA == 1) && (transB == 0 || transB == 1), AG_EINVAL,
           "ag_gemm: trans flags must be 0 or 1");
  if (M == 0 || N == 0) return AG_OK;
  AG_CHECK(C && ldc >= M, AG_EINVAL, "ag_gemm: bad C/ldc (DimensionMismatch)");
  if (K > 0) {
    AG_CHECK(A && B, AG_EINVAL, "ag_gemm: null operand");
    AG_CHECK(lda >= (transA ? K : M) && ldb >= (transB ? N : K), AG_EINVAL,
             "ag_gemm: leading dimension too small (DimensionMismatch)");
  }
  if (K > 0 && g_engine != 2) {
    bool done = false;
    ag_status st = gemm_skinny(dtype, transA, transB, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, &done);
    if (st != AG_OK) return st;
    if (done) {
      g_last_engine = "skinny";
      return AG_OK;
    }
  }
  if (K >= 32 && M * N <= 1024 && K * M * N <= ((int64_t)1 << 22)) {
    Context& c = ctx();
    const unsigned nb = (unsigned)cdiv(M * N, 8);
    if (dtype == AG_F64)
      k_gemm_warp_per_output<double><<<nb, 256, 0, c.stream>>>(transA, transB, M, N, K, alpha, (const double*)A, lda,
                                                               (const double*)B, ldb, beta, (double*)C, ldc);
    else
      k_gemm_warp_per_output<float><<<nb, 256, 0, c.stream>>>(transA, transB, M, N, K, (float)alpha, (const float*)A,
                                                              lda, (const float*)B, ldb, (float)beta, (float*)C, ldc);
    AG_LAUNCHED();
    g_last_engine = "warp-per-output";
    return AG_OK;
  }
  if (dtype == AG_F64) {
    g_last_engine = "dfma";
    return gemm_simt<double>(transA, transB, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
  }
  bool tc_ok = K > 0 && gemm_tc_supported(transA, transB, M, N, K, A, lda, B, ldb, C, ldc);
  if (g_engine == 2) {
    AG_CHECK(tc_ok, AG_EUNSUPPORTED, "ag_gemm: shape does not qualify for the tcgen05 engine");
  }
  if (tc_ok && g_engine != 1) {
    g_last_engine = "tcgen05";
    return gemm_tc(transA, transB, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, nullptr);
  }
  g_last_engine = "ffma";
  return gemm_simt<float>(transA, transB, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
}

// would ag_gemm_epilogue run this product on the tcgen05 engine (bias and activation in its epilogue)?
static bool epilogue_runs_fused(int32_t dtype, int32_t transA, int32_t transB, int64_t M, int64_t N, int64_t K,
                                const void* A, int64_t lda, const void* B, int64_t ldb, const void* C, const void* C2,
                                int64_t ldc) {
  if (!(dtype == AG_F32 && g_engine != 1 && K > 0 && ldc == M && (!C2 || ((uintptr_t)C2 & 15) == 0))) return false;
  if (K >= 32 && M * N <= 1024 && K * M * N <= ((int64_t)1 << 22)) return false;     // warp-per-output shapes
  if (!gemm_tc_supported(transA, transB, M, N, K, A, lda, B, ldb, C, ldc)) return false;
  if (g_engine != 2 && gemm_skinny_takes(dtype, transA, transB, M, N, K)) return false;   // those stay unfused
  return true;
}

/* *fused_out = 1 if ag_gemm_epilogue would run this product as one launch.  The host defers `w*x .+ b` only then;
 * otherwise the sum stays part of the elementwise chain that follows it. */
ag_status ag_gemm_epilogue_query(int32_t dtype, int32_t transA, int32_t transB, int64_t M, int64_t N, int64_t K,
                                 const void* A, int64_t lda, const void* B, int64_t ldb, int32_t* fused_out) {
  AG_REQUIRE_INIT();
  AG_CHECK(fused_out, AG_EINVAL, "ag_gemm_epilogue_query: null");
  // C / C2 come from the pool (512-byte aligned): a 16-byte aligned stand-in
  *fused_out = epilogue_runs_fused(dtype, transA, transB, M, N, K, A, lda, B, ldb, (const void*)16, nullptr, M) ? 1 : 0;
  return AG_OK;
}

static int act_to_tc(int act) {
  return act == AG_U_RELU ? TC_ACT_RELU : act == AG_U_TANH ? TC_ACT_TANH : act == AG_U_SIGM ? TC_ACT_SIGM : TC_ACT_NONE;
}

/* n independent products with their epilogues (each one as ag_gemm_epilogue).  Products the tcgen05 engine takes are
 * issued in groups of one class, up to 8 per launch; the others run one by one.  Same values either way. */
ag_status ag_gemm_group(int32_t dtype, int32_t n, const ag_gemm_desc* d) {
  AG_REQUIRE_INIT();
  AG_CHECK(n >= 0 && (n == 0 || d), AG_EINVAL, "ag_gemm_group: bad arguments");
  std::vector<char> done((size_t)n, 0);
  for (int i = 0; i < n; ++i) {
    if (done[i]) continue;
    const ag_gemm_desc& a = d[i];
    const bool ok_args = a.M > 0 && a.N > 0 && a.K > 0 && a.A && a.B && a.C && a.ldc == a.M &&
                         (a.act == AG_U_IDENTITY || a.act == AG_U_RELU || a.act == AG_U_TANH || a.act == AG_U_SIGM);
    if (!ok_args || !epilogue_runs_fused(dtype, a.transA, a.transB, a.M, a.N, a.K, a.A, a.lda, a.B, a.ldb, a.C, a.C2, a.ldc)) {
      ag_status st = ag_gemm_epilogue(dtype, a.transA, a.transB, a.M, a.N, a.K, a.A, a.lda, a.B, a.ldb, a.bias, a.act, a.C,
                                      a.C2, a.ldc);
      if (st != AG_OK) return st;
      done[i] = 1;
      continue;
    }
    TcJob jobs[8];
    int cnt = 0;
    for (int k = i; k < n && cnt < 8; ++k) {
      if (done[k]) continue;
      const ag_gemm_desc& b = d[k];
      if (k != i) {
        const bool okb = b.M > 0 && b.N > 0 && b.K > 0 && b.A && b.B && b.C && b.ldc == b.M &&
                         (b.act == AG_U_IDENTITY || b.act == AG_U_RELU || b.act == AG_U_TANH || b.act == AG_U_SIGM);
        if (!okb || !epilogue_runs_fused(dtype, b.transA, b.transB, b.M, b.N, b.K, b.A, b.lda, b.B, b.ldb, b.C, b.C2, b.ldc))
          continue;
      }
      TcJob j;
      j.transA = b.transA; j.transB = b.transB; j.M = b.M; j.N = b.N; j.K = b.K; j.alpha = 1.0; j.beta = 0.0;
      j.A = b.A; j.lda = b.lda; j.B = b.B; j.ldb = b.ldb; j.C = b.C; j.ldc = b.ldc;
      j.epi.bias = b.bias; j.epi.act = act_to_tc(b.act); j.epi.out2 = b.C2;
      if (cnt > 0 && !gemm_tc_same_class(jobs[0], j)) continue;
      jobs[cnt++] = j;
      done[k] = 1;
    }
    g_last_engine = cnt > 1 ? "tcgen05+epilogue (group)" : "tcgen05+epilogue";
    ag_status st = gemm_tc_group(cnt, jobs);
    if (st != AG_OK) return st;
  }
  return AG_OK;
}

/* C = op(A)*op(B) .+ bias (bias: M values, broadcast along dim 2), then optionally an activation: one launch when the
 * tcgen05 engine takes the shape (bias and activation in its epilogue), otherwise the same three kernels the unfused
 * tape would run (product, broadcast add, elementwise map) -- results are identical either way. */
ag_status ag_gemm_epilogue(int32_t dtype, int32_t transA, int32_t transB, int64_t M, int64_t N, int64_t K,
                           const void* A, int64_t lda, const void* B, int64_t ldb, const void* bias, int32_t act,
                           void* C, void* C2, int64_t ldc) {
  AG_REQUIRE_INIT();
  AG_CHECK(dtype == AG_F32 || dtype == AG_F64, AG_EUNSUPPORTED, "ag_gemm_epilogue: dtype %d", dtype);
  AG_CHECK(act == AG_U_IDENTITY || act == AG_U_RELU || act == AG_U_TANH || act == AG_U_SIGM, AG_EUNSUPPORTED,
           "ag_gemm_epilogue: activation %d (identity, relu, tanh, sigm)", act);
  AG_CHECK(M >= 0 && N >= 0 && K >= 0, AG_EINVAL, "ag_gemm_epilogue: negative dimension");
  AG_CHECK((transA == 0 || transA == 1) && (transB == 0 || transB == 1), AG_EINVAL,
           "ag_gemm_epilogue: trans flags must be 0 or 1");
  AG_CHECK(!(C2 && act == AG_U_IDENTITY), AG_EINVAL, "ag_gemm_epilogue: second output without an activation");
  if (M == 0 || N == 0) return AG_OK;
  AG_CHECK(C && ldc >= M, AG_EINVAL, "ag_gemm_epilogue: bad C/ldc (DimensionMismatch)");
  if (K > 0) {
    AG_CHECK(A && B, AG_EINVAL, "ag_gemm_epilogue: null operand");
    AG_CHECK(lda >= (transA ? K : M) && ldb >= (transB ? N : K), AG_EINVAL,
             "ag_gemm_epilogue: leading dimension too small (DimensionMismatch)");
  }
  if (epilogue_runs_fused(dtype, transA, transB, M, N, K, A, lda, B, ldb, C, C2, ldc)) {
    TcEpilogue e;
    e.bias = bias;
    e.act = act == AG_U_RELU ? TC_ACT_RELU : act == AG_U_TANH ? TC_ACT_TANH : act == AG_U_SIGM ? TC_ACT_SIGM : TC_ACT_NONE;
    e.out2 = C2;
    g_last_engine = "tcgen05+epilogue";
    return gemm_tc(transA, transB, M, N, K, 1.0, A, lda, B, ldb, 0.0, C, ldc, &e);
  }
  // unfused: product, then z = C .+ bias in place, then the activation (into C2, or in place)
  AG_CHECK(ldc == M, AG_EUNSUPPORTED, "ag_gemm_epilogue: ldc != M");
  ag_status st = ag_gemm(dtype, transA, transB, M, N, K, 1.0, A, lda, B, ldb, 0.0, C, ldc);
  if (st != AG_OK) return st;
  if (bias) {
    const int64_t shape[2] = {M, N}, sc[2] = {1, M}, sb[2] = {1, 0};
    st = ag_binary_fwd(AG_B_ADD, dtype, C, 0.0, sc, bias, 0.0, sb, C, 2, shape);
    if (st != AG_OK) return st;
  }
  if (act != AG_U_IDENTITY) {
    if (act == AG_U_RELU) {   // max.(0, z): the binary map, as the tape would run it
      const int64_t shape[1] = {M * N}, s1[1] = {1}, s0[1] = {0};
      st = ag_binary_fwd(AG_B_MAX, dtype, nullptr, 0.0, s0, C, 0.0, s1, C2 ? C2 : C, 1, shape);
    } else {
      st = ag_unary_fwd(act, dtype, C, C2 ? C2 : C, M * N);
    }
  }
  return st;
}

}  // extern "C"
