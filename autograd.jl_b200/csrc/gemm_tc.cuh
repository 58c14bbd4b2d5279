// gemm_tc.cuh -- interface of the tcgen05 product (gemm_tc.cu) inside the library.
#pragma once
#include "common.cuh"

namespace agb {

enum { TC_ACT_NONE = 0, TC_ACT_RELU = 1, TC_ACT_TANH = 2, TC_ACT_SIGM = 3 };

// Fused epilogue of a product: z = alpha*A*B (+ beta*C) + bias (bias indexed by C's row, i.e. a column vector
// broadcast along dim 2: the `w*x .+ b` of examples/mnist.jl:32, examples/charlm.jl:74-77); with out2 == NULL the
// product's output receives act(z), otherwise it receives z and out2 receives act(z) (same layout and ldc).
struct TcEpilogue {
  const void* bias;
  int act;
  void* out2;
};

// One product of a group launch (gemm_tc_group): BLAS-style description + epilogue.
struct TcJob {
  int transA, transB;
  int64_t M, N, K;
  double alpha, beta;
  const void* A;
  int64_t lda;
  const void* B;
  int64_t ldb;
  void* C;
  int64_t ldc;
  TcEpilogue epi;
};

bool gemm_tc_same_class(const TcJob& a, const TcJob& b);     // may the two run in one group launch?
ag_status gemm_tc_group(int n, const TcJob* jobs);           // n <= 8 products of one class, ONE launch
bool gemm_tc_init();      // allocates the split counters; called from ag_init (never during a capture)
bool gemm_tc_supported(int transA, int transB, int64_t M, int64_t N, int64_t K, const void* A, int64_t lda,
                       const void* B, int64_t ldb, const void* C, int64_t ldc);
ag_status gemm_tc(int transA, int transB, int64_t M, int64_t N, int64_t K, double alpha, const void* A, int64_t lda,
                  const void* B, int64_t ldb, double beta, void* C, int64_t ldc, const TcEpilogue* epi);

}  // namespace agb
