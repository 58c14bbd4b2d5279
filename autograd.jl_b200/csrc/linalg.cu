// linalg.cu -- inverse and determinant of a small dense matrix (src/linearalgebra.jl:18,28,57-58: det, inv, logdet,
// logabsdet and the inv(x)' their rules read).  The reference calls LAPACK (getrf / getri) on the host; here one
// thread block runs Gauss-Jordan elimination with partial pivoting on a double-precision copy [A | I] in the
// workspace: the matrices these rules see are parameter-sized (a few hundred rows at most), the work is O(n^3) on one
// SM and latency-bound, so there is nothing to tile -- it exists so that the rules run without a host round trip of
// the matrix.  n <= 1024.
#include <algorithm>
#include <cstdint>
#include "common.cuh"

namespace agb {

constexpr int kLuThreads = 1024;

// W: n x 2n doubles, column-major (leading dimension n): columns 0..n-1 = A, n..2n-1 = I on entry; on exit columns
// n..2n-1 hold inv(A).  info[0] = det, info[1] = log|det|, info[2] = sign(det).
template <typename T>
__global__ void __launch_bounds__(kLuThreads) k_gauss_jordan(const T* __restrict__ A, int n, double* __restrict__ W,
                                                             T* __restrict__ inv, T* __restrict__ info) {
  __shared__ double s_val[kLuThreads];
  __shared__ int s_idx[kLuThreads];
  __shared__ double s_col[1024];
  __shared__ double s_det[3];
  const int tid = threadIdx.x, nt = blockDim.x;
  const int64_t nn = (int64_t)n * n;
  for (int64_t e = tid; e < nn; e += nt) {
    W[e] = (double)A[e];
    W[nn + e] = (e % n == e / n) ? 1.0 : 0.0;
  }
  if (tid == 0) { s_det[0] = 1.0; s_det[1] = 0.0; s_det[2] = 1.0; }
  __syncthreads();
  for (int k = 0; k < n; ++k) {
    // pivot: the largest |W[i,k]|, i >= k (lowest row on ties, like getrf)
    double best = -1.0;
    int bi = k;
    for (int i = k + tid; i < n; i += nt) {
      const double v = fabs(W[(int64_t)k * n + i]);
      if (v > best) { best = v; bi = i; }
    }
    s_val[tid] = best;
    s_idx[tid] = bi;
    __syncthreads();
    for (int s = nt / 2; s > 0; s >>= 1) {
      if (tid < s) {
        const double o = s_val[tid + s];
        const int oi = s_idx[tid + s];
        if (o > s_val[tid] || (o == s_val[tid] && oi < s_idx[tid])) { s_val[tid] = o; s_idx[tid] = oi; }
      }
      __syncthreads();
    }
    const int p = s_idx[0];
    __syncthreads();
    if (p != k) {
      for (int j = tid; j < 2 * n; j += nt) {
        const double t = W[(int64_t)j * n + k];
        W[(int64_t)j * n + k] = W[(int64_t)j * n + p];
        W[(int64_t)j * n + p] = t;
      }
    }
    __syncthreads();
    const double piv = W[(int64_t)k * n + k];
    if (tid == 0) {
      s_det[0] *= (p != k) ? -piv : piv;
      s_det[1] += log(fabs(piv));
      if ((piv < 0.0) != (p != k)) s_det[2] = -s_det[2];
      if (piv == 0.0) s_det[2] = 0.0;
    }
    for (int i = tid; i < n; i += nt) s_col[i] = (i == k) ? 0.0 : W[(int64_t)k * n + i];   // elimination factors
    __syncthreads();
    const double ip = 1.0 / piv;
    for (int j = tid; j < 2 * n; j += nt) W[(int64_t)j * n + k] *= ip;                      // scale the pivot row
    __syncthreads();
    // rows i != k: W[i,j] -= f_i * W[k,j] for the columns that are not yet (A side) or already (I side) non-trivial
    const int j0 = k, ncols = 2 * n - j0;
    for (int64_t e = tid; e < (int64_t)ncols * n; e += nt) {
      const int j = j0 + (int)(e / n), i = (int)(e % n);
      if (i != k) W[(int64_t)j * n + i] -= s_col[i] * W[(int64_t)j * n + k];
    }
    __syncthreads();
  }
  if (inv)
    for (int64_t e = tid; e < nn; e += nt) inv[e] = (T)W[nn + e];
  if (info && tid < 3) info[tid] = (T)s_det[tid];
}

// ---- argmax / argmin of the vectorised array (@zerograd in src/base.jl:60-61): first index of the extremum, a NaN wins
// (Julia's order), two phases: one candidate per block, then one block over the candidates.
struct Cand { double v; int64_t i; };

__device__ __forceinline__ bool cand_better(const Cand& a, const Cand& b, bool want_max) {   // is a preferred over b?
  if (b.i < 0) return a.i >= 0;
  if (a.i < 0) return false;
  const bool an = a.v != a.v, bn = b.v != b.v;
  if (an || bn) return an && (!bn || a.i < b.i);
  if (a.v == b.v) return a.i < b.i;
  return want_max ? a.v > b.v : a.v < b.v;
}

__device__ __forceinline__ Cand cand_block_reduce(Cand c, bool want_max) {
  __shared__ Cand s[256];
  s[threadIdx.x] = c;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if (threadIdx.x < w && cand_better(s[threadIdx.x + w], s[threadIdx.x], want_max)) s[threadIdx.x] = s[threadIdx.x + w];
    __syncthreads();
  }
  return s[0];
}

template <typename T>
__global__ void __launch_bounds__(256) k_argext_blocks(const T* __restrict__ x, int64_t n, bool want_max, Cand* __restrict__ out) {
  Cand best = {0.0, -1};
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const Cand c = {(double)x[i], i};
    if (cand_better(c, best, want_max)) best = c;
  }
  best = cand_block_reduce(best, want_max);
  if (threadIdx.x == 0) out[blockIdx.x] = best;
}

__global__ void __launch_bounds__(256) k_argext_final(const Cand* __restrict__ in, int nb, bool want_max, int64_t* __restrict__ idx) {
  Cand best = {0.0, -1};
  for (int b = threadIdx.x; b < nb; b += 256)
    if (cand_better(in[b], best, want_max)) best = in[b];
  best = cand_block_reduce(best, want_max);
  if (threadIdx.x == 0) *idx = best.i;
}

// ---- eltype conversion Float32 <-> Float64 (oftype / big / convert: src/base.jl:76,365): four elements per thread,
// 128-bit accesses on the Float32 side, 2 x 128-bit on the Float64 side; HBM-bound, 12 bytes per element.
template <typename S, typename D>
__global__ void __launch_bounds__(256) k_cast(const S* __restrict__ x, D* __restrict__ y, int64_t n) {
  const int64_t i = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4;
  if (i >= n) return;
  if (i + 4 <= n) {
    S v[4];
    if constexpr (sizeof(S) == 4) {
      *reinterpret_cast<float4*>(v) = *reinterpret_cast<const float4*>(x + i);
    } else {
      *reinterpret_cast<double2*>(v) = *reinterpret_cast<const double2*>(x + i);
      *reinterpret_cast<double2*>(v + 2) = *reinterpret_cast<const double2*>(x + i + 2);
    }
    D o[4] = {(D)v[0], (D)v[1], (D)v[2], (D)v[3]};
    if constexpr (sizeof(D) == 4) {
      *reinterpret_cast<float4*>(y + i) = *reinterpret_cast<float4*>(o);
    } else {
      *reinterpret_cast<double2*>(y + i) = *reinterpret_cast<double2*>(o);
      *reinterpret_cast<double2*>(y + i + 2) = *reinterpret_cast<double2*>(o + 2);
    }
  } else {
    for (int64_t j = i; j < n; ++j) y[j] = (D)x[j];
  }
}

}  // namespace agb

using namespace agb;

extern "C" ag_status ag_cast(int32_t src_dtype, int32_t dst_dtype, const void* x, void* y, int64_t n) {
  AG_REQUIRE_INIT();
  if (n == 0) return AG_OK;
  AG_CHECK(x && y && n > 0, AG_EINVAL, "ag_cast: bad arguments");
  AG_CHECK((src_dtype == AG_F32 && dst_dtype == AG_F64) || (src_dtype == AG_F64 && dst_dtype == AG_F32), AG_EUNSUPPORTED,
           "ag_cast: %d -> %d (Float32 <-> Float64 only)", src_dtype, dst_dtype);
  AG_CHECK(((uintptr_t)x % 16 == 0) && ((uintptr_t)y % 16 == 0), AG_EINVAL, "ag_cast: buffers must be 16-byte aligned");
  Context& c = ctx();
  const unsigned nb = (unsigned)cdiv(n, (int64_t)1024);
  if (src_dtype == AG_F32) k_cast<float, double><<<nb, 256, 0, c.stream>>>((const float*)x, (double*)y, n);
  else k_cast<double, float><<<nb, 256, 0, c.stream>>>((const double*)x, (float*)y, n);
  AG_LAUNCHED();
  return AG_OK;
}

extern "C" ag_status ag_argext(int32_t dtype, const void* x, int64_t n, int32_t want_max, int64_t* idx_out) {
  AG_REQUIRE_INIT();
  AG_CHECK(dtype == AG_F32 || dtype == AG_F64, AG_EUNSUPPORTED, "ag_argext: dtype %d", dtype);
  AG_CHECK(x && idx_out && n >= 1, AG_EINVAL, "ag_argext: bad arguments (an empty collection has no extremum)");
  Context& c = ctx();
  const int nb = (int)std::min<int64_t>(cdiv(n, (int64_t)256), 148 * 8);
  Cand* part = (Cand*)workspace((size_t)nb * sizeof(Cand));
  AG_CHECK(part, AG_ENOMEM, "ag_argext: workspace");
  if (dtype == AG_F32) k_argext_blocks<float><<<nb, 256, 0, c.stream>>>((const float*)x, n, want_max != 0, part);
  else k_argext_blocks<double><<<nb, 256, 0, c.stream>>>((const double*)x, n, want_max != 0, part);
  k_argext_final<<<1, 256, 0, c.stream>>>(part, nb, want_max != 0, idx_out);
  AG_LAUNCHED();
  return AG_OK;
}

extern "C" ag_status ag_inv_det(int32_t dtype, const void* a, int64_t n, void* inv_out, void* info_out) {
  AG_REQUIRE_INIT();
  AG_CHECK(dtype == AG_F32 || dtype == AG_F64, AG_EUNSUPPORTED, "ag_inv_det: dtype %d", dtype);
  AG_CHECK(a && n >= 1 && (inv_out || info_out), AG_EINVAL, "ag_inv_det: bad arguments");
  AG_CHECK(n <= 1024, AG_EUNSUPPORTED, "ag_inv_det: n = %lld (one-block elimination, n <= 1024)", (long long)n);
  Context& c = ctx();
  double* W = (double*)workspace((size_t)(2 * n * n) * sizeof(double));
  AG_CHECK(W, AG_ENOMEM, "ag_inv_det: workspace");
  if (dtype == AG_F32)
    k_gauss_jordan<float><<<1, kLuThreads, 0, c.stream>>>((const float*)a, (int)n, W, (float*)inv_out, (float*)info_out);
  else
    k_gauss_jordan<double><<<1, kLuThreads, 0, c.stream>>>((const double*)a, (int)n, W, (double*)inv_out, (double*)info_out);
  AG_LAUNCHED();
  return AG_OK;
}
