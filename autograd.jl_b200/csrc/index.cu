// index.cu -- getindex / Sparse / addtoindex! / cat / axpy! on device.
//
// Reference: getindex forward (src/getindex.jl:32), ungetindex -> Sparse (src/getindex.jl:61-79),
// addto!(dense, Sparse) and full() (src/addto.jl:91-98, src/sparse.jl:57) which run the sequential
// scalar loop addtoindex! (src/addto.jl:107-129); cat/uncat (src/cat.jl:27-31,46-67);
// axpy!/lmul! (src/sparse.jl:50-52, README.md:70-77).
//
// Scatter-add is deterministic: stable LSD radix sort of (destination, source position) pairs, then
// a segmented reduce that sums each destination's contributions in source order and performs ONE
// read-modify-write per touched destination.  Index bookkeeping (which destination each value
// goes to) is bit-exact with the reference; only the floating-point association order inside a
// destination can differ from the reference's interleaved sequential order.
#include "common.cuh"

#include <cub/device/device_radix_sort.cuh>

namespace agb {

template <typename T>
__global__ void __launch_bounds__(256)
    k_gather(const T* __restrict__ x, const int64_t* __restrict__ idx, T* __restrict__ out, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = x[idx[i]];
}

// out[r, j] = x[r, col[j]]
template <typename T>
__global__ void __launch_bounds__(256)
    k_gather_cols(const T* __restrict__ x, int64_t rows, const int64_t* __restrict__ col,
                  T* __restrict__ out, int64_t total) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int64_t j = i / rows, r = i - j * rows;
  out[i] = x[r + col[j] * rows];
}

__global__ void __launch_bounds__(256) k_iota(int32_t* __restrict__ p, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = (int32_t)i;
}

// Validate indices: flag[0] |= 1 if any key is out of [0, nkeys)
__global__ void __launch_bounds__(256)
    k_check_idx(const int64_t* __restrict__ idx, int64_t n, int64_t nkeys, int32_t* __restrict__ flag) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && (idx[i] < 0 || idx[i] >= nkeys)) *flag = 1;
}

// One thread-group per (segment head, row): sums the segment's values in source order.
// keys sorted ascending, pos = source positions (ascending inside a key: stable sort).
// thread (j, b): if j is a segment head, dest[key*block + b] += sum_{members} vals[pos*block + b]
template <typename T>
__global__ void __launch_bounds__(256)
    k_segreduce(const int64_t* __restrict__ keys, const int32_t* __restrict__ pos,
                const T* __restrict__ vals, int64_t n, int64_t block, int64_t nkeys, T* __restrict__ dest) {
  // blockIdx.x walks sorted entries in groups; threads along the value block for coalescing
  const int lanes = block >= 256 ? 256 : (block >= 128 ? 128 : (block >= 64 ? 64 : (block >= 32 ? 32 : (block >= 16 ? 16 : (block >= 8 ? 8 : (block >= 4 ? 4 : (block >= 2 ? 2 : 1)))))));
  const int per_blk = 256 / lanes;  // entries handled concurrently by one CUDA block
  const int lane = threadIdx.x % lanes, slot = threadIdx.x / lanes;
  const int64_t j = (int64_t)blockIdx.x * per_blk + slot;
  if (j >= n) return;
  const int64_t key = keys[j];
  if (key < 0 || key >= nkeys) return;      // out of range: flagged by k_check_idx, never written
  if (j > 0 && keys[j - 1] == key) return;  // not a head
  int64_t e = j + 1;
  while (e < n && keys[e] == key) ++e;
  for (int64_t b = lane; b < block; b += lanes) {
    T s = T(0);
    for (int64_t m = j; m < e; ++m) s += vals[(int64_t)pos[m] * block + b];
    dest[key * block + b] += s;
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
    k_slab_copy(const T* __restrict__ src, int64_t src_red, int64_t src_off, T* __restrict__ dst,
                int64_t dst_red, int64_t dst_off, int64_t inner, int64_t len, int64_t total) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int64_t slab = inner * len;
  int64_t o = i / slab, w = i - o * slab;  // w = ii + inner*r
  dst[o * inner * dst_red + dst_off * inner + w] = src[o * inner * src_red + src_off * inner + w];
}

template <typename T>
__global__ void __launch_bounds__(256)
    k_transpose(const T* __restrict__ x, int64_t rows, int64_t cols, T* __restrict__ out) {
  __shared__ T tile[32][33];
  int64_t r0 = (int64_t)blockIdx.x * 32, c0 = (int64_t)blockIdx.y * 32;
  int tx = threadIdx.x % 32, ty = threadIdx.x / 32;  // 32 x 8
  for (int k = ty; k < 32; k += 8) {
    int64_t r = r0 + tx, c = c0 + k;
    if (r < rows && c < cols) tile[k][tx] = x[r + c * rows];
  }
  __syncthreads();
  for (int k = ty; k < 32; k += 8) {
    int64_t c = c0 + tx, r = r0 + k;
    if (r < rows && c < cols) out[c + r * cols] = tile[tx][k];
  }
}

// out = permutedims(x, perm) for <= 4 dims: thread per output element (out is written coalesced)
struct PermArgs {
  int64_t oshape[4];   // shape of out
  int64_t xstride[4];  // stride in x of each out dimension
  int64_t n;
};
template <typename T>
__global__ void __launch_bounds__(256) k_permute(const T* __restrict__ x, T* __restrict__ out, const PermArgs a) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  int64_t r = i, off = 0;
#pragma unroll
  for (int d = 0; d < 4; ++d) {
    const int64_t q = r / a.oshape[d];
    off += (r - q * a.oshape[d]) * a.xstride[d];
    r = q;
  }
  out[i] = x[off];
}

template <typename T>
__global__ void __launch_bounds__(256)
    k_axpy(T alpha, const T* __restrict__ x, T* __restrict__ y, int64_t n) {
  int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 4 <= n && (((uintptr_t)(x + i) | (uintptr_t)(y + i)) & (4 * sizeof(T) - 1)) == 0 &&
      sizeof(T) == 4) {
    float4 a = *reinterpret_cast<const float4*>(x + i);
    float4 b = *reinterpret_cast<float4*>(y + i);
    b.x = fmaf((float)alpha, a.x, b.x);
    b.y = fmaf((float)alpha, a.y, b.y);
    b.z = fmaf((float)alpha, a.z, b.z);
    b.w = fmaf((float)alpha, a.w, b.w);
    *reinterpret_cast<float4*>(y + i) = b;
  } else {
    for (int k = 0; k < 4 && i + k < n; ++k) y[i + k] = fma(alpha, x[i + k], y[i + k]);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) k_scale(T alpha, T* __restrict__ x, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] *= alpha;
}

template <typename T>
static ag_status scatter_add_impl(T* dest, int64_t dest_len, const int64_t* idx, const T* vals,
                                  int64_t n, int64_t block) {
  Context& c = ctx();
  AG_CHECK(n < ((int64_t)1 << 31), AG_EUNSUPPORTED, "ag_scatter_add: n=%lld too large", (long long)n);
  int64_t nkeys = dest_len / block;
  int end_bit = 1;
  while (end_bit < 63 && ((int64_t)1 << end_bit) < nkeys) ++end_bit;
  size_t tmp_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, (const int64_t*)nullptr, (int64_t*)nullptr,
                                  (const int32_t*)nullptr, (int32_t*)nullptr, (int)n, 0, end_bit,
                                  c.stream);
  size_t off_keys = 256;
  size_t off_pos_in = off_keys + (((size_t)n * 8 + 255) & ~size_t(255));
  size_t off_pos_out = off_pos_in + (((size_t)n * 4 + 255) & ~size_t(255));
  size_t off_tmp = off_pos_out + (((size_t)n * 4 + 255) & ~size_t(255));
  size_t total = off_tmp + tmp_bytes;
  char* ws = (char*)workspace(total);
  if (!ws) return AG_ENOMEM;
  int32_t* flag = oob_flag();
  if (!flag) return AG_ENOMEM;
  int64_t* keys_out = (int64_t*)(ws + off_keys);
  int32_t* pos_in = (int32_t*)(ws + off_pos_in);
  int32_t* pos_out = (int32_t*)(ws + off_pos_out);
  void* tmp = ws + off_tmp;
  unsigned nb = (unsigned)cdiv(n, 256);
  k_check_idx<<<nb, 256, 0, c.stream>>>(idx, n, nkeys, flag);
  AG_LAUNCHED();
  k_iota<<<nb, 256, 0, c.stream>>>(pos_in, n);
  AG_LAUNCHED();
  AG_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, idx, keys_out, (const int32_t*)pos_in,
                                          pos_out, (int)n, 0, end_bit, c.stream));
  ctx().launches += 1;
  // out-of-range destinations are a BoundsError in the reference.  The host wrapper checks the bounds of the
  // index vectors it normalised; the device check is sticky and surfaces at the next ag_sync (no per-call
  // host synchronisation, so a whole step stays capturable); out-of-range entries are skipped, never written.
  int lanes = block >= 256 ? 256 : 1;
  if (block < 256) {
    lanes = 1;
    while (lanes * 2 <= block) lanes *= 2;
  }
  int per_blk = 256 / lanes;
  k_segreduce<T><<<(unsigned)cdiv(n, per_blk), 256, 0, c.stream>>>(keys_out, pos_out, vals, n, block, nkeys, dest);
  AG_LAUNCHED();
  return AG_OK;
}

}  // namespace agb

using namespace agb;

extern "C" {

ag_status ag_gather(int32_t dtype, const void* x, const int64_t* idx, void* out, int64_t n) {
  AG_REQUIRE_INIT();
  if (n == 0) return AG_OK;
  AG_CHECK(x && idx && out && n > 0, AG_EINVAL, "ag_gather: bad arguments");
  Context& c = ctx();
  unsigned nb = (unsigned)cdiv(n, 256);
  if (dtype == AG_F32) k_gather<float><<<nb, 256, 0, c.stream>>>((const float*)x, idx, (float*)out, n);
  else if (dtype == AG_F64) k_gather<double><<<nb, 256, 0, c.stream>>>((const double*)x, idx, (double*)out, n);
  else { set_error("ag_gather: dtype %d", dtype); return AG_EUNSUPPORTED; }
  AG_LAUNCHED();
  return AG_OK;
}

ag_status ag_gather_cols(int32_t dtype, const void* x, int64_t rows, const int64_t* col, void* out,
                         int64_t ncols_out) {
  AG_REQUIRE_INIT();
  int64_t total = rows * ncols_out;
  if (total == 0) return AG_OK;
  AG_CHECK(x && col && out && rows > 0 && ncols_out > 0, AG_EINVAL, "ag_gather_cols: bad arguments");
  Context& c = ctx();
  unsigned nb = (unsigned)cdiv(total, 256);
  if (dtype == AG_F32) k_gather_cols<float><<<nb, 256, 0, c.stream>>>((const float*)x, rows, col, (float*)out, total);
  else if (dtype == AG_F64) k_gather_cols<double><<<nb, 256, 0, c.stream>>>((const double*)x, rows, col, (double*)out, total);
  else { set_error("ag_gather_cols: dtype %d", dtype); return AG_EUNSUPPORTED; }
  AG_LAUNCHED();
  return AG_OK;
}

ag_status ag_scatter_add(int32_t dtype, void* dest, int64_t dest_len, const int64_t* idx,
                         const void* vals, int64_t n, int64_t block) {
  AG_REQUIRE_INIT();
  if (n == 0) return AG_OK;
  AG_CHECK(dest && idx && vals && n > 0 && block >= 1 && dest_len >= 0 && dest_len % block == 0,
           AG_EINVAL, "ag_scatter_add: bad arguments");
  if (dtype == AG_F32) return scatter_add_impl<float>((float*)dest, dest_len, idx, (const float*)vals, n, block);
  if (dtype == AG_F64) return scatter_add_impl<double>((double*)dest, dest_len, idx, (const double*)vals, n, block);
  set_error("ag_scatter_add: dtype %d", dtype);
  return AG_EUNSUPPORTED;
}

ag_status ag_slab_copy(int32_t dtype, const void* src, int64_t src_red, int64_t src_off, void* dst,
                       int64_t dst_red, int64_t dst_off, int64_t inner, int64_t len, int64_t outer) {
  AG_REQUIRE_INIT();
  int64_t total = inner * len * outer;
  if (total == 0) return AG_OK;
  AG_CHECK(src && dst && inner > 0 && len > 0 && outer > 0 && src_off >= 0 && dst_off >= 0 &&
               src_off + len <= src_red && dst_off + len <= dst_red,
           AG_EINVAL, "ag_slab_copy: slab out of range (BoundsError)");
  Context& c = ctx();
  unsigned nb = (unsigned)cdiv(total, 256);
  if (dtype == AG_F32)
    k_slab_copy<float><<<nb, 256, 0, c.stream>>>((const float*)src, src_red, src_off, (float*)dst, dst_red, dst_off, inner, len, total);
  else if (dtype == AG_F64)
    k_slab_copy<double><<<nb, 256, 0, c.stream>>>((const double*)src, src_red, src_off, (double*)dst, dst_red, dst_off, inner, len, total);
  else { set_error("ag_slab_copy: dtype %d", dtype); return AG_EUNSUPPORTED; }
  AG_LAUNCHED();
  return AG_OK;
}

ag_status ag_transpose(int32_t dtype, const void* x, int64_t rows, int64_t cols, void* out) {
  AG_REQUIRE_INIT();
  if (rows * cols == 0) return AG_OK;
  AG_CHECK(x && out && rows > 0 && cols > 0, AG_EINVAL, "ag_transpose: bad arguments");
  AG_CHECK(cdiv(cols, 32) <= 65535, AG_EUNSUPPORTED, "ag_transpose: too many columns");
  Context& c = ctx();
  dim3 grid((unsigned)cdiv(rows, 32), (unsigned)cdiv(cols, 32));
  if (dtype == AG_F32) k_transpose<float><<<grid, 256, 0, c.stream>>>((const float*)x, rows, cols, (float*)out);
  else if (dtype == AG_F64) k_transpose<double><<<grid, 256, 0, c.stream>>>((const double*)x, rows, cols, (double*)out);
  else { set_error("ag_transpose: dtype %d", dtype); return AG_EUNSUPPORTED; }
  AG_LAUNCHED();
  return AG_OK;
}

ag_status ag_permute(int32_t dtype, const void* x, int32_t ndims, const int64_t* shape, const int32_t* perm,
                     void* out) {
  AG_REQUIRE_INIT();
  AG_CHECK(ndims >= 1 && ndims <= 4 && shape && perm, AG_EUNSUPPORTED, "ag_permute: ndims=%d (max 4)", ndims);
  PermArgs a;
  int64_t xs[4] = {1, 1, 1, 1}, n = 1;
  bool seen[4] = {false, false, false, false};
  for (int d = 0; d < ndims; ++d) {
    AG_CHECK(shape[d] >= 0, AG_EINVAL, "ag_permute: negative extent");
    xs[d] = n;
    n *= shape[d];
  }
  for (int d = 0; d < 4; ++d) {
    if (d < ndims) {
      AG_CHECK(perm[d] >= 0 && perm[d] < ndims && !seen[perm[d]], AG_EINVAL, "ag_permute: not a permutation");
      seen[perm[d]] = true;
      a.oshape[d] = shape[perm[d]];
      a.xstride[d] = xs[perm[d]];
    } else {
      a.oshape[d] = 1;
      a.xstride[d] = 0;
    }
  }
  a.n = n;
  if (n == 0) return AG_OK;
  AG_CHECK(x && out, AG_EINVAL, "ag_permute: null pointer");
  Context& c = ctx();
  unsigned nb = (unsigned)cdiv(n, 256);
  if (dtype == AG_F32) k_permute<float><<<nb, 256, 0, c.stream>>>((const float*)x, (float*)out, a);
  else if (dtype == AG_F64) k_permute<double><<<nb, 256, 0, c.stream>>>((const double*)x, (double*)out, a);
  else { set_error("ag_permute: dtype %d", dtype); return AG_EUNSUPPORTED; }
  AG_LAUNCHED();
  return AG_OK;
}

ag_status ag_axpy(int32_t dtype, double alpha, const void* x, void* y, int64_t n) {
  AG_REQUIRE_INIT();
  if (n == 0) return AG_OK;
  AG_CHECK(x && y && n > 0, AG_EINVAL, "ag_axpy: bad arguments");
  Context& c = ctx();
  unsigned nb = (unsigned)cdiv(n, 1024);
  if (dtype == AG_F32) k_axpy<float><<<nb, 256, 0, c.stream>>>((float)alpha, (const float*)x, (float*)y, n);
  else if (dtype == AG_F64) k_axpy<double><<<nb, 256, 0, c.stream>>>(alpha, (const double*)x, (double*)y, n);
  else { set_error("ag_axpy: dtype %d", dtype); return AG_EUNSUPPORTED; }
  AG_LAUNCHED();
  return AG_OK;
}

ag_status ag_scale(int32_t dtype, double alpha, void* x, int64_t n) {
  AG_REQUIRE_INIT();
  if (n == 0) return AG_OK;
  AG_CHECK(x && n > 0, AG_EINVAL, "ag_scale: bad arguments");
  Context& c = ctx();
  unsigned nb = (unsigned)cdiv(n, 256);
  if (dtype == AG_F32) k_scale<float><<<nb, 256, 0, c.stream>>>((float)alpha, (float*)x, n);
  else if (dtype == AG_F64) k_scale<double><<<nb, 256, 0, c.stream>>>(alpha, (double*)x, n);
  else { set_error("ag_scale: dtype %d", dtype); return AG_EUNSUPPORTED; }
  AG_LAUNCHED();
  return AG_OK;
}

}  // extern "C"
