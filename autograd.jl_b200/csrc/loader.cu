// loader.cu -- the input side of the path: minibatches in the dataset's own byte format.
//
// Reference: examples/mnist.jl:77-88 reads idx-ubyte files and builds, on the host,
//   xshape(a) = reshape(a ./ 255f0, 784, :)                       (pixels -> Float32 in [0,1])
//   yshape(a) = (a[a.==0] = 10; full(sparse(Vector{Int}(a), 1:length(a), 1f0)))   (labels -> one-hot (10,N))
// so every training step moves 4 bytes per pixel and 40 bytes per label to the device.  Here the batch crosses
// PCIe as bytes (1 per pixel, 1 per label) and the two conversions run on the device: same values bit for bit
// (IEEE division by 255, exact 0/1), a quarter of the host->device traffic.
// HBM-bound: algorithmic bytes n*(1 + s) for the pixels, ncols*(1 + nclass*s) for the labels.
#include "common.cuh"

namespace agb {

// 16 pixels per thread: one 128-bit load, four 128-bit (Float32) stores
template <typename T>
__global__ void __launch_bounds__(256)
    k_ubyte_to_float(const uint8_t* __restrict__ src, T* __restrict__ dst, int64_t n, T divisor, bool vec) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
  if (i >= n) return;
  if (vec && i + 16 <= n) {
    const uint4 p = *reinterpret_cast<const uint4*>(src + i);
    const uint32_t w[4] = {p.x, p.y, p.z, p.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      T v[4];
#pragma unroll
      for (int b = 0; b < 4; ++b) v[b] = (T)((w[k] >> (8 * b)) & 0xffu) / divisor;
      if (sizeof(T) == 4) {
        *reinterpret_cast<float4*>(dst + i + 4 * k) = make_float4((float)v[0], (float)v[1], (float)v[2], (float)v[3]);
      } else {
        *reinterpret_cast<double2*>(dst + i + 4 * k) = make_double2((double)v[0], (double)v[1]);
        *reinterpret_cast<double2*>(dst + i + 4 * k + 2) = make_double2((double)v[2], (double)v[3]);
      }
    }
  } else {
    for (int64_t j = i; j < n && j < i + 16; ++j) dst[j] = (T)src[j] / divisor;
  }
}

// one thread per column: writes the nclass entries of its column (contiguous), raises the sticky flag on a bad label
template <typename T>
__global__ void __launch_bounds__(256)
    k_onehot_labels(const uint8_t* __restrict__ labels, int64_t ncols, int nclass, int zero_to, T* __restrict__ dst,
                    int32_t* __restrict__ flag) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= ncols) return;
  int cls = labels[j];
  if (cls == 0) cls = zero_to;            // `a[a.==0] = 10`
  if (cls < 1 || cls > nclass) { *flag = 1; cls = 0; }
  T* __restrict__ col = dst + j * nclass;
  for (int r = 0; r < nclass; ++r) col[r] = (r + 1 == cls) ? T(1) : T(0);
}

}  // namespace agb

using namespace agb;

extern "C" {

ag_status ag_ubyte_to_float(int32_t dtype, const void* src, void* dst, int64_t n, double divisor) {
  AG_REQUIRE_INIT();
  if (n == 0) return AG_OK;
  AG_CHECK(src && dst && n > 0 && divisor != 0.0, AG_EINVAL, "ag_ubyte_to_float: bad arguments");
  AG_CHECK(dtype == AG_F32 || dtype == AG_F64, AG_EUNSUPPORTED, "ag_ubyte_to_float: dtype %d", dtype);
  Context& c = ctx();
  const bool vec = (((uintptr_t)src | (uintptr_t)dst) & 15) == 0;
  const unsigned nb = (unsigned)cdiv(cdiv(n, 16), 256);
  if (dtype == AG_F32)
    k_ubyte_to_float<float><<<nb, 256, 0, c.stream>>>((const uint8_t*)src, (float*)dst, n, (float)divisor, vec);
  else
    k_ubyte_to_float<double><<<nb, 256, 0, c.stream>>>((const uint8_t*)src, (double*)dst, n, divisor, vec);
  AG_LAUNCHED();
  return AG_OK;
}

ag_status ag_onehot_labels(int32_t dtype, const void* labels, int64_t ncols, int32_t nclass, int32_t zero_to,
                           void* dst) {
  AG_REQUIRE_INIT();
  if (ncols == 0) return AG_OK;
  AG_CHECK(labels && dst && ncols > 0 && nclass >= 1 && nclass <= 255, AG_EINVAL, "ag_onehot_labels: bad arguments");
  AG_CHECK(dtype == AG_F32 || dtype == AG_F64, AG_EUNSUPPORTED, "ag_onehot_labels: dtype %d", dtype);
  Context& c = ctx();
  const unsigned nb = (unsigned)cdiv(ncols, 256);
  if (dtype == AG_F32)
    k_onehot_labels<float><<<nb, 256, 0, c.stream>>>((const uint8_t*)labels, ncols, nclass, zero_to, (float*)dst, oob_flag());
  else
    k_onehot_labels<double><<<nb, 256, 0, c.stream>>>((const uint8_t*)labels, ncols, nclass, zero_to, (double*)dst, oob_flag());
  AG_LAUNCHED();
  return AG_OK;
}

}  // extern "C"
