// gemm_tc.cu -- tcgen05 (TMA + TMEM) Float32 GEMM engine.  Placeholder until the kernel lands:
// reports "not supported" so ag_gemm uses the CUDA-core engine.
#include "common.cuh"

namespace agb {

bool gemm_tc_supported(int, int, int64_t, int64_t, int64_t, const void*, int64_t, const void*,
                       int64_t, const void*, int64_t) {
  return false;
}

ag_status gemm_tc(int, int, int64_t, int64_t, int64_t, double, const void*, int64_t, const void*,
                  int64_t, double, void*, int64_t) {
  set_error("tcgen05 GEMM engine not built");
  return AG_EUNSUPPORTED;
}

}  // namespace agb
