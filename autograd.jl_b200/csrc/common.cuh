// common.cuh -- shared runtime state and helpers of libagb200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>

#include "../../include/agb200.h"

namespace agb {

struct Context {
  bool ready = false;
  int device = -1;
  int num_sms = 148;
  cudaStream_t stream = nullptr;
  void* ws = nullptr;          // reduction / split-K workspace
  size_t ws_bytes = 0;
  int64_t launches = 0;        // kernels launched by this library
  bool capturing = false;
};

Context& ctx();
void set_error(const char* fmt, ...);
ag_status cuda_fail(cudaError_t e, const char* what, const char* file, int line);
void* workspace(size_t bytes);  // grows the shared workspace (not during graph capture)
int comm_nccl_async_error(char* msg, int msglen);   // comm.cu: ncclCommGetAsyncError of the communicator (0 = healthy)
int comm_take_timeout();       // comm.cu: a peer-memory wait expired since the last check (0 = none, else 1 + rank)
int32_t* oob_flag();            // sticky device flag raised by index checks; read and cleared by ag_sync

bool colprog_init();            // fuse.cu: ticket of the column programs (at ag_init)
bool reduce_scratch_init();     // reduce.cu: counters / second-level partials of the split combine (at ag_init)
ag_status reduce_double_partials(int dtype, const double* part, int64_t nchunks, int64_t inner, void* out);  // reduce.cu

static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace agb

#define AG_CUDA(call)                                                              \
  do {                                                                             \
    cudaError_t _e = (call);                                                       \
    if (_e != cudaSuccess) return agb::cuda_fail(_e, #call, __FILE__, __LINE__);   \
  } while (0)

#define AG_REQUIRE_INIT()                                                          \
  do {                                                                             \
    if (!agb::ctx().ready) {                                                       \
      agb::set_error("libagb200: ag_init has not been called");                    \
      return AG_ESTATE;                                                            \
    }                                                                              \
  } while (0)

#define AG_CHECK(cond, code, ...)                                                  \
  do {                                                                             \
    if (!(cond)) {                                                                 \
      agb::set_error(__VA_ARGS__);                                                 \
      return (code);                                                               \
    }                                                                              \
  } while (0)

// Call after every kernel launch: counts it and surfaces launch-configuration errors.
#define AG_LAUNCHED()                                                              \
  do {                                                                             \
    agb::ctx().launches++;                                                         \
    cudaError_t _e = cudaPeekAtLastError();                                        \
    if (_e != cudaSuccess) {                                                       \
      cudaGetLastError();                                                          \
      return agb::cuda_fail(_e, "kernel launch", __FILE__, __LINE__);              \
    }                                                                              \
  } while (0)
