// gemm_tc.cu -- the Float32 matrix product on the 5th-generation tensor cores.
//
// Reference: forward `*` and its gradient products `dy*x2'` (NT) and `x1'*dy` (TN), src/base.jl:26, which the
// reference hands to OpenBLAS sgemm; with an optional fused epilogue for the `.+ b` / activation nodes that follow
// a product in every model of the reference (examples/mnist.jl:32, examples/charlm.jl:74-77).
//
// Design (sm_100a):
//   * persistent, warp-specialised: warp 0 = TMA producer, warp 1 = tcgen05.mma issuer (one thread),
//     warps 2-9 = operand splitters, warps 10-13 = epilogue.  One CTA per SM; one 6-stage ring (TMA tile in smem,
//     B_lo in smem, A hi/lo in tensor memory).
//   * operands are fed by TMA (cp.async.bulk.tensor.2d) straight from the caller's column-major arrays; either
//     operand may be K-major (SWIZZLE_128B) or MN-major (SWIZZLE_128B_ATOM_32B), so NN / NT / TN / TT all map onto
//     the same kernel without a transpose pass (the instruction descriptor carries the majors).
//   * Float32 accuracy on a TF32 pipe (3xTF32): D += A_hi*B_lo + A_lo*B_hi + A_hi*B_hi with fp32 accumulation in
//     TMEM.  Two accuracy findings of round 2 (tests/test_gpu_baseline_sizes.py, tests/probe_lstm_accuracy.py):
//     (1) A_hi must be ROUNDED to tf32, not truncated: with truncation the residual A_lo has the sign of A, the
//     dropped A_lo*B_lo term has the sign of the product, and every product shrinks systematically;
//     (2) the tensor core accumulates with truncation, a shrink of ~1.7e-7 per accumulated k-block; it is
//     harmless for one product but compounds through a recurrent chain (config 4, T = 100: 2e-5 of the gradient when
//     the TMEM partial is promoted to fp32 registers every 8 k-blocks, 2.7e-6 when promoted every k-block; FFMA
//     6e-7).  The promotion interval is a template parameter: 1 for general shapes, 8 for the HBM-bound skinny
//     shapes (min(M,N) <= 64: the MNIST products, where the epilogue warps would otherwise bound the kernel).
//   * D tile = 128 (mm) x 64 (nn) fp32 in TMEM, double buffered (2 x 64 columns) so the promotion of one chunk
//     overlaps the MMAs of the next.  The large extent of C is mapped to mm.
//   * A goes through tensor memory (TS-mode MMA): with N = 64 an SS-mode MMA re-reads a 4 KB A slice from smem for
//     32 cycles of math and is smem-port bound; the splitters stage exact tf32 hi / lo values with tcgen05.st.
//   * split-K for small outputs with a huge contraction (dW = dy*x', K = batch): partials go to the workspace and
//     are summed in a fixed order (deterministic) by k_tc_splitk_reduce, which also applies the epilogue.
//   * Round-2 experiments that did NOT pay (each measured on the GPU and reverted, gpurun_out/r2_*.txt):
//     (1) separate rings for A (8 stages, released by the splitters), B (6) and the TMEM staging (4) with four
//     accumulators: 63 us against 50 us for the balanced 3-wave case -- and releasing a TMA stage right after
//     generic-proxy loads needs fence.proxy.async before the mbarrier arrive (without it later tiles of a CTA read
//     rows the refill had already overwritten);  (2) work distribution in units of k-blocks ("stream-K": 512 tiles
//     on 148 SMs in 3.46 instead of 4 tile times) with the shared tiles combined in-kernel through the workspace
//     (last arrival adds the partials in k order): the store / fence / atomic / read-back of the combination costs
//     what the balance gains (63.7 us against 60.6 us), also when the shared tiles are processed first;  (3) the
//     split-K reduction in-kernel (every CTA of a tile reduces a slice after a counter barrier): 64.6 us against
//     60.8 us with the separate combine launch.
// Roofline: for the MNIST-MLP shapes (M = 64) the products are HBM-bound: algorithmic bytes
// (MK + KN + MN)*4.
#include "common.cuh"
#include "ewise_ops.cuh"
#include "gemm_tc.cuh"

#include <cuda.h>
#include <stdlib.h>

namespace agb {

constexpr int TC_BM = 128;      // mm tile (TMEM lanes)
constexpr int TC_BN = 64;       // nn tile (TMEM columns per accumulator)
constexpr int TC_BK = 32;       // k block: 32 floats = one 128-byte swizzle row
constexpr int TC_STAGES = 6;    // one ring: TMA tiles (smem), B_lo (smem), A hi/lo (TMEM)
constexpr int TC_A_BYTES = TC_BM * TC_BK * 4;   // 16 KB
constexpr int TC_B_BYTES = TC_BN * TC_BK * 4;   //  8 KB
constexpr int TC_RAW_BYTES = TC_A_BYTES + TC_B_BYTES;          // TMA stage: A and B as loaded (fp32)
constexpr int TC_SPLIT_WARPS = 8;
constexpr int TC_THREADS = 15 * 32;
constexpr int TC_MMA2_WARP = 14;               // second MMA issuer (FLUSH = 1 kernels only)
constexpr int TC_EPI_WARP0 = 10;                // warps 10..13
constexpr int TC_NBARS = 3 * TC_STAGES + 4;
constexpr int TC_SMEM_DATA = TC_STAGES * (TC_RAW_BYTES + TC_B_BYTES);
constexpr int TC_SMEM = TC_SMEM_DATA + 1024 /*align*/ + 8 * TC_NBARS + 64;
constexpr int TC_TMEM_ACC = 2 * TC_BN;          // columns [0,128): two accumulators
constexpr int TC_TMEM_ASTAGE = 2 * TC_BK;       // per stage: 32 columns A_hi + 32 columns A_lo
constexpr int TC_TMEM_COLS = 512;               // 128 + 6*64 = 512: the whole tensor memory of the SM
static_assert(TC_TMEM_ACC + TC_STAGES * TC_TMEM_ASTAGE <= TC_TMEM_COLS, "tensor memory budget");
static_assert(TC_SMEM <= 232448, "shared memory budget");

struct TcParams {
  int mm_tiles, nn_tiles, splits;
  int mm_pitch;            // rows between consecutive mm tiles (<= TC_BM: a tile always loads TC_BM rows, stores mm_pitch)
  int kblocks_total, kblocks_per_split;
  int64_t MM, NN;          // extents along mm / nn
  float* out;              // C, or the split-K workspace
  float* out2;             // optional second output (activation), same layout as C (direct path only)
  const float* bias;       // optional, indexed by C's row (direct path only)
  int act;                 // TC_ACT_*
  int64_t ld_out;          // leading dimension of out (elements)
  int64_t split_stride;    // elements between consecutive split partials
  int transposed;          // 1: mm runs along C's columns (out[nn + ld*mm]); 0: out[mm + ld*nn]
  float alpha, beta;       // beta != 0 goes through the workspace path
};

// Up to TC_GROUP_MAX independent products of one class (same operand majors, same orientation, same promotion
// interval) run as ONE launch: the work items of all problems form one list.  The four gate products of an LSTM step
// (examples/charlm.jl:69-73: "we can probably combine these four operations into one"), their four dx products and
// the dW products of several timesteps are issued this way by the host (lazy._resolve_products); a small product
// pays ~9 us of fixed cost per launch (launch, TMEM allocation, pipeline fill and drain), which dominates at those sizes.
constexpr int TC_GROUP_MAX = 8;
struct TcGroup {
  CUtensorMap tmA[TC_GROUP_MAX], tmB[TC_GROUP_MAX];
  TcParams p[TC_GROUP_MAX];
  int work_start[TC_GROUP_MAX + 1];     // prefix sums of the problems' work items (tiles * splits)
  int nprob;
};

#ifdef AGB_TC_TRACE
// Experiment build (tools/build_variant.sh ... "-DAGB_TC_TRACE"): block 0 records globaltimer stamps per k-block and
// role: [0] producer issued the loads, [1] splitter saw the data, [2] splitter done, [3] MMA issued, [4] epilogue saw the
// partial, [5] epilogue released the accumulator.  Read back with ag_gemm_trace (tests/probe_tc_trace.py).
__device__ unsigned long long g_tc_trace[6][256];
__device__ __forceinline__ unsigned long long tc_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define TC_TRACE(ROLE, IDX) do { if (blockIdx.x == 0 && (IDX) < 256) g_tc_trace[ROLE][IDX] = tc_now(); } while (0)
#else
#define TC_TRACE(ROLE, IDX) do { } while (0)
#endif

// ---- PTX wrappers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n.reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must trap, never hang the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  for (uint32_t spin = 0; !mbar_try(bar, parity); ++spin) {
    if (spin > (1u << 26)) {
      printf("libagb200 gemm_tc: mbarrier timeout (block %d thread %d bar %u parity %u)\n", blockIdx.x,
             threadIdx.x, bar, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// One lane of a fully converged warp.  Keeping the TMA / MMA roles warp-uniform (every lane runs the loop, only
// the instruction itself is predicated on the elected lane) lets ptxas keep descriptors and barrier addresses
// in uniform registers; wrapping the whole role in `if (lane == 0)` instead costs a serialising ELECT/R2UR loop
// around every UTCHMMA/UTMALDG (measured in round 1: ~1000 cycles of issue time per k-block).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// lo = x - trunc_tf32(x) (exact in fp32), then +half-ulp(tf32) on the bit pattern: the tensor core's own
// truncation of the operand then yields round-to-nearest (ties away) -- integer ops instead of cvt.rna.tf32.
// (B's hi part is the TMA tile itself, truncated by the tensor core; A's hi part is rounded, see split_tf32.)
__device__ __forceinline__ float lo_tf32(float x) {
  const float d = x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
  return __uint_as_float(__float_as_uint(d) + 0x1000u);
}
// TS-mode MMA: A (128 x 8 tf32) from tensor memory, B from shared memory through a UMMA descriptor
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// registers -> tensor memory: this thread's TMEM lane, 16 consecutive columns
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
// tf32 split of an fp32 value: hi = rna_tf32(x), lo = rna_tf32(x - hi) (integer rounding on the magnitude bits).
// hi must be ROUNDED, not truncated: with hi = trunc(x) the residual lo always has the sign of x, so the dropped
// term a_lo*b_lo always has the sign of a*b and every product shrinks systematically.  With a rounded hi the
// residual has either sign and the dropped term is zero-mean.
__device__ __forceinline__ void split_tf32(uint32_t x, uint32_t& hi, uint32_t& lo) {
  hi = (x + 0x1000u) & 0xFFFFE000u;
  const float d = __uint_as_float(x) - __uint_as_float(hi);      // exact in fp32
  lo = (__float_as_uint(d) + 0x1000u) & 0xFFFFE000u;
}
// 32 bytes (one full sector) per lane and instruction: the epilogue threads own rows of 256 contiguous bytes at a
// 256-byte pitch, so 16-byte stores would write half sectors, two instructions per sector
__device__ __forceinline__ void st_global_v8(float* p, const float* v) {
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]),
               "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}
// tanh / sigm as CALLED functions for the unrolled epilogue loops: inlined, 64 copies of a 30-40 instruction sequence
// are a 40 KB store path that runs once per tile -- paid for in instruction fetch (config 4 with the gate products
// not split along K: +25 us per launch).  Same expressions as U<>::f, so the values do not change.
__device__ __noinline__ float tc_tanh_call(float z) { return U<AG_U_TANH, float>::f(z); }
__device__ __noinline__ float tc_sigm_call(float z) { return U<AG_U_SIGM, float>::f(z); }
template <int ACT>
__device__ __forceinline__ float tc_activation_compact(float z) {
  if (ACT == TC_ACT_RELU) return relu(z);
  if (ACT == TC_ACT_TANH) return tc_tanh_call(z);
  if (ACT == TC_ACT_SIGM) return tc_sigm_call(z);
  return z;
}
template <int ACT>
__device__ __forceinline__ float tc_activation(float z) {
  if (ACT == TC_ACT_RELU) return relu(z);     // max.(0, z): same values as the FMax map jmax(0, z) of ewise_ops.cuh
  if (ACT == TC_ACT_TANH) return U<AG_U_TANH, float>::f(z);
  if (ACT == TC_ACT_SIGM) return U<AG_U_SIGM, float>::f(z);
  return z;
}
// One finished element of C is alpha, bias, activation -- each operation rounded on its own, so the values are those
// of the unfused sequence `*` -> `.+` -> activation.
// The 64 values of one tile row (mm) held by an epilogue thread.
template <int ACT>
__device__ __forceinline__ void tc_store_row(const TcParams& p, int64_t mm, int64_t nn0, const float* v) {
  if (mm >= p.MM) return;
  if (p.transposed && nn0 + TC_BN <= p.NN && (p.ld_out & 7) == 0 &&
      ((((uintptr_t)p.out) | ((uintptr_t)p.out2)) & 31) == 0) {
    // this thread owns 64 consecutive elements of one column of C: 256-bit stores; bias is indexed by nn (C's row)
    float* dst = p.out + nn0 + p.ld_out * mm;
    float* dst2 = p.out2 ? p.out2 + nn0 + p.ld_out * mm : nullptr;
#pragma unroll
    for (int j = 0; j < TC_BN; j += 8) {
      float z[8], a[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        z[e] = p.alpha == 1.0f ? v[j + e] : __fmul_rn(p.alpha, v[j + e]);
        if (p.bias) z[e] = __fadd_rn(z[e], __ldg(p.bias + nn0 + j + e));
        a[e] = tc_activation_compact<ACT>(z[e]);
      }
      if (dst2) {
        st_global_v8(dst + j, z);
        st_global_v8(dst2 + j, a);
      } else {
        st_global_v8(dst + j, a);
      }
    }
  } else if (p.transposed && nn0 + TC_BN <= p.NN && (p.ld_out & 3) == 0) {
    // the same with 128-bit stores (leading dimension or base not 32-byte aligned)
    float* dst = p.out + nn0 + p.ld_out * mm;
    float* dst2 = p.out2 ? p.out2 + nn0 + p.ld_out * mm : nullptr;
#pragma unroll
    for (int j = 0; j < TC_BN; j += 4) {
      float z[4], a[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        z[e] = p.alpha == 1.0f ? v[j + e] : __fmul_rn(p.alpha, v[j + e]);
        if (p.bias) z[e] = __fadd_rn(z[e], __ldg(p.bias + nn0 + j + e));
        a[e] = tc_activation_compact<ACT>(z[e]);
      }
      if (dst2) {
        *reinterpret_cast<float4*>(dst + j) = make_float4(z[0], z[1], z[2], z[3]);
        *reinterpret_cast<float4*>(dst2 + j) = make_float4(a[0], a[1], a[2], a[3]);
      } else {
        *reinterpret_cast<float4*>(dst + j) = make_float4(a[0], a[1], a[2], a[3]);
      }
    }
  } else if (!p.transposed) {
    // lanes hold consecutive mm = consecutive rows of C: every store of the warp is one coalesced run.  Everything
    // that does not depend on j is hoisted: the 64 stores are unrolled (register array), and a long unrolled body is
    // paid for in instruction fetch -- measured on a 16-tile product (round 2, ncu source page): 36 us of which 30
    // were the epilogue warps fetching a 32 KB store sequence once.
    const int nvalid = (int)min((int64_t)TC_BN, p.NN - nn0);
    const bool hb = p.bias != nullptr, scale = p.alpha != 1.0f;
    const float b = hb ? __ldg(p.bias + mm) : 0.f, alpha = p.alpha;
    const int64_t ld = p.ld_out;
    float* dst = p.out + mm + ld * nn0;
    float* dst2 = p.out2 ? p.out2 + mm + ld * nn0 : nullptr;
    if (ACT == TC_ACT_NONE && nvalid == TC_BN && !scale && !hb && !dst2) {
      // the plain product (every gradient product of the tapes): three instructions per element -- the executed
      // footprint of this path is what a small product pays in instruction fetch every time the tape comes back to it
      // (measured: 13.2 us per launch interleaved with other kernels against 9.9 us back to back)
#pragma unroll
      for (int j = 0; j < TC_BN; ++j) {
        *dst = v[j];
        dst += ld;
      }
      return;
    }
#pragma unroll
    for (int j = 0; j < TC_BN; ++j) {
      if (j < nvalid) {
        float z = scale ? __fmul_rn(alpha, v[j]) : v[j];
        if (hb) z = __fadd_rn(z, b);
        const float a = tc_activation_compact<ACT>(z);
        if (dst2) {
          *dst = z;
          *dst2 = a;
          dst2 += ld;
        } else {
          *dst = a;
        }
        dst += ld;
      }
    }
  } else {
    // transposed tile at the ragged edge or with an unaligned leading dimension: this thread's run of one column of
    // C, element by element (unrolled: v is a register array)
    const int nvalid = (int)min((int64_t)TC_BN, p.NN - nn0);
    const bool hb = p.bias != nullptr, scale = p.alpha != 1.0f;
    const float alpha = p.alpha;
    const float* bias = hb ? p.bias + nn0 : nullptr;
    float* dst = p.out + nn0 + p.ld_out * mm;
    float* dst2 = p.out2 ? p.out2 + nn0 + p.ld_out * mm : nullptr;
#pragma unroll
    for (int j = 0; j < TC_BN; ++j) {
      if (j < nvalid) {
        float z = scale ? __fmul_rn(alpha, v[j]) : v[j];
        if (hb) z = __fadd_rn(z, __ldg(bias + j));
        const float a = tc_activation_compact<ACT>(z);
        if (dst2) {
          dst[j] = z;
          dst2[j] = a;
        } else {
          dst[j] = a;
        }
      }
    }
  }
}
#define TC_DISPATCH_ACT(act, CALL)                 \
  do {                                             \
    switch (act) {                                 \
      case TC_ACT_RELU: CALL(TC_ACT_RELU); break;  \
      case TC_ACT_TANH: CALL(TC_ACT_TANH); break;  \
      case TC_ACT_SIGM: CALL(TC_ACT_SIGM); break;  \
      default: CALL(TC_ACT_NONE); break;           \
    }                                              \
  } while (0)

// A_KMAJOR / B_KMAJOR: memory major of the two operands (K contiguous or mm/nn contiguous).
// FLUSH: k-blocks accumulated in TMEM before the partial is promoted to fp32 registers.
template <bool A_KMAJOR, bool B_KMAJOR, int FLUSH>
__global__ void __launch_bounds__(TC_THREADS, 1)
    k_gemm_tc(const __grid_constant__ TcGroup g) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t blo_base = smem_base + TC_STAGES * TC_RAW_BYTES;          // B_lo ring
  const uint32_t bars = blo_base + TC_STAGES * TC_B_BYTES;
  auto full_bar = [&](int s) { return bars + 8u * s; };                     // TMA bytes landed
  auto ready_bar = [&](int s) { return bars + 8u * (TC_STAGES + s); };      // A in TMEM, B_lo in smem
  auto empty_bar = [&](int s) { return bars + 8u * (2 * TC_STAGES + s); };  // MMAs of the stage retired
  auto tfull = [&](int a) { return bars + 8u * (3 * TC_STAGES + a); };
  auto tempty = [&](int a) { return bars + 8u * (3 * TC_STAGES + 2 + a); };
  const uint32_t tmem_slot = bars + 8u * TC_NBARS;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + TC_SMEM_DATA + 8 * TC_NBARS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int q = 0; q < g.nprob; ++q) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&g.tmA[q]) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&g.tmB[q]) : "memory");
    }
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(ready_bar(s), TC_SPLIT_WARPS);   // one elected arrival per splitter warp
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull(a), 1); mbar_init(tempty(a), 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                 "r"(TC_TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  const int total_work = g.work_start[g.nprob];
  // work item -> (problem, mm tile, nn tile, split)
  auto decode = [&](int w, int& q, int& mt, int& nt, int& kb0, int& kb1) {
    q = 0;
    while (q + 1 < g.nprob && w >= g.work_start[q + 1]) ++q;
    w -= g.work_start[q];
    const TcParams& p = g.p[q];
    mt = w % p.mm_tiles;
    int r = w / p.mm_tiles;
    nt = r % p.nn_tiles;
    int sp = r / p.nn_tiles;
    kb0 = sp * p.kblocks_per_split;
    kb1 = min(kb0 + p.kblocks_per_split, p.kblocks_total);
    return sp;
  };

  if (warp == 0) {
    // ===== TMA producer (warp-uniform; one elected lane issues) =====
    int s = 0;
    uint32_t ph = 0;
    int tr_i = 0;
    (void)tr_i;
    for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
      int q, mt, nt, kb0, kb1;
      decode(w, q, mt, nt, kb0, kb1);
      const CUtensorMap* mapA = &g.tmA[q];
      const CUtensorMap* mapB = &g.tmB[q];
      const int mm0 = mt * g.p[q].mm_pitch;
      // a K-major A tile is loaded with a box of mm_pitch rows (the tile's own rows only); an MN-major one in four
      // 32-row boxes (its last rows overlap the next tile when mm_pitch < TC_BM)
      const uint32_t tx_bytes = (A_KMAJOR ? (uint32_t)g.p[q].mm_pitch * (TC_BK * 4) : (uint32_t)TC_A_BYTES) + TC_B_BYTES;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(empty_bar(s), ph ^ 1);
        const uint32_t st = smem_base + s * TC_RAW_BYTES;
        if (elect_one()) {
          mbar_expect_tx(full_bar(s), tx_bytes);
          if (A_KMAJOR) {
            tma_load_2d(st, mapA, full_bar(s), kb * TC_BK, mm0);
          } else {
#pragma unroll
            for (int c = 0; c < TC_BM / 32; ++c)
              tma_load_2d(st + c * 4096, mapA, full_bar(s), mm0 + c * 32, kb * TC_BK);
          }
          const uint32_t sb = st + TC_A_BYTES;
          if (B_KMAJOR) {
            tma_load_2d(sb, mapB, full_bar(s), kb * TC_BK, nt * TC_BN);
          } else {
#pragma unroll
            for (int c = 0; c < TC_BN / 32; ++c)
              tma_load_2d(sb + c * 4096, mapB, full_bar(s), nt * TC_BN + c * 32, kb * TC_BK);
          }
        }
        __syncwarp();
        if (lane == 0) TC_TRACE(0, tr_i);
        ++tr_i;
        if (++s == TC_STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1 || (FLUSH == 1 && warp == TC_MMA2_WARP)) {
    // ===== MMA issuer (warp-uniform; one elected lane issues) =====
    // When the partial is promoted every k-block (FLUSH = 1) TWO warps issue, one per accumulator: warp 1 takes the
    // even k-blocks (accumulator 0), warp 14 the odd ones (accumulator 1).  Measured with the k-block trace of the
    // experiment build (round 2): one issuer needs 480 ns per k-block -- 200 ns blocked in the 12 tcgen05.mma (the
    // issue of an MMA waits for the pipe) and 280 ns of barrier waits, fences, descriptors and commits during which
    // the tensor pipe idles; with two issuers one warp's overhead runs under the other's MMAs.  No accumulator is
    // shared between the issuers, so the order of the additions (and every bit of the result) is unchanged.
    constexpr bool TWO = FLUSH == 1;
    const int me = warp == 1 ? 0 : 1;
    // instruction descriptor: D=f32, A=B=tf32, A K-major (TMEM), B major, N, M
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((B_KMAJOR ? 0u : 1u) << 16) |
                           ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
    const uint32_t b_lbo = B_KMAJOR ? 16u : 4096u;
    const uint32_t b_kstep = B_KMAJOR ? 32u : 1024u;   // bytes per UMMA_K = 8
    const uint32_t b_sbo = B_KMAJOR ? 1024u : 512u;
    const uint32_t b_lt = B_KMAJOR ? 2u : 1u;          // SWIZZLE_128B / SWIZZLE_128B_BASE32B
    // descriptor words: hi = SBO | version<<14 | layout<<29 ; lo = addr>>4 | (LBO>>4)<<16
    const uint32_t desc_hi_b = ((b_sbo >> 4) & 0x3FFFu) | (1u << 14) | (b_lt << 29);
    const uint32_t desc_lo_bhi = (((smem_base + TC_A_BYTES) >> 4) & 0x3FFFu) | ((b_lbo >> 4) << 16);
    const uint32_t desc_lo_blo = ((blo_base >> 4) & 0x3FFFu) | ((b_lbo >> 4) << 16);
    int s = 0, ci = 0;   // stage, chunk counter (selects the TMEM accumulator)
    uint32_t ph = 0;
    int tr_i = 0;
    (void)tr_i;
    for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
      int q, mt, nt, kb0, kb1;
      decode(w, q, mt, nt, kb0, kb1);
      for (int kc0 = kb0; kc0 < kb1; kc0 += FLUSH, ++ci) {
        const int kc1 = min(kc0 + FLUSH, kb1);
        const int acc = ci & 1;
        const bool mine = !TWO || acc == me;
        if (!mine) {                       // the other issuer's k-block: only the stage bookkeeping advances
          if (++s == TC_STAGES) { s = 0; ph ^= 1; }
          ++tr_i;
          continue;
        }
        mbar_wait(tempty(acc), ((ci >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * TC_BN;
        for (int kb = kc0; kb < kc1; ++kb) {
          mbar_wait(ready_bar(s), ph);
          tc_fence_after();
          // operands: A_hi / A_lo in TMEM (columns), B_hi in the TMA stage, B_lo in the B_lo ring
          const uint32_t a_t = tmem_base + TC_TMEM_ACC + (uint32_t)s * TC_TMEM_ASTAGE;
          const uint32_t rb = desc_lo_bhi + (uint32_t)s * (TC_RAW_BYTES >> 4);
          const uint32_t lb = desc_lo_blo + (uint32_t)s * (TC_B_BYTES >> 4);
          uint64_t dbh[TC_BK / 8], dbl[TC_BK / 8];
#pragma unroll
          for (int k = 0; k < TC_BK / 8; ++k) {
            dbh[k] = ((uint64_t)desc_hi_b << 32) | (uint64_t)(rb + k * (b_kstep >> 4));
            dbl[k] = ((uint64_t)desc_hi_b << 32) | (uint64_t)(lb + k * (b_kstep >> 4));
          }
          const uint32_t first = kb > kc0 ? 1u : 0u;
          const uint32_t bar_e = empty_bar(s);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < TC_BK / 8; ++k) {
              const uint32_t a_hi = a_t + k * 8, a_lo = a_t + TC_BK + k * 8;
              umma_tf32_ts(d_tmem, a_hi, dbl[k], idesc, k > 0 ? 1u : first);   // small terms first
              umma_tf32_ts(d_tmem, a_lo, dbh[k], idesc, 1u);
              umma_tf32_ts(d_tmem, a_hi, dbh[k], idesc, 1u);
            }
            umma_commit(bar_e);               // stage (smem + TMEM) reusable when these MMAs retire
          }
          __syncwarp();
          if (lane == 0) TC_TRACE(3, tr_i);
          ++tr_i;
          if (++s == TC_STAGES) { s = 0; ph ^= 1; }
        }
        if (elect_one()) umma_commit(tfull(acc));        // chunk partial complete
        __syncwarp();
      }
    }
  } else if (warp < 2 + TC_SPLIT_WARPS) {
    // ===== operand splitters =====
    // A: each thread owns one tile row (= its TMEM lane) and half of the k-block: reads 16 fp32 from the TMA
    //    tile, writes hi and lo (exact tf32 values) to tensor memory with tcgen05.st.
    // B: lo = rna_tf32(x - trunc_tf32(x)) to the B_lo ring in shared memory (the tensor core truncates B_hi itself).
    const int tid = threadIdx.x - 64;          // 0..255
    const int lq = warp & 3;                   // TMEM lane quarter of this warp
    const int half = (warp - 2) >> 2;          // which 16 k of the 32
    const int row = lq * 32 + lane;             // tile row
    int s = 0;
    uint32_t ph = 0;
    int tr_i = 0;
    (void)tr_i;
    for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
      int q, mt, nt, kb0, kb1;
      decode(w, q, mt, nt, kb0, kb1);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(full_bar(s), ph);
        if (threadIdx.x == 64) TC_TRACE(1, tr_i);
        tc_fence_after();                   // order the TMEM stores below after the stage hand-over
        const uint8_t* rawA = smem_gen + s * TC_RAW_BYTES;
        uint32_t v[16];
        if (A_KMAJOR) {
          // row r at r*128 B, 16-byte chunk c stored at c ^ (r & 7)   (SWIZZLE_128B)
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int chunk = (half * 4 + c) ^ (row & 7);
            const uint4 t4 = *reinterpret_cast<const uint4*>(rawA + row * 128 + chunk * 16);
            v[4 * c] = t4.x; v[4 * c + 1] = t4.y; v[4 * c + 2] = t4.z; v[4 * c + 3] = t4.w;
          }
        } else {
          // 32-row (mm) chunks of 4 KB; inside: k-row at k*128 B, 32-byte atom a stored at a ^ (k & 3)
          // (SWIZZLE_128B_ATOM_32B); this thread's mm = row -> chunk lq, position lane
          const uint8_t* base = rawA + lq * 4096 + (lane & 7) * 4;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int k = half * 16 + j;
            const int atom = (lane >> 3) ^ (k & 3);
            v[j] = *reinterpret_cast<const uint32_t*>(base + k * 128 + atom * 32);
          }
        }
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) split_tf32(v[j], hi[j], lo[j]);
        const uint32_t a_t = tmem_base + ((uint32_t)(lq * 32) << 16) + TC_TMEM_ACC + (uint32_t)s * TC_TMEM_ASTAGE;
        tmem_st16(a_t + half * 16, hi);
        tmem_st16(a_t + TC_BK + half * 16, lo);
        // B_lo: 512 float4 per stage, 2 per thread
        const float4* rawB = reinterpret_cast<const float4*>(smem_gen + s * TC_RAW_BYTES + TC_A_BYTES);
        float4* blo = reinterpret_cast<float4*>(smem_gen + TC_STAGES * TC_RAW_BYTES + s * TC_B_BYTES);
#pragma unroll
        for (int i = 0; i < TC_B_BYTES / 16 / 256; ++i) {
          const float4 x4 = rawB[tid + i * 256];
          float4 r;
          r.x = lo_tf32(x4.x); r.y = lo_tf32(x4.y); r.z = lo_tf32(x4.z); r.w = lo_tf32(x4.w);
          blo[tid + i * 256] = r;
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        fence_async_smem();                 // generic-proxy smem writes -> visible to the tensor core
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(ready_bar(s));
        if (threadIdx.x == 64) TC_TRACE(2, tr_i);
        ++tr_i;
        if (++s == TC_STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp >= TC_EPI_WARP0 && warp < TC_EPI_WARP0 + 4) {
    // ===== epilogue: promote each chunk partial TMEM -> registers (fp32 RN adds), finish the tile =====
    const int lq = warp & 3;             // TMEM lane quarter this warp may read
    const int r_loc = lq * 32 + lane;     // tile row of this thread
    int ci = 0;
    for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
      int q, mt, nt, kb0, kb1;
      const int sp = decode(w, q, mt, nt, kb0, kb1);
      const TcParams& p = g.p[q];
      float accv[TC_BN];
#pragma unroll
      for (int j = 0; j < TC_BN; ++j) accv[j] = 0.f;
      for (int kc0 = kb0; kc0 < kb1; kc0 += FLUSH, ++ci) {
        const int acc = ci & 1;
        mbar_wait(tfull(acc), (ci >> 1) & 1);
        if (threadIdx.x == TC_EPI_WARP0 * 32) TC_TRACE(4, ci);
        tc_fence_after();
#pragma unroll
        for (int c0 = 0; c0 < TC_BN; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(acc * TC_BN + c0), r);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 32; ++j) accv[c0 + j] += __uint_as_float(r[j]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty(acc));
        if (threadIdx.x == TC_EPI_WARP0 * 32) TC_TRACE(5, ci);
      }
      // rows beyond the tile pitch belong to the next tile (which computes them itself)
      const int64_t mm = r_loc < p.mm_pitch ? (int64_t)mt * p.mm_pitch + r_loc : p.MM;
      const int64_t nn0 = (int64_t)nt * TC_BN;
      if (p.split_stride == 0) {
        // direct path: alpha, bias, activation and the stores
#define TC_CALL(ACT) tc_store_row<ACT>(p, mm, nn0, accv)
        TC_DISPATCH_ACT(p.act, TC_CALL);
#undef TC_CALL
      } else if (mm < p.MM) {
        // workspace path (split-K or beta != 0): raw partial in C's own layout, combined by k_tc_splitk_reduce
        float* __restrict__ out = p.out + (int64_t)sp * p.split_stride;
        if (p.transposed) {
          float* dst = out + nn0 + p.ld_out * mm;
          if (nn0 + TC_BN <= p.NN && (p.ld_out & 7) == 0 && (((uintptr_t)out) & 31) == 0) {
#pragma unroll
            for (int j = 0; j < TC_BN; j += 8) st_global_v8(dst + j, accv + j);
          } else if (nn0 + TC_BN <= p.NN && (p.ld_out & 3) == 0) {
#pragma unroll
            for (int j = 0; j < TC_BN; j += 4)
              *reinterpret_cast<float4*>(dst + j) = make_float4(accv[j], accv[j + 1], accv[j + 2], accv[j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < TC_BN; ++j)
              if (nn0 + j < p.NN) dst[j] = accv[j];
          }
        } else {
          const int nvalid = (int)min((int64_t)TC_BN, p.NN - nn0);
          const int64_t ld = p.ld_out;
          float* dst = out + mm + ld * nn0;
          if (nvalid == TC_BN) {
#pragma unroll
            for (int j = 0; j < TC_BN; ++j) {
              *dst = accv[j];
              dst += ld;
            }
          } else {
#pragma unroll
            for (int j = 0; j < TC_BN; ++j) {
              if (j < nvalid) *dst = accv[j];
              dst += ld;
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS));
  }
}

// C = act(alpha * sum_s ws[s] + beta*C + bias) in a fixed order (deterministic split-K) for every problem of a group
// (blockIdx.y); ws partials are M x N column-major with leading dimension M.  With C2: C receives the pre-activation
// value, C2 the activation.
struct TcReduceGroup {
  const float* ws[TC_GROUP_MAX];
  int splits[TC_GROUP_MAX];
  int64_t M[TC_GROUP_MAX], N[TC_GROUP_MAX], ldc[TC_GROUP_MAX];
  float alpha[TC_GROUP_MAX], beta[TC_GROUP_MAX];
  const float* bias[TC_GROUP_MAX];
  float* C[TC_GROUP_MAX];
  float* C2[TC_GROUP_MAX];
  int act[TC_GROUP_MAX];
  int nprob;
};
__global__ void __launch_bounds__(256) k_tc_splitk_reduce(const __grid_constant__ TcReduceGroup r) {
  const int q = blockIdx.y;
  const int64_t M = r.M[q], N = r.N[q];
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * N) return;
  const float* __restrict__ ws = r.ws[q];
  const int splits = r.splits[q];
  int64_t m = i % M, n = i / M;
  // all partials of a block of 8 are in flight before the (fixed-order) additions
  float s0 = 0.f, s1 = 0.f;
  const int64_t MN = M * N;
  int k = 0;
  for (; k + 8 <= splits; k += 8) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = ws[(int64_t)(k + u) * MN + i];
#pragma unroll
    for (int u = 0; u < 8; u += 2) { s0 += v[u]; s1 += v[u + 1]; }
  }
  for (; k + 1 < splits; k += 2) {
    s0 += ws[(int64_t)k * MN + i];
    s1 += ws[(int64_t)(k + 1) * MN + i];
  }
  if (k < splits) s0 += ws[(int64_t)k * MN + i];
  const int64_t idx = m + n * r.ldc[q];
  float* __restrict__ C = r.C[q];
  float z = __fmul_rn(r.alpha[q], s0 + s1);
  if (r.beta[q] != 0.f) z = __fadd_rn(z, __fmul_rn(r.beta[q], C[idx]));
  if (r.bias[q]) z = __fadd_rn(z, r.bias[q][m]);
  float a = z;
  switch (r.act[q]) {
    case TC_ACT_RELU: a = tc_activation<TC_ACT_RELU>(z); break;
    case TC_ACT_TANH: a = tc_activation<TC_ACT_TANH>(z); break;
    case TC_ACT_SIGM: a = tc_activation<TC_ACT_SIGM>(z); break;
    default: break;
  }
  if (r.C2[q]) {
    C[idx] = z;
    r.C2[q][idx] = a;
  } else {
    C[idx] = a;
  }
}

// ---- host side --------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess &&
        qr == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
    else
      cudaGetLastError();
  }
  return fn;
}

static int tc_l2promo() {        // experiment knob, read once: AGB_TC_L2PROMO = 0 none, 1 64 B, 2 128 B, 3 256 B (default)
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("AGB_TC_L2PROMO");
    v = e ? atoi(e) : 3;
    if (v < 0 || v > 3) v = 3;
  }
  return v;
}

// 2-d fp32 tensor map: inner (contiguous) extent d0, outer extent d1 with `ld` elements between rows
static bool make_map(CUtensorMap* m, const void* base, int64_t d0, int64_t d1, int64_t ld, int box0, int box1,
                     bool kmajor) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {(cuuint64_t)d0, (cuuint64_t)d1};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)box0, (cuuint32_t)box1};
  cuuint32_t es[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE,
                  kmajor ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                  (CUtensorMapL2promotion)tc_l2promo(),
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// Experiment knobs, read once (not on every product): AGB_TC_SPLIT_KB = k-blocks from which a product with fewer
// tiles than SMs is split along K, AGB_TC_SPLIT_MINKB = k-blocks per split at least,
// AGB_TC_FLUSH_SKINNY = 1 promotes every k-block for the skinny shapes as well.
struct TcKnobs {
  int split_kb, min_kb, flush_skinny, even_tiles;
  TcKnobs() {
    const char* t = getenv("AGB_TC_EVEN_TILES");       // 0: mm tiles of 128 rows with a ragged last tile
    even_tiles = t ? atoi(t) : 1;
    const char* f = getenv("AGB_TC_FLUSH_SKINNY");     // 1: promote every k-block for the skinny shapes too
    flush_skinny = f ? atoi(f) : 8;
    const char* e = getenv("AGB_TC_SPLIT_KB");
    split_kb = e ? atoi(e) : 16;
    if (split_kb < 4) split_kb = 4;
    e = getenv("AGB_TC_SPLIT_MINKB");
    min_kb = e ? atoi(e) : 4;
    if (min_kb < 2) min_kb = 2;
  }
};
static const TcKnobs& tc_knobs() {
  static TcKnobs k;
  return k;
}

bool gemm_tc_init() {        // reads the environment knobs once, at ag_init (not on every product)
  (void)tc_knobs();
  (void)encode_fn();
  return true;
}

bool gemm_tc_supported(int transA, int transB, int64_t M, int64_t N, int64_t K, const void* A, int64_t lda,
                       const void* B, int64_t ldb, const void* C, int64_t ldc) {
  (void)transA; (void)transB;
  if (!encode_fn()) return false;
  if (((uintptr_t)A | (uintptr_t)B | (uintptr_t)C) & 15) return false;
  if ((lda | ldb | ldc) & 3) return false;                 // TMA strides are multiples of 16 bytes
  if (M < 8 || N < 8 || K < 8) return false;
  if (M >= ((int64_t)1 << 31) || N >= ((int64_t)1 << 31) || K >= ((int64_t)1 << 31)) return false;
  if (2.0 * (double)M * (double)N * (double)K < 3.0e7) return false;   // tiny: launch-bound either way
  return true;
}

// One class of products: operand majors after the orientation swap, orientation, promotion interval.
struct TcClass {
  bool ak, bk, transposed, skinny;
  bool operator==(const TcClass& o) const { return ak == o.ak && bk == o.bk && transposed == o.transposed && skinny == o.skinny; }
};
static TcClass tc_class(const TcJob& j) {
  TcClass c;
  c.transposed = j.N > j.M;
  const bool a_kmajor = j.transA != 0, b_kmajor = j.transB == 0;
  c.ak = c.transposed ? b_kmajor : a_kmajor;
  c.bk = c.transposed ? a_kmajor : b_kmajor;
  // promotion interval: every k-block, except for the HBM-bound skinny shapes (see the header comment)
  c.skinny = (j.M <= 64 || j.N <= 64) && tc_knobs().flush_skinny != 1;
  return c;
}

bool gemm_tc_same_class(const TcJob& a, const TcJob& b) { return tc_class(a) == tc_class(b); }

// n <= TC_GROUP_MAX products of ONE class as one launch (+ one combine launch when they are split along K).
ag_status gemm_tc_group(int n, const TcJob* jobs) {
  Context& c = ctx();
  AG_CHECK(n >= 1 && n <= TC_GROUP_MAX, AG_EINVAL, "gemm_tc_group: %d problems", n);
  const TcClass cls = tc_class(jobs[0]);
  for (int q = 1; q < n; ++q) AG_CHECK(tc_class(jobs[q]) == cls, AG_EINVAL, "gemm_tc_group: mixed classes");
  static TcGroup g;          // 3 KB: filled per call (one host thread, see the header of include/agb200.h)
  g.nprob = n;
  // Map the larger extent of C onto the 128-row mm tile.
  //   normal     : mm = row index m of C,    Aop = op(A) (M x K),  Bop = op(B)^T (N x K)
  //   transposed : mm = column index n of C, Aop = op(B)^T (N x K), Bop = op(A) (M x K)
  // op(A)[m,k]: transA ? A[k + m*lda] (K-major) : A[m + k*lda] (MN-major)
  // op(B)^T[n,k] = op(B)[k,n]: transB ? B[n + k*ldb] (MN-major) : B[k + n*ldb] (K-major)
  int64_t total_tiles = 0;
  int min_kblocks = 1 << 30;
  bool any_beta = false;
  for (int q = 0; q < n; ++q) {
    const TcJob& j = jobs[q];
    const void *pa, *pb;
    int64_t MM, NN, lda_, ldb_;
    if (!cls.transposed) { pa = j.A; lda_ = j.lda; MM = j.M; pb = j.B; ldb_ = j.ldb; NN = j.N; }
    else { pa = j.B; lda_ = j.ldb; MM = j.N; pb = j.A; ldb_ = j.lda; NN = j.M; }
    TcParams& p = g.p[q];
    p.mm_tiles = (int)cdiv(MM, TC_BM);
    p.nn_tiles = (int)cdiv(NN, TC_BN);
    // Equal tiles.  (1) 784 rows are 7 tiles of 112 (each loads 128 rows, the overlap comes from L2) instead of 6
    // full tiles and one of 16 rows whose CTAs idle while the others stream.  (2) A lone product with more tiles than
    // SMs and a K-major A (w1*x: 512 tiles = 3.46 waves on 148 SMs, i.e. the time of 4) is cut into a whole number
    // of waves of narrower tiles (586 tiles of 112 rows = 4 waves of 0.875 tile times): the TMA box has mm_pitch
    // rows, so a narrower tile moves proportionally fewer bytes.
    p.mm_pitch = TC_BM;
    if (tc_knobs().even_tiles) {
      p.mm_pitch = (int)((cdiv(MM, p.mm_tiles) + 7) / 8 * 8);
      const int64_t tiles = (int64_t)p.mm_tiles * p.nn_tiles;
      if (n == 1 && cls.ak && tiles > c.num_sms) {
        const int64_t waves = cdiv(tiles, c.num_sms);
        const int64_t want = waves * c.num_sms / p.nn_tiles;          // mm tiles that fill `waves` waves
        const int pitch = (int)((cdiv(MM, want) + 7) / 8 * 8);
        if (want > p.mm_tiles && pitch >= 64 && pitch < p.mm_pitch) p.mm_pitch = pitch;
      }
      if (p.mm_pitch > TC_BM) p.mm_pitch = TC_BM;
      p.mm_tiles = (int)cdiv(MM, p.mm_pitch);
    }
    bool ok;
    if (cls.ak) ok = make_map(&g.tmA[q], pa, j.K, MM, lda_, TC_BK, p.mm_pitch, true);
    else ok = make_map(&g.tmA[q], pa, MM, j.K, lda_, 32, TC_BK, false);
    if (cls.bk) ok = ok && make_map(&g.tmB[q], pb, j.K, NN, ldb_, TC_BK, TC_BN, true);
    else ok = ok && make_map(&g.tmB[q], pb, NN, j.K, ldb_, 32, TC_BK, false);
    AG_CHECK(ok, AG_ECUDA, "ag_gemm: cuTensorMapEncodeTiled failed");
    p.kblocks_total = (int)cdiv(j.K, TC_BK);
    p.MM = MM;
    p.NN = NN;
    p.transposed = cls.transposed ? 1 : 0;
    p.alpha = (float)j.alpha;
    p.beta = (float)j.beta;
    p.out2 = (float*)j.epi.out2;
    p.bias = (const float*)j.epi.bias;
    p.act = j.epi.act;
    total_tiles += (int64_t)p.mm_tiles * p.nn_tiles;
    if (p.kblocks_total < min_kblocks) min_kblocks = p.kblocks_total;
    any_beta = any_beta || j.beta != 0.0;
  }
  // split along K when the whole group has fewer tiles than SMs: one work item per CTA (a second partial wave would
  // leave most SMs idle for a whole item); the same number of splits for every problem of the group
  int splits = 1;
  if (total_tiles < c.num_sms && min_kblocks >= tc_knobs().split_kb) {
    splits = (int)(c.num_sms / total_tiles);
    const int maxs = min_kblocks / tc_knobs().min_kb;
    if (splits > maxs) splits = maxs;
    if (splits < 1) splits = 1;
  }
  const bool via_ws = splits > 1 || any_beta;
  size_t ws_floats = 0;
  for (int q = 0; q < n; ++q) {
    TcParams& p = g.p[q];
    p.kblocks_per_split = (int)cdiv(p.kblocks_total, splits);
    p.splits = (int)cdiv(p.kblocks_total, p.kblocks_per_split);
    if (via_ws) ws_floats += (size_t)p.splits * (size_t)(jobs[q].M * jobs[q].N);
  }
  float* ws = nullptr;
  if (via_ws) {
    ws = (float*)workspace(ws_floats * sizeof(float));
    if (!ws) return AG_ENOMEM;
  }
  static TcReduceGroup rg;
  size_t off = 0;
  int64_t work = 0;
  int64_t max_mn = 0;
  for (int q = 0; q < n; ++q) {
    TcParams& p = g.p[q];
    const TcJob& j = jobs[q];
    g.work_start[q] = (int)work;
    work += (int64_t)p.mm_tiles * p.nn_tiles * p.splits;
    if (via_ws) {
      p.out = ws + off;
      p.ld_out = j.M;
      p.split_stride = j.M * j.N;
      rg.ws[q] = ws + off;
      rg.splits[q] = p.splits;
      rg.M[q] = j.M;
      rg.N[q] = j.N;
      rg.alpha[q] = (float)j.alpha;
      rg.beta[q] = (float)j.beta;
      rg.bias[q] = p.bias;
      rg.C[q] = (float*)j.C;
      rg.C2[q] = p.out2;
      rg.ldc[q] = j.ldc;
      rg.act[q] = p.act;
      off += (size_t)p.splits * (size_t)(j.M * j.N);
      if (j.M * j.N > max_mn) max_mn = j.M * j.N;
    } else {
      p.out = (float*)j.C;
      p.ld_out = j.ldc;
      p.split_stride = 0;
    }
  }
  g.work_start[n] = (int)work;
  AG_CHECK(work < ((int64_t)1 << 31), AG_EUNSUPPORTED, "ag_gemm: too many tiles");
  const int64_t grid = work < c.num_sms ? work : c.num_sms;

#define LAUNCH(AK, BK_, FL)                                                                              \
  do {                                                                                                   \
    static bool attr_set = false;                                                                        \
    if (!attr_set) {                                                                                     \
      AG_CUDA(cudaFuncSetAttribute(k_gemm_tc<AK, BK_, FL>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                   TC_SMEM));                                                            \
      attr_set = true;                                                                                   \
    }                                                                                                    \
    k_gemm_tc<AK, BK_, FL><<<(unsigned)grid, TC_THREADS, TC_SMEM, c.stream>>>(g);                        \
  } while (0)
#define LAUNCH2(AK, BK_)                  \
  do {                                    \
    if (cls.skinny) LAUNCH(AK, BK_, 8);   \
    else LAUNCH(AK, BK_, 1);              \
  } while (0)
  if (cls.ak && cls.bk) LAUNCH2(true, true);
  else if (cls.ak && !cls.bk) LAUNCH2(true, false);
  else if (!cls.ak && cls.bk) LAUNCH2(false, true);
  else LAUNCH2(false, false);
#undef LAUNCH2
#undef LAUNCH
  AG_LAUNCHED();
  if (via_ws) {
    rg.nprob = n;
    dim3 grid_r((unsigned)cdiv(max_mn, 256), (unsigned)n);
    k_tc_splitk_reduce<<<grid_r, 256, 0, c.stream>>>(rg);
    AG_LAUNCHED();
  }
  return AG_OK;
}

#ifdef AGB_TC_TRACE
// ---- experiment (trace build only): how fast can ONE SM take data in?  Every CTA streams 16 KB "A" tiles of a large
// array by TMA through a 6-stage ring (as the product does) and, per tile, 8 KB of a small L2-resident "B" array
//   mode 0: not at all      mode 1: by TMA into the same stage      mode 2: by 256 threads with ld.global -> st.shared
// Reports ns per tile (block 0, %globaltimer).  tests/probe_tc_ingest.py.
__global__ void __launch_bounds__(320, 1)
    k_ingest(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const float* __restrict__ Bg,
             int mode, int tiles_per_cta, int pattern, unsigned long long* out_ns) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  constexpr int ST = 6, SB = 24576;
  const uint32_t bars = smem_base + ST * SB;
  auto full = [&](int s) { return bars + 8u * s; };
  auto empty = [&](int s) { return bars + 8u * (ST + s); };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < ST; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  unsigned long long t0 = 0;
  if (threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  if (warp == 0) {
    int s = 0;
    uint32_t ph = 0;
    for (int i = 0; i < tiles_per_cta; ++i) {
      mbar_wait(empty(s), ph ^ 1);
      if (elect_one()) {
        mbar_expect_tx(full(s), 16384 + (mode == 1 ? 8192 : 0));
        if (pattern == 0)       // 16 KB contiguous
          tma_load_2d(smem_base + s * SB, &tmA, full(s), 0, (blockIdx.x * tiles_per_cta + i) * 128);
        else                    // the layout of w1*x: 128 rows of 128 B at a 3136-byte pitch, 24 k-blocks per 128-row tile
          tma_load_2d(smem_base + s * SB, &tmA, full(s), (i % 24) * 32, (blockIdx.x * (tiles_per_cta / 24) + i / 24) * 128);
        if (mode == 1) tma_load_2d(smem_base + s * SB + 16384, &tmB, full(s), 0, (i % 8) * 64);
      }
      __syncwarp();
      if (++s == ST) { s = 0; ph ^= 1; }
    }
  } else if (warp >= 2) {                    // 8 consumer warps (256 threads)
    const int tid = threadIdx.x - 64;
    int s = 0;
    uint32_t ph = 0;
    float acc = 0.f;
    float4 pre[2];
    if (mode == 2) {
      pre[0] = __ldg(reinterpret_cast<const float4*>(Bg) + tid);
      pre[1] = __ldg(reinterpret_cast<const float4*>(Bg) + tid + 256);
    }
    for (int i = 0; i < tiles_per_cta; ++i) {
      mbar_wait(full(s), ph);
      const float4* a4 = reinterpret_cast<const float4*>(smem_gen + s * SB);
      float4 v = a4[tid], w = a4[tid + 256], u = a4[tid + 512], x = a4[tid + 768];      // read the 16 KB
      acc += v.x + w.y + u.z + x.w;
      if (mode == 2) {
        float4* b4 = reinterpret_cast<float4*>(smem_gen + s * SB + 16384);
        b4[tid] = pre[0];
        b4[tid + 256] = pre[1];
        const float4* src = reinterpret_cast<const float4*>(Bg) + (size_t)(((i + 1) % 8) * 512);
        pre[0] = __ldg(src + tid);
        pre[1] = __ldg(src + tid + 256);
      } else if (mode == 1) {
        const float4* b4 = reinterpret_cast<const float4*>(smem_gen + s * SB + 16384);
        acc += b4[tid].x + b4[tid + 256].y;
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(empty(s));
      if (++s == ST) { s = 0; ph ^= 1; }
    }
    if (acc == 123.456f) out_ns[1] = 1;       // keep the reads
  }
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    out_ns[0] = t1 - t0;
  }
}

extern "C" int ag_debug_ingest(int mode, int tiles_per_cta, int pattern, const void* A, long long rowsA, const void* B,
                               double* ns_per_tile) {
  Context& c = ctx();
  CUtensorMap ma, mb;
  // A: rowsA rows of 32 floats (K-major rows, 128 B each, contiguous); B: 512 rows of 32 floats
  const int kA = pattern == 0 ? 32 : 784;
  if (!make_map(&ma, A, kA, rowsA, kA, 32, 128, true) || !make_map(&mb, B, 32, 512, 32, 32, 64, true)) return 3;
  unsigned long long* d = nullptr;
  if (cudaMalloc(&d, 16) != cudaSuccess) return 2;
  cudaMemset(d, 0, 16);
  cudaFuncSetAttribute(k_ingest, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * 24576 + 2048);
  const int grid = c.num_sms;
  for (int rep = 0; rep < 3; ++rep)
    k_ingest<<<grid, 320, 6 * 24576 + 2048, c.stream>>>(ma, mb, (const float*)B, mode, tiles_per_cta, pattern, d);
  unsigned long long h[2] = {0, 0};
  cudaStreamSynchronize(c.stream);
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  cudaFree(d);
  *ns_per_tile = (double)h[0] / tiles_per_cta;
  return cudaGetLastError() == cudaSuccess ? 0 : 3;
}
#endif

#ifdef AGB_TC_TRACE
extern "C" int ag_gemm_trace(unsigned long long* out) {      // 6 x 256 stamps of the last launch's block 0
  cudaDeviceSynchronize();
  return cudaMemcpyFromSymbol(out, g_tc_trace, sizeof(unsigned long long) * 6 * 256) == cudaSuccess ? 0 : 3;
}
#endif

ag_status gemm_tc(int transA, int transB, int64_t M, int64_t N, int64_t K, double alpha, const void* A,
                  int64_t lda, const void* B, int64_t ldb, double beta, void* C, int64_t ldc,
                  const TcEpilogue* epi) {
  TcJob j;
  j.transA = transA; j.transB = transB; j.M = M; j.N = N; j.K = K; j.alpha = alpha; j.beta = beta;
  j.A = A; j.lda = lda; j.B = B; j.ldb = ldb; j.C = C; j.ldc = ldc;
  j.epi.bias = epi ? epi->bias : nullptr;
  j.epi.act = epi ? epi->act : TC_ACT_NONE;
  j.epi.out2 = epi ? epi->out2 : nullptr;
  return gemm_tc_group(1, &j);
}

}  // namespace agb
