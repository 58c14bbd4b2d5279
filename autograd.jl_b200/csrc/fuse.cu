// fuse.cu -- ONE kernel for a whole chain of elementwise tape operations (horizontal fusion).
//
// Reference: every dotted call on a tracked value is its own Base broadcast pass (src/core.jl:68-70 materialises
// each `Broadcasted`), and every gradient rule of src/math.jl:9-74 / src/base.jl:20-49,73-74 is another pass, as is
// the out-of-place `addto!` (src/addto.jl:23).  The host runtime (autograd.jl_b200/lazy.py) defers such operations,
// and when a value is finally needed hands the whole expression DAG to this kernel: every leaf array is read
// once, every value the tape still refers to is written once, temporaries of the gradient expressions never
// reach memory.
//
// The program is a small bytecode held in the kernel parameters (uniform for the grid, decoded once per 128-bit
// vector).  It is a stack machine whose top of stack lives in a fixed register set and whose slots are registers
// selected by warp-uniform switches -- no local memory (check -Xptxas -v: 0 bytes stack frame).  The arithmetic
// of every instruction is the SAME expression as the single-rule kernels (ewise_ops.cuh), so a fused chain is
// bit-identical to the chain of separate launches.
//
// Operands: full arrays of the (collapsed) rows x cols iteration space, column vectors (rows; index i0), row
// vectors (cols; index i1), device scalars and immediates -- the shapes Julia broadcasting produces on the hot
// path (bias add, softmax normaliser, scalar dy).  HBM-bound: algorithmic bytes = (full inputs + outputs)*n*s.
#include "common.cuh"
#include "ewise_ops.cuh"

#include <stdlib.h>
#include <utility>

namespace agb {

constexpr int FZ_MAXCODE = 96, FZ_MAXCONST = 16, FZ_MAXIN = 12, FZ_MAXOUT = 12, FZ_SLOTS = 8;
// what the general interpreter handles; larger programs run only as specialised kernels
constexpr int FZ_DYN_MAXIN = 8, FZ_DYN_MAXOUT = 6;

// instruction word: op | a << 8 | b << 16
enum {
  FZ_LD_IN = 0,      // a = input, b = 1: spill the top of stack first (push), 0: overwrite it
  FZ_LD_CONST = 1,   // a = constant, b = spill
  FZ_LD_TMP = 2,     // a = slot, b = spill
  FZ_ST_TMP = 3,     // slot[a] = top (kept)
  FZ_UNARY = 4,      // top = f_a(top)
  FZ_BIN = 5,        // top = pop() (a) top
  FZ_BIN_IN = 6,     // top = top (a) in[b & 0x7f]   (b & 0x80: reversed operand order)
  FZ_BIN_CONST = 7,  // top = top (a) const[b & 0x7f]
  FZ_UGRAD = 8,      // y = top, x = pop(), dy = pop(): top = dy * g_a(x, y)
  FZ_TERN = 9,       // q = top, p = pop(), dy = pop(): top = G_a(dy, p, q)
  FZ_STORE = 10,     // out[a] = top
  FZ_REDUCE = 11,    // the top is the operand of the fused sum (ag_ewise_fused_reduce); it stays on top
  FZ_NOPS = 12
};
enum { FZ_M_FULL = 0, FZ_M_DEVSCALAR = 1, FZ_M_IMM = 2, FZ_M_COLVEC = 3, FZ_M_ROWVEC = 4,
       // internal (set by the host side below when rows % lanes == 0: a vector never straddles two columns)
       FZ_M_COLVEC_V = 5,   // the lanes are consecutive rows: one aligned vector load
       FZ_M_ROWVEC_1 = 6 }; // the lanes share their column: one scalar load

struct FzArgs {
  int32_t code[FZ_MAXCODE];
  double consts[FZ_MAXCONST];
  const void* in[FZ_MAXIN];
  double imm[FZ_MAXIN];
  int32_t mode[FZ_MAXIN];
  void* out[FZ_MAXOUT];
  int32_t ncode, nin, nout, has_bcast;
  int64_t n, rows;
  int32_t zero;   // always 0, but only known at run time: see fz_opaque
  // fused sum of the FZ_REDUCE operand (specialised kernels only)
  int32_t rmode;  // 0 none, 1 whole array -> double partial per block, 2 contiguous runs of rlen -> final values,
                  // 3 row sums of an rlen x (n/rlen) matrix -> rlen double partials per block
  int32_t chunk;  // elements per block (a multiple of rlen for modes 2 and 3)
  int32_t tps;    // mode 2: threads that share one run (power of two)
  int64_t rlen;
  void* rout;
  // in-kernel finish of the block partials (modes 1 and 3): groups of FZ_FIN_GROUP blocks, two levels of tickets
  int32_t* tickets;   // [0] = groups finished, [1 + g] = blocks of group g finished; null: a combine launch follows
  double* lvl2;       // one set of rlen (mode 3) / 1 (mode 1) sums per group
  void* rfinal;       // the result (rlen or 1 values of T)
};

#define FZ_CASE8(X) X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7)

template <typename T, int N>
__device__ __forceinline__ void fz_coords(const FzArgs& a, int64_t i, int64_t* c0, int64_t* c1) {
  int64_t q, r;
  if (a.n < ((int64_t)1 << 31)) {
    const uint32_t qq = (uint32_t)i / (uint32_t)a.rows;
    q = qq;
    r = (uint32_t)i - qq * (uint32_t)a.rows;
  } else {
    q = i / a.rows;
    r = i - q * a.rows;
  }
#pragma unroll
  for (int j = 0; j < N; ++j) {
    c0[j] = r;
    c1[j] = q;
    if (++r == a.rows) { r = 0; ++q; }
  }
}

template <typename T, int N, bool VEC>
__device__ __forceinline__ void fz_load(const FzArgs& a, int k, int64_t i, const int64_t* c0, const int64_t* c1,
                                        T* dst) {
  const T* __restrict__ p = (const T*)a.in[k];
  const int m = a.mode[k];
  if (m == FZ_M_FULL) {
    if (VEC) {
      const Pack<T> t = *reinterpret_cast<const Pack<T>*>(p + i);
#pragma unroll
      for (int j = 0; j < N; ++j) dst[j] = t.v[j];
    } else {
      dst[0] = p[i];
    }
  } else if (VEC && m == FZ_M_COLVEC_V) {
    const Pack<T> t = *reinterpret_cast<const Pack<T>*>(p + c0[0]);
#pragma unroll
    for (int j = 0; j < N; ++j) dst[j] = t.v[j];
  } else if (m == FZ_M_ROWVEC_1) {
    const T t = p[c1[0]];
#pragma unroll
    for (int j = 0; j < N; ++j) dst[j] = t;
  } else if (m == FZ_M_COLVEC) {
#pragma unroll
    for (int j = 0; j < N; ++j) dst[j] = p[c0[j]];
  } else if (m == FZ_M_ROWVEC) {
#pragma unroll
    for (int j = 0; j < N; ++j) dst[j] = p[c1[j]];
  } else {
    const T sc = m == FZ_M_DEVSCALAR ? p[0] : (T)a.imm[k];
#pragma unroll
    for (int j = 0; j < N; ++j) dst[j] = sc;
  }
}

template <typename T, int N, bool VEC>
__device__ __forceinline__ void fz_store(const FzArgs& a, int o, int64_t i, const T* v) {
  T* __restrict__ p = (T*)a.out[o];
  if (VEC) {
    Pack<T> t;
#pragma unroll
    for (int j = 0; j < N; ++j) t.v[j] = v[j];
    *reinterpret_cast<Pack<T>*>(p + i) = t;
  } else {
    p[i] = v[0];
  }
}

// ---- general evaluator: bytecode interpreted at run time -------------------------------------------------
// NI input register sets and NS slots are template parameters (2/2, 4/4, 8/8): the register budget decides how
// many threads -- and so how many bytes in flight -- an SM holds, which is what bounds an elementwise kernel.
// UN vectors per thread share one decode of every instruction (the interpreter is issue-bound, not memory-bound).
template <typename T, bool VEC, int NI, int NS, int UN, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) k_fused(const __grid_constant__ FzArgs a) {
  constexpr int NV = VEC ? Pack<T>::N : 1;        // elements per vector
  constexpr int N = NV * UN;                      // lanes of this thread
  const int64_t base = ((int64_t)blockIdx.x * (kThreads * UN) + threadIdx.x) * NV;
  if (base >= a.n) return;
  bool on[UN];

  // all leaf loads are issued before any arithmetic
  T inv[NI][N];
#pragma unroll
  for (int u = 0; u < UN; ++u) {
    const int64_t i = base + (int64_t)u * kThreads * NV;
    on[u] = i < a.n;
    if (on[u]) {
      int64_t c0[NV], c1[NV];
      if (a.has_bcast) fz_coords<T, NV>(a, i, c0, c1);
#pragma unroll
      for (int k = 0; k < NI; ++k)
        if (k < a.nin) fz_load<T, NV, VEC>(a, k, i, c0, c1, &inv[k][u * NV]);
    } else {
#pragma unroll
      for (int k = 0; k < NI; ++k)
#pragma unroll
        for (int j = 0; j < NV; ++j) inv[k][u * NV + j] = T(0);
    }
  }

  T tos[N], S[NS][N];
#pragma unroll
  for (int j = 0; j < N; ++j) tos[j] = T(0);
  int sp = 0;

#define FZ_SLOT_ST_CASE(s) case s: if constexpr (s < NS) { _Pragma("unroll") for (int j = 0; j < N; ++j) S[s][j] = tos[j]; } break;
#define FZ_SLOT_ST(idx) switch (idx) { FZ_CASE8(FZ_SLOT_ST_CASE) }
#define FZ_SLOT_LD_CASE(s) case s: if constexpr (s < NS) { _Pragma("unroll") for (int j = 0; j < N; ++j) dst_[j] = S[s][j]; } break;
#define FZ_SLOT_LD(dst, idx) { T* dst_ = dst; switch (idx) { FZ_CASE8(FZ_SLOT_LD_CASE) } }
#define FZ_IN_LD_CASE(s) case s: if constexpr (s < NI) { _Pragma("unroll") for (int j = 0; j < N; ++j) dst_[j] = inv[s][j]; } break;
#define FZ_IN_LD(dst, idx) { T* dst_ = dst; switch (idx) { FZ_CASE8(FZ_IN_LD_CASE) } }
#define FZ_BIN_CASE(ID, F) case ID: _Pragma("unroll") for (int j = 0; j < N; ++j) { const T ab[2] = {l_[j], r_[j]}; tos[j] = F<T>::apply(ab); } break;
#define FZ_APPLY_BIN(b, l, r)                                                                        \
  {                                                                                                  \
    const T* l_ = l; const T* r_ = r;                                                                \
    switch (b) {                                                                                     \
      FZ_BIN_CASE(AG_B_ADD, FAdd) FZ_BIN_CASE(AG_B_SUB, FSub) FZ_BIN_CASE(AG_B_MUL, FMul)            \
      FZ_BIN_CASE(AG_B_DIV, FDiv) FZ_BIN_CASE(AG_B_MAX, FMax) FZ_BIN_CASE(AG_B_MIN, FMin)            \
      FZ_BIN_CASE(AG_B_POW, FPow) FZ_BIN_CASE(AG_B_EQ, FEq) FZ_BIN_CASE(AG_B_LT, FLt)                \
      FZ_BIN_CASE(AG_B_LE, FLe) FZ_BIN_CASE(AG_B_GT, FGt) FZ_BIN_CASE(AG_B_GE, FGe)                  \
      FZ_BIN_CASE(AG_B_ATAN2, FAtan2) FZ_BIN_CASE(AG_B_HYPOT, FHypot)                                \
    }                                                                                                \
  }

  for (int pc = 0; pc < a.ncode; ++pc) {
    const uint32_t w = (uint32_t)a.code[pc];
    const int op = w & 0xff, x = (w >> 8) & 0xff, y = (w >> 16) & 0xff;
    switch (op) {
      case FZ_LD_IN:
        if (y) { FZ_SLOT_ST(sp) ++sp; }
        FZ_IN_LD(tos, x)
        break;
      case FZ_LD_CONST: {
        if (y) { FZ_SLOT_ST(sp) ++sp; }
        const T cv = (T)a.consts[x];
#pragma unroll
        for (int j = 0; j < N; ++j) tos[j] = cv;
        break;
      }
      case FZ_LD_TMP:
        if (y) { FZ_SLOT_ST(sp) ++sp; }
        FZ_SLOT_LD(tos, x)
        break;
      case FZ_ST_TMP:
        FZ_SLOT_ST(x)
        break;
      case FZ_UNARY:
        switch (x) {
#define X(OP) case OP: _Pragma("unroll") for (int j = 0; j < N; ++j) tos[j] = U<OP, T>::f(tos[j]); break;
          AG_FOR_UNARY(X)
#undef X
        }
        break;
      case FZ_BIN: {
        T nos[N];
        --sp;
        FZ_SLOT_LD(nos, sp)
        T cur[N];
#pragma unroll
        for (int j = 0; j < N; ++j) cur[j] = tos[j];
        FZ_APPLY_BIN(x, nos, cur)
        break;
      }
      case FZ_BIN_IN:
      case FZ_BIN_CONST: {
        T o[N], cur[N];
        if (op == FZ_BIN_IN) {
          FZ_IN_LD(o, y & 0x7f)
        } else {
          const T cv = (T)a.consts[y & 0x7f];
#pragma unroll
          for (int j = 0; j < N; ++j) o[j] = cv;
        }
#pragma unroll
        for (int j = 0; j < N; ++j) cur[j] = tos[j];
        if (y & 0x80) { FZ_APPLY_BIN(x, o, cur) } else { FZ_APPLY_BIN(x, cur, o) }
        break;
      }
      case FZ_UGRAD: {
        T gx[N], gdy[N];
        --sp;
        FZ_SLOT_LD(gx, sp)
        --sp;
        FZ_SLOT_LD(gdy, sp)
        switch (x) {
#define X(OP) case OP: _Pragma("unroll") for (int j = 0; j < N; ++j) tos[j] = gdy[j] * U<OP, T>::g(U<OP, T>::needs_x ? gx[j] : T(0), U<OP, T>::needs_y ? tos[j] : T(0)); break;
          AG_FOR_UNARY(X)
#undef X
        }
        break;
      }
      case FZ_TERN: {
        T gp[N], gdy[N];
        --sp;
        FZ_SLOT_LD(gp, sp)
        --sp;
        FZ_SLOT_LD(gdy, sp)
        switch (x) {
#define FZ_TERN_CASE(ID, G) case ID: _Pragma("unroll") for (int j = 0; j < N; ++j) { const T t3[3] = {gdy[j], gp[j], tos[j]}; tos[j] = G<T>::apply(t3); } break;
          FZ_TERN_CASE(0, GDiv2) FZ_TERN_CASE(1, GEqMask) FZ_TERN_CASE(2, GPow1) FZ_TERN_CASE(3, GPow2)
          FZ_TERN_CASE(4, GAtan1) FZ_TERN_CASE(5, GAtan2) FZ_TERN_CASE(6, GHypot)
        }
        break;
      }
      case FZ_STORE: {
        const int o = x < FZ_DYN_MAXOUT ? x : FZ_DYN_MAXOUT - 1;
#pragma unroll
        for (int u = 0; u < UN; ++u)
          if (on[u]) fz_store<T, NV, VEC>(a, o, base + (int64_t)u * kThreads * NV, &tos[u * NV]);
        break;
      }
    }
  }
}

// ---- specialised evaluators: the same machine with the program as a compile-time constant ------------------
// The programs the BASELINE configs produce (generated list, fuse_static.inc) are instantiated ahead of time:
// every switch above folds away, stack slots become plain registers, two vectors per thread are in flight.
struct StaticProg {
  int n, nin;
  uint32_t code[FZ_MAXCODE];
};

#include "fuse_static.inc"

constexpr int fz_sp_before(const StaticProg& p, int pc) {
  int sp = 0;
  for (int i = 0; i < pc; ++i) {
    const int op = p.code[i] & 0xff, y = (p.code[i] >> 16) & 0xff;
    if (op <= FZ_LD_TMP && y) ++sp;
    if (op == FZ_BIN) --sp;
    if (op == FZ_UGRAD || op == FZ_TERN) sp -= 2;
  }
  return sp;
}
constexpr int fz_slots_needed(const StaticProg& p) {
  int need = 1;
  for (int i = 0; i < p.n; ++i) {
    const int op = p.code[i] & 0xff, x = (p.code[i] >> 8) & 0xff;
    const int sp = fz_sp_before(p, i + 1);
    if (sp > need) need = sp;
    if ((op == FZ_LD_TMP || op == FZ_ST_TMP) && x + 1 > need) need = x + 1;
  }
  return need;
}

// A value that crosses an instruction boundary must be the ROUNDED result of that instruction.  Left alone, ptxas
// contracts the `mul.f32` that ends one instruction with the `add.f32` that starts the next into an FFMA (an empty
// asm barrier stops NVVM but not ptxas), and the fused chain would differ in the last bit from the sequence of
// single kernels.  XOR with a zero that is only known at run time is a bit-exact identity no optimiser can remove.
__device__ __forceinline__ void fz_opaque(float& v, int zero) { v = __int_as_float(__float_as_int(v) ^ zero); }
__device__ __forceinline__ void fz_opaque(double& v, int zero) {
  v = __longlong_as_double(__double_as_longlong(v) ^ (long long)zero);
}

template <int B, typename T> struct BinSel;
#define FZ_BINSEL(ID, F) template <typename T> struct BinSel<ID, T> { using type = F<T>; };
FZ_BINSEL(AG_B_ADD, FAdd) FZ_BINSEL(AG_B_SUB, FSub) FZ_BINSEL(AG_B_MUL, FMul) FZ_BINSEL(AG_B_DIV, FDiv)
FZ_BINSEL(AG_B_MAX, FMax) FZ_BINSEL(AG_B_MIN, FMin) FZ_BINSEL(AG_B_POW, FPow) FZ_BINSEL(AG_B_EQ, FEq)
FZ_BINSEL(AG_B_LT, FLt) FZ_BINSEL(AG_B_LE, FLe) FZ_BINSEL(AG_B_GT, FGt) FZ_BINSEL(AG_B_GE, FGe)
FZ_BINSEL(AG_B_ATAN2, FAtan2) FZ_BINSEL(AG_B_HYPOT, FHypot)
template <int G, typename T> struct TernSel;
#define FZ_TERNSEL(ID, F) template <typename T> struct TernSel<ID, T> { using type = F<T>; };
FZ_TERNSEL(0, GDiv2) FZ_TERNSEL(1, GEqMask) FZ_TERNSEL(2, GPow1) FZ_TERNSEL(3, GPow2) FZ_TERNSEL(4, GAtan1)
FZ_TERNSEL(5, GAtan2) FZ_TERNSEL(6, GHypot)

template <int PROG, int PC, typename T, int N, int NIN, int NSL>
__device__ __forceinline__ void fz_step(const FzArgs& a, int64_t i, const T (&inv)[NIN][N], T (&S)[NSL][N], T (&tos)[N],
                                        T* red_at, double& acc) {
  constexpr uint32_t w = kProgs[PROG].code[PC];
  constexpr int op = w & 0xff, x = (w >> 8) & 0xff, y = (w >> 16) & 0xff;
  constexpr int sp = fz_sp_before(kProgs[PROG], PC);
  if constexpr (op <= FZ_LD_TMP) {
    if constexpr (y != 0) {
#pragma unroll
      for (int j = 0; j < N; ++j) S[sp][j] = tos[j];
    }
    if constexpr (op == FZ_LD_IN) {
#pragma unroll
      for (int j = 0; j < N; ++j) tos[j] = inv[x][j];
    } else if constexpr (op == FZ_LD_CONST) {
      const T cv = (T)a.consts[x];
#pragma unroll
      for (int j = 0; j < N; ++j) tos[j] = cv;
    } else {
#pragma unroll
      for (int j = 0; j < N; ++j) tos[j] = S[x][j];
    }
  } else if constexpr (op == FZ_ST_TMP) {
#pragma unroll
    for (int j = 0; j < N; ++j) S[x][j] = tos[j];
  } else if constexpr (op == FZ_UNARY) {
#pragma unroll
    for (int j = 0; j < N; ++j) tos[j] = U<x, T>::f(tos[j]);
  } else if constexpr (op == FZ_BIN) {
#pragma unroll
    for (int j = 0; j < N; ++j) {
      const T ab[2] = {S[sp - 1][j], tos[j]};
      tos[j] = BinSel<x, T>::type::apply(ab);
    }
  } else if constexpr (op == FZ_BIN_IN || op == FZ_BIN_CONST) {
    constexpr int k = y & 0x7f;
    constexpr bool rev = (y & 0x80) != 0;
#pragma unroll
    for (int j = 0; j < N; ++j) {
      T o;
      if constexpr (op == FZ_BIN_IN) o = inv[k][j];
      else o = (T)a.consts[k];
      const T ab[2] = {rev ? o : tos[j], rev ? tos[j] : o};
      tos[j] = BinSel<x, T>::type::apply(ab);
    }
  } else if constexpr (op == FZ_UGRAD) {
#pragma unroll
    for (int j = 0; j < N; ++j)
      tos[j] = S[sp - 2][j] * U<x, T>::g(U<x, T>::needs_x ? S[sp - 1][j] : T(0), U<x, T>::needs_y ? tos[j] : T(0));
  } else if constexpr (op == FZ_TERN) {
#pragma unroll
    for (int j = 0; j < N; ++j) {
      const T t3[3] = {S[sp - 2][j], S[sp - 1][j], tos[j]};
      tos[j] = TernSel<x, T>::type::apply(t3);
    }
  } else if constexpr (op == FZ_STORE) {
    fz_store<T, N, true>(a, x, i, tos);
  } else if constexpr (op == FZ_REDUCE) {
#pragma unroll
    for (int j = 0; j < N; ++j) {
      red_at[j] = tos[j];
      acc += (double)tos[j];
    }
  }
  if constexpr (op == FZ_UNARY || op == FZ_BIN || op == FZ_BIN_IN || op == FZ_BIN_CONST || op == FZ_UGRAD ||
                op == FZ_TERN) {
#pragma unroll
    for (int j = 0; j < N; ++j) fz_opaque(tos[j], a.zero);
  }
}

template <int PROG, typename T, int N, int NIN, int NSL, int... PC>
__device__ __forceinline__ void fz_run(const FzArgs& a, int64_t i, const T (&inv)[NIN][N], T (&S)[NSL][N], T (&tos)[N],
                                       T* red_at, double& acc, std::integer_sequence<int, PC...>) {
  (fz_step<PROG, PC, T, N, NIN, NSL>(a, i, inv, S, tos, red_at, acc), ...);
}

constexpr bool fz_has_reduce(const StaticProg& p) {
  for (int i = 0; i < p.n; ++i)
    if ((p.code[i] & 0xff) == FZ_REDUCE) return true;
  return false;
}

constexpr int FZ_FIN_GROUP = 32, FZ_FIN_MAXGROUPS = 4096;

// The block partials of a fused sum (one double, or R row sums, per block) can be finished inside the kernel: the last
// block of every group of FZ_FIN_GROUP blocks adds the group's partials in block order, the last group to finish adds
// the group sums in group order -- fixed order, so the result does not depend on the schedule, and no second launch.
// OFF by default (AGB_FUSED_FINISH=1 turns it on): measured on the B200 the fence + ticket of every block cost more
// than the combine launch they save inside a CUDA graph (MNIST step 19 -> 16 launches but 183 -> 193 us; the relu
// gradient with its row sums 20.3 -> 24.2 us; sweep 2^28 elements 2.13 -> 2.34 ms).
template <typename T>
__device__ __forceinline__ void fz_finish_partials(const FzArgs& a, int R) {
  __shared__ int s_last;
  const int tid = threadIdx.x;
  const int nblk = gridDim.x, g = blockIdx.x / FZ_FIN_GROUP, ngroups = (nblk + FZ_FIN_GROUP - 1) / FZ_FIN_GROUP;
  const int b0 = g * FZ_FIN_GROUP, b1 = min(nblk, b0 + FZ_FIN_GROUP);
  const double* __restrict__ part = (const double*)a.rout;
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = atomicAdd(&a.tickets[1 + g], 1) == (b1 - b0) - 1;
  __syncthreads();
  if (!s_last) return;
  for (int r = tid; r < R; r += kThreads) {
    double s = 0;
    for (int b = b0; b < b1; ++b) s += __ldcg(part + (int64_t)b * R + r);
    a.lvl2[(int64_t)g * R + r] = s;
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    a.tickets[1 + g] = 0;
    s_last = atomicAdd(&a.tickets[0], 1) == ngroups - 1;
  }
  __syncthreads();
  if (!s_last) return;
  for (int r = tid; r < R; r += kThreads) {
    double s0 = 0, s1 = 0;
    int q = 0;
    for (; q + 1 < ngroups; q += 2) {
      s0 += __ldcg(a.lvl2 + (int64_t)q * R + r);
      s1 += __ldcg(a.lvl2 + (int64_t)(q + 1) * R + r);
    }
    if (q < ngroups) s0 += __ldcg(a.lvl2 + (int64_t)q * R + r);
    ((T*)a.rfinal)[r] = (T)(s0 + s1);
  }
  if (tid == 0) a.tickets[0] = 0;
}

// Sum epilogue of a specialised kernel: `red` holds the block's chunk of the FZ_REDUCE operand (zeros beyond the
// data), `acc` this thread's own sum.  Fixed summation order throughout (deterministic), accumulation in double.
template <typename T, int E>
__device__ __forceinline__ void fz_reduce_epilogue(const FzArgs& a, const T* red, double acc) {
  __shared__ double part[kThreads];
  const int tid = threadIdx.x;
  if (a.rmode == 1) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((tid & 31) == 0) part[tid >> 5] = acc;
    __syncthreads();
    if (tid == 0) {
      double r = 0;
#pragma unroll
      for (int w = 0; w < kThreads / 32; ++w) r += part[w];
      ((double*)a.rout)[blockIdx.x] = r;
    }
    if (a.tickets) fz_finish_partials<T>(a, 1);
  } else if (a.rmode == 2) {
    const int rlen = (int)a.rlen, nrun = a.chunk / rlen, tps = a.tps;
    const int g = tid / tps, q = tid - g * tps;
    double s = 0;
    if (g < nrun)
      for (int k = q; k < rlen; k += tps) s += (double)red[g * rlen + k];
    part[tid] = s;
    __syncthreads();
    const int64_t run = (int64_t)blockIdx.x * nrun + g;
    if (q == 0 && g < nrun && run * a.rlen < a.n) {
      double r = 0;
      for (int k = 0; k < tps; ++k) r += part[g * tps + k];
      ((T*)a.rout)[run] = (T)r;
    }
  } else {
    const int R = (int)a.rlen, ncol = a.chunk / R;
    for (int r = tid; r < R; r += kThreads) {
      double s = 0;
      for (int c = 0; c < ncol; ++c) s += (double)red[r + R * c];
      ((double*)a.rout)[(int64_t)blockIdx.x * R + r] = s;
    }
    if (a.tickets) fz_finish_partials<T>(a, R);
  }
}

// vectors per thread of a specialised kernel: two, or one for programs with many operands (register budget)
constexpr int fz_un(const StaticProg& p) {
  int stores = 0;
  for (int i = 0; i < p.n; ++i)
    if ((p.code[i] & 0xff) == FZ_STORE) ++stores;
  return (p.nin > 6 || stores > 6) ? 1 : 2;
}

template <typename T, int PROG>
__global__ void __launch_bounds__(kThreads) k_fused_static(const __grid_constant__ FzArgs a) {
  constexpr int N = Pack<T>::N;
  constexpr int NIN = kProgs[PROG].nin > 0 ? kProgs[PROG].nin : 1;
  constexpr int NSL = fz_slots_needed(kProgs[PROG]);
  constexpr bool RED = fz_has_reduce(kProgs[PROG]);
  constexpr int FZ_UN = fz_un(kProgs[PROG]);
  constexpr int E = kThreads * FZ_UN * N;                       // elements a block can hold
  __shared__ __align__(16) T red[RED ? E : 1];
  const int chunk = RED ? a.chunk : E;                          // a reducing launch may use a shorter chunk
  const int64_t blk0 = (int64_t)blockIdx.x * chunk;
  T inv[FZ_UN][NIN][N];
  bool on[FZ_UN];
#pragma unroll
  for (int u = 0; u < FZ_UN; ++u) {
    const int l = (u * kThreads + (int)threadIdx.x) * N;
    const int64_t i = blk0 + l;
    on[u] = l < chunk && i < a.n;
    if (on[u]) {
      int64_t c0[N], c1[N];
      if (a.has_bcast) fz_coords<T, N>(a, i, c0, c1);
#pragma unroll
      for (int k = 0; k < kProgs[PROG].nin; ++k) fz_load<T, N, true>(a, k, i, c0, c1, inv[u][k]);
    }
  }
  double acc = 0;
#pragma unroll
  for (int u = 0; u < FZ_UN; ++u) {
    const int l = (u * kThreads + (int)threadIdx.x) * N;
    if (on[u]) {
      T tos[N], S[NSL][N];
#pragma unroll
      for (int j = 0; j < N; ++j) tos[j] = T(0);
      fz_run<PROG, T, N, NIN, NSL>(a, blk0 + l, inv[u], S, tos, RED ? red + l : red, acc,
                                   std::make_integer_sequence<int, kProgs[PROG].n>{});
    } else if (RED) {
#pragma unroll
      for (int j = 0; j < N; ++j) red[l + j] = T(0);
    }
  }
  if constexpr (RED) {
    __syncthreads();
    fz_reduce_epilogue<T, E>(a, red, acc);
  }
}

template <typename T>
static bool launch_static(int prog, const FzArgs& a, unsigned nb, cudaStream_t st) {
  switch (prog) {
#define X(P) case P: k_fused_static<T, P><<<nb, kThreads, 0, st>>>(a); return true;
    FZ_STATIC_LIST(X)
#undef X
  }
  return false;
}

static int static_un(int prog) {
  switch (prog) {
#define X(P) case P: return fz_un(kProgs[P]);
    FZ_STATIC_LIST(X)
#undef X
  }
  return 2;
}

static int find_static(const int32_t* code, int ncode) {
  for (int p = 0; p < FZ_NSTATIC; ++p) {
    if (kProgs[p].n != ncode) continue;
    bool same = true;
    for (int k = 0; k < ncode && same; ++k) same = kProgs[p].code[k] == (uint32_t)code[k];
    if (same) return p;
  }
  return -1;
}

template <typename T, bool VEC>
static void launch_dynamic(int ni, int ns, const FzArgs& a, unsigned nb, cudaStream_t st) {
  // nb is given for one vector per thread
  if (ni <= 2 && ns <= 2) k_fused<T, VEC, 2, 2, 2, 2><<<(nb + 1) / 2, kThreads, 0, st>>>(a);
  else if (ni <= 4 && ns <= 4) k_fused<T, VEC, 4, 4, 1, 3><<<nb, kThreads, 0, st>>>(a);
  else k_fused<T, VEC, 8, 8, 1, 2><<<nb, kThreads, 0, st>>>(a);
}

// ---- column programs (ag_ewise_colprog) ------------------------------------------------------------------
// A chain that contains sums over the rows of an (m, n) matrix and uses them again: the softmax normaliser
// `ypred .- log.(sum(exp.(ypred), dims=1))` of every model of the reference (examples/mnist.jl:40,
// examples/charlm.jl:107-110) and the unbroadcast of its gradient (src/unbroadcast.jl:17-24).  The program is a
// sequence of segments: an ELEMENT segment runs for every row of a column, a COLUMN segment once per column.  G
// lanes share a column, rows along the lanes (G = 8, 16 or 32: the smallest that gives a short column one row per
// lane; consecutive lanes read consecutive rows: coalesced); a column sum is accumulated in double by the element
// segment (CP_CRED), combined over the G lanes and rounded by the next column segment (CP_CFIN).  Instruction set = the one above plus:
enum {
  CP_PHASE = 12,   // a = 0: element segment follows, 1: column segment follows
  CP_CRED = 13,    // column accumulator a += top
  CP_LD_COL = 14,  // top = column value a (b = 1: push first)
  CP_ST_COL = 15,  // column value a = top
  CP_CFIN = 16,    // column value a = top = (T) sum over the lanes of accumulator a; the accumulator is cleared
  CP_NOPS = 17
};
constexpr int CP_COLS = 8;

struct CpArgs {
  int32_t code[FZ_MAXCODE];
  double consts[FZ_MAXCONST];
  const void* in[FZ_DYN_MAXIN];
  int32_t mode[FZ_DYN_MAXIN];
  void* out[FZ_DYN_MAXOUT];
  int32_t ncode, nin, nout;
  int64_t m, n;          // rows per column, columns
  double* part;          // per-block partial of the REDUCE operand (or null)
  int32_t* ticket;
  void* rout;            // final sum (one T)
  int32_t zero;          // always 0, but only known at run time: see fz_opaque
};

template <typename T>
__device__ __forceinline__ T cp_load(const CpArgs& a, int k, int64_t r, int64_t col) {
  const T* __restrict__ p = (const T*)a.in[k];
  switch (a.mode[k]) {
    case FZ_M_FULL: return p[r + a.m * col];
    case FZ_M_COLVEC: return p[r];
    case FZ_M_ROWVEC: return p[col];
    default: return p[0];
  }
}

// one segment [pc0, pc1) for element (r, col) (ELEM) or for column col (!ELEM)
template <typename T, bool ELEM>
__device__ __forceinline__ void cp_run(const CpArgs& a, int pc0, int pc1, int64_t r, int64_t col, bool writer,
                                       T* colval, double* colacc, double& racc) {
  T tos = T(0), S[FZ_SLOTS];
  int sp = 0;
  for (int pc = pc0; pc < pc1; ++pc) {
    const uint32_t w = (uint32_t)a.code[pc];
    const int op = w & 0xff, x = (w >> 8) & 0xff, y = (w >> 16) & 0xff;
    switch (op) {
      case FZ_LD_IN:
        if (y) S[sp++] = tos;
        tos = cp_load<T>(a, x, r, col);
        break;
      case FZ_LD_CONST:
        if (y) S[sp++] = tos;
        tos = (T)a.consts[x];
        break;
      case CP_LD_COL:
        if (y) S[sp++] = tos;
        tos = colval[x];
        break;
      case CP_ST_COL: colval[x] = tos; break;
      case FZ_UNARY:
        switch (x) {
#define X(OP) case OP: tos = U<OP, T>::f(tos); break;
          AG_FOR_UNARY(X)
#undef X
        }
        break;
      case FZ_BIN:
      case FZ_BIN_IN:
      case FZ_BIN_CONST: {
        T l, rr;
        if (op == FZ_BIN) { l = S[--sp]; rr = tos; }
        else {
          const T o = op == FZ_BIN_IN ? cp_load<T>(a, y & 0x7f, r, col) : (T)a.consts[y & 0x7f];
          l = (y & 0x80) ? o : tos;
          rr = (y & 0x80) ? tos : o;
        }
        const T ab[2] = {l, rr};
        switch (x) {
#define CP_BIN_CASE(ID, F) case ID: tos = F<T>::apply(ab); break;
          CP_BIN_CASE(AG_B_ADD, FAdd) CP_BIN_CASE(AG_B_SUB, FSub) CP_BIN_CASE(AG_B_MUL, FMul)
          CP_BIN_CASE(AG_B_DIV, FDiv) CP_BIN_CASE(AG_B_MAX, FMax) CP_BIN_CASE(AG_B_MIN, FMin)
          CP_BIN_CASE(AG_B_POW, FPow) CP_BIN_CASE(AG_B_EQ, FEq) CP_BIN_CASE(AG_B_LT, FLt)
          CP_BIN_CASE(AG_B_LE, FLe) CP_BIN_CASE(AG_B_GT, FGt) CP_BIN_CASE(AG_B_GE, FGe)
          CP_BIN_CASE(AG_B_ATAN2, FAtan2) CP_BIN_CASE(AG_B_HYPOT, FHypot)
#undef CP_BIN_CASE
        }
        break;
      }
      case FZ_UGRAD: {
        const T gx = S[--sp];
        const T gdy = S[--sp];
        switch (x) {
#define X(OP) case OP: tos = gdy * U<OP, T>::g(U<OP, T>::needs_x ? gx : T(0), U<OP, T>::needs_y ? tos : T(0)); break;
          AG_FOR_UNARY(X)
#undef X
        }
        break;
      }
      case FZ_TERN: {
        const T gp = S[--sp];
        const T gdy = S[--sp];
        const T t3[3] = {gdy, gp, tos};
        switch (x) {
#define CP_TERN_CASE(ID, G) case ID: tos = G<T>::apply(t3); break;
          CP_TERN_CASE(0, GDiv2) CP_TERN_CASE(1, GEqMask) CP_TERN_CASE(2, GPow1) CP_TERN_CASE(3, GPow2)
          CP_TERN_CASE(4, GAtan1) CP_TERN_CASE(5, GAtan2) CP_TERN_CASE(6, GHypot)
#undef CP_TERN_CASE
        }
        break;
      }
      case FZ_STORE:
        if (ELEM) ((T*)a.out[x])[r + a.m * col] = tos;
        else if (writer) ((T*)a.out[x])[col] = tos;
        break;
      case CP_CRED: colacc[x] += (double)tos; break;
      case CP_CFIN: {
        double v = colacc[x];
        colacc[x] = 0.0;
        // (the accumulators were combined over the G lanes before the segment: identical totals on every lane)
        tos = (T)v;
        colval[x] = tos;
        break;
      }
      case FZ_REDUCE:
        if (ELEM || writer) racc += (double)tos;
        break;
    }
  }
}

// Whole-space sum of the REDUCE operand: per-block partial in double, the last block to finish adds the partials in
// block order (the result does not depend on the schedule) -- no second launch.
template <typename T>
__device__ __forceinline__ void cp_finish(const CpArgs& a, double racc) {
  if (a.part) {
    __shared__ double sm[kThreads / 32];
    __shared__ int last;
    double v = racc;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      double r = 0;
#pragma unroll
      for (int w = 0; w < kThreads / 32; ++w) r += sm[w];
      a.part[blockIdx.x] = r;
      __threadfence();
      last = atomicAdd(a.ticket, 1) == (int)gridDim.x - 1;
    }
    __syncthreads();
    if (last) {
      // the last block to finish adds the partials in block order (the result does not depend on the schedule)
      double s = 0;
      for (int64_t i = threadIdx.x; i < gridDim.x; i += kThreads) s += __ldcg(a.part + i);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      __syncthreads();
      if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
      __syncthreads();
      if (threadIdx.x == 0) {
        double r = 0;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) r += sm[w];
        *(T*)a.rout = (T)r;
        *a.ticket = 0;
      }
    }
  }
}

template <typename T, int G>
__global__ void __launch_bounds__(kThreads) k_colprog(const __grid_constant__ CpArgs a) {
  const int lane = threadIdx.x % G;
  const int64_t col = ((int64_t)blockIdx.x * kThreads + threadIdx.x) / G;
  const bool on = col < a.n;
  T colval[CP_COLS];
  double colacc[CP_COLS];
#pragma unroll
  for (int k = 0; k < CP_COLS; ++k) { colval[k] = T(0); colacc[k] = 0.0; }
  double racc = 0.0;
  int pc = 0;
  while (pc < a.ncode) {
    const int elem = (((uint32_t)a.code[pc] >> 8) & 0xff) == 0;      // a CP_PHASE word
    int pe = pc + 1;
    while (pe < a.ncode && (a.code[pe] & 0xff) != CP_PHASE) ++pe;
    if (elem) {
      if (on)
        for (int64_t r = lane; r < a.m; r += G) cp_run<T, true>(a, pc + 1, pe, r, col, true, colval, colacc, racc);
    } else {
      // finish the column sums this segment reads: combine the accumulators over the G lanes first
      if (G > 1) {
        for (int q = pc + 1; q < pe; ++q)
          if ((a.code[q] & 0xff) == CP_CFIN) {
            const int x = (a.code[q] >> 8) & 0xff;
            double v = colacc[x];
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            colacc[x] = v;
          }
      }
      if (on) cp_run<T, false>(a, pc + 1, pe, 0, col, lane == 0, colval, colacc, racc);
    }
    pc = pe;
  }
  cp_finish<T>(a, racc);
}

// ---- specialised column programs: the program as a compile-time constant (fuse_static_col.inc) --------------
struct StaticColProg {
  int n;
  uint32_t code[FZ_MAXCODE];
};

#include "fuse_static_col.inc"

constexpr int cp_seg_end(const StaticColProg& p, int pc) {
  while (pc < p.n && (p.code[pc] & 0xff) != CP_PHASE) ++pc;
  return pc;
}
constexpr int cp_sp_before(const StaticColProg& p, int pc) {      // stack depth before instruction pc of its segment
  int sp = 0;
  for (int i = 0; i < pc; ++i) {
    const int op = p.code[i] & 0xff, y = (p.code[i] >> 16) & 0xff;
    if (op == CP_PHASE) sp = 0;
    if ((op == FZ_LD_IN || op == FZ_LD_CONST || op == CP_LD_COL) && y) ++sp;
    if (op == FZ_BIN) --sp;
    if (op == FZ_UGRAD || op == FZ_TERN) sp -= 2;
  }
  return sp;
}

template <int PROG, int PC, bool ELEM, typename T>
__device__ __forceinline__ void cps_step(const CpArgs& a, int64_t r, int64_t col, bool writer, T (&S)[FZ_SLOTS], T& tos,
                                         T (&colval)[CP_COLS], double (&colacc)[CP_COLS], double& racc) {
  constexpr uint32_t w = kColProgs[PROG].code[PC];
  constexpr int op = w & 0xff, x = (w >> 8) & 0xff, y = (w >> 16) & 0xff;
  constexpr int sp = cp_sp_before(kColProgs[PROG], PC);
  if constexpr (op == FZ_LD_IN || op == FZ_LD_CONST || op == CP_LD_COL) {
    if constexpr (y != 0) S[sp] = tos;
    if constexpr (op == FZ_LD_IN) tos = cp_load<T>(a, x, r, col);
    else if constexpr (op == FZ_LD_CONST) tos = (T)a.consts[x];
    else tos = colval[x];
  } else if constexpr (op == CP_ST_COL) {
    colval[x] = tos;
  } else if constexpr (op == FZ_UNARY) {
    tos = U<x, T>::f(tos);
  } else if constexpr (op == FZ_BIN) {
    const T ab[2] = {S[sp - 1], tos};
    tos = BinSel<x, T>::type::apply(ab);
  } else if constexpr (op == FZ_BIN_IN || op == FZ_BIN_CONST) {
    constexpr int k = y & 0x7f;
    constexpr bool rev = (y & 0x80) != 0;
    T o;
    if constexpr (op == FZ_BIN_IN) o = cp_load<T>(a, k, r, col);
    else o = (T)a.consts[k];
    const T ab[2] = {rev ? o : tos, rev ? tos : o};
    tos = BinSel<x, T>::type::apply(ab);
  } else if constexpr (op == FZ_UGRAD) {
    tos = S[sp - 2] * U<x, T>::g(U<x, T>::needs_x ? S[sp - 1] : T(0), U<x, T>::needs_y ? tos : T(0));
  } else if constexpr (op == FZ_TERN) {
    const T t3[3] = {S[sp - 2], S[sp - 1], tos};
    tos = TernSel<x, T>::type::apply(t3);
  } else if constexpr (op == FZ_STORE) {
    if constexpr (ELEM) ((T*)a.out[x])[r + a.m * col] = tos;
    else { if (writer) ((T*)a.out[x])[col] = tos; }
  } else if constexpr (op == CP_CRED) {
    colacc[x] += (double)tos;
  } else if constexpr (op == CP_CFIN) {
    tos = (T)colacc[x];
    colacc[x] = 0.0;
    colval[x] = tos;
  } else if constexpr (op == FZ_REDUCE) {
    if (ELEM || writer) racc += (double)tos;
  }
  if constexpr (op == FZ_UNARY || op == FZ_BIN || op == FZ_BIN_IN || op == FZ_BIN_CONST || op == FZ_UGRAD ||
                op == FZ_TERN)
    fz_opaque(tos, a.zero);            // every instruction's result is rounded on its own (no FMA contraction)
}

template <int PROG, int P0, bool ELEM, typename T, int... I>
__device__ __forceinline__ void cps_seg(const CpArgs& a, int64_t r, int64_t col, bool writer, T (&colval)[CP_COLS],
                                        double (&colacc)[CP_COLS], double& racc, std::integer_sequence<int, I...>) {
  T S[FZ_SLOTS], tos = T(0);
  (cps_step<PROG, P0 + I, ELEM, T>(a, r, col, writer, S, tos, colval, colacc, racc), ...);
}

// combine over the G lanes the accumulators that the column segment [P0, P0 + len) finishes
template <int PROG, int P0, int G, int... I>
__device__ __forceinline__ void cps_combine(double (&colacc)[CP_COLS], std::integer_sequence<int, I...>) {
  (([&] {
     constexpr uint32_t w = kColProgs[PROG].code[P0 + I];
     if constexpr ((w & 0xff) == CP_CFIN) {
       constexpr int x = (w >> 8) & 0xff;
       double v = colacc[x];
#pragma unroll
       for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
       colacc[x] = v;
     }
   }()),
   ...);
}

template <int PROG, int PC, typename T, int G>
__device__ __forceinline__ void cps_all(const CpArgs& a, int lane, int64_t col, bool on, T (&colval)[CP_COLS],
                                        double (&colacc)[CP_COLS], double& racc) {
  if constexpr (PC < kColProgs[PROG].n) {
    constexpr bool elem = ((kColProgs[PROG].code[PC] >> 8) & 0xff) == 0;
    constexpr int pe = cp_seg_end(kColProgs[PROG], PC + 1);
    constexpr int len = pe - PC - 1;
    if constexpr (elem) {
      if (on)
        for (int64_t r = lane; r < a.m; r += G)
          cps_seg<PROG, PC + 1, true, T>(a, r, col, true, colval, colacc, racc, std::make_integer_sequence<int, len>{});
    } else {
      if constexpr (G > 1) cps_combine<PROG, PC + 1, G>(colacc, std::make_integer_sequence<int, len>{});
      if (on)
        cps_seg<PROG, PC + 1, false, T>(a, 0, col, lane == 0, colval, colacc, racc, std::make_integer_sequence<int, len>{});
    }
    cps_all<PROG, pe, T, G>(a, lane, col, on, colval, colacc, racc);
  }
}

template <typename T, int G, int PROG>
__global__ void __launch_bounds__(kThreads) k_colprog_static(const __grid_constant__ CpArgs a) {
  const int lane = threadIdx.x % G;
  const int64_t col = ((int64_t)blockIdx.x * kThreads + threadIdx.x) / G;
  const bool on = col < a.n;
  T colval[CP_COLS];
  double colacc[CP_COLS];
#pragma unroll
  for (int k = 0; k < CP_COLS; ++k) { colval[k] = T(0); colacc[k] = 0.0; }
  double racc = 0.0;
  cps_all<PROG, 0, T, G>(a, lane, col, on, colval, colacc, racc);
  cp_finish<T>(a, racc);
}

template <int G>
static bool launch_col_static(int prog, const CpArgs& a, unsigned nb, cudaStream_t st) {
  switch (prog) {
#define X(P) case P: k_colprog_static<float, G, P><<<nb, kThreads, 0, st>>>(a); return true;
    FZ_COLSTATIC_LIST(X)
#undef X
  }
  return false;
}

static int find_col_static(const int32_t* code, int ncode) {
  for (int p = 0; p < FZ_NCOLSTATIC; ++p) {
    if (kColProgs[p].n != ncode) continue;
    bool same = true;
    for (int k = 0; k < ncode && same; ++k) same = kColProgs[p].code[k] == (uint32_t)code[k];
    if (same) return p;
  }
  return -1;
}

static int32_t* g_fz_tickets = nullptr;    // 1 + FZ_FIN_MAXGROUPS tickets, zero between launches (self-resetting)
static int g_fused_finish = 0;             // AGB_FUSED_FINISH=1: finish the partials in the kernel (measured slower)
static int32_t* g_cp_ticket = nullptr;
bool colprog_init() {
  if (!g_fz_tickets) {
    if (cudaMalloc((void**)&g_fz_tickets, (1 + FZ_FIN_MAXGROUPS) * sizeof(int32_t)) != cudaSuccess) {
      cudaGetLastError();
      set_error("out of device memory allocating the fused-sum tickets");
      return false;
    }
    cudaMemset(g_fz_tickets, 0, (1 + FZ_FIN_MAXGROUPS) * sizeof(int32_t));
    const char* e = getenv("AGB_FUSED_FINISH");
    g_fused_finish = (e && e[0] == '1') ? 1 : 0;
  }
  if (g_cp_ticket) return true;
  if (cudaMalloc((void**)&g_cp_ticket, sizeof(int32_t)) != cudaSuccess) {
    cudaGetLastError();
    set_error("out of device memory allocating the column-program ticket");
    return false;
  }
  cudaMemset(g_cp_ticket, 0, sizeof(int32_t));
  return true;
}

static int g_fused_static_enabled = -1;
static int64_t g_fused_hits = 0, g_fused_misses = 0;   // launches of specialised / interpreted programs

}  // namespace agb

using namespace agb;

extern "C" ag_status ag_ewise_fused_stats(int64_t* out2) {
  if (!out2) return AG_EINVAL;
  out2[0] = g_fused_hits;
  out2[1] = g_fused_misses;
  return AG_OK;
}

extern "C" ag_status ag_ewise_fused_query(const int32_t* code, int32_t ncode, int32_t* is_static) {
  if (!code || !is_static || ncode < 1) return AG_EINVAL;
  if (g_fused_static_enabled < 0) {
    const char* e = getenv("AGB_FUSED_STATIC");
    g_fused_static_enabled = (e && e[0] == '0') ? 0 : 1;
  }
  *is_static = (g_fused_static_enabled && find_static(code, ncode) >= 0) ? 1 : 0;
  return AG_OK;
}

extern "C" ag_status ag_ewise_fused_set_static(int32_t on) {
  g_fused_static_enabled = on ? 1 : 0;
  return AG_OK;
}

// r_out == nullptr: plain fused launch.  Otherwise the FZ_REDUCE operand (viewed as inner x red x outer, column-major)
// is summed over `red` into r_out (inner*outer values), in the same launch plus one small combine kernel.
static ag_status fused_impl(int32_t dtype, const int32_t* code, int32_t ncode, const double* consts, int32_t nconsts,
                            int32_t nin, const void* const* in, const int32_t* in_mode, const double* in_imm,
                            int32_t nout, void* const* out, int64_t rows, int64_t cols, int64_t r_inner, int64_t r_red,
                            int64_t r_outer, void* r_out) {
  AG_REQUIRE_INIT();
  AG_CHECK(dtype == AG_F32 || dtype == AG_F64, AG_EUNSUPPORTED, "ag_ewise_fused: dtype %d", dtype);
  AG_CHECK(code && ncode >= 1 && ncode <= FZ_MAXCODE, AG_EUNSUPPORTED, "ag_ewise_fused: program length %d (max %d)",
           ncode, FZ_MAXCODE);
  AG_CHECK(nconsts >= 0 && nconsts <= FZ_MAXCONST && nin >= 0 && nin <= FZ_MAXIN && nout >= (r_out ? 0 : 1) && nout <= FZ_MAXOUT,
           AG_EUNSUPPORTED, "ag_ewise_fused: %d constants / %d inputs / %d outputs (max %d / %d / %d)", nconsts, nin,
           nout, FZ_MAXCONST, FZ_MAXIN, FZ_MAXOUT);
  AG_CHECK(rows >= 0 && cols >= 0, AG_EINVAL, "ag_ewise_fused: negative extent");
  AG_CHECK((nin == 0 || (in && in_mode && in_imm)) && (out || nout == 0) && (nconsts == 0 || consts), AG_EINVAL,
           "ag_ewise_fused: null argument table");
  const int64_t n = rows * cols;
  if (n == 0) return AG_OK;
  FzArgs a;
  // validate the program: operand indices, slot bounds, stack discipline, every output written
  int sp = 0, tmp_lo = FZ_SLOTS, ns_need = 1, nreduce = 0;
  uint32_t stored = 0;
  for (int pc = 0; pc < ncode; ++pc) {
    const uint32_t w = (uint32_t)code[pc];
    const int op = w & 0xff, x = (w >> 8) & 0xff, y = (w >> 16) & 0xff;
    AG_CHECK(op < FZ_NOPS, AG_EINVAL, "ag_ewise_fused: bad opcode %d at %d", op, pc);
    switch (op) {
      case FZ_LD_IN: AG_CHECK(x < nin, AG_EINVAL, "ag_ewise_fused: input %d of %d", x, nin); break;
      case FZ_LD_CONST: AG_CHECK(x < nconsts, AG_EINVAL, "ag_ewise_fused: constant %d of %d", x, nconsts); break;
      case FZ_LD_TMP:
      case FZ_ST_TMP: AG_CHECK(x < FZ_SLOTS && x >= sp + (op == FZ_LD_TMP && y ? 1 : 0), AG_EINVAL,
                               "ag_ewise_fused: temporary slot %d collides with the stack (depth %d)", x, sp); break;
      case FZ_UNARY:
      case FZ_UGRAD: AG_CHECK(x < AG_U_COUNT, AG_EINVAL, "ag_ewise_fused: unary op %d", x); break;
      case FZ_BIN: AG_CHECK(x < AG_B_COUNT, AG_EINVAL, "ag_ewise_fused: binary op %d", x); break;
      case FZ_BIN_IN: AG_CHECK(x < AG_B_COUNT && (y & 0x7f) < nin, AG_EINVAL, "ag_ewise_fused: bad BIN_IN at %d", pc); break;
      case FZ_BIN_CONST: AG_CHECK(x < AG_B_COUNT && (y & 0x7f) < nconsts, AG_EINVAL, "ag_ewise_fused: bad BIN_CONST at %d", pc); break;
      case FZ_TERN: AG_CHECK(x < 7, AG_EINVAL, "ag_ewise_fused: ternary map %d", x); break;
      case FZ_STORE: AG_CHECK(x < nout, AG_EINVAL, "ag_ewise_fused: output %d of %d", x, nout); stored |= 1u << x; break;
      case FZ_REDUCE: ++nreduce; break;
    }
    if (op <= FZ_LD_TMP && y) ++sp;
    if (op == FZ_BIN) sp -= 1;
    if (op == FZ_UGRAD || op == FZ_TERN) sp -= 2;
    AG_CHECK(sp >= 0, AG_EINVAL, "ag_ewise_fused: stack underflow at %d", pc);
    if (op == FZ_ST_TMP && x < tmp_lo) tmp_lo = x;
    if (sp > ns_need) ns_need = sp;
    if ((op == FZ_ST_TMP || op == FZ_LD_TMP) && x + 1 > ns_need) ns_need = x + 1;
    AG_CHECK(sp <= tmp_lo, AG_EUNSUPPORTED, "ag_ewise_fused: expression too deep (%d slots)", FZ_SLOTS);
    a.code[pc] = code[pc];
  }
  AG_CHECK(stored == (1u << nout) - 1u, AG_EINVAL, "ag_ewise_fused: an output is never stored");
  AG_CHECK(nreduce == (r_out ? 1 : 0), AG_EINVAL, "ag_ewise_fused: %d REDUCE instructions", nreduce);
  for (int k = 0; k < nconsts; ++k) a.consts[k] = consts[k];
  bool vec = true;
  a.has_bcast = 0;
  for (int k = 0; k < nout; ++k) {
    AG_CHECK(out[k], AG_EINVAL, "ag_ewise_fused: null output %d", k);
    a.out[k] = out[k];
    vec = vec && aligned16(out[k]);
  }
  for (int k = 0; k < nin; ++k) {
    a.mode[k] = in_mode[k];
    a.in[k] = in[k];
    a.imm[k] = in_imm[k];
    AG_CHECK(a.mode[k] >= 0 && a.mode[k] <= FZ_M_ROWVEC && (a.mode[k] == FZ_M_IMM || a.in[k]), AG_EINVAL,
             "ag_ewise_fused: bad input %d", k);
    if (a.mode[k] == FZ_M_FULL) vec = vec && aligned16(a.in[k]);
    if (a.mode[k] == FZ_M_COLVEC || a.mode[k] == FZ_M_ROWVEC) a.has_bcast = 1;
  }
  a.ncode = ncode;
  a.nin = nin;
  a.nout = nout;
  a.n = n;
  a.rows = rows > 0 ? rows : 1;
  a.zero = 0;
  Context& c = ctx();
  const int N = dtype == AG_F32 ? 4 : 2;
  // A chain over fewer elements than one vector (the scalar tail of a loss: -(sum)/n) runs as ONE padded vector: every
  // array of the chain is 16-byte aligned here, device allocations are at least 256-byte granular, so the lanes beyond
  // n read and write padding of the arrays' own blocks.  That lets the specialised kernels take it (the interpreter
  // costs a scattered 0.25 MB switch in instruction fetch for one element).  Not for fused sums (the padding lanes
  // would be added) and not with row / column operands.
  if (vec && n < N && !r_out && !a.has_bcast) {
    a.n = N;
    a.rows = N;
  }
  vec = vec && (a.n % N == 0);
  if (g_fused_static_enabled < 0) {
    const char* e = getenv("AGB_FUSED_STATIC");
    g_fused_static_enabled = (e && e[0] == '0') ? 0 : 1;
  }
  if (vec && a.has_bcast && a.rows % N == 0) {
    for (int k = 0; k < nin; ++k) {
      if (a.mode[k] == FZ_M_COLVEC && aligned16(a.in[k])) a.mode[k] = FZ_M_COLVEC_V;
      else if (a.mode[k] == FZ_M_ROWVEC) a.mode[k] = FZ_M_ROWVEC_1;
    }
  }
  const int prog = (vec && g_fused_static_enabled) ? find_static(code, ncode) : -1;
  AG_CHECK(prog >= 0 || (nin <= FZ_DYN_MAXIN && nout <= FZ_DYN_MAXOUT), AG_EUNSUPPORTED,
           "ag_ewise_fused: %d inputs / %d outputs need a specialised kernel (the interpreter takes %d / %d)", nin, nout,
           FZ_DYN_MAXIN, FZ_DYN_MAXOUT);
  a.rmode = 0;
  a.chunk = kThreads * (prog >= 0 ? static_un(prog) : 2) * N;
  a.tps = 1;
  a.rlen = 1;
  a.rout = nullptr;
  if (r_out) {
    // the sum is fused only into specialised kernels, and only for layouts a block's chunk can cover
    AG_CHECK(r_inner >= 1 && r_red >= 1 && r_outer >= 1 && r_inner * r_red * r_outer == n, AG_EINVAL,
             "ag_ewise_fused_reduce: %lld x %lld x %lld is not %lld elements", (long long)r_inner, (long long)r_red,
             (long long)r_outer, (long long)n);
    AG_CHECK(prog >= 0, AG_EUNSUPPORTED, "ag_ewise_fused_reduce: no specialised kernel for this program");
    const int E = a.chunk;
    if (r_inner == 1 && r_outer == 1) {
      a.rmode = 1;
    } else if (r_inner == 1) {                       // contiguous runs of r_red
      AG_CHECK(r_red <= E, AG_EUNSUPPORTED, "ag_ewise_fused_reduce: runs of %lld exceed a block", (long long)r_red);
      int runs = (int)(E / r_red);
      while (runs > 1 && (runs * r_red) % N) --runs;
      AG_CHECK((runs * r_red) % N == 0 && runs <= kThreads, AG_EUNSUPPORTED, "ag_ewise_fused_reduce: run length %lld",
               (long long)r_red);
      a.rmode = 2;
      a.chunk = (int)(runs * r_red);
      a.rlen = r_red;
      int tps = 1;
      while (tps * 2 * runs <= kThreads && tps * 2 <= r_red) tps *= 2;
      a.tps = tps;
    } else {                                         // row sums: value is r_inner x r_red, r_outer == 1
      AG_CHECK(r_outer == 1 && r_inner * 16 <= E, AG_EUNSUPPORTED,
               "ag_ewise_fused_reduce: unsupported layout %lld x %lld x %lld", (long long)r_inner, (long long)r_red,
               (long long)r_outer);
      int colsb = (int)(E / r_inner);
      while (colsb > 1 && (colsb * r_inner) % N) --colsb;
      AG_CHECK((colsb * r_inner) % N == 0, AG_EUNSUPPORTED, "ag_ewise_fused_reduce: %lld rows", (long long)r_inner);
      a.rmode = 3;
      a.chunk = (int)(colsb * r_inner);
      a.rlen = r_inner;
    }
  }
  const int64_t nblk = cdiv(a.n, (int64_t)a.chunk);
  double* part = nullptr;
  a.tickets = nullptr;
  a.lvl2 = nullptr;
  a.rfinal = nullptr;
  bool finish_in_kernel = false;
  if (a.rmode == 1 || a.rmode == 3) {
    const int64_t ngroups = cdiv(nblk, (int64_t)FZ_FIN_GROUP);
    finish_in_kernel = g_fz_tickets && ngroups <= FZ_FIN_MAXGROUPS && g_fused_finish != 0;
    part = (double*)workspace(((size_t)(nblk + ngroups) * (size_t)a.rlen + 64) * sizeof(double));
    if (!part) return AG_ENOMEM;
    a.rout = part;
    if (finish_in_kernel) {
      a.tickets = g_fz_tickets;
      a.lvl2 = part + (size_t)nblk * (size_t)a.rlen;
      a.rfinal = r_out;
    }
  } else if (a.rmode == 2) {
    a.rout = r_out;
  }
  bool done = false;
  if (prog >= 0)
    done = dtype == AG_F32 ? launch_static<float>(prog, a, (unsigned)nblk, c.stream)
                           : launch_static<double>(prog, a, (unsigned)nblk, c.stream);
  if (done) ++g_fused_hits; else ++g_fused_misses;
  if (!done) {
    AG_CHECK(!r_out, AG_EUNSUPPORTED, "ag_ewise_fused_reduce: no specialised kernel for this program");
    const int64_t items = vec ? a.n / N : a.n;
    const unsigned nb = (unsigned)cdiv(items, kThreads);
    if (vec) {
      if (dtype == AG_F32) launch_dynamic<float, true>(nin, ns_need, a, nb, c.stream);
      else launch_dynamic<double, true>(nin, ns_need, a, nb, c.stream);
    } else {
      if (dtype == AG_F32) launch_dynamic<float, false>(nin, ns_need, a, nb, c.stream);
      else launch_dynamic<double, false>(nin, ns_need, a, nb, c.stream);
    }
  }
  AG_LAUNCHED();
  if (part && !finish_in_kernel) return reduce_double_partials(dtype, part, nblk, a.rmode == 1 ? 1 : a.rlen, r_out);
  return AG_OK;
}

extern "C" ag_status ag_ewise_fused(int32_t dtype, const int32_t* code, int32_t ncode, const double* consts,
                                    int32_t nconsts, int32_t nin, const void* const* in, const int32_t* in_mode,
                                    const double* in_imm, int32_t nout, void* const* out, int64_t rows,
                                    int64_t cols) {
  return fused_impl(dtype, code, ncode, consts, nconsts, nin, in, in_mode, in_imm, nout, out, rows, cols, 0, 0, 0,
                    nullptr);
}

extern "C" ag_status ag_ewise_fused_reduce(int32_t dtype, const int32_t* code, int32_t ncode, const double* consts,
                                           int32_t nconsts, int32_t nin, const void* const* in, const int32_t* in_mode,
                                           const double* in_imm, int32_t nout, void* const* out, int64_t rows,
                                           int64_t cols, int64_t r_inner, int64_t r_red, int64_t r_outer,
                                           void* r_out) {
  AG_CHECK(r_out, AG_EINVAL, "ag_ewise_fused_reduce: null result");
  return fused_impl(dtype, code, ncode, consts, nconsts, nin, in, in_mode, in_imm, nout, out, rows, cols, r_inner,
                    r_red, r_outer, r_out);
}

extern "C" ag_status ag_ewise_colprog(int32_t dtype, const int32_t* code, int32_t ncode, const double* consts,
                                      int32_t nconsts, int32_t nin, const void* const* in, const int32_t* in_mode,
                                      const double* in_imm, int32_t nout, void* const* out, int64_t m, int64_t n,
                                      void* r_out) {
  AG_REQUIRE_INIT();
  AG_CHECK(dtype == AG_F32 || dtype == AG_F64, AG_EUNSUPPORTED, "ag_ewise_colprog: dtype %d", dtype);
  AG_CHECK(code && ncode >= 2 && ncode <= FZ_MAXCODE, AG_EUNSUPPORTED, "ag_ewise_colprog: program length %d (max %d)",
           ncode, FZ_MAXCODE);
  AG_CHECK(nconsts >= 0 && nconsts <= FZ_MAXCONST && nin >= 0 && nin <= FZ_DYN_MAXIN && nout >= 0 && nout <= FZ_DYN_MAXOUT,
           AG_EUNSUPPORTED, "ag_ewise_colprog: %d constants / %d inputs / %d outputs (max %d / %d / %d)", nconsts, nin,
           nout, FZ_MAXCONST, FZ_DYN_MAXIN, FZ_DYN_MAXOUT);
  AG_CHECK(m >= 0 && n >= 0, AG_EINVAL, "ag_ewise_colprog: negative extent");
  AG_CHECK((nin == 0 || (in && in_mode)) && (out || nout == 0) && (nconsts == 0 || consts), AG_EINVAL,
           "ag_ewise_colprog: null argument table");
  (void)in_imm;
  if (m * n == 0) return AG_OK;
  AG_CHECK(g_cp_ticket, AG_ESTATE, "ag_ewise_colprog: not initialised");
  CpArgs a;
  // validate: segment structure, operand indices, stack discipline per segment, every output stored
  AG_CHECK((code[0] & 0xff) == CP_PHASE, AG_EINVAL, "ag_ewise_colprog: the program must start with a segment marker");
  int sp = 0, nreduce = 0, elem = 0;
  uint32_t stored = 0;
  for (int pc = 0; pc < ncode; ++pc) {
    const uint32_t w = (uint32_t)code[pc];
    const int op = w & 0xff, x = (w >> 8) & 0xff, y = (w >> 16) & 0xff;
    AG_CHECK(op < CP_NOPS && op != FZ_LD_TMP && op != FZ_ST_TMP, AG_EINVAL, "ag_ewise_colprog: bad opcode %d at %d", op, pc);
    switch (op) {
      case CP_PHASE: AG_CHECK(x <= 1, AG_EINVAL, "ag_ewise_colprog: bad segment marker"); elem = x == 0; sp = 0; break;
      case FZ_LD_IN:
        AG_CHECK(x < nin, AG_EINVAL, "ag_ewise_colprog: input %d of %d", x, nin);
        AG_CHECK(elem || (in_mode[x] != FZ_M_FULL && in_mode[x] != FZ_M_COLVEC), AG_EINVAL,
                 "ag_ewise_colprog: a column segment reads an element-space input");
        break;
      case FZ_LD_CONST: AG_CHECK(x < nconsts, AG_EINVAL, "ag_ewise_colprog: constant %d of %d", x, nconsts); break;
      case FZ_UNARY:
      case FZ_UGRAD: AG_CHECK(x < AG_U_COUNT, AG_EINVAL, "ag_ewise_colprog: unary op %d", x); break;
      case FZ_BIN: AG_CHECK(x < AG_B_COUNT, AG_EINVAL, "ag_ewise_colprog: binary op %d", x); break;
      case FZ_BIN_IN:
        AG_CHECK(x < AG_B_COUNT && (y & 0x7f) < nin, AG_EINVAL, "ag_ewise_colprog: bad BIN_IN at %d", pc);
        AG_CHECK(elem || (in_mode[y & 0x7f] != FZ_M_FULL && in_mode[y & 0x7f] != FZ_M_COLVEC), AG_EINVAL,
                 "ag_ewise_colprog: a column segment reads an element-space input");
        break;
      case FZ_BIN_CONST: AG_CHECK(x < AG_B_COUNT && (y & 0x7f) < nconsts, AG_EINVAL, "ag_ewise_colprog: bad BIN_CONST at %d", pc); break;
      case FZ_TERN: AG_CHECK(x < 7, AG_EINVAL, "ag_ewise_colprog: ternary map %d", x); break;
      case FZ_STORE:
        AG_CHECK(x < nout && !(stored & (1u << x)), AG_EINVAL, "ag_ewise_colprog: output %d of %d (or stored twice)", x, nout);
        stored |= 1u << x;
        break;
      case FZ_REDUCE: ++nreduce; break;
      case CP_CRED: AG_CHECK(elem && x < CP_COLS, AG_EINVAL, "ag_ewise_colprog: bad CRED at %d", pc); break;
      case CP_CFIN:
      case CP_ST_COL: AG_CHECK(!elem && x < CP_COLS, AG_EINVAL, "ag_ewise_colprog: bad column store at %d", pc); break;
      case CP_LD_COL: AG_CHECK(x < CP_COLS, AG_EINVAL, "ag_ewise_colprog: column value %d", x); break;
    }
    if ((op == FZ_LD_IN || op == FZ_LD_CONST || op == CP_LD_COL) && y) ++sp;
    if (op == FZ_BIN) sp -= 1;
    if (op == FZ_UGRAD || op == FZ_TERN) sp -= 2;
    AG_CHECK(sp >= 0 && sp <= FZ_SLOTS, AG_EINVAL, "ag_ewise_colprog: stack depth %d at %d", sp, pc);
    a.code[pc] = code[pc];
  }
  AG_CHECK(stored == (1u << nout) - 1u, AG_EINVAL, "ag_ewise_colprog: an output is never stored");
  AG_CHECK(nreduce == (r_out ? 1 : 0), AG_EINVAL, "ag_ewise_colprog: %d REDUCE instructions", nreduce);
  for (int k = 0; k < nconsts; ++k) a.consts[k] = consts[k];
  for (int k = 0; k < nout; ++k) {
    AG_CHECK(out[k], AG_EINVAL, "ag_ewise_colprog: null output %d", k);
    a.out[k] = out[k];
  }
  for (int k = 0; k < nin; ++k) {
    AG_CHECK(in[k] && (in_mode[k] == FZ_M_FULL || in_mode[k] == FZ_M_DEVSCALAR || in_mode[k] == FZ_M_COLVEC ||
                       in_mode[k] == FZ_M_ROWVEC), AG_EINVAL, "ag_ewise_colprog: bad input %d", k);
    a.in[k] = in[k];
    a.mode[k] = in_mode[k];
  }
  a.ncode = ncode;
  a.nin = nin;
  a.nout = nout;
  a.m = m;
  a.n = n;
  Context& c = ctx();
  // G lanes share a column, rows along the lanes: enough lanes for a short column to take one row each (a thread
  // that walks a whole column serialises one memory latency per row and segment: measured 2x slower at m = 10)
  const int G = m > 16 ? 32 : (m > 8 ? 16 : 8);
  const int64_t nblk = cdiv(n * G, (int64_t)kThreads);
  a.part = nullptr;
  a.ticket = g_cp_ticket;
  a.rout = r_out;
  if (r_out) {
    a.part = (double*)workspace((size_t)nblk * sizeof(double) + 64);
    if (!a.part) return AG_ENOMEM;
  }
  a.zero = 0;
  if (g_fused_static_enabled < 0) {
    const char* e = getenv("AGB_FUSED_STATIC");
    g_fused_static_enabled = (e && e[0] == '0') ? 0 : 1;
  }
  const int prog = (dtype == AG_F32 && g_fused_static_enabled) ? find_col_static(code, ncode) : -1;
  bool done = false;
  if (prog >= 0)
    done = G == 8 ? launch_col_static<8>(prog, a, (unsigned)nblk, c.stream)
                  : (G == 16 ? launch_col_static<16>(prog, a, (unsigned)nblk, c.stream)
                             : launch_col_static<32>(prog, a, (unsigned)nblk, c.stream));
  if (done) ++g_fused_hits; else ++g_fused_misses;
  if (done) {
  } else if (dtype == AG_F32) {
    if (G == 8) k_colprog<float, 8><<<(unsigned)nblk, kThreads, 0, c.stream>>>(a);
    else if (G == 16) k_colprog<float, 16><<<(unsigned)nblk, kThreads, 0, c.stream>>>(a);
    else k_colprog<float, 32><<<(unsigned)nblk, kThreads, 0, c.stream>>>(a);
  } else {
    if (G == 8) k_colprog<double, 8><<<(unsigned)nblk, kThreads, 0, c.stream>>>(a);
    else if (G == 16) k_colprog<double, 16><<<(unsigned)nblk, kThreads, 0, c.stream>>>(a);
    else k_colprog<double, 32><<<(unsigned)nblk, kThreads, 0, c.stream>>>(a);
  }
  AG_LAUNCHED();
  return AG_OK;
}
