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

namespace agb {

constexpr int FZ_MAXCODE = 96, FZ_MAXCONST = 16, FZ_MAXIN = 8, FZ_MAXOUT = 4, FZ_SLOTS = 8;

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
  FZ_NOPS = 11
};
enum { FZ_M_FULL = 0, FZ_M_DEVSCALAR = 1, FZ_M_IMM = 2, FZ_M_COLVEC = 3, FZ_M_ROWVEC = 4 };

struct FzArgs {
  int32_t code[FZ_MAXCODE];
  double consts[FZ_MAXCONST];
  const void* in[FZ_MAXIN];
  double imm[FZ_MAXIN];
  int32_t mode[FZ_MAXIN];
  void* out[FZ_MAXOUT];
  int32_t ncode, nin, nout, has_bcast;
  int64_t n, rows;
};

#define FZ_CASE8(X) X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7)

template <typename T, bool VEC>
__global__ void __launch_bounds__(kThreads) k_fused(const __grid_constant__ FzArgs a) {
  constexpr int N = VEC ? Pack<T>::N : 1;
  const int64_t i = ((int64_t)blockIdx.x * kThreads + threadIdx.x) * N;
  if (i >= a.n) return;

  // (i0, i1) of each lane, only when some operand is a row / column vector
  int64_t c0[N], c1[N];
  if (a.has_bcast) {
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

  // all leaf loads are issued before any arithmetic
  T inv[FZ_MAXIN][N];
#pragma unroll
  for (int k = 0; k < FZ_MAXIN; ++k) {
    if (k < a.nin) {
      const T* __restrict__ p = (const T*)a.in[k];
      const int m = a.mode[k];
      if (m == FZ_M_FULL) {
        if (VEC) {
          const Pack<T> t = *reinterpret_cast<const Pack<T>*>(p + i);
#pragma unroll
          for (int j = 0; j < N; ++j) inv[k][j] = t.v[j];
        } else {
          inv[k][0] = p[i];
        }
      } else if (m == FZ_M_COLVEC) {
#pragma unroll
        for (int j = 0; j < N; ++j) inv[k][j] = p[c0[j]];
      } else if (m == FZ_M_ROWVEC) {
#pragma unroll
        for (int j = 0; j < N; ++j) inv[k][j] = p[c1[j]];
      } else {
        const T sc = m == FZ_M_DEVSCALAR ? p[0] : (T)a.imm[k];
#pragma unroll
        for (int j = 0; j < N; ++j) inv[k][j] = sc;
      }
    }
  }

  T tos[N], S[FZ_SLOTS][N];
#pragma unroll
  for (int j = 0; j < N; ++j) tos[j] = T(0);
  int sp = 0;

#define FZ_SLOT_ST_CASE(s) case s: _Pragma("unroll") for (int j = 0; j < N; ++j) S[s][j] = tos[j]; break;
#define FZ_SLOT_ST(idx) switch (idx) { FZ_CASE8(FZ_SLOT_ST_CASE) }
#define FZ_SLOT_LD_CASE(s) case s: _Pragma("unroll") for (int j = 0; j < N; ++j) dst_[j] = S[s][j]; break;
#define FZ_SLOT_LD(dst, idx) { T* dst_ = dst; switch (idx) { FZ_CASE8(FZ_SLOT_LD_CASE) } }
#define FZ_IN_LD_CASE(s) case s: _Pragma("unroll") for (int j = 0; j < N; ++j) dst_[j] = inv[s][j]; break;
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
        T* __restrict__ o = nullptr;
        switch (x) {
          case 0: o = (T*)a.out[0]; break;
          case 1: o = (T*)a.out[1]; break;
          case 2: o = (T*)a.out[2]; break;
          default: o = (T*)a.out[3]; break;
        }
        if (VEC) {
          Pack<T> t;
#pragma unroll
          for (int j = 0; j < N; ++j) t.v[j] = tos[j];
          *reinterpret_cast<Pack<T>*>(o + i) = t;
        } else {
          o[i] = tos[0];
        }
        break;
      }
    }
  }
}

}  // namespace agb

using namespace agb;

extern "C" ag_status ag_ewise_fused(int32_t dtype, const int32_t* code, int32_t ncode, const double* consts,
                                    int32_t nconsts, int32_t nin, const void* const* in, const int32_t* in_mode,
                                    const double* in_imm, int32_t nout, void* const* out, int64_t rows,
                                    int64_t cols) {
  AG_REQUIRE_INIT();
  AG_CHECK(dtype == AG_F32 || dtype == AG_F64, AG_EUNSUPPORTED, "ag_ewise_fused: dtype %d", dtype);
  AG_CHECK(code && ncode >= 1 && ncode <= FZ_MAXCODE, AG_EUNSUPPORTED, "ag_ewise_fused: program length %d (max %d)",
           ncode, FZ_MAXCODE);
  AG_CHECK(nconsts >= 0 && nconsts <= FZ_MAXCONST && nin >= 0 && nin <= FZ_MAXIN && nout >= 1 && nout <= FZ_MAXOUT,
           AG_EUNSUPPORTED, "ag_ewise_fused: %d constants / %d inputs / %d outputs (max %d / %d / %d)", nconsts, nin,
           nout, FZ_MAXCONST, FZ_MAXIN, FZ_MAXOUT);
  AG_CHECK(rows >= 0 && cols >= 0, AG_EINVAL, "ag_ewise_fused: negative extent");
  AG_CHECK((nin == 0 || (in && in_mode && in_imm)) && out && (nconsts == 0 || consts), AG_EINVAL,
           "ag_ewise_fused: null argument table");
  const int64_t n = rows * cols;
  if (n == 0) return AG_OK;
  FzArgs a;
  // validate the program: operand indices, slot bounds, stack discipline, every output written
  int sp = 0, tmp_lo = FZ_SLOTS;
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
    }
    if (op <= FZ_LD_TMP && y) ++sp;
    if (op == FZ_BIN) sp -= 1;
    if (op == FZ_UGRAD || op == FZ_TERN) sp -= 2;
    AG_CHECK(sp >= 0, AG_EINVAL, "ag_ewise_fused: stack underflow at %d", pc);
    if (op == FZ_ST_TMP && x < tmp_lo) tmp_lo = x;
    AG_CHECK(sp <= tmp_lo, AG_EUNSUPPORTED, "ag_ewise_fused: expression too deep (%d slots)", FZ_SLOTS);
    a.code[pc] = code[pc];
  }
  AG_CHECK(stored == (1u << nout) - 1u, AG_EINVAL, "ag_ewise_fused: an output is never stored");
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
  Context& c = ctx();
  const int N = dtype == AG_F32 ? 4 : 2;
  vec = vec && (n % N == 0);
  const int64_t items = vec ? n / N : n;
  const unsigned nb = (unsigned)cdiv(items, kThreads);
  if (vec) {
    if (dtype == AG_F32) k_fused<float, true><<<nb, kThreads, 0, c.stream>>>(a);
    else k_fused<double, true><<<nb, kThreads, 0, c.stream>>>(a);
  } else {
    if (dtype == AG_F32) k_fused<float, false><<<nb, kThreads, 0, c.stream>>>(a);
    else k_fused<double, false><<<nb, kThreads, 0, c.stream>>>(a);
  }
  AG_LAUNCHED();
  return AG_OK;
}
