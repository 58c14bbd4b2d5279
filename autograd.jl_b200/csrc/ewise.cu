// ewise.cu -- elementwise forward kernels and the fused gradient kernels `dx = dy .* g(x,y)`.
//
// Reference: every rule of src/math.jl:9-74 and src/base.jl:20-49,73-74 is a Julia broadcast
// expression evaluated by Base on the CPU, one pass per dotted call (src/core.jl:68-70 stops
// fusion).  Here each rule is ONE kernel: 128-bit vector loads/stores, 4 vectors in flight per
// thread, exact grids (HBM-bound: algorithmic bytes 3*N*s for a unary gradient).
#include "common.cuh"
#include "ewise_ops.cuh"

namespace agb {


template <typename T>
__device__ __forceinline__ Pack<T> load_pack(const T* __restrict__ p, int64_t i, bool vec) {
  Pack<T> r;
  if (vec) {
    r = *reinterpret_cast<const Pack<T>*>(p + i);
  } else {
#pragma unroll
    for (int j = 0; j < Pack<T>::N; ++j) r.v[j] = p[i + j];
  }
  return r;
}
template <typename T>
__device__ __forceinline__ void store_pack(T* __restrict__ p, int64_t i, const Pack<T>& r, bool vec) {
  if (vec) {
    *reinterpret_cast<Pack<T>*>(p + i) = r;
  } else {
#pragma unroll
    for (int j = 0; j < Pack<T>::N; ++j) p[i + j] = r.v[j];
  }
}

template <int OP, typename T>
__global__ void __launch_bounds__(kThreads) k_unary_fwd(const T* __restrict__ x, T* __restrict__ y,
                                                        int64_t n, bool vec) {
  constexpr int N = Pack<T>::N;
  const int64_t base = (int64_t)blockIdx.x * (kThreads * kUnroll * N);
  Pack<T> in[kUnroll];
#pragma unroll
  for (int u = 0; u < kUnroll; ++u) {
    int64_t i = base + ((int64_t)u * kThreads + threadIdx.x) * N;
    if (i + N <= n) in[u] = load_pack(x, i, vec);
  }
#pragma unroll
  for (int u = 0; u < kUnroll; ++u) {
    int64_t i = base + ((int64_t)u * kThreads + threadIdx.x) * N;
    if (i + N <= n) {
      Pack<T> o;
#pragma unroll
      for (int j = 0; j < N; ++j) o.v[j] = U<OP, T>::f(in[u].v[j]);
      store_pack(y, i, o, vec);
    } else {
      for (; i < n; ++i) y[i] = U<OP, T>::f(x[i]);
    }
  }
}

// dx = dy .* g(x, y);  DYARR: dy is a full array, else a scalar (device scalar or immediate)
template <int OP, typename T, bool DYARR>
__global__ void __launch_bounds__(kThreads)
    k_unary_bwd(const T* __restrict__ dy, T dy_imm, bool dy_dev_scalar, const T* __restrict__ x,
                const T* __restrict__ y, T* __restrict__ dx, int64_t n, bool vec) {
  constexpr int N = Pack<T>::N;
  constexpr bool NX = U<OP, T>::needs_x, NY = U<OP, T>::needs_y;
  const int64_t base = (int64_t)blockIdx.x * (kThreads * kUnroll * N);
  T dys = dy_imm;
  if (!DYARR && dy_dev_scalar) dys = dy[0];
  Pack<T> vdy[kUnroll], vx[kUnroll], vy[kUnroll];
#pragma unroll
  for (int u = 0; u < kUnroll; ++u) {
    int64_t i = base + ((int64_t)u * kThreads + threadIdx.x) * N;
    if (i + N <= n) {
      if (DYARR) vdy[u] = load_pack(dy, i, vec);
      if (NX) vx[u] = load_pack(x, i, vec);
      if (NY) vy[u] = load_pack(y, i, vec);
    }
  }
#pragma unroll
  for (int u = 0; u < kUnroll; ++u) {
    int64_t i = base + ((int64_t)u * kThreads + threadIdx.x) * N;
    if (i + N <= n) {
      Pack<T> o;
#pragma unroll
      for (int j = 0; j < N; ++j) {
        T d = DYARR ? vdy[u].v[j] : dys;
        o.v[j] = d * U<OP, T>::g(NX ? vx[u].v[j] : T(0), NY ? vy[u].v[j] : T(0));
      }
      store_pack(dx, i, o, vec);
    } else {
      for (; i < n; ++i) {
        T d = DYARR ? dy[i] : dys;
        dx[i] = d * U<OP, T>::g(NX ? x[i] : T(0), NY ? y[i] : T(0));
      }
    }
  }
}


template <int OP, typename T>
static ag_status launch_unary_fwd(const void* x, void* y, int64_t n) {
  Context& c = ctx();
  constexpr int N = Pack<T>::N;
  int64_t blocks = cdiv(n, (int64_t)kThreads * kUnroll * N);
  bool vec = aligned16(x) && aligned16(y);
  k_unary_fwd<OP, T><<<(unsigned)blocks, kThreads, 0, c.stream>>>((const T*)x, (T*)y, n, vec);
  AG_LAUNCHED();
  return AG_OK;
}

template <int OP, typename T>
static ag_status launch_unary_bwd(const void* dy, int dy_mode, double dy_scalar, const void* x,
                                  const void* y, void* dx, int64_t n) {
  Context& c = ctx();
  constexpr int N = Pack<T>::N;
  if (U<OP, T>::needs_x && !x) {
    set_error("ag_unary_bwd: op %d needs x", OP);
    return AG_EINVAL;
  }
  if (U<OP, T>::needs_y && !y) {
    set_error("ag_unary_bwd: op %d needs y", OP);
    return AG_EINVAL;
  }
  int64_t blocks = cdiv(n, (int64_t)kThreads * kUnroll * N);
  bool vec = aligned16(dx) && (!U<OP, T>::needs_x || aligned16(x)) &&
             (!U<OP, T>::needs_y || aligned16(y)) && (dy_mode != 0 || aligned16(dy));
  if (dy_mode == 0)
    k_unary_bwd<OP, T, true><<<(unsigned)blocks, kThreads, 0, c.stream>>>(
        (const T*)dy, T(0), false, (const T*)x, (const T*)y, (T*)dx, n, vec);
  else
    k_unary_bwd<OP, T, false><<<(unsigned)blocks, kThreads, 0, c.stream>>>(
        (const T*)dy, (T)dy_scalar, dy_mode == 1, (const T*)x, (const T*)y, (T*)dx, n, vec);
  AG_LAUNCHED();
  return AG_OK;
}


// ---- n-ary broadcast map -------------------------------------------------------------------------
// out[i0,i1,i2,i3] = F(in0[...], in1[...], in2[...]) over a (collapsed) <=4-d column-major shape;
// each operand has per-dimension element strides (0 = broadcast) or is an immediate.
struct Opnd {
  const void* p;
  double imm;
  int64_t s[4];
  int mode;  // 0 immediate, 1 unit stride along dim0 (vectorisable), 2 broadcast along dim0, 3 strided
};

template <int NIN>
struct MapArgs {
  Opnd in[NIN];
  void* out;
  int64_t shape[4];
  int64_t n;
  int nd;
  bool big;  // n >= 2^31: 64-bit index arithmetic
};

// ND = 2: at most two collapsed dims (one division per item, none when shape[1] == 1); ND = 4: general.
// All loads of the kUnroll items of a thread are issued before any arithmetic (memory-level parallelism).
template <typename T, int NIN, typename F, bool VEC, int ND>
__global__ void __launch_bounds__(kThreads) k_map(const MapArgs<NIN> a) {
  constexpr int N = VEC ? Pack<T>::N : 1;
  const int64_t base = ((int64_t)blockIdx.x * kUnroll * kThreads + threadIdx.x) * N;
  T* __restrict__ out = (T*)a.out;
  T v[kUnroll][NIN][N];
  bool ok[kUnroll];
#pragma unroll
  for (int u = 0; u < kUnroll; ++u) {
    const int64_t i = base + (int64_t)u * kThreads * N;
    ok[u] = i < a.n;
    if (!ok[u]) continue;
    int64_t i0 = i, i1 = 0, i2 = 0, i3 = 0;
    if (ND == 2) {
      if (a.shape[1] != 1) {
        if (a.big) { i1 = i / a.shape[0]; i0 = i - i1 * a.shape[0]; }
        else { uint32_t q = (uint32_t)i / (uint32_t)a.shape[0]; i1 = q; i0 = (uint32_t)i - q * (uint32_t)a.shape[0]; }
      }
    } else {
      int64_t r = i;
      int64_t q = r / a.shape[0]; i0 = r - q * a.shape[0]; r = q;
      q = r / a.shape[1]; i1 = r - q * a.shape[1]; r = q;
      q = r / a.shape[2]; i2 = r - q * a.shape[2]; i3 = q;
    }
#pragma unroll
    for (int k = 0; k < NIN; ++k) {
      const Opnd& o = a.in[k];
      if (o.mode == 0) {
#pragma unroll
        for (int j = 0; j < N; ++j) v[u][k][j] = (T)o.imm;
      } else {
        const T* __restrict__ p = (const T*)o.p;
        int64_t off = i0 * o.s[0] + i1 * o.s[1];
        if (ND > 2) off += i2 * o.s[2] + i3 * o.s[3];
        if (VEC && o.mode == 1) {
          const Pack<T> t = *reinterpret_cast<const Pack<T>*>(p + off);
#pragma unroll
          for (int j = 0; j < N; ++j) v[u][k][j] = t.v[j];
        } else if (o.mode == 2) {
          const T t = p[off];
#pragma unroll
          for (int j = 0; j < N; ++j) v[u][k][j] = t;
        } else {
#pragma unroll
          for (int j = 0; j < N; ++j) v[u][k][j] = p[off + j * o.s[0]];
        }
      }
    }
  }
#pragma unroll
  for (int u = 0; u < kUnroll; ++u) {
    if (!ok[u]) continue;
    const int64_t i = base + (int64_t)u * kThreads * N;
    if (VEC) {
      Pack<T> r;
#pragma unroll
      for (int j = 0; j < N; ++j) {
        T in[NIN];
#pragma unroll
        for (int k = 0; k < NIN; ++k) in[k] = v[u][k][j];
        r.v[j] = F::apply(in);
      }
      *reinterpret_cast<Pack<T>*>(out + i) = r;
    } else {
      T in[NIN];
#pragma unroll
      for (int k = 0; k < NIN; ++k) in[k] = v[u][k][0];
      out[i] = F::apply(in);
    }
  }
}


template <typename T, int NIN, typename F>
static ag_status launch_map(MapArgs<NIN>& a) {
  Context& c = ctx();
  constexpr int N = Pack<T>::N;
  // vectorise along dim0 when every N-run stays inside dim0 and every operand is either
  // unit-stride + aligned, broadcast, or immediate
  bool vec = (a.shape[0] % N == 0) && aligned16(a.out);
  for (int k = 0; k < NIN; ++k) {
    Opnd& o = a.in[k];
    if (!o.p) {
      o.mode = 0;
      continue;
    }
    if (o.s[0] == 0) {
      o.mode = 2;
    } else if (o.s[0] == 1) {
      o.mode = 1;
      bool ok = aligned16(o.p);
      for (int d = 1; d < a.nd; ++d) ok = ok && (o.s[d] % N == 0);
      if (!ok) o.mode = 3;
    } else {
      o.mode = 3;
    }
  }
  a.big = a.n >= ((int64_t)1 << 31);
  for (int d = a.nd; d < 4; ++d) a.shape[d] = 1;
  const bool nd2 = a.nd <= 2;
  if (vec) {
    int64_t blocks = cdiv(a.n, (int64_t)kThreads * kUnroll * N);
    if (nd2) k_map<T, NIN, F, true, 2><<<(unsigned)blocks, kThreads, 0, c.stream>>>(a);
    else k_map<T, NIN, F, true, 4><<<(unsigned)blocks, kThreads, 0, c.stream>>>(a);
  } else {
    for (int k = 0; k < NIN; ++k)
      if (a.in[k].mode == 1) a.in[k].mode = 3;
    int64_t blocks = cdiv(a.n, (int64_t)kThreads * kUnroll);
    if (nd2) k_map<T, NIN, F, false, 2><<<(unsigned)blocks, kThreads, 0, c.stream>>>(a);
    else k_map<T, NIN, F, false, 4><<<(unsigned)blocks, kThreads, 0, c.stream>>>(a);
  }
  AG_LAUNCHED();
  return AG_OK;
}

// ---- bytecode evaluator: fused elementwise kernels for user-defined primitives ------------------------
// `@primitive f(x),dy,y (expr)` accepts arbitrary broadcast expressions (src/macros.jl:88-105; e.g. sigm at
// examples/charlm.jl:46-47).  The host compiles such an expression to postfix bytecode once; this kernel
// evaluates it per element in ONE pass over memory (no temporaries).  Elementwise kernels have ~100 spare FP32
// instructions per element at HBM speed, so interpretation (decoded once per 128-bit vector) stays near the
// memory roofline for short programs.
constexpr int VM_MAXCODE = 64, VM_MAXCONST = 16, VM_MAXIN = 4, VM_STACK = 8;
enum { VM_IN = 0, VM_CONST = 1, VM_ADD = 2, VM_SUB = 3, VM_MUL = 4, VM_DIV = 5, VM_NEG = 6, VM_MAX = 7, VM_MIN = 8,
       VM_POW = 9, VM_EQ = 10, VM_LT = 11, VM_LE = 12, VM_GT = 13, VM_GE = 14, VM_UNARY = 15 };
struct VmArgs {
  int32_t code[VM_MAXCODE];
  double consts[VM_MAXCONST];
  const void* in[VM_MAXIN];
  double imm[VM_MAXIN];
  int32_t mode[VM_MAXIN];   // 0 array, 1 device scalar, 2 immediate
  int32_t ncode, nin;
  void* out;
  int64_t n;
};

template <typename T>
__device__ __forceinline__ T vm_unary(int u, T x) {
  switch (u) {
#define X(OP) case OP: return U<OP, T>::f(x);
    AG_FOR_UNARY(X)
#undef X
  }
  return x;
}

template <typename T, bool VEC>
__global__ void __launch_bounds__(kThreads) k_vm(const VmArgs a) {
  constexpr int N = VEC ? Pack<T>::N : 1;
  const int64_t i = ((int64_t)blockIdx.x * kThreads + threadIdx.x) * N;
  if (i >= a.n) return;
  T inv[VM_MAXIN][N];
#pragma unroll
  for (int k = 0; k < VM_MAXIN; ++k) {
    if (k < a.nin) {
      if (a.mode[k] == 0) {
        if (VEC) {
          const Pack<T> t = *reinterpret_cast<const Pack<T>*>((const T*)a.in[k] + i);
#pragma unroll
          for (int j = 0; j < N; ++j) inv[k][j] = t.v[j];
        } else {
          inv[k][0] = ((const T*)a.in[k])[i];
        }
      } else {
        const T sc = a.mode[k] == 1 ? ((const T*)a.in[k])[0] : (T)a.imm[k];
#pragma unroll
        for (int j = 0; j < N; ++j) inv[k][j] = sc;
      }
    }
  }
  T st[VM_STACK][N];
  int sp = 0;
  for (int pc = 0; pc < a.ncode; ++pc) {
    const int op = a.code[pc] & 0xff, arg = a.code[pc] >> 8;
    if (op == VM_IN) {
#pragma unroll
      for (int j = 0; j < N; ++j) st[sp][j] = arg == 0 ? inv[0][j] : (arg == 1 ? inv[1][j] : (arg == 2 ? inv[2][j] : inv[3][j]));
      ++sp;
    } else if (op == VM_CONST) {
      const T cv = (T)a.consts[arg];
#pragma unroll
      for (int j = 0; j < N; ++j) st[sp][j] = cv;
      ++sp;
    } else if (op >= VM_UNARY || op == VM_NEG) {
#pragma unroll
      for (int j = 0; j < N; ++j) st[sp - 1][j] = op == VM_NEG ? -st[sp - 1][j] : vm_unary<T>(op - VM_UNARY, st[sp - 1][j]);
    } else {
      --sp;
#pragma unroll
      for (int j = 0; j < N; ++j) {
        const T x = st[sp - 1][j], y = st[sp][j];
        T r;
        switch (op) {
          case VM_ADD: r = x + y; break;
          case VM_SUB: r = x - y; break;
          case VM_MUL: r = x * y; break;
          case VM_DIV: r = x / y; break;
          case VM_MAX: r = jmax(x, y); break;
          case VM_MIN: r = jmin(x, y); break;
          case VM_POW: r = pow(x, y); break;
          case VM_EQ: r = x == y ? T(1) : T(0); break;
          case VM_LT: r = x < y ? T(1) : T(0); break;
          case VM_LE: r = x <= y ? T(1) : T(0); break;
          case VM_GT: r = x > y ? T(1) : T(0); break;
          default: r = x >= y ? T(1) : T(0); break;
        }
        st[sp - 1][j] = r;
      }
    }
  }
  if (VEC) {
    Pack<T> o;
#pragma unroll
    for (int j = 0; j < N; ++j) o.v[j] = st[0][j];
    *reinterpret_cast<Pack<T>*>((T*)a.out + i) = o;
  } else {
    ((T*)a.out)[i] = st[0][0];
  }
}

static ag_status fill_shape(int32_t ndims, const int64_t* shape, int64_t* out_shape, int64_t* n) {
  AG_CHECK(ndims >= 1 && ndims <= 4 && shape, AG_EUNSUPPORTED,
           "broadcast map: ndims=%d (collapse to <=4 dims on the host)", ndims);
  int64_t total = 1;
  for (int d = 0; d < 4; ++d) {
    out_shape[d] = d < ndims ? shape[d] : 1;
    AG_CHECK(out_shape[d] >= 0, AG_EINVAL, "broadcast map: negative extent");
    total *= out_shape[d];
  }
  *n = total;
  return AG_OK;
}

static void set_opnd(Opnd& o, const void* p, double imm, const int64_t* strides, int nd) {
  o.p = p;
  o.imm = imm;
  for (int d = 0; d < 4; ++d) o.s[d] = (p && strides && d < nd) ? strides[d] : 0;
  o.mode = 0;
}

template <typename T>
static ag_status binary_dispatch(int32_t op, MapArgs<2>& a) {
  switch (op) {
    case AG_B_ADD: return launch_map<T, 2, FAdd<T>>(a);
    case AG_B_SUB: return launch_map<T, 2, FSub<T>>(a);
    case AG_B_MUL: return launch_map<T, 2, FMul<T>>(a);
    case AG_B_DIV: return launch_map<T, 2, FDiv<T>>(a);
    case AG_B_MAX: return launch_map<T, 2, FMax<T>>(a);
    case AG_B_MIN: return launch_map<T, 2, FMin<T>>(a);
    case AG_B_POW: return launch_map<T, 2, FPow<T>>(a);
    case AG_B_EQ: return launch_map<T, 2, FEq<T>>(a);
    case AG_B_LT: return launch_map<T, 2, FLt<T>>(a);
    case AG_B_LE: return launch_map<T, 2, FLe<T>>(a);
    case AG_B_GT: return launch_map<T, 2, FGt<T>>(a);
    case AG_B_GE: return launch_map<T, 2, FGe<T>>(a);
    case AG_B_ATAN2: return launch_map<T, 2, FAtan2<T>>(a);
    case AG_B_HYPOT: return launch_map<T, 2, FHypot<T>>(a);
  }
  set_error("ag_binary_fwd: unknown op %d", op);
  return AG_EUNSUPPORTED;
}

template <typename T>
static ag_status ternary_dispatch(int which, MapArgs<3>& a) {
  switch (which) {
    case 0: return launch_map<T, 3, GDiv2<T>>(a);
    case 1: return launch_map<T, 3, GEqMask<T>>(a);
    case 2: return launch_map<T, 3, GPow1<T>>(a);
    case 3: return launch_map<T, 3, GPow2<T>>(a);
    case 4: return launch_map<T, 3, GAtan1<T>>(a);
    case 5: return launch_map<T, 3, GAtan2<T>>(a);
    case 6: return launch_map<T, 3, GHypot<T>>(a);
  }
  return AG_EUNSUPPORTED;
}

// ---- special functions with an integer parameter (src/specialfunctions.jl:20,25 besselj / bessely of integer order;
// :59 polygamma(m, x); :43 invdigamma): not in the unary table (they carry a parameter), so they run as their own
// pass.  polygamma(m, x) = (-1)^(m+1) m! zeta(m+1, x) with the Hurwitz zeta function by Euler-Maclaurin summation
// (x > 0; evaluated in double for both eltypes); invdigamma by Newton steps on digamma (Minka's start).
enum { AG_SP_BESSELJ = 0, AG_SP_BESSELY = 1, AG_SP_POLYGAMMA = 2, AG_SP_INVDIGAMMA = 3, AG_SP_DAWSON = 4, AG_SP_ERFI = 5,
       AG_SP_AIRYAI = 6, AG_SP_AIRYAIPRIME = 7, AG_SP_AIRYBI = 8, AG_SP_AIRYBIPRIME = 9, AG_SP_BESSELI = 10, AG_SP_BESSELK = 11,
       AG_SP_FLOOR = 12, AG_SP_CEIL = 13, AG_SP_ROUND = 14, AG_SP_TRUNC = 15, AG_SP_EXPONENT = 16, AG_SP_SIGNIFICAND = 17,
       AG_SP_REM2PI = 18, AG_SP_COUNT = 19 };

__device__ __forceinline__ double hurwitz_zeta(int s, double q) {
  // sum_{k<N} (q+k)^-s + tail at a = q + N >= 12
  double sum = 0.0, a = q;
  while (a < 12.0) { sum += pow(a, -(double)s); a += 1.0; }
  const double ia = 1.0 / a, ia2 = ia * ia;
  double t = pow(a, 1.0 - (double)s) / ((double)s - 1.0) + 0.5 * pow(a, -(double)s);
  // B_2j / (2j)! * s (s+1) ... (s+2j-2) * a^(-s-2j+1)
  const double B[7] = {1.0 / 6, -1.0 / 30, 1.0 / 42, -1.0 / 30, 5.0 / 66, -691.0 / 2730, 7.0 / 6};
  double rising = (double)s, fact = 2.0, pw = pow(a, -(double)s - 1.0);
  for (int j = 1; j <= 7; ++j) {
    t += B[j - 1] / fact * rising * pw;
    rising *= ((double)s + 2 * j - 1) * ((double)s + 2 * j);
    fact *= (2.0 * j + 1) * (2.0 * j + 2);
    pw *= ia2;
  }
  return sum + t;
}


// ---- dawson / erfi / Airy functions (src/specialfunctions.jl:5-8,33,39): CUDA's math library has none of them.
// All evaluated in double for both eltypes.
//  dawson: Rybicki's sampling sum, F(x) = pi^-1/2 sum_{n odd} exp(-(x - n h)^2) / n with h = 0.2 (sampling error
//          exp(-(pi/2h)^2) ~ 1e-27), Maclaurin series below 0.2, asymptotic series above 100.
//  erfi:   2/sqrt(pi) exp(x^2) dawson(x).
//  Airy:   Maclaurin series for |x| < 2; beyond, the Laplace-type integral (DLMF 9.5.8)
//            Ai(z) = exp(-zeta) zeta^-1/6 / (sqrt(pi) 48^1/6 Gamma(5/6)) int_0^inf exp(-t) t^-1/6 (2 + t/zeta)^-1/6 dt
//          by generalised Gauss-Laguerre quadrature (alpha = -1/6, 40-point rule, nodes with weight > 1e-22 kept), its
//          derivative from the same nodes; for x < 0 one complex evaluation A = Ai(|x| e^{i pi/3}) gives all four
//          functions through the connection formulas Ai(-z) = 2 Re[e^{i pi/3} A], Bi(-z) = 2 Re[e^{-i pi/6} A];
//          Bi, Bi' for x > 0: the series (all terms positive) up to 8.5, the asymptotic expansion (DLMF 9.7.7-8) above.
constexpr int kAiryNQ = 28;
__constant__ double kAiryT[kAiryNQ] = {
    2.83891417994567609e-02, 1.70985378860034898e-01, 4.35871678341770485e-01, 8.23518257913030904e-01,
    1.33452543254227374e+00, 1.96968293206435074e+00, 2.72998134002859949e+00, 3.61662161916100899e+00,
    4.63102611052654112e+00, 5.77485171830547639e+00, 7.05000568630218716e+00, 8.45866437513237734e+00,
    1.00032955242749395e+01, 1.16866845947722418e+01, 1.35119659344693552e+01, 1.54826596959377145e+01,
    1.76027156808069130e+01, 1.98765656022785429e+01, 2.23091856773962789e+01, 2.49061720212974187e+01,
    2.76738320739497183e+01, 3.06192963295084120e+01, 3.37506560850239978e+01, 3.70771349708391185e+01,
    4.06093049694341346e+01, 4.43593619516066724e+01, 4.83414822434528233e+01, 5.25722917078504892e+01};
__constant__ double kAiryW[kAiryNQ] = {
    1.43720408803314909e-01, 2.30407559241880361e-01, 2.42253045521327898e-01, 2.03636639103440736e-01,
    1.43760630622921298e-01, 8.69128834706077852e-02, 4.54175001832914579e-02, 2.06118031206068975e-02,
    8.14278821268601093e-03, 2.80266075663376368e-03, 8.40337441621719040e-04, 2.19303732907765277e-04,
    4.97401659009251131e-05, 9.78508095920972090e-06, 1.66542824603694159e-06, 2.44502736799656883e-07,
    3.08537034236218771e-08, 3.33296072937279380e-09, 3.06781892365377496e-10, 2.39331309909008319e-11,
    1.57294707676285316e-12, 8.64936013017882508e-14, 3.94819816700652301e-15, 1.48271173048109927e-16,
    4.53390374815054902e-18, 1.11547980452036191e-19, 2.17766660589229045e-21, 3.31878891057973789e-23};

__device__ double dawson_(double x) {
  const double ax = fabs(x);
  if (ax < 0.2) {
    const double x2 = x * x;
    return x * (1.0 - 2.0 / 3 * x2 * (1.0 - 2.0 / 5 * x2 * (1.0 - 2.0 / 7 * x2 * (1.0 - 2.0 / 9 * x2 * (1.0 - 2.0 / 11 * x2 *
               (1.0 - 2.0 / 13 * x2 * (1.0 - 2.0 / 15 * x2)))))));
  }
  if (!(ax <= 100.0)) {        // also NaN
    const double i2 = 1.0 / (2.0 * x * x);
    return (1.0 / (2.0 * x)) * (1.0 + i2 * (1.0 + 3.0 * i2 * (1.0 + 5.0 * i2 * (1.0 + 7.0 * i2))));
  }
  const double H = 0.2;
  const int n0 = 2 * (int)(0.5 * ax / H + 0.5);
  const double xp = ax - n0 * H;
  double e1 = exp(2.0 * xp * H);
  const double e2 = e1 * e1;
  double d1 = n0 + 1.0, d2 = d1 - 2.0, s = 0.0;
  for (int i = 0; i < 18; ++i) {
    const double a = (2 * i + 1) * H;
    s += exp(-a * a) * (e1 / d1 + 1.0 / (d2 * e1));
    d1 += 2.0; d2 -= 2.0; e1 *= e2;
  }
  const double r = 0.56418958354775628695 * exp(-xp * xp) * s;
  return x < 0.0 ? -r : r;
}

struct Cx { double re, im; };
__device__ __forceinline__ Cx cmul(Cx a, Cx b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
__device__ __forceinline__ Cx cpolar(double r, double th) { double s, c; sincos(th, &s, &c); return {r * c, r * s}; }

// out[0..3] = Ai, Ai', Bi, Bi' at x
__device__ void airy_all(double x, double* out) {
  const double C1 = 0.355028053887817239, C2 = 0.258819403792806798, SQ3 = 1.7320508075688772935;
  const double CQ = 0.26218399708832302;           // 1 / (sqrt(pi) 48^(1/6) Gamma(5/6))
  const double ISP = 0.56418958354775628695;       // 1 / sqrt(pi)
  const double PI = 3.14159265358979323846;
  if (x != x) { out[0] = out[1] = out[2] = out[3] = x; return; }
  const bool series = (x > -2.0 && x < 2.0);
  if (series || (x >= 2.0 && x <= 8.5)) {
    const double x3 = x * x * x;
    double a = 1.0, f = 1.0, b = x, g = x, ck = 0.5 * x * x, fp = ck, d = 1.0, gp = 1.0;
    for (int k = 0; k < 120; ++k) {
      a *= x3 / ((3.0 * k + 2) * (3.0 * k + 3)); f += a;
      b *= x3 / ((3.0 * k + 3) * (3.0 * k + 4)); g += b;
      if (k > 0) { ck *= x3 / ((3.0 * k) * (3.0 * k + 2)); fp += ck; }
      d *= x3 / ((3.0 * k + 1) * (3.0 * k + 3)); gp += d;
      if (k > 2 && fabs(a) < 1e-17 * fabs(f) && fabs(b) <= 1e-17 * fabs(g) && fabs(d) < 1e-17 * fabs(gp)) break;
    }
    out[2] = SQ3 * (C1 * f + C2 * g);
    out[3] = SQ3 * (C1 * fp + C2 * gp);
    if (series) { out[0] = C1 * f - C2 * g; out[1] = C1 * fp - C2 * gp; return; }
  }
  if (x >= 2.0) {
    const double sx = sqrt(x), z = (2.0 / 3) * x * sx;
    double I = 0.0, J = 0.0;
    for (int j = 0; j < kAiryNQ; ++j) {
      const double w = 2.0 + kAiryT[j] / z, p = pow(w, -1.0 / 6);
      I += kAiryW[j] * p;
      J += kAiryW[j] * kAiryT[j] * p / w;
    }
    const double pre = CQ * exp(-z) * pow(z, -1.0 / 6);
    out[0] = pre * I;
    out[1] = pre * sx * (-I - I / (6.0 * z) + J / (6.0 * z * z));
    if (x > 8.5) {                                  // Bi, Bi': asymptotic expansion
      double u = 1.0, su = 1.0, sv = 1.0;
      for (int k = 0; k < 40; ++k) {
        const double un = u * ((6.0 * k + 5) * (6.0 * k + 3) * (6.0 * k + 1)) / (216.0 * (k + 1) * (2.0 * k + 1)) / z;
        if (fabs(un) > fabs(u)) break;
        u = un;
        su += u;
        sv += u * (6.0 * (k + 1) + 1) / (1.0 - 6.0 * (k + 1));
        if (fabs(u) < 1e-17) break;
      }
      const double q = sqrt(sx), ez = exp(z);
      out[2] = ISP * ez / q * su;
      out[3] = ISP * ez * q * sv;
    }
    return;
  }
  // x <= -2: A = Ai(Z), Z = |x| e^{i pi/3}, zeta = i zr
  const double zz = -x, sz = sqrt(zz), zr = (2.0 / 3) * zz * sz;
  Cx I = {0.0, 0.0}, J = {0.0, 0.0};
  for (int j = 0; j < kAiryNQ; ++j) {
    const double im = -kAiryT[j] / zr, r2 = 4.0 + im * im, th = atan2(im, 2.0);
    const Cx p = cpolar(pow(r2, -1.0 / 12), -th / 6);                 // w^(-1/6)
    const Cx q = cmul(p, Cx{2.0 / r2, -im / r2});                     // w^(-7/6)
    I.re += kAiryW[j] * p.re; I.im += kAiryW[j] * p.im;
    const double wt = kAiryW[j] * kAiryT[j];
    J.re += wt * q.re; J.im += wt * q.im;
  }
  const Cx pre = cpolar(CQ * pow(zr, -1.0 / 6), -zr - PI / 12);       // C e^{-i zr} zeta^(-1/6)
  const Cx A = cmul(pre, I);
  // -I - I/(6 zeta) + J/(6 zeta^2), zeta = i zr
  const Cx br = {-I.re - I.im / (6.0 * zr) - J.re / (6.0 * zr * zr), -I.im + I.re / (6.0 * zr) - J.im / (6.0 * zr * zr)};
  const Cx Ap = cmul(cmul(pre, cpolar(sz, PI / 6)), br);
  out[0] = 2.0 * cmul(cpolar(1.0, PI / 3), A).re;
  out[2] = 2.0 * cmul(cpolar(1.0, -PI / 6), A).re;
  out[1] = -2.0 * cmul(cpolar(1.0, 2 * PI / 3), Ap).re;
  out[3] = -2.0 * cmul(cpolar(1.0, PI / 6), Ap).re;
}


// ---- modified Bessel functions of integer order (src/specialfunctions.jl:13,24), in double.
//  besseli(m, x): the ascending series sum_k (x/2)^(2k+m) / (k! (k+m)!) -- every term positive, so no cancellation at any
//                 x up to the overflow of the function itself (x ~ 713); I_m(-x) = (-1)^m I_m(x), I_-m = I_m.
//  besselk(m, x): K_0 and K_1 from K_m(x) = int_0^inf exp(-x cosh t) cosh(m t) dt by the trapezoid rule (the integrand is analytic and
//                 decays double-exponentially: the rule converges geometrically), step min(1/4, 0.7/sqrt(x)) so that the
//                 Gaussian peak of width 1/sqrt(x) at large x stays resolved, then the upward recurrence
//                 K_(n+1) = K_(n-1) + (2n/x) K_n (stable for K); x > 0 only (NaN below, Inf at 0).
__device__ double besseli_(int m, double x) {
  if (x != x) return x;
  if (m < 0) m = -m;
  const double ax = fabs(x);
  const double sgn = (x < 0.0 && (m & 1)) ? -1.0 : 1.0;
  if (ax == 0.0) return m == 0 ? 1.0 : 0.0 * sgn;
  if (ax > 745.0 + 0.5 * m) return sgn * INFINITY;
  const double q = 0.25 * ax * ax;
  double t = exp(m * log(0.5 * ax) - lgamma(m + 1.0)), s = t;
  for (int k = 1; k < 1500; ++k) {
    t *= q / ((double)k * (double)(k + m));
    s += t;
    if (t < 1e-17 * s) break;
  }
  return sgn * s;
}

__device__ double besselk_(int m, double x) {
  if (!(x > 0.0)) return x == 0.0 ? INFINITY : NAN;
  if (m < 0) m = -m;
  const double h = fmin(0.25, 0.7 / sqrt(x));
  double s0 = 0.5 * exp(-x), s1 = s0;          // t = 0 with half weight: K0 and K1 integrands
  for (int k = 1; k < 4000; ++k) {
    const double c = cosh(k * h), f = exp(-x * c);
    s0 += f;
    s1 += f * c;
    if (f * c < 1e-18 * s0) break;
  }
  double km = h * s0, k = h * s1;              // K0, K1; upward recurrence (stable for K)
  if (m == 0) return km;
  for (int n = 1; n < m; ++n) {
    const double kn = km + (2.0 * n / x) * k;
    km = k;
    k = kn;
  }
  return k;
}

template <typename T>
__global__ void __launch_bounds__(256) k_special_param(int op, int m, const T* __restrict__ x, T* __restrict__ y, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const T v = x[i];
  T r;
  if (op == AG_SP_BESSELJ || op == AG_SP_BESSELY) {
    // negative integer order: f(-m, x) = (-1)^m f(m, x)   (the rule of order 0 asks for order -1)
    const int am = m < 0 ? -m : m;
    const T sgn = (m < 0 && (am & 1)) ? T(-1) : T(1);
    if (op == AG_SP_BESSELJ) r = sizeof(T) == 4 ? (T)jnf(am, (float)v) : (T)jn(am, (double)v);
    else r = sizeof(T) == 4 ? (T)ynf(am, (float)v) : (T)yn(am, (double)v);
    r *= sgn;
  } else if (op == AG_SP_POLYGAMMA) {
    if (m == 0) r = digamma_<T>(v);
    else if (m == 1) r = trigamma_<T>(v);
    else if (!(v > T(0))) r = (T)NAN;                     // reflection for x <= 0 is not implemented
    else {
      double f = 1.0;
      for (int k = 2; k <= m; ++k) f *= k;
      const double z = hurwitz_zeta(m + 1, (double)v);
      r = (T)(((m & 1) ? 1.0 : -1.0) * f * z);
    }
  } else if (op >= AG_SP_FLOOR && op <= AG_SP_TRUNC) {    // @zerograd floor / ceil / round (ties to even) / trunc
    r = op == AG_SP_FLOOR ? floor(v) : op == AG_SP_CEIL ? ceil(v) : op == AG_SP_ROUND ? rint(v) : trunc(v);
  } else if (op == AG_SP_EXPONENT) {                      // src/math.jl:44 (@zerograd): floor(log2(|x|))
    r = (T)ilogb((double)v);
  } else if (op == AG_SP_SIGNIFICAND) {                   // :64  x / 2^exponent(x), in [1, 2)
    const double xv = (double)v;
    r = (xv == 0.0 || !isfinite(xv)) ? v : (T)scalbn(xv, -ilogb(xv));
  } else if (op == AG_SP_REM2PI) {                        // :56,59  m: 0 RoundNearest, 1 RoundToZero, 2 RoundDown (= mod2pi), 3 RoundUp
    const double P2 = 6.283185307179586476925, xv = (double)v;
    double q = m == 0 ? remainder(xv, P2) : fmod(xv, P2);
    if (m == 2 && q < 0.0) q += P2;
    if (m == 3 && q > 0.0) q -= P2;
    r = (T)q;
  } else if (op == AG_SP_BESSELI) {
    r = (T)besseli_(m, (double)v);
  } else if (op == AG_SP_BESSELK) {
    r = (T)besselk_(m, (double)v);
  } else if (op == AG_SP_DAWSON) {
    r = (T)dawson_((double)v);
  } else if (op == AG_SP_ERFI) {                          // 2/sqrt(pi) exp(x^2) dawson(x)
    const double xv = (double)v;
    r = (T)(1.12837916709551257390 * exp(xv * xv) * dawson_(xv));
  } else if (op >= AG_SP_AIRYAI && op <= AG_SP_AIRYBIPRIME) {
    double o[4];
    airy_all((double)v, o);
    r = (T)o[op - AG_SP_AIRYAI];
  } else {                                                // invdigamma
    double yv = (double)v;
    double xx = yv >= -2.22 ? exp(yv) + 0.5 : -1.0 / (yv + 0.5772156649015329);
    for (int it = 0; it < 8; ++it) xx -= ((double)digamma_<double>(xx) - yv) / (double)trigamma_<double>(xx);
    r = (T)xx;
  }
  y[i] = r;
}

}  // namespace agb

using namespace agb;

extern "C" {

ag_status ag_special_param(int32_t op, int32_t dtype, int32_t m, const void* x, void* y, int64_t n) {
  AG_REQUIRE_INIT();
  if (n == 0) return AG_OK;
  AG_CHECK(x && y && n > 0, AG_EINVAL, "ag_special_param: bad arguments");
  AG_CHECK(dtype == AG_F32 || dtype == AG_F64, AG_EUNSUPPORTED, "ag_special_param: dtype %d", dtype);
  AG_CHECK(op >= AG_SP_BESSELJ && op < AG_SP_COUNT, AG_EUNSUPPORTED, "ag_special_param: op %d", op);
  AG_CHECK(op != AG_SP_POLYGAMMA || (m >= 0 && m <= 20), AG_EUNSUPPORTED, "ag_special_param: polygamma order %d", m);
  Context& c = ctx();
  const unsigned nb = (unsigned)cdiv(n, (int64_t)256);
  if (dtype == AG_F32) k_special_param<float><<<nb, 256, 0, c.stream>>>(op, m, (const float*)x, (float*)y, n);
  else k_special_param<double><<<nb, 256, 0, c.stream>>>(op, m, (const double*)x, (double*)y, n);
  AG_LAUNCHED();
  return AG_OK;
}

ag_status ag_unary_fwd(int32_t op, int32_t dtype, const void* x, void* y, int64_t n) {
  AG_REQUIRE_INIT();
  if (n == 0) return AG_OK;
  AG_CHECK(x && y && n > 0, AG_EINVAL, "ag_unary_fwd: bad arguments");
  AG_CHECK(dtype == AG_F32 || dtype == AG_F64, AG_EUNSUPPORTED, "ag_unary_fwd: dtype %d", dtype);
#define X(OP)                                                                  \
  case OP:                                                                     \
    return dtype == AG_F32 ? launch_unary_fwd<OP, float>(x, y, n)              \
                           : launch_unary_fwd<OP, double>(x, y, n);
  switch (op) { AG_FOR_UNARY(X) }
#undef X
  set_error("ag_unary_fwd: unknown op %d", op);
  return AG_EUNSUPPORTED;
}

ag_status ag_unary_bwd(int32_t op, int32_t dtype, const void* dy, int32_t dy_mode, double dy_scalar,
                       const void* x, const void* y, void* dx, int64_t n) {
  AG_REQUIRE_INIT();
  if (n == 0) return AG_OK;
  AG_CHECK(dx && n > 0, AG_EINVAL, "ag_unary_bwd: bad arguments");
  AG_CHECK(dy_mode >= 0 && dy_mode <= 2, AG_EINVAL, "ag_unary_bwd: dy_mode %d", dy_mode);
  AG_CHECK(dy_mode == 2 || dy, AG_EINVAL, "ag_unary_bwd: dy is null");
  AG_CHECK(dtype == AG_F32 || dtype == AG_F64, AG_EUNSUPPORTED, "ag_unary_bwd: dtype %d", dtype);
#define X(OP)                                                                                  \
  case OP:                                                                                     \
    return dtype == AG_F32 ? launch_unary_bwd<OP, float>(dy, dy_mode, dy_scalar, x, y, dx, n)  \
                           : launch_unary_bwd<OP, double>(dy, dy_mode, dy_scalar, x, y, dx, n);
  switch (op) { AG_FOR_UNARY(X) }
#undef X
  set_error("ag_unary_bwd: unknown op %d", op);
  return AG_EUNSUPPORTED;
}

ag_status ag_binary_fwd(int32_t op, int32_t dtype, const void* a, double a_scalar,
                        const int64_t* a_strides, const void* b, double b_scalar,
                        const int64_t* b_strides, void* y, int32_t ndims, const int64_t* shape) {
  AG_REQUIRE_INIT();
  AG_CHECK(dtype == AG_F32 || dtype == AG_F64, AG_EUNSUPPORTED, "ag_binary_fwd: dtype %d", dtype);
  MapArgs<2> m;
  ag_status s = fill_shape(ndims, shape, m.shape, &m.n);
  if (s != AG_OK) return s;
  if (m.n == 0) return AG_OK;
  AG_CHECK(y, AG_EINVAL, "ag_binary_fwd: null output");
  AG_CHECK((!a || a_strides) && (!b || b_strides), AG_EINVAL, "ag_binary_fwd: missing strides");
  m.nd = ndims;
  m.out = y;
  set_opnd(m.in[0], a, a_scalar, a_strides, ndims);
  set_opnd(m.in[1], b, b_scalar, b_strides, ndims);
  return dtype == AG_F32 ? binary_dispatch<float>(op, m) : binary_dispatch<double>(op, m);
}

ag_status ag_binary_bwd(int32_t op, int32_t argno, int32_t dtype, const void* dy,
                        const int64_t* dy_strides, const void* x1, double x1_scalar,
                        const int64_t* x1_strides, const void* x2, double x2_scalar,
                        const int64_t* x2_strides, const void* y, const int64_t* y_strides,
                        void* dx, int32_t ndims, const int64_t* shape) {
  AG_REQUIRE_INIT();
  AG_CHECK(dtype == AG_F32 || dtype == AG_F64, AG_EUNSUPPORTED, "ag_binary_bwd: dtype %d", dtype);
  AG_CHECK(argno == 1 || argno == 2, AG_EINVAL, "ag_binary_bwd: argno %d", argno);
  AG_CHECK(dy && dy_strides, AG_EINVAL, "ag_binary_bwd: dy must be a device array");
  AG_CHECK((!x1 || x1_strides) && (!x2 || x2_strides) && (!y || y_strides), AG_EINVAL,
           "ag_binary_bwd: missing strides");
  int64_t shp[4], n;
  ag_status s = fill_shape(ndims, shape, shp, &n);
  if (s != AG_OK) return s;
  if (n == 0) return AG_OK;
  AG_CHECK(dx, AG_EINVAL, "ag_binary_bwd: null output");

  // two-operand gradients reuse the forward map: dx = dy (op') other
  int bop = -1;
  const void* o_p = nullptr;
  double o_imm = 0;
  const int64_t* o_s = nullptr;
  switch (op) {
    case AG_B_ADD: bop = AG_B_MUL; o_imm = 1.0; break;                       // base.jl:21
    case AG_B_SUB: bop = AG_B_MUL; o_imm = argno == 1 ? 1.0 : -1.0; break;   // base.jl:23
    case AG_B_MUL:                                                           // base.jl:27
      bop = AG_B_MUL;
      if (argno == 1) { o_p = x2; o_imm = x2_scalar; o_s = x2_strides; }
      else { o_p = x1; o_imm = x1_scalar; o_s = x1_strides; }
      break;
    case AG_B_DIV:                                                           // base.jl:33
      if (argno == 1) { bop = AG_B_DIV; o_p = x2; o_imm = x2_scalar; o_s = x2_strides; }
      break;
    default: break;
  }
  if (bop >= 0) {
    MapArgs<2> m;
    for (int d = 0; d < 4; ++d) m.shape[d] = shp[d];
    m.n = n;
    m.nd = ndims;
    m.out = dx;
    set_opnd(m.in[0], dy, 0, dy_strides, ndims);
    set_opnd(m.in[1], o_p, o_imm, o_s, ndims);
    return dtype == AG_F32 ? binary_dispatch<float>(bop, m) : binary_dispatch<double>(bop, m);
  }

  MapArgs<3> m;
  for (int d = 0; d < 4; ++d) m.shape[d] = shp[d];
  m.n = n;
  m.nd = ndims;
  m.out = dx;
  set_opnd(m.in[0], dy, 0, dy_strides, ndims);
  int which = -1;
  const void* xi = argno == 1 ? x1 : x2;
  double xi_imm = argno == 1 ? x1_scalar : x2_scalar;
  const int64_t* xi_s = argno == 1 ? x1_strides : x2_strides;
  switch (op) {
    case AG_B_DIV:  // arg2
      which = 0;
      set_opnd(m.in[1], x1, x1_scalar, x1_strides, ndims);
      set_opnd(m.in[2], x2, x2_scalar, x2_strides, ndims);
      break;
    case AG_B_MAX:
    case AG_B_MIN:
      AG_CHECK(y, AG_EINVAL, "ag_binary_bwd: max/min need y");
      which = 1;
      set_opnd(m.in[1], y, 0, y_strides, ndims);
      set_opnd(m.in[2], xi, xi_imm, xi_s, ndims);
      break;
    case AG_B_POW:
      if (argno == 1) {
        which = 2;
        set_opnd(m.in[1], x1, x1_scalar, x1_strides, ndims);
        set_opnd(m.in[2], x2, x2_scalar, x2_strides, ndims);
      } else {
        AG_CHECK(y, AG_EINVAL, "ag_binary_bwd: pow arg2 needs y");
        which = 3;
        set_opnd(m.in[1], y, 0, y_strides, ndims);
        set_opnd(m.in[2], x1, x1_scalar, x1_strides, ndims);
      }
      break;
    case AG_B_ATAN2:
      which = argno == 1 ? 4 : 5;
      set_opnd(m.in[1], x1, x1_scalar, x1_strides, ndims);
      set_opnd(m.in[2], x2, x2_scalar, x2_strides, ndims);
      break;
    case AG_B_HYPOT:
      AG_CHECK(y, AG_EINVAL, "ag_binary_bwd: hypot needs y");
      which = 6;
      set_opnd(m.in[1], xi, xi_imm, xi_s, ndims);
      set_opnd(m.in[2], y, 0, y_strides, ndims);
      break;
    default:
      set_error("ag_binary_bwd: op %d has no gradient (zerograd)", op);
      return AG_EUNSUPPORTED;
  }
  return dtype == AG_F32 ? ternary_dispatch<float>(which, m) : ternary_dispatch<double>(which, m);
}

ag_status ag_ewise_vm(int32_t dtype, const int32_t* code, int32_t ncode, const double* consts, int32_t nconsts,
                      int32_t nin, const void* const* in, const int32_t* in_mode, const double* in_imm, void* out,
                      int64_t n) {
  AG_REQUIRE_INIT();
  AG_CHECK(dtype == AG_F32 || dtype == AG_F64, AG_EUNSUPPORTED, "ag_ewise_vm: dtype %d", dtype);
  AG_CHECK(code && ncode >= 1 && ncode <= VM_MAXCODE, AG_EUNSUPPORTED, "ag_ewise_vm: program length %d (max %d)", ncode,
           VM_MAXCODE);
  AG_CHECK(nconsts >= 0 && nconsts <= VM_MAXCONST && nin >= 0 && nin <= VM_MAXIN, AG_EUNSUPPORTED,
           "ag_ewise_vm: %d constants / %d inputs (max %d / %d)", nconsts, nin, VM_MAXCONST, VM_MAXIN);
  if (n == 0) return AG_OK;
  AG_CHECK(out && n > 0, AG_EINVAL, "ag_ewise_vm: bad output");
  VmArgs a;
  // validate the program: operand indices and stack discipline
  int depth = 0;
  for (int pc = 0; pc < ncode; ++pc) {
    const int op = code[pc] & 0xff, arg = code[pc] >> 8;
    if (op == VM_IN) { AG_CHECK(arg >= 0 && arg < nin, AG_EINVAL, "ag_ewise_vm: input %d of %d", arg, nin); ++depth; }
    else if (op == VM_CONST) { AG_CHECK(arg >= 0 && arg < nconsts, AG_EINVAL, "ag_ewise_vm: constant %d of %d", arg, nconsts); ++depth; }
    else if (op == VM_NEG || op >= VM_UNARY) {
      AG_CHECK(depth >= 1 && (op == VM_NEG || op - VM_UNARY < AG_U_COUNT), AG_EINVAL, "ag_ewise_vm: bad unary op %d", op);
    } else { AG_CHECK(depth >= 2, AG_EINVAL, "ag_ewise_vm: stack underflow at %d", pc); --depth; }
    AG_CHECK(depth <= VM_STACK, AG_EUNSUPPORTED, "ag_ewise_vm: expression too deep (stack %d)", VM_STACK);
    a.code[pc] = code[pc];
  }
  AG_CHECK(depth == 1, AG_EINVAL, "ag_ewise_vm: program leaves %d values on the stack", depth);
  for (int k = 0; k < nconsts; ++k) a.consts[k] = consts[k];
  bool vec = aligned16(out);
  for (int k = 0; k < nin; ++k) {
    a.mode[k] = in_mode ? in_mode[k] : 0;
    a.in[k] = in ? in[k] : nullptr;
    a.imm[k] = in_imm ? in_imm[k] : 0.0;
    AG_CHECK(a.mode[k] >= 0 && a.mode[k] <= 2 && (a.mode[k] == 2 || a.in[k]), AG_EINVAL, "ag_ewise_vm: bad input %d", k);
    if (a.mode[k] == 0) vec = vec && aligned16(a.in[k]);
  }
  a.ncode = ncode;
  a.nin = nin;
  a.out = out;
  a.n = n;
  Context& c = ctx();
  const int N = dtype == AG_F32 ? 4 : 2;
  vec = vec && (n % N == 0);
  if (vec) {
    unsigned nb = (unsigned)cdiv(n / N, kThreads);
    if (dtype == AG_F32) k_vm<float, true><<<nb, kThreads, 0, c.stream>>>(a);
    else k_vm<double, true><<<nb, kThreads, 0, c.stream>>>(a);
  } else {
    unsigned nb = (unsigned)cdiv(n, kThreads);
    if (dtype == AG_F32) k_vm<float, false><<<nb, kThreads, 0, c.stream>>>(a);
    else k_vm<double, false><<<nb, kThreads, 0, c.stream>>>(a);
  }
  AG_LAUNCHED();
  return AG_OK;
}

ag_status ag_add(int32_t dtype, const void* a, const void* b, void* out, int64_t n) {
  int64_t st = 1, shape = n;
  return ag_binary_fwd(AG_B_ADD, dtype, a, 0, &st, b, 0, &st, out, 1, &shape);
}

}  // extern "C"
