// ewise_ops.cuh -- the elementwise op tables shared by the single-rule kernels (ewise.cu) and the fused
// evaluator (fuse.cu): unary f / g = df/dx table (src/math.jl:9-74, src/base.jl:73-74), binary maps and the
// ternary gradient maps of the broadcasting binary rules (src/base.jl:21-49, src/math.jl:24,48,54-55).
// One definition of every expression, so a fused program is bit-identical to the chain of single kernels.
#pragma once
#include "common.cuh"

namespace agb {

template <typename T>
struct alignas(16) Pack {
  static constexpr int N = 16 / sizeof(T);
  T v[N];
};

constexpr int kThreads = 256;
constexpr int kUnroll = 4;

// ---- unary op table -----------------------------------------------------------------------------
template <int OP, typename T>
struct U;

#define AG_DEF_U(OP, NX, NY, FEXPR, GEXPR)                         \
  template <typename T>                                            \
  struct U<OP, T> {                                                \
    static constexpr bool needs_x = NX, needs_y = NY;              \
    static __device__ __forceinline__ T f(T x) { return FEXPR; }   \
    static __device__ __forceinline__ T g(T x, T y) { return GEXPR; } \
  };

template <typename T>
__device__ __forceinline__ T sgn(T x) {
  return x > T(0) ? T(1) : (x < T(0) ? T(-1) : x);  // sign(0)=0, sign(NaN)=NaN like Julia
}
template <typename T>
__device__ __forceinline__ T relu(T x) {
  return x > T(0) ? x : (x != x ? x : T(0));
}


// ---- special functions (src/specialfunctions.jl): what CUDA's math library has (erf family, gamma, lgamma, Bessel
// j0/j1/jn, y0/y1/yn) plus digamma / trigamma / polygamma(2,.) for the gamma-family gradients: upward recurrence to
// x >= 8 (12 in double), then the asymptotic series; reflection for x < 0.5 (digamma, trigamma).
template <typename T>
__device__ __forceinline__ T digamma_(T x) {
  const T PI = T(3.14159265358979323846);
  T refl = T(0);
  bool neg = false;
  if (x < T(0.5)) {            // psi(x) = psi(1-x) - pi*cot(pi*x)
    refl = PI / tan(PI * x);
    x = T(1) - x;
    neg = true;
  }
  T r = T(0);
  const T big = sizeof(T) == 8 ? T(12) : T(8);
  while (x < big) { r -= T(1) / x; x += T(1); }
  const T i = T(1) / x, i2 = i * i;
  r += log(x) - T(0.5) * i -
       i2 * (T(1.0 / 12) - i2 * (T(1.0 / 120) - i2 * (T(1.0 / 252) - i2 * (T(1.0 / 240) - i2 * (T(1.0 / 132) - i2 * T(691.0 / 32760))))));
  return neg ? r - refl : r;
}
template <typename T>
__device__ __forceinline__ T trigamma_(T x) {
  const T PI = T(3.14159265358979323846);
  T refl = T(0);
  bool neg = false;
  if (x < T(0.5)) {            // psi1(x) = -psi1(1-x) + pi^2 / sin^2(pi*x)
    const T sp = sin(PI * x);
    refl = PI * PI / (sp * sp);
    x = T(1) - x;
    neg = true;
  }
  T r = T(0);
  const T big = sizeof(T) == 8 ? T(12) : T(8);
  while (x < big) { r += T(1) / (x * x); x += T(1); }
  const T i = T(1) / x, i2 = i * i;
  r += i + T(0.5) * i2 +
       i * i2 * (T(1.0 / 6) - i2 * (T(1.0 / 30) - i2 * (T(1.0 / 42) - i2 * (T(1.0 / 30) - i2 * (T(5.0 / 66) - i2 * T(691.0 / 2730))))));
  return neg ? refl - r : r;
}
template <typename T>
__device__ __forceinline__ T polygamma2_(T x) {   // psi''(x)
  const T PI = T(3.14159265358979323846);
  T refl = T(0);
  if (x < T(0.5)) {            // psi2(x) = psi2(1-x) - 2 pi^3 cos(pi x) / sin^3(pi x)
    const T sp = sin(PI * x);
    refl = T(2) * PI * PI * PI * cos(PI * x) / (sp * sp * sp);
    x = T(1) - x;
  }
  T r = T(0);
  const T big = sizeof(T) == 8 ? T(12) : T(8);
  while (x < big) { r -= T(2) / (x * x * x); x += T(1); }
  const T i = T(1) / x, i2 = i * i;
  r -= i2 * (T(1) + i + i2 * (T(0.5) - i2 * (T(1.0 / 6) - i2 * (T(1.0 / 6) - i2 * (T(3.0 / 10) - i2 * T(5.0 / 6))))));
  return r - refl;
}
__device__ __forceinline__ float bessel_jn2(float x) { return jnf(2, x); }
__device__ __forceinline__ double bessel_jn2(double x) { return jn(2, x); }
__device__ __forceinline__ float bessel_yn2(float x) { return ynf(2, x); }
__device__ __forceinline__ double bessel_yn2(double x) { return yn(2, x); }

//        op            x?     y?     f(x)                          g(x,y)
AG_DEF_U(AG_U_IDENTITY, false, false, x, T(1))
AG_DEF_U(AG_U_NEG, false, false, -x, T(-1))
AG_DEF_U(AG_U_ABS, true, false, fabs(x), sgn(x))
AG_DEF_U(AG_U_ABS2, true, false, x* x, T(2) * x)
AG_DEF_U(AG_U_EXP, false, true, exp(x), y)
AG_DEF_U(AG_U_LOG, true, false, log(x), T(1) / x)
AG_DEF_U(AG_U_TANH, false, true, tanh(x), T(1) - y * y)
AG_DEF_U(AG_U_SIGM, false, true, T(1) / (T(1) + exp(-x)), y*(T(1) - y))
AG_DEF_U(AG_U_SQRT, false, true, sqrt(x), T(1) / (T(2) * y))
AG_DEF_U(AG_U_SIN, true, false, sin(x), cos(x))
AG_DEF_U(AG_U_COS, true, false, cos(x), -sin(x))
AG_DEF_U(AG_U_TAN, false, true, tan(x), T(1) + y * y)
AG_DEF_U(AG_U_SINH, true, false, sinh(x), cosh(x))
AG_DEF_U(AG_U_COSH, true, false, cosh(x), sinh(x))
AG_DEF_U(AG_U_ASIN, true, false, asin(x), T(1) / sqrt(T(1) - x * x))
AG_DEF_U(AG_U_ACOS, true, false, acos(x), T(-1) / sqrt(T(1) - x * x))
AG_DEF_U(AG_U_ATAN, true, false, atan(x), T(1) / (T(1) + x * x))
AG_DEF_U(AG_U_ASINH, true, false, asinh(x), T(1) / sqrt(T(1) + x * x))
AG_DEF_U(AG_U_ACOSH, true, false, acosh(x), T(1) / sqrt(x * x - T(1)))
AG_DEF_U(AG_U_ATANH, true, false, atanh(x), T(1) / (T(1) - x * x))
AG_DEF_U(AG_U_EXP2, false, true, exp2(x), y* T(0.69314718055994530942))
AG_DEF_U(AG_U_EXP10, false, true, exp10(x), y* T(2.30258509299404568402))
AG_DEF_U(AG_U_EXPM1, false, true, expm1(x), T(1) + y)
AG_DEF_U(AG_U_LOG2, true, false, log2(x), T(1) / (T(0.69314718055994530942) * x))
AG_DEF_U(AG_U_LOG10, true, false, log10(x), T(1) / (T(2.30258509299404568402) * x))
AG_DEF_U(AG_U_LOG1P, true, false, log1p(x), T(1) / (T(1) + x))
AG_DEF_U(AG_U_CBRT, false, true, cbrt(x), T(1) / (T(3) * y * y))
AG_DEF_U(AG_U_SIGN, false, false, sgn(x), T(0))
AG_DEF_U(AG_U_RELU, true, false, relu(x), (x >= T(0) ? T(1) : T(0)))
AG_DEF_U(AG_U_ONE, false, false, T(1), T(0))
// src/specialfunctions.jl:35-41,44,63-64 (erf family), :42 (gamma), :63-64 (lgamma / loggamma), :34 (digamma),
// :67 (trigamma), :20-21,29-30 (Bessel functions of order 0 and 1)
AG_DEF_U(AG_U_ERF, true, false, erf(x), T(1.12837916709551257390) * exp(-(x * x)))
AG_DEF_U(AG_U_ERFC, true, false, erfc(x), -T(1.12837916709551257390) * exp(-(x * x)))
AG_DEF_U(AG_U_ERFCX, true, true, erfcx(x), (T(2) * y) * x - T(1.12837916709551257390))
AG_DEF_U(AG_U_ERFINV, false, true, erfinv(x), exp(y * y) * T(0.88622692545275801365))
AG_DEF_U(AG_U_ERFCINV, false, true, erfcinv(x), -(exp(y * y)) * T(0.88622692545275801365))
AG_DEF_U(AG_U_GAMMA, true, true, tgamma(x), y * digamma_(x))
AG_DEF_U(AG_U_LGAMMA, true, false, lgamma(x), digamma_(x))
AG_DEF_U(AG_U_DIGAMMA, true, false, digamma_(x), trigamma_(x))
AG_DEF_U(AG_U_TRIGAMMA, true, false, trigamma_(x), polygamma2_(x))
AG_DEF_U(AG_U_BESSELJ0, true, false, j0(x), -j1(x))
AG_DEF_U(AG_U_BESSELJ1, true, false, j1(x), (j0(x) - bessel_jn2(x)) / T(2))
AG_DEF_U(AG_U_BESSELY0, true, false, y0(x), -y1(x))
AG_DEF_U(AG_U_BESSELY1, true, false, y1(x), (y0(x) - bessel_yn2(x)) / T(2))

#define AG_FOR_UNARY(X)                                                                          \
  X(AG_U_IDENTITY) X(AG_U_NEG) X(AG_U_ABS) X(AG_U_ABS2) X(AG_U_EXP) X(AG_U_LOG) X(AG_U_TANH)     \
  X(AG_U_SIGM) X(AG_U_SQRT) X(AG_U_SIN) X(AG_U_COS) X(AG_U_TAN) X(AG_U_SINH) X(AG_U_COSH)        \
  X(AG_U_ASIN) X(AG_U_ACOS) X(AG_U_ATAN) X(AG_U_ASINH) X(AG_U_ACOSH) X(AG_U_ATANH) X(AG_U_EXP2)  \
  X(AG_U_EXP10) X(AG_U_EXPM1) X(AG_U_LOG2) X(AG_U_LOG10) X(AG_U_LOG1P) X(AG_U_CBRT) X(AG_U_SIGN) \
  X(AG_U_RELU) X(AG_U_ONE) X(AG_U_ERF) X(AG_U_ERFC) X(AG_U_ERFCX) X(AG_U_ERFINV) X(AG_U_ERFCINV) X(AG_U_GAMMA)      \
  X(AG_U_LGAMMA) X(AG_U_DIGAMMA) X(AG_U_TRIGAMMA) X(AG_U_BESSELJ0) X(AG_U_BESSELJ1) X(AG_U_BESSELY0)              \
  X(AG_U_BESSELY1)

// sign bit as a bit test (signbit() compiles to an out-of-line call on some paths: 11 us on the 64 x 65536 relu
// epilogue of the product, tests/probe_epilogue.py)
__device__ __forceinline__ bool sbit(float a) { return (__float_as_uint(a) >> 31) != 0u; }
__device__ __forceinline__ bool sbit(double a) { return (__double2hiint(a) >> 31) != 0; }
template <typename T>
__device__ __forceinline__ T jmax(T a, T b) {  // Julia max: NaN-propagating, max(-0.0, 0.0) = 0.0
  return (a != a || b != b) ? (a + b) : (a > b ? a : (b > a ? b : (sbit(a) ? b : a)));
}
template <typename T>
__device__ __forceinline__ T jmin(T a, T b) {
  return (a != a || b != b) ? (a + b) : (a < b ? a : (b < a ? b : (sbit(a) ? a : b)));
}

#define AG_DEF_F(NAME, EXPR)                                          \
  template <typename T>                                               \
  struct NAME {                                                       \
    static __device__ __forceinline__ T apply(const T* in) {          \
      const T a = in[0], b = in[1];                                   \
      (void)a; (void)b;                                               \
      return EXPR;                                                    \
    }                                                                 \
  };
AG_DEF_F(FAdd, a + b)
AG_DEF_F(FSub, a - b)
AG_DEF_F(FMul, a* b)
AG_DEF_F(FDiv, a / b)
AG_DEF_F(FMax, jmax(a, b))
AG_DEF_F(FMin, jmin(a, b))
AG_DEF_F(FPow, pow(a, b))
AG_DEF_F(FEq, (a == b ? T(1) : T(0)))
AG_DEF_F(FLt, (a < b ? T(1) : T(0)))
AG_DEF_F(FLe, (a <= b ? T(1) : T(0)))
AG_DEF_F(FGt, (a > b ? T(1) : T(0)))
AG_DEF_F(FGe, (a >= b ? T(1) : T(0)))
AG_DEF_F(FAtan2, atan2(a, b))
AG_DEF_F(FHypot, hypot(a, b))

// ternary gradient maps: in[0] = dy, in[1] = p, in[2] = q
#define AG_DEF_F3(NAME, EXPR)                                         \
  template <typename T>                                               \
  struct NAME {                                                       \
    static __device__ __forceinline__ T apply(const T* in) {          \
      const T dy = in[0], p = in[1], q = in[2];                       \
      return EXPR;                                                    \
    }                                                                 \
  };
AG_DEF_F3(GDiv2, -dy* p / (q * q))               // ./ arg2: -dy.*x1./abs2.(x2)   base.jl:33  (p=x1,q=x2)
AG_DEF_F3(GEqMask, dy*(p == q ? T(1) : T(0)))    // max/min: dy.*(y.==xi)         math.jl:54-55 (p=y,q=xi)
AG_DEF_F3(GPow1, dy* q* pow(p, q - T(1)))        // .^ arg1: dy.*x2.*x1.^(x2-1)   base.jl:49  (p=x1,q=x2)
AG_DEF_F3(GPow2, dy* p* log(q))                  // .^ arg2: dy.*y.*log.(x1)      base.jl:45  (p=y,q=x1)
AG_DEF_F3(GAtan1, dy* q / (p * p + q * q))       // atan arg1                     math.jl:24  (p=x1,q=x2)
AG_DEF_F3(GAtan2, dy*(-p) / (p * p + q * q))     // atan arg2
AG_DEF_F3(GHypot, dy* p / q)                     // hypot: dy.*(xi./y)            math.jl:48  (p=xi,q=y)

static inline bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

}  // namespace agb
