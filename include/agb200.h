/* agb200.h -- C ABI of libagb200.so: the B200 (sm_100a) array hot path behind AutoGrad.jl's
 * operator API.
 *
 * AutoGrad.jl (reference, /root/reference) has no FFI: its extension point is Julia multiple
 * dispatch on the *unboxed array type* that sits inside Param/Result (src/core.jl:66 calls
 * f(novalue...); gradient rules in src/base.jl, src/math.jl call generic array functions).  A device
 * array type plugs in by giving those generic functions methods that `ccall` into this library.
 * Every entry point below names the reference call site(s) it replaces.  See INTEGRATION.md for the
 * Julia-side binding.
 *
 * Conventions
 *   - plain C, `extern "C"`; every function returns ag_status (0 = ok); the message of the last
 *     failure on this thread is read with ag_last_error().  Nothing throws, nothing aborts.
 *   - arrays are raw device pointers to COLUMN-MAJOR (Julia layout) dense storage; sizes, strides
 *     and indices are int64_t (Julia Int); strides are in elements; index vectors handed to the
 *     gather/scatter entry points are 0-based flat column-major offsets (the Julia wrapper
 *     subtracts 1 when it normalises an index tuple, src/addto.jl:95).
 *   - dtype: AG_F32 / AG_F64.  Scalars cross the ABI as double and are narrowed on device.
 *   - all work is enqueued on ONE library stream (ag_stream()); only ag_sync, ag_copy_d2h,
 *     ag_event_elapsed and ag_graph_end block the host.  There is no CPU fallback: an unsupported
 *     request returns AG_EUNSUPPORTED.
 */
#ifndef AGB200_H
#define AGB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int32_t ag_status;

enum {
  AG_OK = 0,
  AG_EINVAL = 1,       /* bad argument / shape mismatch  -> Julia DimensionMismatch / ArgumentError */
  AG_ENOMEM = 2,       /* device or pinned allocation failed -> OutOfMemoryError */
  AG_ECUDA = 3,        /* a CUDA runtime/driver call failed */
  AG_ENCCL = 4,        /* an NCCL call failed */
  AG_EUNSUPPORTED = 5, /* dtype/op/rank not implemented on device -> MethodError (no CPU fallback) */
  AG_ESTATE = 6        /* call order violated (not initialised, graph capture state, ...) */
};

enum { AG_F32 = 0, AG_F64 = 1 };

/* ------------------------------------------------------------------------------------------------
 * Unary elementwise primitives.  fwd: y = f(x).  bwd: dx = dy .* g(x, y) -- the
 * `@primitive f(x),dy,y (dy.*(g))` rules of src/math.jl:9-74 and src/base.jl:73-74, plus the user
 * primitive `sigm` of examples/charlm.jl:46-47.
 * ---------------------------------------------------------------------------------------------- */
enum {
  AG_U_IDENTITY = 0, /* +(x): dy                      src/base.jl:20 */
  AG_U_NEG = 1,      /* -(x): -dy                     src/base.jl:22 */
  AG_U_ABS = 2,      /* dy.*sign.(x)                  src/base.jl:73 */
  AG_U_ABS2 = 3,     /* dy.*(2x)                      src/base.jl:74 */
  AG_U_EXP = 4,      /* dy.*y                         src/math.jl:42 */
  AG_U_LOG = 5,      /* dy.*(1 ./ x)                  src/math.jl:50 */
  AG_U_TANH = 6,     /* dy.*(1 .- abs2.(y))           src/math.jl:74 */
  AG_U_SIGM = 7,     /* dy.*y.*(1 .- y)               examples/charlm.jl:47 */
  AG_U_SQRT = 8,     /* dy.*(1 ./ (2 .* y))           src/math.jl:71 */
  AG_U_SIN = 9,      /* dy.*cos.(x)                   src/math.jl:65 */
  AG_U_COS = 10,     /* dy.*(-sin.(x))                src/math.jl:30 */
  AG_U_TAN = 11,     /* dy.*(1 .+ abs2.(y))           src/math.jl:72 */
  AG_U_SINH = 12,    /* dy.*cosh.(x)                  src/math.jl:68 */
  AG_U_COSH = 13,    /* dy.*sinh.(x)                  src/math.jl:33 */
  AG_U_ASIN = 14,    /* dy./sqrt(1-x^2)               src/math.jl:21 */
  AG_U_ACOS = 15,    /* -dy./sqrt(1-x^2)              src/math.jl:9 */
  AG_U_ATAN = 16,    /* dy./(1+x^2)                   src/math.jl:24 */
  AG_U_ASINH = 17,   /* dy./sqrt(1+x^2)               src/math.jl:23 */
  AG_U_ACOSH = 18,   /* dy./sqrt(x^2-1)               src/math.jl:11 */
  AG_U_ATANH = 19,   /* dy./(1-x^2)                   src/math.jl:26 */
  AG_U_EXP2 = 20,    /* dy.*y.*log(2)                 src/math.jl:44 */
  AG_U_EXP10 = 21,   /* dy.*y.*log(10)                src/math.jl:43 */
  AG_U_EXPM1 = 22,   /* dy.*(1 .+ y)                  src/math.jl:45 */
  AG_U_LOG2 = 23,    /* dy./(log(2).*x)               src/math.jl:53 */
  AG_U_LOG10 = 24,   /* dy./(log(10).*x)              src/math.jl:51 */
  AG_U_LOG1P = 25,   /* dy./(1 .+ x)                  src/math.jl:52 */
  AG_U_CBRT = 26,    /* dy./(3 .* abs2.(y))           src/math.jl:27 */
  AG_U_SIGN = 27,    /* @zerograd sign                src/base.jl (zero gradient; fwd only) */
  AG_U_RELU = 28,    /* max.(0,x): dy.*(y.==x)        src/math.jl:54 with x1 = 0 (examples/mnist.jl:32) */
  AG_U_ONE = 29,     /* one.(x) (fwd only; used by dy.*one.(x), src/base.jl:245) */
  /* src/specialfunctions.jl (the functions CUDA's math library provides, plus digamma / trigamma) */
  AG_U_ERF = 30,      /* dy.*exp(-x^2)*2/sqrt(pi)      src/specialfunctions.jl:35 */
  AG_U_ERFC = 31,     /* -dy.*exp(-x^2)*2/sqrt(pi)     :36 */
  AG_U_ERFCX = 32,    /* dy.*(2y.*x - 2/sqrt(pi))      :38 */
  AG_U_ERFINV = 33,   /* dy.*exp(y^2)*sqrt(pi)/2       :40 */
  AG_U_ERFCINV = 34,  /* -dy.*exp(y^2)*sqrt(pi)/2      :37 */
  AG_U_GAMMA = 35,    /* dy.*y.*digamma(x)             :42 */
  AG_U_LGAMMA = 36,   /* dy.*digamma(x)                :63-64 (lgamma / loggamma) */
  AG_U_DIGAMMA = 37,  /* dy.*trigamma(x)               :34 */
  AG_U_TRIGAMMA = 38, /* dy.*polygamma(2,x)            :67 */
  AG_U_BESSELJ0 = 39, /* -dy.*besselj1(x)              :20 */
  AG_U_BESSELJ1 = 40, /* dy.*(besselj0 - besselj(2,x))/2  :21 */
  AG_U_BESSELY0 = 41, /* -dy.*bessely1(x)              :29 */
  AG_U_BESSELY1 = 42, /* dy.*(bessely0 - bessely(2,x))/2  :30 */
  AG_U_COUNT = 43
};

/* Binary broadcasting primitives (src/base.jl:21-49, src/math.jl:24,48,54-55). */
enum {
  AG_B_ADD = 0,
  AG_B_SUB = 1,
  AG_B_MUL = 2,
  AG_B_DIV = 3,
  AG_B_MAX = 4,
  AG_B_MIN = 5,
  AG_B_POW = 6,
  AG_B_EQ = 7,     /* (a .== b) as 1.0/0.0 in the operand dtype: the mask in dy.*(y.==xi) */
  AG_B_LT = 8,     /* comparison masks (@zerograd, src/base.jl:52-60) */
  AG_B_LE = 9,
  AG_B_GT = 10,
  AG_B_GE = 11,
  AG_B_ATAN2 = 12, /* atan(x1,x2)  src/math.jl:24 */
  AG_B_HYPOT = 13, /* hypot        src/math.jl:48 */
  AG_B_COUNT = 14
};

/* Reduction pre-maps: sum(x), sum(abs2,x), sum(abs,x)  (src/base.jl:245-247). */
enum { AG_R_ID = 0, AG_R_ABS2 = 1, AG_R_ABS = 2 };

/* ---- lifecycle ------------------------------------------------------------------------------- */
/* Bind this process to one GPU, create the library stream and the caching pool.  Idempotent for the
 * same device.  Replaces nothing in the reference (CPU only); one rank == one process == one GPU. */
ag_status ag_init(int32_t device);
ag_status ag_shutdown(void);
/* cudaStreamSynchronize on the library stream: the `@gs` hook, src/AutoGrad.jl:11. */
ag_status ag_sync(void);
/* Copies the calling thread's last error message (NUL terminated) into buf. */
ag_status ag_last_error(char* buf, int64_t buflen);
/* info[0]=device, [1]=SM count, [2]=cc major, [3]=cc minor, [4]=total bytes, [5]=free bytes,
 * [6]=bytes held by the pool, [7]=bytes handed out. */
ag_status ag_device_info(int64_t* info8);
/* The library's cudaStream_t, as an integer, for interop (torch.cuda.ExternalStream, NCCL). */
ag_status ag_stream(uint64_t* stream_out);
/* Number of kernels this library has launched since ag_init (bench.py's gpu_launches). */
ag_status ag_launch_count(int64_t* count_out);

/* ---- memory: the storage behind the device array type --------------------------------------- */
/* Stream-ordered caching allocator.  ag_free never synchronises and is safe from a finalizer
 * (Julia GC) and from the gcnode hook (src/core.jl:168,235-237). */
ag_status ag_alloc(int64_t bytes, void** ptr_out);
ag_status ag_free(void* ptr);
ag_status ag_pool_trim(void);
/* Pinned host staging buffers for the h2d/d2h copies. */
ag_status ag_host_alloc(int64_t bytes, void** ptr_out);
ag_status ag_host_free(void* ptr);
ag_status ag_copy_h2d(void* dst_dev, const void* src_host, int64_t bytes); /* async on the stream */
/* Input prefetch (the `minibatch` loop of examples/mnist.jl:96-106 feeds one batch per step): the copy runs on a second
 * stream so that the NEXT batch crosses PCIe while the step on the current one executes.  ag_copy_h2d_prefetch waits
 * (on the device) until the previous contents of the staging buffer were released, then copies; ag_prefetch_wait makes
 * the main stream wait for the prefetched data; ag_prefetch_release (after the consumer of the staging buffer has been
 * enqueued) lets the next prefetch overwrite it.  src must be pinned (ag_host_alloc) for the copy to be asynchronous.
 * ag_prefetch_slot selects which of two (ready, released) event pairs the three calls use (default 0): a loader with
 * two staging buffers alternates the slot, so that the copy of batch i+1 waits for the consumer of batch i-1 only and
 * the copy engine does not idle while batch i is converted. */
ag_status ag_prefetch_slot(int32_t slot);
ag_status ag_copy_h2d_prefetch(void* dst_dev, const void* src_host, int64_t bytes);
ag_status ag_prefetch_wait(void);
ag_status ag_prefetch_release(void);
ag_status ag_copy_d2h(void* dst_host, const void* src_dev, int64_t bytes); /* blocks until done */
ag_status ag_copy_d2h_async(void* dst_host, const void* src_dev, int64_t bytes);
ag_status ag_copy_d2d(void* dst_dev, const void* src_dev, int64_t bytes);  /* copy(x), src/base.jl:338 */
/* fill!(similar(x), v): unbroadcast's scalar branch (src/unbroadcast.jl:14-15), zero(x), ones. */
ag_status ag_fill(void* dst, int64_t n, int32_t dtype, double value);

/* ---- timing / graphs ---------------------------------------------------------------------------- */
ag_status ag_event_create(uint64_t* ev_out);
ag_status ag_event_record(uint64_t ev);
ag_status ag_event_elapsed_ms(uint64_t ev_start, uint64_t ev_stop, float* ms_out); /* syncs on stop */
ag_status ag_event_destroy(uint64_t ev);
/* The reference's `@timer` (AUTOGRAD_TIMER; src/AutoGrad.jl:7-12) labels every forward call "[>ti]f", every `back` call
 * "[ti]f[i]" and every accumulation "[ti]addto![i]" (src/core.jl:165-166,177-182).  These two entries open / close an
 * NVTX range with such a label around the kernels a tape position launches, so nsys / ncu timelines carry the tape
 * position; the host also accumulates per-label wall time and launch counts (core.timer_report). */
ag_status ag_range_push(const char* label);
ag_status ag_range_pop(void);
/* Capture everything enqueued between begin/end (one recorded tape's forward+backward+update)
 * into a CUDA graph and replay it.  Pool blocks handed out during capture stay owned by the graph
 * until ag_graph_destroy.  Replaces the per-op dispatch of src/core.jl:64-79,156-169 on replay. */
ag_status ag_graph_begin(void);
ag_status ag_graph_end(uint64_t* graph_out);
ag_status ag_graph_launch(uint64_t graph);
ag_status ag_graph_destroy(uint64_t graph);

/* ---- elementwise ------------------------------------------------------------------------------ */
/* y[i] = f(x[i]), i < n.  Forward of every `f.(x)` node (src/core.jl:66-70 materialises it). */
ag_status ag_unary_fwd(int32_t op, int32_t dtype, const void* x, void* y, int64_t n);
/* Special functions that carry an integer parameter (not in the unary table): op 0 = besselj(m, x), 1 = bessely(m, x)
 * for integer order m (src/specialfunctions.jl:20,25; their rules (f(m-1,x) - f(m+1,x))/2 are composed by the host),
 * 2 = polygamma(m, x) for 0 <= m <= 20 (:59; m >= 2 needs x > 0, NaN otherwise: no reflection), 3 = invdigamma(x)
 * (:43; m ignored); m is also ignored by 4 = dawson(x) (:33), 5 = erfi(x) (:39), 6 = airyai(x), 7 = airyaiprime(x),
 * 8 = airybi(x), 9 = airybiprime(x) (:5-8), which CUDA's math library does not have (series / Gauss-Laguerre
 * quadrature / asymptotic expansion in double, csrc/ewise.cu); 10 = besseli(m, x), 11 = besselk(m, x) of integer order
 * (:13,24; ascending series / trapezoid rule for K0, K1 + upward recurrence; their rules (f(m-1,x) + f(m+1,x))/(+-2)
 * are composed by the host); 12 = floor, 13 = ceil, 14 = round (ties to even), 15 = trunc (the @zerograd rounding
 * functions of src/base.jl:79,100,145,151; m ignored); 16 = exponent(x), 17 = significand(x) (src/math.jl:44,64),
 * 18 = rem2pi(x, mode m: 0 nearest, 1 to zero, 2 down = mod2pi(x), 3 up; :56,59; reduction by the double 2*pi).
 * One pass, y[i] = f(m, x[i]). */
ag_status ag_special_param(int32_t op, int32_t dtype, int32_t m, const void* x, void* y, int64_t n);
/* inv(a) and the determinant of an n x n column-major matrix (n <= 1024) in one launch: Gauss-Jordan elimination with
 * partial pivoting in double.  inv_out (n x n, may be NULL) receives inv(a); info_out (3 values of the matrix dtype, may
 * be NULL) receives det(a), log|det(a)|, sign(det(a)).  A singular matrix gives det 0 and non-finite inverse entries.
 * Replaces the LAPACK calls behind det / inv / logdet / logabsdet and the inv(x)' of their rules
 * (src/linearalgebra.jl:18,28,57-58). */
ag_status ag_inv_det(int32_t dtype, const void* a, int64_t n, void* inv_out, void* info_out);
/* *idx_out (device int64) = 0-based position of the first maximum (want_max != 0) or minimum of x[0..n); a NaN wins,
 * like Julia's argmax / argmin (@zerograd, src/base.jl:60-61).  Two launches. */
ag_status ag_argext(int32_t dtype, const void* x, int64_t n, int32_t want_max, int64_t* idx_out);
/* y[i] = (dst eltype) x[i]: Float32 <-> Float64 conversion of a device array (oftype / big / Float64.(x),
 * src/base.jl:76,365).  Buffers 16-byte aligned (pool blocks are). */
ag_status ag_cast(int32_t src_dtype, int32_t dst_dtype, const void* x, void* y, int64_t n);
/* dx[i] = dy[i] * g(x[i], y[i]).  dy: device array (dy_mode 0), device scalar (1), or the
 * immediate dy_scalar (2).  x / y may be NULL when g does not use them.
 * Replaces the broadcast expression in each rule of src/math.jl:9-74. */
ag_status ag_unary_bwd(int32_t op, int32_t dtype, const void* dy, int32_t dy_mode, double dy_scalar,
                       const void* x, const void* y, void* dx, int64_t n);
/* y = op.(a, b) with Julia broadcasting.  shape[ndims] (ndims <= 4) is the result shape after the
 * caller has collapsed adjacent dimensions; a_strides/b_strides give each operand's element
 * stride per dimension (0 = broadcast along it).  a == NULL (b == NULL) means the operand is the
 * immediate a_scalar (b_scalar).  Forward of `+ - .* ./ max min .^` nodes (src/base.jl:21-45,
 * src/math.jl:54-55) and the building block of the unfused (higher-order-safe) gradient path. */
ag_status ag_binary_fwd(int32_t op, int32_t dtype, const void* a, double a_scalar,
                        const int64_t* a_strides, const void* b, double b_scalar,
                        const int64_t* b_strides, void* y, int32_t ndims, const int64_t* shape);
/* Fused gradient of a binary node w.r.t. argument argno (1 or 2), BEFORE unbroadcast:
 *   dx = dy .* d(op)/d(x_argno)(x1, x2, y), all operands broadcast to `shape` through strides
 * (x1/x2 may be immediates as above; y may be NULL where unused).  Only valid when no outer tape
 * is recording (src/core.jl:60-61), see SURVEY.md 3.4. */
ag_status ag_binary_bwd(int32_t op, int32_t argno, int32_t dtype, const void* dy,
                        const int64_t* dy_strides, const void* x1, double x1_scalar,
                        const int64_t* x1_strides, const void* x2, double x2_scalar,
                        const int64_t* x2_strides, const void* y, const int64_t* y_strides,
                        void* dx, int32_t ndims, const int64_t* shape);

/* Fused evaluator for user-defined primitives (`@primitive f(x),dy,y (expr)`, src/macros.jl:88-105; e.g. sigm,
 * examples/charlm.jl:46-47): out[i] = program(in0[i], ..., in3[i]) in one pass.  `code` is postfix bytecode, one
 * int32 per instruction = op | (arg << 8): 0 push input[arg], 1 push consts[arg], 2 + 3 - 4 * 5 / 6 neg 7 max 8 min
 * 9 pow 10 == 11 < 12 <= 13 > 14 >= (comparisons give 1.0/0.0), 15+u = unary f of AG_U_* id u.  <= 64
 * instructions, <= 16 constants, <= 4 inputs (in_mode: 0 array of n, 1 device scalar, 2 immediate in_imm),
 * stack depth <= 8.  Inputs must already have the output's shape (the caller unbroadcasts afterwards). */
ag_status ag_ewise_vm(int32_t dtype, const int32_t* code, int32_t ncode, const double* consts, int32_t nconsts,
                      int32_t nin, const void* const* in, const int32_t* in_mode, const double* in_imm, void* out,
                      int64_t n);

/* Horizontal fusion of a chain of elementwise tape operations: the host runtime defers dotted
 * calls (src/core.jl:68-70 evaluates each one as its own Base broadcast pass), gradient rules (src/math.jl:9-74,
 * src/base.jl:20-49,73-74) and the dense `addto!` (src/addto.jl:23), and evaluates the whole expression DAG with ONE
 * launch: each leaf read once, each value the tape still holds written once (up to 6 outputs; 12 inputs / 12 outputs for programs that have a specialised kernel), temporaries never
 * stored.  Iteration space rows x cols (column-major, n = rows*cols).  in_mode: 0 full array of n, 1 device scalar,
 * 2 immediate in_imm, 3 column vector (rows; broadcast along dim 2, e.g. a bias), 4 row vector (cols; broadcast along
 * dim 1, e.g. the softmax normaliser).  `code`: one int32 per instruction = op | a << 8 | b << 16, a stack machine
 * with 8 register slots (expression stack growing from slot 0; temporaries in higher slots):
 *   0 LD_IN a (b = 1: push, 0: overwrite top)   1 LD_CONST a   2 LD_TMP slot a   3 ST_TMP slot a (top kept)
 *   4 UNARY f of AG_U_* id a                    5 BIN: top = pop() (AG_B_* a) top
 *   6 BIN_IN: top = top (a) in[b & 127] (b & 128: operands reversed)   7 BIN_CONST: same with consts[b & 127]
 *   8 UGRAD: y = top, x = pop(), dy = pop(); top = dy * g_a(x, y)      (the rule table of ag_unary_bwd)
 *   9 TERN a: q = top, p = pop(), dy = pop(); top = G_a(dy, p, q)      (0 ./ arg2, 1 max/min mask, 2 .^ arg1,
 *     3 .^ arg2, 4 atan arg1, 5 atan arg2, 6 hypot: the ternary maps of ag_binary_bwd)
 *   10 STORE: out[a] = top.
 * <= 96 instructions, 16 constants, 8 inputs, 6 outputs.  Every instruction evaluates the same expression as the
 * corresponding single kernel, so results are bit-identical to the unfused sequence. */
ag_status ag_ewise_fused(int32_t dtype, const int32_t* code, int32_t ncode, const double* consts, int32_t nconsts,
                         int32_t nin, const void* const* in, const int32_t* in_mode, const double* in_imm,
                         int32_t nout, void* const* out, int64_t rows, int64_t cols);

/* The same launch followed by a sum of one of its values: unbroadcast(x, expr) = sum(expr; dims) (src/unbroadcast.jl:17-24),
 * sum / sum(abs2,.) of a chain (src/base.jl:245-247) without a second pass over the data.  The program contains exactly
 * one instruction 11 REDUCE (the top of stack at that point is the operand; nout may be 0).  The operand, viewed as
 * r_inner x r_red x r_outer (column-major, = rows*cols elements), is summed over r_red into r_out (r_inner*r_outer
 * values): whole array (1, n, 1), contiguous runs (1, red, outer) with red <= 2048, row sums (inner, red, 1) with
 * inner <= 128.  Deterministic (fixed order, double accumulation).  AG_EUNSUPPORTED when the layout or the program has no
 * specialised kernel -- the caller then evaluates the chain and reduces with ag_sum_dim. */
ag_status ag_ewise_fused_reduce(int32_t dtype, const int32_t* code, int32_t ncode, const double* consts,
                                int32_t nconsts, int32_t nin, const void* const* in, const int32_t* in_mode,
                                const double* in_imm, int32_t nout, void* const* out, int64_t rows, int64_t cols,
                                int64_t r_inner, int64_t r_red, int64_t r_outer, void* r_out);

/* A chain that sums over the rows of an m x n (column-major) matrix and USES the sums again, as ONE launch: the
 * softmax normaliser `ypred .- log.(sum(exp.(ypred),dims=1))` of the reference's models (examples/mnist.jl:40,
 * examples/charlm.jl:107-110, examples/rnn_lonely_integer.jl:52-55) and the unbroadcast of its gradient
 * (src/unbroadcast.jl:17-24), which the reference runs as one broadcast / reduction pass per node.
 * The program uses the instruction words of ag_ewise_fused (without the temporary-slot instructions 2 and 3) in
 * SEGMENTS introduced by instruction 12 PHASE: a = 0, an element segment, runs for every row of every column;
 * a = 1, a column segment, runs once per column.  Further instructions: 13 CRED a (element segment: column accumulator
 * a += top, in double), 16 CFIN a (column segment: column value a = top = the rounded accumulator a, which is cleared),
 * 15 ST_COL a (column value a = top), 14 LD_COL a (top = column value a; b = 1 pushes first; either kind of segment;
 * a < 8).  STORE writes an m x n output from an element segment and a 1 x n output from a column segment.  Inputs:
 * in_mode 0 = m x n array, 3 = m values (one per row), 4 = n values (one per column; the only array kind a column
 * segment may read), 1 = device scalar.  r_out != NULL: the program holds one instruction 11 REDUCE and the sum of
 * its operand over the whole segment space goes to r_out (one value; deterministic: per-block partials in double,
 * added in block order by the last block to finish).  Every instruction evaluates the expression of the single
 * kernels; a column sum is taken in double: every lane of the column's 8, 16 or 32 lanes adds its rows in order, the lanes
 * combine pairwise. */
ag_status ag_ewise_colprog(int32_t dtype, const int32_t* code, int32_t ncode, const double* consts, int32_t nconsts,
                           int32_t nin, const void* const* in, const int32_t* in_mode, const double* in_imm,
                           int32_t nout, void* const* out, int64_t m, int64_t n, void* r_out);

/* Programs that the BASELINE configs produce are also instantiated ahead of time (csrc/fuse_static.inc, generated
 * by tools/gen_fused_programs.py) and picked by an exact match of `code`; on = 0 forces the general interpreter
 * kernel for every program (default 1, or environment AGB_FUSED_STATIC=0).  Results are identical either way. */
ag_status ag_ewise_fused_set_static(int32_t on);
/* is_static = 1 when `code` has a specialised kernel.  The host uses it for GB-size arrays: a chain without one is
 * evaluated as smaller chains / single kernels there, because the general interpreter is issue-bound (about 40 % of the
 * memory roofline) and only pays off where launches, not bytes, dominate. */
ag_status ag_ewise_fused_query(const int32_t* code, int32_t ncode, int32_t* is_static);
/* out2[0] = launches that ran a specialised program, out2[1] = launches on the general interpreter. */
ag_status ag_ewise_fused_stats(int64_t* out2);

/* ---- reductions -------------------------------------------------------------------------------- */
/* out[0] = sum(map.(x)) over n elements; deterministic two-phase tree.  sum(x), sum(abs2,x),
 * sum(abs,x) forward (rules at src/base.jl:245-247); unbroadcast's Number branch
 * (src/unbroadcast.jl:12-13). */
ag_status ag_sum_all(int32_t map, int32_t dtype, const void* x, int64_t n, void* out_scalar);
/* x viewed as (inner, red, outer) column-major; out[i,o] = sum_r map(x[i,r,o]), out is
 * (inner, outer).  One call per reduced dimension: sum(x;dims=d) and the general branch of
 * unbroadcast (src/unbroadcast.jl:17-24). */
ag_status ag_sum_dim(int32_t map, int32_t dtype, const void* x, int64_t inner, int64_t red,
                     int64_t outer, void* out);
/* The same reduction for n arrays of ONE layout, handed over together so that the library can run them as one launch
 * (up to 32 arrays per launch for the row sums of a tall matrix with a short reduced extent: the bias gradients
 * `unbroadcast(b, dy)` = sum(dy, dims=2) of the four LSTM gates over many timesteps, src/unbroadcast.jl:17-24;
 * other layouts run one by one).  Values are those of n ag_sum_dim calls, bit for bit. */
ag_status ag_sum_dim_batch(int32_t map, int32_t dtype, int32_t n, const void* const* xs, int64_t inner, int64_t red,
                           int64_t outer, void* const* outs);

/* out[i,o] = op over r of x[i,r,o], op: 0 = maximum, 1 = minimum, 2 = prod (forward of the rules at
 * src/base.jl:202-206,221; their gradients `dy.*(y.==x)` / `dy.*(y./x)` go through ag_binary_bwd / ag_binary_fwd). */
ag_status ag_reduce_dim(int32_t op, int32_t dtype, const void* x, int64_t inner, int64_t red, int64_t outer,
                        void* out);

/* Gradient clipping on the device (examples/charlm.jl:116-118,148-152: gnorm = sqrt(sum over params of sum(abs2, g));
 * if gnorm > gclip the step is scaled by gclip/gnorm) without a host round trip, so a clipped step still replays as
 * one CUDA graph.  ag_multi_sumabs2: out_scalar = sum_k sum(abs2, x_k) for up to 16 tensors, one launch + one combine
 * (deterministic, double accumulation).  ag_multi_axpy_scaled: y_k += alpha * (*scale_dev) * x_k, the update with the
 * clipping factor min(1, gclip/gnorm) read from device memory. */
ag_status ag_multi_sumabs2(int32_t dtype, int32_t nseg, void* const* x, const int64_t* sizes, void* out_scalar);
ag_status ag_multi_axpy_scaled(int32_t dtype, int32_t nseg, double alpha, const void* scale_dev, void* const* x,
                               void* const* y, const int64_t* sizes);

/* ---- accumulation / update ---------------------------------------------------------------------- */
/* out = a + b (out may alias a or b).  addto!(a::AbstractArray,b) = a+b, src/addto.jl:23. */
ag_status ag_add(int32_t dtype, const void* a, const void* b, void* out, int64_t n);
/* y += alpha*x.  LinearAlgebra.axpy! in the SGD loop (README.md:70-77; src/core.jl:55). */
ag_status ag_axpy(int32_t dtype, double alpha, const void* x, void* y, int64_t n);
/* y[k] += alpha*x[k] for up to 16 tensors in one launch: the loop `for w in params(f); axpy!(-lr, grad(J,w), w); end`
 * (README.md:73-76). */
ag_status ag_multi_axpy(int32_t dtype, int32_t nseg, double alpha, void* const* x, void* const* y,
                        const int64_t* sizes);
/* dst[k] = src[k] for up to 16 tensors in one launch: packs the Param gradients into the flat allreduce bucket
 * (cat1d order, src/cat.jl:150-182). */
ag_status ag_multi_copy(int32_t dtype, int32_t nseg, void* const* src, void* const* dst, const int64_t* sizes);
/* x *= alpha (lmul!, gradient clipping examples/charlm.jl:148-149). */
ag_status ag_scale(int32_t dtype, double alpha, void* x, int64_t n);

/* ---- matrix product ------------------------------------------------------------------------------ */
/* C = alpha*op(A)*op(B) + beta*C, column-major BLAS convention, op = N (0) or T (1).
 * Forward `*` (src/core.jl:66 -> LinearAlgebra) and its rules dy*x2' (NT) and x1'*dy (TN),
 * src/base.jl:26 with the lazy adjoint of src/linearalgebra.jl:16.  Float32 uses tcgen05
 * (3xTF32 error-compensated, fp32 TMEM accumulation) where the shape qualifies and an FFMA
 * kernel otherwise; Float64 uses DFMA. */
ag_status ag_gemm(int32_t dtype, int32_t transA, int32_t transB, int64_t M, int64_t N, int64_t K,
                  double alpha, const void* A, int64_t lda, const void* B, int64_t ldb, double beta,
                  void* C, int64_t ldc);
/* The product followed by the `.+ b` and the activation that every model of the reference applies to it
 * (examples/mnist.jl:32 `max.(0, w*x .+ b)`, examples/charlm.jl:74-77 `sigm.(W*x .+ b)`, `tanh.(W*x .+ b)`), i.e. the tape
 * nodes `*` (src/base.jl:26), broadcast `+` (src/base.jl:21) and the elementwise map (src/math.jl:54,74) evaluated by
 * ONE launch: z = op(A)*op(B) .+ bias (bias: M values or NULL), a = act.(z), act an AG_U_* id out of
 * {AG_U_IDENTITY, AG_U_RELU (max.(0,z)), AG_U_TANH, AG_U_SIGM}.  C2 == NULL: C receives a.  C2 != NULL: C receives z and
 * C2 receives a (the tape keeps both when a gradient rule reads both, e.g. relu's mask y .== x).  Each step is rounded
 * on its own, so the values are bit-identical to ag_gemm + ag_binary_fwd + ag_unary_fwd.  Float32 shapes the tcgen05
 * engine takes run fused (bias and activation in its epilogue); anything else runs exactly those three kernels. */
ag_status ag_gemm_epilogue(int32_t dtype, int32_t transA, int32_t transB, int64_t M, int64_t N, int64_t K,
                           const void* A, int64_t lda, const void* B, int64_t ldb, const void* bias, int32_t act,
                           void* C, void* C2, int64_t ldc);
/* *fused_out = 1 when ag_gemm_epilogue would run this product as ONE launch (tcgen05 engine), 0 when it would run the
 * separate kernels.  The host defers `w*x .+ b` into the product only in the first case; otherwise the sum stays part of
 * the fused elementwise chain that follows it (ag_ewise_fused). */
ag_status ag_gemm_epilogue_query(int32_t dtype, int32_t transA, int32_t transB, int64_t M, int64_t N, int64_t K,
                                 const void* A, int64_t lda, const void* B, int64_t ldb, int32_t* fused_out);
/* Several independent products (each with the optional bias / activation epilogue of ag_gemm_epilogue) handed over
 * together, so that the library can run them as ONE launch: the four gate products of an LSTM step -- the reference's
 * own comment, examples/charlm.jl:69-73: "we can probably combine these four operations into one" -- their four
 * gradient products, and the weight-gradient products of several timesteps.  Products the tcgen05 engine takes are
 * grouped by class (operand majors, orientation), up to 8 per launch; the rest run one by one.  Values are those of n
 * separate ag_gemm_epilogue calls up to the summation order along K (a group that fills the SMs is not split along K
 * where a lone product would be); deterministic for a given group. */
typedef struct {
  int32_t transA, transB;
  int64_t M, N, K;
  const void* A;
  int64_t lda;
  const void* B;
  int64_t ldb;
  const void* bias; /* M values or NULL */
  int32_t act;      /* AG_U_IDENTITY, AG_U_RELU, AG_U_TANH or AG_U_SIGM */
  void* C;
  void* C2;         /* NULL, or the activation goes here and C receives the pre-activation sum */
  int64_t ldc;
} ag_gemm_desc;
ag_status ag_gemm_group(int32_t dtype, int32_t n, const ag_gemm_desc* descs);
/* Select the Float32 GEMM engine: 0 = auto, 1 = force FFMA, 2 = force tcgen05 (error if the
 * shape does not qualify).  Test hook. */
ag_status ag_gemm_set_engine(int32_t engine);
/* Name of the engine the last ag_gemm call used ("ffma", "dfma", "tcgen05", "tcgen05+epilogue", "skinny", ...). */
ag_status ag_gemm_last_engine(char* buf, int64_t buflen);

/* ---- indexing: getindex / Sparse / addtoindex! ------------------------------------------------- */
/* Minibatches in the dataset's byte format (examples/mnist.jl:77-88): the host builds `a ./ 255f0` (xshape, :80)
 * and the one-hot label matrix (yshape, :81) and every step would move 4 bytes per pixel; here the bytes cross
 * PCIe and the conversions run on the device with identical results (IEEE division; exact 0/1).
 * ag_ubyte_to_float: dst[i] = T(src[i]) / divisor for n bytes on the device.
 * ag_onehot_labels: dst is (nclass, ncols) column-major; column j has a 1 in row (labels[j] == 0 ? zero_to :
 * labels[j]) - 1 (1-based classes, `a[a.==0] = 10`), 0 elsewhere; a label outside 1..nclass raises the sticky
 * bounds flag (reported by the next ag_sync, like an index out of range). */
ag_status ag_ubyte_to_float(int32_t dtype, const void* src, void* dst, int64_t n, double divisor);
ag_status ag_onehot_labels(int32_t dtype, const void* labels, int64_t ncols, int32_t nclass, int32_t zero_to,
                           void* dst);

/* out[j] = x[idx[j]], j < n; idx are 0-based flat offsets on device.  getindex forward,
 * src/getindex.jl:32. */
ag_status ag_gather(int32_t dtype, const void* x, const int64_t* idx, void* out, int64_t n);
/* Blocked gather: out[:, j] = x[:, col[j]] for (rows x ncols_out); the W[:, ids] embedding lookup. */
ag_status ag_gather_cols(int32_t dtype, const void* x, int64_t rows, const int64_t* col,
                         void* out, int64_t ncols_out);
/* dest[idx[j]] += vals[j] for j < n with repeated indices summed (addtoindex!,
 * src/addto.jl:107-129; addto!(dense, Sparse), src/addto.jl:91-98; full(), src/sparse.jl:57).
 * Deterministic: stable radix sort by destination + segmented reduce in source order, one
 * read-modify-write per touched destination.  block > 1 treats idx as block-granular: value j
 * covers dest[idx[j]*block .. +block) (column indices of an (rows x cols) container with
 * block = rows). */
ag_status ag_scatter_add(int32_t dtype, void* dest, int64_t dest_len, const int64_t* idx,
                         const void* vals, int64_t n, int64_t block);

/* ---- cat / uncat ---------------------------------------------------------------------------------- */
/* Copy a (inner, len, outer) slab: dst[i, dst_off + r, o] = src[i, src_off + r, o] where src has
 * middle extent src_red and dst has dst_red.  cat forward / uncat gradient slices,
 * src/cat.jl:27-31,46-67; also getindex/ungetindex with range indices. */
ag_status ag_slab_copy(int32_t dtype, const void* src, int64_t src_red, int64_t src_off, void* dst,
                       int64_t dst_red, int64_t dst_off, int64_t inner, int64_t len, int64_t outer);
/* cat(xs...; dims) of up to 8 arrays in ONE launch (src/cat.jl:27-31; `vcat(input, hidden)` of examples/charlm.jl:68):
 * dst viewed as inner x sum(len) x outer (column-major), source j as inner x len[j] x outer.  128-bit accesses when
 * every inner*len[j] is a multiple of the vector width and all bases are 16-byte aligned. */
ag_status ag_cat(int32_t dtype, int32_t n, const void* const* src, const int64_t* len, void* dst, int64_t inner,
                 int64_t outer);
/* permutedims(x, perm) for ndims <= 4 (src/base.jl:217-218); perm is 0-based: out dim d = x dim perm[d]. */
ag_status ag_permute(int32_t dtype, const void* x, int32_t ndims, const int64_t* shape, const int32_t* perm,
                     void* out);
/* permutedims(x, (2,1)) materialised: out (cols x rows) = transpose of x (rows x cols). */
ag_status ag_transpose(int32_t dtype, const void* x, int64_t rows, int64_t cols, void* out);

/* ---- C-side tape ------------------------------------------------------------------ */
/* The reference records a Node per Result on a host tape (src/core.jl:110-129) and sweeps it in reverse, one `back` and
 * one `addto!` per tracked parent (src/core.jl:153-170).  For tapes made only of device primitives the same bookkeeping
 * lives behind the ABI: one call per recorded operation (which also runs its forward kernel) and ONE call for the whole
 * backward sweep.  Values are small integer ids local to the tape.  Node order, the order of the `back` calls and
 * the out-of-place accumulation are the reference's, so gradients are bit-identical to the host-recorded tape running
 * the same (unfused) kernels.  First order only.  Arrays have <= 4 dims. */
enum {
  AG_T_MATMUL = 0,     /* x1*x2: dy*x2', x1'*dy                            src/base.jl:26 */
  AG_T_ADD = 1,        /* x1 .+ x2 (Julia broadcasting): unbroadcast(xi, dy)        :21 */
  AG_T_SUB = 2,        /* x1 .- x2                                                  :23 */
  AG_T_MUL = 3,        /* x1 .* x2: unbroadcast(x1, dy.*x2), unbroadcast(x2, x1.*dy) :27 */
  AG_T_DIV_SCALAR = 4, /* x ./ scalar: dy ./ scalar                                 :33 */
  AG_T_MUL_SCALAR = 5, /* scalar .* x                                               :27 */
  AG_T_UNARY = 6,      /* f.(x), f = AG_U_* id in `arg`: dy .* g(x,y)     src/math.jl:9-74 */
  AG_T_SUM = 7,        /* sum(x): dy .* one.(x)                           src/base.jl:245 */
  AG_T_SUMABS2 = 8,    /* sum(abs2, x): dy .* 2x                                    :247 */
  AG_T_SUM_DIM = 9     /* sum(x; dims = arg + 1), keeps the dim with extent 1       :245 */
};
ag_status ag_tape_new(uint64_t* tape_out);
/* Frees the tape, the values it computed and every gradient it produced (leaf arrays belong to the caller). */
ag_status ag_tape_free(uint64_t tape);
/* Register an array the caller owns: a Param (tracked = 1, src/core.jl:201) or a constant (tracked = 0). */
ag_status ag_tape_leaf(uint64_t tape, int32_t dtype, void* ptr, int32_t ndims, const int64_t* shape, int32_t tracked,
                       int32_t* id_out);
/* forw + record! (src/core.jl:64-79,110-129): run primitive `op` on the values in_ids[0..nin) and append a node. */
ag_status ag_tape_record(uint64_t tape, int32_t op, int32_t arg, int32_t nin, const int32_t* in_ids, double scalar,
                         int32_t* id_out);
/* The backward sweep of `differentiate` (src/core.jl:153-170) from the scalar value result_id. */
ag_status ag_tape_backward(uint64_t tape, int32_t result_id);
/* grad(tape, x) (src/core.jl:215-216): device pointer of the accumulated gradient, NULL when none reached it. */
ag_status ag_tape_grad(uint64_t tape, int32_t id, void** ptr_out);
/* value(x): pointer, number of dims and shape (4 entries) of a value of the tape. */
ag_status ag_tape_value(uint64_t tape, int32_t id, void** ptr_out, int32_t* ndims_out, int64_t* shape4_out);

/* ---- data-parallel exchange (no reference counterpart: BASELINE.json north_star) --------------- */
/* NCCL communicator over the ranks of one node; id is the 128-byte ncclUniqueId produced on rank
 * 0 by ag_comm_unique_id and distributed by the host launcher. */
ag_status ag_comm_unique_id(uint8_t* id128);
ag_status ag_comm_init(int32_t nranks, int32_t rank, const uint8_t* id128);
ag_status ag_comm_destroy(void);
/* In-place sum-allreduce of a flat gradient bucket on the library stream. */
ag_status ag_allreduce_sum(int32_t dtype, void* buf, int64_t n);
/* One-shot peer-memory allreduce for latency-bound buckets (single node, NVLink/NVSwitch): each rank
 * exports a symmetric buffer (CUDA IPC handle, 64 bytes), the launcher gathers the handles of all ranks and
 * hands them to ag_p2p_connect; ag_p2p_allreduce_sum then sums `buf` over all ranks in place with ONE kernel
 * (publish, flag every peer, sum all peers' buckets in rank order: bit-identical on every rank).
 * n*sizeof(dtype) must not exceed cap_bytes.  Float32 buckets up to 1 MB use a flag-in-data (LL) push protocol:
 * no fence and no flag round trip on the critical path.  Graph-capturable. */
ag_status ag_p2p_export(int64_t cap_bytes, int32_t nranks, uint8_t* handle64);
ag_status ag_p2p_connect(int32_t nranks, int32_t rank, const uint8_t* handles /* nranks x 64 bytes */);
ag_status ag_p2p_allreduce_sum(int32_t dtype, void* buf, int64_t n);
/* Multi-tensor form (<= 16 segments, no bucket copies): src[k] (n = sizes[k]) is replaced by its sum over all
 * ranks; when dst is non-NULL, dst[k] += alpha * sum is applied in the same kernel -- the SGD update
 * `axpy!(-lr, grad(J,w), w)` (README.md:70-77) fused behind the collective. */
ag_status ag_p2p_allreduce_multi(int32_t dtype, int32_t nseg, void* const* src, void* const* dst,
                                 const int64_t* sizes, double alpha);
ag_status ag_p2p_destroy(void);

#ifdef __cplusplus
}
#endif
#endif /* AGB200_H */
