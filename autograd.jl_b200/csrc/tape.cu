// tape.cu -- a C-side tape: record array primitives, run their forward kernels, sweep backward.
//
// The reference keeps its tape on the host (Julia): `forw` runs a primitive and `record!`s a Node per Result
// (src/core.jl:64-129); `differentiate` walks the nodes in reverse record order, calls `back` per tracked parent and
// accumulates with `addto!` (src/core.jl:153-170).  For tapes made only of device primitives the same bookkeeping can
// live behind the C ABI: the host then pays one foreign call per recorded operation and ONE for the whole backward
// sweep, instead of re-entering its own dispatch for every rule and every accumulation.  The rules are
// those of the reference, each as the fused kernel the host path uses outside of recording:
//   *            dy*x2', x1'*dy                          src/base.jl:26 (lazy adjoint: src/linearalgebra.jl:16)
//   .+ .-        unbroadcast(xi, +-dy)                   src/base.jl:21,23
//   .*           unbroadcast(x1, dy.*x2), (x2, x1.*dy)   src/base.jl:27
//   ./ Number    dy./c                                   src/base.jl:33
//   f.(x)        dy.*g(x,y) for the unary table          src/math.jl:9-74, src/base.jl:73-74
//   sum, sum(;dims), sum(abs2,.)   dy.*one.(x), dy.*2x   src/base.jl:245-247
//   unbroadcast  sum over the broadcast dims + reshape   src/unbroadcast.jl:9-26
//   addto!       a+b out of place                        src/addto.jl:23
// Node order, the order of `back` calls (reverse record order; parents in argument order) and the out-of-place
// accumulation follow the reference, so gradients are bit-identical to the host-recorded tape running the same
// (unfused) kernels.  First-order only: a tape cannot be recorded on another tape.
#include "common.cuh"

#include <mutex>
#include <unordered_map>
#include <vector>

namespace agb {

struct TVal {
  void* ptr = nullptr;
  int nd = 0;
  int64_t shape[4] = {1, 1, 1, 1};
  int dtype = AG_F32;
  bool owned = false;      // allocated by the tape (freed with it)
  bool tracked = false;    // a Param, or computed from one
  void* grad = nullptr;    // accumulated outgrad (owned)
  int64_t size() const {
    int64_t n = 1;
    for (int d = 0; d < nd; ++d) n *= shape[d];
    return n;
  }
  size_t bytes() const { return (size_t)size() * (dtype == AG_F32 ? 4 : 8); }
};

struct TNode {
  int op = 0, arg = 0;     // AG_T_*; arg: unary op id / reduced dim
  int in[2] = {-1, -1};
  int out = -1;
  double scalar = 0;
};

struct CTape {
  std::vector<TVal> vals;
  std::vector<TNode> nodes;
  std::vector<void*> scratch;    // temporaries of the backward sweep, freed with the tape
};

static std::mutex g_tmu;
static std::unordered_map<uint64_t, CTape*> g_tapes;
static uint64_t g_next_tape = 1;

static CTape* find_tape(uint64_t h) {
  std::lock_guard<std::mutex> lk(g_tmu);
  auto it = g_tapes.find(h);
  return it == g_tapes.end() ? nullptr : it->second;
}

static ag_status talloc(CTape* t, const TVal& like, void** out, bool keep_in_scratch) {
  ag_status st = ag_alloc((int64_t)(like.bytes() ? like.bytes() : 1), out);
  if (st == AG_OK && keep_in_scratch) t->scratch.push_back(*out);
  return st;
}

// Julia broadcasting: align leading dims, pad trailing 1s.  Fills the result shape and both operands' element strides.
static ag_status bcast(const TVal& a, const TVal& b, TVal* out, int64_t* sa, int64_t* sb) {
  const int nd = a.nd > b.nd ? a.nd : b.nd;
  out->nd = nd;
  int64_t acca = 1, accb = 1;
  for (int d = 0; d < nd; ++d) {
    const int64_t ea = d < a.nd ? a.shape[d] : 1, eb = d < b.nd ? b.shape[d] : 1;
    AG_CHECK(ea == eb || ea == 1 || eb == 1, AG_EINVAL,
             "ag_tape_record: arrays could not be broadcast to a common size (dim %d: %lld vs %lld)", d + 1,
             (long long)ea, (long long)eb);
    const int64_t e = ea == 1 ? eb : ea;
    out->shape[d] = e;
    sa[d] = (ea == 1 && e != 1) ? 0 : acca;
    sb[d] = (eb == 1 && e != 1) ? 0 : accb;
    acca *= ea;
    accb *= eb;
  }
  for (int d = nd; d < 4; ++d) out->shape[d] = 1;
  return AG_OK;
}

static void own_strides(const TVal& v, int nd, int64_t* s) {
  int64_t acc = 1;
  for (int d = 0; d < nd; ++d) {
    s[d] = acc;
    acc *= d < v.nd ? v.shape[d] : 1;
  }
}

// unbroadcast(x, dx): sum dx over the dims where x has extent 1 (src/unbroadcast.jl:17-24); returns dx itself when
// the shapes agree (the reference returns dx: aliasing is safe because accumulation is out of place).
static ag_status unbroadcast(CTape* t, const TVal& x, const TVal& dxv, void* dx, void** out) {
  if (x.size() == dxv.size()) { *out = dx; return AG_OK; }
  TVal cur = dxv;
  void* cp = dx;
  int d = 0;
  while (d < cur.nd) {
    const int64_t ex = d < x.nd ? x.shape[d] : 1;
    if (!(ex == 1 && cur.shape[d] > 1)) { ++d; continue; }
    int hi = d;
    while (hi + 1 < cur.nd && (hi + 1 < x.nd ? x.shape[hi + 1] : 1) == 1 && cur.shape[hi + 1] > 1) ++hi;
    int64_t inner = 1, red = 1, outer = 1;
    for (int k = 0; k < d; ++k) inner *= cur.shape[k];
    for (int k = d; k <= hi; ++k) red *= cur.shape[k];
    for (int k = hi + 1; k < cur.nd; ++k) outer *= cur.shape[k];
    TVal nv = cur;
    for (int k = d; k <= hi; ++k) nv.shape[k] = 1;
    void* np = nullptr;
    ag_status st = talloc(t, nv, &np, true);
    if (st != AG_OK) return st;
    st = ag_sum_dim(AG_R_ID, cur.dtype, cp, inner, red, outer, np);
    if (st != AG_OK) return st;
    cur = nv;
    cp = np;
    d = hi + 1;
  }
  AG_CHECK(cur.size() == x.size(), AG_EINVAL, "unbroadcast: DimensionMismatch");
  *out = cp;
  return AG_OK;
}

// p.outgrad = addto!(p.outgrad, g): adopt g when there is none, else a+b out of place (src/addto.jl:23,27-31)
static ag_status accumulate(CTape* t, TVal& p, void* g) {
  if (!p.grad) { p.grad = g; return AG_OK; }
  void* sum = nullptr;
  ag_status st = talloc(t, p, &sum, true);
  if (st != AG_OK) return st;
  st = ag_add(p.dtype, p.grad, g, sum, p.size());
  p.grad = sum;
  return st;
}

}  // namespace agb

using namespace agb;

extern "C" {

ag_status ag_tape_new(uint64_t* tape_out) {
  AG_REQUIRE_INIT();
  AG_CHECK(tape_out, AG_EINVAL, "ag_tape_new: null");
  std::lock_guard<std::mutex> lk(g_tmu);
  *tape_out = g_next_tape++;
  g_tapes[*tape_out] = new CTape();
  return AG_OK;
}

ag_status ag_tape_free(uint64_t tape) {
  CTape* t = nullptr;
  {
    std::lock_guard<std::mutex> lk(g_tmu);
    auto it = g_tapes.find(tape);
    if (it == g_tapes.end()) return AG_OK;
    t = it->second;
    g_tapes.erase(it);
  }
  for (auto& v : t->vals)
    if (v.owned && v.ptr) ag_free(v.ptr);
  for (void* p : t->scratch) ag_free(p);     // includes every accumulated gradient
  delete t;
  return AG_OK;
}

ag_status ag_tape_leaf(uint64_t tape, int32_t dtype, void* ptr, int32_t ndims, const int64_t* shape, int32_t tracked,
                       int32_t* id_out) {
  AG_REQUIRE_INIT();
  CTape* t = find_tape(tape);
  AG_CHECK(t, AG_EINVAL, "ag_tape_leaf: unknown tape");
  AG_CHECK(dtype == AG_F32 || dtype == AG_F64, AG_EUNSUPPORTED, "ag_tape_leaf: dtype %d", dtype);
  AG_CHECK(ptr && ndims >= 0 && ndims <= 4 && (ndims == 0 || shape) && id_out, AG_EINVAL, "ag_tape_leaf: bad arguments");
  TVal v;
  v.ptr = ptr;
  v.nd = ndims;
  for (int d = 0; d < ndims; ++d) v.shape[d] = shape[d];
  v.dtype = dtype;
  v.tracked = tracked != 0;
  t->vals.push_back(v);
  *id_out = (int32_t)t->vals.size() - 1;
  return AG_OK;
}

ag_status ag_tape_record(uint64_t tape, int32_t op, int32_t arg, int32_t nin, const int32_t* in_ids, double scalar,
                         int32_t* id_out) {
  AG_REQUIRE_INIT();
  CTape* t = find_tape(tape);
  AG_CHECK(t, AG_EINVAL, "ag_tape_record: unknown tape");
  AG_CHECK(in_ids && id_out && nin >= 1 && nin <= 2, AG_EINVAL, "ag_tape_record: bad arguments");
  for (int i = 0; i < nin; ++i)
    AG_CHECK(in_ids[i] >= 0 && in_ids[i] < (int)t->vals.size(), AG_EINVAL, "ag_tape_record: value id %d", in_ids[i]);
  const TVal a = t->vals[in_ids[0]];
  const TVal b = nin > 1 ? t->vals[in_ids[1]] : TVal();
  if (nin > 1) AG_CHECK(a.dtype == b.dtype, AG_EUNSUPPORTED, "ag_tape_record: mixed eltypes");
  TVal y;
  y.dtype = a.dtype;
  y.owned = true;
  y.tracked = a.tracked || (nin > 1 && b.tracked);
  ag_status st = AG_OK;
  int64_t sa[4], sb[4];
  switch (op) {
    case AG_T_MATMUL: {
      AG_CHECK(nin == 2 && a.nd == 2 && b.nd == 2 && a.shape[1] == b.shape[0], AG_EINVAL,
               "ag_tape_record(*): A has dimensions (%lld,%lld) but B has dimensions (%lld,%lld)", (long long)a.shape[0],
               (long long)a.shape[1], (long long)b.shape[0], (long long)b.shape[1]);
      y.nd = 2;
      y.shape[0] = a.shape[0];
      y.shape[1] = b.shape[1];
      if ((st = talloc(t, y, &y.ptr, false)) != AG_OK) return st;
      st = ag_gemm(a.dtype, 0, 0, a.shape[0], b.shape[1], a.shape[1], 1.0, a.ptr, a.shape[0], b.ptr, b.shape[0], 0.0, y.ptr,
                   a.shape[0]);
      break;
    }
    case AG_T_ADD: case AG_T_SUB: case AG_T_MUL: {
      AG_CHECK(nin == 2, AG_EINVAL, "ag_tape_record: binary op needs two inputs");
      if ((st = bcast(a, b, &y, sa, sb)) != AG_OK) return st;
      if ((st = talloc(t, y, &y.ptr, false)) != AG_OK) return st;
      const int bop = op == AG_T_ADD ? AG_B_ADD : op == AG_T_SUB ? AG_B_SUB : AG_B_MUL;
      st = ag_binary_fwd(bop, a.dtype, a.ptr, 0.0, sa, b.ptr, 0.0, sb, y.ptr, y.nd ? y.nd : 1, y.shape);
      break;
    }
    case AG_T_DIV_SCALAR: case AG_T_MUL_SCALAR: {
      y.nd = a.nd;
      for (int d = 0; d < 4; ++d) y.shape[d] = a.shape[d];
      if ((st = talloc(t, y, &y.ptr, false)) != AG_OK) return st;
      own_strides(a, a.nd ? a.nd : 1, sa);
      for (int d = 0; d < 4; ++d) sb[d] = 0;
      st = ag_binary_fwd(op == AG_T_DIV_SCALAR ? AG_B_DIV : AG_B_MUL, a.dtype, a.ptr, 0.0, sa, nullptr, scalar, sb, y.ptr,
                         a.nd ? a.nd : 1, a.shape);
      break;
    }
    case AG_T_UNARY: {
      AG_CHECK(arg >= 0 && arg < AG_U_COUNT, AG_EINVAL, "ag_tape_record: unary op %d", arg);
      y.nd = a.nd;
      for (int d = 0; d < 4; ++d) y.shape[d] = a.shape[d];
      if ((st = talloc(t, y, &y.ptr, false)) != AG_OK) return st;
      st = ag_unary_fwd(arg, a.dtype, a.ptr, y.ptr, a.size());
      break;
    }
    case AG_T_SUM: case AG_T_SUMABS2: {
      y.nd = 0;
      if ((st = talloc(t, y, &y.ptr, false)) != AG_OK) return st;
      st = ag_sum_all(op == AG_T_SUM ? AG_R_ID : AG_R_ABS2, a.dtype, a.ptr, a.size(), y.ptr);
      break;
    }
    case AG_T_SUM_DIM: {
      AG_CHECK(arg >= 0 && arg < a.nd, AG_EINVAL, "ag_tape_record(sum; dims): dim %d of %d", arg + 1, a.nd);
      y.nd = a.nd;
      for (int d = 0; d < 4; ++d) y.shape[d] = a.shape[d];
      y.shape[arg] = 1;
      int64_t inner = 1, outer = 1;
      for (int d = 0; d < arg; ++d) inner *= a.shape[d];
      for (int d = arg + 1; d < a.nd; ++d) outer *= a.shape[d];
      if ((st = talloc(t, y, &y.ptr, false)) != AG_OK) return st;
      st = ag_sum_dim(AG_R_ID, a.dtype, a.ptr, inner, a.shape[arg], outer, y.ptr);
      break;
    }
    default:
      AG_CHECK(false, AG_EUNSUPPORTED, "ag_tape_record: op %d", op);
  }
  if (st != AG_OK) {
    if (y.ptr) ag_free(y.ptr);
    return st;
  }
  t->vals.push_back(y);
  TNode n;
  n.op = op;
  n.arg = arg;
  n.in[0] = in_ids[0];
  n.in[1] = nin > 1 ? in_ids[1] : -1;
  n.out = (int)t->vals.size() - 1;
  n.scalar = scalar;
  t->nodes.push_back(n);
  *id_out = n.out;
  return AG_OK;
}

ag_status ag_tape_backward(uint64_t tape, int32_t result_id) {
  AG_REQUIRE_INIT();
  CTape* t = find_tape(tape);
  AG_CHECK(t, AG_EINVAL, "ag_tape_backward: unknown tape");
  AG_CHECK(result_id >= 0 && result_id < (int)t->vals.size(), AG_EINVAL, "ag_tape_backward: value id %d", result_id);
  AG_CHECK(t->vals[result_id].size() == 1, AG_EINVAL, "Only scalar valued functions supported.");   // src/core.jl:153
  for (auto& v : t->vals) v.grad = nullptr;
  ag_status st;
  {
    TVal& r = t->vals[result_id];
    if ((st = talloc(t, r, &r.grad, true)) != AG_OK) return st;
    if ((st = ag_fill(r.grad, 1, r.dtype, 1.0)) != AG_OK) return st;           // one(result)
  }
  for (int ni = (int)t->nodes.size() - 1; ni >= 0; --ni) {
    const TNode n = t->nodes[ni];
    const TVal y = t->vals[n.out];
    if (!y.grad) continue;                                                     // src/core.jl:158
    for (int i = 0; i < 2; ++i) {
      if (n.in[i] < 0 || !t->vals[n.in[i]].tracked) continue;
      const TVal x = t->vals[n.in[i]];
      const TVal other = n.in[1 - i] >= 0 ? t->vals[n.in[1 - i]] : TVal();
      void* g = nullptr;
      int64_t s0[4], s1[4], s2[4], s3[4];
      switch (n.op) {
        case AG_T_MATMUL: {
          if ((st = talloc(t, x, &g, true)) != AG_OK) return st;
          const int64_t M = t->vals[n.in[0]].shape[0], K = t->vals[n.in[0]].shape[1], N = t->vals[n.in[1]].shape[1];
          if (i == 0)      // dy*x2'  (M x N)*(N x K)
            st = ag_gemm(x.dtype, 0, 1, M, K, N, 1.0, y.grad, M, other.ptr, K, 0.0, g, M);
          else             // x1'*dy  (K x M)*(M x N)
            st = ag_gemm(x.dtype, 1, 0, K, N, M, 1.0, other.ptr, M, y.grad, M, 0.0, g, K);
          break;
        }
        case AG_T_ADD: case AG_T_SUB: {
          void* src = y.grad;
          if (n.op == AG_T_SUB && i == 1) {                                    // -dy
            if ((st = talloc(t, y, &src, true)) != AG_OK) return st;
            if ((st = ag_unary_fwd(AG_U_NEG, y.dtype, y.grad, src, y.size())) != AG_OK) return st;
          }
          st = unbroadcast(t, x, y, src, &g);
          break;
        }
        case AG_T_MUL: {
          void* full = nullptr;
          if ((st = talloc(t, y, &full, true)) != AG_OK) return st;
          TVal tmp;
          const TVal& x1 = t->vals[n.in[0]];
          const TVal& x2 = t->vals[n.in[1]];
          if ((st = bcast(x1, x2, &tmp, s1, s2)) != AG_OK) return st;
          own_strides(y, y.nd ? y.nd : 1, s0);
          for (int d = 0; d < 4; ++d) s3[d] = s0[d];
          st = ag_binary_bwd(AG_B_MUL, i + 1, y.dtype, y.grad, s0, x1.ptr, 0.0, s1, x2.ptr, 0.0, s2, y.ptr, s3, full,
                             y.nd ? y.nd : 1, y.shape);
          if (st != AG_OK) return st;
          st = unbroadcast(t, x, y, full, &g);
          break;
        }
        case AG_T_DIV_SCALAR: case AG_T_MUL_SCALAR: {
          if ((st = talloc(t, x, &g, true)) != AG_OK) return st;
          own_strides(y, y.nd ? y.nd : 1, s0);
          for (int d = 0; d < 4; ++d) s1[d] = 0;
          st = ag_binary_fwd(n.op == AG_T_DIV_SCALAR ? AG_B_DIV : AG_B_MUL, y.dtype, y.grad, 0.0, s0, nullptr, n.scalar, s1,
                             g, y.nd ? y.nd : 1, y.shape);
          break;
        }
        case AG_T_UNARY:
          if ((st = talloc(t, x, &g, true)) != AG_OK) return st;
          st = ag_unary_bwd(n.arg, x.dtype, y.grad, 0, 0.0, x.ptr, y.ptr, g, x.size());
          break;
        case AG_T_SUM:        // dy .* one.(x): the device scalar dy filled over x's shape
          if ((st = talloc(t, x, &g, true)) != AG_OK) return st;
          st = ag_unary_bwd(AG_U_IDENTITY, x.dtype, y.grad, x.size() == 1 ? 0 : 1, 0.0, nullptr, nullptr, g, x.size());
          break;
        case AG_T_SUMABS2:    // dy .* 2x
          if ((st = talloc(t, x, &g, true)) != AG_OK) return st;
          st = ag_unary_bwd(AG_U_ABS2, x.dtype, y.grad, x.size() == 1 ? 0 : 1, 0.0, x.ptr, nullptr, g, x.size());
          break;
        case AG_T_SUM_DIM: {  // dy (keep-dims) broadcast back over the reduced dim
          if ((st = talloc(t, x, &g, true)) != AG_OK) return st;
          int64_t acc = 1;
          for (int d = 0; d < x.nd; ++d) {
            s0[d] = (y.shape[d] == 1 && x.shape[d] != 1) ? 0 : acc;
            acc *= y.shape[d];
            s1[d] = 0;
          }
          st = ag_binary_fwd(AG_B_MUL, x.dtype, y.grad, 0.0, s0, nullptr, 1.0, s1, g, x.nd, x.shape);
          break;
        }
        default:
          AG_CHECK(false, AG_EUNSUPPORTED, "ag_tape_backward: op %d", n.op);
      }
      if (st != AG_OK) return st;
      if ((st = accumulate(t, t->vals[n.in[i]], g)) != AG_OK) return st;
    }
  }
  return AG_OK;
}

ag_status ag_tape_grad(uint64_t tape, int32_t id, void** ptr_out) {
  AG_REQUIRE_INIT();
  CTape* t = find_tape(tape);
  AG_CHECK(t && ptr_out, AG_EINVAL, "ag_tape_grad: unknown tape");
  AG_CHECK(id >= 0 && id < (int)t->vals.size(), AG_EINVAL, "ag_tape_grad: value id %d", id);
  *ptr_out = t->vals[id].grad;          // NULL = `nothing`: no gradient reached this value (src/core.jl:215-216)
  return AG_OK;
}

ag_status ag_tape_value(uint64_t tape, int32_t id, void** ptr_out, int32_t* ndims_out, int64_t* shape4_out) {
  AG_REQUIRE_INIT();
  CTape* t = find_tape(tape);
  AG_CHECK(t && ptr_out, AG_EINVAL, "ag_tape_value: unknown tape");
  AG_CHECK(id >= 0 && id < (int)t->vals.size(), AG_EINVAL, "ag_tape_value: value id %d", id);
  *ptr_out = t->vals[id].ptr;
  if (ndims_out) *ndims_out = t->vals[id].nd;
  if (shape4_out)
    for (int d = 0; d < 4; ++d) shape4_out[d] = t->vals[id].shape[d];
  return AG_OK;
}

}  // extern "C"
