// index.cu -- getindex / Sparse / addtoindex! / cat / axpy! on device.
//
// Reference: getindex forward (src/getindex.jl:32), ungetindex -> Sparse (src/getindex.jl:61-79),
// addto!(dense, Sparse) and full() (src/addto.jl:91-98, src/sparse.jl:57) which run the sequential
// scalar loop addtoindex! (src/addto.jl:107-129); cat/uncat (src/cat.jl:27-31,46-67);
// axpy!/lmul! (src/sparse.jl:50-52, README.md:70-77).
//
// Scatter-add is deterministic: an in-tree stable LSD radix sort (8 bits per pass; only as many passes as the number
// of destinations needs: one for the 128-column embedding of config 4) of (destination, source position) pairs,
// then a segmented reduce that sums each destination's contributions in a fixed order derived from the source order
// (warps take consecutive runs of the segment, 128-bit loads along the value block) and performs ONE
// read-modify-write per touched destination.  Index bookkeeping (which destination each value goes to) is bit-exact
// with the reference; only the floating-point association order inside a destination can differ from the
// reference's interleaved sequential order.
// Algorithmic bytes of a scatter-add: n*block*s (values) + 8n (indices) + 2*touched*block*s (destination RMW).
#include "common.cuh"
#include "ewise_ops.cuh"

namespace agb {

template <typename T>
__global__ void __launch_bounds__(256)
    k_gather(const T* __restrict__ x, const int64_t* __restrict__ idx, T* __restrict__ out, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = x[idx[i]];
}

// out[r, j] = x[r, col[j]]
template <typename T>
__global__ void __launch_bounds__(256)
    k_gather_cols(const T* __restrict__ x, int64_t rows, const int64_t* __restrict__ col,
                  T* __restrict__ out, int64_t total) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int64_t j = i / rows, r = i - j * rows;
  out[i] = x[r + col[j] * rows];
}
// the same with 128-bit accesses (rows a multiple of the vector width, both arrays 16-byte aligned): a thread moves
// one 16-byte piece of a column
template <typename T>
__global__ void __launch_bounds__(256)
    k_gather_cols_v(const T* __restrict__ x, int64_t rows, const int64_t* __restrict__ col,
                    T* __restrict__ out, int64_t total_vec) {
  constexpr int N = 16 / sizeof(T);
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total_vec) return;
  const int64_t rv = rows / N;
  int64_t j = i / rv, r = (i - j * rv) * N;
  *reinterpret_cast<uint4*>(out + j * rows + r) = __ldg(reinterpret_cast<const uint4*>(x + r + col[j] * rows));
}

// ---- stable LSD radix sort of (key, source position), 8 bits per pass ------------------------------------
// A block owns a contiguous tile of the input (tile = a multiple of 256 entries).  Pass = three launches:
// k_radix_hist (per-block digit counts, stored digit-major), k_radix_scan (exclusive scan of that table: where each
// (digit, block) run starts), k_radix_scatter (stable ranks inside the tile -> final positions).  The first pass reads
// the caller's int64 indices (position = entry number), checks their range and narrows them to uint32; entries out
// of range get the key `nkeys` (sorted behind every valid key, never written) and raise the sticky flag.
constexpr int kRadixBits = 8, kRadixBins = 1 << kRadixBits;

template <bool FIRST>
__device__ __forceinline__ uint32_t radix_key(const void* keys, int64_t i, int64_t nkeys, int32_t* flag) {
  if (FIRST) {
    const int64_t k = ((const int64_t*)keys)[i];
    if (k < 0 || k >= nkeys) {
      *flag = 1;
      return (uint32_t)nkeys;
    }
    return (uint32_t)k;
  }
  return ((const uint32_t*)keys)[i];
}

template <bool FIRST>
__global__ void __launch_bounds__(256)
    k_radix_hist(const void* __restrict__ keys, int64_t n, int64_t tile, int shift, int64_t nkeys,
                 int32_t* __restrict__ flag, uint32_t* __restrict__ hist) {
  __shared__ uint32_t h[kRadixBins];
  h[threadIdx.x] = 0;
  __syncthreads();
  const int64_t lo = (int64_t)blockIdx.x * tile, hi = min(n, lo + tile);
  for (int64_t i = lo + threadIdx.x; i < hi; i += 256)
    atomicAdd(&h[(radix_key<FIRST>(keys, i, nkeys, flag) >> shift) & (kRadixBins - 1)], 1u);
  __syncthreads();
  hist[(int64_t)threadIdx.x * gridDim.x + blockIdx.x] = h[threadIdx.x];
}

// exclusive scan of `m` counters in place (one block): every thread sums a contiguous run, one block-wide scan of
// the 1024 run totals, then every thread rewrites its run
__global__ void __launch_bounds__(1024) k_radix_scan(uint32_t* __restrict__ hist, int64_t m) {
  __shared__ uint32_t wsum[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t per = (m + 1023) / 1024;
  const int64_t lo = min(m, (int64_t)threadIdx.x * per), hi = min(m, lo + per);
  uint32_t tot = 0;
  for (int64_t i = lo; i < hi; ++i) tot += hist[i];
  uint32_t x = tot;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) wsum[w] = x;
  __syncthreads();
  if (w == 0) {
    uint32_t t = wsum[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, t, o);
      if (lane >= o) t += y;
    }
    wsum[lane] = t;                       // inclusive sums of the warps
  }
  __syncthreads();
  uint32_t run = (w ? wsum[w - 1] : 0u) + x - tot;
  for (int64_t i = lo; i < hi; ++i) {
    const uint32_t v = hist[i];
    hist[i] = run;
    run += v;
  }
}

template <bool FIRST>
__global__ void __launch_bounds__(256)
    k_radix_scatter(const void* __restrict__ keys_in, const int32_t* __restrict__ pos_in, int64_t n, int64_t tile,
                    int shift, int64_t nkeys, const uint32_t* __restrict__ offs, uint32_t* __restrict__ keys_out,
                    int32_t* __restrict__ pos_out) {
  __shared__ uint32_t run[kRadixBins];          // next free position of every digit for this block
  __shared__ uint32_t cnt[8][kRadixBins];       // this round's entries per warp and digit
  int32_t dummy = 0;
  run[threadIdx.x] = offs[(int64_t)threadIdx.x * gridDim.x + blockIdx.x];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t lo = (int64_t)blockIdx.x * tile, hi = min(n, lo + tile);
  for (int64_t base = lo; base < hi; base += 256) {
#pragma unroll
    for (int q = 0; q < 8; ++q) cnt[q][threadIdx.x] = 0;
    __syncthreads();
    const int64_t i = base + threadIdx.x;
    const bool on = i < hi;
    uint32_t key = 0, d = 0;
    if (on) {
      key = radix_key<FIRST>(keys_in, i, nkeys, &dummy);
      d = (key >> shift) & (kRadixBins - 1);
    }
    // entries of the warp with the same digit, in lane (= source) order
    const uint32_t act = __ballot_sync(0xffffffffu, on);
    uint32_t peers = 0;
    if (on) peers = __match_any_sync(act, d);
    const uint32_t below = peers & ((1u << lane) - 1u);
    if (on && below == 0) cnt[w][d] = __popc(peers);
    __syncthreads();
    if (on) {
      uint32_t at = run[d] + __popc(below);
      for (int q = 0; q < w; ++q) at += cnt[q][d];
      keys_out[at] = key;
      pos_out[at] = FIRST ? (int32_t)i : pos_in[i];
    }
    __syncthreads();
    uint32_t tot = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) tot += cnt[q][threadIdx.x];
    run[threadIdx.x] += tot;
    __syncthreads();
  }
}

// Narrow value blocks (fewer than 32 vector lanes): a group of `lanes` threads per sorted entry, lanes along the value
// block; the group of a segment head walks its segment in source order, one read-modify-write per destination.
template <typename T>
__global__ void __launch_bounds__(256)
    k_segreduce_narrow(const uint32_t* __restrict__ keys, const int32_t* __restrict__ pos, const T* __restrict__ vals,
                       int64_t n, int64_t block, int64_t nkeys, int lanes, T* __restrict__ dest) {
  const int per_blk = 256 / lanes;
  const int lane = threadIdx.x % lanes, slot = threadIdx.x / lanes;
  const int64_t j = (int64_t)blockIdx.x * per_blk + slot;
  if (j >= n || lane >= block) return;
  const uint32_t key = keys[j];
  if ((int64_t)key >= nkeys) return;          // out of range: flagged by the first sort pass, never written
  if (j > 0 && keys[j - 1] == key) return;    // not a head
  int64_t e = j + 1;
  while (e < n && keys[e] == key) ++e;
  for (int64_t b = lane; b < block; b += lanes) {
    T s = T(0);
    for (int64_t m = j; m < e; ++m) s += vals[(int64_t)pos[m] * block + b];
    dest[(int64_t)key * block + b] += s;
  }
}

// First position of the sorted keys that is not below `key`, found by the whole block: 256 probes per round narrow
// the range 256-fold (two or three dependent loads instead of the ~17 of a per-thread binary search).  All 256
// threads of the block call it; all return the same value.
__device__ __forceinline__ int64_t block_lower_bound(const uint32_t* __restrict__ keys, int64_t n, uint32_t key) {
  __shared__ int64_t s_lo, s_hi;      // the answer lies in [s_lo, s_hi]
  if (threadIdx.x == 0) { s_lo = 0; s_hi = n; }
  __syncthreads();
  while (s_lo < s_hi) {
    const int64_t lo = s_lo, hi = s_hi;
    const int64_t stride = (hi - lo + 255) / 256;
    const int64_t p = lo + (int64_t)threadIdx.x * stride;
    const int below = (p < hi && keys[p] < key) ? 1 : 0;
    const int cnt = __syncthreads_count(below);          // the probes below the key form a prefix (sorted keys)
    if (threadIdx.x == 0) {
      if (cnt == 0) {
        s_hi = lo;                                       // keys[lo] is not below: the answer is lo
      } else if (stride == 1) {
        s_lo = s_hi = lo + cnt;                          // every position was probed
      } else {
        s_lo = lo + (int64_t)(cnt - 1) * stride + 1;     // behind the last probe below the key ...
        const int64_t q = lo + (int64_t)cnt * stride;    // ... and not behind the first probe that is not
        s_hi = q < hi ? q : hi;
      }
    }
    __syncthreads();
  }
  const int64_t r = s_lo;
  __syncthreads();
  return r;
}

// ---- segmented reduce over the sorted pairs -----------------------------------------------------------------
// One CUDA block per (segment, slice of 32*V values of the block): the 8 warps take consecutive eighths of the
// segment, each sums its run in source order (4 entries in flight), the 8 partial sums are added in warp order and
// the destination is read-modified-written once.  BYKEY: blockIdx.x is a destination, its segment is found by two
// binary searches (few destinations, many entries: the embedding gradient); otherwise blockIdx.x is a sorted entry
// and only segment heads work (many destinations).
template <typename T, int V, bool BYKEY>
__global__ void __launch_bounds__(256)
    k_segreduce(const uint32_t* __restrict__ keys, const int32_t* __restrict__ pos, const T* __restrict__ vals,
                int64_t n, int64_t block, int64_t nkeys, T* __restrict__ dest) {
  __shared__ T part[8][32 * V];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int64_t j, e;
  uint32_t key;
  if (BYKEY) {
    key = blockIdx.x;
    j = block_lower_bound(keys, n, key);
    e = block_lower_bound(keys, n, key + 1);
    if (j == e) return;
  } else {
    j = blockIdx.x;
    key = keys[j];
    if ((int64_t)key >= nkeys) return;          // out of range: flagged by the first sort pass, never written
    if (j > 0 && keys[j - 1] == key) return;    // not a head
    e = j + 1;
    while (e < n && keys[e] == key) ++e;
  }
  const int64_t b = ((int64_t)blockIdx.y * 32 + lane) * V;
  const int64_t len = e - j, per = (len + 7) / 8;
  const int64_t m0 = j + w * per, m1 = min(e, m0 + per);
  T acc[V];
#pragma unroll
  for (int v = 0; v < V; ++v) acc[v] = T(0);
  // the positions of the run are fetched 32 at a time (one per lane) and handed round by shuffle, so that the row
  // loads do not wait on an index load each; 8 rows in flight per lane, added in source order
  for (int64_t base = m0; base < m1; base += 32) {
    const int cnt = (int)min((int64_t)32, m1 - base);
    const int32_t pl = lane < cnt ? pos[base + lane] : 0;
    for (int u0 = 0; u0 < cnt; u0 += 8) {
      T x[8][V];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int32_t pu = __shfl_sync(0xffffffffu, pl, (u0 + u) & 31);
        if (u0 + u < cnt && b < block) {
          const T* src = vals + (int64_t)pu * block + b;
          if (V == 4 && sizeof(T) == 4) {
            const float4 t4 = __ldg(reinterpret_cast<const float4*>(src));
            x[u][0] = (T)t4.x; x[u][1 % V] = (T)t4.y; x[u][2 % V] = (T)t4.z; x[u][3 % V] = (T)t4.w;
          } else if (V == 2 && sizeof(T) == 8) {
            const double2 t2 = __ldg(reinterpret_cast<const double2*>(src));
            x[u][0] = (T)t2.x; x[u][1 % V] = (T)t2.y;
          } else {
#pragma unroll
            for (int v = 0; v < V; ++v) x[u][v] = __ldg(src + v);
          }
        } else {
#pragma unroll
          for (int v = 0; v < V; ++v) x[u][v] = T(0);
        }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int v = 0; v < V; ++v) acc[v] += x[u][v];
    }
  }
#pragma unroll
  for (int v = 0; v < V; ++v) part[w][lane * V + v] = acc[v];
  __syncthreads();
  if (w == 0 && b < block) {
    T* d = dest + (int64_t)key * block + b;
#pragma unroll
    for (int v = 0; v < V; ++v) {
      T s = part[0][lane * V + v];
#pragma unroll
      for (int q = 1; q < 8; ++q) s += part[q][lane * V + v];
      d[v] += s;
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
    k_slab_copy(const T* __restrict__ src, int64_t src_red, int64_t src_off, T* __restrict__ dst,
                int64_t dst_red, int64_t dst_off, int64_t inner, int64_t len, int64_t total) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int64_t slab = inner * len;
  int64_t o = i / slab, w = i - o * slab;  // w = ii + inner*r
  dst[o * inner * dst_red + dst_off * inner + w] = src[o * inner * src_red + src_off * inner + w];
}

// 128-bit form: every slab (inner*len elements) and both slab origins are multiples of the vector width
template <typename T>
__global__ void __launch_bounds__(256)
    k_slab_copy_v(const T* __restrict__ src, int64_t src_red, int64_t src_off, T* __restrict__ dst,
                  int64_t dst_red, int64_t dst_off, int64_t inner, int64_t len, int64_t total_vec) {
  constexpr int N = 16 / sizeof(T);
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total_vec) return;
  const int64_t slab = inner * len, sv = slab / N;
  int64_t o = i / sv, w = (i - o * sv) * N;
  *reinterpret_cast<uint4*>(dst + o * inner * dst_red + dst_off * inner + w) =
      *reinterpret_cast<const uint4*>(src + o * inner * src_red + src_off * inner + w);
}

// cat(xs...; dims) in ONE launch (src/cat.jl:27-31): dst viewed as inner x tot x outer, source j as inner x len[j] x
// outer; a thread owns one 16-byte piece (V elements) of the destination and finds its source by the running offsets.
constexpr int kCatMax = 8;
struct CatArgs {
  const void* src[kCatMax];
  int64_t len[kCatMax];      // extent of source j along the concatenated dimension
  int64_t off[kCatMax + 1];  // running offsets along that dimension; off[n] = tot
  int n;
  int64_t inner, outer;
};
template <typename T, int V>
__global__ void __launch_bounds__(256) k_cat(const __grid_constant__ CatArgs a, T* __restrict__ dst, int64_t total_vec) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total_vec) return;
  const int64_t e = i * V, slab = a.inner * a.off[a.n];
  const int64_t o = e / slab, w = e - o * slab;
  const int64_t r = w / a.inner;
  int j = 0;
#pragma unroll
  for (int q = 1; q < kCatMax; ++q)
    if (q < a.n && r >= a.off[q]) j = q;
  const T* __restrict__ s = (const T*)a.src[j] + o * a.inner * a.len[j] + (w - a.off[j] * a.inner);
  if (V == 1) dst[e] = s[0];
  else *reinterpret_cast<uint4*>(dst + e) = *reinterpret_cast<const uint4*>(s);
}

template <typename T>
__global__ void __launch_bounds__(256)
    k_transpose(const T* __restrict__ x, int64_t rows, int64_t cols, T* __restrict__ out) {
  __shared__ T tile[32][33];
  int64_t r0 = (int64_t)blockIdx.x * 32, c0 = (int64_t)blockIdx.y * 32;
  int tx = threadIdx.x % 32, ty = threadIdx.x / 32;  // 32 x 8
  for (int k = ty; k < 32; k += 8) {
    int64_t r = r0 + tx, c = c0 + k;
    if (r < rows && c < cols) tile[k][tx] = x[r + c * rows];
  }
  __syncthreads();
  for (int k = ty; k < 32; k += 8) {
    int64_t c = c0 + tx, r = r0 + k;
    if (r < rows && c < cols) out[c + r * cols] = tile[tx][k];
  }
}

// out = permutedims(x, perm) for <= 4 dims: thread per output element (out is written coalesced)
struct PermArgs {
  int64_t oshape[4];   // shape of out
  int64_t xstride[4];  // stride in x of each out dimension
  int64_t n;
};
template <typename T>
__global__ void __launch_bounds__(256) k_permute(const T* __restrict__ x, T* __restrict__ out, const PermArgs a) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  int64_t r = i, off = 0;
#pragma unroll
  for (int d = 0; d < 4; ++d) {
    const int64_t q = r / a.oshape[d];
    off += (r - q * a.oshape[d]) * a.xstride[d];
    r = q;
  }
  out[i] = x[off];
}

template <typename T>
__global__ void __launch_bounds__(256)
    k_axpy(T alpha, const T* __restrict__ x, T* __restrict__ y, int64_t n) {
  int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 4 <= n && (((uintptr_t)(x + i) | (uintptr_t)(y + i)) & (4 * sizeof(T) - 1)) == 0 &&
      sizeof(T) == 4) {
    float4 a = *reinterpret_cast<const float4*>(x + i);
    float4 b = *reinterpret_cast<float4*>(y + i);
    b.x = fmaf((float)alpha, a.x, b.x);
    b.y = fmaf((float)alpha, a.y, b.y);
    b.z = fmaf((float)alpha, a.z, b.z);
    b.w = fmaf((float)alpha, a.w, b.w);
    *reinterpret_cast<float4*>(y + i) = b;
  } else {
    for (int k = 0; k < 4 && i + k < n; ++k) y[i + k] = fma(alpha, x[i + k], y[i + k]);
  }
}

// x .*= alpha; a thread owns one aligned 16-byte piece (the ragged head / tail elements go one by one)
template <typename T>
__global__ void __launch_bounds__(256) k_scale(T alpha, T* __restrict__ x, int64_t n) {
  constexpr int N = 16 / sizeof(T);
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * N;
  if (i >= n) return;
  if (i + N <= n && (((uintptr_t)(x + i)) & 15) == 0) {
    Pack<T> v = *reinterpret_cast<Pack<T>*>(x + i);
#pragma unroll
    for (int k = 0; k < N; ++k) v.v[k] *= alpha;
    *reinterpret_cast<Pack<T>*>(x + i) = v;
  } else {
    for (int k = 0; k < N && i + k < n; ++k) x[i + k] *= alpha;
  }
}

template <typename T>
static ag_status scatter_add_impl(T* dest, int64_t dest_len, const int64_t* idx, const T* vals,
                                  int64_t n, int64_t block) {
  Context& c = ctx();
  AG_CHECK(n < ((int64_t)1 << 31), AG_EUNSUPPORTED, "ag_scatter_add: n=%lld too large", (long long)n);
  const int64_t nkeys = dest_len / block;
  AG_CHECK(nkeys < ((int64_t)1 << 31), AG_EUNSUPPORTED, "ag_scatter_add: %lld destinations", (long long)nkeys);
  // bits that distinguish the keys 0..nkeys (nkeys itself marks an out-of-range entry)
  int bits = 1;
  while (bits < 32 && ((int64_t)1 << bits) <= nkeys) ++bits;
  const int passes = (bits + kRadixBits - 1) / kRadixBits;
  // tile: 2048 entries per block, more when that would need more than 256 blocks (the scan is one block)
  int64_t tile = 2048;
  if (cdiv(n, tile) > 256) tile = cdiv(cdiv(n, (int64_t)256), (int64_t)256) * 256;
  const int64_t nblk = cdiv(n, tile);
  auto al = [](size_t b) { return (b + 255) & ~size_t(255); };
  const size_t sz_k = al((size_t)n * 4), sz_h = al((size_t)nblk * kRadixBins * 4);
  char* ws = (char*)workspace(256 + 4 * sz_k + sz_h);
  if (!ws) return AG_ENOMEM;
  int32_t* flag = oob_flag();
  if (!flag) return AG_ENOMEM;
  uint32_t* kbuf[2] = {(uint32_t*)(ws + 256), (uint32_t*)(ws + 256 + sz_k)};
  int32_t* pbuf[2] = {(int32_t*)(ws + 256 + 2 * sz_k), (int32_t*)(ws + 256 + 3 * sz_k)};
  uint32_t* hist = (uint32_t*)(ws + 256 + 4 * sz_k);
  // out-of-range destinations are a BoundsError in the reference.  The host wrapper checks the bounds of the
  // index vectors it normalised; the device check is sticky and surfaces at the next ag_sync (no per-call
  // host synchronisation, so a whole step stays capturable); out-of-range entries are skipped, never written.
  int cur = 0;
  for (int p = 0; p < passes; ++p) {
    const int shift = p * kRadixBits;
    if (p == 0) {
      k_radix_hist<true><<<(unsigned)nblk, 256, 0, c.stream>>>(idx, n, tile, shift, nkeys, flag, hist);
      AG_LAUNCHED();
      k_radix_scan<<<1, 1024, 0, c.stream>>>(hist, nblk * kRadixBins);
      AG_LAUNCHED();
      k_radix_scatter<true><<<(unsigned)nblk, 256, 0, c.stream>>>(idx, nullptr, n, tile, shift, nkeys, hist, kbuf[0], pbuf[0]);
      AG_LAUNCHED();
      cur = 0;
    } else {
      k_radix_hist<false><<<(unsigned)nblk, 256, 0, c.stream>>>(kbuf[cur], n, tile, shift, nkeys, flag, hist);
      AG_LAUNCHED();
      k_radix_scan<<<1, 1024, 0, c.stream>>>(hist, nblk * kRadixBins);
      AG_LAUNCHED();
      k_radix_scatter<false><<<(unsigned)nblk, 256, 0, c.stream>>>(kbuf[cur], pbuf[cur], n, tile, shift, nkeys, hist,
                                                                  kbuf[cur ^ 1], pbuf[cur ^ 1]);
      AG_LAUNCHED();
      cur ^= 1;
    }
  }
  const uint32_t* keys = kbuf[cur];
  const int32_t* pos = pbuf[cur];
  constexpr int VMAX = 16 / sizeof(T);
  const bool vec = block % VMAX == 0 && ((uintptr_t)vals & 15) == 0 && ((uintptr_t)dest & 15) == 0;
  const int V = vec ? VMAX : 1;
  if (cdiv(block, (int64_t)V) < 32) {           // narrow value blocks
    int lanes = 1;
    while (lanes < block && lanes < 128) lanes *= 2;
    k_segreduce_narrow<T><<<(unsigned)cdiv(n, (int64_t)(256 / lanes)), 256, 0, c.stream>>>(keys, pos, vals, n, block, nkeys,
                                                                                       lanes, dest);
    AG_LAUNCHED();
    return AG_OK;
  }
  const bool bykey = nkeys * 8 <= n;
  dim3 grid((unsigned)(bykey ? nkeys : n), (unsigned)cdiv(block, (int64_t)32 * V));
  AG_CHECK(grid.y <= 65535, AG_EUNSUPPORTED, "ag_scatter_add: block of %lld values", (long long)block);
  if (vec) {
    if (bykey) k_segreduce<T, VMAX, true><<<grid, 256, 0, c.stream>>>(keys, pos, vals, n, block, nkeys, dest);
    else k_segreduce<T, VMAX, false><<<grid, 256, 0, c.stream>>>(keys, pos, vals, n, block, nkeys, dest);
  } else {
    if (bykey) k_segreduce<T, 1, true><<<grid, 256, 0, c.stream>>>(keys, pos, vals, n, block, nkeys, dest);
    else k_segreduce<T, 1, false><<<grid, 256, 0, c.stream>>>(keys, pos, vals, n, block, nkeys, dest);
  }
  AG_LAUNCHED();
  return AG_OK;
}

}  // namespace agb

using namespace agb;

extern "C" {

ag_status ag_gather(int32_t dtype, const void* x, const int64_t* idx, void* out, int64_t n) {
  AG_REQUIRE_INIT();
  if (n == 0) return AG_OK;
  AG_CHECK(x && idx && out && n > 0, AG_EINVAL, "ag_gather: bad arguments");
  Context& c = ctx();
  unsigned nb = (unsigned)cdiv(n, 256);
  if (dtype == AG_F32) k_gather<float><<<nb, 256, 0, c.stream>>>((const float*)x, idx, (float*)out, n);
  else if (dtype == AG_F64) k_gather<double><<<nb, 256, 0, c.stream>>>((const double*)x, idx, (double*)out, n);
  else { set_error("ag_gather: dtype %d", dtype); return AG_EUNSUPPORTED; }
  AG_LAUNCHED();
  return AG_OK;
}

ag_status ag_gather_cols(int32_t dtype, const void* x, int64_t rows, const int64_t* col, void* out,
                         int64_t ncols_out) {
  AG_REQUIRE_INIT();
  int64_t total = rows * ncols_out;
  if (total == 0) return AG_OK;
  AG_CHECK(x && col && out && rows > 0 && ncols_out > 0, AG_EINVAL, "ag_gather_cols: bad arguments");
  Context& c = ctx();
  AG_CHECK(dtype == AG_F32 || dtype == AG_F64, AG_EUNSUPPORTED, "ag_gather_cols: dtype %d", dtype);
  const int N = dtype == AG_F32 ? 4 : 2;
  if (rows % N == 0 && aligned16(x) && aligned16(out)) {
    const int64_t tv = total / N;
    unsigned nbv = (unsigned)cdiv(tv, 256);
    if (dtype == AG_F32) k_gather_cols_v<float><<<nbv, 256, 0, c.stream>>>((const float*)x, rows, col, (float*)out, tv);
    else k_gather_cols_v<double><<<nbv, 256, 0, c.stream>>>((const double*)x, rows, col, (double*)out, tv);
    AG_LAUNCHED();
    return AG_OK;
  }
  unsigned nb = (unsigned)cdiv(total, 256);
  if (dtype == AG_F32) k_gather_cols<float><<<nb, 256, 0, c.stream>>>((const float*)x, rows, col, (float*)out, total);
  else k_gather_cols<double><<<nb, 256, 0, c.stream>>>((const double*)x, rows, col, (double*)out, total);
  AG_LAUNCHED();
  return AG_OK;
}

ag_status ag_scatter_add(int32_t dtype, void* dest, int64_t dest_len, const int64_t* idx,
                         const void* vals, int64_t n, int64_t block) {
  AG_REQUIRE_INIT();
  if (n == 0) return AG_OK;
  AG_CHECK(dest && idx && vals && n > 0 && block >= 1 && dest_len >= 0 && dest_len % block == 0,
           AG_EINVAL, "ag_scatter_add: bad arguments");
  if (dtype == AG_F32) return scatter_add_impl<float>((float*)dest, dest_len, idx, (const float*)vals, n, block);
  if (dtype == AG_F64) return scatter_add_impl<double>((double*)dest, dest_len, idx, (const double*)vals, n, block);
  set_error("ag_scatter_add: dtype %d", dtype);
  return AG_EUNSUPPORTED;
}

ag_status ag_slab_copy(int32_t dtype, const void* src, int64_t src_red, int64_t src_off, void* dst,
                       int64_t dst_red, int64_t dst_off, int64_t inner, int64_t len, int64_t outer) {
  AG_REQUIRE_INIT();
  int64_t total = inner * len * outer;
  if (total == 0) return AG_OK;
  AG_CHECK(src && dst && inner > 0 && len > 0 && outer > 0 && src_off >= 0 && dst_off >= 0 &&
               src_off + len <= src_red && dst_off + len <= dst_red,
           AG_EINVAL, "ag_slab_copy: slab out of range (BoundsError)");
  Context& c = ctx();
  AG_CHECK(dtype == AG_F32 || dtype == AG_F64, AG_EUNSUPPORTED, "ag_slab_copy: dtype %d", dtype);
  {
    const int N = dtype == AG_F32 ? 4 : 2;
    const int64_t slab = inner * len;
    if (slab % N == 0 && (inner * src_red) % N == 0 && (inner * dst_red) % N == 0 && (inner * src_off) % N == 0 &&
        (inner * dst_off) % N == 0 && aligned16(src) && aligned16(dst)) {
      const int64_t tv = total / N;
      unsigned nbv = (unsigned)cdiv(tv, 256);
      if (dtype == AG_F32)
        k_slab_copy_v<float><<<nbv, 256, 0, c.stream>>>((const float*)src, src_red, src_off, (float*)dst, dst_red, dst_off, inner, len, tv);
      else
        k_slab_copy_v<double><<<nbv, 256, 0, c.stream>>>((const double*)src, src_red, src_off, (double*)dst, dst_red, dst_off, inner, len, tv);
      AG_LAUNCHED();
      return AG_OK;
    }
  }
  unsigned nb = (unsigned)cdiv(total, 256);
  if (dtype == AG_F32)
    k_slab_copy<float><<<nb, 256, 0, c.stream>>>((const float*)src, src_red, src_off, (float*)dst, dst_red, dst_off, inner, len, total);
  else if (dtype == AG_F64)
    k_slab_copy<double><<<nb, 256, 0, c.stream>>>((const double*)src, src_red, src_off, (double*)dst, dst_red, dst_off, inner, len, total);
  else { set_error("ag_slab_copy: dtype %d", dtype); return AG_EUNSUPPORTED; }
  AG_LAUNCHED();
  return AG_OK;
}

ag_status ag_cat(int32_t dtype, int32_t n, const void* const* src, const int64_t* len, void* dst, int64_t inner,
                 int64_t outer) {
  AG_REQUIRE_INIT();
  AG_CHECK(dtype == AG_F32 || dtype == AG_F64, AG_EUNSUPPORTED, "ag_cat: dtype %d", dtype);
  AG_CHECK(n >= 1 && n <= kCatMax && src && len && inner >= 0 && outer >= 0, AG_EINVAL, "ag_cat: bad arguments (n = %d, max %d)",
           n, kCatMax);
  CatArgs a;
  a.n = n;
  a.inner = inner;
  a.outer = outer;
  int64_t tot = 0;
  const int N = dtype == AG_F32 ? 4 : 2;
  bool vec = aligned16(dst);
  for (int j = 0; j < n; ++j) {
    AG_CHECK(len[j] >= 0 && (len[j] == 0 || src[j]), AG_EINVAL, "ag_cat: bad source %d", j);
    a.src[j] = src[j];
    a.len[j] = len[j];
    a.off[j] = tot;
    tot += len[j];
    vec = vec && aligned16(src[j]) && (len[j] * inner) % N == 0;
  }
  for (int j = n; j <= kCatMax; ++j) a.off[j] = tot;
  for (int j = n; j < kCatMax; ++j) { a.src[j] = nullptr; a.len[j] = 0; }
  const int64_t total = inner * tot * outer;
  if (total == 0) return AG_OK;
  AG_CHECK(dst, AG_EINVAL, "ag_cat: null destination");
  Context& c = ctx();
  if (vec) {
    const int64_t tv = total / N;
    if (dtype == AG_F32) k_cat<float, 4><<<(unsigned)cdiv(tv, (int64_t)256), 256, 0, c.stream>>>(a, (float*)dst, tv);
    else k_cat<double, 2><<<(unsigned)cdiv(tv, (int64_t)256), 256, 0, c.stream>>>(a, (double*)dst, tv);
  } else {
    if (dtype == AG_F32) k_cat<float, 1><<<(unsigned)cdiv(total, (int64_t)256), 256, 0, c.stream>>>(a, (float*)dst, total);
    else k_cat<double, 1><<<(unsigned)cdiv(total, (int64_t)256), 256, 0, c.stream>>>(a, (double*)dst, total);
  }
  AG_LAUNCHED();
  return AG_OK;
}

ag_status ag_transpose(int32_t dtype, const void* x, int64_t rows, int64_t cols, void* out) {
  AG_REQUIRE_INIT();
  if (rows * cols == 0) return AG_OK;
  AG_CHECK(x && out && rows > 0 && cols > 0, AG_EINVAL, "ag_transpose: bad arguments");
  AG_CHECK(cdiv(cols, 32) <= 65535, AG_EUNSUPPORTED, "ag_transpose: too many columns");
  Context& c = ctx();
  dim3 grid((unsigned)cdiv(rows, 32), (unsigned)cdiv(cols, 32));
  if (dtype == AG_F32) k_transpose<float><<<grid, 256, 0, c.stream>>>((const float*)x, rows, cols, (float*)out);
  else if (dtype == AG_F64) k_transpose<double><<<grid, 256, 0, c.stream>>>((const double*)x, rows, cols, (double*)out);
  else { set_error("ag_transpose: dtype %d", dtype); return AG_EUNSUPPORTED; }
  AG_LAUNCHED();
  return AG_OK;
}

ag_status ag_permute(int32_t dtype, const void* x, int32_t ndims, const int64_t* shape, const int32_t* perm,
                     void* out) {
  AG_REQUIRE_INIT();
  AG_CHECK(ndims >= 1 && ndims <= 4 && shape && perm, AG_EUNSUPPORTED, "ag_permute: ndims=%d (max 4)", ndims);
  PermArgs a;
  int64_t xs[4] = {1, 1, 1, 1}, n = 1;
  bool seen[4] = {false, false, false, false};
  for (int d = 0; d < ndims; ++d) {
    AG_CHECK(shape[d] >= 0, AG_EINVAL, "ag_permute: negative extent");
    xs[d] = n;
    n *= shape[d];
  }
  for (int d = 0; d < 4; ++d) {
    if (d < ndims) {
      AG_CHECK(perm[d] >= 0 && perm[d] < ndims && !seen[perm[d]], AG_EINVAL, "ag_permute: not a permutation");
      seen[perm[d]] = true;
      a.oshape[d] = shape[perm[d]];
      a.xstride[d] = xs[perm[d]];
    } else {
      a.oshape[d] = 1;
      a.xstride[d] = 0;
    }
  }
  a.n = n;
  if (n == 0) return AG_OK;
  AG_CHECK(x && out, AG_EINVAL, "ag_permute: null pointer");
  Context& c = ctx();
  unsigned nb = (unsigned)cdiv(n, 256);
  if (dtype == AG_F32) k_permute<float><<<nb, 256, 0, c.stream>>>((const float*)x, (float*)out, a);
  else if (dtype == AG_F64) k_permute<double><<<nb, 256, 0, c.stream>>>((const double*)x, (double*)out, a);
  else { set_error("ag_permute: dtype %d", dtype); return AG_EUNSUPPORTED; }
  AG_LAUNCHED();
  return AG_OK;
}

ag_status ag_axpy(int32_t dtype, double alpha, const void* x, void* y, int64_t n) {
  AG_REQUIRE_INIT();
  if (n == 0) return AG_OK;
  AG_CHECK(x && y && n > 0, AG_EINVAL, "ag_axpy: bad arguments");
  Context& c = ctx();
  unsigned nb = (unsigned)cdiv(n, 1024);
  if (dtype == AG_F32) k_axpy<float><<<nb, 256, 0, c.stream>>>((float)alpha, (const float*)x, (float*)y, n);
  else if (dtype == AG_F64) k_axpy<double><<<nb, 256, 0, c.stream>>>(alpha, (const double*)x, (double*)y, n);
  else { set_error("ag_axpy: dtype %d", dtype); return AG_EUNSUPPORTED; }
  AG_LAUNCHED();
  return AG_OK;
}

ag_status ag_scale(int32_t dtype, double alpha, void* x, int64_t n) {
  AG_REQUIRE_INIT();
  if (n == 0) return AG_OK;
  AG_CHECK(x && n > 0, AG_EINVAL, "ag_scale: bad arguments");
  Context& c = ctx();
  unsigned nb = (unsigned)cdiv(n, dtype == AG_F32 ? 1024 : 512);
  if (dtype == AG_F32) k_scale<float><<<nb, 256, 0, c.stream>>>((float)alpha, (float*)x, n);
  else if (dtype == AG_F64) k_scale<double><<<nb, 256, 0, c.stream>>>(alpha, (double*)x, n);
  else { set_error("ag_scale: dtype %d", dtype); return AG_EUNSUPPORTED; }
  AG_LAUNCHED();
  return AG_OK;
}

}  // extern "C"
