// scan.cu -- running folds and mask-driven selection:
//   cumsum / cummax / cummin along an axis and the flat cumsum   (Mat.scala:3635-3745, 929-964)
//   Mat.where(cond, x, y) / Mat.where(cond, sx, sy)             (Mat.scala:1305-1336)
//   m(mask): the selected elements in row-major order, 1 x k    (Mat.scala:4008-4021)
//
// The reference's running folds are sequential per lane, and the fixture pins them bit for bit
// (cum0fnv cum1fnv cmax0fnv cmin1fnv cummaxfnv cumlast, MatParityGen.scala:199,220,278-281,847-850).
// Floating-point addition does not re-associate, so a lane is never split here: every lane is folded by
// ONE thread in increasing index order and the parallelism is across lanes (and across the loads of a
// lane, which are issued ahead of the dependent adds).  The flat cumsum is a single lane: one warp loads
// coalesced and folds in order, bound by the 8-cycle DADD latency by construction.
// cummax / cummin use java.lang.Double.compare's total order (NaN greatest, -0.0 < +0.0) and replace on
// a strictly better value only, like cumExtremumD.
#include <math.h>

#include "common.cuh"

namespace um {

enum { CUM_SUM = 0, CUM_MAX = 1, CUM_MIN = 2 };

// running state of one lane: the value, and for cummax / cummin its order key (kept so that only the incoming
// element's key is computed per step)
template <int KIND, class T>
struct CumAcc {
  T v;
  i64 k;
  __device__ __forceinline__ void seed(T first) {
    if (KIND == CUM_SUM) { v = T(0); k = 0; }
    else { v = first; k = total_key(first); }
  }
  __device__ __forceinline__ void step(T x) {
    if (KIND == CUM_SUM) {
      v += x;
    } else {
      const i64 kx = total_key(x);
      const bool better = KIND == CUM_MAX ? kx > k : kx < k;
      if (better) { v = x; k = kx; }
    }
  }
};

// Lanes adjacent in memory (axis 0 of a row-major matrix): thread = lane.  The loads of the next 16 elements are
// issued before the current 16 are folded, so 16-32 loads per thread are in flight behind the dependent chain.
constexpr int CLT_THREADS = 64;
template <int KIND, class T>
__global__ void __launch_bounds__(CLT_THREADS) k_cum_lane_per_thread(const T* __restrict__ p, i64 s_lane, i64 s_red, i64 n_lane, i64 n_red,
                                                                      T* __restrict__ out, i64 o_lane, i64 o_red) {
  constexpr int U = 16;
  const i64 l = (i64)blockIdx.x * CLT_THREADS + threadIdx.x;
  if (l >= n_lane) return;
  const T* q = p + l * s_lane;
  T* o = out + l * o_lane;
  CumAcc<KIND, T> acc;
  acc.seed(q[0]);
  T xa[U], xb[U];
  i64 i = 0;
  if (n_red >= U) {
#pragma unroll
    for (int u = 0; u < U; ++u) xa[u] = q[u * s_red];
  }
  // invariant: xa holds elements [i, i + U) whenever i + U <= n_red
  while (i + U <= n_red) {
    const bool more = i + 2 * U <= n_red;
    if (more) {
#pragma unroll
      for (int u = 0; u < U; ++u) xb[u] = q[(i + U + u) * s_red];
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      acc.step(xa[u]);
      o[(i + u) * o_red] = acc.v;
    }
    i += U;
    if (!more) break;
    const bool more2 = i + 2 * U <= n_red;
    if (more2) {
#pragma unroll
      for (int u = 0; u < U; ++u) xa[u] = q[(i + U + u) * s_red];
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      acc.step(xb[u]);
      o[(i + u) * o_red] = acc.v;
    }
    i += U;
    if (!more2) break;
  }
  for (; i < n_red; ++i) {
    acc.step(q[i * s_red]);
    o[i * o_red] = acc.v;
  }
}

// Lanes whose elements are adjacent in memory (axis 1 of a row-major matrix): a warp owns 32 lanes and walks
// them in 32 x 32 tiles -- coalesced loads (a whole tile per warp in flight, the next tile's issued before this
// one is folded) into a padded shared tile, thread j folds lane j's 32 elements in order carrying its accumulator
// across tiles, coalesced stores back out.
constexpr int CTL_WARPS = 2;
template <int KIND, class T>
__global__ void __launch_bounds__(CTL_WARPS * 32) k_cum_lane_tiles(const T* __restrict__ p, i64 s_lane, i64 s_red, i64 n_lane, i64 n_red,
                                                                    T* __restrict__ out, i64 o_lane, i64 o_red) {
  __shared__ T tile[CTL_WARPS][32][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const i64 l0 = ((i64)blockIdx.x * CTL_WARPS + w) * 32;
  if (l0 >= n_lane) return;
  const int nl = (int)min((i64)32, n_lane - l0);
  CumAcc<KIND, T> acc;
  acc.seed(lane < nl ? p[(l0 + lane) * s_lane] : T(0));
  T x[32];
  auto fetch = [&](i64 t0) {
    const bool col = t0 + lane < n_red;
#pragma unroll
    for (int r = 0; r < 32; ++r)
      if (col && r < nl) x[r] = p[(l0 + r) * s_lane + (t0 + lane) * s_red];
  };
  fetch(0);
  for (i64 t0 = 0; t0 < n_red; t0 += 32) {
    const int nt = (int)min((i64)32, n_red - t0);
    if (lane < nt) {
#pragma unroll
      for (int r = 0; r < 32; ++r)
        if (r < nl) tile[w][r][lane] = x[r];
    }
    __syncwarp();
    if (t0 + 32 < n_red) fetch(t0 + 32);
    if (lane < nl) {
#pragma unroll 8
      for (int k = 0; k < nt; ++k) {
        acc.step(tile[w][lane][k]);
        tile[w][lane][k] = acc.v;
      }
    }
    __syncwarp();
    if (lane < nt) {
#pragma unroll 8
      for (int r = 0; r < nl; ++r) out[(l0 + r) * o_lane + (t0 + lane) * o_red] = tile[w][r][lane];
    }
    __syncwarp();
  }
}

// Flat running fold over the row-major element order (m.cumsum with no axis): ONE lane.  A warp loads 4 x 32
// elements ahead, every thread folds all of them in order (the adds are the dependent chain; the shuffles are
// not on it) and thread j keeps the values of its own positions.
template <int KIND, class T, bool FLAT>
__global__ void __launch_bounds__(32) k_cum_flat(const T* __restrict__ p, i64 rs, i64 cs, i64 rows, i64 cols, T* __restrict__ out) {
  constexpr int U = 4;
  const int lane = threadIdx.x;
  const i64 n = rows * cols;
  auto at = [&](i64 e) { return FLAT ? p[e] : p[(e / cols) * rs + (e % cols) * cs]; };
  CumAcc<KIND, T> acc;
  acc.seed(n > 0 ? at(0) : T(0));
  T x[U], xn[U];
  auto fetch = [&](T* dst, i64 e0) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const i64 e = e0 + u * 32 + lane;
      dst[u] = e < n ? at(e) : T(0);
    }
  };
  fetch(x, 0);
  for (i64 e0 = 0; e0 < n; e0 += 32 * U) {
    fetch(xn, e0 + 32 * U);   // the next 128 elements load while these are folded
    T mine[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      mine[u] = T(0);
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        const T v = __shfl_sync(0xffffffffu, x[u], k);
        if (e0 + u * 32 + k < n) acc.step(v);   // uniform over the warp
        if (lane == k) mine[u] = acc.v;
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const i64 e = e0 + u * 32 + lane;
      if (e < n) out[e] = mine[u];
      x[u] = xn[u];
    }
  }
}

template <int KIND, class T>
static int cumulative_typed(const um_view* a, int axis, const um_view* out) {
  const T* p = static_cast<const T*>(a->data) + a->offset;
  T* o = static_cast<T*>(out->data) + out->offset;
  cudaStream_t st = ctx()->stream;
  if (axis < 0) {
    if ((a->cs == 1 || a->cols == 1) && (a->rs == a->cols || a->rows == 1)) k_cum_flat<KIND, T, true><<<1, 32, 0, st>>>(p, a->rs, a->cs, a->rows, a->cols, o);
    else k_cum_flat<KIND, T, false><<<1, 32, 0, st>>>(p, a->rs, a->cs, a->rows, a->cols, o);
    UM_LAUNCHED();
    return UM_OK;
  }
  const i64 n_lane = axis == 0 ? a->cols : a->rows, n_red = axis == 0 ? a->rows : a->cols;
  const i64 s_lane = axis == 0 ? a->cs : a->rs, s_red = axis == 0 ? a->rs : a->cs;
  const i64 o_lane = axis == 0 ? out->cs : out->rs, o_red = axis == 0 ? out->rs : out->cs;
  auto mag = [](i64 s) { return s < 0 ? -s : s; };
  // which index is adjacent in memory?  (a stride-0 broadcast lane axis counts as adjacent)
  if (mag(s_lane) <= mag(s_red) || n_red == 1) {
    k_cum_lane_per_thread<KIND, T><<<(unsigned)((n_lane + CLT_THREADS - 1) / CLT_THREADS), CLT_THREADS, 0, st>>>(p, s_lane, s_red, n_lane, n_red, o, o_lane, o_red);
  } else {
    k_cum_lane_tiles<KIND, T><<<(unsigned)((n_lane + CTL_WARPS * 32 - 1) / (CTL_WARPS * 32)), CTL_WARPS * 32, 0, st>>>(p, s_lane, s_red, n_lane, n_red, o, o_lane, o_red);
  }
  UM_LAUNCHED();
  return UM_OK;
}

template <class T>
static int cumulative_kind(int kind, const um_view* a, int axis, const um_view* out) {
  switch (kind) {
    case CUM_SUM: return cumulative_typed<CUM_SUM, T>(a, axis, out);
    case CUM_MAX: return cumulative_typed<CUM_MAX, T>(a, axis, out);
    case CUM_MIN: return cumulative_typed<CUM_MIN, T>(a, axis, out);
  }
  return fail(UM_ERR_ARG, "cumulative: unknown kind %d", kind);
}

// ---- where ---------------------------------------------------------------------------------------------
template <class T, bool SCALAR>
__global__ void __launch_bounds__(256) k_where(In<uint8_t> c, In<T> x, In<T> y, T sx, T sy, i64 rows, i64 cols, Out<T> o) {
  const i64 j = (i64)blockIdx.x * 64 + (threadIdx.x & 63);
  if (j >= cols) return;
  for (i64 i = (i64)blockIdx.y * 4 + (threadIdx.x >> 6); i < rows; i += (i64)gridDim.y * 4) {
    const bool t = c.p[i * c.rs + j * c.cs] != 0;
    T v;
    if (SCALAR) v = t ? sx : sy;
    else v = t ? x.p[i * x.rs + j * x.cs] : y.p[i * y.rs + j * y.cs];
    o.p[i * o.rs + j * o.cs] = v;
  }
}

// contiguous FP64 operands: two columns per thread, four rows in flight, both sources always loaded
template <bool SCALAR>
__global__ void __launch_bounds__(256) k_where_vec(const uint8_t* __restrict__ c, i64 c_rs, const double* __restrict__ x, i64 x_rs,
                                                    const double* __restrict__ y, i64 y_rs, double sx, double sy, i64 rows, i64 cols,
                                                    double* __restrict__ o, i64 o_rs) {
  const i64 j = ((i64)blockIdx.x * 64 + (threadIdx.x & 63)) * 2;
  if (j >= cols) return;
  const i64 step = (i64)gridDim.y * 4;
  i64 i = (i64)blockIdx.y * 4 + (threadIdx.x >> 6);
  auto one = [&](i64 r, uchar2 t, double2 a, double2 b) {
    __stcs(reinterpret_cast<double2*>(o + r * o_rs + j), make_double2(t.x ? a.x : b.x, t.y ? a.y : b.y));
  };
  for (; i + 3 * step < rows; i += 4 * step) {
    uchar2 t[4];
    double2 a[4], b[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const i64 r = i + u * step;
      t[u] = *reinterpret_cast<const uchar2*>(c + r * c_rs + j);
      if (SCALAR) { a[u] = make_double2(sx, sx); b[u] = make_double2(sy, sy); }
      else {
        a[u] = *reinterpret_cast<const double2*>(x + r * x_rs + j);
        b[u] = *reinterpret_cast<const double2*>(y + r * y_rs + j);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) one(i + u * step, t[u], a[u], b[u]);
  }
  for (; i < rows; i += step) {
    const uchar2 t = *reinterpret_cast<const uchar2*>(c + i * c_rs + j);
    double2 a = make_double2(sx, sx), b = make_double2(sy, sy);
    if (!SCALAR) {
      a = *reinterpret_cast<const double2*>(x + i * x_rs + j);
      b = *reinterpret_cast<const double2*>(y + i * y_rs + j);
    }
    one(i, t, a, b);
  }
}

template <class T>
static int where_typed(const um_view* c, const um_view* x, const um_view* y, double sx, double sy, const um_view* out) {
  const i64 rows = c->rows, cols = c->cols;
  if (rows == 0 || cols == 0) return UM_OK;
  i64 gy = (rows + 3) / 4;
  const i64 gx = (cols + 63) / 64;
  const i64 want = ((i64)ctx()->sm_count * 16 + gx - 1) / gx;
  if (gy > want) gy = want;
  if (gy > 65535) gy = 65535;
  if (sizeof(T) == 8 && cols % 2 == 0) {
    auto rowvec = [&](const um_view* v, int esz) {   // rows readable as aligned 2-element vectors?
      const char* q = static_cast<const char*>(v->data) + (size_t)v->offset * esz;
      return v->cs == 1 && (reinterpret_cast<uintptr_t>(q) % (2 * esz)) == 0 && (rows == 1 || v->rs % 2 == 0);
    };
    if (rowvec(c, 1) && rowvec(out, 8) && (!x || (rowvec(x, 8) && rowvec(y, 8)))) {
      const i64 gxv = (cols / 2 + 63) / 64;
      i64 gyv = (rows + 15) / 16;
      const i64 wantv = ((i64)ctx()->sm_count * 16 + gxv - 1) / gxv;
      if (gyv > wantv) gyv = wantv;
      if (gyv > 65535) gyv = 65535;
      if (gyv < 1) gyv = 1;
      dim3 gridv((unsigned)gxv, (unsigned)gyv);
      const uint8_t* cp = static_cast<const uint8_t*>(c->data) + c->offset;
      double* op = static_cast<double*>(out->data) + out->offset;
      if (x) k_where_vec<false><<<gridv, 256, 0, ctx()->stream>>>(cp, c->rs, static_cast<const double*>(x->data) + x->offset, x->rs,
                                                                 static_cast<const double*>(y->data) + y->offset, y->rs, 0, 0, rows, cols, op, out->rs);
      else k_where_vec<true><<<gridv, 256, 0, ctx()->stream>>>(cp, c->rs, nullptr, 0, nullptr, 0, sx, sy, rows, cols, op, out->rs);
      UM_LAUNCHED();
      return UM_OK;
    }
  }
  dim3 grid((unsigned)gx, (unsigned)gy);
  if (x) k_where<T, false><<<grid, 256, 0, ctx()->stream>>>(in_of<uint8_t>(c), in_of<T>(x), in_of<T>(y), T(0), T(0), rows, cols, out_of<T>(out));
  else k_where<T, true><<<grid, 256, 0, ctx()->stream>>>(in_of<uint8_t>(c), In<T>{nullptr, 0, 0}, In<T>{nullptr, 0, 0}, (T)sx, (T)sy, rows, cols, out_of<T>(out));
  UM_LAUNCHED();
  return UM_OK;
}

// ---- m(mask): stable compaction in row-major order ---------------------------------------------------------
constexpr int SEL_CHUNK = 4096;   // flat elements per block
constexpr int SEL_THREADS = 256;

// element e of the row-major order; FLAT = the view is contiguous (no 64-bit division)
template <bool FLAT, class T>
__device__ __forceinline__ T flat_at(const In<T>& m, i64 cols, i64 e) { return FLAT ? m.p[e] : m.p[(e / cols) * m.rs + (e % cols) * m.cs]; }

template <bool FLAT>
__global__ void __launch_bounds__(SEL_THREADS) k_sel_count(In<uint8_t> m, i64 cols, i64 n, i64* __restrict__ counts) {
  __shared__ int sm[SEL_THREADS / 32];
  const i64 e0 = (i64)blockIdx.x * SEL_CHUNK;
  int c = 0;
  for (int k = threadIdx.x; k < SEL_CHUNK; k += SEL_THREADS) {
    const i64 e = e0 + k;
    if (e < n && flat_at<FLAT>(m, cols, e) != 0) ++c;
  }
  for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < SEL_THREADS / 32; ++w) t += sm[w];
    counts[blockIdx.x] = t;
  }
}

// exclusive scan of the block counts in place; counts[nb] = total.  One block (integer sums are exact in any order):
// each thread folds 8 consecutive counts, warps scan with shuffles, one shared-memory step across the 32 warps; the
// next 8192 counts are loaded before the current ones are scanned.
// Block b scans counts[b * chunk, min(nb, (b + 1) * chunk)) and stores its total in total_out[b]; with one block and
// chunk >= nb that is the whole array and total_out = counts + nb.
__global__ void __launch_bounds__(1024) k_sel_scan(i64* __restrict__ counts_all, i64 nb_all, i64 chunk, i64* __restrict__ total_out) {
  constexpr int PER = 8;
  i64* __restrict__ counts = counts_all + (i64)blockIdx.x * chunk;
  const i64 nb = min(chunk, nb_all - (i64)blockIdx.x * chunk);
  __shared__ i64 wsum[32];
  __shared__ i64 carry_s;
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  if (t == 0) carry_s = 0;
  i64 v[PER], nxt[PER];
  auto fetch = [&](i64* dst, i64 b0) {
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      const i64 b = b0 + (i64)t * PER + k;
      dst[k] = b < nb ? counts[b] : 0;
    }
  };
  fetch(v, 0);
  __syncthreads();
  for (i64 b0 = 0; b0 < nb; b0 += 1024 * PER) {
    fetch(nxt, b0 + 1024 * PER);
    i64 tot = 0;
#pragma unroll
    for (int k = 0; k < PER; ++k) tot += v[k];
    i64 inc = tot;                                   // inclusive scan of the thread totals inside the warp
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const i64 o = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += o;
    }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    i64 wbase = 0;
    for (int q = 0; q < w; ++q) wbase += wsum[q];
    i64 run = carry_s + wbase + inc - tot;           // exclusive prefix of this thread's first count
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      const i64 b = b0 + (i64)t * PER + k;
      if (b < nb) counts[b] = run;
      run += v[k];
    }
    __syncthreads();
    if (t == 1023) carry_s = run;                    // the last thread ends on the running total
    __syncthreads();
#pragma unroll
    for (int k = 0; k < PER; ++k) v[k] = nxt[k];
  }
  if (t == 0) total_out[blockIdx.x] = carry_s;
}
// second level: add the scanned chunk totals back; the grand total goes to counts[nb]
__global__ void __launch_bounds__(1024) k_sel_scan_add(i64* __restrict__ counts, i64 nb, i64 chunk, const i64* __restrict__ chunk_off,
                                                        i64 nchunk) {
  const i64 add = chunk_off[blockIdx.x];
  const i64 b0 = (i64)blockIdx.x * chunk, b1 = min(nb, b0 + chunk);
  for (i64 b = b0 + threadIdx.x; b < b1; b += 1024) counts[b] += add;
  if (blockIdx.x == 0 && threadIdx.x == 0) counts[nb] = chunk_off[nchunk];
}
constexpr i64 SEL_SCAN_CHUNK = 8192;  // one pass of the 1024-thread scan

template <class T, bool FLAT_A, bool FLAT_M>
__global__ void __launch_bounds__(SEL_THREADS) k_sel_scatter(In<T> a, In<uint8_t> m, i64 cols, i64 n, const i64* __restrict__ offsets,
                                                              T* __restrict__ out, i64 out_n) {
  __shared__ int sm[SEL_THREADS / 32];
  const i64 e0 = (i64)blockIdx.x * SEL_CHUNK;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  i64 base = offsets[blockIdx.x];
  for (int k0 = 0; k0 < SEL_CHUNK; k0 += SEL_THREADS) {
    const i64 e = e0 + k0 + threadIdx.x;
    const bool t = e < n && flat_at<FLAT_M>(m, cols, e) != 0;
    const unsigned bal = __ballot_sync(0xffffffffu, t);
    if (lane == 0) sm[w] = __popc(bal);
    __syncthreads();
    int before = 0, total = 0;
#pragma unroll
    for (int q = 0; q < SEL_THREADS / 32; ++q) {
      const int c = sm[q];
      if (q < w) before += c;
      total += c;
    }
    if (t) {
      const i64 pos = base + before + __popc(bal & ((1u << lane) - 1u));
      if (pos < out_n) out[pos] = flat_at<FLAT_A>(a, cols, e);
    }
    base += total;
    __syncthreads();
  }
}

// Contiguous operand and mask: a warp owns 512 consecutive elements.  All 16 coalesced loads of the mask bytes and of
// the operand are issued up front; the mask bits stay in a register, positions come from ballots, and the only
// block-level step is the prefix over the 8 warp totals.
template <class T>
__global__ void __launch_bounds__(SEL_THREADS) k_sel_scatter_flat(const T* __restrict__ a, const uint8_t* __restrict__ m, i64 n,
                                                                   const i64* __restrict__ offsets, T* __restrict__ out, i64 out_n) {
  __shared__ int sm[SEL_THREADS / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const i64 e0 = (i64)blockIdx.x * SEL_CHUNK + (i64)w * 512;
  unsigned bits = 0;
  T av[16];
  int total = 0;
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const i64 e = e0 + k * 32 + lane;
    const bool in = e < n;
    const bool t = in && m[e] != 0;
    av[k] = in ? a[e] : T(0);
    bits |= (t ? 1u : 0u) << k;
  }
#pragma unroll
  for (int k = 0; k < 16; ++k) total += __popc(__ballot_sync(0xffffffffu, (bits >> k) & 1u));
  if (lane == 0) sm[w] = total;
  __syncthreads();
  i64 base = offsets[blockIdx.x];
  for (int q = 0; q < w; ++q) base += sm[q];
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const bool t = (bits >> k) & 1u;
    const unsigned bal = __ballot_sync(0xffffffffu, t);
    if (t) {
      const i64 pos = base + __popc(bal & ((1u << lane) - 1u));
      if (pos < out_n) out[pos] = av[k];
    }
    base += __popc(bal);
  }
}

// Contiguous operand and mask, staged through shared memory so that BOTH directions are coalesced: the block's 4096
// operand elements are loaded with unit-stride warp loads into a staging buffer, every thread owns 16 consecutive mask
// bytes (one 16-byte load), its selected elements are moved to their place (exclusive scan of the per-thread
// counts), and the block's k_b selected elements leave with unit-stride stores.  Needs a 16-byte aligned
// mask pointer; 32 KB of shared memory for 8-byte elements (compaction in place).
template <class T>
__global__ void __launch_bounds__(SEL_THREADS) k_sel_scatter_staged(const T* __restrict__ a, const uint8_t* __restrict__ m, i64 n,
                                                                     const i64* __restrict__ offsets, T* __restrict__ out, i64 out_n) {
  extern __shared__ __align__(16) unsigned char sel_smem[];
  T* stage = reinterpret_cast<T*>(sel_smem);
  __shared__ int wsum[SEL_THREADS / 32];
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  const i64 e0 = (i64)blockIdx.x * SEL_CHUNK;
  // operand: unit-stride loads
#pragma unroll
  for (int k = 0; k < SEL_CHUNK / SEL_THREADS; ++k) {
    const i64 e = e0 + k * SEL_THREADS + t;
    if (e < n) stage[k * SEL_THREADS + t] = a[e];
  }
  // mask: 16 consecutive bytes per thread -> one bit each
  unsigned bits = 0;
  const i64 me = e0 + (i64)t * 16;
  if (me + 16 <= n) {
    const uint4 v = *reinterpret_cast<const uint4*>(m + me);
    const unsigned wds[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const unsigned nz = __vcmpne4(wds[q], 0u);  // 0xff per non-zero byte
      bits |= (((nz >> 0) & 1u) | ((nz >> 7) & 2u) | ((nz >> 14) & 4u) | ((nz >> 21) & 8u)) << (4 * q);
    }
  } else {
    for (int k = 0; k < 16; ++k)
      if (me + k < n && m[me + k] != 0) bits |= 1u << k;
  }
  const int c = __popc(bits);
  int inc = c;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int o = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += o;
  }
  if (lane == 31) wsum[w] = inc;
  __syncthreads();  // staging buffer and warp totals complete
  int off = inc - c, total = 0;
#pragma unroll
  for (int q = 0; q < SEL_THREADS / 32; ++q) {
    if (q < w) off += wsum[q];
    total += wsum[q];
  }
  // compact IN PLACE: every thread first takes its 16 elements into registers (the rotation (k + t) keeps the
  // 16-element-strided reads off one bank), then -- after a barrier -- drops the selected ones at its offset, which
  // is never beyond its own first element
  T v[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) v[k] = stage[t * 16 + ((k + t) & 15)];
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const int kk = (k + t) & 15;
    if ((bits >> kk) & 1u) stage[off + __popc(bits & ((1u << kk) - 1u))] = v[k];
  }
  __syncthreads();
  const i64 base = offsets[blockIdx.x];
  for (int i = t; i < total; i += SEL_THREADS)
    if (base + i < out_n) out[base + i] = stage[i];
}

// count for a contiguous mask: 16 bytes per thread per load
__global__ void __launch_bounds__(SEL_THREADS) k_sel_count_flat(const uint8_t* __restrict__ m, i64 n, i64* __restrict__ counts) {
  __shared__ int sm[SEL_THREADS / 32];
  const i64 e = (i64)blockIdx.x * SEL_CHUNK + (i64)threadIdx.x * 16;
  int c = 0;
  if (e + 16 <= n && (reinterpret_cast<uintptr_t>(m + e) & 15) == 0) {
    const uint4 v = *reinterpret_cast<const uint4*>(m + e);
    c = (__popc(__vcmpne4(v.x, 0)) + __popc(__vcmpne4(v.y, 0)) + __popc(__vcmpne4(v.z, 0)) + __popc(__vcmpne4(v.w, 0))) >> 3;
  } else {
    for (int k = 0; k < 16; ++k)
      if (e + k < n && m[e + k] != 0) ++c;
  }
  for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int q = 0; q < SEL_THREADS / 32; ++q) t += sm[q];
    counts[blockIdx.x] = t;
  }
}

static bool view_is_flat(const um_view* v) { return (v->cs == 1 || v->cols == 1) && (v->rs == v->cols || v->rows == 1); }

template <class T>
static void sel_scatter(const um_view* a, const um_view* mask, i64 n, const i64* offs, void* o, i64 out_n, unsigned g) {
  cudaStream_t st = ctx()->stream;
  const bool fa = view_is_flat(a), fm = view_is_flat(mask);
  T* op = static_cast<T*>(o);
  if (fa && fm && sizeof(T) >= 4 && aligned16(in_of<uint8_t>(mask).p)) {
    const size_t smem = (size_t)SEL_CHUNK * sizeof(T);
    cudaFuncSetAttribute(k_sel_scatter_staged<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);  // per device; cheap
    k_sel_scatter_staged<T><<<g, SEL_THREADS, smem, st>>>(in_of<T>(a).p, in_of<uint8_t>(mask).p, n, offs, op, out_n);
  } else if (fa && fm) k_sel_scatter_flat<T><<<g, SEL_THREADS, 0, st>>>(in_of<T>(a).p, in_of<uint8_t>(mask).p, n, offs, op, out_n);
  else if (fa) k_sel_scatter<T, true, false><<<g, SEL_THREADS, 0, st>>>(in_of<T>(a), in_of<uint8_t>(mask), a->cols, n, offs, op, out_n);
  else if (fm) k_sel_scatter<T, false, true><<<g, SEL_THREADS, 0, st>>>(in_of<T>(a), in_of<uint8_t>(mask), a->cols, n, offs, op, out_n);
  else k_sel_scatter<T, false, false><<<g, SEL_THREADS, 0, st>>>(in_of<T>(a), in_of<uint8_t>(mask), a->cols, n, offs, op, out_n);
}

// m(mask) is two calls (count, then fill the caller-allocated row).  The block offsets the count produced are still in
// this thread's scratch buffer when the fill follows immediately: remembered with the mask's view and the thread's
// call counter, reused only if no other device-touching call was made by this thread in between.
struct SelCache {
  const void* data = nullptr;
  i64 offset = 0, rows = 0, cols = 0, rs = 0, cs = 0, nb = 0;
  i64* offsets = nullptr;
  unsigned long long epoch = 0;
};
static thread_local SelCache t_sel;
static bool sel_cache_hit(const um_view* mask, unsigned long long epoch_now) {
  return t_sel.offsets && t_sel.data == mask->data && t_sel.offset == mask->offset && t_sel.rows == mask->rows &&
         t_sel.cols == mask->cols && t_sel.rs == mask->rs && t_sel.cs == mask->cs && t_sel.epoch + 1 == epoch_now &&
         t_sel.offsets == ctx()->scratch;
}

static int sel_offsets(const um_view* mask, i64** offsets_out, i64* nb_out) {
  const i64 n = mask->rows * mask->cols;
  const i64 nb = (n + SEL_CHUNK - 1) / SEL_CHUNK;
  const i64 nchunk = (nb + SEL_SCAN_CHUNK - 1) / SEL_SCAN_CHUNK;
  i64* counts = static_cast<i64*>(scratch((size_t)(nb + 1 + nchunk + 1) * sizeof(i64)));
  if (!counts) return UM_ERR_OOM;
  if (nb > 0) {
    if (view_is_flat(mask)) k_sel_count_flat<<<(unsigned)nb, SEL_THREADS, 0, ctx()->stream>>>(in_of<uint8_t>(mask).p, n, counts);
    else k_sel_count<false><<<(unsigned)nb, SEL_THREADS, 0, ctx()->stream>>>(in_of<uint8_t>(mask), mask->cols, n, counts);
    UM_LAUNCHED();
  }
  if (nchunk <= 1) {
    k_sel_scan<<<1, 1024, 0, ctx()->stream>>>(counts, nb, nb > 0 ? nb : 1, counts + nb);
    UM_LAUNCHED();
  } else {  // chunks scanned in parallel, their totals scanned by one block, then added back
    i64* tot = counts + nb + 1;
    k_sel_scan<<<(unsigned)nchunk, 1024, 0, ctx()->stream>>>(counts, nb, SEL_SCAN_CHUNK, tot);
    UM_LAUNCHED();
    k_sel_scan<<<1, 1024, 0, ctx()->stream>>>(tot, nchunk, nchunk, tot + nchunk);
    UM_LAUNCHED();
    k_sel_scan_add<<<(unsigned)nchunk, 1024, 0, ctx()->stream>>>(counts, nb, SEL_SCAN_CHUNK, tot, nchunk);
    UM_LAUNCHED();
  }
  *offsets_out = counts;
  *nb_out = nb;
  return UM_OK;
}

}  // namespace um

using namespace um;

extern "C" {

int32_t um_cumulative(int32_t kind, const um_view* a, int32_t axis, const um_view* out) {
  UM_CTX();
  UM_VIEW(a); UM_VIEW(out);
  UM_REQUIRE(axis == 0 || axis == 1 || axis == -1, "axis must be 0 or 1, got %d", axis);   // Mat.scala:3636,3693,3720
  UM_REQUIRE(a->dtype == out->dtype && (a->dtype == UM_F64 || a->dtype == UM_F32), "cumulative needs f64 or f32 input and a result of the same type");
  if (axis == -1) {
    UM_REQUIRE(out->rows == 1 && out->cols == a->rows * a->cols && (out->cs == 1 || out->cols <= 1), "flat cumulative result must be a contiguous 1 x (rows*cols) row");
  } else {
    UM_REQUIRE(out->rows == a->rows && out->cols == a->cols, "cumulative result must have the operand's shape");
  }
  if (a->rows == 0 || a->cols == 0) {
    // cummax / cummin seed from m(0, j) / m(i, 0): with no such element the reference's apply fails its `require`
    // (Mat.scala:3694-3699); when the outer loop is the empty one it returns an empty matrix
    const bool seeds = kind != CUM_SUM && axis >= 0 && (axis == 0 ? (a->rows == 0 && a->cols > 0) : (a->cols == 0 && a->rows > 0));
    if (seeds) return fail(UM_ERR_ARG, "cummax/cummin of an empty lane");
    return UM_OK;
  }
  return a->dtype == UM_F64 ? cumulative_kind<double>(kind, a, axis, out) : cumulative_kind<float>(kind, a, axis, out);
}

int32_t um_where(const um_view* cond, const um_view* x, const um_view* y, const um_view* out) {
  UM_CTX();
  UM_VIEW(cond); UM_VIEW(x); UM_VIEW(y); UM_VIEW(out);
  UM_REQUIRE(cond->dtype == UM_BOOL, "where: condition must be a boolean matrix");
  UM_REQUIRE(cond->rows == x->rows && cond->cols == x->cols && x->rows == y->rows && x->cols == y->cols,
             "where: shape mismatch: condition=%lldx%lld x=%lldx%lld y=%lldx%lld", (i64)cond->rows, (i64)cond->cols, (i64)x->rows, (i64)x->cols,
             (i64)y->rows, (i64)y->cols);   // Mat.scala:1306-1310
  UM_REQUIRE(x->dtype == y->dtype && x->dtype == out->dtype && out->rows == x->rows && out->cols == x->cols, "where: result must match x and y");
  switch (x->dtype) {
    case UM_F64: return where_typed<double>(cond, x, y, 0, 0, out);
    case UM_F32: return where_typed<float>(cond, x, y, 0, 0, out);
    case UM_I32: return where_typed<int32_t>(cond, x, y, 0, 0, out);
    default: return where_typed<uint8_t>(cond, x, y, 0, 0, out);
  }
}

int32_t um_where_scalar(const um_view* cond, double x, double y, const um_view* out) {
  UM_CTX();
  UM_VIEW(cond); UM_VIEW(out);
  UM_REQUIRE(cond->dtype == UM_BOOL, "where: condition must be a boolean matrix");
  UM_REQUIRE(out->rows == cond->rows && out->cols == cond->cols, "where: result must have the condition's shape");
  switch (out->dtype) {
    case UM_F64: return where_typed<double>(cond, nullptr, nullptr, x, y, out);
    case UM_F32: return where_typed<float>(cond, nullptr, nullptr, x, y, out);
    case UM_I32: return where_typed<int32_t>(cond, nullptr, nullptr, x, y, out);
    default: return where_typed<uint8_t>(cond, nullptr, nullptr, x, y, out);
  }
}

int32_t um_mask_count(const um_view* mask, int64_t* host_count) {
  UM_CTX();
  UM_VIEW(mask);
  UM_REQUIRE(mask->dtype == UM_BOOL && host_count, "mask_count needs a boolean matrix");
  i64* offs = nullptr;
  i64 nb = 0;
  int s = sel_offsets(mask, &offs, &nb);
  if (s != UM_OK) return s;
  i64 total = 0;
  UM_CUDA(cudaMemcpyAsync(&total, offs + nb, sizeof(i64), cudaMemcpyDeviceToHost, ctx()->stream));
  UM_CUDA(cudaStreamSynchronize(ctx()->stream));
  *host_count = total;
  t_sel = SelCache{mask->data, mask->offset, mask->rows, mask->cols, mask->rs, mask->cs, nb, offs, op_epoch()};
  return UM_OK;
}

int32_t um_mask_select(const um_view* a, const um_view* mask, const um_view* out) {
  UM_CTX();
  const unsigned long long epoch_now = op_epoch();
  UM_VIEW(a); UM_VIEW(mask); UM_VIEW(out);
  UM_REQUIRE(mask->dtype == UM_BOOL, "mask select needs a boolean mask");
  UM_REQUIRE(mask->rows == a->rows && mask->cols == a->cols, "Mask shape %lldx%lld must match matrix shape %lldx%lld", (i64)mask->rows,
             (i64)mask->cols, (i64)a->rows, (i64)a->cols);   // Mat.scala:4010-4011
  UM_REQUIRE(out->dtype == a->dtype && (out->rows == 1 || out->cols == 0) && (out->cs == 1 || out->cols <= 1), "mask select result must be a contiguous 1 x k row of the operand's type");
  const i64 n = a->rows * a->cols;
  if (n == 0 || out->cols == 0) return UM_OK;
  i64* offs = nullptr;
  i64 nb = 0;
  if (sel_cache_hit(mask, epoch_now)) {
    offs = t_sel.offsets;
    nb = t_sel.nb;
  } else {
    int s = sel_offsets(mask, &offs, &nb);
    if (s != UM_OK) return s;
  }
  t_sel.offsets = nullptr;
  const unsigned g = (unsigned)nb;
  void* o = static_cast<char*>(out->data) + (size_t)out->offset * dtype_size(out->dtype);
  switch (a->dtype) {
    case UM_F64: sel_scatter<double>(a, mask, n, offs, o, out->cols, g); break;
    case UM_F32: sel_scatter<float>(a, mask, n, offs, o, out->cols, g); break;
    case UM_I32: sel_scatter<int32_t>(a, mask, n, offs, o, out->cols, g); break;
    default: sel_scatter<uint8_t>(a, mask, n, offs, o, out->cols, g); break;
  }
  UM_LAUNCHED();
  return UM_OK;
}

}  // extern "C"
