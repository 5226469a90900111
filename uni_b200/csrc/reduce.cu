// reduce.cu -- axis and whole-matrix reductions over strided views.
//
// Every reduction is a fixed two-level tree: stage 1 reduces (lane, chunk) blocks with
// per-thread sequential accumulation -> warp shuffles -> shared memory, stage 2 combines the
// chunk partials of each lane in chunk order.  The geometry depends only on the shape, so
// results are run-to-run reproducible (no atomics).  HBM-bound: each element is read once,
// also for var/std (shifted sums merged with Chan's formula) -- the reference's two-pass
// form (Mat.scala:719-739) would read twice.
//
// Ordering for min/max/arg*/idx* is java.lang.Double.compare with "earliest wins ties"
// (Mat.scala:470-548,892-895; MatPandasOps.scala:20-64): reduce (total-order key, logical
// index) pairs, which is associative and commutative, so any tree gives the scan's answer.
#include <stdlib.h>

#include "common.cuh"

namespace um {

constexpr int RD_THREADS = 256;
constexpr int RD_UNROLL = 4;

template <class T, int VW>
struct alignas(sizeof(T) * VW) RVec {
  T v[VW];
};

// ---- policies --------------------------------------------------------------------------------
struct SumP {
  struct Local { double s; };
  struct Acc { double s; };
  __device__ static void linit(Local& l) { l.s = 0.0; }
  template <class T> __device__ static void push(Local& l, T x, i64, double) { l.s += (double)x; }
  __device__ static Acc finish(const Local& l, double, double) { return Acc{l.s}; }
  __device__ static Acc identity() { return Acc{0.0}; }
  __device__ static void merge(Acc& a, const Acc& b) { a.s += b.s; }
  __device__ static Acc shfl(const Acc& a, int d) { return Acc{shfl_down_d(a.s, d)}; }
};

struct VarP {
  struct Local { double s1, s2; };
  struct Acc { double n, mean, m2; };
  __device__ static void linit(Local& l) { l.s1 = l.s2 = 0.0; }
  template <class T> __device__ static void push(Local& l, T x, i64, double k) {
    double d = (double)x - k;
    l.s1 += d;
    l.s2 = fma(d, d, l.s2);
  }
  // n = number of elements pushed into this local (the kernels count loop trips, not elements)
  __device__ static Acc finish(const Local& l, double k, double n) {
    if (n == 0.0) return Acc{0.0, 0.0, 0.0};
    double m = l.s1 / n;
    return Acc{n, k + m, l.s2 - l.s1 * m};
  }
  __device__ static Acc identity() { return Acc{0.0, 0.0, 0.0}; }
  __device__ static void merge(Acc& a, const Acc& b) {  // Chan et al. pairwise update
    if (b.n == 0.0) return;
    if (a.n == 0.0) { a = b; return; }
    double n = a.n + b.n;
    double d = b.mean - a.mean;
    a.mean += d * (b.n / n);
    a.m2 += b.m2 + d * d * (a.n * b.n / n);
    a.n = n;
  }
  __device__ static Acc shfl(const Acc& a, int d) {
    return Acc{shfl_down_d(a.n, d), shfl_down_d(a.mean, d), shfl_down_d(a.m2, d)};
  }
};

template <bool MAX>
struct ArgP {
  struct Acc { i64 key, idx; };
  typedef Acc Local;
  __device__ static void linit(Local& l) { l.key = 0; l.idx = -1; }
  template <class T> __device__ static void push(Local& l, T x, i64 idx, double) {
    i64 k = total_key(x);
    bool better = l.idx < 0 || (MAX ? k > l.key : k < l.key);  // strict: first occurrence stays
    if (better) { l.key = k; l.idx = idx; }
  }
  __device__ static Acc finish(const Local& l, double, double) { return l; }
  __device__ static Acc identity() { return Acc{0, -1}; }
  __device__ static void merge(Acc& a, const Acc& b) {
    if (b.idx < 0) return;
    bool take = a.idx < 0 || (MAX ? b.key > a.key : b.key < a.key) || (b.key == a.key && b.idx < a.idx);
    if (take) a = b;
  }
  __device__ static Acc shfl(const Acc& a, int d) { return Acc{shfl_down_l(a.key, d), shfl_down_l(a.idx, d)}; }
};

// block-wide merge in a fixed order: lanes of a warp by shuffle tree, then warps in order.
template <class P>
__device__ __forceinline__ typename P::Acc block_merge(typename P::Acc a, typename P::Acc* sm) {
  for (int d = 16; d > 0; d >>= 1) {
    typename P::Acc o = P::shfl(a, d);
    P::merge(a, o);
  }
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int nwarps = (blockDim.x * blockDim.y + 31) >> 5;
  if (lane == 0) sm[warp] = a;
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < nwarps; ++w) P::merge(a, sm[w]);
  }
  return a;  // valid in thread 0
}

// ---- stage 1, reduction along the CONTIGUOUS dimension ------------------------------------------
// lane l, chunk c: elements p[l*s_lane + i], i in [c*chunk_len, min(n_red, (c+1)*chunk_len)).
// logical index of element i of lane l = l*ils + i*irs (for arg policies).
template <class P, class T, int VW, bool WARP_PER_LANE>
__global__ void __launch_bounds__(RD_THREADS) k_reduce_contig(const T* __restrict__ p, i64 s_lane, i64 n_lane, i64 n_red,
                                                               i64 chunk_len, int chunks, i64 ils, i64 irs,
                                                               typename P::Acc* __restrict__ partials) {
  __shared__ typename P::Acc sm[RD_THREADS / 32];
  const int tid = threadIdx.x;
  i64 lane;
  int nthr, t;
  if (WARP_PER_LANE) {
    lane = (i64)blockIdx.x * (RD_THREADS / 32) + (tid >> 5);
    nthr = 32;
    t = tid & 31;
  } else {
    lane = blockIdx.x;
    nthr = RD_THREADS;
    t = tid;
  }
  const int chunk = blockIdx.y;
  const bool active = lane < n_lane;
  typename P::Acc acc = P::identity();
  if (active) {
    const T* base = p + lane * s_lane;
    const double k = (double)base[0];
    const i64 lo = (i64)chunk * chunk_len;
    i64 hi = lo + chunk_len;
    if (hi > n_red) hi = n_red;
    const i64 ibase = lane * ils;
    typename P::Local loc[RD_UNROLL];
#pragma unroll
    for (int u = 0; u < RD_UNROLL; ++u) P::linit(loc[u]);
    const i64 stride = (i64)nthr * VW;
    i64 i = lo + (i64)t * VW;
    int trips = 0, tail = 0;  // loc[u] received trips*VW elements, loc[0] `tail` more
    if (VW > 1) {
      typedef RVec<T, VW> V;
      for (; i + (RD_UNROLL - 1) * stride + VW <= hi; i += RD_UNROLL * stride) {
        ++trips;
        V x[RD_UNROLL];
#pragma unroll
        for (int u = 0; u < RD_UNROLL; ++u) x[u] = *reinterpret_cast<const V*>(base + i + u * stride);
#pragma unroll
        for (int u = 0; u < RD_UNROLL; ++u)
#pragma unroll
          for (int j = 0; j < VW; ++j) P::push(loc[u], x[u].v[j], ibase + (i + u * stride + j) * irs, k);
      }
      for (; i < hi; i += stride) {
        if (i + VW <= hi) {
          V x = *reinterpret_cast<const V*>(base + i);
          tail += VW;
#pragma unroll
          for (int j = 0; j < VW; ++j) P::push(loc[0], x.v[j], ibase + (i + j) * irs, k);
        } else {
          for (i64 ii = i; ii < hi; ++ii) { P::push(loc[0], base[ii], ibase + ii * irs, k); ++tail; }
        }
      }
    } else {
      for (; i + (RD_UNROLL - 1) * stride < hi; i += RD_UNROLL * stride) {
        ++trips;
        T x[RD_UNROLL];
#pragma unroll
        for (int u = 0; u < RD_UNROLL; ++u) x[u] = base[i + u * stride];
#pragma unroll
        for (int u = 0; u < RD_UNROLL; ++u) P::push(loc[u], x[u], ibase + (i + u * stride) * irs, k);
      }
      for (; i < hi; i += stride) { P::push(loc[0], base[i], ibase + i * irs, k); ++tail; }
    }
    acc = P::finish(loc[0], k, (double)(trips * VW + tail));
#pragma unroll
    for (int u = 1; u < RD_UNROLL; ++u) {
      typename P::Acc o = P::finish(loc[u], k, (double)(trips * VW));
      P::merge(acc, o);
    }
  }
  if (WARP_PER_LANE) {
    for (int d = 16; d > 0; d >>= 1) {
      typename P::Acc o = P::shfl(acc, d);
      P::merge(acc, o);
    }
    if (active && t == 0) partials[(i64)chunk * n_lane + lane] = acc;
  } else {
    acc = block_merge<P>(acc, sm);
    if (tid == 0) partials[(i64)chunk * n_lane + lane] = acc;
  }
}

// ---- stage 1, reduction along the STRIDED dimension (lanes contiguous) ---------------------------
// block (32, 8): x -> VW adjacent lanes, y -> row phase.  elements p[i*s_red + l].
template <class P, class T, int VW>
__global__ void __launch_bounds__(RD_THREADS) k_reduce_strided(const T* __restrict__ p, i64 s_red, i64 n_lane, i64 n_red,
                                                                i64 chunk_len, int chunks, i64 ils, i64 irs,
                                                                typename P::Acc* __restrict__ partials) {
  __shared__ typename P::Acc sm[8][32 * VW];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const i64 l0 = ((i64)blockIdx.x * 32 + tx) * VW;
  const int chunk = blockIdx.y;
  const i64 lo = (i64)chunk * chunk_len;
  i64 hi = lo + chunk_len;
  if (hi > n_red) hi = n_red;
  typename P::Acc acc[VW];
#pragma unroll
  for (int j = 0; j < VW; ++j) acc[j] = P::identity();
  if (l0 < n_lane) {  // VW > 1 only when n_lane % VW == 0, so all VW lanes are valid
    typedef RVec<T, VW> V;
    double k[VW];
#pragma unroll
    for (int j = 0; j < VW; ++j) k[j] = (double)p[l0 + j];
    typename P::Local loc[RD_UNROLL][VW];
#pragma unroll
    for (int u = 0; u < RD_UNROLL; ++u)
#pragma unroll
      for (int j = 0; j < VW; ++j) P::linit(loc[u][j]);
    i64 i = lo + ty;
    int trips = 0, tail = 0;
    for (; i + (RD_UNROLL - 1) * 8 < hi; i += RD_UNROLL * 8) {
      ++trips;
      V x[RD_UNROLL];
#pragma unroll
      for (int u = 0; u < RD_UNROLL; ++u) x[u] = *reinterpret_cast<const V*>(p + (i + u * 8) * s_red + l0);
#pragma unroll
      for (int u = 0; u < RD_UNROLL; ++u)
#pragma unroll
        for (int j = 0; j < VW; ++j) P::push(loc[u][j], x[u].v[j], (l0 + j) * ils + (i + u * 8) * irs, k[j]);
    }
    for (; i < hi; i += 8) {
      ++tail;
      V x = *reinterpret_cast<const V*>(p + i * s_red + l0);
#pragma unroll
      for (int j = 0; j < VW; ++j) P::push(loc[0][j], x.v[j], (l0 + j) * ils + i * irs, k[j]);
    }
#pragma unroll
    for (int j = 0; j < VW; ++j) {
      acc[j] = P::finish(loc[0][j], k[j], (double)(trips + tail));
#pragma unroll
      for (int u = 1; u < RD_UNROLL; ++u) {
        typename P::Acc o = P::finish(loc[u][j], k[j], (double)trips);
        P::merge(acc[j], o);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < VW; ++j) sm[ty][tx * VW + j] = acc[j];
  __syncthreads();
  if (ty == 0 && l0 < n_lane) {
#pragma unroll
    for (int j = 0; j < VW; ++j) {
      typename P::Acc a = sm[0][tx * VW + j];
      for (int y = 1; y < 8; ++y) P::merge(a, sm[y][tx * VW + j]);
      partials[(i64)chunk * n_lane + l0 + j] = a;
    }
  }
}

// ---- stage 1, generic strides: one thread per lane, sequential (slow path) ------------------------
template <class P, class T>
__global__ void __launch_bounds__(RD_THREADS) k_reduce_generic(const T* __restrict__ p, i64 s_lane, i64 s_red, i64 n_lane,
                                                                i64 n_red, i64 ils, i64 irs,
                                                                typename P::Acc* __restrict__ partials) {
  i64 lane = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (lane >= n_lane) return;
  const T* base = p + lane * s_lane;
  double k = (double)base[0];
  typename P::Local loc;
  P::linit(loc);
  for (i64 i = 0; i < n_red; ++i) P::push(loc, base[i * s_red], lane * ils + i * irs, k);
  partials[lane] = P::finish(loc, k, (double)n_red);
}

// ---- stage 2 ---------------------------------------------------------------------------------
enum Fin { FIN_SUM, FIN_MEAN, FIN_VAR, FIN_STD, FIN_VALUE, FIN_INDEX, FIN_ALL, FIN_ANY };

template <class P> struct FinOf;
template <> struct FinOf<SumP> {
  template <class T, class TO>
  __device__ static TO fin(int kind, const SumP::Acc& a, i64 n_red, const T*, i64, i64, i64) {
    if (kind == FIN_MEAN) return (TO)(a.s / (double)n_red);
    if (kind == FIN_ALL) return (TO)(a.s == (double)n_red);
    if (kind == FIN_ANY) return (TO)(a.s > 0.0);
    return (TO)a.s;
  }
};
template <> struct FinOf<VarP> {
  template <class T, class TO>
  __device__ static TO fin(int kind, const VarP::Acc& a, i64 n_red, const T*, i64, i64, i64) {
    double v = a.m2 / (double)n_red;
    if (v < 0.0) v = 0.0;  // rounding in the shifted form; a NaN stays NaN
    return (TO)(kind == FIN_STD ? sqrt(v) : v);
  }
};
template <bool MAX> struct FinOf<ArgP<MAX>> {
  template <class T, class TO>
  __device__ static TO fin(int kind, const typename ArgP<MAX>::Acc& a, i64, const T* p, i64 lane, i64 s_lane, i64 s_red) {
    // per-lane idx is the position inside the lane (ils = 0, irs = 1)
    if (kind == FIN_INDEX) return (TO)a.idx;
    return (TO)p[lane * s_lane + a.idx * s_red];
  }
};

template <class P, class T, class TO>
__global__ void __launch_bounds__(RD_THREADS) k_finalize_lanes(const typename P::Acc* __restrict__ partials, int chunks,
                                                                i64 n_lane, i64 n_red, int kind, const T* __restrict__ p,
                                                                i64 s_lane, i64 s_red, TO* __restrict__ out, i64 out_stride) {
  i64 lane = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (lane >= n_lane) return;
  typename P::Acc a = partials[lane];
  for (int c = 1; c < chunks; ++c) P::merge(a, partials[(i64)c * n_lane + lane]);
  out[lane * out_stride] = FinOf<P>::template fin<T, TO>(kind, a, n_red, p, lane, s_lane, s_red);
}

// whole-matrix: merge ALL partials (count of them) into one Acc at result[0]
template <class P>
__global__ void __launch_bounds__(1024) k_merge_all(const typename P::Acc* __restrict__ partials, i64 count,
                                                     typename P::Acc* __restrict__ result) {
  __shared__ typename P::Acc sm[32];
  typename P::Acc a = P::identity();
  for (i64 i = threadIdx.x; i < count; i += blockDim.x) P::merge(a, partials[i]);
  a = block_merge<P>(a, sm);
  if (threadIdx.x == 0) result[0] = a;
}

// ---- host-side planning ------------------------------------------------------------------------
struct Plan {
  const void* p;
  i64 n_lane, n_red, s_lane, s_red, ils, irs;
  int mode;  // 0 contiguous reduction, 1 strided reduction, 2 generic
};

static i64 ceil_div(i64 a, i64 b) { return (a + b - 1) / b; }

// runs stage 1; returns partial count per lane (chunks) through *chunks_out; partials in scratch
template <class P, class T>
static int stage1(const Plan& pl, typename P::Acc** partials_out, int* chunks_out, size_t extra_accs) {
  typedef typename P::Acc Acc;
  Ctx& c = *ctx();
  const T* p = static_cast<const T*>(pl.p);
  const i64 target = (i64)c.sm_count * 16;
  cudaStream_t st = c.stream;
  if (pl.mode == 0) {
    bool warp = pl.n_red <= 1024;
    i64 nblk = warp ? ceil_div(pl.n_lane, RD_THREADS / 32) : pl.n_lane;
    i64 chunks = 1;
    if (!warp && nblk < target) {
      chunks = ceil_div(target, nblk);
      i64 maxc = ceil_div(pl.n_red, 8192);
      if (chunks > maxc) chunks = maxc;
      if (chunks > 4096) chunks = 4096;
      if (chunks < 1) chunks = 1;
    }
    constexpr int VW = 16 / sizeof(T);
    i64 chunk_len = ceil_div(pl.n_red, chunks);
    chunk_len = ceil_div(chunk_len, VW) * VW;  // keep chunk starts vector-aligned
    chunks = ceil_div(pl.n_red, chunk_len);
    Acc* part = static_cast<Acc*>(scratch(sizeof(Acc) * (size_t)(chunks * pl.n_lane + extra_accs)));
    if (!part) return UM_ERR_OOM;
    bool vec = aligned16(p) && (pl.n_lane == 1 || pl.s_lane % VW == 0);
    dim3 grid((unsigned)nblk, (unsigned)chunks);
    if (warp) {
      if (vec) k_reduce_contig<P, T, VW, true><<<grid, RD_THREADS, 0, st>>>(p, pl.s_lane, pl.n_lane, pl.n_red, chunk_len, (int)chunks, pl.ils, pl.irs, part);
      else k_reduce_contig<P, T, 1, true><<<grid, RD_THREADS, 0, st>>>(p, pl.s_lane, pl.n_lane, pl.n_red, chunk_len, (int)chunks, pl.ils, pl.irs, part);
    } else {
      if (vec) k_reduce_contig<P, T, VW, false><<<grid, RD_THREADS, 0, st>>>(p, pl.s_lane, pl.n_lane, pl.n_red, chunk_len, (int)chunks, pl.ils, pl.irs, part);
      else k_reduce_contig<P, T, 1, false><<<grid, RD_THREADS, 0, st>>>(p, pl.s_lane, pl.n_lane, pl.n_red, chunk_len, (int)chunks, pl.ils, pl.irs, part);
    }
    UM_LAUNCHED();
    *partials_out = part;
    *chunks_out = (int)chunks;
    return UM_OK;
  }
  if (pl.mode == 1) {
    constexpr int VWMAX = sizeof(T) == 8 ? 2 : 1;  // 128-bit lanes for f64 only
    bool vec = VWMAX > 1 && aligned16(p) && (pl.n_lane % VWMAX == 0) && (pl.s_red % VWMAX == 0);
    int vw = vec ? VWMAX : 1;
    i64 tiles = ceil_div(pl.n_lane, 32 * vw);
    i64 chunks = 1;
    if (tiles < target) {
      chunks = ceil_div(target, tiles);
      i64 maxc = ceil_div(pl.n_red, 64);
      if (chunks > maxc) chunks = maxc;
      if (chunks > 4096) chunks = 4096;
      if (chunks < 1) chunks = 1;
    }
    i64 chunk_len = ceil_div(pl.n_red, chunks);
    chunks = ceil_div(pl.n_red, chunk_len);
    Acc* part = static_cast<Acc*>(scratch(sizeof(Acc) * (size_t)(chunks * pl.n_lane + extra_accs)));
    if (!part) return UM_ERR_OOM;
    dim3 grid((unsigned)tiles, (unsigned)chunks), block(32, 8);
    if (vec) k_reduce_strided<P, T, VWMAX><<<grid, block, 0, st>>>(p, pl.s_red, pl.n_lane, pl.n_red, chunk_len, (int)chunks, pl.ils, pl.irs, part);
    else k_reduce_strided<P, T, 1><<<grid, block, 0, st>>>(p, pl.s_red, pl.n_lane, pl.n_red, chunk_len, (int)chunks, pl.ils, pl.irs, part);
    UM_LAUNCHED();
    *partials_out = part;
    *chunks_out = (int)chunks;
    return UM_OK;
  }
  Acc* part = static_cast<Acc*>(scratch(sizeof(Acc) * (size_t)(pl.n_lane + extra_accs)));
  if (!part) return UM_ERR_OOM;
  k_reduce_generic<P, T><<<(unsigned)ceil_div(pl.n_lane, RD_THREADS), RD_THREADS, 0, st>>>(p, pl.s_lane, pl.s_red, pl.n_lane, pl.n_red, pl.ils, pl.irs, part);
  UM_LAUNCHED();
  *partials_out = part;
  *chunks_out = 1;
  return UM_OK;
}

template <class T>
static Plan plan_axis(const um_view* a, int axis) {
  Plan pl;
  pl.p = static_cast<const T*>(a->data) + a->offset;
  if (axis == 0) { pl.n_lane = a->cols; pl.n_red = a->rows; pl.s_lane = a->cs; pl.s_red = a->rs; }
  else { pl.n_lane = a->rows; pl.n_red = a->cols; pl.s_lane = a->rs; pl.s_red = a->cs; }
  pl.ils = 0;
  pl.irs = 1;
  if (pl.s_red == 1 || pl.n_red == 1) { pl.mode = 0; if (pl.n_red == 1) pl.s_red = 1; }
  else if (pl.s_lane == 1 || pl.n_lane == 1) { pl.mode = 1; if (pl.n_lane == 1) pl.s_lane = 1; }
  else pl.mode = 2;
  return pl;
}

// whole matrix: pick the reduction direction that is contiguous in memory; logical row-major
// index of (r, c) is r*cols + c.
template <class T>
static Plan plan_all(const um_view* a) {
  Plan pl;
  pl.p = static_cast<const T*>(a->data) + a->offset;
  const i64 rows = a->rows, cols = a->cols;
  if (a->cs == 1 && (a->rs == cols || rows == 1)) {  // dense: one long lane
    pl.n_lane = 1; pl.n_red = rows * cols; pl.s_lane = 0; pl.s_red = 1; pl.ils = 0; pl.irs = 1; pl.mode = 0;
  } else if (a->cs == 1 || cols == 1) {              // row-contiguous view: lanes = rows
    pl.n_lane = rows; pl.n_red = cols; pl.s_lane = a->rs; pl.s_red = 1; pl.ils = cols; pl.irs = 1; pl.mode = 0;
  } else if (a->rs == 1 || rows == 1) {              // transposed view: lanes = logical columns
    pl.n_lane = cols; pl.n_red = rows; pl.s_lane = a->cs; pl.s_red = 1; pl.ils = 1; pl.irs = cols; pl.mode = 0;
  } else {
    pl.n_lane = rows; pl.n_red = cols; pl.s_lane = a->rs; pl.s_red = a->cs; pl.ils = cols; pl.irs = 1; pl.mode = 2;
  }
  return pl;
}

template <class P, class T>
static int reduce_all_acc(const um_view* a, typename P::Acc* host_acc) {
  typedef typename P::Acc Acc;
  Plan pl = plan_all<T>(a);
  Acc* part = nullptr;
  int chunks = 0;
  int s = stage1<P, T>(pl, &part, &chunks, 1);
  if (s != UM_OK) return s;
  i64 count = (i64)chunks * pl.n_lane;
  Acc* result = part + count;
  k_merge_all<P><<<1, 1024, 0, ctx()->stream>>>(part, count, result);
  UM_LAUNCHED();
  UM_CUDA(cudaMemcpyAsync(host_acc, result, sizeof(Acc), cudaMemcpyDeviceToHost, ctx()->stream));
  UM_CUDA(cudaStreamSynchronize(ctx()->stream));
  return UM_OK;
}

template <class P, class T, class TO>
static int reduce_axis_run(const um_view* a, int axis, int kind, const um_view* out) {
  typedef typename P::Acc Acc;
  Plan pl = plan_axis<T>(a, axis);
  Acc* part = nullptr;
  int chunks = 0;
  int s = stage1<P, T>(pl, &part, &chunks, 0);
  if (s != UM_OK) return s;
  TO* o = static_cast<TO*>(out->data) + out->offset;
  i64 ostride = axis == 0 ? out->cs : out->rs;
  k_finalize_lanes<P, T, TO><<<(unsigned)ceil_div(pl.n_lane, RD_THREADS), RD_THREADS, 0, ctx()->stream>>>(
      part, chunks, pl.n_lane, pl.n_red, kind, static_cast<const T*>(pl.p), pl.s_lane, pl.s_red, o, ostride);
  UM_LAUNCHED();
  return UM_OK;
}

template <class T>
static int reduce_axis_typed(int kind, const um_view* a, int axis, const um_view* out) {
  if (out->dtype != a->dtype) return fail(UM_ERR_ARG, "reduce(axis): output dtype must match input");
  switch (kind) {
    case UM_SUM: return reduce_axis_run<SumP, T, T>(a, axis, FIN_SUM, out);
    case UM_MEAN: return reduce_axis_run<SumP, T, T>(a, axis, FIN_MEAN, out);
    case UM_VAR: return reduce_axis_run<VarP, T, T>(a, axis, FIN_VAR, out);
    case UM_STD: return reduce_axis_run<VarP, T, T>(a, axis, FIN_STD, out);
    case UM_MIN: return reduce_axis_run<ArgP<false>, T, T>(a, axis, FIN_VALUE, out);
    case UM_MAX: return reduce_axis_run<ArgP<true>, T, T>(a, axis, FIN_VALUE, out);
  }
  return fail(UM_ERR_ARG, "unknown reduction %d", kind);
}

static int check_axis_out(const um_view* a, int axis, const um_view* out) {
  if (axis != 0 && axis != 1) return fail(UM_ERR_ARG, "axis must be 0 or 1, got %d", axis);
  i64 er = axis == 0 ? 1 : a->rows, ec = axis == 0 ? a->cols : 1;
  if (out->rows != er || out->cols != ec)
    return fail(UM_ERR_ARG, "reduce(axis=%d) of %lldx%lld needs a %lldx%lld output, got %lldx%lld", axis, (i64)a->rows,
                (i64)a->cols, er, ec, (i64)out->rows, (i64)out->cols);
  return UM_OK;
}

template <class T>
static int reduce_all_typed(int kind, const um_view* a, double* host_out) {
  const i64 n = a->rows * a->cols;
  if (kind == UM_SUM || kind == UM_MEAN) {
    SumP::Acc r;
    int s = reduce_all_acc<SumP, T>(a, &r);
    if (s != UM_OK) return s;
    *host_out = kind == UM_MEAN ? r.s / (double)n : r.s;
    return UM_OK;
  }
  if (kind == UM_VAR || kind == UM_STD) {
    VarP::Acc r;
    int s = reduce_all_acc<VarP, T>(a, &r);
    if (s != UM_OK) return s;
    double v = r.m2 / (double)n;
    if (v < 0.0) v = 0.0;
    *host_out = kind == UM_STD ? sqrt(v) : v;
    return UM_OK;
  }
  return fail(UM_ERR_ARG, "unknown reduction %d", kind);
}

template <class T>
static int arg_all_typed(int want_max, const um_view* a, i64* idx_out) {
  if (want_max) {
    ArgP<true>::Acc r;
    int s = reduce_all_acc<ArgP<true>, T>(a, &r);
    if (s != UM_OK) return s;
    *idx_out = r.idx;
  } else {
    ArgP<false>::Acc r;
    int s = reduce_all_acc<ArgP<false>, T>(a, &r);
    if (s != UM_OK) return s;
    *idx_out = r.idx;
  }
  return UM_OK;
}


// =====================================================================================================
// softmax(axis) (MatMathOps.scala:156-226): per lane  mx = max, e = exp(x - mx), e / sum(e).
//  * lanes contiguous in memory (axis 1 of a row-major matrix): the whole lane is held in
//    REGISTERS across a thread-block cluster (<= 8 CTAs x 512 threads x 16 doubles), max and sum
//    exchanged through distributed shared memory -> one HBM read + one write per element.
//  * lanes strided (axis 0): online (max, sum) reduction pass + normalise pass (two reads).
// NaN / Inf follow the reference's arithmetic: any NaN, any +Inf, or an all -Inf lane gives an
// all-NaN lane (exp(x - mx) produces the NaN that poisons the sum).
// =====================================================================================================
struct SoftP {
  struct Acc { double m, s; };
  typedef Acc Local;
  __device__ static void linit(Local& l) { l.m = -INFINITY; l.s = 0.0; }
  template <class T> __device__ static void push(Local& l, T xi, i64, double) {
    double x = (double)xi;
    if (x > l.m) {
      l.s = l.s * fast_exp(l.m - x) + 1.0;   // exp(-inf) = 0 on the first element
      l.m = x;
      if (x == INFINITY) l.s = NAN;     // reference: inf - inf = NaN poisons the lane
    } else if (x == -INFINITY) {
      // contributes exp(-inf) = 0 against any finite max; an all -inf lane ends with s = 0 -> NaN
    } else {
      l.s += fast_exp(x - l.m);         // NaN input -> NaN sum
    }
  }
  __device__ static Acc finish(const Local& l, double, double) { return l; }
  __device__ static Acc identity() { return Acc{-INFINITY, 0.0}; }
  __device__ static void merge(Acc& a, const Acc& b) {
    if (b.m == -INFINITY && b.s == 0.0) return;
    if (a.m == -INFINITY && a.s == 0.0) { a = b; return; }
    double M = a.m > b.m ? a.m : b.m;
    a.s = a.s * exp(a.m - M) + b.s * exp(b.m - M);
    a.m = M;
  }
  __device__ static Acc shfl(const Acc& a, int d) { return Acc{shfl_down_d(a.m, d), shfl_down_d(a.s, d)}; }
};

// Branch-free variant for the streaming pass: shift by the lane's first element k instead of a
// running max, s = sum exp(x - k), mx = max x.  Exact unless exp(x - k) overflows; that (mx - k >
// 600, or a non-finite k) is flagged with s = -1 and the flagged lane is redone with the online
// form by k_softmax_lane_stats.
struct SoftKP {
  struct Local { double s, mx; };
  typedef SoftP::Acc Acc;
  __device__ static void linit(Local& l) { l.s = 0.0; l.mx = -INFINITY; }
  template <class T> __device__ static void push(Local& l, T xi, i64, double k) {
    double x = (double)xi;
    l.s += fast_exp(x - k);
    l.mx = fmax(l.mx, x);
  }
  __device__ static Acc finish(const Local& l, double k, double n) {
    if (n == 0.0) return Acc{-INFINITY, 0.0};
    double d = l.mx - k;
    // |d| > 600: exp(x - k) may have overflowed (d >> 0) or the whole chunk underflowed to s = 0 against a huge k
    // (d << 0: 0 * exp(k - mx) = 0 * Inf would poison the merge) -- not a plain NaN lane: redo it with the online form
    if (!(d <= 600.0 && d >= -600.0) && !(l.s != l.s && d == d)) return Acc{0.0, -1.0};
    if (!(fabs(k) <= 1.7e308)) return Acc{0.0, -1.0};
    return Acc{l.mx, l.s * fast_exp(k - l.mx)};   // re-base on the local max
  }
  __device__ static Acc identity() { return Acc{-INFINITY, 0.0}; }
  __device__ static void merge(Acc& a, const Acc& b) {
    if (a.s < 0.0) return;
    if (b.s < 0.0) { a = b; return; }
    SoftP::merge(a, b);
  }
  __device__ static Acc shfl(const Acc& a, int d) { return SoftP::shfl(a, d); }
};

// normalise pass for strided lanes: out[i][l] = exp(x[i][l] - m[l]) / s[l]
template <class T>
__global__ void __launch_bounds__(256) k_softmax_norm(const T* __restrict__ p, i64 s_lane, i64 s_red, i64 n_lane, i64 n_red,
                                                       const SoftP::Acc* __restrict__ part, int chunks,
                                                       double* __restrict__ out, i64 o_lane, i64 o_red) {
  i64 l = (i64)blockIdx.x * 32 + threadIdx.x;
  if (l >= n_lane) return;
  SoftP::Acc a = part[l];
  for (int c = 1; c < chunks; ++c) SoftP::merge(a, part[(i64)c * n_lane + l]);
  const double inv = 1.0 / a.s;
  for (i64 i = (i64)blockIdx.y * 8 + threadIdx.y; i < n_red; i += (i64)gridDim.y * 8) {
    double x = (double)p[l * s_lane + i * s_red];
    out[l * o_lane + i * o_red] = fast_exp(x - a.m) * inv;
  }
}

// lane statistics (m, 1/s) from the chunk partials, one thread per lane
template <class T>
__global__ void __launch_bounds__(256) k_softmax_lane_stats(const SoftP::Acc* __restrict__ part, int chunks, i64 n_lane,
                                                             const T* __restrict__ p, i64 s_red, i64 n_red,
                                                             double* __restrict__ m_out, double* __restrict__ inv_out) {
  i64 l = (i64)blockIdx.x * 256 + threadIdx.x;
  if (l >= n_lane) return;
  SoftP::Acc a = part[l];
  for (int c = 1; c < chunks; ++c) SoftKP::merge(a, part[(i64)c * n_lane + l]);
  if (a.s < 0.0) {  // flagged: the shifted sum could overflow -> online form, this thread alone (rare)
    SoftP::Local loc;
    SoftP::linit(loc);
    for (i64 i = 0; i < n_red; ++i) SoftP::push(loc, p[i * s_red + l], 0, 0.0);
    a = loc;
  }
  m_out[l] = a.m;
  inv_out[l] = 1.0 / a.s;
}

// vectorised normalise pass for row-major f64 input with unit-stride lanes (axis 0 of a
// row-major matrix): block (32, 8), each thread 2 adjacent lanes x 4 rows in flight.
__global__ void __launch_bounds__(256) k_softmax_norm_vec(const double* __restrict__ p, i64 s_red, i64 n_lane, i64 n_red,
                                                           const double* __restrict__ m_in, const double* __restrict__ inv_in,
                                                           double* __restrict__ out, i64 o_red) {
  __shared__ double s_tab[32];   // 2^(j/32) for exp_tab_nonpos (x - m <= 0 here)
  if (threadIdx.y == 0) s_tab[threadIdx.x] = k_exp2_32[threadIdx.x];
  __syncthreads();
  const unsigned tab = (unsigned)__cvta_generic_to_shared(s_tab);
  const i64 l = ((i64)blockIdx.x * 32 + threadIdx.x) * 2;
  if (l >= n_lane) return;  // n_lane is even on this path
  const double2 m = *reinterpret_cast<const double2*>(m_in + l);
  const double2 inv = *reinterpret_cast<const double2*>(inv_in + l);
  const i64 step = (i64)gridDim.y * 8;
  i64 i = (i64)blockIdx.y * 8 + threadIdx.y;
  for (; i + 3 * step < n_red; i += 4 * step) {
    double2 x[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) x[u] = *reinterpret_cast<const double2*>(p + (i + u * step) * s_red + l);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      double2 o = make_double2(exp_tab_nonpos(x[u].x - m.x, tab) * inv.x, exp_tab_nonpos(x[u].y - m.y, tab) * inv.y);
      __stcs(reinterpret_cast<double2*>(out + (i + u * step) * o_red + l), o);
    }
  }
  for (; i < n_red; i += step) {
    double2 x = *reinterpret_cast<const double2*>(p + i * s_red + l);
    double2 o = make_double2(exp_tab_nonpos(x.x - m.x, tab) * inv.x, exp_tab_nonpos(x.y - m.y, tab) * inv.y);
    __stcs(reinterpret_cast<double2*>(out + i * o_red + l), o);
  }
}

// Long contiguous lanes (a 256 KiB row does not fit one CTA's registers): ONE CTA of 512 threads
// per lane (two resident per SM), two passes.  Pass A folds an online (max, sum) over the lane from HBM; pass B re-reads
// the lane -- 296 resident CTAs x 256 KiB = 76 MB, so it comes back from the 126 MB L2, not HBM --
// and writes exp(x - m) / s with streaming stores.  DRAM traffic stays one read + one write.
template <class T, int VW>
__global__ void __launch_bounds__(512, 2) k_softmax_row2pass(const T* __restrict__ p, i64 s_lane, i64 n_red,
                                                               double* __restrict__ out, i64 o_lane) {
  __shared__ SoftP::Acc sm[32];
  __shared__ double bc[2];
  typedef RVec<T, VW> V;
  typedef RVec<double, VW> VO;
  const int t = threadIdx.x;
  const T* base = p + (i64)blockIdx.x * s_lane;
  double* obase = out + (i64)blockIdx.x * o_lane;
  constexpr int U = 8;                 // 8 x 16 B in flight per thread, 2 CTAs per SM
  const i64 stride = 512 * (i64)VW;
  const i64 nvec = (n_red / VW) * VW;  // elements covered by whole vectors
  // ---- pass A: online max / sum -------------------------------------------------------------
  double m = -INFINITY, s = 0.0;
  auto fold = [&](const double* x, int n) {
    double bm = x[0];
    for (int q = 1; q < n; ++q) bm = fmax(bm, x[q]);   // fmax drops NaNs; the exp below keeps them
    if (bm > m) {
      s *= fast_exp(m - bm);   // exp(-inf) = 0 the first time
      m = bm;
    }
    for (int q = 0; q < n; ++q) {
      // x = -inf against a finite max contributes 0; (+inf) - (+inf) and NaN give NaN, as the reference
      double e = (x[q] == -INFINITY && m == -INFINITY) ? 0.0 : fast_exp(x[q] - m);
      s += e;
    }
  };
  i64 i = (i64)t * VW;
  for (; i + (U - 1) * stride + VW <= nvec; i += U * stride) {
    V v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = *reinterpret_cast<const V*>(base + i + u * stride);
    double x[U * VW];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int q = 0; q < VW; ++q) x[u * VW + q] = (double)v[u].v[q];
    fold(x, U * VW);
  }
  for (; i + VW <= nvec; i += stride) {
    V v = *reinterpret_cast<const V*>(base + i);
    double x[VW];
#pragma unroll
    for (int q = 0; q < VW; ++q) x[q] = (double)v.v[q];
    fold(x, VW);
  }
  if (t == 0)
    for (i64 ii = nvec; ii < n_red; ++ii) {
      double x[1] = {(double)base[ii]};
      fold(x, 1);
    }
  SoftP::Acc a{m, s};
  if (m == INFINITY) a.s = NAN;  // +inf in the lane: inf - inf poisons the reference's sum
  a = block_merge<SoftP>(a, sm);
  if (t == 0) {
    bc[0] = a.m;
    bc[1] = 1.0 / a.s;
  }
  __syncthreads();
  const double mx = bc[0], inv = bc[1];
  // ---- pass B: normalise (lane re-read from L2) ---------------------------------------------------
  i = (i64)t * VW;
  for (; i + (U - 1) * stride + VW <= nvec; i += U * stride) {
    V v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = *reinterpret_cast<const V*>(base + i + u * stride);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      VO o;
#pragma unroll
      for (int q = 0; q < VW; ++q) o.v[q] = fast_exp((double)v[u].v[q] - mx) * inv;
      if (VW == 2) __stcs(reinterpret_cast<double2*>(obase + i + u * stride), make_double2(o.v[0], o.v[VW - 1]));
      else *reinterpret_cast<VO*>(obase + i + u * stride) = o;
    }
  }
  for (; i + VW <= nvec; i += stride) {
    V v = *reinterpret_cast<const V*>(base + i);
    VO o;
#pragma unroll
    for (int q = 0; q < VW; ++q) o.v[q] = fast_exp((double)v.v[q] - mx) * inv;
    *reinterpret_cast<VO*>(obase + i) = o;
  }
  if (t == 0)
    for (i64 ii = nvec; ii < n_red; ++ii) obase[ii] = fast_exp((double)base[ii] - mx) * inv;
}

// register-resident lane softmax for lanes up to 4096 elements: one CTA per lane, thread t holds
// vectors j*THREADS + t; max and sum by shuffle + shared memory in a fixed order.
template <class T, int NV, int THREADS, int VW>
__global__ void __launch_bounds__(THREADS) k_softmax_lane(const T* __restrict__ p, i64 s_lane, i64 n_red,
                                                           double* __restrict__ out, i64 o_lane) {
  __shared__ double sm_red[THREADS / 32];
  __shared__ double sm_bc[2];
  const i64 lane = blockIdx.x;
  const int t = threadIdx.x;
  const T* base = p + lane * s_lane;
  double* obase = out + lane * o_lane;
  typedef RVec<T, VW> V;
  double x[NV][VW];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    i64 i = ((i64)j * THREADS + t) * VW;
    if (VW > 1 && i + VW <= n_red) {
      V v = *reinterpret_cast<const V*>(base + i);
#pragma unroll
      for (int q = 0; q < VW; ++q) x[j][q] = (double)v.v[q];
    } else {
#pragma unroll
      for (int q = 0; q < VW; ++q) x[j][q] = (i + q < n_red) ? (double)base[i + q] : -INFINITY;
    }
  }
  // ---- max (fmax drops NaNs; a NaN element still poisons the sum below) ----
  double mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < NV; ++j)
#pragma unroll
    for (int q = 0; q < VW; ++q) mx = fmax(mx, x[j][q]);
  for (int d = 16; d > 0; d >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, d));
  if ((t & 31) == 0) sm_red[t >> 5] = mx;
  __syncthreads();
  if (t == 0) {
    double m = sm_red[0];
    for (int w = 1; w < THREADS / 32; ++w) m = fmax(m, sm_red[w]);
    sm_bc[0] = m;
  }
  __syncthreads();
  mx = sm_bc[0];
  // ---- exp and sum ----
  double s = 0.0;
#pragma unroll
  for (int j = 0; j < NV; ++j)
#pragma unroll
    for (int q = 0; q < VW; ++q) {
      i64 i = ((i64)j * THREADS + t) * VW + q;
      double e = (i < n_red) ? fast_exp(x[j][q] - mx) : 0.0;
      x[j][q] = e;
      s += e;
    }
  for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
  __syncthreads();  // sm_red reuse
  if ((t & 31) == 0) sm_red[t >> 5] = s;
  __syncthreads();
  if (t == 0) {
    double a = sm_red[0];
    for (int w = 1; w < THREADS / 32; ++w) a += sm_red[w];
    sm_bc[1] = 1.0 / a;
  }
  __syncthreads();
  s = sm_bc[1];
  // ---- normalise and store (e * (1/s): <= 1.5 ulp from the reference's e / s) ----
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    i64 i = ((i64)j * THREADS + t) * VW;
    if (VW > 1 && i + VW <= n_red) {
      RVec<double, VW> v;
#pragma unroll
      for (int q = 0; q < VW; ++q) v.v[q] = x[j][q] * s;
      *reinterpret_cast<RVec<double, VW>*>(obase + i) = v;
    } else {
#pragma unroll
      for (int q = 0; q < VW; ++q)
        if (i + q < n_red) obase[i + q] = x[j][q] * s;
    }
  }
}

// Rows of 4097..32768 doubles, 16-byte aligned: the WHOLE lane is staged on chip, so HBM sees one
// read and one write and every load of the lane is in flight at once.  A CTA of 512 threads owns up
// to 16384 elements: 112 KB of shared memory (cp.async; each thread later reads back only the chunks
// it copied, so the data needs no barrier) plus 4 doubles per thread in registers.  Two CTAs fit an
// SM, so one row's exp phase overlaps another row's loads/stores.  Rows longer than 16384 use a
// 2-CTA cluster: max and sum are exchanged through distributed shared memory in rank order.
constexpr int OC_THREADS = 512;
constexpr int OC_REG = 2;                                        // double2 per thread held in registers
constexpr int OC_SMEM_CHUNKS = 14;                               // 16-byte chunks per thread in shared memory
constexpr int OC_SMEM_BYTES = OC_THREADS * OC_SMEM_CHUNKS * 16;  // 114688
constexpr int OC_CTA_VECS = OC_THREADS * (OC_REG + OC_SMEM_CHUNKS);  // 8192 double2 = 16384 doubles
constexpr int OC_CAP = 2 * OC_CTA_VECS * 2;                      // 32768 with a 2-CTA cluster

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Cluster barrier whose release is paid by ONE thread: the writer of the exchanged value fences, everybody
// arrives relaxed.  (A releasing arrive makes every thread wait for its own outstanding streaming stores.)
__device__ __forceinline__ void cluster_sync_light(bool i_wrote) {
  if (i_wrote) asm volatile("fence.acq_rel.cluster;" ::: "memory");
  asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ double ld_peer_f64(const double* local_smem, unsigned rank) {
  unsigned la = (unsigned)__cvta_generic_to_shared(local_smem), ra;
  double v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(rank));
  asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(ra));
  return v;
}

// FULL = every chunk of every thread is inside the row (n_red == CS * 16384): the per-chunk bounds checks and their 64-bit
// index arithmetic (a third of the issued instructions, ncu source page) drop out and addresses become base + immediate.
template <int CS, bool FULL>
__global__ void __launch_bounds__(OC_THREADS, 2) k_softmax_row_onchip(const double* __restrict__ p, i64 s_lane, i64 n_lane, i64 n_red,
                                                                       double* __restrict__ out, i64 o_lane) {
  // Persistent: cluster k takes rows k, k + #clusters, ...  While a thread normalises and stores chunk c of
  // row i it refills the same shared-memory slot with chunk c of row i+1 (cp.async), so the next row's
  // loads are in flight during this row's stores and have landed by the time its max pass starts; the
  // exp phase of one CTA overlaps the load/store phase of the other CTA on the SM.
  extern __shared__ __align__(16) double2 sx[];   // chunk c of thread t at sx[c * OC_THREADS + t]
  __shared__ double sm_red[OC_THREADS / 32];
  __shared__ double s_tab[32];                    // 2^(j/32): the table exp of common.cuh (exp_tab_nonpos)
  if (threadIdx.x < 32) s_tab[threadIdx.x] = k_exp2_32[threadIdx.x];   // published by the first barrier below
  const unsigned tab = (unsigned)__cvta_generic_to_shared(s_tab);
  __shared__ double sm_x[2][2];                   // [row parity][0] this CTA's max, [..][1] its sum (read by the peer)
  const int t = threadIdx.x;
  unsigned rank = 0;
  if (CS > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const i64 nclusters = gridDim.x / CS;
  const i64 nv = n_red >> 1;                       // whole double2 vectors of the row
  const i64 v0 = (i64)rank * OC_CTA_VECS;          // first vector owned by this CTA
  const bool odd_tail = !FULL && (n_red & 1) && t == 0 && rank == 0;

  auto issue_smem = [&](const double* base, int c) {
    const i64 v = v0 + (i64)(OC_REG + c) * OC_THREADS + t;
    if (FULL || v < nv) {
      unsigned sa = (unsigned)__cvta_generic_to_shared(&sx[c * OC_THREADS + t]);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(base + 2 * v) : "memory");
    }
  };
  i64 row = blockIdx.x / CS;
  if (row >= n_lane) return;   // whole cluster at once: nobody is left waiting at a cluster barrier
  {
    const double* base = p + row * s_lane;
#pragma unroll
    for (int c = 0; c < OC_SMEM_CHUNKS; ++c) issue_smem(base, c);
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  int par = 0;
  double2 rn[OC_REG];   // register-held chunks of the NEXT row, loaded during this row's store phase
  double xtn;
  {
    const double* base = p + row * s_lane;
#pragma unroll
    for (int c = 0; c < OC_REG; ++c) {
      const i64 v = v0 + (i64)c * OC_THREADS + t;
      rn[c] = (FULL || v < nv) ? *reinterpret_cast<const double2*>(base + 2 * v) : make_double2(-INFINITY, -INFINITY);
    }
    xtn = odd_tail ? base[n_red - 1] : -INFINITY;
  }
  for (; row < n_lane; row += nclusters, par ^= 1) {
    double* obase = out + row * o_lane;
    const i64 nrow = row + nclusters;
    const bool have_next = nrow < n_lane;
    const double* nbase = p + (have_next ? nrow : row) * s_lane;
    double2 r[OC_REG];
#pragma unroll
    for (int c = 0; c < OC_REG; ++c) r[c] = rn[c];
    double xt = xtn;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    // ---- max (fmax drops NaNs; a NaN still poisons the sum) ----
    double mx = fmax(xt, fmax(fmax(r[0].x, r[0].y), fmax(r[1].x, r[1].y)));
#pragma unroll
    for (int c = 0; c < OC_SMEM_CHUNKS; ++c) {
      const i64 v = v0 + (i64)(OC_REG + c) * OC_THREADS + t;
      if (FULL || v < nv) {
        double2 x = sx[c * OC_THREADS + t];
        mx = fmax(mx, fmax(x.x, x.y));
      }
    }
    for (int d = 16; d > 0; d >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, d));
    if ((t & 31) == 0) sm_red[t >> 5] = mx;
    __syncthreads();
    if (t == 0) {
      double m = sm_red[0];
      for (int w = 1; w < OC_THREADS / 32; ++w) m = fmax(m, sm_red[w]);
      sm_x[par][0] = m;
    }
    __syncthreads();
    if (CS > 1) {
      cluster_sync_light(t == 0);
      mx = fmax(ld_peer_f64(&sm_x[par][0], 0), ld_peer_f64(&sm_x[par][0], 1));
    } else {
      mx = sm_x[par][0];
    }
    // ---- exp (kept on chip) and sum ----
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < OC_REG; ++c) {
      const i64 v = v0 + (i64)c * OC_THREADS + t;
      if (FULL || v < nv) {
        r[c].x = exp_tab_nonpos(r[c].x - mx, tab);
        r[c].y = exp_tab_nonpos(r[c].y - mx, tab);
        s += r[c].x + r[c].y;
      }
    }
    if (odd_tail) { xt = exp_tab_nonpos(xt - mx, tab); s += xt; }
#pragma unroll
    for (int c = 0; c < OC_SMEM_CHUNKS; ++c) {
      const i64 v = v0 + (i64)(OC_REG + c) * OC_THREADS + t;
      if (FULL || v < nv) {
        double2 x = sx[c * OC_THREADS + t];
        x.x = exp_tab_nonpos(x.x - mx, tab);
        x.y = exp_tab_nonpos(x.y - mx, tab);
        s += x.x + x.y;
        sx[c * OC_THREADS + t] = x;
      }
    }
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    __syncthreads();
    if ((t & 31) == 0) sm_red[t >> 5] = s;
    __syncthreads();
    if (t == 0) {
      double a = sm_red[0];
      for (int w = 1; w < OC_THREADS / 32; ++w) a += sm_red[w];
      sm_x[par][1] = a;
    }
    __syncthreads();
    if (CS > 1) {
      cluster_sync_light(t == 0);
      s = ld_peer_f64(&sm_x[par][1], 0) + ld_peer_f64(&sm_x[par][1], 1);   // fixed rank order
    } else {
      s = sm_x[par][1];
    }
    const double inv = 1.0 / s;
    // ---- normalise, streaming stores; refill each slot with the next row's chunk as soon as it is stored ----
#pragma unroll
    for (int c = 0; c < OC_REG; ++c) {
      const i64 v = v0 + (i64)c * OC_THREADS + t;
      if (FULL || v < nv) __stcs(reinterpret_cast<double2*>(obase + 2 * v), make_double2(r[c].x * inv, r[c].y * inv));
      rn[c] = (have_next && (FULL || v < nv)) ? *reinterpret_cast<const double2*>(nbase + 2 * v) : make_double2(-INFINITY, -INFINITY);
    }
    xtn = (have_next && odd_tail) ? nbase[n_red - 1] : -INFINITY;
#pragma unroll
    for (int c = 0; c < OC_SMEM_CHUNKS; ++c) {
      const i64 v = v0 + (i64)(OC_REG + c) * OC_THREADS + t;
      if (FULL || v < nv) {
        double2 x = sx[c * OC_THREADS + t];
        __stcs(reinterpret_cast<double2*>(obase + 2 * v), make_double2(x.x * inv, x.y * inv));
      }
      if (have_next) issue_smem(nbase, c);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    if (odd_tail) obase[n_red - 1] = xt * inv;
    // sm_x is double-buffered by row parity: the peer reads [par] while this CTA may already write [par ^ 1];
    // [par] is rewritten two rows later, after two more cluster barriers.
  }
  if (CS > 1) cluster_sync_all();   // the peer may still be reading sm_x
}

template <int CS, bool FULL>
static int launch_row_onchip_impl(const double* p, i64 s_lane, i64 n_lane, i64 n_red, double* out, i64 o_lane) {
  {
    std::lock_guard<std::mutex> plan_lock(plan_mutex());
    static bool attr = false;
    if (!attr) {
      UM_CUDA(cudaFuncSetAttribute(k_softmax_row_onchip<CS, FULL>, cudaFuncAttributeMaxDynamicSharedMemorySize, OC_SMEM_BYTES));
      attr = true;
    }
  }
  cudaLaunchConfig_t cfg = {};
  i64 nclusters = (i64)ctx()->sm_count * 2 / CS;   // two CTAs per SM, all resident: the kernel is persistent
  if (nclusters > n_lane) nclusters = n_lane;
  cfg.gridDim = dim3((unsigned)(nclusters * CS), 1, 1);
  cfg.blockDim = dim3(OC_THREADS, 1, 1);
  cfg.dynamicSmemBytes = OC_SMEM_BYTES;
  cfg.stream = ctx()->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CS;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  UM_CUDA(cudaLaunchKernelEx(&cfg, k_softmax_row_onchip<CS, FULL>, p, s_lane, n_lane, n_red, out, o_lane));
  return UM_OK;
}

template <int CS>
static int launch_row_onchip(const double* p, i64 s_lane, i64 n_lane, i64 n_red, double* out, i64 o_lane) {
  return n_red == (i64)CS * OC_CTA_VECS * 2 ? launch_row_onchip_impl<CS, true>(p, s_lane, n_lane, n_red, out, o_lane)
                                            : launch_row_onchip_impl<CS, false>(p, s_lane, n_lane, n_red, out, o_lane);
}

template <class T, int VW>
static int softmax_contig(const T* p, i64 s_lane, i64 n_lane, i64 n_red, double* out, i64 o_lane) {
  cudaStream_t st = ctx()->stream;
  const unsigned g = (unsigned)n_lane;
  if (n_red <= 128 * VW) k_softmax_lane<T, 1, 128, VW><<<g, 128, 0, st>>>(p, s_lane, n_red, out, o_lane);
  else if (n_red <= 256 * 2 * VW) k_softmax_lane<T, 2, 256, VW><<<g, 256, 0, st>>>(p, s_lane, n_red, out, o_lane);
  else if (n_red <= 256 * 8 * VW) k_softmax_lane<T, 8, 256, VW><<<g, 256, 0, st>>>(p, s_lane, n_red, out, o_lane);
  else if (sizeof(T) == 8 && VW == 2 && n_red <= OC_CAP) {
    const double* pd = reinterpret_cast<const double*>(p);
    int rc = n_red <= OC_CTA_VECS * 2 ? launch_row_onchip<1>(pd, s_lane, n_lane, n_red, out, o_lane)
                                      : launch_row_onchip<2>(pd, s_lane, n_lane, n_red, out, o_lane);
    if (rc != UM_OK) return rc;
  } else k_softmax_row2pass<T, VW><<<g, 512, 0, st>>>(p, s_lane, n_red, out, o_lane);
  UM_LAUNCHED();
  return UM_OK;
}

}  // namespace um
#include "softmax_strip.cuh"
namespace um {
// softmax over strided lanes (axis 0 of a row-major matrix): 0 = automatic (the single-read strip kernel for large
// FP64 matrices it fits, the two-read kernels otherwise), 1 = two-read only, 2 = strip kernel or UM_ERR_ARG
static std::atomic<int> g_softmax0_path{0};
// share of the columns the automatic path hands to the two-read kernels next to the strip kernel (UM_SMS_SPLIT overrides)
static const double g_softmax0_split = [] { const char* e = getenv("UM_SMS_SPLIT"); return e ? atof(e) : 0.15; }();

// two-read softmax over strided (or generic) lanes: online (max, sum) per lane, then normalise
template <class T>
static int softmax_two_read(const Plan& pl, const T* p, double* o, i64 o_lane, i64 o_red) {
  SoftP::Acc* part = nullptr;
  int chunks = 0;
  // two extra lanes' worth of Acc slots hold the per-lane (m, 1/s) vectors of the vectorised pass
  const bool vecnorm = sizeof(T) == 8 && pl.mode == 1 && pl.s_lane == 1 && o_lane == 1 && pl.n_lane % 2 == 0 &&
                       pl.s_red % 2 == 0 && o_red % 2 == 0 && aligned16(p) && aligned16(o);
  int s = vecnorm ? stage1<SoftKP, T>(pl, &part, &chunks, (size_t)pl.n_lane + 2)
                  : stage1<SoftP, T>(pl, &part, &chunks, (size_t)pl.n_lane + 2);
  if (s != UM_OK) return s;
  if (vecnorm && sizeof(T) == 8 && pl.mode == 1 && pl.s_lane == 1 && o_lane == 1 && pl.n_lane % 2 == 0 && pl.s_red % 2 == 0 &&
      o_red % 2 == 0 && aligned16(p) && aligned16(o)) {
    double* mvec = reinterpret_cast<double*>(part + (i64)chunks * pl.n_lane);
    double* ivec = mvec + pl.n_lane;
    k_softmax_lane_stats<T><<<(unsigned)ceil_div(pl.n_lane, 256), 256, 0, ctx()->stream>>>(part, chunks, pl.n_lane, p, pl.s_red, pl.n_red, mvec, ivec);
    UM_LAUNCHED();
    i64 gyv = ceil_div(pl.n_red, 8 * 32);
    i64 tiles = ceil_div(pl.n_lane, 64);
    i64 want = ceil_div((i64)ctx()->sm_count * 32, tiles);
    if (gyv > want) gyv = want;
    if (gyv < 1) gyv = 1;
    dim3 gridv((unsigned)tiles, (unsigned)gyv), blockv(32, 8);
    k_softmax_norm_vec<<<gridv, blockv, 0, ctx()->stream>>>(reinterpret_cast<const double*>(p), pl.s_red, pl.n_lane, pl.n_red,
                                                             mvec, ivec, o, o_red);
    UM_LAUNCHED();
    return UM_OK;
  }
  i64 gy = ceil_div(pl.n_red, 8 * 16);
  if (gy > 2048) gy = 2048;
  if (gy < 1) gy = 1;
  dim3 grid((unsigned)ceil_div(pl.n_lane, 32), (unsigned)gy), block(32, 8);
  k_softmax_norm<T><<<grid, block, 0, ctx()->stream>>>(p, pl.s_lane, pl.s_red, pl.n_lane, pl.n_red, part, chunks, o, o_lane, o_red);
  UM_LAUNCHED();
  return UM_OK;
}

template <class T>
static int softmax_typed(const um_view* a, int axis, const um_view* out) {
  Plan pl = plan_axis<T>(a, axis);
  const T* p = static_cast<const T*>(pl.p);
  double* o = static_cast<double*>(out->data) + out->offset;
  const i64 o_lane = axis == 0 ? out->cs : out->rs;
  const i64 o_red = axis == 0 ? out->rs : out->cs;
  if (pl.mode == 0 && (o_red == 1 || pl.n_red == 1)) {
    constexpr int VW = 16 / sizeof(double);  // output vectors are double2
    bool vec = (reinterpret_cast<uintptr_t>(p) % (sizeof(T) * VW) == 0) && (pl.n_lane == 1 || pl.s_lane % VW == 0) &&
               aligned16(o) && (pl.n_lane == 1 || o_lane % VW == 0);
    return vec ? softmax_contig<T, VW>(p, pl.s_lane, pl.n_lane, pl.n_red, o, o_lane)
               : softmax_contig<T, 1>(p, pl.s_lane, pl.n_lane, pl.n_red, o, o_lane);
  }
  if (sizeof(T) == 8 && pl.mode == 1 && pl.s_lane == 1 && o_lane == 1) {
    const int path = g_softmax0_path.load();
    // automatic: the matrix must be large enough to be HBM-bound (the two-read kernels re-read from L2 below that)
    const bool want = path == 2 || (path == 0 && pl.n_red >= 512 && pl.n_lane >= 1024 && pl.n_red * pl.n_lane >= (i64(1) << 23));
    if (want) {
      // Split: the 16-CTA clusters of the strip kernel reach 112 of the 148 SMs and leave half of the FP64 issue slots
      // of those unused, so (automatic path, wide matrices) the last columns go to the two-read kernels on a second
      // stream at the same time; they fill the idle SMs and the gaps.  The strip kernel is launched first.
      i64 n2 = 0;
      if (path == 0 && pl.n_lane >= 8192 && g_softmax0_split > 0.0) n2 = (i64)((double)pl.n_lane * g_softmax0_split) / 16 * 16;
      const i64 n1 = pl.n_lane - n2;
      static thread_local cudaStream_t st2 = nullptr;
      static thread_local cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
      if (n2 > 0 && !st2) {
        UM_CUDA(cudaStreamCreateWithFlags(&st2, cudaStreamNonBlocking));
        UM_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
        UM_CUDA(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
      }
      if (n2 > 0) UM_CUDA(cudaEventRecord(ev_fork, ctx()->stream));
      const int rc = sms::launch(reinterpret_cast<const double*>(p), pl.s_red, pl.n_red, n1, o, o_red);
      if (rc != -1) {
        if (rc != UM_OK || n2 == 0) return rc;
        Plan pr = pl;
        pr.p = p + n1;
        pr.n_lane = n2;
        cudaStream_t main_st = ctx()->stream;
        UM_CUDA(cudaStreamWaitEvent(st2, ev_fork, 0));
        ctx()->stream = st2;
        const int rc2 = softmax_two_read<T>(pr, p + n1, o + n1, o_lane, o_red);
        ctx()->stream = main_st;
        if (rc2 != UM_OK) return rc2;
        UM_CUDA(cudaEventRecord(ev_join, st2));
        UM_CUDA(cudaStreamWaitEvent(main_st, ev_join, 0));
        return UM_OK;
      }
    }
    if (path == 2) return fail(UM_ERR_ARG, "softmax: operands do not fit the single-read strip kernel (FP64, row-major, even columns and pitches, 16-byte aligned, <= 32768 rows)");
  } else if (g_softmax0_path.load() == 2 && axis == 0) {
    return fail(UM_ERR_ARG, "softmax: operands do not fit the single-read strip kernel (FP64 row-major input and output)");
  }
  return softmax_two_read<T>(pl, p, o, o_lane, o_red);
}

}  // namespace um

using namespace um;

extern "C" {

int32_t um_reduce_axis(int32_t kind, const um_view* a, int32_t axis, const um_view* out) {
  UM_CTX();
  UM_VIEW(a); UM_VIEW(out);
  int s = check_axis_out(a, axis, out);
  if (s != UM_OK) return s;
  const i64 n_red = axis == 0 ? a->rows : a->cols;
  const i64 n_lane = axis == 0 ? a->cols : a->rows;
  if (n_lane == 0) return UM_OK;
  if (n_red == 0) {
    if (kind == UM_MIN || kind == UM_MAX) return fail(UM_ERR_EMPTY, "min/max of an empty lane");
    return um_fill(out, kind == UM_SUM ? 0.0 : nan(""));
  }
  if (a->dtype == UM_F64) return reduce_axis_typed<double>(kind, a, axis, out);
  if (a->dtype == UM_F32) return reduce_axis_typed<float>(kind, a, axis, out);
  return fail(UM_ERR_ARG, "reduce(axis) needs f64 or f32 input");
}

int32_t um_reduce_all(int32_t kind, const um_view* a, double* host_out) {
  UM_CTX();
  UM_VIEW(a);
  if (!host_out) return fail(UM_ERR_ARG, "null host_out");
  const i64 n = a->rows * a->cols;
  if (n == 0) {
    // sum of nothing = 0; mean of an empty matrix = 0 (Mat.scala:3470); min/max throw (:2216)
    if (kind == UM_MIN || kind == UM_MAX) return fail(UM_ERR_EMPTY, "min/max of an empty matrix");
    *host_out = (kind == UM_SUM || kind == UM_MEAN) ? 0.0 : nan("");
    return UM_OK;
  }
  if (kind == UM_MIN || kind == UM_MAX) {
    i64 idx = -1;
    int s = a->dtype == UM_F64 ? arg_all_typed<double>(kind == UM_MAX, a, &idx)
            : a->dtype == UM_F32 ? arg_all_typed<float>(kind == UM_MAX, a, &idx)
                                 : fail(UM_ERR_ARG, "min/max need f64 or f32 input");
    if (s != UM_OK) return s;
    return um_get(a, idx / a->cols, idx % a->cols, host_out);
  }
  if (a->dtype == UM_F64) return reduce_all_typed<double>(kind, a, host_out);
  if (a->dtype == UM_F32) return reduce_all_typed<float>(kind, a, host_out);
  return fail(UM_ERR_ARG, "reduce needs f64 or f32 input");
}

int32_t um_arg_all(int32_t want_max, const um_view* a, int64_t* row, int64_t* col) {
  UM_CTX();
  UM_VIEW(a);
  if (!row || !col) return fail(UM_ERR_ARG, "null output");
  if (a->rows * a->cols == 0) return fail(UM_ERR_EMPTY, "argmin/argmax of an empty matrix");
  i64 idx = -1;
  int s = a->dtype == UM_F64 ? arg_all_typed<double>(want_max, a, &idx)
          : a->dtype == UM_F32 ? arg_all_typed<float>(want_max, a, &idx)
                               : fail(UM_ERR_ARG, "argmin/argmax need f64 or f32 input");
  if (s != UM_OK) return s;
  *row = idx / a->cols;
  *col = idx % a->cols;
  return UM_OK;
}

int32_t um_idx_axis(int32_t want_max, const um_view* a, int32_t axis, const um_view* out) {
  UM_CTX();
  UM_VIEW(a); UM_VIEW(out);
  int s = check_axis_out(a, axis, out);
  if (s != UM_OK) return s;
  UM_REQUIRE(out->dtype == UM_I32, "idxmin/idxmax output must be i32");
  const i64 n_red = axis == 0 ? a->rows : a->cols;
  const i64 n_lane = axis == 0 ? a->cols : a->rows;
  if (n_lane == 0) return UM_OK;
  if (n_red == 0) return fail(UM_ERR_EMPTY, "idxmin/idxmax of an empty lane");
  if (a->dtype == UM_F64) {
    return want_max ? reduce_axis_run<ArgP<true>, double, int32_t>(a, axis, FIN_INDEX, out)
                    : reduce_axis_run<ArgP<false>, double, int32_t>(a, axis, FIN_INDEX, out);
  }
  if (a->dtype == UM_F32) {
    return want_max ? reduce_axis_run<ArgP<true>, float, int32_t>(a, axis, FIN_INDEX, out)
                    : reduce_axis_run<ArgP<false>, float, int32_t>(a, axis, FIN_INDEX, out);
  }
  return fail(UM_ERR_ARG, "idxmin/idxmax need f64 or f32 input");
}

int32_t um_mask_reduce(int32_t want_all, const um_view* a, int32_t axis, const um_view* out, int32_t* host_out) {
  UM_CTX();
  UM_VIEW(a);
  UM_REQUIRE(a->dtype == UM_BOOL, "all/any need a bool mask");
  const int kind = want_all ? FIN_ALL : FIN_ANY;
  if (axis < 0) {
    if (!host_out) return fail(UM_ERR_ARG, "null host_out");
    const i64 n = a->rows * a->cols;
    if (n == 0) { *host_out = want_all ? 1 : 0; return UM_OK; }
    SumP::Acc r;
    int s = reduce_all_acc<SumP, uint8_t>(a, &r);
    if (s != UM_OK) return s;
    *host_out = want_all ? (r.s == (double)n) : (r.s > 0.0);
    return UM_OK;
  }
  UM_VIEW(out);
  int s = check_axis_out(a, axis, out);
  if (s != UM_OK) return s;
  UM_REQUIRE(out->dtype == UM_BOOL, "all/any(axis) output must be bool");
  const i64 n_red = axis == 0 ? a->rows : a->cols;
  const i64 n_lane = axis == 0 ? a->cols : a->rows;
  if (n_lane == 0) return UM_OK;
  if (n_red == 0) return um_fill(out, want_all ? 1.0 : 0.0);
  return reduce_axis_run<SumP, uint8_t, uint8_t>(a, axis, kind, out);
}

}  // extern "C"

extern "C" int32_t um_softmax_set_axis0_path(int32_t path) {
  if (path < 0 || path > 2) return fail(UM_ERR_ARG, "um_softmax_set_axis0_path: 0 (automatic), 1 (two-read kernels), 2 (single-read strip kernel)");
  g_softmax0_path.store(path);
  return UM_OK;
}

extern "C" int32_t um_softmax(const um_view* a, int32_t axis, const um_view* out) {
  UM_CTX();
  UM_VIEW(a); UM_VIEW(out);
  UM_REQUIRE(axis == 0 || axis == 1, "axis must be 0 (columns) or 1 (rows)");
  UM_REQUIRE(out->rows == a->rows && out->cols == a->cols, "softmax: output shape mismatch");
  UM_REQUIRE(out->dtype == UM_F64, "softmax returns Mat[Double]: output must be f64");
  if (a->rows == 0 || a->cols == 0) return UM_OK;
  if (a->dtype == UM_F64) return softmax_typed<double>(a, axis, out);
  if (a->dtype == UM_F32) return softmax_typed<float>(a, axis, out);
  return fail(UM_ERR_ARG, "softmax needs f64 or f32 input");
}
