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

template <int KIND, class T>
__device__ __forceinline__ void cum_step(T& acc, T v) {
  if (KIND == CUM_SUM) acc += v;
  else if (KIND == CUM_MAX) { if (total_key(v) > total_key(acc)) acc = v; }
  else { if (total_key(v) < total_key(acc)) acc = v; }
}
template <int KIND, class T>
__device__ __forceinline__ T cum_seed(T first) { return KIND == CUM_SUM ? T(0) : first; }

// Lanes adjacent in memory (axis 0 of a row-major matrix): thread = lane, U loads in flight per thread.
template <int KIND, class T>
__global__ void __launch_bounds__(128) k_cum_lane_per_thread(const T* __restrict__ p, i64 s_lane, i64 s_red, i64 n_lane, i64 n_red,
                                                              T* __restrict__ out, i64 o_lane, i64 o_red) {
  constexpr int U = 16;
  const i64 l = (i64)blockIdx.x * 128 + threadIdx.x;
  if (l >= n_lane) return;
  const T* q = p + l * s_lane;
  T* o = out + l * o_lane;
  T acc = cum_seed<KIND, T>(q[0]);
  i64 i = 0;
  for (; i + U <= n_red; i += U) {
    T x[U];
#pragma unroll
    for (int u = 0; u < U; ++u) x[u] = q[(i + u) * s_red];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      cum_step<KIND, T>(acc, x[u]);
      o[(i + u) * o_red] = acc;
    }
  }
  for (; i < n_red; ++i) {
    cum_step<KIND, T>(acc, q[i * s_red]);
    o[i * o_red] = acc;
  }
}

// Lanes whose elements are adjacent in memory (axis 1 of a row-major matrix): a warp owns 32 lanes and walks
// them in 32 x 32 tiles -- coalesced loads into a padded shared tile, thread j folds lane j's 32 elements in
// order carrying its accumulator across tiles, coalesced stores back out.
template <int KIND, class T>
__global__ void __launch_bounds__(128) k_cum_lane_tiles(const T* __restrict__ p, i64 s_lane, i64 s_red, i64 n_lane, i64 n_red,
                                                         T* __restrict__ out, i64 o_lane, i64 o_red) {
  __shared__ T tile[4][32][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const i64 l0 = ((i64)blockIdx.x * 4 + w) * 32;
  if (l0 >= n_lane) return;
  const int nl = (int)min((i64)32, n_lane - l0);
  T acc = T(0);
  if (lane < nl) acc = cum_seed<KIND, T>(p[(l0 + lane) * s_lane]);
  for (i64 t0 = 0; t0 < n_red; t0 += 32) {
    const int nt = (int)min((i64)32, n_red - t0);
    if (lane < nt) {
#pragma unroll 8
      for (int r = 0; r < nl; ++r) tile[w][r][lane] = p[(l0 + r) * s_lane + (t0 + lane) * s_red];
    }
    __syncwarp();
    if (lane < nl) {
#pragma unroll 8
      for (int k = 0; k < nt; ++k) {
        cum_step<KIND, T>(acc, tile[w][lane][k]);
        tile[w][lane][k] = acc;
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
template <int KIND, class T>
__global__ void __launch_bounds__(32) k_cum_flat(const T* __restrict__ p, i64 rs, i64 cs, i64 rows, i64 cols, T* __restrict__ out) {
  constexpr int U = 4;
  const int lane = threadIdx.x;
  const i64 n = rows * cols;
  auto at = [&](i64 e) { return p[(e / cols) * rs + (e % cols) * cs]; };
  T acc = cum_seed<KIND, T>(n > 0 ? at(0) : T(0));
  for (i64 e0 = 0; e0 < n; e0 += 32 * U) {
    T x[U], mine[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const i64 e = e0 + u * 32 + lane;
      x[u] = e < n ? at(e) : T(0);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      mine[u] = T(0);
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        const T v = __shfl_sync(0xffffffffu, x[u], k);
        if (e0 + u * 32 + k < n) cum_step<KIND, T>(acc, v);   // uniform over the warp
        if (lane == k) mine[u] = acc;
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const i64 e = e0 + u * 32 + lane;
      if (e < n) out[e] = mine[u];
    }
  }
}

template <int KIND, class T>
static int cumulative_typed(const um_view* a, int axis, const um_view* out) {
  const T* p = static_cast<const T*>(a->data) + a->offset;
  T* o = static_cast<T*>(out->data) + out->offset;
  cudaStream_t st = ctx()->stream;
  if (axis < 0) {
    k_cum_flat<KIND, T><<<1, 32, 0, st>>>(p, a->rs, a->cs, a->rows, a->cols, o);
    UM_LAUNCHED();
    return UM_OK;
  }
  const i64 n_lane = axis == 0 ? a->cols : a->rows, n_red = axis == 0 ? a->rows : a->cols;
  const i64 s_lane = axis == 0 ? a->cs : a->rs, s_red = axis == 0 ? a->rs : a->cs;
  const i64 o_lane = axis == 0 ? out->cs : out->rs, o_red = axis == 0 ? out->rs : out->cs;
  auto mag = [](i64 s) { return s < 0 ? -s : s; };
  // which index is adjacent in memory?  (a stride-0 broadcast lane axis counts as adjacent)
  if (mag(s_lane) <= mag(s_red) || n_red == 1) {
    k_cum_lane_per_thread<KIND, T><<<(unsigned)((n_lane + 127) / 128), 128, 0, st>>>(p, s_lane, s_red, n_lane, n_red, o, o_lane, o_red);
  } else {
    k_cum_lane_tiles<KIND, T><<<(unsigned)((n_lane + 127) / 128), 128, 0, st>>>(p, s_lane, s_red, n_lane, n_red, o, o_lane, o_red);
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

template <class T>
static int where_typed(const um_view* c, const um_view* x, const um_view* y, double sx, double sy, const um_view* out) {
  const i64 rows = c->rows, cols = c->cols;
  if (rows == 0 || cols == 0) return UM_OK;
  i64 gy = (rows + 3) / 4;
  const i64 gx = (cols + 63) / 64;
  const i64 want = ((i64)ctx()->sm_count * 16 + gx - 1) / gx;
  if (gy > want) gy = want;
  if (gy > 65535) gy = 65535;
  dim3 grid((unsigned)gx, (unsigned)gy);
  if (x) k_where<T, false><<<grid, 256, 0, ctx()->stream>>>(in_of<uint8_t>(c), in_of<T>(x), in_of<T>(y), T(0), T(0), rows, cols, out_of<T>(out));
  else k_where<T, true><<<grid, 256, 0, ctx()->stream>>>(in_of<uint8_t>(c), In<T>{nullptr, 0, 0}, In<T>{nullptr, 0, 0}, (T)sx, (T)sy, rows, cols, out_of<T>(out));
  UM_LAUNCHED();
  return UM_OK;
}

// ---- m(mask): stable compaction in row-major order ---------------------------------------------------------
constexpr int SEL_CHUNK = 4096;   // flat elements per block
constexpr int SEL_THREADS = 256;

__device__ __forceinline__ bool mask_at(const In<uint8_t>& m, i64 cols, i64 e) { return m.p[(e / cols) * m.rs + (e % cols) * m.cs] != 0; }

__global__ void __launch_bounds__(SEL_THREADS) k_sel_count(In<uint8_t> m, i64 cols, i64 n, i64* __restrict__ counts) {
  __shared__ int sm[SEL_THREADS / 32];
  const i64 e0 = (i64)blockIdx.x * SEL_CHUNK;
  int c = 0;
  for (int k = threadIdx.x; k < SEL_CHUNK; k += SEL_THREADS) {
    const i64 e = e0 + k;
    if (e < n && mask_at(m, cols, e)) ++c;
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

// exclusive scan of the block counts in place; counts[nb] = total.  One block; integer sums are exact in any order.
__global__ void __launch_bounds__(1024) k_sel_scan(i64* __restrict__ counts, i64 nb) {
  __shared__ i64 sm[1024];
  __shared__ i64 carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (i64 b0 = 0; b0 < nb; b0 += 1024) {
    const i64 b = b0 + threadIdx.x;
    const i64 v = b < nb ? counts[b] : 0;
    sm[threadIdx.x] = v;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
      i64 add = threadIdx.x >= d ? sm[threadIdx.x - d] : 0;
      __syncthreads();
      sm[threadIdx.x] += add;
      __syncthreads();
    }
    if (b < nb) counts[b] = carry + sm[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += sm[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) counts[nb] = carry;
}

template <class T>
__global__ void __launch_bounds__(SEL_THREADS) k_sel_scatter(In<T> a, In<uint8_t> m, i64 cols, i64 n, const i64* __restrict__ offsets,
                                                              T* __restrict__ out, i64 out_n) {
  __shared__ int sm[SEL_THREADS / 32];
  const i64 e0 = (i64)blockIdx.x * SEL_CHUNK;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  i64 base = offsets[blockIdx.x];
  for (int k0 = 0; k0 < SEL_CHUNK; k0 += SEL_THREADS) {
    const i64 e = e0 + k0 + threadIdx.x;
    const bool t = e < n && mask_at(m, cols, e);
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
      if (pos < out_n) out[pos] = a.p[(e / cols) * a.rs + (e % cols) * a.cs];
    }
    base += total;
    __syncthreads();
  }
}

static int sel_offsets(const um_view* mask, i64** offsets_out, i64* nb_out) {
  const i64 n = mask->rows * mask->cols;
  const i64 nb = (n + SEL_CHUNK - 1) / SEL_CHUNK;
  i64* counts = static_cast<i64*>(scratch((size_t)(nb + 1) * sizeof(i64)));
  if (!counts) return UM_ERR_OOM;
  if (nb > 0) {
    k_sel_count<<<(unsigned)nb, SEL_THREADS, 0, ctx()->stream>>>(in_of<uint8_t>(mask), mask->cols, n, counts);
    UM_LAUNCHED();
  }
  k_sel_scan<<<1, 1024, 0, ctx()->stream>>>(counts, nb);
  UM_LAUNCHED();
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
  return UM_OK;
}

int32_t um_mask_select(const um_view* a, const um_view* mask, const um_view* out) {
  UM_CTX();
  UM_VIEW(a); UM_VIEW(mask); UM_VIEW(out);
  UM_REQUIRE(mask->dtype == UM_BOOL, "mask select needs a boolean mask");
  UM_REQUIRE(mask->rows == a->rows && mask->cols == a->cols, "Mask shape %lldx%lld must match matrix shape %lldx%lld", (i64)mask->rows,
             (i64)mask->cols, (i64)a->rows, (i64)a->cols);   // Mat.scala:4010-4011
  UM_REQUIRE(out->dtype == a->dtype && (out->rows == 1 || out->cols == 0) && (out->cs == 1 || out->cols <= 1), "mask select result must be a contiguous 1 x k row of the operand's type");
  const i64 n = a->rows * a->cols;
  if (n == 0 || out->cols == 0) return UM_OK;
  i64* offs = nullptr;
  i64 nb = 0;
  int s = sel_offsets(mask, &offs, &nb);
  if (s != UM_OK) return s;
  cudaStream_t st = ctx()->stream;
  const unsigned g = (unsigned)nb;
  void* o = static_cast<char*>(out->data) + (size_t)out->offset * dtype_size(out->dtype);
  switch (a->dtype) {
    case UM_F64: k_sel_scatter<double><<<g, SEL_THREADS, 0, st>>>(in_of<double>(a), in_of<uint8_t>(mask), a->cols, n, offs, static_cast<double*>(o), out->cols); break;
    case UM_F32: k_sel_scatter<float><<<g, SEL_THREADS, 0, st>>>(in_of<float>(a), in_of<uint8_t>(mask), a->cols, n, offs, static_cast<float*>(o), out->cols); break;
    case UM_I32: k_sel_scatter<int32_t><<<g, SEL_THREADS, 0, st>>>(in_of<int32_t>(a), in_of<uint8_t>(mask), a->cols, n, offs, static_cast<int32_t*>(o), out->cols); break;
    default: k_sel_scatter<uint8_t><<<g, SEL_THREADS, 0, st>>>(in_of<uint8_t>(a), in_of<uint8_t>(mask), a->cols, n, offs, static_cast<uint8_t*>(o), out->cols); break;
  }
  UM_LAUNCHED();
  return UM_OK;
}

}  // extern "C"
