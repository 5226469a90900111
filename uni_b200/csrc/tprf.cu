// tprf.cu -- Kelly-Pruitt Three-Pass Regression Filter on the device (stats/Tprf3.scala).
//
// The reference's closed form builds an N x N `Sxx` and an N x T factor (Tprf3.scala:206-207):
// infeasible at N = 262144.  Both fits are re-associated here onto ONE
// set of shard-additive accumulators that never exceed T x L:
//
//   per column j :  mu_j, sd_j (sample std, Tprf3.scala:469-502), W0_j = Zc' x_j / sd_j   (L)
//   PX  = sum_j (x_j / sd_j) W0_j'      (T x L)      h0  = sum_j x_j / sd_j             (T)
//   sw  = sum_j W0_j    SWW = sum_j W0_j W0_j'    smw = sum_j (mu_j/sd_j) W0_j    smu = sum_j mu_j/sd_j
//
// with Zc = J(T) Z.  Closed form (Tprf3.scala:199-209, rust/src/t3prf.rs:861-869):
//   wbar = sw/N,  G = PX - 1 smw' - (h0 - smu) wbar',  s = (G'G)^-1 G'y,  yhat = G s + ybar,
//   alpha_j = (W0_j - wbar) s.
// Three-pass fit (Tprf3.scala:260-289): Phi = W0 Szz^-1 (pass-1 slopes with intercept),
//   [1|Phi]'[1|Phi] from (N, sw, SWW),  Xn [1|Phi] = [h0 | PX Szz^-1],  then passes 2-3 are O(T L^2).
//
// Kernels (all batched over independent panels through blockIdx.z):
//   k_tprf_prep      Zc, zbar, Szz, ybar, ssy                       (tiny)
//   k_tprf_colstats  sweep 1 over X: per column  sum(x-k), sum(x-k)^2 (FP64 pipe) and Zc' x
//                    (DMMA m8n8k4, M = 8 columns, N = L<=8, K = rows), cp.async 3-stage tiles
//   k_tprf_colfinal  per column mu/sd/W0/V + block partials of sw, SWW, smw, smu
//   k_tprf_project   sweep 2 over X: PX, h0 (DMMA, M = 8 rows, N = L, K = columns)
//   k_tprf_gather    fixed-order sum of the chunk partials into the packed all-reduce buffer
//   k_tprf_finish    the O(T L^2) finish of either fit (one CTA per panel)
//   k_tprf_outputs   alpha / phi / phi_hat / xstd per column
// Multi-GPU: X is sharded by predictor column; the packed buffer (T*(L8+1)+81 doubles) is the
// only thing exchanged (um_comm_allreduce_sum_f64), then every rank runs the same finish.
#include "common.cuh"
#include "small_linalg.cuh"

namespace um {

int comm_allreduce_sum_f64_impl(double* buf, i64 count);  // comm.cu
int comm_check_error();                                     // comm.cu: did a bounded peer wait give up? (call after a sync)
bool comm_active();
int comm_world();

constexpr int LP = 8;            // padded proxy count (DMMA N)
constexpr int ZPITCH = 12;       // smem pitch of an (rows x 8) operand: conflict-free B fragments
constexpr int NSMALL = 8 + 64 + 8 + 1;  // sw | SWW | smw | smu = 81

// ---- small helpers -----------------------------------------------------------------------------------
__device__ __forceinline__ void t_cp16(void* smem, const void* g, int bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(g), "r"(bytes) : "memory");
}
__device__ __forceinline__ void t_cp8(void* smem, const void* g, int bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(s), "l"(g), "r"(bytes) : "memory");
}
__device__ __forceinline__ void t_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void t_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void t_dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// copy a (rows x cols) box of doubles (global row pitch ld) into smem (pitch), zero-filling
// anything outside [0,row_lim) x [0,col_lim).  cols multiple of 2.
template <bool VEC16>
__device__ __forceinline__ void stage_box(double* sm, int pitch, const double* __restrict__ g, i64 ld, i64 r0, i64 c0,
                                          int rows, int cols, i64 row_lim, i64 col_lim, int tid, int nthreads) {
  const int ch = VEC16 ? 2 : 1;
  const int cpr = cols / ch;
  const int total = rows * cpr;
  for (int c = tid; c < total; c += nthreads) {
    int r = c / cpr, q = (c % cpr) * ch;
    i64 gr = r0 + r, gc = c0 + q;
    i64 rem = (gr < row_lim) ? (col_lim - gc) : 0;
    int nv = rem <= 0 ? 0 : (rem >= ch ? ch : (int)rem);
    const double* src = nv ? (g + gr * ld + gc) : g;
    if (VEC16) t_cp16(sm + r * pitch + q, src, nv * 8);
    else t_cp8(sm + r * pitch + q, src, nv * 8);
  }
}

// ---- prep: Zc (T x 8, zero padded), zbar, Szz, ybar, ssy ------------------------------------------------
// small[] layout per panel: zbar[8] | Szz[64] | ybar | ssy | (pad to 80)
constexpr int SMALL_PREP = 80;
constexpr int FIN_CS = 8;   // cluster size of the row-parallel prep / finish kernels
__device__ __forceinline__ unsigned t_cluster_rank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void t_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void t_st_cluster(double* local_addr, unsigned rank, double v) {
  unsigned la = (unsigned)__cvta_generic_to_shared(local_addr), ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(rank));
  asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(ra), "d"(v) : "memory");
}
template <int CS>   // CS > 1: a cluster of CS CTAs, each a slice of the rows (one panel only)
__global__ void __launch_bounds__(256) k_tprf_prep(const double* __restrict__ Z, i64 z_rs, i64 z_cs, i64 z_bs,
                                                    const double* __restrict__ y, i64 y_s, i64 y_bs, i64 T, int L,
                                                    double* __restrict__ zc, double* __restrict__ small) {
  __shared__ double red[8][48];
  __shared__ double cbuf[2][CS][48];
  const int crank = CS > 1 ? (int)t_cluster_rank() : 0;
  constexpr int STRIDE = 256 * CS;
  __shared__ double zbar[LP];
  __shared__ double ybar_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const i64 b = blockIdx.z;   // CS > 1 is launched for one panel
  Z += b * z_bs;
  y += b * y_bs;
  zc += b * T * LP;
  small += b * SMALL_PREP;
  // pass 1: column sums of Z and sum of y
  double acc[LP + 1];
#pragma unroll
  for (int l = 0; l <= LP; ++l) acc[l] = 0.0;
  for (i64 t = (i64)crank * 256 + tid; t < T; t += STRIDE) {
#pragma unroll
    for (int l = 0; l < LP; ++l)
      if (l < L) acc[l] += Z[t * z_rs + l * z_cs];
    acc[LP] += y[t * y_s];
  }
#pragma unroll
  for (int l = 0; l <= LP; ++l) {
    double v = acc[l];
    for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
    if (lane == 0) red[warp][l] = v;
  }
  __syncthreads();
  if (CS > 1) {   // CTA totals -> every CTA of the cluster, then a rank-order sum (identical everywhere)
    if (tid <= LP) {
      double v = 0.0;
      for (int w = 0; w < 8; ++w) v += red[w][tid];
      for (unsigned r = 0; r < (unsigned)CS; ++r) t_st_cluster(&cbuf[0][crank][tid], r, v);
    }
    t_cluster_sync();
  }
  if (tid <= LP) {
    double v = 0.0;
    if (CS > 1) for (int r = 0; r < CS; ++r) v += cbuf[0][r][tid];
    else for (int w = 0; w < 8; ++w) v += red[w][tid];
    v /= (double)T;
    if (tid < LP) zbar[tid] = v;
    else ybar_s = v;
  }
  __syncthreads();
  // pass 2: Zc, Szz (upper triangle, 36 values), ssy
  double sz[36];
#pragma unroll
  for (int i = 0; i < 36; ++i) sz[i] = 0.0;
  double ssy = 0.0;
  const double yb = ybar_s;
  for (i64 t = (i64)crank * 256 + tid; t < T; t += STRIDE) {
    double z[LP];
#pragma unroll
    for (int l = 0; l < LP; ++l) {
      z[l] = l < L ? Z[t * z_rs + l * z_cs] - zbar[l] : 0.0;
      zc[t * LP + l] = z[l];
    }
    int q = 0;
#pragma unroll
    for (int a = 0; a < LP; ++a)
#pragma unroll
      for (int c = a; c < LP; ++c) sz[q++] += z[a] * z[c];
    double d = y[t * y_s] - yb;
    ssy += d * d;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 37; ++i) {
    double v = i < 36 ? sz[i] : ssy;
    for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
    if (lane == 0) red[warp][i] = v;
  }
  __syncthreads();
  if (tid < 37) {
    double v = 0.0;
    for (int w = 0; w < 8; ++w) v += red[w][tid];
    if (CS > 1) for (unsigned r = 0; r < (unsigned)CS; ++r) t_st_cluster(&cbuf[1][crank][tid], r, v);
    else red[0][tid] = v;
  }
  __syncthreads();
  if (CS > 1) {
    t_cluster_sync();
    if (tid < 37) {
      double v = 0.0;
      for (int r = 0; r < CS; ++r) v += cbuf[1][r][tid];
      red[0][tid] = v;
    }
    __syncthreads();
  }
  if (tid == 0 && crank == 0) {
    for (int l = 0; l < LP; ++l) small[l] = zbar[l];
    int q = 0;
    for (int a = 0; a < LP; ++a)
      for (int c = a; c < LP; ++c) {
        double v = red[0][q++];
        small[8 + a * LP + c] = v;
        small[8 + c * LP + a] = v;
      }
    small[72] = ybar_s;
    small[73] = red[0][36];
  }
}

// ---- sweep 1: column statistics + Zc' x ------------------------------------------------------------------
constexpr int C1_ROWS = 32, C1_COLS = 64, C1_PITCH = 68, C1_STAGES = 3, C1_THREADS = 256;
constexpr int C1_STAGE_DBL = C1_ROWS * C1_PITCH + C1_ROWS * ZPITCH;  // 2176 + 384 = 2560 doubles
constexpr int C1_SMEM = C1_STAGES * C1_STAGE_DBL * 8;                // 61440 B

// outputs per (panel, chunk, column): q[8], s1, s2   (column index padded to strips*64)
template <bool VEC16>
__global__ void __launch_bounds__(C1_THREADS, 3)
k_tprf_colstats(const double* __restrict__ X, i64 ldx, i64 x_bs, const double* __restrict__ zc, i64 T, i64 N, i64 npad,
                i64 rows_per_chunk, double* __restrict__ qp, double* __restrict__ s1p, double* __restrict__ s2p) {
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const i64 b = blockIdx.z;
  const int chunk = blockIdx.y, nchunk = gridDim.y;
  X += b * x_bs;
  zc += b * T * LP;
  const i64 c0 = (i64)blockIdx.x * C1_COLS;
  const i64 r_lo = (i64)chunk * rows_per_chunk;
  i64 r_hi = r_lo + rows_per_chunk;
  if (r_hi > T) r_hi = T;
  const int ntiles = (int)((r_hi - r_lo + C1_ROWS - 1) / C1_ROWS);

  auto stage = [&](int s, int tile) {
    double* xs = smem + (size_t)s * C1_STAGE_DBL;
    double* zs = xs + C1_ROWS * C1_PITCH;
    const i64 r0 = r_lo + (i64)tile * C1_ROWS;
    stage_box<VEC16>(xs, C1_PITCH, X, ldx, r0, c0, C1_ROWS, C1_COLS, r_hi, N, tid, C1_THREADS);
    stage_box<true>(zs, ZPITCH, zc, LP, r0, 0, C1_ROWS, LP, r_hi, LP, tid, C1_THREADS);
  };

  const int fj = lane >> 2, ft = lane & 3;         // fragment column / row-in-k4
  const i64 col = c0 + warp * 8 + fj;
  const double kshift = (col < N) ? X[col] : 0.0;  // row 0 of the panel: the variance shift
  double acc0[2] = {0.0, 0.0}, acc1[2] = {0.0, 0.0};
  double s1 = 0.0, s2 = 0.0;

#pragma unroll
  for (int s = 0; s < C1_STAGES - 1; ++s) {
    if (s < ntiles) stage(s, s);
    t_commit();
  }
  for (int tile = 0; tile < ntiles; ++tile) {
    t_wait<C1_STAGES - 2>();
    __syncthreads();
    {
      int nt = tile + C1_STAGES - 1;
      if (nt < ntiles) stage(nt % C1_STAGES, nt);
      t_commit();
    }
    const double* xs = smem + (size_t)(tile % C1_STAGES) * C1_STAGE_DBL;
    const double* zs = xs + C1_ROWS * C1_PITCH;
    const i64 r0 = r_lo + (i64)tile * C1_ROWS;
    const bool full = r0 + C1_ROWS <= r_hi;
#pragma unroll
    for (int ks = 0; ks < C1_ROWS / 4; ++ks) {
      const int t = ks * 4 + ft;
      const double a = xs[t * C1_PITCH + warp * 8 + fj];
      const double bz = zs[t * ZPITCH + fj];
      if (ks & 1) t_dmma(acc1[0], acc1[1], a, bz);
      else t_dmma(acc0[0], acc0[1], a, bz);
      double d = a - kshift;
      if (!full && r0 + t >= r_hi) d = 0.0;
      s1 += d;
      s2 = fma(d, d, s2);
    }
  }
  t_wait<0>();
  // fold the 4 row-phases of each column
  s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
  s2 += __shfl_xor_sync(0xffffffffu, s2, 1);
  s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
  s2 += __shfl_xor_sync(0xffffffffu, s2, 2);
  const i64 pidx = (b * nchunk + chunk) * npad + col;  // col < npad always
  double2 q = make_double2(acc0[0] + acc1[0], acc0[1] + acc1[1]);
  *reinterpret_cast<double2*>(qp + pidx * LP + 2 * ft) = q;  // C[j = fj][l = 2*ft, 2*ft+1]
  if (ft == 0) {
    s1p[pidx] = s1;
    s2p[pidx] = s2;
  }
}

// ---- per-column finalisation ------------------------------------------------------------------------------
// cols arrays per panel (npad entries): w0[8], v[8] (= w0/sd), invsd, mun (= mu/sd)
__global__ void __launch_bounds__(256)
k_tprf_colfinal(const double* __restrict__ X, i64 x_bs, const double* __restrict__ qp, const double* __restrict__ s1p,
                const double* __restrict__ s2p, int nchunk, i64 T, i64 N, i64 npad, double* __restrict__ w0,
                double* __restrict__ v, double* __restrict__ invsd, double* __restrict__ mun,
                double* __restrict__ blockpart /* [panel][gridDim.x][81] */) {
  __shared__ double red[8][NSMALL];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const i64 b = blockIdx.z;
  X += b * x_bs;
  double loc[NSMALL];
#pragma unroll
  for (int i = 0; i < NSMALL; ++i) loc[i] = 0.0;
  for (i64 j = (i64)blockIdx.x * 256 + tid; j < N; j += (i64)gridDim.x * 256) {
    double S = 0.0, SS = 0.0, q[LP];
#pragma unroll
    for (int l = 0; l < LP; ++l) q[l] = 0.0;
    for (int c = 0; c < nchunk; ++c) {
      const i64 pidx = (b * nchunk + c) * npad + j;
      S += s1p[pidx];
      SS += s2p[pidx];
      const double2* qq = reinterpret_cast<const double2*>(qp + pidx * LP);
#pragma unroll
      for (int l = 0; l < LP / 2; ++l) {
        double2 t2 = qq[l];
        q[2 * l] += t2.x;
        q[2 * l + 1] += t2.y;
      }
    }
    const double k = X[j];
    const double dm = S / (double)T;
    const double mu = k + dm;
    double sd = 1.0;
    if (T > 1) {
      double var = (SS - S * dm) / (double)(T - 1);
      if (var < 0.0) var = 0.0;
      sd = sqrt(var);       // NaN input stays NaN (the reference's matmuls would spread it too)
      if (sd == 0.0) sd = 1.0;  // Tprf3.scala:493-501
    }
    const double inv = 1.0 / sd;
    const double m_n = mu * inv;
    const i64 cidx = b * npad + j;
    double w[LP];
#pragma unroll
    for (int l = 0; l < LP; ++l) {
      w[l] = q[l] * inv;
      w0[cidx * LP + l] = w[l];
      v[cidx * LP + l] = w[l] * inv;
    }
    invsd[cidx] = inv;
    mun[cidx] = m_n;
#pragma unroll
    for (int a = 0; a < LP; ++a) {
      loc[a] += w[a];
#pragma unroll
      for (int c = 0; c < LP; ++c) loc[8 + a * LP + c] += w[a] * w[c];
      loc[72 + a] += m_n * w[a];
    }
    loc[80] += m_n;
  }
  // columns beyond N up to npad: zero V / invsd so sweep 2 ignores the padding
  for (i64 j = N + (i64)blockIdx.x * 256 + tid; j < npad; j += (i64)gridDim.x * 256) {
    const i64 cidx = b * npad + j;
#pragma unroll
    for (int l = 0; l < LP; ++l) { w0[cidx * LP + l] = 0.0; v[cidx * LP + l] = 0.0; }
    invsd[cidx] = 0.0;
    mun[cidx] = 0.0;
  }
#pragma unroll
  for (int i = 0; i < NSMALL; ++i) {
    double x = loc[i];
    for (int d = 16; d > 0; d >>= 1) x += __shfl_down_sync(0xffffffffu, x, d);
    if (lane == 0) red[warp][i] = x;
  }
  __syncthreads();
  if (tid < NSMALL) {
    double x = 0.0;
    for (int w = 0; w < 8; ++w) x += red[w][tid];
    blockpart[(b * gridDim.x + blockIdx.x) * NSMALL + tid] = x;
  }
}

// ---- sw / SWW / smw / smu from the per-column arrays (single-read sweeps with small groups, where the
// finaliser warp owns up to 16 columns per piece and cannot afford the 8 x 8 outer products) ----------------
__global__ void __launch_bounds__(256)
k_tprf_smallsums(const double* __restrict__ w0, const double* __restrict__ mun, i64 N, i64 npad, int nblk,
                 double* __restrict__ blockpart /* [panel][nblk][81]: block x < gridDim.x adds into slot x */) {
  __shared__ double red[8][NSMALL];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const i64 b = blockIdx.z;
  double loc[NSMALL];
#pragma unroll
  for (int i = 0; i < NSMALL; ++i) loc[i] = 0.0;
  for (i64 j = (i64)blockIdx.x * 256 + tid; j < N; j += (i64)gridDim.x * 256) {
    const i64 cidx = b * npad + j;
    double w[LP];
#pragma unroll
    for (int l = 0; l < LP; ++l) w[l] = w0[cidx * LP + l];
    const double m_n = mun[cidx];
#pragma unroll
    for (int a = 0; a < LP; ++a) {
      loc[a] += w[a];
#pragma unroll
      for (int c = 0; c < LP; ++c) loc[8 + a * LP + c] += w[a] * w[c];
      loc[72 + a] += m_n * w[a];
    }
    loc[80] += m_n;
  }
#pragma unroll
  for (int i = 0; i < NSMALL; ++i) {
    double x = loc[i];
    for (int d = 16; d > 0; d >>= 1) x += __shfl_down_sync(0xffffffffu, x, d);
    if (lane == 0) red[warp][i] = x;
  }
  __syncthreads();
  if (tid < NSMALL) {
    double x = 0.0;
    for (int w = 0; w < 8; ++w) x += red[w][tid];
    blockpart[(b * nblk + blockIdx.x) * NSMALL + tid] += x;   // the finalisers wrote zeros there
  }
}

// ---- sweep 2: PX = Xn W0, h0 = Xn 1 ----------------------------------------------------------------------
constexpr int C2_ROWS = 128, C2_COLS = 32, C2_PITCH = 36, C2_STAGES = 4, C2_THREADS = 256;
constexpr int C2_STAGE_DBL = C2_ROWS * C2_PITCH + C2_COLS * ZPITCH + C2_COLS;  // 4608 + 384 + 32 = 5024
constexpr int C2_SMEM = C2_STAGES * C2_STAGE_DBL * 8;                          // 120576 B

template <bool VEC16>
__global__ void __launch_bounds__(C2_THREADS, 1)
k_tprf_project(const double* __restrict__ X, i64 ldx, i64 x_bs, const double* __restrict__ v,
               const double* __restrict__ invsd, i64 T, i64 N, i64 npad, i64 cols_per_chunk, i64 tpad,
               double* __restrict__ pxp /* [panel][chunk][tpad][8] */, double* __restrict__ h0p /* [panel][chunk][tpad] */) {
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const i64 b = blockIdx.z;
  const int chunk = blockIdx.y, nchunk = gridDim.y;
  X += b * x_bs;
  v += b * npad * LP;
  invsd += b * npad;
  const i64 r0 = (i64)blockIdx.x * C2_ROWS;
  const i64 c_lo = (i64)chunk * cols_per_chunk;
  i64 c_hi = c_lo + cols_per_chunk;
  if (c_hi > npad) c_hi = npad;
  const int ntiles = c_hi > c_lo ? (int)((c_hi - c_lo + C2_COLS - 1) / C2_COLS) : 0;

  auto stage = [&](int s, int tile) {
    double* xs = smem + (size_t)s * C2_STAGE_DBL;
    double* vs = xs + C2_ROWS * C2_PITCH;
    double* is = vs + C2_COLS * ZPITCH;
    const i64 c0 = c_lo + (i64)tile * C2_COLS;
    stage_box<VEC16>(xs, C2_PITCH, X, ldx, r0, c0, C2_ROWS, C2_COLS, T, N, tid, C2_THREADS);
    stage_box<true>(vs, ZPITCH, v, LP, c0, 0, C2_COLS, LP, npad, LP, tid, C2_THREADS);
    stage_box<true>(is, C2_COLS, invsd, npad, 0, c0, 1, C2_COLS, 1, npad, tid, C2_THREADS);
  };

  const int fr = lane >> 2, fk = lane & 3;
  double acc[2][2][2];  // [m-tile][k parity][2]
  double h[2] = {0.0, 0.0};
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int p = 0; p < 2; ++p) acc[i][p][0] = acc[i][p][1] = 0.0;

#pragma unroll
  for (int s = 0; s < C2_STAGES - 1; ++s) {
    if (s < ntiles) stage(s, s);
    t_commit();
  }
  for (int tile = 0; tile < ntiles; ++tile) {
    t_wait<C2_STAGES - 2>();
    __syncthreads();
    {
      int nt = tile + C2_STAGES - 1;
      if (nt < ntiles) stage(nt % C2_STAGES, nt);
      t_commit();
    }
    const double* xs = smem + (size_t)(tile % C2_STAGES) * C2_STAGE_DBL;
    const double* vs = xs + C2_ROWS * C2_PITCH;
    const double* is = vs + C2_COLS * ZPITCH;
#pragma unroll
    for (int ks = 0; ks < C2_COLS / 4; ++ks) {
      const int j = ks * 4 + fk;
      const double bv = vs[j * ZPITCH + fr];   // V[j][l = fr]
      const double si = is[j];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const double a = xs[(warp * 16 + i * 8 + fr) * C2_PITCH + j];
        t_dmma(acc[i][ks & 1][0], acc[i][ks & 1][1], a, bv);
        h[i] = fma(a, si, h[i]);
      }
    }
  }
  t_wait<0>();
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    h[i] += __shfl_xor_sync(0xffffffffu, h[i], 1);
    h[i] += __shfl_xor_sync(0xffffffffu, h[i], 2);
    const i64 t = r0 + warp * 16 + i * 8 + fr;   // < tpad always
    const i64 pidx = (b * nchunk + chunk) * tpad + t;
    double2 q = make_double2(acc[i][0][0] + acc[i][1][0], acc[i][0][1] + acc[i][1][1]);
    *reinterpret_cast<double2*>(pxp + pidx * LP + 2 * fk) = q;
    if (fk == 0) h0p[pidx] = h[i];
  }
}

// ---- gather chunk partials into the packed buffer: PX[T][8] | h0[T] | small[81] --------------------------
__global__ void __launch_bounds__(256)
k_tprf_gather(const double* __restrict__ pxp, const double* __restrict__ h0p, int nchunk, i64 T, i64 tpad,
              const double* __restrict__ blockpart, int nblk, double* __restrict__ packed, i64 packed_stride,
              int accumulate) {
  // the next kernel in the stream (exchange or finish, launched with programmatic stream serialisation) may be made
  // resident while this one runs; it waits (griddepcontrol.wait) for this grid's completion before it reads anything
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const i64 b = blockIdx.z;
  double* out = packed + b * packed_stride;
  const i64 total = T * (LP + 1) + NSMALL;
  for (i64 e = (i64)blockIdx.x * 256 + threadIdx.x; e < total; e += (i64)gridDim.x * 256) {
    double s = 0.0;
    if (e < T * LP) {
      const i64 t = e / LP;
      const int l = (int)(e % LP);
      for (int c = 0; c < nchunk; ++c) s += pxp[((b * nchunk + c) * tpad + t) * LP + l];
    } else if (e < T * (LP + 1)) {
      const i64 t = e - T * LP;
      for (int c = 0; c < nchunk; ++c) s += h0p[(b * nchunk + c) * tpad + t];
    } else {
      const int i = (int)(e - T * (LP + 1));
      for (int k = 0; k < nblk; ++k) s += blockpart[(b * nblk + k) * NSMALL + i];
    }
    out[e] = accumulate ? out[e] + s : s;
  }
}

// ---- finish: one CTA per panel -----------------------------------------------------------------------------
struct FinishArgs {
  const double* packed; i64 packed_stride;
  const double* small;          // prep output (zbar | Szz | ybar | ssy), SMALL_PREP per panel
  const double* y; i64 y_s, y_bs;
  i64 T; int L; double n_total; int mode;
  int want3;                                           // CLOSED only: also run the three-pass finish for sigma / beta3 / phi (Tprf3.scala:211-231)
  double* forecasts; double* residuals; i64 f_bs;     // per panel stride (elements); may be null
  double* sigma; double* beta3;                        // single-panel outputs (may be null)
  double* sigma_ws;                                    // T x 8 workspace per panel
  double* params;                                      // per panel: s[8] | wbar[8] | szzinv[64] | zbar[8] | r2 | singular
  double* r2_out;                                      // per panel r2 (device), may be null
  double* host_res;                                    // mapped pinned host memory (r2 | singular) of a single-panel fit, may be null:
                                                       // the fit's two host scalars arrive with the kernel, no copy-engine round trip
};
constexpr int PARAMS = 96;

// Sum of NV per-thread values over the CTA and, when the kernel runs as a cluster of CS CTAs (each CTA a slice
// of the rows), over the cluster: every CTA writes its CTA totals into slot [rank] of EVERY CTA's `cbuf`
// (distributed shared memory), one cluster barrier, then everyone adds the CS slots in rank order -- all
// CTAs end up with bit-identical totals and redo the tiny serial algebra that follows redundantly.
template <int NV, int CS>
__device__ __forceinline__ void block_sum(double* loc, double (*red)[64], double* result, int tid, double (*cbuf)[64] = nullptr) {
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    double x = loc[i];
    for (int d = 16; d > 0; d >>= 1) x += __shfl_down_sync(0xffffffffu, x, d);
    if (lane == 0) red[warp][i] = x;
  }
  __syncthreads();
  if (CS == 1) {
    if (tid < NV) {
      double x = 0.0;
      for (int w = 0; w < 8; ++w) x += red[w][tid];
      result[tid] = x;
    }
    __syncthreads();
  } else {
    if (tid < NV) {
      double x = 0.0;
      for (int w = 0; w < 8; ++w) x += red[w][tid];
      const unsigned me = t_cluster_rank();
      for (unsigned r = 0; r < (unsigned)CS; ++r) t_st_cluster(&cbuf[me][tid], r, x);
    }
    t_cluster_sync();
    if (tid < NV) {
      double x = 0.0;
      for (int r = 0; r < CS; ++r) x += cbuf[r][tid];
      result[tid] = x;
    }
    __syncthreads();
  }
}

// THREE = false: the closed form alone (tprfClosedForm without the pass-1/2/3 fields): the three-pass block is compiled
// out -- the kernel runs once per fit on eight CTAs, so what it costs is mostly instructions fetched, not executed
template <int CS, bool THREE>
__global__ void __launch_bounds__(256) k_tprf_finish(FinishArgs g) {
  asm volatile("griddepcontrol.wait;" ::: "memory");   // no-op unless launched as a programmatic dependent
  __shared__ double red[8][64];
  __shared__ double tot[64];
  __shared__ double cbuf[5][CS][64];   // one exchange buffer per reduction stage (no reuse, no second barrier)
  __shared__ double lu_a[81], lu_inv[81], lu_x[81];   // warp-cooperative small inverses (small_linalg.cuh)
  __shared__ int lu_piv[9];
  const int crank = CS > 1 ? (int)t_cluster_rank() : 0;
  constexpr int STRIDE = 256 * CS;
  __shared__ double sw[LP], smw[LP], wbar[LP], s_cl[LP], szzinv[LP * LP], m2[(LP + 1) * LP], beta[LP + 1];
  __shared__ double smu_s, ybar_s, ssy_s;
  __shared__ int singular_s;
  const int tid = threadIdx.x;
  const i64 b = blockIdx.x / CS;
  const i64 T = g.T;
  const int L = g.L, L1 = g.L + 1;
  const double* px = g.packed + b * g.packed_stride;
  const double* h0 = px + T * LP;
  const double* sm = h0 + T;  // sw[8] | SWW[64] | smw[8] | smu
  const double* small = g.small + b * SMALL_PREP;
  const double* y = g.y + b * g.y_bs;
  double* par = g.params + b * PARAMS;
  double* fc = g.forecasts ? g.forecasts + b * g.f_bs : nullptr;
  double* rs = g.residuals ? g.residuals + b * g.f_bs : nullptr;
  double* sig_ws = g.sigma_ws + b * T * LP;
  if (tid < LP) {
    sw[tid] = sm[tid];
    smw[tid] = sm[72 + tid];
    wbar[tid] = sm[tid] / g.n_total;
    s_cl[tid] = 0.0;
  }
  if (tid < LP * LP) szzinv[tid] = 0.0;
  if (tid == 0) {
    smu_s = sm[80];
    ybar_s = small[72];
    ssy_s = small[73];
    singular_s = 0;
  }
  __syncthreads();
  const double ybar = ybar_s, ssy = ssy_s, smu = smu_s;
  double r2 = 0.0;

  if (g.mode != UM_TPRF_ITER && g.mode != UM_TPRF_WINDOW) {
    // ---- closed form: G = PX - 1 smw' - (h0 - smu) wbar' ; s = (G'G)^-1 G'y ----
    double loc[44];
#pragma unroll
    for (int i = 0; i < 44; ++i) loc[i] = 0.0;
    for (i64 t = (i64)crank * 256 + tid; t < T; t += STRIDE) {
      double gl[LP];
      const double r = h0[t] - smu;
#pragma unroll
      for (int l = 0; l < LP; ++l) gl[l] = l < L ? px[t * LP + l] - smw[l] - r * wbar[l] : 0.0;
      int q = 0;
#pragma unroll
      for (int a = 0; a < LP; ++a)
#pragma unroll
        for (int c = a; c < LP; ++c) loc[q++] += gl[a] * gl[c];
      const double yt = y[t * g.y_s];
#pragma unroll
      for (int l = 0; l < LP; ++l) loc[36 + l] += gl[l] * yt;
    }
    block_sum<44, CS>(loc, red, tot, tid, cbuf[0]);
    if (tid < 32) {
      if (tid == 0) {
        int q = 0;
        double full[LP][LP];
        for (int a = 0; a < LP; ++a)
          for (int c = a; c < LP; ++c) { full[a][c] = tot[q]; full[c][a] = tot[q]; ++q; }
        for (int a = 0; a < L; ++a)
          for (int c = 0; c < L; ++c) lu_a[a * L + c] = full[a][c];
      }
      __syncwarp();
      const bool ok = lu_inverse_warp(lu_a, L, lu_inv, lu_x, lu_piv, tid);
      if (tid == 0 && !ok) singular_s = 1;
      if (tid < LP) {
        double s = 0.0;
        if (ok && tid < L)
          for (int c = 0; c < L; ++c) s += lu_inv[tid * L + c] * tot[36 + c];
        s_cl[tid] = tid >= L ? 0.0 : (ok ? s : nan(""));
      }
    }
    __syncthreads();
    if (g.mode == UM_TPRF_CLOSED) {
      double see = 0.0;
      for (i64 t = (i64)crank * 256 + tid; t < T; t += STRIDE) {
        const double r = h0[t] - smu;
        double f = 0.0;
#pragma unroll
        for (int l = 0; l < LP; ++l)
          if (l < L) f += (px[t * LP + l] - smw[l] - r * wbar[l]) * s_cl[l];
        f += ybar;
        const double e = y[t * g.y_s] - f;
        if (fc) fc[t] = f;
        if (rs) rs[t] = e;
        see += e * e;
      }
      double l1[1] = {see};
      block_sum<1, CS>(l1, red, tot, tid, cbuf[1]);
      r2 = ssy == 0.0 ? 0.0 : 1.0 - tot[0] / ssy;  // Tprf3.scala:239-241
    }
  }

  // tprfClosedForm also returns the pass-1/2/3 quantities (phi, sigma, beta3, phiHat: Tprf3.scala:211-231,243-252),
  // which are those of t3prf: the same block runs for CLOSED when they were asked for, without touching yhat / R^2
  const bool three = THREE && (g.mode != UM_TPRF_CLOSED || g.want3 != 0);
  const bool three_fc = g.mode != UM_TPRF_CLOSED;
  if constexpr (THREE) if (three) {
    // ---- three-pass fit from the same accumulators ----
    // The small dense algebra between the sweeps' sums and pass 2 used to run on ONE thread with its operands in
    // global memory (about 2 000 dependent multiply-adds, ~90 us of a 120 us kernel): it now runs across warp 0, 64
    // matrix entries two per lane, operands staged in shared memory, one __syncwarp per stage.
    __shared__ double sww_s[LP * LP], swwi_s[LP * LP], pp_s[LP * LP], phisum_s[LP];
    if (tid < LP * LP) sww_s[tid] = sm[8 + tid];
    __syncthreads();
    if (tid < 32) {
      for (int e = tid; e < L * L; e += 32) lu_a[e] = small[8 + (e / L) * LP + (e % L)];
      __syncwarp();
      const bool ok = lu_inverse_warp(lu_a, L, lu_inv, lu_x, lu_piv, tid);
      if (tid == 0 && !ok) singular_s = 1;
      for (int e = tid; e < LP * LP; e += 32) {
        const int a = e / LP, c = e % LP;
        szzinv[e] = (ok && a < L && c < L) ? lu_inv[a * L + c] : (ok ? 0.0 : nan(""));
      }
      __syncwarp();
      // [1|Phi]'[1|Phi] with Phi = W0 Szz^-1  (Tprf3.scala:265-266)
      if (tid < L) {
        double s = 0.0;
        for (int c = 0; c < L; ++c) s += szzinv[tid * LP + c] * sw[c];
        phisum_s[tid] = s;
      }
      for (int e = tid; e < LP * LP; e += 32) {      // sww_i = Szz^-1 * SWW
        const int a = e / LP, c = e % LP;
        double s = 0.0;
        if (a < L && c < L)
          for (int k = 0; k < L; ++k) s += szzinv[a * LP + k] * sww_s[k * LP + c];
        swwi_s[e] = s;
      }
      __syncwarp();
      for (int e = tid; e < LP * LP; e += 32) {      // pp = sww_i * Szz^-1' (Szz^-1 symmetric up to rounding)
        const int a = e / LP, c = e % LP;
        double s = 0.0;
        if (a < L && c < L)
          for (int k = 0; k < L; ++k) s += swwi_s[a * LP + k] * szzinv[c * LP + k];
        pp_s[e] = s;
      }
      __syncwarp();
      double* ptp = lu_a;
      if (tid == 0) ptp[0] = g.n_total;
      if (tid < L) {
        ptp[0 * L1 + 1 + tid] = phisum_s[tid];
        ptp[(1 + tid) * L1 + 0] = phisum_s[tid];
      }
      for (int e = tid; e < L * L; e += 32) {
        const int a = e / L, c = e % L;
        ptp[(1 + a) * L1 + 1 + c] = 0.5 * (pp_s[a * LP + c] + pp_s[c * LP + a]);
      }
      __syncwarp();
      const bool ok2 = lu_inverse_warp(lu_a, L1, lu_inv, lu_x, lu_piv, tid);
      if (tid == 0 && !ok2) singular_s = 1;
      const double* ptpinv = lu_inv;
      // Sigma[t][l] = sum_m XP[t][m] PtPinv[1+l][m],  XP[t][0] = h0[t],  XP[t][1+a] = sum_c PX[t][c] szzinv[c][a]
      // fold into m2: Sigma[t][l] = m2[0][l] h0[t] + sum_c m2[1+c][l] PX[t][c]
      if (tid < LP) m2[0 * LP + tid] = (tid < L) ? (ok2 ? ptpinv[(1 + tid) * L1 + 0] : nan("")) : 0.0;
      for (int e = tid; e < LP * LP; e += 32) {
        const int c = e / LP, l = e % LP;
        double s = 0.0;
        if (l < L && c < L)
          for (int a = 0; a < L; ++a) s += szzinv[c * LP + a] * ptpinv[(1 + l) * L1 + 1 + a];
        m2[(1 + c) * LP + l] = (ok2 || !(l < L && c < L)) ? s : nan("");
      }
    }
    __syncthreads();
    // estimate3prf's pass 3 is nanOls(y, [1 | Sigma], minObs = 1) (Tprf3.scala:929-931, 807-830): rows with a NaN in y
    // or Sigma are left out of the normal equations, fewer than one usable row -> NaN coefficients.  t3prf and
    // tprfClosedForm take the plain product (:268-269, :229-231), where a NaN poisons the fit.
    const bool nan_rows = g.mode == UM_TPRF_IS_FULL || g.mode == UM_TPRF_WINDOW;
    double loc[55];
#pragma unroll
    for (int i = 0; i < 55; ++i) loc[i] = 0.0;
    for (i64 t = (i64)crank * 256 + tid; t < T; t += STRIDE) {
      double d[LP + 1];
      d[0] = 1.0;
      const double h = h0[t];
      double pxr[LP];
#pragma unroll
      for (int c = 0; c < LP; ++c) pxr[c] = px[t * LP + c];
#pragma unroll
      for (int l = 0; l < LP; ++l) {
        double s = m2[l] * h;
#pragma unroll
        for (int c = 0; c < LP; ++c) s += m2[(1 + c) * LP + l] * pxr[c];
        d[1 + l] = l < L ? s : 0.0;
        sig_ws[t * LP + l] = d[1 + l];
      }
      const double yt = y[t * g.y_s];
      if (nan_rows) {
        bool ok = yt == yt;
#pragma unroll
        for (int a = 1; a <= LP; ++a) ok = ok && d[a] == d[a];
        if (!ok) continue;
        loc[54] += 1.0;
      }
      int q = 0;
#pragma unroll
      for (int a = 0; a <= LP; ++a)
#pragma unroll
        for (int c = a; c <= LP; ++c) loc[q++] += d[a] * d[c];
#pragma unroll
      for (int a = 0; a <= LP; ++a) loc[45 + a] += d[a] * yt;
    }
    block_sum<55, CS>(loc, red, tot, tid, cbuf[2]);
    if (tid < 32) {
      if (tid == 0) {
        double full[LP + 1][LP + 1];
        int q = 0;
        for (int a = 0; a <= LP; ++a)
          for (int c = a; c <= LP; ++c) { full[a][c] = tot[q]; full[c][a] = tot[q]; ++q; }
        for (int a = 0; a < L1; ++a)
          for (int c = 0; c < L1; ++c) lu_a[a * L1 + c] = full[a][c];
      }
      __syncwarp();
      const bool none = nan_rows && tot[54] < 1.0;   // no usable row: NaN coefficients, not a singular-matrix error
      const bool ok = none ? false : lu_inverse_warp(lu_a, L1, lu_inv, lu_x, lu_piv, tid);
      if (tid == 0 && !ok && !none) singular_s = 1;
      if (tid <= LP) {
        double s = 0.0;
        if (tid < L1 && ok)
          for (int c = 0; c < L1; ++c) s += lu_inv[tid * L1 + c] * tot[45 + c];
        beta[tid] = tid < L1 ? (ok ? s : nan("")) : 0.0;
      }
    }
    __syncthreads();
    double see = 0.0, sy = 0.0, cnt = 0.0;
    for (i64 t = (i64)crank * 256 + tid; t < T; t += STRIDE) {
      double f = beta[0];
#pragma unroll
      for (int l = 0; l < LP; ++l)
        if (l < L) f += sig_ws[t * LP + l] * beta[1 + l];
      const double yt = y[t * g.y_s];
      const double e = yt - f;
      if (three_fc) {
        if (fc) fc[t] = f;
        if (rs) rs[t] = e;
      }
      if (!nan_rows || e == e) {   // IS Full: R^2 over the rows with a residual (`loc`, Tprf3.scala:1207)
        see += e * e;
        sy += yt;
        cnt += 1.0;
      }
      if (g.sigma)
        for (int l = 0; l < L; ++l) g.sigma[(b * T + t) * L + l] = sig_ws[t * LP + l];
    }
    double l3[3] = {see, sy, cnt};
    block_sum<3, CS>(l3, red, tot, tid, cbuf[3]);
    const double see_tot = tot[0];
    if (g.mode == UM_TPRF_ITER || g.mode == UM_TPRF_WINDOW) r2 = ssy == 0.0 ? 0.0 : 1.0 - see_tot / ssy;  // Tprf3.scala:274-275
    else if (g.mode == UM_TPRF_IS_FULL) {
      // 1 - sum e^2 / sum (y - mean)^2 over `loc`, mean over `loc` too, no zero guard (Tprf3.scala:1214-1225)
      const double mu = tot[1] / tot[2];
      __syncthreads();   // everyone has read tot[] before the next reduction overwrites it
      double syy = 0.0;
      for (i64 t = (i64)crank * 256 + tid; t < T; t += STRIDE) {
        double f = beta[0];
#pragma unroll
        for (int l = 0; l < LP; ++l)
          if (l < L) f += sig_ws[t * LP + l] * beta[1 + l];
        const double yt = y[t * g.y_s];
        const double e = yt - f;
        if (e == e) {
          const double dd = yt - mu;
          syy += dd * dd;
        }
      }
      double l1[1] = {syy};
      block_sum<1, CS>(l1, red, tot, tid, cbuf[4]);
      r2 = 1.0 - see_tot / tot[0];
    }
    if (crank == 0 && g.beta3 && tid < L1) g.beta3[b * L1 + tid] = beta[tid];
  }
  __syncthreads();
  if (crank != 0) return;
  if (tid < LP) {
    par[tid] = s_cl[tid];
    par[8 + tid] = wbar[tid];
    par[80 + tid] = small[tid];
  }
  if (tid < LP * LP) par[16 + tid] = three ? szzinv[tid] : 0.0;
  if (tid == 0) {
    par[88] = r2;
    par[89] = (double)singular_s;
    if (g.r2_out) g.r2_out[b] = r2;
    if (g.host_res) {
      g.host_res[0] = r2;
      g.host_res[1] = (double)singular_s;
    }
  }
}

// ---- per-column outputs ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_tprf_outputs(const double* __restrict__ w0, const double* __restrict__ invsd, const double* __restrict__ mun,
               const double* __restrict__ params, i64 N, i64 npad, int L, int mode, double* __restrict__ alpha,
               double* __restrict__ phi, double* __restrict__ phi_hat, double* __restrict__ xstd, i64 out_bs) {
  __shared__ double par[PARAMS];
  const i64 b = blockIdx.z;
  if (threadIdx.x < PARAMS) par[threadIdx.x] = params[b * PARAMS + threadIdx.x];
  __syncthreads();
  const double* s = par;
  const double* wbar = par + 8;
  const double* szzinv = par + 16;
  const double* zbar = par + 80;
  for (i64 j = (i64)blockIdx.x * 256 + threadIdx.x; j < N; j += (i64)gridDim.x * 256) {
    const i64 cidx = b * npad + j;
    double w[LP];
#pragma unroll
    for (int l = 0; l < LP; ++l) w[l] = w0[cidx * LP + l];
    if (alpha) {
      double a = 0.0;
#pragma unroll
      for (int l = 0; l < LP; ++l)
        if (l < L) a += (w[l] - wbar[l]) * s[l];
      alpha[b * out_bs + j] = a;
    }
    if (xstd) xstd[b * out_bs + j] = 1.0 / invsd[cidx];
    if (phi || phi_hat) {
      double icpt = mun[cidx];
      for (int l = 0; l < L; ++l) {
        double p = 0.0;
        for (int c = 0; c < L; ++c) p += szzinv[l * LP + c] * w[c];
        if (phi) phi[(b * out_bs + j) * L + l] = p;
        if (phi_hat) phi_hat[b * out_bs * (L + 1) + (i64)(1 + l) * N + j] = p;
        icpt -= zbar[l] * p;
      }
      if (phi_hat) phi_hat[b * out_bs * (L + 1) + j] = icpt;
    }
  }
}

__global__ void k_fill_nan(double* p, i64 n) {
  i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = nan("");
}

}  // namespace um
#include "tprf_fused.cuh"
#include "tprf_grid.cuh"
namespace um {

// ---- host orchestration ------------------------------------------------------------------------------------------
static i64 cdiv(i64 a, i64 b) { return (a + b - 1) / b; }
static size_t al(size_t x) { return (x + 255) & ~size_t(255); }

struct Plan3 {
  i64 T, N, L, batch;
  i64 strips, npad, rchunks, rows_per_chunk;     // sweep 1
  i64 rblocks, tpad, cchunks, cols_per_chunk;    // sweep 2
  int cf_blocks;                                 // colfinal blocks per panel
  // single-read fused sweep (tprf_fused.cuh): cluster size, rows per piece, chunks per panel, clusters
  bool fused = false;
  bool grid_sweep = false;   // tprf_grid.cuh (virtual groups over all SMs, TMEM stash) instead of hardware clusters
  int fG = 0, fR = 0, fcpp = 0, fncl = 0;
  size_t o_gpart = 0, o_grec = 0, o_gcnt = 0;
  // split sweep: the 16-CTA cluster shape strands SMs; a grid sweep of `split_groups` virtual groups on them takes
  // the last `split_tiles` column tiles concurrently (second stream)
  int split_groups = 0;
  int fcpp_out = 0;          // chunk slots per panel in the partial buffers (= fcpp, + split_groups when split)
  // workspace offsets (bytes)
  size_t o_zc, o_small, o_qp, o_s1, o_s2, o_w0, o_v, o_invsd, o_mun, o_bp, o_pxp, o_h0p, o_packed, o_sig, o_par, total;
  i64 packed_len;
};

static void carve(Plan3& p) {
  const i64 T = p.T, batch = p.batch;
  p.packed_len = T * (LP + 1) + NSMALL;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t at = o; o += al(bytes); return at; };
  p.o_zc = take((size_t)batch * T * LP * 8);
  p.o_small = take((size_t)batch * SMALL_PREP * 8);
  p.o_qp = take((size_t)batch * p.rchunks * p.npad * LP * 8);
  p.o_s1 = take((size_t)batch * p.rchunks * p.npad * 8);
  p.o_s2 = take((size_t)batch * p.rchunks * p.npad * 8);
  p.o_w0 = take((size_t)batch * p.npad * LP * 8);
  p.o_v = take((size_t)batch * p.npad * LP * 8);
  p.o_invsd = take((size_t)batch * p.npad * 8);
  p.o_mun = take((size_t)batch * p.npad * 8);
  p.o_bp = take((size_t)batch * p.cf_blocks * NSMALL * 8);
  p.o_pxp = take((size_t)batch * p.cchunks * p.tpad * LP * 8);
  p.o_h0p = take((size_t)batch * p.cchunks * p.tpad * 8);
  p.o_packed = take((size_t)batch * p.packed_len * 8);
  p.o_sig = take((size_t)batch * T * LP * 8);
  p.o_par = take((size_t)batch * PARAMS * 8);
  if (p.grid_sweep || p.split_groups) {
    const size_t groups = p.grid_sweep ? (size_t)p.fncl : (size_t)p.split_groups;
    p.o_gpart = take(groups * gridk::NX * gridk::PART_ENTRY * 16);
    p.o_grec = take(groups * gridk::NX * gridk::REC_ENTRY * 16);   // directly after o_gpart: zeroed together
  }
  p.total = o;
}

static Plan3 make_plan(i64 T, i64 N, i64 L, i64 batch, int sms) {
  Plan3 p;
  p.T = T; p.N = N; p.L = L; p.batch = batch;
  const i64 target = (i64)sms * 16;
  p.strips = cdiv(N, C1_COLS);
  p.npad = p.strips * C1_COLS;
  i64 rc = 1;
  if (p.strips * batch < target) {
    rc = cdiv(target, p.strips * batch);
    i64 maxc = cdiv(T, 4 * C1_ROWS);
    if (rc > maxc) rc = maxc;
    if (rc < 1) rc = 1;
  }
  p.rows_per_chunk = cdiv(cdiv(T, rc), C1_ROWS) * C1_ROWS;
  p.rchunks = cdiv(T, p.rows_per_chunk);
  p.rblocks = cdiv(T, C2_ROWS);
  p.tpad = p.rblocks * C2_ROWS;
  i64 cc = 1;
  if (p.rblocks * batch < target) {
    cc = cdiv(target, p.rblocks * batch);
    i64 maxc = cdiv(p.npad, 4 * C2_COLS);
    if (cc > maxc) cc = maxc;
    if (cc < 1) cc = 1;
  }
  p.cols_per_chunk = cdiv(cdiv(p.npad, cc), C2_COLS) * C2_COLS;
  p.cchunks = cdiv(p.npad, p.cols_per_chunk);
  i64 cfb = cdiv(N, 256 * 4);
  if (cfb > 1024) cfb = 1024;
  if (cfb < 1) cfb = 1;
  p.cf_blocks = (int)cfb;
  carve(p);
  return p;
}

static void plan_fused(Plan3& p, const double* X, i64 ldx, i64 x_bs);
// plan for a panel resident at X: the fused single-read sweep when it applies, else two sweeps
static Plan3 make_plan_for(i64 T, i64 N, i64 L, i64 batch, const double* X, i64 ldx, i64 x_bs) {
  Plan3 p = make_plan(T, N, L, batch, ctx()->sm_count);
  plan_fused(p, X, ldx, x_bs);
  if (p.fused) {
    // same workspace carving, with the fused sweep's partial counts: no row-chunk partials,
    // one PX / h0 partial per chunk, one small-sum partial per (chunk, cluster rank)
    Plan3 q = p;
    q.rchunks = 0;
    q.cchunks = p.fcpp_out;
    q.tpad = (i64)p.fG * p.fR;
    q.cf_blocks = p.fcpp_out * p.fG;
    carve(q);
    return q;
  }
  return p;
}

static int set_smem_attrs() {
  std::lock_guard<std::mutex> plan_lock(plan_mutex());
  static bool done = false;
  if (done) return UM_OK;
  UM_CUDA(cudaFuncSetAttribute(k_tprf_colstats<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, C1_SMEM));
  UM_CUDA(cudaFuncSetAttribute(k_tprf_colstats<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, C1_SMEM));
  UM_CUDA(cudaFuncSetAttribute(k_tprf_project<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, C2_SMEM));
  UM_CUDA(cudaFuncSetAttribute(k_tprf_project<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, C2_SMEM));
  done = true;
  return UM_OK;
}

// ---- path selection: 0 = auto (fused when it applies), 1 = two-sweep only, 2 = fused or fail ----
static std::atomic<int> g_tprf_path{0};
static long long* g_fused_dbg = nullptr;  // device buffer for clock stamps (um_tprf_debug_stamps)
// relative speed of one virtual group of the grid sweep vs one hardware cluster (measured at config 4: 5.1 ms on 9
// groups vs 4.8 ms on 7 clusters); UM_TPRF_SPLIT_SPEED overrides, 0 disables the split
static double g_split_share = [] {
  const char* e = getenv("UM_TPRF_SPLIT_SPEED");
  return e ? atof(e) : 0.73;
}();

static fused::FusedKernel cluster_kernel_for(int R) { return fused::kernel_for(R); }
static int cluster_threads() { return fused::NTHREADS; }
static int cluster_smem() { return fused::SMEM_ALLOC; }

typedef CUresult (*StreamWaitFn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
static StreamWaitFn stream_wait_fn() {
  std::lock_guard<std::mutex> plan_lock(plan_mutex());
  static StreamWaitFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &p, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<StreamWaitFn>(p);
  }
  return fn;
}

static int fused_clusters(int G, int R) {
  std::lock_guard<std::mutex> plan_lock(plan_mutex());
  // resident clusters of size G of the fused kernel (rows-per-piece variant R/128) on this device, cached
  static int cache[5][fused::GMAX + 1] = {{0}};
  static bool attr_done[5] = {false, false, false, false, false};
  const int j = R / 128;
  fused::FusedKernel kern = cluster_kernel_for(R);
  if (!attr_done[j]) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, cluster_smem()) != cudaSuccess ||
        cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
      cudaGetLastError();
      return 0;
    }
    attr_done[j] = true;
  }
  if (cache[j][G]) return cache[j][G] > 0 ? cache[j][G] : 0;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(G * 64), 1, 1);
  cfg.blockDim = dim3((unsigned)cluster_threads(), 1, 1);
  cfg.dynamicSmemBytes = (size_t)cluster_smem();
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)G; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { cudaGetLastError(); n = 0; }
  cache[j][G] = n > 0 ? n : -1;
  return n > 0 ? n : 0;
}

static int grid_groups(int G, int R) {
  std::lock_guard<std::mutex> plan_lock(plan_mutex());
  // groups of G CTAs of the grid-wide sweep, one CTA per SM (cooperative launch: all co-resident)
  static bool attr_done[5] = {false, false, false, false, false};
  const int j = R / 128;
  gridk::GridKernel kern = gridk::grid_kernel_for(R);
  if (!attr_done[j]) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, gridk::GSMEM_ALLOC) != cudaSuccess) {
      cudaGetLastError();
      return 0;
    }
    attr_done[j] = true;
  }
  int coop = 0;
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx()->device);
  if (!coop) return 0;
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, gridk::G_NTHREADS, gridk::GSMEM_ALLOC) != cudaSuccess || per_sm < 1) {
    cudaGetLastError();
    return 0;
  }
  return ctx()->sm_count / G;
}

// decide whether the fused sweep applies to this panel; fill p.fused / fG / fR / fcpp / fncl
static void plan_fused(Plan3& p, const double* X, i64 ldx, i64 x_bs) {
  p.fused = false;
  if (g_tprf_path.load() == 1) return;
  if (!fused::encode_fn()) return;
  if (!aligned16(X) || (ldx % 2) != 0 || (x_bs % 2) != 0 || ldx < p.N) return;
  if (p.N >= (i64(1) << 31) || p.batch >= 65536) return;
  int G = 0, R = 0;
  fused::pick_shape(p.T, &G, &R);
  if (!G) return;
  const int path = g_tprf_path.load();
  int ncl = 0;
  p.grid_sweep = false;
  if (path == 3) {
    ncl = grid_groups(G, R);
    p.grid_sweep = ncl > 0;
  }
  if (!p.grid_sweep) {
    if (path == 3) return;
    ncl = fused_clusters(G, R);
  }
  if (ncl <= 0) return;
  const i64 ntile = cdiv(p.N, fused::CW);
  i64 cpp;
  if (p.batch == 1) {
    cpp = ntile < ncl ? ntile : ncl;
  } else {
    // chunks = batch * cpp go round-robin over the clusters: pick the smallest cpp that balances
    cpp = 1;
    while (cpp * 2 * 8 <= ntile) {
      const i64 tot = p.batch * cpp;
      const double eff = (double)tot / (double)(cdiv(tot, ncl) * ncl);
      if (eff >= 0.95) break;
      cpp *= 2;
    }
  }
  if (cpp < 1) cpp = 1;
  p.fused = true;
  p.fG = G; p.fR = R; p.fcpp = (int)cpp; p.fncl = ncl;
  p.fcpp_out = p.fcpp;
  p.split_groups = 0;
  // Split sweep (auto only): hardware clusters of G CTAs leave SMs stranded (7 x 16 = 112 of 148 on a B200);
  // a concurrent grid sweep of virtual groups on those SMs takes a share of the column tiles.
  if (path == 0 && !p.grid_sweep && p.batch == 1 && G >= 8 && g_split_share > 0.0) {
    const int spare = (ctx()->sm_count - ncl * G) / G;
    if (spare >= 1 && grid_groups(G, R) > 0 && ntile >= 32 * (i64)(ncl + spare) && cpp == ncl) {
      p.split_groups = spare;
      p.fcpp_out = p.fcpp + spare;
    }
  }
}

static int run_fused(const Plan3& p, char* ws, const double* X, i64 ldx, i64 x_bs) {
  cudaStream_t st = ctx()->stream;
  fused::Args a;
  a.T = p.T; a.N = p.N; a.npad = p.npad; a.tpad = p.tpad;
  a.L = (int)p.L; a.G = p.fG; a.R = p.fR; a.batch = (int)p.batch; a.cpp = p.fcpp;
  a.ntile = cdiv(p.N, fused::CW);
  a.zc = reinterpret_cast<const double*>(ws + p.o_zc);
  a.w0 = reinterpret_cast<double*>(ws + p.o_w0);
  a.invsd = reinterpret_cast<double*>(ws + p.o_invsd);
  a.mun = reinterpret_cast<double*>(ws + p.o_mun);
  a.blockpart = reinterpret_cast<double*>(ws + p.o_bp);
  a.pxp = reinterpret_cast<double*>(ws + p.o_pxp);
  a.h0p = reinterpret_cast<double*>(ws + p.o_h0p);
  a.dbg = g_fused_dbg;
  a.started = nullptr;
  CUtensorMap mx, m0;
  const i64 bs = p.batch > 1 ? x_bs : p.T * ldx;
  cuuint64_t dims[3] = {(cuuint64_t)p.N, (cuuint64_t)p.T, (cuuint64_t)p.batch};
  cuuint64_t strides[2] = {(cuuint64_t)ldx * 8, (cuuint64_t)bs * 8};
  cuuint32_t es[3] = {1, 1, 1};
  cuuint32_t box[3] = {(cuuint32_t)fused::CW, (cuuint32_t)fused::BOX_ROWS, 1};
  cuuint32_t box0[3] = {(cuuint32_t)fused::CW, 1, 1};
  fused::EncodeTiledFn enc = fused::encode_fn();
  CUresult r1 = enc(&mx, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(X), dims, strides, box, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CUresult r2 = enc(&m0, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(X), dims, strides, box0, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS) return fail(UM_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d, %d)", (int)r1, (int)r2);
  i64 nclu = (i64)p.batch * p.fcpp;
  if (nclu > p.fncl) nclu = p.fncl;
  a.tile_lo = 0;
  a.cpp_out = p.fcpp_out;
  a.cidx_off = 0;
  auto launch_grid = [&](const fused::Args& ga_args, i64 groups, cudaStream_t gst, bool dependent) -> int {
    gridk::GArgs ga;
    ga.a = ga_args;
    ga.part = reinterpret_cast<uint4*>(ws + p.o_gpart);
    ga.rec = reinterpret_cast<uint4*>(ws + p.o_grec);
    if (!dependent)   // tags restart at 1 with every launch: clear the rings of the groups that run
      UM_CUDA(cudaMemsetAsync(ws + p.o_gpart, 0, (p.o_grec - p.o_gpart) + (size_t)groups * gridk::NX * gridk::REC_ENTRY * 16, gst));
    cudaLaunchConfig_t gc = {};
    gc.gridDim = dim3((unsigned)(groups * p.fG), 1, 1);
    gc.blockDim = dim3(gridk::G_NTHREADS, 1, 1);
    gc.dynamicSmemBytes = gridk::GSMEM_ALLOC;
    gc.stream = gst;
    cudaLaunchAttribute ga_at[1];
    ga_at[0].id = cudaLaunchAttributeCooperative;       // every CTA co-resident: they wait on one another through L2
    ga_at[0].val.cooperative = 1;
    gc.attrs = ga_at;
    gc.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&gc, gridk::grid_kernel_for(p.fR), mx, m0, ga);
    if (e != cudaSuccess) return fail(UM_ERR_CUDA, "grid-wide 3PRF sweep launch failed: %s", cudaGetErrorString(e));
    UM_LAUNCHED();
    return UM_OK;
  };
  if (p.grid_sweep) {
    prof_begin(0);
    int s = launch_grid(a, nclu, st, false);
    prof_end(0);
    return s;
  }
  // split sweep: the last `tiles_g` column tiles go to virtual groups on the SMs the clusters leave idle
  static thread_local cudaStream_t st2 = nullptr;
  static thread_local cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  static thread_local unsigned* started = nullptr;   // device word: clusters resident in the current launch
  fused::Args a2 = a;
  a.started = nullptr;
  a2.started = nullptr;
  bool split = p.split_groups > 0 && a.ntile >= 32 * (i64)(p.fncl + p.split_groups) && stream_wait_fn() != nullptr;
  if (split) {
    if (!st2) {
      UM_CUDA(cudaStreamCreateWithFlags(&st2, cudaStreamNonBlocking));
      UM_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
      UM_CUDA(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
      UM_CUDA(cudaMalloc(&started, 256));
    }
    a.started = started;
    const double wg = p.split_groups * g_split_share;
    i64 tiles_g = (i64)((double)a.ntile * wg / ((double)p.fncl + wg) + 0.5);
    if (tiles_g < p.split_groups) tiles_g = p.split_groups;
    a2.tile_lo = a.ntile - tiles_g;
    a2.cpp = p.split_groups;
    a2.cidx_off = p.fcpp;
    a.ntile = a2.tile_lo;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(nclu * p.fG), 1, 1);
  cfg.blockDim = dim3((unsigned)cluster_threads(), 1, 1);
  cfg.dynamicSmemBytes = (size_t)cluster_smem();
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)p.fG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  prof_begin(0);
  if (split) {
    UM_CUDA(cudaMemsetAsync(started, 0, 4, st));
    UM_CUDA(cudaEventRecord(ev_fork, st));          // everything before the sweep (prep, the counter reset)
    UM_CUDA(cudaStreamWaitEvent(st2, ev_fork, 0));
  }
  cudaError_t e = cudaLaunchKernelEx(&cfg, cluster_kernel_for(p.fR), mx, m0, a);
  if (e != cudaSuccess) return fail(UM_ERR_CUDA, "fused 3PRF sweep launch failed: %s", cudaGetErrorString(e));
  UM_LAUNCHED();
  if (split) {
    // second stream: held back by a stream memory wait until every cluster has reported itself resident, so
    // the cooperative launch is placed on exactly the SMs the clusters leave idle
    CUresult wr = stream_wait_fn()(reinterpret_cast<CUstream>(st2), reinterpret_cast<CUdeviceptr>(started), (cuuint32_t)nclu,
                                   CU_STREAM_WAIT_VALUE_GEQ);
    if (wr != CUDA_SUCCESS) return fail(UM_ERR_CUDA, "cuStreamWaitValue32 failed (%d)", (int)wr);
    int s2 = launch_grid(a2, p.split_groups, st2, false);
    if (s2 != UM_OK) return s2;
    UM_CUDA(cudaEventRecord(ev_join, st2));
    UM_CUDA(cudaStreamWaitEvent(st, ev_join, 0));
  }
  prof_end(0);
  return UM_OK;
}

// sweeps over one resident X block (T x N, row pitch ldx), writing / accumulating the packed buffer.
// `ws` is the carved workspace; prep must already have run.
static int run_sweeps(const Plan3& p, char* ws, const double* X, i64 ldx, i64 x_bs, int accumulate, bool gather = true) {
  cudaStream_t st = ctx()->stream;
  double* zc = reinterpret_cast<double*>(ws + p.o_zc);
  double* qp = reinterpret_cast<double*>(ws + p.o_qp);
  double* s1 = reinterpret_cast<double*>(ws + p.o_s1);
  double* s2 = reinterpret_cast<double*>(ws + p.o_s2);
  double* w0 = reinterpret_cast<double*>(ws + p.o_w0);
  double* v = reinterpret_cast<double*>(ws + p.o_v);
  double* invsd = reinterpret_cast<double*>(ws + p.o_invsd);
  double* mun = reinterpret_cast<double*>(ws + p.o_mun);
  double* bp = reinterpret_cast<double*>(ws + p.o_bp);
  double* pxp = reinterpret_cast<double*>(ws + p.o_pxp);
  double* h0p = reinterpret_cast<double*>(ws + p.o_h0p);
  double* packed = reinterpret_cast<double*>(ws + p.o_packed);
  const bool vec16 = aligned16(X) && (ldx % 2 == 0) && (x_bs % 2 == 0);
  if (p.fused) {
    int s = run_fused(p, ws, X, ldx, x_bs);
    if (s != UM_OK) return s;
    if (!p.grid_sweep && fused::CW / p.fG > 4) {   // cluster sweep, small groups: the finalisers left the small sums to this kernel
      i64 nb = cdiv(p.N, 256 * 4);
      if (nb > p.cf_blocks) nb = p.cf_blocks;
      k_tprf_smallsums<<<dim3((unsigned)nb, 1, (unsigned)p.batch), 256, 0, st>>>(w0, mun, p.N, p.npad, p.cf_blocks, bp);
      UM_LAUNCHED();
    }
  } else {
  {
    dim3 grid((unsigned)p.strips, (unsigned)p.rchunks, (unsigned)p.batch);
    prof_begin(0);
    if (vec16) k_tprf_colstats<true><<<grid, C1_THREADS, C1_SMEM, st>>>(X, ldx, x_bs, zc, p.T, p.N, p.npad, p.rows_per_chunk, qp, s1, s2);
    else k_tprf_colstats<false><<<grid, C1_THREADS, C1_SMEM, st>>>(X, ldx, x_bs, zc, p.T, p.N, p.npad, p.rows_per_chunk, qp, s1, s2);
    prof_end(0);
    UM_LAUNCHED();
  }
  {
    dim3 grid((unsigned)p.cf_blocks, 1, (unsigned)p.batch);
    prof_begin(1);
    k_tprf_colfinal<<<grid, 256, 0, st>>>(X, x_bs, qp, s1, s2, (int)p.rchunks, p.T, p.N, p.npad, w0, v, invsd, mun, bp);
    prof_end(1);
    UM_LAUNCHED();
  }
  {
    dim3 grid((unsigned)p.rblocks, (unsigned)p.cchunks, (unsigned)p.batch);
    prof_begin(2);
    if (vec16) k_tprf_project<true><<<grid, C2_THREADS, C2_SMEM, st>>>(X, ldx, x_bs, v, invsd, p.T, p.N, p.npad, p.cols_per_chunk, p.tpad, pxp, h0p);
    else k_tprf_project<false><<<grid, C2_THREADS, C2_SMEM, st>>>(X, ldx, x_bs, v, invsd, p.T, p.N, p.npad, p.cols_per_chunk, p.tpad, pxp, h0p);
    prof_end(2);
    UM_LAUNCHED();
  }
  }
  if (gather) {
    i64 gx = cdiv(p.packed_len, 256);
    if (gx > 1024) gx = 1024;
    dim3 grid((unsigned)gx, 1, (unsigned)p.batch);
    prof_begin(3);
    k_tprf_gather<<<grid, 256, 0, st>>>(pxp, h0p, (int)p.cchunks, p.T, p.tpad, bp, p.cf_blocks, packed, p.packed_len, accumulate);
    prof_end(3);
    UM_LAUNCHED();
  }
  return UM_OK;
}

static int run_prep(const Plan3& p, char* ws, const double* Z, i64 z_rs, i64 z_cs, i64 z_bs, const double* y, i64 y_s,
                    i64 y_bs) {
  dim3 grid(1, 1, (unsigned)p.batch);
    prof_begin(7);
  double* zc_out = reinterpret_cast<double*>(ws + p.o_zc);
  double* small_out = reinterpret_cast<double*>(ws + p.o_small);
  if (p.batch == 1 && p.T >= 2048) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(FIN_CS, 1, 1);
    cfg.blockDim = dim3(256, 1, 1);
    cfg.stream = ctx()->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = FIN_CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    const int Li = (int)p.L;
    UM_CUDA(cudaLaunchKernelEx(&cfg, k_tprf_prep<FIN_CS>, Z, z_rs, z_cs, z_bs, y, y_s, y_bs, p.T, Li, zc_out, small_out));
  } else {
    k_tprf_prep<1><<<grid, 256, 0, ctx()->stream>>>(Z, z_rs, z_cs, z_bs, y, y_s, y_bs, p.T, (int)p.L, zc_out, small_out);
  }
  prof_end(7);
    UM_LAUNCHED();
  return UM_OK;
}

static int run_finish(const Plan3& p, char* ws, int mode, const double* y, i64 y_s, i64 y_bs, double n_total,
                      double* forecasts, double* residuals, i64 f_bs, double* sigma, double* beta3, double* r2_dev,
                      bool want3 = false, double* host_res = nullptr) {
  FinishArgs g;
  g.host_res = host_res;
  g.want3 = (mode == UM_TPRF_CLOSED && want3) ? 1 : 0;
  g.packed = reinterpret_cast<double*>(ws + p.o_packed);
  g.packed_stride = p.packed_len;
  g.small = reinterpret_cast<double*>(ws + p.o_small);
  g.y = y; g.y_s = y_s; g.y_bs = y_bs;
  g.T = p.T; g.L = (int)p.L; g.n_total = n_total; g.mode = mode;
  g.forecasts = forecasts; g.residuals = residuals; g.f_bs = f_bs;
  g.sigma = sigma; g.beta3 = beta3;
  g.sigma_ws = reinterpret_cast<double*>(ws + p.o_sig);
  g.params = reinterpret_cast<double*>(ws + p.o_par);
  g.r2_out = r2_dev;
    prof_begin(4);
  if (p.batch == 1 && p.T >= 2048) {
    // one panel: a cluster of FIN_CS CTAs, each a slice of the rows (the kernel is latency-bound on its
    // row loops: 111 -> ~30 us at T = 8192, which is what the 8-GPU scaling of a 0.5 ms sweep needs)
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(FIN_CS, 1, 1);
    cfg.blockDim = dim3(256, 1, 1);
    cfg.stream = ctx()->stream;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = FIN_CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // launch latency hidden behind the gather / exchange
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 2;
    if (mode == UM_TPRF_CLOSED && !g.want3) UM_CUDA(cudaLaunchKernelEx(&cfg, k_tprf_finish<FIN_CS, false>, g));
    else UM_CUDA(cudaLaunchKernelEx(&cfg, k_tprf_finish<FIN_CS, true>, g));
  } else {
    if (mode == UM_TPRF_CLOSED && !g.want3) k_tprf_finish<1, false><<<(unsigned)p.batch, 256, 0, ctx()->stream>>>(g);
    else k_tprf_finish<1, true><<<(unsigned)p.batch, 256, 0, ctx()->stream>>>(g);
  }
  prof_end(4);
    UM_LAUNCHED();
  return UM_OK;
}

static int run_outputs(const Plan3& p, char* ws, int mode, double* alpha, double* phi, double* phi_hat, double* xstd,
                       i64 out_bs) {
  if (!alpha && !phi && !phi_hat && !xstd) return UM_OK;
  i64 gx = cdiv(p.N, 256);
  if (gx > 2048) gx = 2048;
  dim3 grid((unsigned)gx, 1, (unsigned)p.batch);
    prof_begin(5);
  k_tprf_outputs<<<grid, 256, 0, ctx()->stream>>>(reinterpret_cast<double*>(ws + p.o_w0), reinterpret_cast<double*>(ws + p.o_invsd),
                                                   reinterpret_cast<double*>(ws + p.o_mun), reinterpret_cast<double*>(ws + p.o_par),
                                                   p.N, p.npad, (int)p.L, mode, alpha, phi, phi_hat, xstd, out_bs);
  prof_end(5);
    UM_LAUNCHED();
  return UM_OK;
}

// 64 bytes of mapped pinned host memory per host thread: the finish kernel of a single-panel fit stores r2 / singular
// there, so the host reads them after the stream synchronisation without a separate device-to-host copy
static int host_result_block(double** host, double** dev) {
  static thread_local double* h = nullptr;
  static thread_local double* d = nullptr;
  if (!h) {
    UM_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&h), 64, cudaHostAllocMapped));
    UM_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&d), h, 0));
  }
  *host = h; *dev = d;
  return UM_OK;
}

// L > LP: the accumulators are column-separable in Z, so the fused sweep runs once per group of 8 proxies (X is read
// ceil(L / 8) times) and its W0 / PX blocks are copied out of the workspace before anything else may reuse it
static int fit_wide(int mode, const um_view* y, const um_view* X, const um_view* Z, double n_total, bool sharded, um_tprf_out* out) {
  const i64 T = X->rows, N = X->cols, L = Z->cols;
  const double* xp = static_cast<const double*>(X->data) + X->offset;
  const i64 ldx = N == 1 ? 1 : X->rs;
  const double* yp = static_cast<const double*>(y->data) + y->offset;
  const double* zp = static_cast<const double*>(Z->data) + Z->offset;
  auto even = [](i64 n) { return (n + 1) & ~i64(1); };
  const i64 n_w0 = even(N * L), n_px = even(T * L), n_h0 = even(T), n_v = even(N);
  void* blk = nullptr;
  int s = um_malloc((size_t)(n_w0 + n_px + n_h0 + 2 * n_v) * 8, &blk);
  if (s != UM_OK) return s;
  struct Free { void* p; ~Free() { um_free(p); } } guard{blk};
  double* base = static_cast<double*>(blk);
  TprfWideAcc acc;
  um_view_init(&acc.W0, base, N, L, UM_F64);
  um_view_init(&acc.PX, base + n_w0, T, L, UM_F64);
  um_view_init(&acc.h0, base + n_w0 + n_px, T, 1, UM_F64);
  um_view_init(&acc.isd, base + n_w0 + n_px + n_h0, 1, N, UM_F64);
  um_view_init(&acc.m, base + n_w0 + n_px + n_h0 + n_v, 1, N, UM_F64);
  s = set_smem_attrs();
  if (s != UM_OK) return s;
  auto strided = [](double* p, i64 rows, i64 cols, i64 rs) {
    um_view v;
    um_view_init(&v, p, rows, cols, UM_F64);
    v.rs = rs;
    return v;
  };
  for (i64 g0 = 0; g0 < L; g0 += LP) {
    const i64 Lg = L - g0 < LP ? L - g0 : LP;
    Plan3 p = make_plan_for(T, N, Lg, 1, xp, ldx, 0);
    char* ws = static_cast<char*>(scratch(p.total));
    if (!ws) return UM_ERR_OOM;
    s = run_prep(p, ws, zp + g0 * Z->cs, Z->rs, Z->cs, 0, yp, y->rs, 0);
    if (s != UM_OK) return s;
    s = run_sweeps(p, ws, xp, ldx, 0, 0);
    if (s != UM_OK) return s;
    double* packed = reinterpret_cast<double*>(ws + p.o_packed);
    const um_view w_src = strided(reinterpret_cast<double*>(ws + p.o_w0), N, Lg, LP);
    const um_view p_src = strided(packed, T, Lg, LP);
    um_view w_dst, p_dst;
    um_view_slice(&acc.W0, 0, N, g0, g0 + Lg, &w_dst, nullptr);
    um_view_slice(&acc.PX, 0, T, g0, g0 + Lg, &p_dst, nullptr);
    s = um_copy(&w_src, &w_dst);
    if (s == UM_OK) s = um_copy(&p_src, &p_dst);
    if (s == UM_OK && g0 == 0) {
      const um_view h_src = strided(packed + T * LP, T, 1, 1);
      const um_view i_src = strided(reinterpret_cast<double*>(ws + p.o_invsd), 1, N, N);
      const um_view m_src = strided(reinterpret_cast<double*>(ws + p.o_mun), 1, N, N);
      s = um_copy(&h_src, &acc.h0);
      if (s == UM_OK) s = um_copy(&i_src, &acc.isd);
      if (s == UM_OK) s = um_copy(&m_src, &acc.m);
    }
    if (s != UM_OK) return s;
  }
  s = tprf_fit_wide(mode, y, Z, N, acc, n_total, sharded, out);
  if (s == UM_OK && sharded) s = comm_check_error();   // tprf_fit_wide ends with a stream synchronisation
  return s;
}

static int fill_nan(double* p, i64 n) {
  if (!p || n == 0) return UM_OK;
  k_fill_nan<<<(unsigned)cdiv(n, 256), 256, 0, ctx()->stream>>>(p, n);
  UM_LAUNCHED();
  return UM_OK;
}

}  // namespace um

using namespace um;

extern "C" {

int32_t um_tprf_set_path(int32_t path) {
  if (path < 0 || path > 3) return fail(UM_ERR_ARG, "3PRF path %d outside {0 auto, 1 two-sweep, 2 cluster sweep, 3 grid sweep}", path);
  g_tprf_path.store(path);
  return UM_OK;
}

int32_t um_tprf_debug_stamps(long long* device_buf) {
  g_fused_dbg = device_buf;
  return UM_OK;
}

int32_t um_tprf_fused_info(int64_t T, int32_t* cluster_size, int32_t* rows_per_piece, int32_t* resident_clusters) {
  UM_CTX();
  int G = 0, R = 0;
  fused::pick_shape(T, &G, &R);
  if (cluster_size) *cluster_size = G;
  if (rows_per_piece) *rows_per_piece = R;
  if (resident_clusters) {
    const int path = g_tprf_path.load();
    int n = 0;
    if (G && fused::encode_fn()) {
      if (path == 3) n = grid_groups(G, R);
      if (n <= 0) n = fused_clusters(G, R);
    }
    *resident_clusters = n;
  }
  return UM_OK;
}

int32_t um_tprf_work(int64_t T, int64_t N, int64_t L, double* flops, double* bytes) {
  // X'[Z] 2TNL, Xc W0 2TNL, row sums + stats ~6TN -> T*N*(4L+6); X read once = 8TN bytes
  if (flops) *flops = (double)T * (double)N * (4.0 * (double)L + 6.0);
  if (bytes) *bytes = 8.0 * (double)T * (double)N;
  return UM_OK;
}

int32_t um_tprf_fit(int32_t mode, const um_view* y, const um_view* X, const um_view* Z, int64_t n_total, um_tprf_out* out) {
  UM_CTX();
  UM_VIEW(y); UM_VIEW(X); UM_VIEW(Z);
  if (!out) return fail(UM_ERR_ARG, "null out");
  UM_REQUIRE(mode >= UM_TPRF_CLOSED && mode <= UM_TPRF_WINDOW, "unknown 3PRF mode %d", mode);
  UM_REQUIRE(X->dtype == UM_F64 && y->dtype == UM_F64 && Z->dtype == UM_F64, "3PRF needs f64 inputs");
  const i64 T = X->rows, N = X->cols, L = Z->cols;
  UM_REQUIRE(y->rows == T && y->cols == 1, "y must be %lld x 1, got %lldx%lld", T, (i64)y->rows, (i64)y->cols);
  UM_REQUIRE(Z->rows == T, "Z must have %lld rows, got %lld", T, (i64)Z->rows);
  UM_REQUIRE(L >= 1, "number of proxies L = %lld < 1", L);
  UM_REQUIRE(T >= 1 && N >= 1, "empty panel");
  UM_REQUIRE(X->cs == 1 || N == 1, "X must be row-contiguous (cs == 1); materialise the view first");
  // n_total > 0 declares a column-sharded fit (partials are summed over the communicator); n_total <= 0 is a fit of
  // the panel this rank holds, communicator or not (replicated rows must never be summed across ranks)
  const bool sharded = n_total > 0 && comm_active() && comm_world() > 1;
  if (n_total <= 0) n_total = N;
  out->r_squared = nan("");
  out->singular = 0;
  if (mode == UM_TPRF_IS_FULL && T < 10) {
    // T < minObs => all-NaN forecasts (Tprf3.scala:893-894)
    // every requested output is defined on return (the reference has no fit to report either)
    int s = fill_nan(out->forecasts, T);
    if (s == UM_OK) s = fill_nan(out->residuals, T);
    if (s == UM_OK) s = fill_nan(out->alpha, N);
    if (s == UM_OK) s = fill_nan(out->phi, N * L);
    if (s == UM_OK) s = fill_nan(out->phi_hat, N * (L + 1));
    if (s == UM_OK) s = fill_nan(out->sigma, T * L);
    if (s == UM_OK) s = fill_nan(out->beta3, L + 1);
    if (s == UM_OK) s = fill_nan(out->xstd, N);
    if (s != UM_OK) return s;
    UM_CUDA(cudaStreamSynchronize(ctx()->stream));
    return UM_OK;
  }
  if (L > LP) {
    // more proxies than one DMMA fragment holds: one fused sweep per group of 8, finish by operators (tprf_wide.cu)
    const int sw = fit_wide(mode, y, X, Z, (double)n_total, sharded, out);
    if (sw == UM_ERR_SINGULAR) {
      out->singular = 1;
      int sf = fill_nan(out->forecasts, T);
      if (sf == UM_OK) sf = fill_nan(out->residuals, T);
      if (sf == UM_OK) UM_CUDA(cudaStreamSynchronize(ctx()->stream));
      return fail(UM_ERR_SINGULAR, "Matrix is singular or nearly singular");
    }
    return sw;
  }
  const double* xp = static_cast<const double*>(X->data) + X->offset;
  const i64 ldx = N == 1 ? 1 : X->rs;
  Plan3 p = make_plan_for(T, N, L, 1, xp, ldx, 0);
  if (g_tprf_path.load() >= 2 && !p.fused)
    return fail(UM_ERR_ARG, "fused 3PRF sweep requested but not applicable (T = %lld > %d*%d rows, X not 16-byte aligned / odd row pitch, or no cluster launch)", T, fused::GMAX, fused::RMAX);
  char* ws = static_cast<char*>(scratch(p.total));
  if (!ws) return UM_ERR_OOM;
  int s = set_smem_attrs();
  if (s != UM_OK) return s;
  const double* yp = static_cast<const double*>(y->data) + y->offset;
  const double* zp = static_cast<const double*>(Z->data) + Z->offset;
  s = run_prep(p, ws, zp, Z->rs, Z->cs, 0, yp, y->rs, 0);
  if (s != UM_OK) return s;
  // (gathering the chunk partials inside the finish kernel, or inside the exchange kernel, was tried in round 2: the
  // 8- and 32-CTA kernels take longer over it than the 1024-CTA gather launch they save)
  s = run_sweeps(p, ws, xp, ldx, 0, 0);
  if (s != UM_OK) return s;
  if (sharded) {
    prof_begin(6);
    s = comm_allreduce_sum_f64_impl(reinterpret_cast<double*>(ws + p.o_packed), p.packed_len);
    prof_end(6);
    if (s != UM_OK) return s;
  }
  const bool want3 = out->sigma || out->beta3 || out->phi || out->phi_hat;
  double *res = nullptr, *res_dev = nullptr;
  s = host_result_block(&res, &res_dev);
  if (s != UM_OK) return s;
  s = run_finish(p, ws, mode, yp, y->rs, 0, (double)n_total, out->forecasts, out->residuals, 0, out->sigma, out->beta3, nullptr, want3, res_dev);
  if (s != UM_OK) return s;
  s = run_outputs(p, ws, mode, out->alpha, out->phi, out->phi_hat, out->xstd, N);
  if (s != UM_OK) return s;
  UM_CUDA(cudaStreamSynchronize(ctx()->stream));
  out->r_squared = res[0];
  out->singular = res[1] != 0.0;
  if (comm_active()) {
    const int cs = comm_check_error();
    if (cs != UM_OK) return cs;
  }
  // an exact-zero pivot in any of the small inverses: the reference's `.inverse` throws (Mat.scala:990); the
  // outputs that depended on it hold NaN
  if (out->singular) return fail(UM_ERR_SINGULAR, "Matrix is singular or nearly singular");
  return UM_OK;
}

static int fit_batched_chunk(int32_t mode, const double* y, const double* X, const double* Z, int64_t T, int64_t N, int64_t L,
                             int64_t batch, double* forecasts, double* alpha, double* r2) {
  Plan3 p = make_plan_for(T, N, L, batch, X, N, T * N);
  char* ws = static_cast<char*>(scratch(p.total));
  if (!ws) return UM_ERR_OOM;
  int s = set_smem_attrs();
  if (s != UM_OK) return s;
  s = run_prep(p, ws, Z, L, 1, T * L, y, 1, T);
  if (s != UM_OK) return s;
  s = run_sweeps(p, ws, X, N, T * N, 0);
  if (s != UM_OK) return s;
  s = run_finish(p, ws, mode, y, 1, T, (double)N, forecasts, nullptr, T, nullptr, nullptr, r2);
  if (s != UM_OK) return s;
  return run_outputs(p, ws, mode, alpha, nullptr, nullptr, nullptr, N);
}

int32_t um_tprf_fit_batched(int32_t mode, const double* y, const double* X, const double* Z, int64_t T, int64_t N, int64_t L,
                            int64_t batch, double* forecasts, double* alpha, double* r2) {
  UM_CTX();
  UM_REQUIRE(mode >= UM_TPRF_CLOSED && mode <= UM_TPRF_WINDOW, "unknown 3PRF mode %d", mode);
  UM_REQUIRE(y && X && Z, "null input");
  UM_REQUIRE(L >= 1, "number of proxies L = %lld < 1", (i64)L);
  UM_REQUIRE(T >= 1 && N >= 1 && batch >= 0, "bad panel shape");
  if (L > LP) {
    // more than 8 proxies: panel by panel through um_tprf_fit (one sweep per group of 8 proxies + operator finish)
    for (int64_t b = 0; b < batch; ++b) {
      um_view yv, xv, zv;
      um_view_init(&yv, const_cast<double*>(y) + b * T, T, 1, UM_F64);
      um_view_init(&xv, const_cast<double*>(X) + b * T * N, T, N, UM_F64);
      um_view_init(&zv, const_cast<double*>(Z) + b * T * L, T, L, UM_F64);
      um_tprf_out o = {};
      o.forecasts = forecasts ? forecasts + b * T : nullptr;
      o.alpha = alpha ? alpha + b * N : nullptr;
      const int s = um_tprf_fit(mode, &yv, &xv, &zv, 0, &o);
      if (s != UM_OK) return s;
      if (r2) UM_CUDA(cudaMemcpy(r2 + b, &o.r_squared, 8, cudaMemcpyHostToDevice));
    }
    return UM_OK;
  }
  // the panel index is a grid dimension (<= 65535): any number of panels goes through in slices of that, in stream order
  constexpr int64_t SLICE = 65535;
  for (int64_t b0 = 0; b0 < batch; b0 += SLICE) {
    const int64_t nb = batch - b0 < SLICE ? batch - b0 : SLICE;
    const int s = fit_batched_chunk(mode, y + b0 * T, X + b0 * T * N, Z + b0 * T * L, T, N, L, nb, forecasts ? forecasts + b0 * T : nullptr,
                                    alpha ? alpha + b0 * N : nullptr, r2 ? r2 + b0 : nullptr);
    if (s != UM_OK) return s;
  }
  return UM_OK;
}

int32_t um_tprf_fit_host(int32_t mode, const double* y, const double* X, int64_t ldx, const double* Z, int64_t T,
                         int64_t N_local, int64_t L, int64_t n_total, double* forecasts_host, double* alpha_host,
                         double* r_squared) {
  UM_CTX();
  UM_REQUIRE(mode >= UM_TPRF_CLOSED && mode <= UM_TPRF_WINDOW, "unknown 3PRF mode %d", mode);
  UM_REQUIRE(y && X && Z, "null input");
  UM_REQUIRE(L >= 1, "number of proxies L = %lld < 1", (i64)L);
  UM_REQUIRE(T >= 1 && N_local >= 1 && ldx >= N_local, "bad panel shape");
  if (L > LP) {
    // more than 8 proxies need ceil(L / 8) reads of X: the panel is made resident once instead of streamed per group
    void* blk = nullptr;
    const size_t nx = (size_t)T * N_local, nz = (size_t)T * L;
    int s = um_malloc((nx + nz + 2 * (size_t)T + (size_t)N_local) * 8, &blk);
    if (s != UM_OK) return s;
    struct Free { void* p; ~Free() { um_free(p); } } guard{blk};
    double* dx = static_cast<double*>(blk);
    double* dzz = dx + nx;
    double* dyy = dzz + nz;
    double* dff = dyy + T;
    double* daa = dff + T;
    cudaStream_t st = ctx()->stream;
    UM_CUDA(cudaMemcpy2DAsync(dx, (size_t)N_local * 8, X, (size_t)ldx * 8, (size_t)N_local * 8, (size_t)T, cudaMemcpyHostToDevice, st));
    UM_CUDA(cudaMemcpyAsync(dzz, Z, nz * 8, cudaMemcpyHostToDevice, st));
    UM_CUDA(cudaMemcpyAsync(dyy, y, (size_t)T * 8, cudaMemcpyHostToDevice, st));
    um_view yv, xv, zv;
    um_view_init(&yv, dyy, T, 1, UM_F64);
    um_view_init(&xv, dx, T, N_local, UM_F64);
    um_view_init(&zv, dzz, T, L, UM_F64);
    um_tprf_out o = {};
    o.forecasts = dff;
    o.alpha = alpha_host ? daa : nullptr;
    s = um_tprf_fit(mode, &yv, &xv, &zv, n_total, &o);
    if (s != UM_OK) return s;
    if (forecasts_host) UM_CUDA(cudaMemcpyAsync(forecasts_host, dff, (size_t)T * 8, cudaMemcpyDeviceToHost, st));
    if (alpha_host) UM_CUDA(cudaMemcpyAsync(alpha_host, daa, (size_t)N_local * 8, cudaMemcpyDeviceToHost, st));
    UM_CUDA(cudaStreamSynchronize(st));
    if (r_squared) *r_squared = o.r_squared;
    return UM_OK;
  }
  const bool sharded = n_total > 0 && comm_active() && comm_world() > 1;
  if (n_total <= 0) n_total = N_local;
  Ctx& c = *ctx();
  // column blocks of CB columns are staged through two device buffers: H2D of block k+1 (copy
  // stream) overlaps the two sweeps of block k (compute stream).
  i64 CB = 8192;
  if (CB > N_local) CB = cdiv(N_local, 2) * 2;
  const i64 nblocks = cdiv(N_local, CB);
  Plan3 p = make_plan_for(T, CB, L, 1, nullptr, CB, 0);  // staging buffers are 256-byte aligned
  const size_t stage_bytes = al((size_t)T * CB * 8);
  const size_t w0_all = al((size_t)N_local * LP * 8);  // keep every block's W0 for alpha
  const size_t yz_bytes = al((size_t)T * (L + 1) * 8) + al((size_t)T * 8);
  char* base = static_cast<char*>(scratch(p.total + 2 * stage_bytes + w0_all + yz_bytes + al((size_t)N_local * 8)));
  if (!base) return UM_ERR_OOM;
  char* ws = base;
  double* stage_buf[2] = {reinterpret_cast<double*>(base + p.total), reinterpret_cast<double*>(base + p.total + stage_bytes)};
  double* w0_keep = reinterpret_cast<double*>(base + p.total + 2 * stage_bytes);
  double* dz = reinterpret_cast<double*>(base + p.total + 2 * stage_bytes + w0_all);
  double* dy = dz + T * L;
  double* dfc = reinterpret_cast<double*>(base + p.total + 2 * stage_bytes + w0_all + al((size_t)T * (L + 1) * 8));
  double* dalpha = reinterpret_cast<double*>(base + p.total + 2 * stage_bytes + w0_all + yz_bytes);
  int s = set_smem_attrs();
  if (s != UM_OK) return s;
  static thread_local cudaStream_t copy_st = nullptr;
  static thread_local cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr};
  if (!copy_st) {
    UM_CUDA(cudaStreamCreateWithFlags(&copy_st, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      UM_CUDA(cudaEventCreateWithFlags(&ev_copied[i], cudaEventDisableTiming));
      UM_CUDA(cudaEventCreateWithFlags(&ev_free[i], cudaEventDisableTiming));
    }
  }
  cudaStream_t st = c.stream;
  UM_CUDA(cudaMemcpyAsync(dz, Z, (size_t)T * L * 8, cudaMemcpyHostToDevice, st));
  UM_CUDA(cudaMemcpyAsync(dy, y, (size_t)T * 8, cudaMemcpyHostToDevice, st));
  s = run_prep(p, ws, dz, L, 1, 0, dy, 1, 0);
  if (s != UM_OK) return s;
  UM_CUDA(cudaEventRecord(ev_free[0], st));
  UM_CUDA(cudaEventRecord(ev_free[1], st));
  for (i64 k = 0; k < nblocks; ++k) {
    const int bsel = (int)(k & 1);
    const i64 c0 = k * CB;
    const i64 w = (c0 + CB <= N_local) ? CB : (N_local - c0);
    UM_CUDA(cudaStreamWaitEvent(copy_st, ev_free[bsel], 0));
    UM_CUDA(cudaMemcpy2DAsync(stage_buf[bsel], (size_t)CB * 8, X + c0, (size_t)ldx * 8, (size_t)w * 8, (size_t)T,
                              cudaMemcpyHostToDevice, copy_st));
    UM_CUDA(cudaEventRecord(ev_copied[bsel], copy_st));
    UM_CUDA(cudaStreamWaitEvent(st, ev_copied[bsel], 0));
    Plan3 pb = p;
    pb.N = w;  // narrower last block; strides / padding stay those of the CB plan
    s = run_sweeps(pb, ws, stage_buf[bsel], CB, 0, k > 0);
    if (s != UM_OK) return s;
    UM_CUDA(cudaMemcpyAsync(w0_keep + c0 * LP, ws + p.o_w0, (size_t)w * LP * 8, cudaMemcpyDeviceToDevice, st));
    UM_CUDA(cudaEventRecord(ev_free[bsel], st));
  }
  if (sharded) {
    s = comm_allreduce_sum_f64_impl(reinterpret_cast<double*>(ws + p.o_packed), p.packed_len);
    if (s != UM_OK) return s;
  }
  s = run_finish(p, ws, mode, dy, 1, 0, (double)n_total, dfc, nullptr, 0, nullptr, nullptr, nullptr);
  if (s != UM_OK) return s;
  if (alpha_host) {
    i64 gx = cdiv(N_local, 256);
    if (gx > 2048) gx = 2048;
    k_tprf_outputs<<<dim3((unsigned)gx, 1, 1), 256, 0, st>>>(w0_keep, nullptr, nullptr, reinterpret_cast<double*>(ws + p.o_par),
                                                              N_local, N_local, (int)L, mode, dalpha, nullptr, nullptr, nullptr, N_local);
    UM_LAUNCHED();
    UM_CUDA(cudaMemcpyAsync(alpha_host, dalpha, (size_t)N_local * 8, cudaMemcpyDeviceToHost, st));
  }
  if (forecasts_host) UM_CUDA(cudaMemcpyAsync(forecasts_host, dfc, (size_t)T * 8, cudaMemcpyDeviceToHost, st));
  double res[2];
  UM_CUDA(cudaMemcpyAsync(res, reinterpret_cast<double*>(ws + p.o_par) + 88, 16, cudaMemcpyDeviceToHost, st));
  UM_CUDA(cudaStreamSynchronize(st));
  if (r_squared) *r_squared = res[0];
  if (comm_active()) {
    const int cs = comm_check_error();
    if (cs != UM_OK) return cs;
  }
  if (res[1] != 0.0) return fail(UM_ERR_SINGULAR, "Matrix is singular or nearly singular");
  return UM_OK;
}

}  // extern "C"
