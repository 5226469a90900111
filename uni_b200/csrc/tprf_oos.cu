// tprf_oos.cu -- the out-of-sample point forecast of one 3PRF window.
//
// The OOS procedures of estimate3prf (Tprf3.scala:1112-1199) fit a three-pass filter on each window's kept
// rows and forecast ONE held-out row x_t (t3prfTail, Tprf3.scala:913-943):
//     bo    = (x_t * D^-1 * [1 | Phi]) * PtPinv^T          1 x (L+1),  PtP = [1 | Phi]^T [1 | Phi]
//     yhatt = beta_0 + sum_j bo_{j+1} beta_{j+1}
// The window fit itself is um_tprf_fit(UM_TPRF_ITER) on the kept rows (single-read sweep); this kernel takes the
// fit's Phi (N x L), column std (N) and beta (L+1) and folds the N-long products in one launch.
#include "common.cuh"
#include "small_linalg.cuh"
#include <algorithm>
#include <vector>

namespace um {

constexpr int OOS_LP = 8;                     // proxies, as tprf.cu
constexpr int OOS_D = OOS_LP + 1;
constexpr int OOS_TRI = OOS_D * (OOS_D + 1) / 2;
constexpr int OOS_THREADS = 256;

__global__ void __launch_bounds__(OOS_THREADS) k_tprf_oos_forecast(const double* __restrict__ phi, const double* __restrict__ xstd,
                                                                   const double* __restrict__ beta, const double* __restrict__ xt, i64 xt_cs,
                                                                   i64 N, int L, double* __restrict__ out) {
  const int D = L + 1;
  double d[OOS_D], P[OOS_TRI];
#pragma unroll
  for (int a = 0; a < OOS_D; ++a) d[a] = 0.0;
#pragma unroll
  for (int a = 0; a < OOS_TRI; ++a) P[a] = 0.0;
  for (i64 j = threadIdx.x; j < N; j += OOS_THREADS) {
    double v[OOS_D];
    v[0] = 1.0;
#pragma unroll
    for (int a = 1; a < OOS_D; ++a) v[a] = a <= L ? phi[j * L + (a - 1)] : 0.0;
    const double xs = xt[j * xt_cs] * (1.0 / xstd[j]);   // scaleRows(designPhi, invSd): the reciprocal, then a multiply
    int q = 0;
#pragma unroll
    for (int a = 0; a < OOS_D; ++a) {
      d[a] += xs * v[a];
#pragma unroll
      for (int b = a; b < OOS_D; ++b) P[q++] += v[a] * v[b];
    }
  }
  // block reduction in a fixed order: warp shuffles, then the 8 warp partials in warp order
  __shared__ double sm[OOS_THREADS / 32][OOS_D + OOS_TRI];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int a = 0; a < OOS_D; ++a) {
    double x = d[a];
    for (int s = 16; s > 0; s >>= 1) x += __shfl_down_sync(0xffffffffu, x, s);
    if (lane == 0) sm[w][a] = x;
  }
#pragma unroll
  for (int a = 0; a < OOS_TRI; ++a) {
    double x = P[a];
    for (int s = 16; s > 0; s >>= 1) x += __shfl_down_sync(0xffffffffu, x, s);
    if (lane == 0) sm[w][OOS_D + a] = x;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot[OOS_D + OOS_TRI];
    for (int a = 0; a < OOS_D + OOS_TRI; ++a) {
      double x = sm[0][a];
      for (int q = 1; q < OOS_THREADS / 32; ++q) x += sm[q][a];
      tot[a] = x;
    }
    double ptp[OOS_D * OOS_D], inv[OOS_D * OOS_D];
    int q = 0;
    for (int a = 0; a < OOS_D; ++a)
      for (int b = a; b < OOS_D; ++b) {
        const double x = tot[OOS_D + q++];
        if (a < D && b < D) { ptp[a * D + b] = x; ptp[b * D + a] = x; }
      }
    double y = nan("");
    out[1] = 1.0;                              // singular unless the inverse goes through
    if (lu_inverse_seq(ptp, D, inv)) {
      out[1] = 0.0;
      y = beta[0];
      for (int b = 1; b < D; ++b) {
        double bo = 0.0;                       // bo_b = sum_a d_a * PtPinv[b][a]
        for (int a = 0; a < D; ++a) bo += tot[a] * inv[b * D + a];
        y += bo * beta[b];
      }
    }
    out[0] = y;
  }
}

}  // namespace um

using namespace um;

extern "C" int32_t um_tprf_oos_forecast(const um_view* phi, const um_view* xstd, const um_view* beta3, const um_view* xt, double* host_out) {
  UM_CTX();
  UM_VIEW(phi); UM_VIEW(xstd); UM_VIEW(beta3); UM_VIEW(xt);
  UM_REQUIRE(host_out, "null output");
  const i64 N = phi->rows, L = phi->cols;
  UM_REQUIRE(phi->dtype == UM_F64 && xstd->dtype == UM_F64 && beta3->dtype == UM_F64 && xt->dtype == UM_F64, "3PRF needs f64 inputs");
  UM_REQUIRE(L >= 1 && L <= OOS_LP && N >= 1, "phi must be N x L with 1 <= L <= %d", OOS_LP);
  UM_REQUIRE(phi->cs == 1 && phi->rs == L, "phi must be contiguous row-major");
  UM_REQUIRE(xstd->rows * xstd->cols == N && (xstd->cols == 1 ? xstd->rs : xstd->cs) == 1, "xstd must be a contiguous vector of N = %lld", N);
  UM_REQUIRE(beta3->rows * beta3->cols == L + 1 && (beta3->cols == 1 ? (beta3->rs == 1 || L == 0) : beta3->cs == 1), "beta3 must be a contiguous vector of L+1");
  UM_REQUIRE(xt->rows == 1 && xt->cols == N, "the held-out row must be 1 x %lld, got %lldx%lld", N, (i64)xt->rows, (i64)xt->cols);
  double* dev = static_cast<double*>(scratch(2 * sizeof(double)));
  if (!dev) return UM_ERR_OOM;
  k_tprf_oos_forecast<<<1, OOS_THREADS, 0, ctx()->stream>>>(static_cast<const double*>(phi->data) + phi->offset,
                                                            static_cast<const double*>(xstd->data) + xstd->offset,
                                                            static_cast<const double*>(beta3->data) + beta3->offset,
                                                            static_cast<const double*>(xt->data) + xt->offset, xt->cs, N, (int)L, dev);
  UM_LAUNCHED();
  double res[2];
  UM_CUDA(cudaMemcpyAsync(res, dev, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx()->stream));
  UM_CUDA(cudaStreamSynchronize(ctx()->stream));
  *host_out = res[0];
  // [1 | Phi]'[1 | Phi] with an exact-zero pivot: `.inverse` throws in t3prfTail (Tprf3.scala:913-943, Mat.scala:990)
  if (res[1] != 0.0) return fail(UM_ERR_SINGULAR, "Matrix is singular or nearly singular");
  return UM_OK;
}

// =====================================================================================================================
// ALL out-of-sample windows of estimate3prf in ONE launch.
//
// The reference runs its T windows as `.par` tasks over runT3prf (Tprf3.scala:1112-1199), each window a three-pass fit
// on its kept rows plus the point forecast of one held-out row (t3prfFast / t3prfTail, :883-943).  Here one CTA owns one
// window and walks the panel (L2-resident for the panel sizes these procedures are used on) three times:
//   A  per column j (thread per column, rows in order): window mean / sample std (two passes, as stdCols :550-578),
//      dZ'x_j; B1_j = (dZ'dZ)^-1 (dZ'x_j / sd_j); the scaled design row dps_j = [1 | Phi_j] / sd_j -> scratch; [1|Phi]'[1|Phi]
//   C  per kept row (warp per row, lanes over columns): x_t . dps -> Sigma_t; normal equations of pass 3 (nanOls: rows
//      with a NaN in y or Sigma are left out, :929-931)
//   D  beta, then the held-out row: yhatt = beta_0 + sum_l bo_{l+1} beta_{l+1}
// With automatic proxies (Z = Left(L)) the three passes repeat L times, each round's residual column going in FRONT of
// the proxy matrix (:1134-1139).  Every sum has a fixed order (thread-sequential, then a fixed tree), so a window's
// forecast does not depend on the grid.  No per-window launch, no per-window synchronisation.
// =====================================================================================================================
namespace um {

constexpr int OW_THREADS = 256;
constexpr int OW_WARPS = OW_THREADS / 32;

struct OosArgs {
  const double* X; i64 ldx;            // Xn, T x N, row pitch ldx (cs == 1)
  const double* y; i64 y_s;
  const double* Z; i64 z_rs, z_cs;     // proxies (T x L) or null: automatic proxies
  i64 T, N;
  int L;                               // proxy columns (Z given) or the number of automatic proxies
  int autop;                           // 1: automatic proxies (L rounds)
  int kind;                            // 0: the window is rows [lo, hi); 1: every row but [lo, hi)
  int min_obs, nan_mean;
  const int* lo; const int* hi; const int* held;
  double* dps;                         // [window][N][9]
  double* zw;                          // [window][T][8]  automatic proxies: slot 0 = y, slot s = residual of round s-1
  double* sig;                         // [window][T][8]  Sigma rows of the current round (automatic proxies)
  double* fore; double* roll; int* flag;
};

// sum of `nv` per-thread values over the CTA in a fixed order: warp shuffles, then the warp partials in warp order
template <int NV>
__device__ __forceinline__ void ow_block_sum(double (&v)[NV], int nv, double* red /*[OW_WARPS][NV]*/, double* tot /*[NV]*/) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int a = 0; a < NV; ++a) {
    if (a < nv) {
      double x = v[a];
      for (int s = 16; s > 0; s >>= 1) x += __shfl_down_sync(0xffffffffu, x, s);
      if (lane == 0) red[w * NV + a] = x;
    }
  }
  __syncthreads();
  if ((int)threadIdx.x < nv) {
    double x = red[threadIdx.x];
    for (int q = 1; q < OW_WARPS; ++q) x += red[q * NV + threadIdx.x];
    tot[threadIdx.x] = x;
  }
  __syncthreads();
}

// DM = the compiled bound on D = L + 1 (3, 5 or 9): two proxies -- the common case -- need a sixth of the accumulators of
// eight, and the smaller kernels fit several CTAs per SM, which is what hides the L2 latency of the row walks
template <int DM>
__global__ void __launch_bounds__(OW_THREADS, DM <= 3 ? 4 : (DM <= 5 ? 2 : 1)) k_tprf_oos_windows(OosArgs g) {
  constexpr int TRI = DM * (DM + 1) / 2;
  __shared__ double red[OW_WARPS * TRI], tot[TRI + DM + 1];
  __shared__ double mat[DM * DM], inv1[DM * DM], pinv[DM * DM], lux[DM * DM], beta[DM];
  __shared__ double wacc[OW_WARPS][TRI + DM + 1];
  __shared__ double part_raw[OW_THREADS * (1 + DM)];   // pass 1: [slice][column of the sweep][ss | dZ'x]
  __shared__ int piv[DM];
  __shared__ int bad;
  const int w = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int lo = g.lo[w], hi = g.hi[w], held = g.held[w];
  const i64 T = g.T, N = g.N;
  const int Tw = g.kind == 0 ? hi - lo : (int)T - (hi - lo);
  auto row = [&](int i) -> i64 { return g.kind == 0 ? (i64)lo + i : (i < lo ? (i64)i : (i64)i + (hi - lo)); };
  double* dps = g.dps + (size_t)w * N * OOS_D;   // rows of DM <= OOS_D doubles
  double* zw = g.zw ? g.zw + (size_t)w * T * OOS_LP : nullptr;
  double* sg = g.sig ? g.sig + (size_t)w * T * OOS_LP : nullptr;
  if (tid == 0) bad = 0;

  // ---- rolling mean of the window's y (colMean(yt, hasNan), :1145) ----
  {
    double v[2] = {0.0, 0.0};
    for (int i = tid; i < Tw; i += OW_THREADS) {
      const double yv = g.y[row(i) * g.y_s];
      if (!g.nan_mean || yv == yv) { v[0] += yv; v[1] += 1.0; }
    }
    ow_block_sum<2>(v, 2, red, tot);
    if (tid == 0) g.roll[w] = tot[1] > 0.0 ? tot[0] / tot[1] : nan("");
  }
  if (Tw < g.min_obs) {   // T < minObs: NaN fit, NaN forecast (:893-894)
    if (tid == 0) { g.fore[w] = nan(""); g.flag[w] = 0; }
    return;
  }
  if (g.autop) {
    for (int i = tid; i < Tw; i += OW_THREADS) zw[(size_t)i * OOS_LP] = g.y[row(i) * g.y_s];   // r0 = yt
    __syncthreads();
  }
  const int rounds = g.autop ? g.L : 1;
  double yhatt = nan("");
  for (int round = 0; round < rounds; ++round) {
    const int L = g.autop ? round + 1 : g.L, D = L + 1;
    const int ntri = D * (D + 1) / 2;
    // proxy value (window row i, logical column c): given matrix, or the residual slots, newest in front
    auto zval = [&](int i, int c) -> double {
      return g.autop ? zw[(size_t)i * OOS_LP + (round - c)] : g.Z[row(i) * g.z_rs + c * g.z_cs];
    };
    // ---- dZ'dZ -> A1 = (dZ'dZ)^-1 ----
    {
      double v[TRI];
#pragma unroll
      for (int a = 0; a < TRI; ++a) v[a] = 0.0;
      for (int i = tid; i < Tw; i += OW_THREADS) {
        double d[DM];
        d[0] = 1.0;
#pragma unroll
        for (int c = 1; c < DM; ++c) d[c] = c < D ? zval(i, c - 1) : 0.0;
        int q = 0;
#pragma unroll
        for (int a = 0; a < DM; ++a)
#pragma unroll
          for (int b = a; b < DM; ++b) { v[q] += d[a] * d[b]; ++q; }
      }
      ow_block_sum<TRI>(v, TRI, red, tot);
      if (tid == 0) {
        int q = 0;
        for (int a = 0; a < DM; ++a)
          for (int b = a; b < DM; ++b) {
            if (a < D && b < D) { mat[a * D + b] = tot[q]; mat[b * D + a] = tot[q]; }
            ++q;
          }
      }
      __syncthreads();
      if (warp == 0) {
        const bool ok = lu_inverse_warp(mat, D, inv1, lux, piv, lane);
        if (!ok && lane == 0) bad = 1;
      }
      __syncthreads();
    }
    // ---- pass 1 per column: window std, dZ'x, B1, scaled design row; [1|Phi]'[1|Phi] ----
    // Thread (slice s, column c) sums the rows i = s, s + S, ... of its column; the S slice partials are added in slice
    // order.  S = 256 / N slices when the panel is narrower than the CTA (the published 650 x 40 shape: 6 x 40 threads
    // instead of 40), 1 otherwise.  Row loops are unrolled by four: the loads are independent, the latency is L2's.
    {
      double v[TRI];
#pragma unroll
      for (int a = 0; a < TRI; ++a) v[a] = 0.0;
      int S = N >= OW_THREADS ? 1 : (int)(OW_THREADS / N);
      if (S > 8) S = 8;
      const int C = OW_THREADS / S;                       // columns per sweep of the CTA
      const int sl = tid / C, cj = tid % C;
      for (i64 j0 = 0; j0 < N; j0 += C) {
        const i64 j = j0 + cj;
        const bool on = sl < S && j < N;
        const double* xc = g.X + (on ? j : 0);
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        if (on) {
          int i = sl;
          for (; i + 3 * S < Tw; i += 4 * S) {
            s0 += xc[row(i) * g.ldx]; s1 += xc[row(i + S) * g.ldx]; s2 += xc[row(i + 2 * S) * g.ldx]; s3 += xc[row(i + 3 * S) * g.ldx];
          }
          for (; i < Tw; i += S) s0 += xc[row(i) * g.ldx];
        }
        part_raw[((sl) * C + (cj)) * (1 + DM) + (0)] = (s0 + s1) + (s2 + s3);
        __syncthreads();
        double sum = 0.0;
        for (int q = 0; q < S; ++q) sum += part_raw[((q) * C + (cj)) * (1 + DM) + (0)];
        const double mu = Tw > 1 ? sum / (double)Tw : 0.0;
        __syncthreads();
        double ss = 0.0, zx[DM];
#pragma unroll
        for (int c = 0; c < DM; ++c) zx[c] = 0.0;
        if (on) {
          for (int i = sl; i < Tw; i += S) {
            const double x = xc[row(i) * g.ldx];
            const double dd = x - mu;
            ss += dd * dd;
            zx[0] += x;
#pragma unroll
            for (int c = 1; c < DM; ++c)
              if (c < D) zx[c] += zval(i, c - 1) * x;
          }
        }
        part_raw[((sl) * C + (cj)) * (1 + DM) + (0)] = ss;
#pragma unroll
        for (int c = 0; c < DM; ++c) part_raw[((sl) * C + (cj)) * (1 + DM) + (1 + c)] = zx[c];
        __syncthreads();
        if (on && sl == 0) {
          ss = 0.0;
#pragma unroll
          for (int c = 0; c < DM; ++c) zx[c] = 0.0;
          for (int q = 0; q < S; ++q) {
            ss += part_raw[((q) * C + (cj)) * (1 + DM) + (0)];
#pragma unroll
            for (int c = 0; c < DM; ++c) zx[c] += part_raw[((q) * C + (cj)) * (1 + DM) + (1 + c)];
          }
          double isd = 1.0;
          if (Tw > 1) {
            const double sd = sqrt(ss / (double)(Tw - 1));
            isd = sd == 0.0 ? 1.0 : 1.0 / sd;     // sd == 0 -> 1 (:571-574); NaN stays NaN
          }
          double dphi[DM];
          dphi[0] = 1.0;
#pragma unroll
          for (int c = 1; c < DM; ++c) {
            double b1 = 0.0;
            if (c < D)
              for (int k = 0; k < D; ++k) b1 += inv1[c * D + k] * (zx[k] * isd);   // B1 = A1 (dZ'X D^-1), row c
            dphi[c] = c < D ? b1 : 0.0;
          }
#pragma unroll
          for (int c = 0; c < DM; ++c) dps[j * OOS_D + c] = dphi[c] * isd;             // scaleRows(designPhi, invSd)
          int q = 0;
#pragma unroll
          for (int a = 0; a < DM; ++a)
#pragma unroll
            for (int b = a; b < DM; ++b) { v[q] += dphi[a] * dphi[b]; ++q; }
        }
        __syncthreads();
      }
      ow_block_sum<TRI>(v, TRI, red, tot);
      if (tid == 0) {
        int q = 0;
        for (int a = 0; a < DM; ++a)
          for (int b = a; b < DM; ++b) {
            if (a < D && b < D) { mat[a * D + b] = tot[q]; mat[b * D + a] = tot[q]; }
            ++q;
          }
      }
      __syncthreads();   // also publishes dps (global) to the whole CTA
      if (warp == 0) {
        const bool ok = lu_inverse_warp(mat, D, pinv, lux, piv, lane);
        if (!ok && lane == 0) bad = 1;
      }
      __syncthreads();
    }
    // ---- pass 2 per kept row (warp per row) + the normal equations of pass 3 ----
    const int nacc = ntri + D + 1;
    for (int a = lane; a < TRI + DM + 1; a += 32) wacc[warp][a] = 0.0;
    __syncwarp();
    auto row_dot = [&](i64 r, double (&acc)[DM]) {   // x_r . dps, lanes over columns, fixed xor tree
#pragma unroll
      for (int c = 0; c < DM; ++c) acc[c] = 0.0;
      const double* xr = g.X + r * g.ldx;
      for (i64 j = lane; j < N; j += 32) {
        const double x = xr[j];
#pragma unroll
        for (int c = 0; c < DM; ++c)
          if (c < D) acc[c] += x * dps[j * OOS_D + c];
      }
#pragma unroll
      for (int c = 0; c < DM; ++c)
        if (c < D)
          for (int s = 16; s > 0; s >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], s);
    };
    for (int i = warp; i < Tw; i += OW_WARPS) {
      double acc[DM];
      row_dot(row(i), acc);
      if (lane == 0) {
        double xa[DM];
        xa[0] = 1.0;
        bool ok = true;
        for (int l = 1; l < D; ++l) {          // Sigma_t = ((x_t dps) PtPinv')[1:]
          double sgm = 0.0;
          for (int k = 0; k < D; ++k) sgm += acc[k] * pinv[l * D + k];
          xa[l] = sgm;
          ok = ok && sgm == sgm;
          if (sg) sg[(size_t)i * OOS_LP + (l - 1)] = sgm;
        }
        const double yv = g.y[row(i) * g.y_s];
        ok = ok && yv == yv;
        if (ok) {
          int q = 0;
          for (int a = 0; a < D; ++a)
            for (int b = a; b < D; ++b) wacc[warp][q++] += xa[a] * xa[b];
          for (int a = 0; a < D; ++a) wacc[warp][ntri + a] += xa[a] * yv;
          wacc[warp][ntri + D] += 1.0;
        }
      }
    }
    __syncthreads();
    if (warp == 0) {
      for (int a = lane; a < nacc; a += 32) {
        double x = wacc[0][a];
        for (int q = 1; q < OW_WARPS; ++q) x += wacc[q][a];
        tot[a] = x;
      }
      __syncwarp();
      if (lane == 0) {
        int q = 0;
        for (int a = 0; a < D; ++a)
          for (int b = a; b < D; ++b) { mat[a * D + b] = tot[q]; mat[b * D + a] = tot[q]; ++q; }
      }
      __syncwarp();
      const bool none = tot[ntri + D] < 1.0;                       // nanOls(minObs = 1): no usable row -> NaN coefficients
      const bool ok = none ? false : lu_inverse_warp(mat, D, inv1, lux, piv, lane);
      if (!ok && !none && lane == 0) bad = 1;
      if (lane < D) {
        double s = 0.0;
        if (ok)
          for (int k = 0; k < D; ++k) s += inv1[lane * D + k] * tot[ntri + k];
        beta[lane] = ok ? s : nan("");
      }
      __syncwarp();
      // ---- the held-out row (t3prfTail :933-940) ----
      double acc[DM];
      row_dot((i64)held, acc);
      if (lane == 0) {
        double yh = beta[0];
        for (int l = 1; l < D; ++l) {
          double bo = 0.0;
          for (int k = 0; k < D; ++k) bo += acc[k] * pinv[l * D + k];
          yh += bo * beta[l];
        }
        yhatt = yh;
      }
    }
    __syncthreads();
    if (g.autop && round + 1 < rounds) {   // the next proxy column: this round's in-sample residual, in front
      for (int i = tid; i < Tw; i += OW_THREADS) {
        double f = beta[0];
        for (int l = 1; l < D; ++l) f += sg[(size_t)i * OOS_LP + (l - 1)] * beta[l];
        zw[(size_t)i * OOS_LP + (round + 1)] = g.y[row(i) * g.y_s] - f;
      }
      __syncthreads();
    }
  }
  if (tid == 0) {
    g.fore[w] = yhatt;
    g.flag[w] = bad;
  }
}

}  // namespace um

namespace um {
// Xn = X / nanStdCols(X) (Tprf3.scala:469-502, 1078-1081): per column over its non-NaN rows n, mean (0 if n <= 1), sample std,
// n <= 1 or sd == 0 -> 1.  Thread per column, rows in order (coalesced across the warp); one launch.
__global__ void __launch_bounds__(256) k_nanstd_scale(const double* __restrict__ X, i64 ldx, i64 T, i64 N, double* __restrict__ Xn) {
  const i64 j = (i64)blockIdx.x * 256 + threadIdx.x;
  if (j >= N) return;
  double s = 0.0, n = 0.0;
  for (i64 t = 0; t < T; ++t) {
    const double x = X[t * ldx + j];
    if (x == x) { s += x; n += 1.0; }
  }
  const double mu = n > 1.0 ? s / n : 0.0;
  double ss = 0.0;
  for (i64 t = 0; t < T; ++t) {
    const double x = X[t * ldx + j];
    if (x == x) { const double d = x - mu; ss += d * d; }
  }
  double sd = 1.0;
  if (n > 1.0) {
    sd = sqrt(ss / (n - 1.0));
    if (sd == 0.0) sd = 1.0;
  }
  for (i64 t = 0; t < T; ++t) Xn[t * N + j] = X[t * ldx + j] / sd;
}
}  // namespace um

extern "C" int32_t um_tprf_oos_windows(const um_view* y, const um_view* Xn, const um_view* Z, int32_t n_auto, int32_t kind,
                                       const int32_t* lo, const int32_t* hi, const int32_t* held, int64_t nwin, int32_t min_obs,
                                       int32_t standardize, double* fore_host, double* roll_host) {
  const int32_t nan_mean = 1;   // colMean(yt, hasNan): on NaN-free y the NaN-aware mean IS the mean
  UM_CTX();
  UM_VIEW(y); UM_VIEW(Xn);
  UM_REQUIRE(lo && hi && held && fore_host && roll_host, "null window / output array");
  UM_REQUIRE(kind == 0 || kind == 1, "window kind %d outside {0 rows [lo, hi), 1 all rows but [lo, hi)}", kind);
  UM_REQUIRE(Xn->dtype == UM_F64 && y->dtype == UM_F64, "3PRF needs f64 inputs");
  const i64 T = Xn->rows, N = Xn->cols;
  UM_REQUIRE(y->rows == T && y->cols == 1, "y must be %lld x 1", T);
  UM_REQUIRE(Xn->cs == 1 || N == 1, "X must be row-contiguous (cs == 1)");
  UM_REQUIRE(T < (i64(1) << 31) && nwin >= 0, "bad window count / panel height");
  int L;
  if (Z) {
    UM_VIEW(Z);
    UM_REQUIRE(Z->dtype == UM_F64 && Z->rows == T, "Z must be f64 with %lld rows", T);
    L = (int)Z->cols;
  } else {
    L = n_auto;
  }
  UM_REQUIRE(L >= 1 && L <= OOS_LP, "number of proxies L = %d outside [1, %d]", L, OOS_LP);
  if (nwin == 0) return UM_OK;
  for (i64 k = 0; k < nwin; ++k)
    UM_REQUIRE(lo[k] >= 0 && lo[k] <= hi[k] && hi[k] <= T && held[k] >= 0 && held[k] < T, "window %lld [%d, %d) / row %d outside the panel", k, lo[k], hi[k], held[k]);
  // windows in batches that keep the scratch (dps N x 9 per window, + 2 T x 8 with automatic proxies) under ~2 GiB
  const size_t per_win = (size_t)N * OOS_D * 8 + (Z ? 0 : (size_t)T * OOS_LP * 16) + 64;
  i64 batch = (i64)((size_t(2) << 30) / per_win);
  if (batch < 1) batch = 1;
  if (batch > nwin) batch = nwin;
  const size_t meta = ((size_t)batch * (3 * sizeof(int) + 2 * sizeof(double) + sizeof(int)) + 1023) & ~size_t(255);
  const size_t xn_bytes = standardize ? (((size_t)T * N * 8 + 255) & ~size_t(255)) : 0;
  char* ws = static_cast<char*>(scratch(meta + (size_t)batch * per_win + 4096 + xn_bytes));
  if (!ws) return UM_ERR_OOM;
  int* d_lo = reinterpret_cast<int*>(ws);
  int* d_hi = d_lo + batch;
  int* d_held = d_hi + batch;
  int* d_flag = d_held + batch;
  double* d_fore = reinterpret_cast<double*>(ws + (((size_t)batch * 4 * sizeof(int) + 255) & ~size_t(255)));
  double* d_roll = d_fore + batch;
  char* big = ws + meta + 1024;
  OosArgs a;
  a.X = static_cast<const double*>(Xn->data) + Xn->offset; a.ldx = N == 1 ? 1 : Xn->rs;
  if (standardize) {   // the raw panel came in: Xn = X / nanStdCols(X) into scratch, one launch
    double* xn = reinterpret_cast<double*>(ws + meta + (size_t)batch * per_win + 4096);
    k_nanstd_scale<<<(unsigned)((N + 255) / 256), 256, 0, ctx()->stream>>>(a.X, a.ldx, T, N, xn);
    UM_LAUNCHED();
    a.X = xn; a.ldx = N;
  }
  a.y = static_cast<const double*>(y->data) + y->offset; a.y_s = y->rs;
  a.Z = Z ? static_cast<const double*>(Z->data) + Z->offset : nullptr;
  a.z_rs = Z ? Z->rs : 0; a.z_cs = Z ? Z->cs : 0;
  a.T = T; a.N = N; a.L = L; a.autop = Z ? 0 : 1; a.kind = kind; a.min_obs = min_obs; a.nan_mean = nan_mean;
  a.lo = d_lo; a.hi = d_hi; a.held = d_held; a.flag = d_flag; a.fore = d_fore; a.roll = d_roll;
  a.dps = reinterpret_cast<double*>(big);
  a.zw = Z ? nullptr : reinterpret_cast<double*>(big + (size_t)batch * N * OOS_D * 8);
  a.sig = Z ? nullptr : a.zw + (size_t)batch * T * OOS_LP;
  cudaStream_t st = ctx()->stream;
  std::vector<int> flags((size_t)batch);
  int any_bad = 0;
  for (i64 w0 = 0; w0 < nwin; w0 += batch) {
    const i64 nb = std::min(batch, nwin - w0);
    UM_CUDA(cudaMemcpyAsync(d_lo, lo + w0, (size_t)nb * sizeof(int), cudaMemcpyHostToDevice, st));
    UM_CUDA(cudaMemcpyAsync(d_hi, hi + w0, (size_t)nb * sizeof(int), cudaMemcpyHostToDevice, st));
    UM_CUDA(cudaMemcpyAsync(d_held, held + w0, (size_t)nb * sizeof(int), cudaMemcpyHostToDevice, st));
    if (L <= 2) k_tprf_oos_windows<3><<<(unsigned)nb, OW_THREADS, 0, st>>>(a);
    else if (L <= 4) k_tprf_oos_windows<5><<<(unsigned)nb, OW_THREADS, 0, st>>>(a);
    else k_tprf_oos_windows<9><<<(unsigned)nb, OW_THREADS, 0, st>>>(a);
    UM_LAUNCHED();
    UM_CUDA(cudaMemcpyAsync(fore_host + w0, d_fore, (size_t)nb * sizeof(double), cudaMemcpyDeviceToHost, st));
    UM_CUDA(cudaMemcpyAsync(roll_host + w0, d_roll, (size_t)nb * sizeof(double), cudaMemcpyDeviceToHost, st));
    UM_CUDA(cudaMemcpyAsync(flags.data(), d_flag, (size_t)nb * sizeof(int), cudaMemcpyDeviceToHost, st));
    UM_CUDA(cudaStreamSynchronize(st));   // one synchronisation per batch of windows (one in all for ordinary panels)
    for (i64 k = 0; k < nb; ++k) any_bad |= flags[(size_t)k];
  }
  // an exact-zero pivot in any window's small inverse: the reference's `.inverse` throws out of the parallel loop (Mat.scala:990)
  if (any_bad) return fail(UM_ERR_SINGULAR, "Matrix is singular or nearly singular");
  return UM_OK;
}
