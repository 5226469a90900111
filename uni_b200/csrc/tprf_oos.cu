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
