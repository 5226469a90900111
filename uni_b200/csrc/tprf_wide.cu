// tprf_wide.cu -- the finish of the 3PRF fits for MORE than 8 proxies (L > LP of tprf.cu).
//
// The reference's fits take any L (Tprf3.scala:199-289: `Z: MatD` of T x L).  The fused sweeps of tprf.cu keep the
// proxies in the N = 8 dimension of one DMMA fragment, and their finish kernel keeps every small system in registers
// and shared memory sized for 8.  Wider proxy sets are rare (Kelly-Pruitt use 1-3, the reference's tests at most 5):
// um_tprf_fit runs the fused sweep once per group of 8 proxies -- the accumulators of DESIGN.md section 4.1 are
// column-separable in Z: W0 and PX group by group, h0 / 1/sd / mu/sd the same in every group -- and hands them to
// this file, which forms the small sums (sw, SWW, smw, smu) and both finishes with the library's own device
// operators (`*@` on the FP64 tensor pipe, axis reductions, broadcast elementwise ops, `.inverse` in the reference's
// LU order).  X is read ceil(L / 8) times.  Results: closed form (Tprf3.scala:199-209), the three OLS passes
// (:211-231,260-289), IS Full (:883-943,1214-1250).  A singular small system surfaces as UM_ERR_SINGULAR exactly where
// the reference's `.inverse` would throw (Mat.scala:990).
#include <cmath>
#include <vector>

#include "common.cuh"

namespace um {

int comm_allreduce_sum_f64_impl(double* buf, i64 count);  // comm.cu

namespace {

#define W_(call)                  \
  do {                            \
    int _s = (call);              \
    if (_s != UM_OK) return _s;   \
  } while (0)

// temporaries from the stream-ordered pool, returned when the fit leaves (frees are stream-ordered too)
struct Pool {
  std::vector<void*> held;
  ~Pool() {
    for (void* p : held) um_free(p);
  }
  int mat(i64 r, i64 c, um_view* v, int dtype = UM_F64) {
    void* p = nullptr;
    size_t bytes = (size_t)r * (size_t)c * dtype_size(dtype);
    int s = um_malloc(bytes ? bytes : 8, &p);
    if (s != UM_OK) return s;
    held.push_back(p);
    return um_view_init(v, p, r, c, dtype);
  }
};

inline um_view tr(const um_view& v) {
  um_view t;
  um_view_transpose(&v, &t);
  return t;
}

inline um_view sl(const um_view& v, i64 r0, i64 r1, i64 c0, i64 c1) {
  um_view s;
  um_view_slice(&v, r0, r1, c0, c1, &s, nullptr);
  return s;
}

// caller-owned contiguous output, or a temporary when the caller does not want it
int out_or_tmp(Pool& pool, double* p, i64 r, i64 c, um_view* v) {
  if (p) return um_view_init(v, p, r, c, UM_F64);
  return pool.mat(r, c, v);
}

// out (m x n) = A' B for tall operands (A: K x m, B: K x n, K >> m, n): `*@` splits such a contraction along K
// (matmul.cu launch_dgemm) and adds the partial products in a fixed tree
int tall_gram(const um_view& A, const um_view& B, const um_view& out) {
  const um_view At = tr(A);
  return um_matmul(&At, &B, &out);
}

}  // namespace

int tprf_fit_wide(int mode, const um_view* y, const um_view* Z, i64 N, const TprfWideAcc& acc, double n_total, bool sharded,
                  um_tprf_out* out) {
  const i64 T = Z->rows, L = Z->cols, D = L + 1;
  Pool pool;
  const bool closed_wanted = mode == UM_TPRF_CLOSED || (mode == UM_TPRF_IS_FULL && out->alpha);
  const bool three_wanted = mode != UM_TPRF_CLOSED || out->phi || out->phi_hat || out->sigma || out->beta3;
  // from the sweeps: W0 = diag(1/sd) (X - mu)' Zc (N x L), PX = Xn W0 (T x L), h0 = Xn 1 (T x 1), 1/sd and mu/sd (1 x N)
  um_view W0 = acc.W0, PX = acc.PX, h0 = acc.h0, isd = acc.isd, m = acc.m;
  // column-sharded fit: every sum over predictors is completed across the ranks (rank-order sums: all ranks hold
  // the same bits); W0, 1/sd, mu/sd, alpha, Phi stay this rank's columns
  auto allsum = [&](const um_view& v) -> int {
    if (!sharded) return UM_OK;
    return comm_allreduce_sum_f64_impl(static_cast<double*>(v.data) + v.offset, v.rows * v.cols);   // contiguous buffers only
  };
  W_(allsum(PX));
  W_(allsum(h0));
  if (out->xstd) {   // column sample std (n <= 1 or sd == 0 -> 1, Tprf3.scala:469-502)
    um_view xs;
    W_(um_view_init(&xs, out->xstd, 1, N, UM_F64));
    W_(um_binary_scalar(UM_DIV, &isd, 1.0, 1, &xs));
  }

  // ---- proxies centred; y centred ----
  um_view zbar, Zc, ybar, yc, NL;
  W_(pool.mat(1, L, &zbar));
  W_(pool.mat(T, L, &Zc));
  W_(pool.mat(1, 1, &ybar));
  W_(pool.mat(T, 1, &yc));
  W_(pool.mat(N, L, &NL));
  W_(um_reduce_axis(UM_MEAN, Z, 0, &zbar));
  W_(um_binary(UM_SUB, Z, &zbar, &Zc));
  W_(um_reduce_axis(UM_MEAN, y, 0, &ybar));
  W_(um_binary(UM_SUB, y, &ybar, &yc));

  um_view fc, rs;
  W_(out_or_tmp(pool, out->forecasts, T, 1, &fc));
  W_(out_or_tmp(pool, out->residuals, T, 1, &rs));

  // ---- closed form: G = J(T) Xn Wxz = PX - 1 smw' - (h0 - smu) wbar',  s = (G'G)^-1 G'yc ----
  if (closed_wanted) {
    um_view sw, smw, smu, G, r, RW, core, core_i, rhs, s;
    W_(pool.mat(1, L, &sw));
    W_(pool.mat(1, L, &smw));
    W_(pool.mat(1, 1, &smu));
    W_(pool.mat(T, L, &G));
    W_(pool.mat(T, 1, &r));
    W_(pool.mat(T, L, &RW));
    W_(pool.mat(L, L, &core));
    W_(pool.mat(L, L, &core_i));
    W_(pool.mat(L, 1, &rhs));
    W_(pool.mat(L, 1, &s));
    W_(um_reduce_axis(UM_SUM, &W0, 0, &sw));
    W_(allsum(sw));
    W_(um_binary_scalar(UM_DIV, &sw, n_total, 0, &sw));                         // wbar
    W_(tall_gram(tr(m), W0, smw));                                        // smw = (mu/sd) W0
    W_(allsum(smw));
    W_(um_reduce_axis(UM_SUM, &m, 1, &smu));
    W_(allsum(smu));
    W_(um_binary(UM_SUB, &PX, &smw, &G));
    W_(um_binary(UM_SUB, &h0, &smu, &r));
    W_(um_matmul(&r, &sw, &RW));
    W_(um_binary(UM_SUB, &G, &RW, &G));
    W_(tall_gram(G, G, core));
    W_(tall_gram(G, yc, rhs));
    W_(um_inverse(&core, &core_i));
    W_(um_matmul(&core_i, &rhs, &s));
    if (mode == UM_TPRF_CLOSED) {
      W_(um_matmul(&G, &s, &fc));
      W_(um_binary(UM_ADD, &fc, &ybar, &fc));
    }
    if (out->alpha) {                                                           // alpha = Wxz s, Wxz = W0 - wbar
      um_view al;
      W_(um_view_init(&al, out->alpha, N, 1, UM_F64));
      W_(um_binary(UM_SUB, &W0, &sw, &NL));
      W_(um_matmul(&NL, &s, &al));
    }
  }

  // ---- the three passes from the same accumulators ----
  if (three_wanted) {
    um_view Szz, Szz_i, Phi, dP, PtP, PtP_i, XdP, Xaug, XtX, XtX_i, Xty, beta;
    W_(pool.mat(L, L, &Szz));
    W_(pool.mat(L, L, &Szz_i));
    W_(out_or_tmp(pool, out->phi, N, L, &Phi));
    W_(pool.mat(N, D, &dP));
    W_(pool.mat(D, D, &PtP));
    W_(pool.mat(D, D, &PtP_i));
    W_(pool.mat(T, D, &XdP));
    W_(pool.mat(T, D, &Xaug));
    W_(pool.mat(D, D, &XtX));
    W_(pool.mat(D, D, &XtX_i));
    W_(pool.mat(D, 1, &Xty));
    W_(out_or_tmp(pool, out->beta3, D, 1, &beta));
    // pass 1: Phi = W0 (Zc'Zc)^-1 (the slopes of x_j/sd_j on [1 | Z], Tprf3.scala:263-264)
    W_(tall_gram(Zc, Zc, Szz));
    W_(um_inverse(&Szz, &Szz_i));
    W_(um_matmul(&W0, &Szz_i, &Phi));
    const um_view Phit = tr(Phi);
    if (out->phi_hat) {                                                         // B1 = [mu/sd - zbar Phi' ; Phi']
      um_view B1;
      W_(um_view_init(&B1, out->phi_hat, D, N, UM_F64));
      const um_view b0 = sl(B1, 0, 1, 0, N), rest = sl(B1, 1, D, 0, N);
      W_(um_matmul(&zbar, &Phit, &b0));
      W_(um_binary(UM_SUB, &m, &b0, &b0));
      W_(um_copy(&Phit, &rest));
    }
    // pass 2: Sigma = (Xn [1 | Phi]) ([1 | Phi]'[1 | Phi])^-T, columns 1..L (Tprf3.scala:266-267)
    const um_view dP0 = sl(dP, 0, N, 0, 1), dP1 = sl(dP, 0, N, 1, D);
    W_(um_fill(&dP0, 1.0));
    W_(um_copy(&Phi, &dP1));
    W_(tall_gram(dP, dP, PtP));
    W_(allsum(PtP));
    W_(um_inverse(&PtP, &PtP_i));
    const um_view XdP0 = sl(XdP, 0, T, 0, 1), XdP1 = sl(XdP, 0, T, 1, D);
    W_(um_copy(&h0, &XdP0));
    W_(um_matmul(&PX, &Szz_i, &XdP1));
    const um_view PtP_it = tr(PtP_i);
    W_(um_matmul(&XdP, &PtP_it, &Xaug));
    const um_view Xa0 = sl(Xaug, 0, T, 0, 1), Sig = sl(Xaug, 0, T, 1, D);
    if (out->sigma) {
      um_view so;
      W_(um_view_init(&so, out->sigma, T, L, UM_F64));
      W_(um_copy(&Sig, &so));
    }
    // pass 3: y on [1 | Sigma] (Tprf3.scala:268-269)
    W_(um_fill(&Xa0, 1.0));
    W_(tall_gram(Xaug, Xaug, XtX));
    W_(tall_gram(Xaug, *y, Xty));
    W_(um_inverse(&XtX, &XtX_i));
    W_(um_matmul(&XtX_i, &Xty, &beta));
    if (mode != UM_TPRF_CLOSED) W_(um_matmul(&Xaug, &beta, &fc));
  }

  // ---- residuals and R^2 (closed / t3prf: ssy == 0 -> 0, Tprf3.scala:239-241,274-275; IS Full: no guard, :1214-1225) ----
  W_(um_binary(UM_SUB, y, &fc, &rs));
  um_view sq;
  W_(pool.mat(T, 1, &sq));
  double see = 0.0, ssy = 0.0;
  W_(um_binary(UM_MUL, &rs, &rs, &sq));
  W_(um_reduce_all(UM_SUM, &sq, &see));
  W_(um_binary(UM_MUL, &yc, &yc, &sq));
  W_(um_reduce_all(UM_SUM, &sq, &ssy));
  const bool guarded = mode == UM_TPRF_CLOSED || mode == UM_TPRF_ITER;
  out->r_squared = (guarded && ssy == 0.0) ? 0.0 : 1.0 - see / ssy;
  W_(um_sync());
  return UM_OK;
}

}  // namespace um
