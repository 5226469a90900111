"""CPU oracle for the uni 3PRF path -- TEST INFRASTRUCTURE ONLY.

NumPy restatement of `uni.stats.Tprf3` (Scala) for the fits SURVEY.md section 8a rows
a15-a17 name.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this module; the product path never does.

Parity is PINNED three ways (tests/test_oracle_tprf.py):
  * Kelly-Pruitt MATLAB known answers `t3prf-validation/kp_{yhat,alpha,rsquare}.csv`;
  * every `*/isfull` and `*/closed` row of `test-data/tprf3-parity/scala-reference.txt`
    (the reference's own cross-language fixture, tolerance rule
    src/test/scala/uni/stats/Tprf3ParitySuite.scala:28-45);
  * `tests/golden/tprf_pyref/*.npz`: outputs of the reference's own NumPy implementation
    (`/root/reference/py/tprf3fast.py`) imported and run in the build container by
    `tests/golden/gen_tprf_pyref.py`.

File:line citations are into /root/reference/.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import numpy as np


# ---------------------------------------------------------------------------------------
# helpers
# ---------------------------------------------------------------------------------------
def nan_std_cols(X: np.ndarray) -> np.ndarray:
    """Sample std per column over non-NaN rows, two-pass; n<=1 or sd==0 -> 1.0.
    src/main/scala/uni/stats/Tprf3.scala:469-502 (and stdCols :550-578)."""
    X = np.asarray(X, dtype=np.float64)
    ok = ~np.isnan(X)
    n = ok.sum(axis=0)
    s = np.where(ok, X, 0.0).sum(axis=0)
    mu = np.where(n > 1, s / np.maximum(n, 1), 0.0)
    d = np.where(ok, X - mu, 0.0)
    ss = (d * d).sum(axis=0)
    with np.errstate(all="ignore"):
        sd = np.sqrt(ss / (n - 1))
    sd = np.where((n > 1) & (sd != 0.0), sd, 1.0)
    return sd.reshape(1, -1)


def center_columns(M: np.ndarray) -> np.ndarray:
    """J(rows) @ M as `m - m.sum(0)/rows` (Tprf3.scala:458-459)."""
    return M - M.sum(axis=0, keepdims=True) / float(M.shape[0])


def with_intercept(M: np.ndarray) -> np.ndarray:
    """[1 | M] (Tprf3.scala:835-846)."""
    return np.column_stack([np.ones(M.shape[0]), M])


def lu_inverse(A: np.ndarray) -> np.ndarray:
    """`Mat.inverse`: LU with partial pivoting + n substitutions
    (src/main/scala/uni/data/Mat.scala:970-1001,2979-3006).  Singular -> ArithmeticError."""
    from .mat_oracle import from2d, inverse  # local import: same test-only package

    return inverse(from2d(np.asarray(A, dtype=np.float64))).to2d()


@dataclass
class TprfResult:
    forecasts: np.ndarray  # T x 1
    residuals: np.ndarray  # T x 1
    r_squared: float
    phi: Optional[np.ndarray] = None  # N x L
    sigma: Optional[np.ndarray] = None  # T x L
    beta3: Optional[np.ndarray] = None  # (L+1) x 1
    phi_hat: Optional[np.ndarray] = None  # (L+1) x N
    alpha: Optional[np.ndarray] = None  # N x 1
    extra: dict = field(default_factory=dict)


def _r2_guarded(y, resid):
    """ssy == 0 -> 0 else 1 - sum(e^2)/ssy (Tprf3.scala:239-241, 274-275)."""
    ssy = float(((y - y.mean()) ** 2).sum())
    return 0.0 if ssy == 0.0 else 1.0 - float((resid**2).sum()) / ssy


# ---------------------------------------------------------------------------------------
# literal restatements (small N only: they form the N x N Sxx)
# ---------------------------------------------------------------------------------------
def tprf_closed_form_literal(y, X, Z) -> TprfResult:
    """`Tprf3.tprfClosedForm` exactly as written (Tprf3.scala:199-252), including the
    N x N `Sxx` and the per-column / per-row pass-1/2 `Lm` fits -- usable for small N only."""
    y = np.asarray(y, dtype=np.float64).reshape(-1, 1)
    X = np.asarray(X, dtype=np.float64)
    Z = np.asarray(Z, dtype=np.float64)
    Xn = X / nan_std_cols(X)
    T, N = Xn.shape
    L = Z.shape[1]
    JtX = center_columns(Xn)
    XtJt = JtX.T
    Wxz = center_columns(XtJt @ Z)
    Sxx = XtJt @ Xn
    core = Wxz.T @ Sxx @ Wxz
    factor = Wxz @ lu_inverse(core) @ Wxz.T @ XtJt
    alpha = factor @ y
    yhat = JtX @ alpha + y.mean()
    # pass 1 / pass 2 / pass 3 fields (:211-231): plain OLS with intercept
    dZ = with_intercept(Z)
    B1 = lu_inverse(dZ.T @ dZ) @ (dZ.T @ Xn)
    phi = B1[1:, :].T
    dP = with_intercept(phi)
    B2 = lu_inverse(dP.T @ dP) @ (dP.T @ Xn.T)
    sigma = B2[1:, :].T
    Xaug = with_intercept(sigma)
    beta3 = lu_inverse(Xaug.T @ Xaug) @ (Xaug.T @ y)
    resid = y - yhat
    return TprfResult(yhat, resid, _r2_guarded(y, resid), phi=phi, sigma=sigma, beta3=beta3,
                      phi_hat=B1, alpha=alpha)


def t3prf(y, X, Z) -> TprfResult:
    """`Tprf3.t3prf` (Tprf3.scala:260-289): three batch OLS passes on Xn = X / sd."""
    y = np.asarray(y, dtype=np.float64).reshape(-1, 1)
    X = np.asarray(X, dtype=np.float64)
    Z = np.asarray(Z, dtype=np.float64)
    Xn = X / nan_std_cols(X)
    dZ = with_intercept(Z)
    B1 = lu_inverse(dZ.T @ dZ) @ (dZ.T @ Xn)
    phi = B1[1:, :].T
    dP = with_intercept(phi)
    B2 = lu_inverse(dP.T @ dP) @ (dP.T @ Xn.T)
    sigma = B2[1:, :].T
    Xaug = with_intercept(sigma)
    beta = lu_inverse(Xaug.T @ Xaug) @ (Xaug.T @ y)
    yhat = Xaug @ beta
    resid = y - yhat
    return TprfResult(yhat, resid, _r2_guarded(y, resid), phi=phi, sigma=sigma, beta3=beta,
                      phi_hat=B1)


def estimate3prf_is_full(y, X, Z, min_obs: int = 10) -> TprfResult:
    """`Tprf3.estimate3prf(y, X, Right(Z), "IS Full")` for NaN-free input
    (Tprf3.scala:1056-1110 -> runT3prf :973 -> t3prfFast :883-901 -> t3prfTail :913-943;
    R^2 :1214-1225 -- no zero guard; alpha :1240-1250 -- the N x N form)."""
    y = np.asarray(y, dtype=np.float64).reshape(-1, 1)
    X = np.asarray(X, dtype=np.float64)
    Z = np.asarray(Z, dtype=np.float64)
    T, N = X.shape
    L = Z.shape[1]
    Xn = X / nan_std_cols(X)
    if T < min_obs:
        f = np.full((T, 1), np.nan)
        return TprfResult(f, y - f, float("nan"))
    dZ = with_intercept(Z)
    B1 = lu_inverse(dZ.T @ dZ) @ (dZ.T @ Xn)
    dP = with_intercept(B1[1:, :].T)
    PtPinv = lu_inverse(dP.T @ dP)
    sigma = ((Xn @ dP) @ PtPinv.T)[:, 1 : L + 1]
    Xaug = with_intercept(sigma)
    beta = lu_inverse(Xaug.T @ Xaug) @ (Xaug.T @ y)
    f = Xaug @ beta
    resid = y - f
    with np.errstate(all="ignore"):
        r2 = 1.0 - float((resid**2).sum()) / float(((y - y.mean()) ** 2).sum())
    XtJt = center_columns(Xn).T
    Wxz = center_columns(XtJt @ Z)
    Sxx = XtJt @ Xn
    alpha = Wxz @ lu_inverse(Wxz.T @ Sxx @ Wxz) @ Wxz.T @ (XtJt @ y)
    return TprfResult(f, resid, r2, phi=B1[1:, :].T, sigma=sigma, beta3=beta, phi_hat=B1, alpha=alpha)


# ---------------------------------------------------------------------------------------
# estimate3prf with the out-of-sample procedures
# ---------------------------------------------------------------------------------------
def _window_fit(yt, Xt, inv_sd, Zt, xt=None, min_obs=10):
    """t3prfFast + t3prfTail on one window (Tprf3.scala:883-943): `Xt` unscaled, `inv_sd` = 1/column std of
    the window; returns (yhat or None, yhatt).  `T < minObs` -> all-NaN fit, NaN forecast (:893-894)."""
    T = yt.shape[0]
    if T < min_obs:
        return np.full((T, 1), np.nan), float("nan")
    L = Zt.shape[1]
    dZ = with_intercept(Zt)
    ZtX = (dZ.T @ Xt) * inv_sd.reshape(1, -1)
    B1 = lu_inverse(dZ.T @ dZ) @ ZtX
    dP = with_intercept(B1[1:, :].T)
    PtPinv = lu_inverse(dP.T @ dP)
    dps = dP * inv_sd.reshape(-1, 1)
    sigma = ((Xt @ dps) @ PtPinv.T)[:, 1 : L + 1]
    Xaug = with_intercept(sigma)
    # pass 3 is nanOls(y, Xaug, minObs = 1) (Tprf3.scala:929-931): rows with a NaN in y or Sigma are left out; no
    # usable row -> NaN coefficients.  On NaN-free data it is the plain normal-equations solve.
    beta = nan_ols(yt, Xaug, 1)
    if beta is None:
        beta = np.full((L + 1, 1), np.nan)
    yhatt = float("nan")
    if xt is not None:
        bo = (xt.reshape(1, -1) @ dps) @ PtPinv.T
        yhatt = float(beta[0, 0] + bo[0, 1:] @ beta[1:, 0])
    return Xaug @ beta, yhatt


def nan_ols(y, X, min_obs):
    """nanOls (Tprf3.scala:807-830): rows with a NaN in y or X dropped; fewer than min_obs valid rows -> None;
    (X'X)^-1 X'y with the LU inverse."""
    y = np.asarray(y, dtype=np.float64).reshape(-1, 1)
    valid = ~np.isnan(y[:, 0]) & ~np.isnan(X).any(axis=1)
    if valid.sum() < min_obs:
        return None
    Xv, yv = X[valid], y[valid]
    return lu_inverse(Xv.T @ Xv) @ (Xv.T @ yv)


def nan_mean_cols(M):
    """nanMeanCols (Tprf3.scala:506-521): mean over the non-NaN rows, NaN for an empty column."""
    ok = ~np.isnan(M)
    n = ok.sum(axis=0)
    with np.errstate(all="ignore"):
        return (np.where(ok, M, 0.0).sum(axis=0) / np.where(n > 0, n, np.nan)).reshape(1, -1)


def t3prf_pls(yt, Xs, Zt, xt=None, min_obs=10):
    """runT3prf's pls branch (Tprf3.scala:985-1048): `Xs` already scaled by the window's 1/std; passes 1 and 2 are
    per-column / per-row nanOls fits WITHOUT intercept on the column-centred data, so gaps (NaN) only remove the
    observation they sit in.  Returns (yhat T x 1, yhatt)."""
    T, N = Xs.shape
    L = Zt.shape[1]
    Phi = np.full((N, L), np.nan)
    Sigma = np.full((T, L), np.nan)
    cm = nan_mean_cols(Xs)
    Xc = Xs - cm
    for i in range(N):                                   # pass 1 (:1001-1005)
        b = nan_ols(Xc[:, i], Zt, min_obs)
        if b is not None:
            Phi[i, :] = b[:, 0]
    for t in range(T):                                   # pass 2, identification floor L + 1 (:1013-1019)
        b = nan_ols(Xc[t, :], Phi, L + 1)
        if b is not None:
            Sigma[t, :] = b[:, 0]
    Xaug = with_intercept(Sigma)
    beta = nan_ols(yt, Xaug, 1)                          # pass 3 (:1021-1023)
    if beta is None:
        beta = np.full((L + 1, 1), np.nan)
    yhatt = float("nan")
    if xt is not None:                                   # :1025-1036
        b = nan_ols((np.asarray(xt).reshape(1, -1) - cm).ravel(), Phi, L + 1)
        if b is not None:
            yhatt = float(beta[0, 0] + b[:, 0] @ beta[1:, 0])
    return Xaug @ beta, yhatt


def _inv_std(M):
    """invStdCols (Tprf3.scala:618-642 family): reciprocal sample std, sd == 0 -> 1."""
    return 1.0 / nan_std_cols(M).reshape(-1)


def encnew(e1, e2) -> float:
    """Tprf3.scala:854-860, called as encnew(rollfore[loc], residuals[loc]) (:1235-1238)."""
    ok = ~np.isnan(e1) & ~np.isnan(e2)
    a, b = e1[ok], e2[ok]
    with np.errstate(all="ignore"):
        return float(a.size * (a * a - a * b).sum() / (b * b).sum())


def estimate3prf(y, X, Z, procedure="IS Full", window=(0, 1), mintrain=(-1, 0), rollwin=(30, 20, 0), pls=False,
                 compute_avar=False) -> TprfResult:
    """`Tprf3.estimate3prf` (Tprf3.scala:1056-1250), every window fitted directly (the reference's FullSample
    downdating is an O(eps) re-association of the same sums, :742-757).  `Z` = proxy matrix (Right) or an int L
    (Left: automatic proxies).  `pls = True` (effective with automatic proxies only, :1076) takes the per-column /
    per-row nanOls passes and tolerates NaN gaps in X and y; without it the input must be NaN-free.  Result:
    forecasts, residuals, r_squared, extra = {rollfore, enc}; with `compute_avar` ("IS Full" only, :1252-1277) also
    extra["avar_alpha"] (N x N) and extra["avar_forecasts"] (T x T)."""
    y = np.asarray(y, dtype=np.float64).reshape(-1, 1)
    X = np.asarray(X, dtype=np.float64)
    T, N = X.shape
    auto = isinstance(Z, (int, np.integer))
    L = int(Z) if auto else np.asarray(Z).shape[1]
    Zm = None if auto else np.asarray(Z, dtype=np.float64)
    z_final = None
    mt = (T // 2, mintrain[1]) if mintrain[0] < 0 else tuple(mintrain)
    win, min_nona, gap = rollwin
    Xn = X / nan_std_cols(X)
    forecasts = np.full((T, 1), np.nan)
    rollfore = np.full((T, 1), np.nan)

    eff_pls = bool(pls) and auto
    has_nan = bool(np.isnan(X).any() or np.isnan(y).any() or (Zm is not None and np.isnan(Zm).any()))   # Tprf3.scala:1078

    def nanmean(v):
        ok = ~np.isnan(v)
        return float(v[ok].sum() / ok.sum()) if ok.any() else float("nan")

    def window_forecast(yt, Xt, Zt, xt, min_obs=10):
        inv = _inv_std(Xt)
        if eff_pls:                                       # runT3prf(..., effPls): scaled window, nanOls passes
            Xs, xs = Xt * inv.reshape(1, -1), xt * inv
            r0, tp = yt * 1.0, float("nan")
            for _ in range(L):
                yh, tp = t3prf_pls(yt, Xs, r0, xs, min_obs)
                r0 = np.column_stack([yt - yh, r0])
            return tp
        if not auto:
            return _window_fit(yt, Xt, inv, Zt, xt, 10)[1]
        r0, tp = yt * 1.0, float("nan")
        for _ in range(L):  # :1134-1139: each new residual column goes in FRONT
            yh, tp = _window_fit(yt, Xt, inv, r0, xt, min_obs)
            r0 = np.column_stack([yt - yh, r0])
        return tp

    if procedure == "IS Full":
        ones = np.ones(N)
        if auto:
            r0, fore = y * 1.0, np.full((T, 1), np.nan)
            for _ in range(L):  # :1097-1106: residual columns are appended at the BACK
                fore = t3prf_pls(y, Xn, r0, None, 10)[0] if eff_pls else _window_fit(y, Xn, ones, r0, None, 10)[0]
                z_final = r0            # the proxies of the LAST fit (:1102), before its residual is appended
                r0 = np.column_stack([r0, y - fore])
            forecasts[:] = fore
        else:
            forecasts[:] = _window_fit(y, Xn, ones, Zm, None, 10)[0]
            z_final = Zm
    elif procedure == "OOS Cross Val":
        for t in range(T):
            lo, hi = max(t - window[0], 0), min(t - window[0] + window[1], T)
            keep = np.r_[0:lo, hi:T]
            yt = y[keep]
            forecasts[t, 0] = window_forecast(yt, Xn[keep], None if auto else Zm[keep], Xn[t])
            rollfore[t, 0] = nanmean(yt[:, 0]) if (eff_pls or has_nan) else yt.mean()   # colMean(yt, hasNan)
    elif procedure == "OOS Recursive":
        for t in range(mt[0] + 1 + mt[1], T):
            end = t - 1 - mt[1]
            yt = y[:end]
            forecasts[t, 0] = window_forecast(yt, Xn[:end], None if auto else Zm[:end], Xn[t], mt[0])
            rollfore[t, 0] = nanmean(yt[:, 0]) if (eff_pls or has_nan) else yt.mean()   # colMean(yt, hasNan)
    elif procedure == "OOS Rolling":
        for t in range(win + 1 + gap, T):
            lo, hi = max(t - win - gap, 0), min(t - 1 - gap, T)
            yt = y[lo:hi]
            forecasts[t, 0] = window_forecast(yt, Xn[lo:hi], None if auto else Zm[lo:hi], Xn[t], min_nona)
            rollfore[t, 0] = nanmean(yt[:, 0]) if (eff_pls or has_nan) else yt.mean()   # colMean(yt, hasNan)
    else:
        raise ValueError(f"Unknown procedure '{procedure}'. Choose: 'IS Full', 'OOS Recursive', 'OOS Cross Val', 'OOS Rolling'")

    resid = y - forecasts
    loc = ~np.isnan(resid[:, 0])
    with np.errstate(all="ignore"):
        if procedure == "IS Full":  # :1214-1225
            mu = np.float64(y[loc, 0].sum()) / np.float64(loc.sum())          # no usable row: 0/0 = NaN, as the JVM's
            r2 = float(1.0 - np.float64((resid[loc, 0] ** 2).sum()) / np.float64(((y[loc, 0] - mu) ** 2).sum()))
        else:  # :1226-1233: rolling-mean denominator, ssT == 0 -> NaN
            sst = float(((y[loc, 0] - rollfore[loc, 0]) ** 2).sum())
            r2 = 1.0 - float((resid[loc, 0] ** 2).sum()) / sst if sst != 0.0 else float("nan")
    enc = encnew(rollfore[loc, 0], resid[loc, 0]) if procedure == "OOS Recursive" else float("nan")
    extra = {"rollfore": rollfore, "enc": enc}
    if compute_avar and procedure == "IS Full" and z_final is not None:
        extra["avar_alpha"], extra["avar_forecasts"] = avar_estimates(Xn, z_final, resid)
    return TprfResult(forecasts, resid, r2, extra=extra)


def avar_estimates(Xn, zf, resid):
    """The asymptotic-variance block of estimate3prf("IS Full", computeAvar = true) (Tprf3.scala:1252-1277): Xn the
    standardised panel (T x N), zf the proxies of the final fit (T x L), resid the in-sample residuals (T x 1).
    Every J(T) / J(N) factor of the formulas is a column / row centring (:1255-1262); the T-term sum of weighted outer
    products (:1266-1270) is accumulated row by row as the reference does.  Returns (alpha avar N x N, forecast avar T x T)."""
    Xn = np.asarray(Xn, dtype=np.float64); zf = np.asarray(zf, dtype=np.float64)
    T, N = Xn.shape
    centre_rows = lambda M: M - M.mean(axis=1, keepdims=True)
    Xc, Zc = center_columns(Xn), center_columns(zf)
    xz = Xc.T @ zf                                  # Xn' J(T) zf          N x L
    zx = centre_rows(Zc.T @ Xn)                     # zf' J(T) Xn J(N)     L x N
    xx = centre_rows(Xc.T @ Xn)                     # Xn' J(T) Xn J(N)     N x N
    a = xz * (1.0 / T)
    b = (zx @ xx @ xz) * (float(T) ** -3 * float(N) ** -2)
    c = zx * (1.0 / T / N)
    omega_a = center_columns(a) @ lu_inverse(b) @ c
    xm = Xn.sum(axis=0) / float(T)
    acc = np.zeros((N, N))
    for t in range(T):
        d = (Xn[t] - xm).reshape(1, -1)
        acc = acc + (d.T @ d) * (float(resid[t, 0]) ** 2 / T)
    alpha_avar = omega_a @ acc @ omega_a.T
    return alpha_avar, (Xc @ alpha_avar @ Xc.T) * float(N) ** -2


# ---------------------------------------------------------------------------------------
# PLS-variant closed form + predict (Tprf3.scala:299-401)
# ---------------------------------------------------------------------------------------
@dataclass
class PlsModel:
    phi: np.ndarray        # N x 1
    sigma: np.ndarray      # T x 1
    beta: np.ndarray       # 2 x 1
    forecasts: np.ndarray  # T x 1
    col_mean: np.ndarray   # 1 x N  means of the scaled X
    col_std: np.ndarray    # 1 x N  std of the raw X
    r_squared: float

    def predict(self, row) -> float:
        """Pls3prfModel.predict (Tprf3.scala:325-333): passes 2 and 3 for one raw row."""
        row = np.asarray(row, dtype=np.float64).reshape(-1)
        if row.size != self.phi.shape[0]:
            raise ValueError(f"row length {row.size} != {self.phi.shape[0]} predictors")
        phi = self.phi[:, 0]
        dot = float(((row / self.col_std[0] - self.col_mean[0]) * phi).sum())
        return float(self.beta[0, 0] + self.beta[1, 0] * (dot / float((phi * phi).sum())))


def pls_closed_form(y, X) -> PlsModel:
    """`Tprf3.plsClosedForm` (Tprf3.scala:360-390): no intercept in passes 1 and 2, proxy = y."""
    y = np.asarray(y, dtype=np.float64).reshape(-1, 1)
    X = np.asarray(X, dtype=np.float64)
    if X.shape[0] != y.shape[0]:
        raise ValueError(f"X has {X.shape[0]} rows but y has {y.shape[0]}")
    if np.isnan(X).any() or np.isnan(y).any():
        raise ValueError("plsClosedForm requires NaN-free input")
    col_std = nan_std_cols(X)
    Xn = X / col_std
    col_mean = Xn.sum(axis=0, keepdims=True) / X.shape[0]
    Xc = Xn - col_mean
    yss = float((y.T @ y)[0, 0])
    with np.errstate(all="ignore"):
        phi = (Xc.T @ y) / yss
        sigma = (Xc @ phi) / float((phi.T @ phi)[0, 0])
    Xaug = with_intercept(sigma)
    beta = lu_inverse(Xaug.T @ Xaug) @ (Xaug.T @ y)
    f = Xaug @ beta
    return PlsModel(phi, sigma, beta, f, col_mean, col_std, _r2_guarded(y, y - f))


# ---------------------------------------------------------------------------------------
# scalable (Gram-reassociated) closed form -- the reference's own Rust port does this
# ---------------------------------------------------------------------------------------
def tprf_closed_form_gram(y, X, Z) -> TprfResult:
    """Closed form with `Wxz' Sxx Wxz = G'G`, `G = J(T) Xn Wxz` -- never forms N x N.
    rust/src/t3prf.rs:847-879 (`tprfClosedForm`); same result as Tprf3.scala:199-209."""
    y = np.asarray(y, dtype=np.float64).reshape(-1, 1)
    X = np.asarray(X, dtype=np.float64)
    Z = np.asarray(Z, dtype=np.float64)
    Xn = X / nan_std_cols(X)
    JtX = center_columns(Xn)
    Wxz = center_columns(JtX.T @ Z)
    G = JtX @ Wxz
    core = G.T @ G
    rhs = Wxz.T @ (JtX.T @ y)
    s = np.linalg.solve(core, rhs)
    alpha = Wxz @ s
    ybar = float(y.mean())
    yhat = JtX @ alpha + ybar
    resid = y - yhat
    return TprfResult(yhat, resid, _r2_guarded(y, resid), alpha=alpha)


def closed_form_partials(X_shard, Z, T_total=None):
    """The shard-local accumulators of the fused device sweep, in
    NumPy: used by the world_size-2 gloo test and by `bench.py --impl reference` to time a
    column-sharded CPU fit.  Returns a dict of partial sums that add across shards."""
    X = np.asarray(X_shard, dtype=np.float64)
    Z = np.asarray(Z, dtype=np.float64)
    T = X.shape[0]
    sd = nan_std_cols(X).reshape(-1)
    mu = X.mean(axis=0)
    Zc = Z - Z.mean(axis=0, keepdims=True)
    W0 = ((X - mu) / sd).T @ Zc  # N_s x L  (= Xn' J(T) Z)
    Xn = X / sd
    return {
        "PX": Xn @ W0,  # T x L
        "h0": Xn.sum(axis=1),  # T
        "sw": W0.sum(axis=0),  # L
        "SWW": W0.T @ W0,  # L x L
        "smw": (mu / sd) @ W0,  # L
        "smu": float((mu / sd).sum()),
        "n": X.shape[1],
        "W0": W0,
    }


def closed_form_from_partials(y, parts_sum, n_total):
    """Finish step shared by every rank after the all-reduce."""
    y = np.asarray(y, dtype=np.float64).reshape(-1, 1)
    wbar = parts_sum["sw"] / n_total
    XcW0 = parts_sum["PX"] - parts_sum["smw"][None, :]
    r = parts_sum["h0"] - parts_sum["smu"]
    G = XcW0 - np.outer(r, wbar)
    core = G.T @ G
    rhs = G.T @ (y - y.mean())
    s = np.linalg.solve(core, rhs)
    yhat = G @ s + float(y.mean())
    return yhat, s, wbar
