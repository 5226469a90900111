"""Host-side mirror of `uni.stats.Tprf3` (src/main/scala/uni/stats/Tprf3.scala) over libunimat.

`tprfClosedForm`, `t3prf`, `estimate3prf` (IS Full and the three out-of-sample procedures, proxy
matrix or automatic proxies) and `forecast3prf` keep the reference's names, argument meaning and
error behaviour; every fit is one fused device sweep plus an O(T L^2) finish (uni_b200/csrc/tprf.cu),
an out-of-sample window adds one small launch for its held-out row (uni_b200/csrc/tprf_oos.cu).  Under an active communicator
(`uni_b200.dist`) X is this rank's column shard and `n_total` the global predictor count.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Tuple, Union

import numpy as np

from . import _lib as L
from ._lib import UmTprfOut, check, lib
from .mat import Mat, MatD


@dataclass
class AvarEstimates:
    """AvarEstimates (Tprf3.scala:94-99): asymptotic variance of alpha (N x N) and of the forecasts (T x T)."""
    alpha: Mat
    forecasts: Mat


@dataclass
class Tprf3Result:
    """Tprf3Result (Tprf3.scala:72-92): device-resident fields."""
    forecasts: Mat
    residuals: Mat
    rSquared: float
    phi: Optional[Mat] = None       # N x L
    sigma: Optional[Mat] = None     # T x L
    beta3: Optional[Mat] = None     # (L+1) x 1
    phiHat: Optional[Mat] = None    # (L+1) x N
    alpha: Optional[Mat] = None     # N x 1
    xstd: Optional[Mat] = None      # 1 x N  column sample std used to normalise X
    singular: bool = False
    rollfore: Optional[Mat] = None  # T x 1  mean of each window's y (OOS procedures, Tprf3.scala:1145)
    encnew: float = float("nan")    # ENC-NEW statistic, OOS Recursive only (Tprf3.scala:854-860,1235-1238)
    avar: Optional[AvarEstimates] = None   # estimate3prf("IS Full", computeAvar = true) only (Tprf3.scala:1252-1277)

    @property
    def y_hat(self) -> Mat:         # TprfResult alias (Tprf3.scala:93-95)
        return self.forecasts


def _fit(mode: int, y: Mat, X: Mat, Z: Mat, n_total: int = 0, want_all: bool = True) -> Tprf3Result:
    for name, m in (("y", y), ("X", X), ("Z", Z)):
        if not isinstance(m, Mat) or m.dtype != L.F64:
            raise ValueError(f"{name} must be a MatD")
    T, N, Lp = X.rows, X.cols, Z.cols
    if y.rows != T or y.cols != 1:
        raise ValueError(f"y must be {T}x1, got {y.shape}")
    if Z.rows != T:
        raise ValueError(f"Z must have {T} rows, got {Z.rows}")
    if X.cs != 1 and N != 1:
        X = X.copy()  # the sweep wants row-contiguous X; a transposed view is materialised once
    f = Mat._alloc(T, 1, L.F64)
    r = Mat._alloc(T, 1, L.F64)
    out = UmTprfOut()
    out.forecasts, out.residuals = f._buf.ptr, r._buf.ptr
    res = Tprf3Result(f, r, float("nan"))
    if want_all:
        res.xstd = Mat._alloc(1, N, L.F64)
        out.xstd = res.xstd._buf.ptr
        if mode in (L.TPRF_CLOSED, L.TPRF_IS_FULL):
            res.alpha = Mat._alloc(N, 1, L.F64)
            out.alpha = res.alpha._buf.ptr
        if True:   # phi / phiHat / sigma / beta3 are part of every result, tprfClosedForm included (Tprf3.scala:211-231,243-252)
            res.phi = Mat._alloc(N, Lp, L.F64)
            res.phiHat = Mat._alloc(Lp + 1, N, L.F64)
            res.sigma = Mat._alloc(T, Lp, L.F64)
            res.beta3 = Mat._alloc(Lp + 1, 1, L.F64)
            out.phi, out.phi_hat = res.phi._buf.ptr, res.phiHat._buf.ptr
            out.sigma, out.beta3 = res.sigma._buf.ptr, res.beta3._buf.ptr
    # a singular Gram (exact-zero pivot) raises SingularMatrixError, as every `.inverse` of the reference's fits does
    # (ArithmeticException, Mat.scala:990)
    check(lib.um_tprf_fit(mode, y._ref(), X._ref(), Z._ref(), int(n_total), C.byref(out)))
    res.rSquared = out.r_squared
    res.singular = bool(out.singular)
    return res


def tprfClosedForm(y: Mat, X: Mat, Z: Mat, n_total: int = 0) -> Tprf3Result:
    """Tprf3.tprfClosedForm (Tprf3.scala:199-252)."""
    return _fit(L.TPRF_CLOSED, y, X, Z, n_total)


def t3prf(y: Mat, X: Mat, Z: Mat, n_total: int = 0) -> Tprf3Result:
    """Tprf3.t3prf (Tprf3.scala:260-289)."""
    return _fit(L.TPRF_ITER, y, X, Z, n_total)


_PROCEDURES = ("IS Full", "OOS Recursive", "OOS Cross Val", "OOS Rolling")


def _avar(Xn: Mat, zf: Mat, resid: Mat) -> AvarEstimates:
    """The asymptotic-variance block of estimate3prf("IS Full", computeAvar = true) (Tprf3.scala:1252-1277) with the
    device operators the reference spells it in: every J(T) / J(N) factor is a column / row centring (:1255-1262),
    `*@`, one L x L `.inverse`.  The reference's T-term loop of weighted outer products (:1266-1270) is ONE product:
    sum_t (r_t^2 / T) d_t' d_t = (Xc o r)' (Xc o r) / T with d_t the centred row.  N x N and T x T by definition: meant
    for the small panels the reference uses it on."""
    T, N = Xn.rows, Xn.cols
    Xc = Xn - Xn.mean(0)
    Zc = zf - zf.mean(0)
    xz = Xc.T @ zf                                   # Xn' J(T) zf          N x L
    t1 = Zc.T @ Xn
    zx = t1 - t1.mean(1)                             # zf' J(T) Xn J(N)     L x N
    t2 = Xc.T @ Xn
    xx = t2 - t2.mean(1)                             # Xn' J(T) Xn J(N)     N x N
    a = xz * (1.0 / T)
    b = (zx @ xx @ xz) * (float(T) ** -3 * float(N) ** -2)
    c = zx * (1.0 / T / N)
    omega_a = (a - a.mean(0)) @ b.inverse() @ c
    w = Xc * resid                                   # rows scaled by their residual (T x 1 broadcast)
    acc = (w.T @ w) * (1.0 / T)
    alpha_avar = omega_a @ acc @ omega_a.T
    return AvarEstimates(alpha_avar, (Xc @ alpha_avar @ Xc.T) * float(N) ** -2)


def _std_cols(X: Mat) -> Mat:
    """stdCols (Tprf3.scala:550-578) for NaN-free data: sample std per column, 0 -> 1."""
    T = X.rows
    sd = X.std(0) * float(np.sqrt(T / (T - 1.0))) if T > 1 else MatD.ones(1, X.cols)
    return Mat.where(sd.eq(0.0), MatD.ones(1, X.cols), sd)


def _oos_forecast(res: Tprf3Result, xt: Mat) -> float:
    if res.phi.cols > 8:
        # more proxies than the one-launch forecast holds: t3prfTail's `bo = (xt o 1/sd) [1|Phi] ([1|Phi]'[1|Phi])^-T`,
        # `yhatt = b0 + bo[1:] . b[1:]` (Tprf3.scala:933-941) with the device operators
        N, Lz = res.phi.rows, res.phi.cols
        dP = MatD.hstack(MatD.ones(N, 1), res.phi)
        bo = ((xt / res.xstd) @ dP) @ (dP.T @ dP).inverse().T
        b = res.beta3
        return float(b[0, 0]) + float((bo._raw_slice(0, 1, 1, Lz + 1) @ b._raw_slice(1, Lz + 1, 0, 1))[0, 0])
    out = C.c_double(0)
    check(lib.um_tprf_oos_forecast(res.phi._ref(), res.xstd._ref(), res.beta3._ref(), xt._ref(), C.byref(out)))
    return out.value


def _window_forecast_pls(yt: Mat, Xt: Mat, n_auto: int, xt: Mat, min_obs: int) -> float:
    """One out-of-sample window of the pls branch (Tprf3.scala:1122-1139): the window re-standardised by its own
    NaN-aware std, automatic proxies rebuilt from the window's residuals, each new column in front."""
    inv = 1.0 / _nan_std_cols(Xt)
    Xs, xs = Xt * inv, xt * inv
    r0, tp = yt.copy(), float("nan")
    for _ in range(n_auto):
        yh, tp = _t3prf_pls(yt, Xs, r0, xs, min_obs)
        r0 = MatD.hstack(yt - yh, r0)
    return tp


def _window_forecast(yt: Mat, Xt: Mat, Zt: Optional[Mat], n_auto: int, xt: Mat, min_obs: int) -> float:
    """One out-of-sample window: t3prfFast + t3prfTail's yhatt (Tprf3.scala:883-943).  The fit re-standardises the
    window's columns by their own sample std (invStdCols of the window); automatic proxies rebuild Z from the
    window's residuals, each new column in front (:1134-1139)."""
    if Zt is not None:
        if yt.rows < 10:                       # runT3prf / t3prfDowndated minObs = 10 (Tprf3.scala:893-894,951-968)
            return float("nan")
        return _oos_forecast(_fit(L.TPRF_WINDOW, yt, Xt, Zt), xt)
    if yt.rows < min_obs:
        return float("nan")
    r0, tp = yt.copy(), float("nan")
    for _ in range(n_auto):
        res = _fit(L.TPRF_WINDOW, yt, Xt, r0)
        tp = _oos_forecast(res, xt)
        r0 = MatD.hstack(yt - res.forecasts, r0)
    return tp


# ---------------------------------------------------------------------------------------------------------------
# pls = true: per-column / per-row nanOls passes (Tprf3.scala:985-1048), gap tolerant.  The reference fits N + T
# small regressions one after the other; here every pass is ONE masked problem: with W the 0/1 matrix of usable
# observations and D the design (rows with a NaN zeroed), all per-lane Grams are the columns of (D (x) D)' W and all
# right-hand sides the columns of D' (W o X) -- two skinny products over the panel -- followed by one batched
# small inverse (um_inverse_batched) and p^2 lane-wise multiply-adds.
# ---------------------------------------------------------------------------------------------------------------
def _zeros_like(m: Mat) -> Mat:
    return MatD.zeros(m.rows, m.cols)


def _nan_to_zero(m: Mat) -> Mat:
    return Mat.where(m.isnan(), _zeros_like(m), m)


def _col(m: Mat, j: int) -> Mat:
    return m._raw_slice(0, m.rows, j, j + 1)


def _nan_std_cols(X: Mat) -> Mat:
    """nanStdCols (Tprf3.scala:469-502): per column over the non-NaN rows, sample std; n <= 1 or sd == 0 -> 1."""
    ok = ~X.isnan()
    w = ok.astype(L.F64)
    Xz = Mat.where(ok, X, _zeros_like(X))
    n = w.sum(0)
    one = MatD.ones(1, X.cols)
    few = n.lte(1.0)
    mu = Mat.where(few, MatD.zeros(1, X.cols), Xz.sum(0) / Mat.where(few, one, n))
    d = (Xz - mu) * w
    sd = ((d * d).sum(0) / Mat.where(few, one, n - 1.0)).sqrt()
    return Mat.where(few | sd.eq(0.0), one, sd)


def _inverse_batched(G: Mat, p: int) -> Mat:
    """G: lanes x p^2, row i = a p x p matrix (row-major) -> the inverses in the same layout."""
    out = Mat._alloc(G.rows, p * p, L.F64)
    a0, o0 = L.UmView(), L.UmView()
    check(lib.um_view_init(C.byref(a0), C.c_void_p(G._buf.ptr), p, p, L.F64))
    check(lib.um_view_init(C.byref(o0), C.c_void_p(out._buf.ptr), p, p, L.F64))
    check(lib.um_inverse_batched(C.byref(a0), p * p, C.byref(o0), p * p, G.rows))
    return out


def _masked_ols(Y0: Mat, W: Mat, D0: Mat, min_obs: int) -> Mat:
    """nanOls for every column (lane) of Y0 at once.  Y0: obs x lanes with zeros where the observation is unusable,
    W: the 0/1 usability matrix, D0: obs x p design with unusable rows zeroed.  Returns lanes x p coefficients,
    NaN rows for lanes with fewer than min_obs usable observations (Tprf3.scala:807-830)."""
    p, lanes = D0.cols, Y0.cols
    DD = MatD.hstack(*[_col(D0, a) * _col(D0, b) for a in range(p) for b in range(p)])   # obs x p^2
    G = (DD.T @ W).T.copy()                                                               # lanes x p^2
    R = (D0.T @ Y0).T.copy()                                                              # lanes x p
    ok = W.sum(0).T.copy().gte(float(min_obs))                                            # lanes x 1
    okf = ok.astype(L.F64)
    eye = Mat.from_numpy(np.eye(p).reshape(1, p * p))
    Ginv = _inverse_batched(G * okf + eye * (1.0 - okf), p)     # lanes without a fit get the identity (never singular)
    cols = []
    for a in range(p):
        acc = _col(Ginv, a * p) * _col(R, 0)
        for b in range(1, p):
            acc = acc + _col(Ginv, a * p + b) * _col(R, b)
        cols.append(acc)
    return MatD.hstack(*cols) + Mat.where(ok, 0.0, float("nan"))


def _t3prf_pls(yt: Mat, Xs: Mat, Z: Mat, xt: Optional[Mat], min_obs: int):
    """runT3prf's pls branch on the device (Tprf3.scala:985-1048): `Xs` = the window scaled by its 1/std, `xt` the
    held-out row scaled the same way (or None).  Returns (yhat T x 1, yhatt)."""
    T, N, Lz = Xs.rows, Xs.cols, Z.cols
    okx = ~Xs.isnan()
    wx = okx.astype(L.F64)
    Xz = Mat.where(okx, Xs, _zeros_like(Xs))
    n = wx.sum(0)
    has = n.gt(0.0)                                            # nanMeanCols: an empty column has no mean (:506-521)
    hasf = has.astype(L.F64)
    cm0 = Mat.where(has, Xz.sum(0) / Mat.where(has, n, MatD.ones(1, N)), MatD.zeros(1, N))
    Xc0 = (Xz - cm0) * wx                                      # centred; zero where missing
    # pass 1: column i on Z over the rows where x_i and the whole proxy row are present
    zf = (~Z.isnan()).all(1).astype(L.F64)                     # T x 1
    Z0 = _nan_to_zero(Z) * zf
    Phi = _masked_ols(Xc0 * zf, wx * zf, Z0, min_obs)          # N x L
    # pass 2: row t on Phi over the columns where x_t and the loading are present; floor L + 1 (:1013)
    pf = (~Phi.isnan()).all(1).astype(L.F64)                   # N x 1
    Phi0 = _nan_to_zero(Phi) * pf
    pft = pf.T
    Sigma = _masked_ols((Xc0 * pft).T.copy(), (wx * pft).T.copy(), Phi0, Lz + 1)    # T x L
    # pass 3: y on [1 | Sigma] over the rows where both are present, minObs = 1 (:1021-1023)
    Xaug = MatD.hstack(MatD.ones(T, 1), Sigma)
    rf = ((~yt.isnan()) & (~Sigma.isnan()).all(1)).astype(L.F64)
    if float(rf.sum()) < 1.0:
        beta = Mat.from_numpy(np.full((Lz + 1, 1), np.nan))
    else:
        A0 = _nan_to_zero(Xaug) * rf
        beta = (A0.T @ A0).inverse() @ (A0.T @ (_nan_to_zero(yt) * rf))
    yhat = Xaug @ beta
    yhatt = float("nan")
    if xt is not None:                                         # :1025-1036: pass 2 applied to the held-out row
        okt = ~xt.isnan()
        wt = okt.astype(L.F64) * hasf * pft                    # 1 x N
        xc0 = (Mat.where(okt, xt, MatD.zeros(1, N)) - cm0) * wt
        sig = _masked_ols(xc0.T.copy(), wt.T.copy(), Phi0, Lz + 1)                   # 1 x L
        yhatt = float((MatD.hstack(MatD.ones(1, 1), sig) @ beta)[0, 0])
    return yhat, yhatt


def _nan_mean(v: Mat) -> float:
    ok = ~v.isnan()
    k = float(ok.astype(L.F64).sum())
    return float(Mat.where(ok, v, _zeros_like(v)).sum()) / k if k > 0 else float("nan")


def _rows(m: Mat, lo: int, hi: int) -> Mat:
    return m._raw_slice(lo, hi, 0, m.cols)


def _drop_rows(m: Mat, lo: int, hi: int) -> Mat:
    """dropRows (Tprf3.scala:680-690): every row but the contiguous block [lo, hi), copied."""
    parts = [p for p in (_rows(m, 0, lo), _rows(m, hi, m.rows)) if p.rows > 0]
    return MatD.vstack(*parts) if len(parts) > 1 else parts[0].copy()


def estimate3prf(y: Mat, X: Mat, Z: Union[Mat, int], procedure: str = "IS Full", window: Tuple[int, int] = (0, 1),
                 mintrain: Tuple[int, int] = (-1, 0), rollwin: Tuple[int, int, int] = (30, 20, 0), pls: bool = False,
                 computeAvar: bool = False, n_total: int = 0) -> Tprf3Result:
    """Tprf3.estimate3prf (Tprf3.scala:1056-1250) on the device (NaN-free input, or gappy input with pls = true and
    automatic proxies -- the nanOls passes, :985-1048): `Z` a proxy matrix (Right) or an
    int L (Left: automatic proxies); procedures "IS Full", "OOS Recursive", "OOS Cross Val", "OOS Rolling" with the
    reference's `window`, `mintrain`, `rollwin` meaning.  Every window is one fused sweep over its kept rows plus
    one launch for the held-out row; R^2 uses the rolling-mean denominator for the OOS procedures (:1226-1233).
    Unknown procedure -> ValueError, as the reference's IllegalArgumentException (:1201-1204)."""
    if procedure not in _PROCEDURES:
        raise ValueError(f"Unknown procedure '{procedure}'. Choose: 'IS Full', 'OOS Recursive', "
                         "'OOS Cross Val', 'OOS Rolling'")
    auto = not isinstance(Z, Mat)
    eff_pls = bool(pls) and auto                                    # `val effPls = pls && autoproxy` (:1076)
    want_avar = bool(computeAvar) and procedure == "IS Full"        # `if computeAvar && procedure == "IS Full"` (:1253)
    if want_avar and n_total:
        raise NotImplementedError("computeAvar is N x N by definition: not for column-sharded panels")
    if procedure == "IS Full" and not auto:
        res = _fit(L.TPRF_IS_FULL, y, X, Z, n_total)
        if want_avar:
            nan_in = bool(X.isnan().any() or y.isnan().any() or Z.isnan().any())
            res.avar = _avar(X / (_nan_std_cols(X) if nan_in else _std_cols(X)), Z, res.residuals)
        return res
    if n_total:
        raise NotImplementedError("column-sharded panels run IS Full with a proxy matrix only")
    n_auto = int(Z) if auto else 0
    if auto and n_auto < 1:
        raise ValueError(f"number of automatic proxies L = {n_auto} < 1")
    wide = (n_auto if auto else Z.cols) > 8                           # beyond the one-launch window kernel: per-window fits
    # NaN input without pls takes the same matrix passes as clean input (Tprf3.scala:883-943): a NaN in X or Z inside a
    # window spreads to that window's whole fit (every product it touches), a NaN in y only removes its row from pass 3
    # (nanOls, :929-931); `hasNan` just switches the column std to the NaN-aware one (:1078-1080)
    T = y.rows
    oos_batched = procedure != "IS Full" and not eff_pls and not wide # um_tprf_oos_windows standardises the panel itself
    has_nan = False if oos_batched else bool(X.isnan().any() or y.isnan().any() or (not auto and Z.isnan().any()))
    Xn = None if oos_batched else X / (_nan_std_cols(X) if (eff_pls or has_nan) else _std_cols(X))   # Tprf3.scala:1080-1081

    if procedure == "IS Full" and eff_pls:                           # :1097-1106 with runT3prf's pls branch
        r0, fore, zf = y.copy(), None, None
        for _ in range(n_auto):
            fore, _ = _t3prf_pls(y, Xn, r0, None, 10)
            zf = r0                                                  # zFinal: the proxies of the last fit (:1102)
            r0 = MatD.hstack(r0, y - fore)
        resid = y - fore
        loc = ~resid.isnan()
        e, yl = resid[loc], y[loc]
        r2 = float("nan")
        if e.cols > 0:                                               # :1214-1225, no zero guard
            d = yl - yl.mean()
            with np.errstate(all="ignore"):
                r2 = float(1.0 - np.float64((e * e).sum()) / np.float64((d * d).sum()))
        return Tprf3Result(fore, resid, r2, avar=_avar(Xn, zf, resid) if want_avar else None)

    if procedure == "IS Full":                                       # automatic proxies, :1097-1106
        r0, res, zf = y.copy(), None, None
        for j in range(n_auto):
            last = j == n_auto - 1
            res = _fit(L.TPRF_IS_FULL if last else L.TPRF_WINDOW, y, Xn, r0, want_all=last)
            zf = r0                                                  # zFinal: the proxies of the last fit (:1102)
            r0 = MatD.hstack(r0, y - res.forecasts)
        if want_avar:
            res.avar = _avar(Xn, zf, res.residuals)
        return res

    mt = (T // 2, mintrain[1]) if mintrain[0] < 0 else (int(mintrain[0]), int(mintrain[1]))
    win, min_nona, gap = (int(v) for v in rollwin)
    fore = np.full((T, 1), np.nan)
    roll = np.full((T, 1), np.nan)
    if oos_batched:
        # every window of the procedure in ONE launch (um_tprf_oos_windows: one CTA per window, no per-window
        # synchronisation) -- the reference's `.par` loop over the windows (:1117,1151,1177)
        if procedure == "OOS Cross Val":                             # the window drops rows [lo, hi), :1112-1146
            ts = np.arange(T)
            lo = np.maximum(ts - window[0], 0); hi = np.minimum(ts - window[0] + window[1], T)
            kind, min_obs = 1, 10
        elif procedure == "OOS Recursive":                           # rows [0, end), :1148-1179
            ts = np.arange(mt[0] + 1 + mt[1], T)
            lo = np.zeros_like(ts); hi = ts - 1 - mt[1]
            kind, min_obs = 0, (mt[0] if auto else 10)
        else:                                                        # OOS Rolling: rows [lo, hi), :1181-1199
            ts = np.arange(win + 1 + gap, T)
            lo = np.maximum(ts - win - gap, 0); hi = np.minimum(ts - 1 - gap, T)
            kind, min_obs = 0, (min_nona if auto else 10)
        if ts.size:
            hi = np.maximum(hi, lo)
            lo32, hi32, ts32 = (np.ascontiguousarray(a, dtype=np.int32) for a in (lo, hi, ts))
            f = np.empty(ts.size); rl = np.empty(ts.size)
            ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
            dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
            Xc = X if (X.cs == 1 or X.cols == 1) else X.copy()      # the RAW panel: the call standardises it (nanStdCols) itself
            check(lib.um_tprf_oos_windows(y._ref(), Xc._ref(), None if auto else Z._ref(), n_auto, kind, ip(lo32), ip(hi32), ip(ts32),
                                          ts.size, int(min_obs), 1, dp(f), dp(rl)))
            fore[ts, 0] = f
            roll[ts, 0] = rl
    elif procedure == "OOS Cross Val":                               # :1112-1146
        for t in range(T):
            lo, hi = max(t - window[0], 0), min(t - window[0] + window[1], T)
            yt = _drop_rows(y, lo, hi)
            fore[t, 0] = (_window_forecast_pls(yt, _drop_rows(Xn, lo, hi), n_auto, _rows(Xn, t, t + 1), 10) if eff_pls else
                          _window_forecast(yt, _drop_rows(Xn, lo, hi), None if auto else _drop_rows(Z, lo, hi), n_auto,
                                           _rows(Xn, t, t + 1), 10))
            roll[t, 0] = _nan_mean(yt) if (eff_pls or has_nan) else yt.mean()   # colMean(yt, hasNan)
    elif procedure == "OOS Recursive":                               # :1148-1179
        for t in range(mt[0] + 1 + mt[1], T):
            end = t - 1 - mt[1]
            yt = _rows(y, 0, end)
            fore[t, 0] = (_window_forecast_pls(yt, _rows(Xn, 0, end), n_auto, _rows(Xn, t, t + 1), mt[0]) if eff_pls else
                          _window_forecast(yt, _rows(Xn, 0, end), None if auto else _rows(Z, 0, end), n_auto,
                                           _rows(Xn, t, t + 1), mt[0]))
            roll[t, 0] = _nan_mean(yt) if (eff_pls or has_nan) else yt.mean()   # colMean(yt, hasNan)
    else:                                                            # OOS Rolling, :1181-1199
        for t in range(win + 1 + gap, T):
            lo, hi = max(t - win - gap, 0), min(t - 1 - gap, T)
            yt = _rows(y, lo, hi)
            fore[t, 0] = (_window_forecast_pls(yt, _rows(Xn, lo, hi), n_auto, _rows(Xn, t, t + 1), min_nona) if eff_pls else
                          _window_forecast(yt, _rows(Xn, lo, hi), None if auto else _rows(Z, lo, hi), n_auto,
                                           _rows(Xn, t, t + 1), min_nona))
            roll[t, 0] = _nan_mean(yt) if (eff_pls or has_nan) else yt.mean()   # colMean(yt, hasNan)

    # point estimates over the rows that have a forecast (`loc`, :1207): T-long vectors, summed on the host in the
    # reference's order (sequential while loops, :1226-1238) -- one copy of y down, three small vectors up
    yh = y.to_numpy()
    res = yh - fore
    loc = ~np.isnan(res[:, 0])
    r2 = enc = float("nan")
    if loc.any():
        e, rl = res[loc, 0], roll[loc, 0]
        with np.errstate(all="ignore"):
            see = float(np.add.reduce(e * e))
            d = yh[loc, 0] - rl
            sst = float(np.add.reduce(d * d))
            if sst != 0.0:
                r2 = 1.0 - see / sst                                 # :1226-1233 (ssT == 0 -> NaN)
            if procedure == "OOS Recursive":                         # encnew(rollfore[loc], residuals[loc]), :854-860
                enc = float(np.float64(e.size) * np.float64(np.add.reduce(rl * rl - rl * e)) / np.float64(see))
    forecasts, rollfore, residuals = Mat.from_numpy(fore), Mat.from_numpy(roll), Mat.from_numpy(res)
    return Tprf3Result(forecasts, residuals, r2, rollfore=rollfore, encnew=enc)


def forecast3prf(y: Mat, X: Mat, Z: Union[Mat, int], procedure: str = "IS Full", **kw) -> Mat:
    """Tprf3.forecast3prf (Tprf3.scala:1290-1299)."""
    return estimate3prf(y, X, Z, procedure, **kw).forecasts


@dataclass
class Pls3prfModel:
    """Pls3prfModel (Tprf3.scala:299-339): the fitted PLS-variant 3PRF with the column mean / scale it normalised X
    by, so `predict` takes a raw row."""
    phi: Mat          # N x 1  pass-1 loadings
    sigma: Mat        # T x 1  pass-2 factor scores
    beta: Mat         # 2 x 1  pass-3 [intercept, slope]
    forecasts: Mat    # T x 1
    colMean: Mat      # 1 x N  column means of the scaled X
    colStd: Mat       # 1 x N  column std of the raw X
    rSquared: float

    def __post_init__(self):
        self._w = self.phi.T / self.colStd                      # 1 x N: phi_j / std_j
        self._shift = float((self.colMean @ self.phi)[0, 0])    # sum_j mean_j phi_j
        self._phi_ss = float((self.phi * self.phi).sum())
        b = self.beta.to_numpy()
        self._b0, self._b1 = float(b[0, 0]), float(b[1, 0])

    def predictAll(self, rows) -> Mat:
        """Forecast for each raw predictor row (Tprf3.scala:335-339): sigma_new = (row/std - mean).phi / phi'phi,
        yhat = b0 + b1 sigma_new; `rows` is a k x N device matrix (or anything Mat.from_numpy takes)."""
        r = rows if isinstance(rows, Mat) else Mat.from_numpy(np.asarray(rows, dtype=np.float64).reshape(-1, self.phi.rows))
        if r.cols != self.phi.rows:
            raise ValueError(f"row length {r.cols} != {self.phi.rows} predictors")
        return ((r @ self._w.T) - self._shift) * (self._b1 / self._phi_ss) + self._b0

    def predict(self, row) -> float:
        """Forecast for one raw (un-normalised) predictor row (Tprf3.scala:325-333)."""
        a = np.asarray(row, dtype=np.float64).reshape(1, -1) if not isinstance(row, Mat) else row
        n = a.cols if isinstance(a, Mat) else a.shape[1]
        if n != self.phi.rows:
            raise ValueError(f"row length {n} != {self.phi.rows} predictors")
        return float(self.predictAll(a)[0, 0])


def plsClosedForm(y: Mat, X: Mat) -> Pls3prfModel:
    """Tprf3.plsClosedForm (Tprf3.scala:360-390): the 3PRF with one automatic proxy (= y) and no intercept in passes
    1 and 2, which coincides with one-component PLS-1.  Device work: one fused sweep with Z = y (its pass-1 slopes
    are cov(x_j, y) / ssy_centred; the PLS loading is the same cross product over y'y), one column-mean pass and
    one T x N by N x 1 product for the factor scores; no T x N temporary is formed."""
    if X.rows != y.rows:
        raise ValueError(f"X has {X.rows} rows but y has {y.rows}")
    if X.isnan().any() or y.isnan().any():
        raise ValueError("plsClosedForm requires NaN-free input; use estimate3prf(pls = true) for gappy data")
    T = X.rows
    fit = _fit(L.TPRF_ITER, y, X, y)
    col_std = fit.xstd                                            # stdCols(X): 1 x N
    col_mean = X.mean(0) / col_std                                # nanMeanCols(X / colStd)
    yc = y - y.mean()
    yss = (y * y).sum()
    phi = fit.phi * ((yc * yc).sum() / yss)                       # Xc'y / y'y (Xc is column-centred: Xc'y = Xc'(y - ybar))
    phi_ss = (phi * phi).sum()
    w = phi / col_std.T                                           # N x 1
    sigma = ((X @ w) - float((col_mean @ phi)[0, 0])) / phi_ss    # Xc phi / phi'phi
    xaug = MatD.hstack(MatD.ones(T, 1), sigma)
    beta = (xaug.T @ xaug).inverse() @ (xaug.T @ y)
    forecasts = xaug @ beta
    resid = y - forecasts
    ssy = (yc * yc).sum()
    rsq = 0.0 if ssy == 0.0 else 1.0 - (resid * resid).sum() / ssy
    return Pls3prfModel(phi, sigma, beta, forecasts, col_mean, col_std, rsq)


def pls1Fit(x, y) -> Pls3prfModel:
    """Tprf3.pls1Fit (Tprf3.scala:392-401): plsClosedForm from host arrays, `x` row-major T x N."""
    xa = np.asarray(x, dtype=np.float64)
    ya = np.asarray(y, dtype=np.float64).reshape(-1)
    if xa.size == 0:
        raise ValueError("x is empty")
    if xa.shape[0] != ya.size:
        raise ValueError(f"x has {xa.shape[0]} rows but y has {ya.size}")
    return plsClosedForm(Mat.from_numpy(ya.reshape(-1, 1)), Mat.from_numpy(xa))


TPRF_PATH_AUTO, TPRF_PATH_TWO_SWEEP, TPRF_PATH_FUSED, TPRF_PATH_GRID = 0, 1, 2, 3


def set_tprf_path(path: int) -> None:
    """0 = auto (single-read sweeps whenever they apply), 1 = two sweeps only, 2 = cluster sweep or
    ValueError, 3 = grid sweep (virtual CTA groups, tensor-memory stash) or ValueError.  Process-wide switch for tests and profiling (um_tprf_set_path)."""
    check(lib.um_tprf_set_path(int(path)))
