"""Host-side mirror of `uni.stats.Tprf3` (src/main/scala/uni/stats/Tprf3.scala) over libunimat.

`tprfClosedForm`, `t3prf`, `estimate3prf(..., "IS Full")` and `forecast3prf` keep the
reference's names, argument meaning and error behaviour; the work is one fused device sweep
plus an O(T L^2) finish (uni_b200/csrc/tprf.cu).  Under an active communicator
(`uni_b200.dist`) X is this rank's column shard and `n_total` the global predictor count.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Union

from . import _lib as L
from ._lib import UmTprfOut, check, lib
from .mat import Mat


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
        if mode in (L.TPRF_ITER, L.TPRF_IS_FULL):
            res.phi = Mat._alloc(N, Lp, L.F64)
            res.phiHat = Mat._alloc(Lp + 1, N, L.F64)
            res.sigma = Mat._alloc(T, Lp, L.F64)
            res.beta3 = Mat._alloc(Lp + 1, 1, L.F64)
            out.phi, out.phi_hat = res.phi._buf.ptr, res.phiHat._buf.ptr
            out.sigma, out.beta3 = res.sigma._buf.ptr, res.beta3._buf.ptr
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


def estimate3prf(y: Mat, X: Mat, Z: Union[Mat, int], procedure: str = "IS Full", n_total: int = 0) -> Tprf3Result:
    """Tprf3.estimate3prf (Tprf3.scala:1056-1250), `Right(Z)` + "IS Full" on the device.
    Unknown procedure -> ValueError, as the reference's IllegalArgumentException (:1201-1204)."""
    if procedure not in _PROCEDURES:
        raise ValueError(f"Unknown procedure '{procedure}'. Choose: 'IS Full', 'OOS Recursive', "
                         "'OOS Cross Val', 'OOS Rolling'")
    if procedure != "IS Full":
        raise NotImplementedError(f"'{procedure}' is a SURVEY.md section 8(f) 'next' row; only IS Full runs on the device")
    if not isinstance(Z, Mat):
        raise NotImplementedError("autoproxy (Left(L)) is not on the device path yet")
    return _fit(L.TPRF_IS_FULL, y, X, Z, n_total)


def forecast3prf(y: Mat, X: Mat, Z: Union[Mat, int], procedure: str = "IS Full") -> Mat:
    """Tprf3.forecast3prf (Tprf3.scala:1290-1299)."""
    return estimate3prf(y, X, Z, procedure).forecasts


TPRF_PATH_AUTO, TPRF_PATH_TWO_SWEEP, TPRF_PATH_FUSED, TPRF_PATH_GRID = 0, 1, 2, 3


def set_tprf_path(path: int) -> None:
    """0 = auto (single-read sweeps whenever they apply), 1 = two sweeps only, 2 = cluster sweep or
    ValueError, 3 = grid sweep (virtual CTA groups, tensor-memory stash) or ValueError.  Process-wide switch for tests and profiling (um_tprf_set_path)."""
    check(lib.um_tprf_set_path(int(path)))
