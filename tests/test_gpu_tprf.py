"""GPU parity tests for the 3PRF fits: CUDA vs the reference's golden vectors and the CPU oracle.
Tolerance: |a-b| <= 1e-12 + 1e-9 max(|a|,|b|) (Tprf3ParitySuite.scala:28-45; north_star 1e-9)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import tprf_oracle as to


@pytest.fixture(scope="module")
def u():
    import uni_b200
    if uni_b200.device_count() == 0:
        pytest.fail("no CUDA device: the product path has no CPU fallback")
    uni_b200.init(0)
    return uni_b200


def close(a, b, rtol=1e-9, atol=1e-12):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    both_nan = np.isnan(a) & np.isnan(b)
    ok = np.abs(a - b) <= atol + rtol * np.maximum(np.abs(a), np.abs(b))
    return bool(np.all(ok | both_nan))


def devs(u, y, X, Z):
    return u.Mat.from_numpy(y), u.Mat.from_numpy(X), u.Mat.from_numpy(Z)


def test_kp_known_answers(u, golden_dir):
    """BASELINE config 1: t3prf-validation reference panel vs the K&P MATLAB outputs."""
    d = os.path.join(golden_dir, "t3prf-validation")
    X = np.loadtxt(os.path.join(d, "ref_X.csv"), delimiter=",")
    Z = np.loadtxt(os.path.join(d, "ref_Z.csv"), delimiter=",")
    y = np.loadtxt(os.path.join(d, "ref_y.csv"), delimiter=",").reshape(-1, 1)
    kp_yhat = np.loadtxt(os.path.join(d, "kp_yhat.csv"), delimiter=",").reshape(-1, 1)
    kp_alpha = np.loadtxt(os.path.join(d, "kp_alpha.csv"), delimiter=",").reshape(-1, 1)
    kp_r2 = float(np.loadtxt(os.path.join(d, "kp_rsquare.csv"), delimiter=","))
    dy, dX, dZ = devs(u, y, X, Z)
    for fit in (u.tprfClosedForm, u.t3prf, lambda a, b, c: u.estimate3prf(a, b, c, "IS Full")):
        r = fit(dy, dX, dZ)
        assert close(r.forecasts.to_numpy(), kp_yhat)
        assert close(r.rSquared, kp_r2)
        assert close(r.residuals.to_numpy(), y - kp_yhat, atol=1e-11)
        if r.alpha is not None:
            assert close(r.alpha.to_numpy(), kp_alpha)


@pytest.mark.parametrize("case", ["T100N10L2", "T140N12L3", "T60N25L2"])
def test_scala_parity_rows(u, golden_dir, case):
    d = os.path.join(golden_dir, "tprf3-parity")
    ref = {}
    with open(os.path.join(d, "scala-reference.txt")) as fh:
        for line in fh:
            if line.startswith("#") or not line.strip():
                continue
            tag, key, val = line.split()
            ref[(tag, key)] = float(val)
    X = np.loadtxt(os.path.join(d, f"{case}_X.csv"), delimiter=",", ndmin=2)
    y = np.loadtxt(os.path.join(d, f"{case}_y.csv"), delimiter=",", ndmin=2).reshape(-1, 1)
    Z = np.loadtxt(os.path.join(d, f"{case}_Z.csv"), delimiter=",", ndmin=2)
    T = X.shape[0]
    dy, dX, dZ = devs(u, y, X, Z)
    isf = u.estimate3prf(dy, dX, dZ, "IS Full")
    want = np.array([ref[(f"{case}/isfull", f"f{i}")] for i in range(T)]).reshape(-1, 1)
    assert close(isf.forecasts.to_numpy(), want) and close(isf.rSquared, ref[(f"{case}/isfull", "r2")])
    want = np.array([ref[(f"{case}/closed", f"f{i}")] for i in range(T)]).reshape(-1, 1)
    for fit in (u.tprfClosedForm, u.t3prf):
        r = fit(dy, dX, dZ)
        assert close(r.forecasts.to_numpy(), want) and close(r.rSquared, ref[(f"{case}/closed", "r2")])


def test_python_reference_vectors(u, golden_dir):
    d = os.path.join(golden_dir, "tprf_pyref")
    for f in sorted(x for x in os.listdir(d) if x.endswith(".npz")):
        g = np.load(os.path.join(d, f))
        rng = np.random.default_rng(int(g["seed"]))
        T, N, L = int(g["T"]), int(g["N"]), int(g["L"])
        X = rng.standard_normal((T, N)); y = rng.standard_normal((T, 1)); Z = rng.standard_normal((T, L))
        r = u.estimate3prf(*devs(u, y, X, Z))
        assert close(r.forecasts.to_numpy(), g["forecasts"]), f
        assert close(r.alpha.to_numpy(), g["alpha"], rtol=1e-8), f
        assert abs(r.rSquared - float(g["rsquare"])) < 1e-10, f


def test_compute_avar_vs_python_reference(u, golden_dir):
    """estimate3prf("IS Full", computeAvar = true) on the device (Tprf3.scala:1252-1277) against the reference's NumPy
    outputs: avar.alpha (N x N) and avar.forecasts (T x T), proxy matrix and automatic proxies; rule of
    Tprf3ParitySuite (atol 1e-12 + rtol 1e-9), loosened to 1e-8 for the products of products."""
    d = os.path.join(golden_dir, "tprf_avar_pyref")
    names = sorted(x for x in os.listdir(d) if x.endswith(".npz"))
    assert len(names) >= 5
    for f in names:
        g = np.load(os.path.join(d, f))
        rng = np.random.default_rng(int(g["seed"]))
        T, N, L = int(g["T"]), int(g["N"]), int(g["L"])
        X = rng.standard_normal((T, N)); y = rng.standard_normal((T, 1)); Z = rng.standard_normal((T, L))
        dy, dX, dZ = devs(u, y, X, Z)
        r = u.estimate3prf(dy, dX, L if bool(g["auto"]) else dZ, "IS Full", computeAvar=True)
        assert close(r.forecasts.to_numpy(), g["forecasts"]), f
        assert r.avar is not None and r.avar.alpha.shape == (N, N) and r.avar.forecasts.shape == (T, T)
        assert close(r.avar.alpha.to_numpy(), g["avar_alpha"], rtol=1e-8), f
        assert close(r.avar.forecasts.to_numpy(), g["avar_forecasts"], rtol=1e-8), f
    # ignored outside "IS Full", absent by default
    assert u.estimate3prf(dy, dX, dZ, "OOS Recursive", computeAvar=True).avar is None
    assert u.estimate3prf(dy, dX, dZ).avar is None


SHAPES = [(10, 3, 1), (33, 7, 1), (50, 6, 2), (100, 64, 3), (127, 65, 2), (129, 200, 8), (512, 2048, 4), (300, 1000, 5), (1000, 130, 7)]


@pytest.mark.parametrize("tnl", SHAPES)
def test_all_outputs_vs_oracle(u, tnl):
    T, N, L = tnl
    rng = np.random.default_rng(T + N + L)
    X = rng.standard_normal((T, N)) * rng.uniform(0.5, 3.0, size=(1, N)) + rng.uniform(-2, 2, size=(1, N))
    y = rng.standard_normal((T, 1)) + 0.3 * X[:, :1]
    Z = rng.standard_normal((T, L)) + 0.2 * X[:, : min(L, N)].mean(axis=1, keepdims=True)
    dy, dX, dZ = devs(u, y, X, Z)
    oc = to.tprf_closed_form_literal(y, X, Z) if N <= 300 else to.tprf_closed_form_gram(y, X, Z)
    c = u.tprfClosedForm(dy, dX, dZ)
    assert close(c.forecasts.to_numpy(), oc.forecasts) and close(c.rSquared, oc.r_squared)
    assert close(c.alpha.to_numpy(), oc.alpha, rtol=1e-8, atol=1e-13)
    assert close(c.xstd.to_numpy(), to.nan_std_cols(X), rtol=1e-12)
    oi = to.t3prf(y, X, Z)
    it = u.t3prf(dy, dX, dZ)
    assert close(it.forecasts.to_numpy(), oi.forecasts) and close(it.rSquared, oi.r_squared)
    assert close(it.phi.to_numpy(), oi.phi, rtol=1e-8, atol=1e-12)
    assert close(it.phiHat.to_numpy(), oi.phi_hat, rtol=1e-8, atol=1e-12)
    assert close(it.sigma.to_numpy(), oi.sigma, rtol=1e-8, atol=1e-11)
    assert close(it.beta3.to_numpy(), oi.beta3, rtol=1e-7, atol=1e-11)
    # tprfClosedForm carries the pass-1/2/3 fields too (Tprf3.scala:211-231,243-252): they are t3prf's
    assert close(c.phi.to_numpy(), oi.phi, rtol=1e-8, atol=1e-12) and close(c.phiHat.to_numpy(), oi.phi_hat, rtol=1e-8, atol=1e-12)
    assert close(c.sigma.to_numpy(), oi.sigma, rtol=1e-8, atol=1e-11) and close(c.beta3.to_numpy(), oi.beta3, rtol=1e-7, atol=1e-11)
    if N <= 300:
        assert close(c.phi.to_numpy(), oc.phi, rtol=1e-8, atol=1e-12) and close(c.sigma.to_numpy(), oc.sigma, rtol=1e-8, atol=1e-11)
        assert close(c.beta3.to_numpy(), oc.beta3, rtol=1e-7, atol=1e-11) and close(c.phiHat.to_numpy(), oc.phi_hat, rtol=1e-8, atol=1e-12)
    # tprfClosedForm == t3prf == IS Full (TprfSuite.scala:20-34)
    assert close(c.forecasts.to_numpy(), it.forecasts.to_numpy())
    if T >= 10:
        isf = u.estimate3prf(dy, dX, dZ)
        assert close(isf.forecasts.to_numpy(), it.forecasts.to_numpy(), rtol=1e-12)
    # strided inputs: Z as a column slice of a wider matrix, y as a column of a T x 3 matrix
    ZZ = u.Mat.from_numpy(np.column_stack([np.zeros(T), Z, np.ones(T)]))._raw_slice(0, T, 1, 1 + L)
    yy = u.Mat.from_numpy(np.column_stack([y, y * 2, y * 3]))._raw_slice(0, T, 0, 1)
    c2 = u.tprfClosedForm(yy, dX, ZZ)
    assert np.array_equal(c2.forecasts.to_numpy(), c.forecasts.to_numpy())


WIDE_SHAPES = [(60, 40, 9), (120, 300, 12), (200, 64, 16), (512, 2048, 10), (1000, 130, 24), (40, 100, 9), (1100, 1300, 17), (300, 4096, 33)]


@pytest.mark.parametrize("tnl", WIDE_SHAPES)
def test_more_than_eight_proxies_vs_oracle(u, tnl):
    """Tprf3.scala:199-289 take any L; beyond the 8 proxies of the fused sweeps the fit runs on the library's operators
    (uni_b200/csrc/tprf_wide.cu): every field of all three fits against the oracle."""
    T, N, L = tnl
    rng = np.random.default_rng(T + N + L)
    X = rng.standard_normal((T, N)) * rng.uniform(0.5, 3.0, size=(1, N)) + rng.uniform(-2, 2, size=(1, N))
    y = rng.standard_normal((T, 1)) + 0.3 * X[:, :1]
    Z = rng.standard_normal((T, L)) + 0.2 * X[:, : min(L, N)].mean(axis=1, keepdims=True)
    dy, dX, dZ = devs(u, y, X, Z)
    oc = to.tprf_closed_form_literal(y, X, Z) if N <= 300 else to.tprf_closed_form_gram(y, X, Z)
    oi = to.t3prf(y, X, Z)
    c = u.tprfClosedForm(dy, dX, dZ)
    assert close(c.forecasts.to_numpy(), oc.forecasts) and close(c.rSquared, oc.r_squared)
    assert close(c.residuals.to_numpy(), y - oc.forecasts, atol=1e-11)
    assert close(c.alpha.to_numpy(), oc.alpha, rtol=1e-8, atol=1e-13)
    assert close(c.xstd.to_numpy(), to.nan_std_cols(X), rtol=1e-12)
    it = u.t3prf(dy, dX, dZ)
    assert close(it.forecasts.to_numpy(), oi.forecasts) and close(it.rSquared, oi.r_squared)
    for r in (it, c):                      # tprfClosedForm carries the pass-1/2/3 fields too (Tprf3.scala:211-231)
        assert r.phi.shape == (N, L) and r.phiHat.shape == (L + 1, N) and r.sigma.shape == (T, L) and r.beta3.shape == (L + 1, 1)
        assert close(r.phi.to_numpy(), oi.phi, rtol=1e-8, atol=1e-12)
        assert close(r.phiHat.to_numpy(), oi.phi_hat, rtol=1e-8, atol=1e-12)
        assert close(r.sigma.to_numpy(), oi.sigma, rtol=1e-8, atol=1e-11)
        assert close(r.beta3.to_numpy(), oi.beta3, rtol=1e-7, atol=1e-11)
    of = to.estimate3prf_is_full(y, X, Z)
    isf = u.estimate3prf(dy, dX, dZ)
    assert close(isf.forecasts.to_numpy(), of.forecasts) and close(isf.rSquared, of.r_squared)
    assert close(isf.alpha.to_numpy(), of.alpha, rtol=1e-8, atol=1e-13)
    # strided inputs: X and Z as column slices of wider matrices (row pitch > cols), y as a column of a T x 3 matrix
    XX = u.Mat.from_numpy(np.column_stack([np.full(T, 7.0), X, np.zeros((T, 2))]))._raw_slice(0, T, 1, 1 + N)
    ZZ = u.Mat.from_numpy(np.column_stack([np.zeros(T), Z, np.ones(T)]))._raw_slice(0, T, 1, 1 + L)
    yy = u.Mat.from_numpy(np.column_stack([y, y * 2, y * 3]))._raw_slice(0, T, 0, 1)
    c2 = u.tprfClosedForm(yy, XX, ZZ)
    assert close(c2.forecasts.to_numpy(), oc.forecasts) and close(c2.alpha.to_numpy(), oc.alpha, rtol=1e-8, atol=1e-13)


def test_more_than_eight_proxies_procedures_and_errors(u):
    """Out-of-sample procedures and automatic proxies with L > 8 (per-window fits on the operator path), the
    reference's singular-matrix error, constant columns and T < minObs."""
    rng = np.random.default_rng(5)
    T, N, L = 80, 60, 9
    X = rng.standard_normal((T, N)); y = rng.standard_normal((T, 1)) + 0.5 * X[:, :3].sum(axis=1, keepdims=True)
    Z = rng.standard_normal((T, L))
    dy, dX, dZ = devs(u, y, X, Z)
    r = u.estimate3prf(dy, dX, dZ, "OOS Recursive", mintrain=(60, 0))
    o = to.estimate3prf(y, X, Z, "OOS Recursive", mintrain=(60, 0))
    assert close(r.forecasts.to_numpy(), o.forecasts) and close(r.rSquared, o.r_squared) and close(r.encnew, o.extra["enc"], rtol=1e-8)
    r = u.estimate3prf(dy, dX, dZ, "OOS Cross Val", window=(1, 3))
    o = to.estimate3prf(y, X, Z, "OOS Cross Val", window=(1, 3))
    assert close(r.forecasts.to_numpy(), o.forecasts) and close(r.rSquared, o.r_squared)
    r = u.estimate3prf(dy, dX, 9, "IS Full")                    # nine automatic proxies (Tprf3.scala:1097-1106)
    o = to.estimate3prf(y, X, 9, "IS Full")
    assert close(r.forecasts.to_numpy(), o.forecasts, rtol=1e-7, atol=1e-9) and close(r.rSquared, o.r_squared, rtol=1e-7)
    # duplicated proxy -> ArithmeticException("Matrix is singular or nearly singular") (Mat.scala:990)
    z = np.round(rng.standard_normal((T, 1)) * 8) / 8
    Zdup = np.column_stack([Z[:, :8], z, z])
    for fit in (u.tprfClosedForm, u.t3prf, u.estimate3prf):
        with pytest.raises(u.SingularMatrixError):
            fit(*devs(u, y, X, Zdup))
    # constant column -> sd := 1.0; T < minObs -> NaN fit; ssy == 0 -> R^2 = 0
    Xc = X.copy(); Xc[:, 3] = 2.5
    rc = u.tprfClosedForm(*devs(u, y, Xc, Z))
    assert rc.xstd.to_numpy()[0, 3] == 1.0
    assert close(rc.forecasts.to_numpy(), to.tprf_closed_form_literal(y, Xc, Z).forecasts)
    s = u.estimate3prf(*devs(u, y[:5], X[:5], Z[:5]))
    assert np.isnan(s.forecasts.to_numpy()).all() and np.isnan(s.rSquared)
    assert u.t3prf(*devs(u, np.ones((T, 1)), X, Z)).rSquared == 0.0
    Xn = X.copy(); Xn[7, 2] = np.nan
    assert np.isnan(u.tprfClosedForm(*devs(u, y, Xn, Z)).forecasts.to_numpy()).all()


def test_more_than_eight_proxies_batched_and_host_calls(u):
    """um_tprf_fit_batched and um_tprf_fit_host with L > 8: panel by panel / the panel made resident once."""
    import ctypes as C
    from uni_b200 import _lib as L_
    B, T, N, L = 3, 64, 200, 11
    rng = [np.random.default_rng(40 + p) for p in range(B)]
    Xs = np.stack([g.standard_normal((T, N)) for g in rng]); ys = np.stack([g.standard_normal((T, 1)) for g in rng])
    Zs = np.stack([g.standard_normal((T, L)) for g in rng])
    dX = u.Mat.from_numpy(Xs.reshape(B * T, N)); dy = u.Mat.from_numpy(ys.reshape(B * T, 1)); dZ = u.Mat.from_numpy(Zs.reshape(B * T, L))
    f = u.MatD.zeros(B * T, 1); a = u.MatD.zeros(B * N, 1); r2 = u.MatD.zeros(B, 1)
    L_.check(L_.lib.um_tprf_fit_batched(L_.TPRF_CLOSED, dy._buf.ptr, dX._buf.ptr, dZ._buf.ptr, T, N, L, B, f._buf.ptr, a._buf.ptr, r2._buf.ptr))
    F = f.to_numpy().reshape(B, T); A = a.to_numpy().reshape(B, N); R = r2.to_numpy().ravel()
    for p in range(B):
        o = to.tprf_closed_form_gram(ys[p], Xs[p], Zs[p])
        assert close(F[p], o.forecasts.ravel()) and close(A[p], o.alpha.ravel(), rtol=1e-8) and close(R[p], o.r_squared)
    T, N, L = 96, 5000, 10
    g = np.random.default_rng(9)
    Xw = g.standard_normal((T, N + 6)); X = Xw[:, 3 : N + 3]            # a host panel with a row pitch > N
    y = g.standard_normal((T, 1)); Z = g.standard_normal((T, L))
    fh = np.zeros(T); ah = np.zeros(N); r2h = C.c_double(0)
    vp = lambda arr: arr.ctypes.data_as(C.c_void_p)
    L_.check(L_.lib.um_tprf_fit_host(L_.TPRF_CLOSED, vp(y), C.c_void_p(Xw.ctypes.data + 3 * 8), N + 6, vp(Z), T, N, L, N, vp(fh), vp(ah), C.byref(r2h)))
    o = to.tprf_closed_form_gram(y, np.ascontiguousarray(X), Z)
    assert close(fh, o.forecasts.ravel()) and close(ah, o.alpha.ravel(), rtol=1e-8, atol=1e-13) and close(r2h.value, o.r_squared)


def test_edge_semantics(u):
    """TprfCoverageSuite.scala:176-230."""
    rng = np.random.default_rng(4)
    T, N, L = 40, 9, 2
    X = rng.standard_normal((T, N)); X[:, 3] = 2.5          # constant column -> sd := 1.0
    y = rng.standard_normal((T, 1)); Z = rng.standard_normal((T, L))
    r = u.tprfClosedForm(*devs(u, y, X, Z))
    assert r.xstd.to_numpy()[0, 3] == 1.0
    assert close(r.forecasts.to_numpy(), to.tprf_closed_form_literal(y, X, Z).forecasts)
    # T < minObs (10) -> all-NaN forecasts, NaN R^2 (Tprf3.scala:893-894)
    s = u.estimate3prf(*devs(u, y[:5], X[:5], Z[:5]))
    assert np.isnan(s.forecasts.to_numpy()).all() and np.isnan(s.rSquared)
    # ssy == 0 -> R^2 = 0 in tprfClosedForm / t3prf (Tprf3.scala:239-241)
    r0 = u.t3prf(*devs(u, np.ones((T, 1)), X, Z))
    assert r0.rSquared == 0.0
    with pytest.raises(ValueError):
        u.estimate3prf(*devs(u, y, X, Z), procedure="bogus")
    with pytest.raises(ValueError):
        u.tprfClosedForm(*devs(u, y[:-1], X, Z))
    # a NaN in X poisons the fit, as the reference's matmuls do
    Xn = X.copy(); Xn[7, 2] = np.nan
    assert np.isnan(u.tprfClosedForm(*devs(u, y, Xn, Z)).forecasts.to_numpy()).all()


def test_batched_matches_single(u):
    import ctypes as C
    from uni_b200 import _lib as L_
    B, T, N, L = 6, 64, 200, 4
    rng = [np.random.default_rng(p) for p in range(B)]
    Xs = np.stack([g.standard_normal((T, N)) for g in rng]); ys = np.stack([g.standard_normal((T, 1)) for g in rng])
    Zs = np.stack([g.standard_normal((T, L)) for g in rng])
    dX = u.Mat.from_numpy(Xs.reshape(B * T, N)); dy = u.Mat.from_numpy(ys.reshape(B * T, 1)); dZ = u.Mat.from_numpy(Zs.reshape(B * T, L))
    f = u.MatD.zeros(B * T, 1); a = u.MatD.zeros(B * N, 1); r2 = u.MatD.zeros(B, 1)
    L_.check(L_.lib.um_tprf_fit_batched(L_.TPRF_CLOSED, dy._buf.ptr, dX._buf.ptr, dZ._buf.ptr, T, N, L, B, f._buf.ptr, a._buf.ptr, r2._buf.ptr))
    F = f.to_numpy().reshape(B, T); A = a.to_numpy().reshape(B, N); R = r2.to_numpy().ravel()
    for p in range(B):
        o = to.tprf_closed_form_gram(ys[p], Xs[p], Zs[p])
        assert close(F[p], o.forecasts.ravel()) and close(A[p], o.alpha.ravel(), rtol=1e-8) and close(R[p], o.r_squared)


def test_batched_more_panels_than_a_grid_dimension(u):
    """70 000 tiny panels in one call (the panel index is a grid dimension <= 65535: the call slices): sampled panels on
    either side of the slice boundary against the oracle, every R^2 finite."""
    from uni_b200 import _lib as L_
    B, T, N, L = 70000, 16, 32, 2
    dX = u.MatD.randn(B * T, N, seed=5); dy = u.MatD.randn(B * T, 1, seed=6); dZ = u.MatD.randn(B * T, L, seed=7)
    f = u.MatD.zeros(B * T, 1); r2 = u.MatD.zeros(B, 1)
    L_.check(L_.lib.um_tprf_fit_batched(L_.TPRF_CLOSED, dy._buf.ptr, dX._buf.ptr, dZ._buf.ptr, T, N, L, B, f._buf.ptr, None, r2._buf.ptr))
    R = r2.to_numpy().ravel()
    assert np.isfinite(R).all()
    for p in (0, 1, 65534, 65535, 65536, 69999):
        Xp = dX.slice(p * T, (p + 1) * T, 0, N).to_numpy(); yp = dy.slice(p * T, (p + 1) * T, 0, 1).to_numpy()
        Zp = dZ.slice(p * T, (p + 1) * T, 0, L).to_numpy()
        o = to.tprf_closed_form_gram(yp, Xp, Zp)
        assert close(f.slice(p * T, (p + 1) * T, 0, 1).to_numpy().ravel(), o.forecasts.ravel()) and close(R[p], o.r_squared)


def test_host_streamed_fit(u):
    import ctypes as C
    from uni_b200 import _lib as L_
    T, N, L = 96, 20000, 3
    rng = np.random.default_rng(8)
    X = rng.standard_normal((T, N)); y = rng.standard_normal((T, 1)); Z = rng.standard_normal((T, L))
    f = np.zeros(T); a = np.zeros(N); r2 = C.c_double(0)
    vp = lambda arr: arr.ctypes.data_as(C.c_void_p)
    L_.check(L_.lib.um_tprf_fit_host(L_.TPRF_CLOSED, vp(y), vp(X), N, vp(Z), T, N, L, N, vp(f), vp(a), C.byref(r2)))
    o = to.tprf_closed_form_gram(y, X, Z)
    assert close(f, o.forecasts.ravel()) and close(a, o.alpha.ravel(), rtol=1e-8, atol=1e-13) and close(r2.value, o.r_squared)


def test_full_size_properties(u):
    """BASELINE config 4 at full size (T=8192, N=262144, L=8, 16 GiB panel generated in HBM):
    closed form and three-pass fit must agree (TprfSuite.scala:20-34); alpha reproduces yhat."""
    T, N, L = 8192, 262144, 8
    X = u.MatD.randn(T, N, seed=1); y = u.MatD.randn(T, 1, seed=2); Z = u.MatD.randn(T, L, seed=3)
    c = u.tprfClosedForm(y, X, Z)
    it = u.t3prf(y, X, Z)
    fc, fi = c.forecasts.to_numpy(), it.forecasts.to_numpy()
    assert np.isfinite(fc).all() and close(fc, fi, rtol=1e-9, atol=1e-12)
    assert close(c.rSquared, it.rSquared, rtol=1e-9)
    c2 = u.tprfClosedForm(y, X, Z)
    assert np.array_equal(fc, c2.forecasts.to_numpy()), "fixed reduction tree: bit-reproducible"
    # column statistics at full size against the axis reducers (independent kernels)
    sd = c.xstd.to_numpy()
    v = X.var(0).to_numpy() * (T / (T - 1.0))
    assert close(sd * sd, v, rtol=1e-11)


# ---- both sweeps: the single-read cluster sweep (tprf_fused.cuh) and the two-sweep path (tprf.cu) ----
FUSED_SHAPES = [(50, 6, 2), (130, 48, 1), (256, 1000, 4), (512, 2048, 4), (600, 130, 3), (1000, 512, 8),
                (2048, 256, 8), (4096, 320, 5), (8000, 200, 8), (8192, 1024, 8)]


@pytest.mark.parametrize("tnl", FUSED_SHAPES)
def test_single_read_sweep_vs_two_sweep_vs_oracle(u, tnl):
    """Every cluster size (1..16 CTAs), ragged tails in T (rows past T zero-filled by TMA, masked in the
    variance sums) and in N (columns past N), data with large column means (shifted sums)."""
    from uni_b200 import tprf3
    T, N, L = tnl
    rng = np.random.default_rng(T + N)
    X = rng.standard_normal((T, N)) * (1 + 3 * rng.random(N)) + 5 * rng.standard_normal(N)
    y = rng.standard_normal((T, 1)); Z = rng.standard_normal((T, L))
    o = to.tprf_closed_form_gram(y, X, Z)
    oi = to.t3prf(y, X, Z)
    res = {}
    try:
        for name, path in (("two", tprf3.TPRF_PATH_TWO_SWEEP), ("fused", tprf3.TPRF_PATH_FUSED), ("grid", tprf3.TPRF_PATH_GRID)):
            tprf3.set_tprf_path(path)
            c = u.tprfClosedForm(*devs(u, y, X, Z))
            it = u.t3prf(*devs(u, y, X, Z))
            res[name] = (c, it)
            assert close(c.forecasts.to_numpy(), o.forecasts), name
            assert close(c.alpha.to_numpy(), o.alpha, rtol=1e-8, atol=1e-13), name
            assert close(c.rSquared, o.r_squared), name
            assert close(it.forecasts.to_numpy(), oi.forecasts), name
            assert close(it.phi.to_numpy(), oi.phi, rtol=1e-8, atol=1e-13), name
            assert close(it.sigma.to_numpy(), oi.sigma, rtol=1e-7, atol=1e-12), name
            assert close(c.xstd.to_numpy().ravel(), X.std(axis=0, ddof=1), rtol=1e-11), name
    finally:
        tprf3.set_tprf_path(tprf3.TPRF_PATH_AUTO)
    assert close(res["fused"][0].forecasts.to_numpy(), res["two"][0].forecasts.to_numpy())
    assert close(res["grid"][0].forecasts.to_numpy(), res["two"][0].forecasts.to_numpy())


def test_single_read_sweep_requirements(u):
    """An odd row pitch cannot be described by a TMA tensor map: auto falls back to two sweeps, forcing
    the single-read sweep raises IllegalArgumentException's twin."""
    from uni_b200 import tprf3
    rng = np.random.default_rng(3)
    T, N, L = 64, 33, 2
    X = rng.standard_normal((T, N)); y = rng.standard_normal((T, 1)); Z = rng.standard_normal((T, L))
    o = to.tprf_closed_form_gram(y, X, Z)
    assert close(u.tprfClosedForm(*devs(u, y, X, Z)).forecasts.to_numpy(), o.forecasts)
    try:
        tprf3.set_tprf_path(tprf3.TPRF_PATH_FUSED)
        with pytest.raises(ValueError):
            u.tprfClosedForm(*devs(u, y, X, Z))
    finally:
        tprf3.set_tprf_path(tprf3.TPRF_PATH_AUTO)
    with pytest.raises(ValueError):
        tprf3.set_tprf_path(7)


def test_single_read_sweeps_bit_reproducible_and_match_two_sweep_at_scale(u):
    """T = 8192 (16-CTA clusters / groups), N = 65536: every path run-to-run bit-identical (auto = clusters +
    a concurrent grid sweep on the SMs they strand); all paths agree to 1e-9."""
    from uni_b200 import tprf3
    T, N, L = 8192, 65536, 8
    X = u.MatD.randn(T, N, seed=11); y = u.MatD.randn(T, 1, seed=12); Z = u.MatD.randn(T, L, seed=13)
    out = {}
    try:
        for path in (tprf3.TPRF_PATH_TWO_SWEEP, tprf3.TPRF_PATH_FUSED, tprf3.TPRF_PATH_GRID, tprf3.TPRF_PATH_AUTO):
            tprf3.set_tprf_path(path)
            a = u.tprfClosedForm(y, X, Z); b = u.tprfClosedForm(y, X, Z)
            assert np.array_equal(a.forecasts.to_numpy(), b.forecasts.to_numpy()), path
            assert np.array_equal(a.alpha.to_numpy(), b.alpha.to_numpy()), path
            out[path] = a
    finally:
        tprf3.set_tprf_path(tprf3.TPRF_PATH_AUTO)
    ref = out[tprf3.TPRF_PATH_TWO_SWEEP]
    for path, a in out.items():
        assert close(a.forecasts.to_numpy(), ref.forecasts.to_numpy()), path
        assert close(a.alpha.to_numpy(), ref.alpha.to_numpy(), rtol=1e-8, atol=1e-14), path
        assert close(a.xstd.to_numpy(), ref.xstd.to_numpy(), rtol=1e-12), path


def test_batched_config5_sample(u):
    """BASELINE config 5 shape (T=512, N=2048, L=4): 64 panels, every panel against the oracle
    (single-CTA 'clusters': no exchange between CTAs)."""
    from uni_b200 import _lib as L_
    B, T, N, L = 64, 512, 2048, 4
    Xs = np.empty((B, T, N)); ys = np.empty((B, T, 1)); Zs = np.empty((B, T, L))
    for p in range(B):
        g = np.random.default_rng(p)
        Xs[p] = g.standard_normal((T, N)); ys[p] = g.standard_normal((T, 1)); Zs[p] = g.standard_normal((T, L))
    dX = u.Mat.from_numpy(Xs.reshape(B * T, N)); dy = u.Mat.from_numpy(ys.reshape(B * T, 1)); dZ = u.Mat.from_numpy(Zs.reshape(B * T, L))
    f = u.MatD.zeros(B * T, 1); a = u.MatD.zeros(B * N, 1); r2 = u.MatD.zeros(B, 1)
    L_.check(L_.lib.um_tprf_fit_batched(L_.TPRF_CLOSED, dy._buf.ptr, dX._buf.ptr, dZ._buf.ptr, T, N, L, B, f._buf.ptr, a._buf.ptr, r2._buf.ptr))
    F = f.to_numpy().reshape(B, T); A = a.to_numpy().reshape(B, N); R = r2.to_numpy().ravel()
    for p in range(B):
        o = to.tprf_closed_form_gram(ys[p], Xs[p], Zs[p])
        assert close(F[p], o.forecasts.ravel()), p
        assert close(A[p], o.alpha.ravel(), rtol=1e-8, atol=1e-14), p
        assert close(R[p], o.r_squared), p


# ---------------------------------------------------------------- out-of-sample procedures
@pytest.mark.parametrize("case", ["T100N10L2", "T140N12L3", "T60N25L2"])
def test_scala_oos_rows(u, golden_dir, case):
    """`*/oosrec`, `*/ooscv01`, `*/ooscv23`: r2, enc, every forecast and rolling mean (apps/Tprf3ParityGen.scala:50-62,95-102)."""
    d = os.path.join(golden_dir, "tprf3-parity")
    ref = {}
    with open(os.path.join(d, "scala-reference.txt")) as fh:
        for line in fh:
            if line.startswith("#") or not line.strip():
                continue
            tag, key, val = line.split()
            ref[(tag, key)] = float(val)
    X = np.loadtxt(os.path.join(d, f"{case}_X.csv"), delimiter=",", ndmin=2)
    y = np.loadtxt(os.path.join(d, f"{case}_y.csv"), delimiter=",", ndmin=2).reshape(-1, 1)
    Z = np.loadtxt(os.path.join(d, f"{case}_Z.csv"), delimiter=",", ndmin=2)
    T = X.shape[0]
    dy, dX, dZ = devs(u, y, X, Z)
    for tag, proc, kw in (("oosrec", "OOS Recursive", {"mintrain": (T // 2, 0)}), ("ooscv01", "OOS Cross Val", {"window": (0, 1)}),
                          ("ooscv23", "OOS Cross Val", {"window": (2, 3)})):
        r = u.estimate3prf(dy, dX, dZ, proc, **kw)
        lab = f"{case}/{tag}"
        assert close(r.rSquared, ref[(lab, "r2")]), lab
        assert close(r.encnew, ref[(lab, "enc")]), lab
        assert close(r.forecasts.to_numpy()[:, 0], [ref[(lab, f"f{i}")] for i in range(T)]), lab
        assert close(r.rollfore.to_numpy()[:, 0], [ref[(lab, f"roll{i}")] for i in range(T)]), lab
        assert close(r.residuals.to_numpy(), y - r.forecasts.to_numpy())


def test_python_reference_oos_vectors(u, golden_dir):
    """Rolling windows, gaps and the automatic-proxy loops against the reference's NumPy implementation
    (tests/golden/gen_tprf_oos_pyref.py) and against the oracle."""
    from tests.test_oracle_tprf import oos_pyref_cases
    n = 0
    for name, y, X, Z, proc, kw, g in oos_pyref_cases(golden_dir):
        dy, dX = u.Mat.from_numpy(y), u.Mat.from_numpy(X)
        r = u.estimate3prf(dy, dX, Z if isinstance(Z, int) else u.Mat.from_numpy(Z), proc, **kw)
        want = g["forecasts"]
        got = r.forecasts.to_numpy()
        assert np.array_equal(np.isnan(got), np.isnan(want)), name
        # automatic proxies feed each fit's residuals into the next fit's Z: rounding differences compound
        tol = 1e-7 if isinstance(Z, int) else 1e-9
        assert close(got, want, rtol=tol, atol=tol), name
        assert close(r.rSquared, float(g["rsquare"]), rtol=tol, atol=tol), name
        o = to.estimate3prf(y, X, Z, proc, **kw)
        assert close(got, o.forecasts, rtol=tol, atol=tol), name
        if proc != "IS Full":
            assert close(r.rollfore.to_numpy(), o.extra["rollfore"]), name
        if proc == "OOS Recursive":
            assert close(r.encnew, float(g["enc"]), rtol=max(tol, 1e-8)), name
        n += 1
    assert n >= 10


def test_oos_windows_on_a_wide_panel(u):
    """A window count small enough to run here, a panel wide enough for the single-read sweeps (N = 4096)."""
    rng = np.random.default_rng(77)
    T, N, Lp = 72, 4096, 3
    X = rng.standard_normal((T, N)); y = rng.standard_normal((T, 1)); Z = rng.standard_normal((T, Lp))
    dy, dX, dZ = devs(u, y, X, Z)
    r = u.estimate3prf(dy, dX, dZ, "OOS Recursive", mintrain=(60, 0))
    o = to.estimate3prf(y, X, Z, "OOS Recursive", mintrain=(60, 0))
    assert close(r.forecasts.to_numpy(), o.forecasts) and close(r.rSquared, o.r_squared) and close(r.encnew, o.extra["enc"], rtol=1e-8)
    r = u.estimate3prf(dy, dX, dZ, "OOS Rolling", rollwin=(40, 20, 1))
    o = to.estimate3prf(y, X, Z, "OOS Rolling", rollwin=(40, 20, 1))
    assert close(r.forecasts.to_numpy(), o.forecasts) and close(r.rSquared, o.r_squared)
    # NaN input without pls (Tprf3.scala:883-943): a NaN in X spreads through every window that holds its row and
    # through the forecast of its own row; windows clear of it still forecast
    Xn = X.copy(); Xn[3, 5] = np.nan
    r = u.estimate3prf(dy, u.Mat.from_numpy(Xn), dZ, "OOS Rolling", rollwin=(40, 20, 1))
    o = to.estimate3prf(y, Xn, Z, "OOS Rolling", rollwin=(40, 20, 1))
    f = r.forecasts.to_numpy()
    assert np.array_equal(np.isnan(f), np.isnan(o.forecasts)) and np.isfinite(f).any() and np.isnan(f[42:45]).all()
    assert close(f, o.forecasts) and close(r.rSquared, o.r_squared)
    # a NaN in y only takes its row out of pass 3 (nanOls, :929-931) and out of R^2 (`loc`, :1207)
    yn = y.copy(); yn[50, 0] = np.nan
    r = u.estimate3prf(u.Mat.from_numpy(yn), dX, dZ, "OOS Recursive", mintrain=(60, 0))
    o = to.estimate3prf(yn, X, Z, "OOS Recursive", mintrain=(60, 0))
    assert close(r.forecasts.to_numpy(), o.forecasts) and close(r.rSquared, o.r_squared)
    assert np.isfinite(r.forecasts.to_numpy()[61:]).all()


def test_nan_input_is_full(u):
    """estimate3prf "IS Full" with NaN input and no pls: X or Z -> every forecast NaN, R^2 NaN; y -> that row dropped
    from pass 3 and from R^2, the other rows forecast (Tprf3.scala:883-943, 1206-1225)."""
    rng = np.random.default_rng(5)
    T, N, Lp = 64, 300, 3
    X = rng.standard_normal((T, N)); y = rng.standard_normal((T, 1)) + X[:, :1]; Z = rng.standard_normal((T, Lp)) + X[:, :Lp]
    Xn = X.copy(); Xn[10, 7] = np.nan
    r = u.estimate3prf(*devs(u, y, Xn, Z))
    assert np.isnan(r.forecasts.to_numpy()).all() and np.isnan(r.rSquared) and not r.singular
    Zn = Z.copy(); Zn[5, 1] = np.nan
    r = u.estimate3prf(*devs(u, y, X, Zn))
    assert np.isnan(r.forecasts.to_numpy()).all() and np.isnan(r.rSquared)
    yn = y.copy(); yn[20, 0] = np.nan
    r = u.estimate3prf(*devs(u, yn, X, Z))
    o = to.estimate3prf(yn, X, Z)
    f = r.forecasts.to_numpy()
    assert np.isfinite(f).all() and close(f, o.forecasts) and close(r.rSquared, o.r_squared)
    assert np.isnan(r.residuals.to_numpy()[20, 0]) and np.isnan(r.alpha.to_numpy()).all()   # alpha = ... (XtJt y): NaN (:1245-1248)
    # t3prf / tprfClosedForm take the plain products: a NaN anywhere poisons them (Tprf3.scala:229-231,268-269)
    assert np.isnan(u.t3prf(*devs(u, yn, X, Z)).forecasts.to_numpy()).all()
    assert np.isnan(u.tprfClosedForm(*devs(u, yn, X, Z)).forecasts.to_numpy()).all()


def test_singular_gram_raises(u):
    """An exact-zero pivot in a fit's small inverse is the reference's ArithmeticException("Matrix is singular or nearly
    singular") (Mat.scala:990): duplicated / constant proxy columns, in every fit and in an out-of-sample window."""
    rng = np.random.default_rng(6)
    T, N = 48, 200
    X = rng.standard_normal((T, N)); y = rng.standard_normal((T, 1))
    z = np.round(rng.standard_normal((T, 1)) * 8) / 8          # exactly representable: the duplicate cancels exactly
    Zdup = np.column_stack([z, z])
    for fit in (u.tprfClosedForm, u.t3prf, u.estimate3prf):
        with pytest.raises(u.SingularMatrixError):
            fit(*devs(u, y, X, Zdup))
    Zc = np.column_stack([z, np.zeros(T)])                      # a constant (zero) proxy: dZ'dZ has a zero pivot
    with pytest.raises(u.SingularMatrixError):
        u.t3prf(*devs(u, y, X, Zc))
    with pytest.raises(u.SingularMatrixError):
        u.estimate3prf(*devs(u, y, X, Zdup), "OOS Recursive", mintrain=(40, 0))
    ok = u.t3prf(*devs(u, y, X, np.column_stack([z, rng.standard_normal(T)])))
    assert not ok.singular and np.isfinite(ok.forecasts.to_numpy()).all()


# ---------------------------------------------------------------- PLS variant
@pytest.mark.parametrize("case", ["T100N10L2", "T140N12L3", "T60N25L2"])
def test_scala_pls_rows(u, golden_dir, case):
    """`*/pls`: r2, b0, b1, f<i>, pred0 (apps/Tprf3ParityGen.scala:69-79,104)."""
    d = os.path.join(golden_dir, "tprf3-parity")
    ref = {}
    with open(os.path.join(d, "scala-reference.txt")) as fh:
        for line in fh:
            if line.startswith("#") or not line.strip():
                continue
            tag, key, val = line.split()
            ref[(tag, key)] = float(val)
    X = np.loadtxt(os.path.join(d, f"{case}_X.csv"), delimiter=",", ndmin=2)
    y = np.loadtxt(os.path.join(d, f"{case}_y.csv"), delimiter=",", ndmin=2).reshape(-1, 1)
    T = X.shape[0]
    m = u.plsClosedForm(u.Mat.from_numpy(y), u.Mat.from_numpy(X))
    lab = f"{case}/pls"
    assert close(m.rSquared, ref[(lab, "r2")])
    assert close(m.beta.to_numpy()[:, 0], [ref[(lab, "b0")], ref[(lab, "b1")]])
    assert close(m.forecasts.to_numpy()[:, 0], [ref[(lab, f"f{i}")] for i in range(T)])
    assert close(m.predict(X[0]), ref[(lab, "pred0")])
    o = to.pls_closed_form(y, X)
    assert close(m.phi.to_numpy(), o.phi) and close(m.sigma.to_numpy(), o.sigma)
    assert close(m.colMean.to_numpy(), o.col_mean, atol=1e-15) and close(m.colStd.to_numpy(), o.col_std)
    assert close(m.predictAll(X[:7]).to_numpy()[:, 0], [o.predict(X[i]) for i in range(7)])
    m2 = u.pls1Fit(X.tolist(), y[:, 0].tolist())
    assert np.array_equal(m2.forecasts.to_numpy(), m.forecasts.to_numpy())
    with pytest.raises(ValueError):
        m.predict(X[0, :-1])
    with pytest.raises(ValueError):
        u.plsClosedForm(u.Mat.from_numpy(y[:-1]), u.Mat.from_numpy(X))


def test_pls_wide_panel_vs_oracle(u):
    rng = np.random.default_rng(12)
    T, N = 512, 4096
    X = rng.standard_normal((T, N)) * rng.uniform(0.5, 3.0, size=(1, N)) + rng.uniform(-2, 2, size=(1, N))
    y = (X[:, :5] @ rng.standard_normal((5, 1))) + rng.standard_normal((T, 1)) + 0.7
    m = u.plsClosedForm(u.Mat.from_numpy(y), u.Mat.from_numpy(X))
    o = to.pls_closed_form(y, X)
    assert close(m.forecasts.to_numpy(), o.forecasts) and close(m.rSquared, o.r_squared)
    assert close(m.beta.to_numpy(), o.beta) and close(m.phi.to_numpy(), o.phi)
    assert close(m.predict(X[11] * 1.01), o.predict(X[11] * 1.01))


# ---------------------------------------------------------------- multi-GPU (runs when the box has >= 2 GPUs)
def test_estimate3prf_pls_gappy_panels_vs_numpy_reference_and_oracle(u, golden_dir):
    """estimate3prf(pls = true, automatic proxies): the per-column / per-row nanOls passes (Tprf3.scala:985-1048) as
    masked batched problems on the device, NaN gaps included, against the reference's NumPy implementation
    (tests/golden/tprf_pls_pyref, generator beside it) and the oracle, all four procedures."""
    import glob
    from tests.test_oracle_tprf import pls_kwargs, pls_panel
    files = sorted(glob.glob(os.path.join(golden_dir, "tprf_pls_pyref", "*.npz")))
    assert len(files) == 10
    for f in files:
        g = np.load(f)
        y, X = pls_panel(int(g["seed"]), int(g["T"]), int(g["N"]), bool(g["gaps"]))
        proc, kw = str(g["procedure"]), pls_kwargs(g)
        r = u.estimate3prf(u.Mat.from_numpy(y), u.Mat.from_numpy(X), int(g["L"]), proc, pls=True, **kw)
        got, want = r.forecasts.to_numpy(), g["forecasts"]
        assert np.array_equal(np.isnan(got), np.isnan(want)), f
        ok = ~np.isnan(want)
        assert np.allclose(got[ok], want[ok], rtol=1e-9, atol=1e-12), (f, np.max(np.abs(got[ok] - want[ok])))
        o = to.estimate3prf(y, X, int(g["L"]), proc, pls=True, **kw)
        assert np.isclose(r.rSquared, o.r_squared, rtol=1e-9, atol=1e-12, equal_nan=True), f
    # the same gappy panel without pls: the matrix passes spread the NaNs to every forecast (Tprf3.scala:883-943)
    y, X = pls_panel(43, 90, 15, True)
    r = u.estimate3prf(u.Mat.from_numpy(y), u.Mat.from_numpy(X), 2, "IS Full")
    o = to.estimate3prf(y, X, 2, "IS Full")
    assert np.isnan(r.forecasts.to_numpy()).all() and np.isnan(o.forecasts).all() and np.isnan(r.rSquared)


def test_pls_closed_form_agrees_with_estimate3prf_pls(u):
    """Pls3prfSuite.scala:94-110: the closed form and the iterative pls = true fit give the same forecasts (L = 1)."""
    for T, N, seed in ((60, 8, 1), (50, 3, 2), (200, 40, 3)):
        rng = np.random.default_rng(seed)
        X = rng.standard_normal((T, N)); y = rng.standard_normal((T, 1)) + X[:, :1]
        dy, dX = u.Mat.from_numpy(y), u.Mat.from_numpy(X)
        it = u.estimate3prf(dy, dX, 1, pls=True)
        cf = u.plsClosedForm(dy, dX)
        assert np.allclose(it.forecasts.to_numpy(), cf.forecasts.to_numpy(), rtol=1e-9, atol=1e-10)


def test_sharded_fit_across_ranks(u):
    """tools/dist_check.py under torchrun, one rank per GPU: sharded closed-form / three-pass fits vs the oracle on the
    full panel, bit-identical forecasts on every rank, the peer-memory exchange against a host-side sum."""
    import subprocess
    import sys
    n = u.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs (the 1-GPU path is covered above; host-side sharding logic in tests/test_dist_cpu.py)")
    world = 2 if n < 4 else 4
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for env_p2p in ("1", "0"):
        env = dict(os.environ, UM_COMM_P2P=env_p2p)
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                            "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(root, "tools", "dist_check.py")],
                           cwd=root, env=env, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        assert r.stdout.count("DIST_CHECK PASS") == world
        assert ("peer memory" in r.stdout) == (env_p2p == "1") or "NCCL" in r.stdout
