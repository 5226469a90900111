"""Pins oracle/tprf_oracle.py against the reference's golden vectors (CPU only)."""
import os

import numpy as np
import pytest

from oracle import tprf_oracle as to

CASES = ["T100N10L2", "T140N12L3", "T60N25L2"]


def _close(a, b):
    """src/test/scala/uni/stats/Tprf3ParitySuite.scala:28-45"""
    if np.isnan(a) and np.isnan(b):
        return True
    return abs(a - b) <= 1e-12 + 1e-9 * max(abs(a), abs(b))


def _load_scala_ref(golden_dir):
    ref = {}
    with open(os.path.join(golden_dir, "tprf3-parity", "scala-reference.txt")) as fh:
        for line in fh:
            if line.startswith("#") or not line.strip():
                continue
            tag, key, val = line.split()
            ref[(tag, key)] = float(val)
    return ref


def _panel(golden_dir, case):
    d = os.path.join(golden_dir, "tprf3-parity")
    X = np.loadtxt(os.path.join(d, f"{case}_X.csv"), delimiter=",", ndmin=2)
    y = np.loadtxt(os.path.join(d, f"{case}_y.csv"), delimiter=",", ndmin=2).reshape(-1, 1)
    Z = np.loadtxt(os.path.join(d, f"{case}_Z.csv"), delimiter=",", ndmin=2)
    return y, X, Z


def test_kp_known_answers(golden_dir):
    d = os.path.join(golden_dir, "t3prf-validation")
    X = np.loadtxt(os.path.join(d, "ref_X.csv"), delimiter=",")
    Z = np.loadtxt(os.path.join(d, "ref_Z.csv"), delimiter=",")
    y = np.loadtxt(os.path.join(d, "ref_y.csv"), delimiter=",").reshape(-1, 1)
    kp_yhat = np.loadtxt(os.path.join(d, "kp_yhat.csv"), delimiter=",").reshape(-1, 1)
    kp_alpha = np.loadtxt(os.path.join(d, "kp_alpha.csv"), delimiter=",").reshape(-1, 1)
    kp_r2 = float(np.loadtxt(os.path.join(d, "kp_rsquare.csv"), delimiter=","))
    # fixture inputs are numpy default_rng(1234): Z, X, y (jsrc/generate3prfData.sc:21-25)
    rng = np.random.default_rng(1234)
    assert np.array_equal(rng.standard_normal((50, 2)), Z)
    assert np.array_equal(rng.standard_normal((50, 6)), X)
    for fit in (to.estimate3prf_is_full, to.t3prf, to.tprf_closed_form_literal, to.tprf_closed_form_gram):
        r = fit(y, X, Z)
        assert np.max(np.abs(r.forecasts - kp_yhat)) < 1e-10, fit.__name__
        assert abs(r.r_squared - kp_r2) < 1e-10, fit.__name__
        if r.alpha is not None:
            assert np.max(np.abs(r.alpha - kp_alpha)) < 1e-10, fit.__name__


@pytest.mark.parametrize("case", CASES)
def test_scala_parity_rows(golden_dir, case):
    ref = _load_scala_ref(golden_dir)
    y, X, Z = _panel(golden_dir, case)
    T = X.shape[0]
    isf = to.estimate3prf_is_full(y, X, Z)
    assert _close(isf.r_squared, ref[(f"{case}/isfull", "r2")])
    for i in range(T):
        assert _close(float(isf.forecasts[i, 0]), ref[(f"{case}/isfull", f"f{i}")]), i
    for fit in (to.tprf_closed_form_literal, to.tprf_closed_form_gram, to.t3prf):
        r = fit(y, X, Z)
        assert _close(r.r_squared, ref[(f"{case}/closed", "r2")]), fit.__name__
        for i in range(T):
            assert _close(float(r.forecasts[i, 0]), ref[(f"{case}/closed", f"f{i}")]), (fit.__name__, i)


def test_python_reference_vectors(golden_dir):
    d = os.path.join(golden_dir, "tprf_pyref")
    names = sorted(f for f in os.listdir(d) if f.endswith(".npz"))
    assert names
    for f in names:
        g = np.load(os.path.join(d, f))
        rng = np.random.default_rng(int(g["seed"]))
        T, N, L = int(g["T"]), int(g["N"]), int(g["L"])
        X = rng.standard_normal((T, N)); y = rng.standard_normal((T, 1)); Z = rng.standard_normal((T, L))
        for fit in (to.estimate3prf_is_full, to.tprf_closed_form_gram):
            r = fit(y, X, Z)
            np.testing.assert_allclose(r.forecasts, g["forecasts"], rtol=1e-9, atol=1e-12, err_msg=f)
            np.testing.assert_allclose(r.alpha, g["alpha"], rtol=1e-8, atol=1e-12, err_msg=f)
            assert abs(r.r_squared - float(g["rsquare"])) < 1e-10


def test_python_reference_avar_vectors(golden_dir):
    """estimate3prf("IS Full", computeAvar = true) (Tprf3.scala:1252-1277) against outputs of the reference's own NumPy
    implementation (tests/golden/gen_tprf_avar_pyref.py): proxy matrix and automatic proxies."""
    d = os.path.join(golden_dir, "tprf_avar_pyref")
    names = sorted(f for f in os.listdir(d) if f.endswith(".npz"))
    assert len(names) >= 5
    for f in names:
        g = np.load(os.path.join(d, f))
        rng = np.random.default_rng(int(g["seed"]))
        T, N, L = int(g["T"]), int(g["N"]), int(g["L"])
        X = rng.standard_normal((T, N)); y = rng.standard_normal((T, 1)); Z = rng.standard_normal((T, L))
        r = to.estimate3prf(y, X, L if bool(g["auto"]) else Z, "IS Full", compute_avar=True)
        np.testing.assert_allclose(r.forecasts, g["forecasts"], rtol=1e-9, atol=1e-12, err_msg=f)
        assert r.extra["avar_alpha"].shape == (N, N) and r.extra["avar_forecasts"].shape == (T, T)
        np.testing.assert_allclose(r.extra["avar_alpha"], g["avar_alpha"], rtol=1e-9, atol=1e-12, err_msg=f)
        np.testing.assert_allclose(r.extra["avar_forecasts"], g["avar_forecasts"], rtol=1e-9, atol=1e-12, err_msg=f)
        # both are symmetric positive semi-definite by construction
        assert np.allclose(r.extra["avar_alpha"], r.extra["avar_alpha"].T, rtol=1e-9, atol=1e-12)
    # computeAvar is ignored outside "IS Full" (:1253)
    assert "avar_alpha" not in to.estimate3prf(y, X, Z, "OOS Recursive", compute_avar=True).extra


def test_sharded_partials_sum_to_full_fit():
    rng = np.random.default_rng(3)
    T, N, L = 120, 48, 3
    X = rng.standard_normal((T, N)) * 3 + 1.5; y = rng.standard_normal((T, 1)); Z = rng.standard_normal((T, L))
    full = to.tprf_closed_form_gram(y, X, Z)
    parts = [to.closed_form_partials(X[:, a:b], Z) for a, b in ((0, 10), (10, 31), (31, 48))]
    tot = {k: sum(p[k] for p in parts) for k in ("PX", "h0", "sw", "SWW", "smw", "smu")}
    yhat, s, wbar = to.closed_form_from_partials(y, tot, N)
    np.testing.assert_allclose(yhat, full.forecasts, rtol=1e-10, atol=1e-12)
    alpha = np.concatenate([(p["W0"] - wbar) @ s for p in parts])
    np.testing.assert_allclose(alpha, full.alpha, rtol=1e-9, atol=1e-13)


def test_edge_semantics():
    """TprfCoverageSuite.scala:176-230: constant column -> sd 1.0; T < minObs -> NaN."""
    X = np.ones((12, 3)); X[:, 1] = np.arange(12.0)
    sd = to.nan_std_cols(X)
    assert sd[0, 0] == 1.0 and sd[0, 2] == 1.0 and sd[0, 1] > 0
    Xn = np.array([[1.0, np.nan], [3.0, np.nan]])
    sd = to.nan_std_cols(Xn)
    assert sd[0, 1] == 1.0 and abs(sd[0, 0] - np.sqrt(2.0)) < 1e-15
    rng = np.random.default_rng(0)
    r = to.estimate3prf_is_full(rng.standard_normal((5, 1)), rng.standard_normal((5, 4)), rng.standard_normal((5, 1)))
    assert np.isnan(r.forecasts).all() and np.isnan(r.r_squared)


# ---------------------------------------------------------------- out-of-sample procedures
_SCALA_OOS = [("oosrec", "OOS Recursive", {}), ("ooscv01", "OOS Cross Val", {"window": (0, 1)}),
              ("ooscv23", "OOS Cross Val", {"window": (2, 3)})]


@pytest.mark.parametrize("case", CASES)
def test_scala_oos_rows(golden_dir, case):
    """Every `*/oosrec`, `*/ooscv01`, `*/ooscv23` row: r2, enc, f<i>, roll<i> (apps/Tprf3ParityGen.scala:50-62,95-102)."""
    ref = _load_scala_ref(golden_dir)
    y, X, Z = _panel(golden_dir, case)
    T = X.shape[0]
    for tag, proc, kw in _SCALA_OOS:
        if proc == "OOS Recursive":
            kw = {"mintrain": (T // 2, 0)}
        r = to.estimate3prf(y, X, Z, proc, **kw)
        lab = f"{case}/{tag}"
        assert _close(r.r_squared, ref[(lab, "r2")]), lab
        assert _close(r.extra["enc"], ref[(lab, "enc")]), lab
        for i in range(T):
            assert _close(float(r.forecasts[i, 0]), ref[(lab, f"f{i}")]), (lab, i)
            assert _close(float(r.extra["rollfore"][i, 0]), ref[(lab, f"roll{i}")]), (lab, i)
    isf = to.estimate3prf(y, X, Z, "IS Full")
    assert _close(isf.r_squared, ref[(f"{case}/isfull", "r2")])
    assert all(_close(float(isf.forecasts[i, 0]), ref[(f"{case}/isfull", f"f{i}")]) for i in range(T))


def oos_pyref_cases(golden_dir):
    d = os.path.join(golden_dir, "tprf_oos_pyref")
    for f in sorted(x for x in os.listdir(d) if x.endswith(".npz")):
        g = np.load(os.path.join(d, f))
        rng = np.random.default_rng(int(g["seed"]))
        T, N, L = int(g["T"]), int(g["N"]), int(g["L"])
        X = rng.standard_normal((T, N)); y = rng.standard_normal((T, 1)); Z = rng.standard_normal((T, L))
        kw = {k[3:]: tuple(int(v) for v in g[k]) for k in g.files if k.startswith("kw_")}
        yield f[:-4], y, X, (L if bool(g["auto"]) else Z), str(g["procedure"]), kw, g


def test_python_reference_oos_vectors(golden_dir):
    """Rolling windows and the automatic-proxy loops, which the Scala fixture does not cover: outputs of the reference's
    NumPy implementation (tests/golden/gen_tprf_oos_pyref.py)."""
    n = 0
    for name, y, X, Z, proc, kw, g in oos_pyref_cases(golden_dir):
        r = to.estimate3prf(y, X, Z, proc, **kw)
        want = g["forecasts"]
        assert np.array_equal(np.isnan(r.forecasts), np.isnan(want)), name
        ok = ~np.isnan(want)
        assert np.max(np.abs(r.forecasts[ok] - want[ok]) / np.maximum(1.0, np.abs(want[ok]))) < 1e-9, name
        assert _close(r.r_squared, float(g["rsquare"])) or abs(r.r_squared - float(g["rsquare"])) < 1e-9, name
        if proc == "OOS Recursive":
            assert abs(r.extra["enc"] - float(g["enc"])) <= 1e-8 * max(1.0, abs(float(g["enc"]))), name
        n += 1
    assert n >= 10


def test_estimate3prf_unknown_procedure():
    y, X, Z = np.zeros((20, 1)), np.ones((20, 3)), np.ones((20, 1))
    with pytest.raises(ValueError):
        to.estimate3prf(y, X, Z, "OOS Bootstrap")      # Tprf3.scala:1201-1204


# ---------------------------------------------------------------- PLS variant
@pytest.mark.parametrize("case", CASES)
def test_scala_pls_rows(golden_dir, case):
    """`*/pls`: r2, b0, b1, every forecast and the raw-row prediction pred0 (apps/Tprf3ParityGen.scala:69-79,104)."""
    ref = _load_scala_ref(golden_dir)
    y, X, _ = _panel(golden_dir, case)
    m = to.pls_closed_form(y, X)
    lab = f"{case}/pls"
    assert _close(m.r_squared, ref[(lab, "r2")])
    assert _close(float(m.beta[0, 0]), ref[(lab, "b0")]) and _close(float(m.beta[1, 0]), ref[(lab, "b1")])
    for i in range(X.shape[0]):
        assert _close(float(m.forecasts[i, 0]), ref[(lab, f"f{i}")]), i
    assert _close(m.predict(X[0]), ref[(lab, "pred0")])
    # predicting a training row reproduces its fitted value
    assert abs(m.predict(X[3]) - float(m.forecasts[3, 0])) < 1e-10
    with pytest.raises(ValueError):
        m.predict(X[0, :-1])
    with pytest.raises(ValueError):
        to.pls_closed_form(y[:-1], X)


# ---- estimate3prf(pls = True): nanOls passes, gappy panels -- pinned to the reference's NumPy implementation ----
def _pls_cases(golden_dir):
    import glob
    return sorted(glob.glob(os.path.join(golden_dir, "tprf_pls_pyref", "*.npz")))


def pls_panel(seed, T, N, gaps):
    """tests/golden/gen_tprf_pls_pyref.py::panel"""
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((T, N))
    y = rng.standard_normal((T, 1)) + 0.5 * X[:, :1]
    if gaps:
        X[rng.random((T, N)) < 0.04] = np.nan
        y[rng.choice(T, 3, replace=False), 0] = np.nan
    return y, X


def pls_kwargs(g):
    return {k[3:]: tuple(int(v) for v in g[k]) for k in g.files if k.startswith("kw_")}


def test_oracle_pls_estimate_matches_numpy_reference(golden_dir):
    files = _pls_cases(golden_dir)
    assert len(files) == 10
    for f in files:
        g = np.load(f)
        y, X = pls_panel(int(g["seed"]), int(g["T"]), int(g["N"]), bool(g["gaps"]))
        r = to.estimate3prf(y, X, int(g["L"]), str(g["procedure"]), pls=True, **pls_kwargs(g))
        want = g["forecasts"]
        assert np.array_equal(np.isnan(r.forecasts), np.isnan(want)), f
        ok = ~np.isnan(want)
        assert np.allclose(r.forecasts[ok], want[ok], rtol=1e-9, atol=1e-12), f
        if str(g["procedure"]) != "IS Full":
            rf = g["rollfore"]
            assert np.allclose(r.extra["rollfore"][~np.isnan(rf)], rf[~np.isnan(rf)], rtol=1e-12, atol=0), f
