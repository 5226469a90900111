#!/usr/bin/env python3
"""Golden vectors for estimate3prf(pls = True) -- per-column / per-row nanOls passes, gappy (NaN) panels -- from the
REFERENCE's own NumPy implementation (build container only; needs /root/reference):

    python tests/golden/gen_tprf_pls_pyref.py

Imports `/root/reference/py/tprf3.py::estimate3prf` unchanged (the loop / lstsq twin of Tprf3.scala's runT3prf pls
branch, :985-1048) and runs the four procedures with automatic proxies on seeded panels
(numpy.random.default_rng(seed), draw order X, y as apps/Tprf3Bench.scala:146-149), without and with NaN gaps
(a seeded 4 % of X and a few y entries).  Stored under tests/golden/tprf_pls_pyref/; inputs are regenerated from
the seed by `panel` (copied by the tests)."""
import os
import sys

sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference/py")
import numpy as np  # noqa: E402
import tprf3  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
# (name, seed, T, N, L, procedure, kwargs, gaps)
CASES = [
    ("isfull_T80N12L1", 41, 80, 12, 1, "IS Full", dict(), False),
    ("isfull_T80N12L2", 42, 80, 12, 2, "IS Full", dict(), False),
    ("isfull_gaps_T90N15L2", 43, 90, 15, 2, "IS Full", dict(), True),
    ("isfull_smallN_T60N3L1", 44, 60, 3, 1, "IS Full", dict(), False),
    ("rec_T60N10L1", 45, 60, 10, 1, "OOS Recursive", dict(mintrain=(30, 0)), False),
    ("rec_gaps_T60N10L2", 46, 60, 10, 2, "OOS Recursive", dict(mintrain=(30, 1)), True),
    ("cv_T40N8L1", 47, 40, 8, 1, "OOS Cross Val", dict(window=(1, 3)), False),
    ("cv_gaps_T40N8L2", 48, 40, 8, 2, "OOS Cross Val", dict(window=(0, 1)), True),
    ("roll_T70N9L1", 49, 70, 9, 1, "OOS Rolling", dict(rollwin=(30, 20, 0)), False),
    ("roll_gaps_T70N9L2", 50, 70, 9, 2, "OOS Rolling", dict(rollwin=(30, 20, 2)), True),
]


def panel(seed, T, N, gaps):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((T, N))
    y = rng.standard_normal((T, 1)) + 0.5 * X[:, :1]
    if gaps:
        X[rng.random((T, N)) < 0.04] = np.nan
        y[rng.choice(T, 3, replace=False), 0] = np.nan
    return y, X


def main():
    out = os.path.join(HERE, "tprf_pls_pyref")
    os.makedirs(out, exist_ok=True)
    for name, seed, T, N, L, proc, kw, gaps in CASES:
        y, X = panel(seed, T, N, gaps)
        f, pe, _ = tprf3.estimate3prf(y, X, L, procedure=proc, pls=True, **kw)
        np.savez(os.path.join(out, name + ".npz"), seed=seed, T=T, N=N, L=L, procedure=proc, gaps=gaps,
                 forecasts=f, rollfore=pe.rollfore, **{"kw_" + k: np.array(v) for k, v in kw.items()})
        print(name, "nan forecasts", int(np.isnan(f).sum()), "of", T)


if __name__ == "__main__":
    main()
