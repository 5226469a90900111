#!/usr/bin/env python3
"""Golden vectors for estimate3prf(..., "IS Full", computeAvar = true) from the REFERENCE's own NumPy implementation.

Run in the build container only (needs /root/reference):

    python tests/golden/gen_tprf_avar_pyref.py

Imports `/root/reference/py/tprf3fast.py` unchanged and runs `estimate3prf_fast(..., compute_avar=True)` on seeded
panels (numpy.random.default_rng(seed), draw order X, y, Z as apps/Tprf3Bench.scala:146-149), with a proxy matrix and
with automatic proxies; stores the N x N `avar.alpha`, the T x T `avar.forecasts`, the forecasts and R^2 under
tests/golden/tprf_avar_pyref/.  Inputs are NOT stored: the tests regenerate them from the seed.
"""
import os
import sys

sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference/py")
import numpy as np  # noqa: E402
import tprf3fast  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = [  # (name, seed, T, N, L, automatic proxies?)
    ("T40N8L2", 21, 40, 8, 2, False),
    ("T64N16L3", 22, 64, 16, 3, False),
    ("T120N30L2", 23, 120, 30, 2, False),
    ("T64N16A1", 24, 64, 16, 1, True),
    ("T90N24A2", 25, 90, 24, 2, True),
]


def panel(seed, T, N, L):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((T, N))
    y = rng.standard_normal((T, 1))
    Z = rng.standard_normal((T, L))
    return y, X, Z


def main():
    out = os.path.join(HERE, "tprf_avar_pyref")
    os.makedirs(out, exist_ok=True)
    for name, seed, T, N, L, auto in CASES:
        y, X, Z = panel(seed, T, N, L)
        f, pe, av = tprf3fast.estimate3prf_fast(y, X, L if auto else Z, procedure="IS Full", compute_avar=True)
        assert av is not None
        np.savez(os.path.join(out, name + ".npz"), seed=seed, T=T, N=N, L=L, auto=auto, forecasts=f, rsquare=pe.rsquare,
                 avar_alpha=av.alpha, avar_forecasts=av.forecasts)
        print(name, "r2", pe.rsquare, "avar alpha", av.alpha.shape, float(np.abs(av.alpha).max()))


if __name__ == "__main__":
    main()
