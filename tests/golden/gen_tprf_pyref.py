#!/usr/bin/env python3
"""Generate golden 3PRF vectors by running the REFERENCE's own NumPy implementation.

Run in the build container only (needs /root/reference):

    python tests/golden/gen_tprf_pyref.py

Imports `/root/reference/py/tprf3fast.py` (`estimate3prf_fast`, IS Full) and
`/root/reference/py/tprf3.py` (`estimate3prf`, loop form) unchanged, runs them on seeded
synthetic panels (numpy.random.default_rng(seed), draw order X, y, Z as
apps/Tprf3Bench.scala:146-149) and stores forecasts / alpha / R^2 under
tests/golden/tprf_pyref/.  Inputs are NOT stored: the tests regenerate them from the seed.
"""
import os
import sys

sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference/py")
import numpy as np  # noqa: E402
import tprf3  # noqa: E402
import tprf3fast  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = [  # (name, seed, T, N, L)
    ("T64N16L2", 11, 64, 16, 2),
    ("T200N30L2", 0, 200, 30, 2),
    ("T650N40L2", 0, 650, 40, 2),
    ("T256N512L4", 5, 256, 512, 4),
    ("T512N2048L4", 7, 512, 2048, 4),
    ("T300N96L8", 3, 300, 96, 8),
]


def panel(seed, T, N, L):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((T, N))
    y = rng.standard_normal((T, 1))
    Z = rng.standard_normal((T, L))
    return y, X, Z


def main():
    out = os.path.join(HERE, "tprf_pyref")
    os.makedirs(out, exist_ok=True)
    for name, seed, T, N, L in CASES:
        y, X, Z = panel(seed, T, N, L)
        f, pe, _ = tprf3fast.estimate3prf_fast(y, X, Z, procedure="IS Full")
        d = dict(seed=seed, T=T, N=N, L=L, forecasts=f, alpha=pe.alpha, rsquare=pe.rsquare)
        if N <= 64:
            f2, pe2, _ = tprf3.estimate3prf(y, X, Z, procedure="IS Full")
            d["loop_forecasts"] = f2
            d["loop_rsquare"] = pe2.rsquare
        np.savez(os.path.join(out, name + ".npz"), **d)
        print(name, "r2", pe.rsquare)


if __name__ == "__main__":
    main()
