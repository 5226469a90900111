#!/usr/bin/env python3
"""Generate golden vectors for estimate3prf's OUT-OF-SAMPLE procedures by running the REFERENCE's own NumPy
implementation (build container only; needs /root/reference):

    python tests/golden/gen_tprf_oos_pyref.py

Imports `/root/reference/py/tprf3fast.py::estimate3prf_fast` unchanged and runs OOS Recursive / Cross Val /
Rolling with a proxy matrix and with automatic proxies (Z = L) on seeded synthetic panels
(numpy.random.default_rng(seed), draw order X, y, Z as apps/Tprf3Bench.scala:146-149).  The Scala fixture
(test-data/tprf3-parity) pins Recursive and Cross Val with a proxy matrix; these add Rolling and the
automatic-proxy loops.  Stored under tests/golden/tprf_oos_pyref/; inputs are regenerated from the seed.
"""
import os
import sys

sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference/py")
import numpy as np  # noqa: E402
import tprf3fast  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
# (name, seed, T, N, L, procedure, kwargs, autoproxy)
CASES = [
    ("rec_T90N14L2", 21, 90, 14, 2, "OOS Recursive", dict(mintrain=(45, 0)), False),
    ("rec_gap_T80N10L2", 22, 80, 10, 2, "OOS Recursive", dict(mintrain=(40, 2)), False),
    ("cv01_T70N9L3", 23, 70, 9, 3, "OOS Cross Val", dict(window=(0, 1)), False),
    ("cv25_T70N9L3", 23, 70, 9, 3, "OOS Cross Val", dict(window=(2, 5)), False),
    ("roll_T120N16L2", 24, 120, 16, 2, "OOS Rolling", dict(rollwin=(30, 20, 0)), False),
    ("roll_gap_T100N8L1", 25, 100, 8, 1, "OOS Rolling", dict(rollwin=(40, 20, 3)), False),
    ("isfull_auto_T80N12L2", 26, 80, 12, 2, "IS Full", dict(), True),
    ("rec_auto_T70N10L2", 27, 70, 10, 2, "OOS Recursive", dict(mintrain=(35, 0)), True),
    ("cv_auto_T60N8L2", 28, 60, 8, 2, "OOS Cross Val", dict(window=(1, 3)), True),
    ("roll_auto_T90N10L2", 29, 90, 10, 2, "OOS Rolling", dict(rollwin=(30, 20, 0)), True),
]


def panel(seed, T, N, L):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((T, N))
    y = rng.standard_normal((T, 1))
    Z = rng.standard_normal((T, L))
    return y, X, Z


def main():
    out = os.path.join(HERE, "tprf_oos_pyref")
    os.makedirs(out, exist_ok=True)
    for name, seed, T, N, L, proc, kw, auto in CASES:
        y, X, Z = panel(seed, T, N, L)
        f, pe, _ = tprf3fast.estimate3prf_fast(y, X, L if auto else Z, procedure=proc, **kw)
        np.savez(os.path.join(out, name + ".npz"), seed=seed, T=T, N=N, L=L, procedure=proc, auto=auto,
                 forecasts=f, rsquare=pe.rsquare, enc=getattr(pe, "encnew", np.nan),
                 **{"kw_" + k: np.array(v) for k, v in kw.items()})
        print(name, "r2", pe.rsquare, "enc", getattr(pe, "encnew", None), "nan forecasts", int(np.isnan(f).sum()))


if __name__ == "__main__":
    main()
