"""Config-4 fit time through each 3PRF sweep (1 two-sweep, 2 cluster sweep, 3 grid sweep), same process, same panel."""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uni_b200 as u
from uni_b200 import tprf3, _lib as L_
from uni_b200._lib import lib, check
T, N, Lp = 8192, int(sys.argv[1]) if len(sys.argv) > 1 else 262144, 8
u.init(0)
X = u.MatD.randn(T, N, seed=1000); y = u.MatD.randn(T, 1, seed=2); Z = u.MatD.randn(T, Lp, seed=3)
ref = None
for path, name in ((1, "two-sweep"), (2, "cluster sweep"), (3, "grid sweep"), (0, "auto (split)"), (2, "cluster sweep"), (0, "auto (split)")):
    tprf3.set_tprf_path(path)
    fit = lambda: tprf3._fit(L_.TPRF_CLOSED, y, X, Z, n_total=N, want_all=False)
    for _ in range(3):
        r = fit()
    ms = C.c_float(0)
    check(lib.um_timer_start())
    for _ in range(10):
        r = fit()
    check(lib.um_timer_stop(C.byref(ms)))
    f = r.forecasts.to_numpy()
    if ref is None:
        ref = f
    import numpy as np
    print(f"{name:14s}: {ms.value / 10:.3f} ms/fit   max |yhat - two-sweep| = {np.max(np.abs(f - ref)):.2e}   r2 = {r.rSquared:.9g}")
tprf3.set_tprf_path(0)
