"""Time the config-4 3PRF fit through one sweep path: python tools/sweep_time.py [path=2] [N=65536] [T=8192] [reps=10]
(UNIMAT_LIB selects a side-by-side build of the library, UM_TPRF_SYM=0 the round-1 A/B-warp cluster kernel)."""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uni_b200 as u
from uni_b200 import tprf3, _lib as L_
from uni_b200._lib import lib, check
path = int(sys.argv[1]) if len(sys.argv) > 1 else 2
N = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
T = int(sys.argv[3]) if len(sys.argv) > 3 else 8192
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 10
u.init(0)
X = u.MatD.randn(T, N, seed=1000); y = u.MatD.randn(T, 1, seed=2); Z = u.MatD.randn(T, 8, seed=3)
tprf3.set_tprf_path(path)
fit = lambda: tprf3._fit(L_.TPRF_CLOSED, y, X, Z, n_total=N, want_all=False)
for _ in range(3):
    r = fit()
ms = C.c_float(0)
check(lib.um_timer_start())
for _ in range(reps):
    r = fit()
check(lib.um_timer_stop(C.byref(ms)))
t = ms.value / reps
print(f"{os.environ.get('UNIMAT_LIB', 'libunimat.so').split('/')[-1]:28s} path {path} T={T} N={N}: {t:.3f} ms/fit  {8.0 * T * N / t / 1e6:.0f} GB/s  r2={r.rSquared:.9g}")
