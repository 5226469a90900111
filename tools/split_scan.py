"""Config-4 fit time through the auto (split) path, repeated, for the current UM_TPRF_SPLIT_SPEED."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uni_b200 as u
from uni_b200 import tprf3, _lib as L_
from uni_b200._lib import lib, check
T, N, Lp = 8192, int(sys.argv[1]) if len(sys.argv) > 1 else 262144, 8
u.init(0)
X = u.MatD.randn(T, N, seed=1000); y = u.MatD.randn(T, 1, seed=2); Z = u.MatD.randn(T, Lp, seed=3)
fit = lambda: tprf3._fit(L_.TPRF_CLOSED, y, X, Z, n_total=N, want_all=False)
out = []
for rep in range(4):
    for _ in range(3):
        fit()
    ms = C.c_float(0)
    check(lib.um_timer_start())
    for _ in range(10):
        fit()
    check(lib.um_timer_stop(C.byref(ms)))
    out.append(ms.value / 10)
print(os.environ.get("UM_TPRF_SPLIT_SPEED", "default"), " ".join(f"{x:.3f}" for x in out))
