"""Closed-form vs three-pass fit of the config-4 panel, with the per-kernel breakdown: python tools/iter_vs_closed.py [N=262144]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uni_b200 as u
from uni_b200 import tprf3, _lib as L_
from uni_b200._lib import lib, check
T, N = 8192, int(sys.argv[1]) if len(sys.argv) > 1 else 262144
u.init(0)
X = u.MatD.randn(T, N, seed=1000); y = u.MatD.randn(T, 1, seed=2); Z = u.MatD.randn(T, 8, seed=3)
names = ["sweep", "colfinal", "project", "gather", "finish", "outputs", "allreduce", "prep"]
ms = C.c_float(0)
for rep in range(2):
    for mode, nm in ((L_.TPRF_CLOSED, "closed"), (L_.TPRF_ITER, "t3prf")):
        fit = lambda: tprf3._fit(mode, y, X, Z, n_total=N, want_all=False)
        for _ in range(3): fit()
        check(lib.um_timer_start())
        for _ in range(10): fit()
        check(lib.um_timer_stop(C.byref(ms)))
        line = f"{nm:7s} {ms.value / 10:.4f} ms/fit |"
        check(lib.um_prof_reset()); check(lib.um_prof_enable(1))
        for _ in range(5): fit()
        check(lib.um_prof_enable(0))
        for s, k in enumerate(names):
            t = C.c_double(0); c = C.c_int64(0)
            check(lib.um_prof_read(s, C.byref(t), C.byref(c)))
            if c.value: line += f" {k} {t.value / 5:.4f}"
        print(line, flush=True)
