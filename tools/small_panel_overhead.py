"""Where does the time of a SMALL fit go (the per-GPU share of config 4 at 8 GPUs is T=8192 x N=32768)?
Event time per fit vs wall time per fit vs the sum of per-kernel event times."""
import ctypes as C
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uni_b200 as u
from uni_b200 import tprf3, _lib as L_
from uni_b200._lib import lib, check
T, N, Lp = 8192, int(sys.argv[1]) if len(sys.argv) > 1 else 32768, 8
u.init(0)
X = u.MatD.randn(T, N, seed=1000); y = u.MatD.randn(T, 1, seed=2); Z = u.MatD.randn(T, Lp, seed=3)
names = ["colstats/sweep", "colfinal", "project", "gather", "finish", "outputs", "allreduce", "prep"]
for path, name in ((0, "auto"), (2, "cluster only")):
    tprf3.set_tprf_path(path)
    MODE = int(os.environ.get("UM_MODE", L_.TPRF_CLOSED))
    fit = lambda: tprf3._fit(MODE, y, X, Z, n_total=N, want_all=False)
    for _ in range(5):
        fit()
    ms = C.c_float(0)
    reps = 30
    u.sync(); t0 = time.perf_counter()
    check(lib.um_timer_start())
    for _ in range(reps):
        fit()
    check(lib.um_timer_stop(C.byref(ms)))
    wall = (time.perf_counter() - t0) * 1e3
    print(f"{name}: N={N}: {ms.value / reps:.4f} ms/fit by events, {wall / reps:.4f} ms/fit wall")
    check(lib.um_prof_reset()); check(lib.um_prof_enable(1))
    for _ in range(10):
        fit()
    check(lib.um_prof_enable(0))
    tot = 0.0
    for s, nm in enumerate(names):
        t = C.c_double(0); c = C.c_int64(0)
        check(lib.um_prof_read(s, C.byref(t), C.byref(c)))
        if c.value:
            print(f"    {nm:16s} {t.value / 10:.4f} ms/fit ({c.value // 10} launches)")
            tot += t.value / 10
    print(f"    kernels sum      {tot:.4f} ms/fit")
tprf3.set_tprf_path(0)
