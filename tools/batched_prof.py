"""Per-kernel device times of the config-5 batched fit (4096 panels of 512 x 2048 x 4)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uni_b200 as u
from uni_b200 import _lib as L_
from uni_b200._lib import lib, check
u.init(0)
T, N, Lp, nb = 512, 2048, 4, int(sys.argv[1]) if len(sys.argv) > 1 else 4096
X = u.MatD.randn(nb * T, N, seed=7); y = u.MatD.randn(nb * T, 1, seed=8); Z = u.MatD.randn(nb * T, Lp, seed=9)
f = u.MatD.zeros(nb * T, 1); r2 = u.MatD.zeros(nb, 1)
p = lambda m: C.c_void_p(m._buf.ptr)
call = lambda: check(lib.um_tprf_fit_batched(L_.TPRF_CLOSED, p(y), p(X), p(Z), T, N, Lp, nb, p(f), None, p(r2)))
for _ in range(3):
    call()
ms = C.c_float(0)
check(lib.um_timer_start())
for _ in range(5):
    call()
check(lib.um_timer_stop(C.byref(ms)))
print(f"{nb} panels: {ms.value / 5:.3f} ms per call, {8.0 * T * N * nb / (ms.value / 5 * 1e-3) / 1e9:.0f} GB/s")
check(lib.um_prof_reset()); check(lib.um_prof_enable(1))
for _ in range(5):
    call()
check(lib.um_prof_enable(0))
names = ["colstats/sweep", "colfinal/smallsums", "project", "gather", "finish", "outputs", "allreduce", "prep"]
for s, nm in enumerate(names):
    t = C.c_double(0); c = C.c_int64(0)
    check(lib.um_prof_read(s, C.byref(t), C.byref(c)))
    if c.value:
        print(f"    {nm:20s} {t.value / 5:.4f} ms per call ({c.value // 5} launches)")
