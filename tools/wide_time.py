"""3PRF fits with more than 8 proxies (the operator path, csrc/tprf_wide.cu) next to the fused sweep at L = 8:
python tools/wide_time.py [N=262144] [T=8192]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uni_b200 as u
from uni_b200._lib import lib, check
N = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
T = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
u.init(0)
X = u.MatD.randn(T, N, seed=1000); y = u.MatD.randn(T, 1, seed=2)
ms = C.c_float(0)
for L in (8, 9, 16, 32):
    Z = u.MatD.randn(T, L, seed=3)
    for fit, name in ((u.tprfClosedForm, "tprfClosedForm"), (u.t3prf, "t3prf")):
        r = fit(y, X, Z)
        check(lib.um_timer_start())
        for _ in range(3):
            r = fit(y, X, Z)
        check(lib.um_timer_stop(C.byref(ms)))
        print(f"T={T} N={N} L={L:2d} {name:15s}: {ms.value / 3:8.2f} ms per fit   r2={r.rSquared:.6g}", flush=True)
