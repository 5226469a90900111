"""sigmoid / exp timing at n x n FP64 (min of 20 after 5 warm-up calls) + accuracy against NumPy: python tools/sigmoid_time.py [n=32768]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import uni_b200 as u
from uni_b200._lib import lib, check
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
u.init(0)
m = u.MatD.randn(n, n, seed=42)
ms = C.c_float(0)
for name, fn in (("sigmoid", lambda: m.sigmoid()), ("neg", lambda: -m)):
    for _ in range(5): r = fn()
    best = 1e30
    for _ in range(20):
        check(lib.um_timer_start()); r = fn(); check(lib.um_timer_stop(C.byref(ms))); best = min(best, ms.value)
    print(f"{name}: {best:.3f} ms  {16.0 * n * n / best / 1e6:.0f} GB/s", flush=True)
x = np.concatenate([np.random.default_rng(0).uniform(-750, 750, 100000), [0.0, -0.0, np.inf, -np.inf, np.nan, 700.0, -700.0, -700.0000001, -708.3, -708.5, -745.0, -746.0]])[None, :]
got = u.Mat.from_numpy(x).sigmoid().to_numpy()
with np.errstate(all="ignore"):
    want = np.where(x >= 0, 1.0 / (1.0 + np.exp(-x)), np.exp(x) / (1.0 + np.exp(x)))
ok = np.isfinite(want) & (want != 0)
print("max rel err", np.max(np.abs(got[ok] - want[ok]) / np.abs(want[ok])), "nan ok", np.array_equal(np.isnan(got), np.isnan(want)), "zeros ok", np.array_equal(got[want == 0], want[want == 0]))
rel = np.where(ok, np.abs(got - want) / np.abs(np.where(ok, want, 1.0)), 0.0)
bad = np.argsort(rel.ravel())[-5:]
for i in bad: print("x", x.ravel()[i], "got", got.ravel()[i], "want", want.ravel()[i], "rel", rel.ravel()[i])
