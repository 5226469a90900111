"""softmax(axis) timing + check against a two-pass NumPy evaluation on sampled lanes: python tools/softmax_time.py [n=32768]
(axis 0 is timed through the two-read kernels and through the single-read strip kernel)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import uni_b200 as u
from uni_b200._lib import lib, check
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
u.init(0)
m = u.MatD.randn(n, n, seed=42)
for ax, path, name in ((0, 1, "two-read"), (0, 2, "strip"), (0, 0, "auto: strip + two-read split"), (1, 0, "row on chip")):
    u.set_softmax_axis0_path(path)
    for _ in range(3):
        r = m.softmax(ax)
    best = 1e30; ms = C.c_float(0)
    for _ in range(10):
        check(lib.um_timer_start()); r = m.softmax(ax); check(lib.um_timer_stop(C.byref(ms))); best = min(best, ms.value)
    idx = [0, 1, 63, 64, 65, n // 2 + 3, n - 2, n - 1]
    err = 0.0
    for j in idx:
        lane = (m.slice(0, n, j, j + 1) if ax == 0 else m.slice(j, j + 1, 0, n)).to_numpy().ravel()
        got = (r.slice(0, n, j, j + 1) if ax == 0 else r.slice(j, j + 1, 0, n)).to_numpy().ravel()
        e = np.exp(lane - lane.max()); want = e / e.sum()
        err = max(err, float(np.max(np.abs(got - want) / want)))
    print(f"softmax axis {ax} ({name}) n={n}: {best:.3f} ms  {16.0 * n * n / best / 1e6:.0f} GB/s (16E)  max rel err on sampled lanes {err:.2e}", flush=True)
u.set_softmax_axis0_path(0)
