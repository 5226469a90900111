"""estimate3prf out-of-sample procedures on the device: python tools/oos_bench.py  (650 x 40 x 2 = the reference's published
shape, README.md:315-317, and a wide panel)"""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import uni_b200 as u
u.init(0)
for T, N, L in ((650, 40, 2), (200, 16384, 4)):
    rng = np.random.default_rng(1)
    X = rng.standard_normal((T, N)); y = rng.standard_normal((T, 1)); Z = rng.standard_normal((T, L))
    dy, dX, dZ = u.Mat.from_numpy(y), u.Mat.from_numpy(X), u.Mat.from_numpy(Z)
    for proc, kw in (("OOS Recursive", {}), ("OOS Cross Val", {}), ("OOS Rolling", {"rollwin": (60, 20, 0)})):
        for zarg, zn in ((dZ, "Z matrix"), (2, "2 automatic proxies")):
            f = lambda: u.estimate3prf(dy, dX, zarg, proc, **kw)
            f(); u.sync()
            best = 1e30
            for _ in range(5):
                t0 = time.perf_counter(); r = f(); u.sync(); best = min(best, time.perf_counter() - t0)
            print(f"T={T} N={N} L={L} {proc:14s} {zn:20s}: {best * 1e3:8.3f} ms   r2={r.rSquared:.6g}")
