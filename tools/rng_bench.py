"""Times the device generator: MatD.rand / randn fills of n x n FP64 (CUDA events around the whole call; randn's
figure includes its host synchronisation) and NumPy's default_rng on one host core beside it."""
import ctypes as C
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import uni_b200 as u
from uni_b200._lib import check, lib

u.init(0)
ms = C.c_float(0)
for n in (4096, 16384, 32768):
    g = u.NumPyRNG(0)
    for name, fn in (("rand", lambda: g.rand(n, n)), ("randn", lambda: g.randn(n, n))):
        best = 1e30
        for _ in range(5):
            check(lib.um_timer_start())
            m = fn()
            check(lib.um_timer_stop(C.byref(ms)))
            best = min(best, ms.value)
            del m
        E = float(n) * n
        print(f"{name:6s} {n}x{n}: {best:8.3f} ms  {8 * E / best / 1e6:8.1f} GB/s written  {E / best / 1e6:7.2f} Gdraw/s", flush=True)
r = np.random.default_rng(0)
t0 = time.perf_counter(); r.standard_normal(1 << 24); t1 = time.perf_counter(); r.random(1 << 24); t2 = time.perf_counter()
print(f"numpy 1 core: standard_normal {(1 << 24) / (t1 - t0) / 1e9:.3f} Gdraw/s, random {(1 << 24) / (t2 - t1) / 1e9:.3f} Gdraw/s")
