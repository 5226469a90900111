"""FP64 `*@` time per CTA tile: python tools/dgemm_time.py  (n = 512 .. 4096; tile 0 = automatic)"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uni_b200 as u
from uni_b200._lib import lib, check
u.init(0)
ms = C.c_float(0)
for n in (512, 1024, 1536, 2048, 3072, 4096):
    a = u.MatD.randn(n, n, seed=1); b = u.MatD.randn(n, n, seed=2)
    row = []
    for tile in (128, 64, 0):
        u.set_matd_matmul_tile(tile)
        for _ in range(5):
            c = a @ b
        best = 1e30
        for _ in range(20):
            check(lib.um_timer_start()); c = a @ b; check(lib.um_timer_stop(C.byref(ms))); best = min(best, ms.value)
        row.append(f"tile {tile:3d}: {best:7.3f} ms {2.0 * n ** 3 / best / 1e9:6.2f} TF")
    u.set_matd_matmul_tile(0)
    print(f"n={n:5d}  " + "   ".join(row), flush=True)
