"""MatF `a @ b` on the tcgen05 3xTF32 path vs float64 NumPy; timing vs the FFMA kernel (UM_MATF_TENSOR=0)."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uni_b200 as u
from uni_b200._lib import lib, check
u.init(0)
shapes = [(128, 128, 32), (128, 128, 128), (256, 384, 512), (1024, 1024, 1024), (2048, 1024, 4096)]
if len(sys.argv) > 1:
    shapes = [tuple(int(x) for x in a.split("x")) for a in sys.argv[1:]]
rng = np.random.default_rng(0)
for (m, n, k) in shapes:
    a = rng.standard_normal((m, k)).astype(np.float32); b = rng.standard_normal((k, n)).astype(np.float32)
    ref = a.astype(np.float64) @ b.astype(np.float64)
    f32 = a @ b
    got = (u.Mat.from_numpy(a) @ u.Mat.from_numpy(b)).to_numpy()
    scale = np.sqrt(k)
    print(f"{m}x{n}x{k}: max |ours - f64| / sqrt(k) = {np.max(np.abs(got - ref)) / scale:.3e}   numpy f32: {np.max(np.abs(f32 - ref)) / scale:.3e}")
for n in (4096, 8192):
    a = u.MatF.randn(n, n, seed=1); b = u.MatF.randn(n, n, seed=2)
    for _ in range(2):
        a @ b
    ms = C.c_float(0); best = 1e30
    for _ in range(3):
        check(lib.um_timer_start()); a @ b; check(lib.um_timer_stop(C.byref(ms))); best = min(best, ms.value)
    print(f"f32 {n}^3: {best:.3f} ms  {2.0 * n ** 3 / best / 1e9:.1f} TFLOP/s")
