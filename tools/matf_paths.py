"""FP32 `a @ b`: the tensor-core kernel with operands split in shared memory (path 2) against operands split once in
global memory (path 3): bit-equality on odd shapes and time per size.  python tools/matf_paths.py"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uni_b200 as u
from uni_b200._lib import lib, check
from uni_b200.mat import set_matf_matmul_path
u.init(0)
rng = np.random.default_rng(0)
for (m, n, k) in [(128, 128, 32), (200, 132, 77), (256, 384, 513), (1000, 1028, 1030), (2048, 1024, 4096), (3000, 2500, 130)]:
    kp = (k + 3) // 4 * 4          # K itself need not be a multiple of 4, the row pitch must: A is a column slice
    af = rng.standard_normal((m, kp)).astype(np.float32); b = rng.standard_normal((k, n)).astype(np.float32)
    a = np.ascontiguousarray(af[:, :k])
    da, db = u.Mat.from_numpy(af)._raw_slice(0, m, 0, k), u.Mat.from_numpy(b)
    set_matf_matmul_path(2); c2 = (da @ db).to_numpy()
    set_matf_matmul_path(3); c3 = (da @ db).to_numpy()
    ref = a.astype(np.float64) @ b.astype(np.float64)
    print(f"{m}x{n}x{k}: identical {np.array_equal(c2, c3)}  max |c3 - f64| / sqrt(k) = {np.max(np.abs(c3 - ref)) / np.sqrt(k):.3e}", flush=True)
ms = C.c_float(0)
for n in (1024, 1536, 2048, 3072, 4096, 8192, 16384):
    a = u.MatF.randn(n, n, seed=1); b = u.MatF.randn(n, n, seed=2)
    row = []
    for path in (2, 3, 0):
        set_matf_matmul_path(path)
        for _ in range(3):
            a @ b
        best = 1e30
        for _ in range(5 if n > 4096 else 20):
            check(lib.um_timer_start()); c = a @ b; check(lib.um_timer_stop(C.byref(ms))); best = min(best, ms.value)
        row.append(f"path {path}: {best:8.3f} ms {2.0 * n ** 3 / best / 1e9:6.1f} TF")
    print(f"n={n:6d}  " + "   ".join(row), flush=True)
set_matf_matmul_path(0)
