"""Compares the single-read cluster sweep with the two-sweep path and the CPU oracle on a few shapes,
and times both on a mid-size panel.  Run on a GPU box:  python tools/fused_check.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uni_b200 as u
from uni_b200 import tprf3
from oracle import tprf_oracle as to

u.init(0)


def relerr(a, b):
    return float(np.max(np.abs(a - b) / (1e-12 + 1e-9 * np.maximum(np.abs(a), np.abs(b)))))


shapes = [(256, 1000, 4), (512, 2048, 4), (50, 6, 2), (600, 130, 3), (1000, 512, 8), (2048, 256, 8), (4096, 320, 5),
          (8192, 1024, 8), (8000, 200, 8), (130, 48, 1)]
if len(sys.argv) > 1:
    shapes = [tuple(int(x) for x in a.split("x")) for a in sys.argv[1:]]
for (T, N, L) in shapes:
    rng = np.random.default_rng(T + N)
    X = rng.standard_normal((T, N)) * (1 + 3 * rng.random(N)) + 5 * rng.standard_normal(N)
    y = rng.standard_normal((T, 1)); Z = rng.standard_normal((T, L))
    o = to.tprf_closed_form_gram(y, X, Z)
    dy, dX, dZ = u.Mat.from_numpy(y), u.Mat.from_numpy(X), u.Mat.from_numpy(Z)
    out = {}
    for name, path in (("two", 1), ("fused", 2), ("grid", 3)):
        tprf3.set_tprf_path(path)
        try:
            r = u.tprfClosedForm(dy, dX, dZ)
            it = u.t3prf(dy, dX, dZ)
            out[name] = (r.forecasts.to_numpy(), r.alpha.to_numpy(), r.rSquared, it.forecasts.to_numpy(), it.phi.to_numpy(), it.sigma.to_numpy(), r.xstd.to_numpy())
        except Exception as e:  # noqa
            out[name] = None
            print(f"T={T} N={N} L={L} {name}: {type(e).__name__}: {e}")
    tprf3.set_tprf_path(0)
    for name, v in out.items():
        if v is None:
            continue
        print(f"T={T} N={N} L={L} {name:5s}: yhat err/tol {relerr(v[0], o.forecasts):.3g}  alpha {relerr(v[1], o.alpha):.3g}  "
              f"r2 {abs(v[2] - o.r_squared):.2e}  iter-yhat {relerr(v[3], o.forecasts):.3g}")
    if out.get("two") and out.get("grid"):
        print(f"    grid vs two-sweep: yhat {relerr(out['grid'][0], out['two'][0]):.3g}  alpha {relerr(out['grid'][1], out['two'][1]):.3g}"
              f"  phi {relerr(out['grid'][4], out['two'][4]):.3g}  sigma {relerr(out['grid'][5], out['two'][5]):.3g}  xstd {relerr(out['grid'][6], out['two'][6]):.3g}")
    if out.get("two") and out.get("fused"):
        print(f"    fused vs two-sweep: yhat {relerr(out['fused'][0], out['two'][0]):.3g}  alpha {relerr(out['fused'][1], out['two'][1]):.3g}"
              f"  phi {relerr(out['fused'][4], out['two'][4]):.3g}  sigma {relerr(out['fused'][5], out['two'][5]):.3g}  xstd {relerr(out['fused'][6], out['two'][6]):.3g}")

# timing
import ctypes as C
from uni_b200._lib import lib, check
g_, r_, n_ = C.c_int32(), C.c_int32(), C.c_int32()
for Tq in (512, 1024, 2048, 4096, 8192):
    check(lib.um_tprf_fused_info(Tq, C.byref(g_), C.byref(r_), C.byref(n_)))
    print(f"fused shape T={Tq}: cluster {g_.value} x {r_.value} rows, resident clusters {n_.value}")
T, N, L = 8192, 65536, 8
rng = np.random.default_rng(1)
X = rng.standard_normal((T, N)); y = rng.standard_normal((T, 1)); Z = rng.standard_normal((T, L))
dy, dX, dZ = u.Mat.from_numpy(y), u.Mat.from_numpy(X), u.Mat.from_numpy(Z)
for name, path in (("two", 1), ("fused", 2), ("grid", 3)):
    tprf3.set_tprf_path(path)
    for _ in range(3):
        r = u.tprfClosedForm(dy, dX, dZ)
    u.sync()
    check(lib.um_prof_enable(1)); check(lib.um_prof_reset())
    t0 = time.perf_counter()
    for _ in range(10):
        r = u.tprfClosedForm(dy, dX, dZ)
    u.sync()
    dt = (time.perf_counter() - t0) / 10
    names = ["sweep1/fused", "colfinal", "project", "gather", "finish", "outputs", "allreduce", "prep"]
    for i, nm in enumerate(names):
        tot, cnt = C.c_double(), C.c_int64()
        check(lib.um_prof_read(i, C.byref(tot), C.byref(cnt)))
        if cnt.value:
            print(f"      {nm:14s} {tot.value / cnt.value:.4f} ms x{cnt.value}")
    check(lib.um_prof_enable(0))
    print(f"T={T} N={N} L={L} {name}: {dt * 1e3:.3f} ms/fit (wall)  -> {T * N * 8 / dt / 1e9:.0f} GB/s  r2={r.rSquared:.6g}")
tprf3.set_tprf_path(0)
