"""Accuracy of the device exp (MatD.exp -> fast_exp) against NumPy's correctly-rounded-ish exp, in ulp and relative."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uni_b200 as u
u.init(0)
rng = np.random.default_rng(0)
for lo, hi in ((-1, 1), (-20, 20), (-700, 700), (-745, -700), (700, 709.7)):
    x = rng.uniform(lo, hi, size=(2048, 2048))
    got = u.Mat.from_numpy(x).exp().to_numpy()
    want = np.exp(x.astype(np.longdouble)).astype(np.float64)
    rel = np.abs(got - want) / want
    ulp = np.abs(got - want) / np.spacing(want)
    print(f"[{lo}, {hi}]: max rel {rel.max():.3e}  max ulp {ulp.max():.2f}  mean ulp {ulp.mean():.3f}")
x = np.array([[0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, np.nan, 710.0, -746.0, 1e-300, 0.021660849392498290, -0.021660849392498290]])
print(u.Mat.from_numpy(x).exp().to_numpy(), np.exp(x))
