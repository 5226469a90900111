"""A few three-pass fits of a T x N panel (ncu target for k_tprf_finish): python tools/finish_prof.py [T] [N]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uni_b200 as u
from uni_b200 import tprf3, _lib as L_
T = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
u.init(0)
X = u.MatD.randn(T, N, seed=1000); y = u.MatD.randn(T, 1, seed=2); Z = u.MatD.randn(T, 8, seed=3)
for _ in range(4):
    r = tprf3._fit(L_.TPRF_ITER, y, X, Z, n_total=N, want_all=False)
u.sync()
print("r2", r.rSquared)
