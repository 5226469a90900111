"""Small driver for ncu: a few closed-form fits of a T x N panel through one chosen sweep.
usage: python tools/fused_prof.py [T] [N] [L] [path 0|1|2] [reps]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uni_b200 as u
from uni_b200 import tprf3

T = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
N = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
L = int(sys.argv[3]) if len(sys.argv) > 3 else 8
path = int(sys.argv[4]) if len(sys.argv) > 4 else 2
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
u.init(0)
rng = np.random.default_rng(0)
X = rng.standard_normal((T, N)); y = rng.standard_normal((T, 1)); Z = rng.standard_normal((T, L))
dy, dX, dZ = u.Mat.from_numpy(y), u.Mat.from_numpy(X), u.Mat.from_numpy(Z)
tprf3.set_tprf_path(path)
for _ in range(reps):
    r = u.tprfClosedForm(dy, dX, dZ)
u.sync()
print("r2", r.rSquared)
