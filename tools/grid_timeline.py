"""Per-piece timeline of the grid-wide single-read 3PRF sweep (group 0) from the kernel's clock64 stamps.
usage: python tools/grid_timeline.py [T] [N] [L]"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uni_b200 as u
from uni_b200 import tprf3
from uni_b200._lib import lib, check
from uni_b200.mat import Mat

T = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
N = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
L = int(sys.argv[3]) if len(sys.argv) > 3 else 8
u.init(0)
rng = np.random.default_rng(0)
X = rng.standard_normal((T, N)); y = rng.standard_normal((T, 1)); Z = rng.standard_normal((T, L))
dy, dX, dZ = u.Mat.from_numpy(y), u.Mat.from_numpy(X), u.Mat.from_numpy(Z)
tprf3.set_tprf_path(3)
for _ in range(2):
    u.tprfClosedForm(dy, dX, dZ)
buf = Mat.from_numpy(np.zeros((16 * 16 * 64 * 8, 1)))
check(lib.um_tprf_debug_stamps(C.c_void_p(buf._buf.ptr)))
u.tprfClosedForm(dy, dX, dZ)
u.sync()
check(lib.um_tprf_debug_stamps(None))
st = buf.to_numpy().view(np.int64).reshape(16, 16, 64, 8)
its = list(range(10, 40))
for r in (0, 15):
    a, b, f, red = st[r, 0], st[r, 4], st[r, 9], st[r, 10]
    per = np.mean([a[i + 1, 0] - a[i, 0] for i in its])
    m = lambda x: float(np.mean(x))
    print(f"rank {r:2d}: {per:7.0f} clk/piece | A-warp: wait TMA {m([a[i,1]-a[i,0] for i in its]):6.0f}, wait stash free {m([a[i,2]-a[i,1] for i in its]):6.0f}, "
          f"phase A {m([a[i,3]-a[i,2] for i in its]):6.0f} | B-warp: wait records {m([b[i,1]-b[i,0] for i in its]):6.0f}, wait stash {m([b[i,2]-b[i,1] for i in its]):6.0f}, "
          f"phase B {m([b[i,3]-b[i,2] for i in its]):6.0f}")
    ref = a[:, 3]
    print(f"         relative to end of A(i): reducer saw scratch {m([red[i,1]-ref[i] for i in its]):6.0f}, partial flagged {m([red[i,2]-ref[i] for i in its]):6.0f}, "
          f"finaliser saw all partials {m([f[i,1]-ref[i] for i in its]):6.0f}, record flagged {m([f[i,2]-ref[i] for i in its]):6.0f}, "
          f"B saw records {m([b[i,1]-ref[i] for i in its]):6.0f}, B done {m([b[i,3]-ref[i] for i in its]):6.0f}")
