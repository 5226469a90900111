"""Per-piece timeline of the single-read 3PRF sweep (cluster 0) from the kernel's clock64 stamps.
usage: python tools/fused_timeline.py [T] [N] [L]"""
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
N = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
L = int(sys.argv[3]) if len(sys.argv) > 3 else 8
u.init(0)
rng = np.random.default_rng(0)
X = rng.standard_normal((T, N)); y = rng.standard_normal((T, 1)); Z = rng.standard_normal((T, L))
dy, dX, dZ = u.Mat.from_numpy(y), u.Mat.from_numpy(X), u.Mat.from_numpy(Z)
tprf3.set_tprf_path(2)
for _ in range(2):
    u.tprfClosedForm(dy, dX, dZ)
n = 16 * 16 * 64 * 8
buf = Mat.from_numpy(np.zeros((n, 1)))      # 8-byte cells, reinterpreted as int64
check(lib.um_tprf_debug_stamps(C.c_void_p(buf._buf.ptr)))
u.tprfClosedForm(dy, dX, dZ)
u.sync()
check(lib.um_tprf_debug_stamps(None))
st = buf.to_numpy().view(np.int64).reshape(16, 16, 64, 8)
names = ["iter start", "full landed", "A done", "bar1+scratch+bar2", "reduce+push done", "vready seen", "B done"]
g_ = C.c_int32(); r_ = C.c_int32(); n_ = C.c_int32()
check(lib.um_tprf_fused_info(T, C.byref(g_), C.byref(r_), C.byref(n_)))
G = g_.value
print(f"T={T} N={N}: cluster {G} x {r_.value} rows, {n_.value} resident clusters")
for rank in sorted({0, G - 1}):
    for warp in (0, 15):
        s = st[rank, warp]
        its = [i for i in range(8, 40) if s[i, 0] and s[i + 1, 0]]
        if not its:
            continue
        per = np.mean([s[i + 1, 0] - s[i, 0] for i in its])
        segs = []
        for k in range(1, 7):
            d = [s[i, k] - s[i, k - 1] for i in its if s[i, k] and s[i, k - 1]]
            segs.append(np.mean(d) if d else float("nan"))
        print(f"rank {rank:2d} warp {warp:2d}: {per:7.0f} clk/piece | wait full {segs[0]:6.0f} | A {segs[1]:6.0f} | bars {segs[2]:6.0f} | "
              f"reduce+push {segs[3]:6.0f} | wait vready {segs[4]:6.0f} | B {segs[5]:6.0f}")
