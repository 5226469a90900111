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
SYM = False   # (the symmetric-warp experiment of round 2 used a 32-warp stamp layout; see git history)
WPR = 32 if SYM else 16                      # warps per rank in the stamp buffer
n = 16 * WPR * 64 * 8
buf = Mat.from_numpy(np.zeros((n, 1)))      # 8-byte cells, reinterpreted as int64
check(lib.um_tprf_debug_stamps(C.c_void_p(buf._buf.ptr)))
u.tprfClosedForm(dy, dX, dZ)
u.sync()
check(lib.um_tprf_debug_stamps(None))
st = buf.to_numpy().view(np.int64).reshape(16, WPR, 64, 8)
g_ = C.c_int32(); r_ = C.c_int32(); n_ = C.c_int32()
check(lib.um_tprf_fused_info(T, C.byref(g_), C.byref(r_), C.byref(n_)))
G = g_.value
print(f"T={T} N={N}: cluster {G} x {r_.value} rows, {n_.value} resident clusters, {'symmetric' if SYM else 'A/B'} warps")
if SYM:
    # tprf_sym.cuh: warps 0-7 = A (stamps 0 wait, 1 landed, 2 done), 8-15 = B (3 wait, 4 records seen, 5 done), 17 finaliser, 18 reducer
    for rank in sorted({0, G - 1}):
        for warp in (0, 7):
            s = st[rank, warp]
            its = [i for i in range(8, 40) if s[i, 0] and s[i + 1, 0]]
            if not its:
                continue
            per = np.mean([s[i + 1, 0] - s[i, 0] for i in its])
            wa = np.mean([s[i, 1] - s[i, 0] for i in its])
            pa = np.mean([s[i, 2] - s[i, 1] for i in its])
            print(f"rank {rank:2d} warp {warp:2d} (A): {per:7.0f} clk/piece | wait TMA {wa:6.0f} | phase A + partials {pa:6.0f}")
        for warp in (8, 15):
            s = st[rank, warp]
            its = [i for i in range(8, 40) if s[i, 3] and s[i + 1, 3]]
            if not its:
                continue
            per = np.mean([s[i + 1, 3] - s[i, 3] for i in its])
            wb = np.mean([s[i, 4] - s[i, 3] for i in its])
            pb = np.mean([s[i, 5] - s[i, 4] for i in its])
            print(f"rank {rank:2d} warp {warp:2d} (B): {per:7.0f} clk/piece | wait records {wb:6.0f} | phase B {pb:6.0f}")
    r = 0
    its = list(range(8, 40))
    ref = st[r, 0, :, 2]
    def rel(w, k):
        return np.mean([st[r, w, i, k] - ref[i] for i in its])
    slowA = np.mean([max(st[r, w, i, 2] for w in range(8)) - ref[i] for i in its])
    print(f"rank 0, relative to end of A(i) on warp 0: last A-warp done {slowA:6.0f} | reducer sees scratch {rel(18, 1):6.0f} | reducer pushed {rel(18, 2):6.0f} | "
          f"finaliser sees all partials {rel(17, 1):6.0f} | finaliser pushed {rel(17, 2):6.0f} | B-warp 8 sees records {rel(8, 4):6.0f} | B done {rel(8, 5):6.0f}")
    sys.exit(0)
for rank in sorted({0, G - 1}):
    for warp, (a, b, c, nm) in ((0, (0, 1, 2, "A")), (3, (0, 1, 2, "A")), (4, (3, 4, 5, "B")), (7, (3, 4, 5, "B"))):
        s = st[rank, warp]
        its = [i for i in range(8, 40) if s[i, a] and s[i + 1, a]]
        if not its:
            continue
        per = np.mean([s[i + 1, a] - s[i, a] for i in its])
        wait = np.mean([s[i, b] - s[i, a] for i in its])
        work = np.mean([s[i, c] - s[i, b] for i in its])
        lag = np.mean([st[rank, 4, i, 5] - st[rank, 0, i, 2] for i in its]) if nm == "A" else float("nan")
        print(f"rank {rank:2d} warp {warp:2d} ({nm}): {per:7.0f} clk/piece | wait {wait:6.0f} | phase {nm} {work:6.0f}" +
              (f" | end of A(i) -> end of B(i): {lag:6.0f}" if nm == "A" else ""))

# exchange chain of piece i on rank 0, relative to the end of A(i) on A-warp 0
r = 0
its = [i for i in range(8, 40)]
ref = st[r, 0, :, 2]
def rel(w, k):
    return np.mean([st[r, w, i, k] - ref[i] for i in its])
slowA = np.mean([max(st[r, w, i, 2] for w in range(4)) - ref[i] for i in its])
print(f"rank 0, relative to end of A(i) on warp 0: last A-warp done {slowA:6.0f} | reducer sees scratch {rel(10, 1):6.0f} | reducer pushed {rel(10, 2):6.0f} | finaliser sees all partials {rel(9, 1):6.0f} | "
      f"finaliser pushed {rel(9, 2):6.0f} | B-warp sees records {rel(4, 4):6.0f} | B done {rel(4, 5):6.0f}")
