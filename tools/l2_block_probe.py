"""Probe: does a column block of X stay L2-resident between the two sweeps?  (UM_TPRF_BLOCK_COLS)"""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uni_b200 as u
from uni_b200 import _lib as L_
from uni_b200._lib import lib, check
u.init(0)
T, N, Lp = 8192, int(os.environ.get("PROBE_N", 65536)), 8
X = u.MatD.randn(T, N, seed=1); y = u.MatD.randn(T, 1, seed=2); Z = u.MatD.randn(T, Lp, seed=3)
fit = lambda: u.tprf3._fit(L_.TPRF_CLOSED, y, X, Z, want_all=False)
for blk in [int(b) for b in os.environ.get("PROBE_BLOCKS", "0,2048,1024,512,256").split(",")]:
    if blk: os.environ["UM_TPRF_BLOCK_COLS"] = str(blk)
    else: os.environ.pop("UM_TPRF_BLOCK_COLS", None)
    for _ in range(2): r = fit()
    ms = C.c_float(0)
    check(lib.um_timer_start())
    for _ in range(3): r = fit()
    check(lib.um_timer_stop(C.byref(ms)))
    print(f"block_cols {blk:6d} ({blk*T*8/2**20:7.1f} MiB)  {ms.value/3:.3f} ms/fit  r2 {r.rSquared:.9f}", flush=True)
