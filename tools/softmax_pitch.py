"""Does the power-of-two row pitch (256 KB at 32768 columns) cost the strip kernel anything?  softmax(0) of a 32768 x 32768 view of
a buffer whose pitch is 32768 + pad columns, forced through the strip kernel."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uni_b200 as u
from uni_b200._lib import lib, check
u.init(0)
n = 32768
ms = C.c_float(0)
for pad in (0, 16, 48, 272):
    big = u.MatD.randn(n, n + pad, seed=1)
    m = big if pad == 0 else big._raw_slice(0, n, 0, n)
    for path, name in ((2, "strip"), (1, "two-read")):
        u.set_softmax_axis0_path(path)
        for _ in range(3): r = m.softmax(0)
        best = 1e30
        for _ in range(8):
            check(lib.um_timer_start()); r = m.softmax(0); check(lib.um_timer_stop(C.byref(ms))); best = min(best, ms.value)
        print(f"input pitch {n + pad} columns, {name}: {best:.3f} ms", flush=True)
    del r, m, big
u.set_softmax_axis0_path(0)
