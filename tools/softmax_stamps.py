import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import uni_b200 as u
from uni_b200._lib import lib, check
n = 32768
u.init(0)
m = u.MatD.randn(n, n, seed=42)
u.set_softmax_axis0_path(2)
for _ in range(3): r = m.softmax(0)
u.sync()
buf = (C.c_longlong * 1024)()
lib.um_dbg_softmax_stamps.argtypes = [C.POINTER(C.c_longlong)]
print("rc", lib.um_dbg_softmax_stamps(buf))
a = np.array(buf[:]).reshape(256, 4)
d = np.diff(a, axis=1)
print("phase A, exchange, phase B (ns), first 12 strips:\n", d[:12])
print("medians ns: A %.0f  exch %.0f  B %.0f ; strip period %.0f" % (np.median(d[:, 0]), np.median(d[:, 1]), np.median(d[:, 2]), np.median(np.diff(a[:, 0]))))
