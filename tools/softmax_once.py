"""One softmax(0) through the strip kernel and one softmax(1) (ncu target): python tools/softmax_once.py [n=32768]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uni_b200 as u
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
u.init(0)
m = u.MatD.randn(n, n, seed=42)
u.set_softmax_axis0_path(2)
for _ in range(2):
    r = m.softmax(0)
r1 = m.softmax(1)
u.sync()
print("ok")
