"""One 1024^2 (64 x 64 tile) and one 4096^2 (128 x 128 tile) FP64 product (ncu target)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uni_b200 as u
u.init(0)
for n in (1024, 4096):
    a = u.MatD.randn(n, n, seed=1); b = u.MatD.randn(n, n, seed=2)
    c = a @ b
u.sync()
print("ok")
