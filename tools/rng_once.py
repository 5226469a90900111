"""One rand and one randn fill of n x n (for ncu launch lists / --set full captures)."""
import sys

sys.path.insert(0, ".")
import uni_b200 as u

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
u.init(0)
g = u.NumPyRNG(0)
a = g.rand(n, n)
del a
b = g.randn(n, n)
u.sync()
print("ok", b.shape)
