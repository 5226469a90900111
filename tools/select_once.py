"""One m(mask) on an n x n FP64 matrix with half of the elements selected (for ncu launch lists)."""
import sys

sys.path.insert(0, ".")
import uni_b200 as u

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
u.init(0)
m = u.MatD.randn(n, n, seed=42)
mask = m.gt(0.0)
r = m[mask]
u.sync()
print("ok", r.shape)
