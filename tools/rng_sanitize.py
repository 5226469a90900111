"""Small generator workload for compute-sanitizer (memcheck / racecheck): every kernel of csrc/rng.cu."""
import sys

sys.path.insert(0, ".")
import numpy as np

import uni_b200 as u

u.init(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 150000
want = np.random.default_rng(5).standard_normal(n)
g = u.NumPyRNG(5)
z = g.randn(1, n).to_numpy().ravel()
tail = np.abs(want) > 3.6541528853610088
assert np.array_equal(z[~tail], want[~tail])
g.normal(1.0, 2.0, 33, 7)
g = u.NumPyRNG(6)
r = np.random.default_rng(6)
assert np.array_equal(g.rand(1, n).to_numpy().ravel(), r.random(n))
g.uniform_mat(-1.0, 1.0, 17, 5); g.randint(0, 100, 9, 7); g.randint(0, 100, 3, 3)
u.sync()
print("rng sanitize workload ok")
