"""Times the UNMODIFIED reference NumPy implementation (/root/reference/py/tprf3fast.py::estimate3prf_fast) on the build
container's host cores and records the figures under profiles/ -- the reference tree does not travel to the GPU box, so
bench.py cannot run it there; bench.py's `--impl reference` arm times the oracle restatement instead and DESIGN.md quotes
this file beside it.   python tools/ref_python_timing.py [out.json]"""
import json
import os
import sys
import time

import numpy as np

sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference/py")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tprf3fast  # noqa: E402  (the reference's own module)
from oracle import tprf_oracle as to  # noqa: E402


def best(fn, reps=3):
    fn()
    b = 1e30
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); b = min(b, time.perf_counter() - t0)
    return b


out = {"cores": len(os.sched_getaffinity(0)), "numpy": np.__version__, "module": "/root/reference/py/tprf3fast.py::estimate3prf_fast", "rows": []}
rng = np.random.default_rng(0)
# config-4-shaped: T = 8192, L = 8, as many columns as the module's N x N alpha allows in a few seconds
for T, N, L in ((8192, 1024, 8), (8192, 4096, 8), (650, 40, 2)):
    X = rng.standard_normal((T, N)); y = rng.standard_normal((T, 1)); Z = rng.standard_normal((T, L))
    t_ref = best(lambda: tprf3fast.estimate3prf_fast(y, X, Z, "IS Full"))
    t_port = best(lambda: to.tprf_closed_form_gram(y, X, Z))
    f_ref = tprf3fast.estimate3prf_fast(y, X, Z, "IS Full")[0]
    f_port = to.tprf_closed_form_gram(y, X, Z).forecasts
    flop = T * N * (4.0 * L + 6.0)
    out["rows"].append({"T": T, "N": N, "L": L, "procedure": "IS Full", "reference_python_ms": round(t_ref * 1e3, 2),
                        "reference_python_tflops_algorithmic": round(flop / t_ref / 1e12, 6),
                        "oracle_gram_closed_form_ms": round(t_port * 1e3, 2), "oracle_tflops_algorithmic": round(flop / t_port / 1e12, 6),
                        "max_abs_forecast_diff_reference_vs_oracle": float(np.max(np.abs(np.asarray(f_ref).ravel() - f_port.ravel())))})
    print(out["rows"][-1], flush=True)
# the published OOS shape (README.md:315-317): 650 x 40 x 2
T, N, L = 650, 40, 2
X = rng.standard_normal((T, N)); y = rng.standard_normal((T, 1)); Z = rng.standard_normal((T, L))
for proc, kw in (("OOS Recursive", {}), ("OOS Cross Val", {})):
    t_ref = best(lambda: tprf3fast.estimate3prf_fast(y, X, Z, proc, **kw), reps=2)
    out["rows"].append({"T": T, "N": N, "L": L, "procedure": proc, "reference_python_ms": round(t_ref * 1e3, 2)})
    print(out["rows"][-1], flush=True)
path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r2_reference_python_cpu.json")
json.dump(out, open(path, "w"), indent=1)
