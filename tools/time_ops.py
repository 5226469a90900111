#!/usr/bin/env python3
"""CUDA-event timings (min of 7) of chosen MatD ops on an n x n FP64 matrix: python tools/time_ops.py [n] [op,op,...]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uni_b200 as u
from uni_b200._lib import lib, check

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
ops = sys.argv[2].split(",") if len(sys.argv) > 2 else ["softmax1", "softmax0", "sigmoid"]
u.init(0)
m = u.MatD.randn(n, n, seed=42)
E = float(n) * n
mask = m.gt(0.0)
table = {"softmax1": (lambda: m.softmax(1), 16), "softmax0": (lambda: m.softmax(0), 16), "sigmoid": (lambda: m.sigmoid(), 16),
         "cumsum0": (lambda: m.cumsum(0), 16), "cumsum1": (lambda: m.cumsum(1), 16), "cummax0": (lambda: m.cummax(0), 16),
         "cummax1": (lambda: m.cummax(1), 16), "where": (lambda: u.Mat.where(mask, m, m), 17), "select": (lambda: m[mask], 13),
         "var0": (lambda: m.var(0), 8), "var1": (lambda: m.var(1), 8), "sum0": (lambda: m.sum(0), 8), "relu": (lambda: m.relu(), 16)}
for op in ops:
    fn, bpe = table[op]
    for _ in range(3):
        fn()
    best = 1e30
    ms = C.c_float(0)
    for _ in range(7):
        check(lib.um_timer_start()); fn(); check(lib.um_timer_stop(C.byref(ms)))
        best = min(best, ms.value)
    print(f"{op:10s} {best:8.4f} ms  {bpe * E / best / 1e6:8.1f} GB/s  ({bpe * E / best / 1e6 / 6532.2:.3f} of 6532)")
