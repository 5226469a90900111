#!/usr/bin/env python3
"""Runs a few MatD ops once each on an n x n FP64 matrix (ncu target: `ncu ... python tools/profile_ops.py`)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import uni_b200 as u

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
ops = sys.argv[2].split(",") if len(sys.argv) > 2 else ["sigmoid", "softmax1", "softmax0", "var0", "var1", "sum0", "matmul"]
u.init(0)
m = u.MatD.randn(n, n, seed=42)
for _ in range(2):
    if "sigmoid" in ops: m.sigmoid()
    if "softmax1" in ops: m.softmax(1)
    if "softmax0" in ops: m.softmax(0)
    if "var0" in ops: m.var(0)
    if "var1" in ops: m.var(1)
    if "sum0" in ops: m.sum(0)
    if "sub" in ops: m - m.mean(0)
if "matmul" in ops:
    a = u.MatD.randn(4096, 4096, seed=1); b = u.MatD.randn(4096, 4096, seed=2)
    for _ in range(2):
        a @ b
u.sync()
print("ok")
