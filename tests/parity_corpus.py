"""Rebuilds the inputs of the reference's mat-parity fixture (apps/MatParityGen.scala).

corpus():    NumPyRNG(20260814), x(i) = uniform(-1e6,1e6) if i even else uniform(-1e-6,1e-6)
             (MatParityGen.scala:176-180)
corpus_exp(): NumPyRNG(20260815).uniform(-5.0, 0.5) (MatParityGen.scala:189-191)
NumPyRNG is bit-identical to numpy.random.default_rng (src/main/scala/uni/data/NumPyRNG.scala:22-140).
"""
import functools

import numpy as np

SEED = 20260814
SIZES = [0, 1, 7, 8, 9, 4095, 4096, 4097, 65535, 65536, 65537, 98304, 200000]
SHAPES = [(1, 1), (3, 5), (5, 3), (8, 9), (9, 8), (64, 64), (65, 63), (256, 256), (257, 256),
          (300, 400), (400, 300)]
NEG_NAN = np.array([0xFFF8000000000000], dtype=np.uint64).view(np.float64)[0]
ADVERSARIAL = {
    "nanmid": [1.0, np.nan, -1.0, 2.0],
    "nanfirst": [np.nan, 1.0, -1.0],
    "nanonly": [np.nan, np.nan],
    "negnan": [NEG_NAN, 1.0, -1.0],
    "zeros": [0.0, -0.0],
    "zerosrev": [-0.0, 0.0],
    "zerosmix": [1.0, -0.0, 0.0, -1.0],
    "infs": [np.inf, -np.inf, 0.0],
    "infnan": [np.inf, np.nan, -np.inf],
    "ties": [3.0, 1.0, 1.0, 3.0],
}


@functools.lru_cache(maxsize=None)
def corpus(n=200000):
    u = np.random.default_rng(SEED).random(n)
    lo = np.where(np.arange(n) % 2 == 0, -1e6, -1e-6)
    hi = -lo
    return lo + (hi - lo) * u


@functools.lru_cache(maxsize=None)
def corpus_exp(n=200000):
    u = np.random.default_rng(SEED + 1).random(n)
    return -5.0 + (0.5 - (-5.0)) * u


def load_reference(path):
    """-> dict[(shape_label, case)] = int word"""
    ref = {}
    with open(path) as fh:
        for line in fh:
            if line.startswith("#") or not line.strip():
                continue
            shape, case, word = line.split()
            ref[(shape, case)] = int(word, 16)
    return ref
