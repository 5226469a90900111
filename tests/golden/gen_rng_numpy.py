"""Golden vectors for the NumPy-compatible generator: outputs of numpy.random (PCG64 / default_rng) -- the library
the reference's NumPyRNG.scala is pinned to -- generated in the build container (NumPy 2.3.5) and committed, so the
tests do not depend on the NumPy version installed where they run.

    python tests/golden/gen_rng_numpy.py
"""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SEEDS = [0, 1, 42, 1234, 20260814, 2**32 + 5, 2**63 - 1]


def main():
    out = {}
    for seed in SEEDS:
        st = np.random.PCG64(seed).state["state"]
        out[f"s{seed}_state"] = np.array([st["state"] >> 64, st["state"] & (2**64 - 1), st["inc"] >> 64, st["inc"] & (2**64 - 1)], dtype=np.uint64)
        out[f"s{seed}_raw"] = np.random.PCG64(seed).random_raw(64)
        out[f"s{seed}_random"] = np.random.default_rng(seed).random(256)
        out[f"s{seed}_randn"] = np.random.default_rng(seed).standard_normal(6000)
        g = np.random.default_rng(seed)  # one generator, mixed calls: state carries over
        out[f"s{seed}_mix"] = np.concatenate([g.standard_normal(777), g.random(100), g.uniform(-2.5, 4.0, 50), g.normal(3.0, 0.5, 333), g.standard_normal(5)])
    # the draws that reach the ziggurat tail (|z| > R) in a long stream: index and value
    z = np.random.default_rng(1234).standard_normal(1 << 22)
    idx = np.nonzero(np.abs(z) > 3.6541528853610088)[0]
    out["tail_idx_seed1234"] = idx.astype(np.int64)
    out["tail_val_seed1234"] = z[idx]
    # a sparse sample of the same stream, and its 64-bit checksum
    out["stride_val_seed1234"] = z[::4099].copy()
    out["xor_seed1234"] = np.array([np.bitwise_xor.reduce(z.view(np.uint64))], dtype=np.uint64)
    out["sum_seed1234"] = np.array([int(z.view(np.uint64).astype(object).sum()) & (2**64 - 1)], dtype=np.uint64)
    np.savez_compressed(os.path.join(HERE, "rng_numpy", "rng_numpy.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
