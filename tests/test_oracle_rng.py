"""CPU: the generator oracle (oracle/rng_oracle.py) and the host half of the C ABI (um_rng_seed / um_rng_advance /
um_rng_next_u64 -- pure host arithmetic, no device) against the committed NumPy vectors (tests/golden/rng_numpy/,
generator script beside it).  NumPy's PCG64 / default_rng is what the reference's NumPyRNG.scala is pinned to."""
import os

import numpy as np
import pytest

from oracle import rng_oracle as ro

SEEDS = [0, 1, 42, 1234, 20260814, 2**32 + 5, 2**63 - 1]


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "rng_numpy", "rng_numpy.npz"))


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


@pytest.mark.parametrize("seed", SEEDS)
def test_oracle_seeding_and_raw_stream(gold, seed):
    g = ro.NumPyRNG(seed)
    st = gold[f"s{seed}_state"]
    assert (g.state >> 64, g.state & ro.M64, g.inc >> 64, g.inc & ro.M64) == tuple(int(v) for v in st)
    assert [g.nextLong() for _ in range(64)] == [int(v) for v in gold[f"s{seed}_raw"]]


@pytest.mark.parametrize("seed", SEEDS)
def test_oracle_random_and_ziggurat_bit_exact(gold, seed):
    g = ro.NumPyRNG(seed)
    assert np.array_equal(bits(g.rand_mat(1, 256)), bits(gold[f"s{seed}_random"]).reshape(1, -1))
    g = ro.NumPyRNG(seed)
    assert np.array_equal(bits(g.randn_mat(1, 6000)), bits(gold[f"s{seed}_randn"]).reshape(1, -1))


@pytest.mark.parametrize("seed", SEEDS[:4])
def test_oracle_mixed_calls_share_one_stream(gold, seed):
    g = ro.NumPyRNG(seed)
    got = np.concatenate([g.randn_mat(1, 777).ravel(), g.rand_mat(1, 100).ravel(), g.uniform_mat(-2.5, 4.0, 1, 50).ravel(),
                          g.normal_mat(3.0, 0.5, 1, 333).ravel(), g.randn_mat(1, 5).ravel()])
    assert np.array_equal(bits(got), bits(gold[f"s{seed}_mix"]))


def test_oracle_bounded_int_halves():
    # nextBoundedInt (NumPyRNG.scala:82-110): low half of a draw first, then its high half, (bits32 * bound) >>> 32
    g, h = ro.NumPyRNG(5), ro.NumPyRNG(5)
    raws = [h.nextLong() for _ in range(8)]
    want = []
    for r in raws:
        want += [((r & 0xFFFFFFFF) * 1000) >> 32, ((r >> 32) * 1000) >> 32]
    assert [g.nextBoundedInt(1000) for _ in range(16)] == want
    with pytest.raises(ValueError):
        g.nextBoundedInt(0)
    with pytest.raises(ValueError):
        ro.NumPyRNG(-3)


@pytest.mark.parametrize("seed", SEEDS)
def test_abi_host_generator_matches_golden(gold, seed):
    import uni_b200 as u
    g = u.NumPyRNG(seed)
    st = gold[f"s{seed}_state"]
    assert g.state == ((int(st[0]) << 64) | int(st[1]), (int(st[2]) << 64) | int(st[3]))
    assert [g.nextLong() & ro.M64 for _ in range(64)] == [int(v) for v in gold[f"s{seed}_raw"]]
    g = u.NumPyRNG(seed)
    assert [g.nextDouble() for _ in range(256)] == list(gold[f"s{seed}_random"])
    # jump-ahead == stepping
    a, b = u.NumPyRNG(seed), u.NumPyRNG(seed)
    for _ in range(1237):
        a.nextLong()
    b.advance(1237)
    assert a.state == b.state


def test_abi_negative_seed_is_an_argument_error():
    import uni_b200 as u
    with pytest.raises(ValueError, match="non-negative"):
        u.NumPyRNG(-1)


def test_ziggurat_tables_shape():
    ki, wi, fi = ro.ziggurat_tables()
    assert ki[1] == 0 and fi[0] == 1.0 and abs(wi[255] * 2.0**52 - ro.ZIG_R) < 1e-12
    assert abs(1.0 / ro.ZIG_R - ro.ZIG_INV_R) < 1e-16


def test_abi_jump_ahead_is_additive_and_matches_numpy_advance():
    """um_rng_advance (the host twin of the per-thread jumps the kernels take): advance(a) then advance(b) equals
    advance(a + b) equals NumPy's PCG64.advance, for deltas spanning every bit of a 64-bit count."""
    import uni_b200 as u
    rng = np.random.default_rng(77)
    for seed in (0, 5, 2**40 + 1):
        for _ in range(20):
            a = int(rng.integers(0, 2**62)) >> int(rng.integers(0, 62))
            b = int(rng.integers(0, 2**62)) >> int(rng.integers(0, 62))
            g1, g2 = u.NumPyRNG(seed), u.NumPyRNG(seed)
            g1.advance(a); g1.advance(b)
            g2.advance(a + b)
            ref = np.random.PCG64(seed); ref.advance(a + b)
            assert g1.state == g2.state == (ref.state["state"]["state"], ref.state["state"]["inc"])
            assert g1.nextLong() & ro.M64 == int(ref.random_raw())
