"""GPU parity for the device generator (csrc/rng.cu) through the C ABI: PCG64 draws,
uniform / normal transforms and the ziggurat are compared bit for bit with the committed NumPy vectors
(tests/golden/rng_numpy/), with NumPy itself on longer streams (NumPy's generator is what the reference's
NumPyRNG.scala is pinned to), and randint with the oracle's restatement of nextBoundedInt.

Tolerance: bit-exact everywhere except ziggurat TAIL draws (|z| > 3.6541528853610088, 2.7e-4 of the stream), whose
magnitude goes through log1p: CUDA's and the host libm's agree to 1 ulp, asserted as <= 1 ulp."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import rng_oracle as ro

R = 3.6541528853610088


@pytest.fixture(scope="module")
def u():
    import uni_b200
    if uni_b200.device_count() == 0:
        pytest.fail("no CUDA device: the product path has no CPU fallback")
    uni_b200.init(0)
    return uni_b200


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "rng_numpy", "rng_numpy.npz"))


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def assert_normal_stream_equal(got, want):
    """bit-exact outside the tail; tail draws within 1 ulp with the same sign."""
    got, want = np.ravel(got), np.ravel(want)
    assert got.shape == want.shape
    tail = np.abs(want) > R
    assert np.array_equal(bits(got[~tail]), bits(want[~tail])), "non-tail draws differ"
    if tail.any():
        d = np.abs(bits(got[tail]).astype(np.int64) - bits(want[tail]).astype(np.int64))
        assert d.max() <= 1, f"tail draws differ by {d.max()} ulp"
    return int(tail.sum())


@pytest.mark.parametrize("n", [1, 255, 256, 257, 4096, 4097, 100003])
def test_rand_bit_exact_and_state_carries_over(u, n):
    g = u.NumPyRNG(42)
    r = np.random.default_rng(42)
    assert np.array_equal(bits(g.rand(1, n).to_numpy()), bits(r.random(n)).reshape(1, -1))
    # the generator is left where the sequential one would be
    assert g.nextDouble() == r.random()
    assert np.array_equal(bits(g.rand(3, 5).to_numpy()), bits(r.random((3, 5))))


@pytest.mark.parametrize("seed", [0, 1, 42, 1234, 20260814, 2**32 + 5, 2**63 - 1])
def test_golden_random_and_randn(u, gold, seed):
    assert np.array_equal(bits(u.NumPyRNG(seed).rand(1, 256).to_numpy()).ravel(), bits(gold[f"s{seed}_random"]))
    assert_normal_stream_equal(u.NumPyRNG(seed).randn(1, 6000).to_numpy(), gold[f"s{seed}_randn"])


@pytest.mark.parametrize("seed", [0, 1, 42, 1234, 20260814, 2**32 + 5, 2**63 - 1])
def test_golden_mixed_calls_share_one_stream(u, gold, seed):
    g = u.NumPyRNG(seed)
    got = np.concatenate([g.randn(777, 1).to_numpy().ravel(), g.rand(10, 10).to_numpy().ravel(),
                          g.uniform_mat(-2.5, 4.0, 5, 10).to_numpy().ravel(), g.normal(3.0, 0.5, 333, 1).to_numpy().ravel(),
                          g.randn(5).to_numpy().ravel()])
    want = gold[f"s{seed}_mix"]
    # standard_normal / normal parts tolerate 1 ulp on tail draws only; the uniform parts are exact
    assert np.array_equal(bits(got[777:927]), bits(want[777:927]))
    assert_normal_stream_equal(got[:777], want[:777])
    assert_normal_stream_equal(got[1260:], want[1260:])
    z = (want[927:1260] - 3.0) / 0.5
    ok = np.abs(z) <= R - 1e-6
    assert np.array_equal(bits(got[927:1260][ok]), bits(want[927:1260][ok]))
    assert np.allclose(got[927:1260], want[927:1260], rtol=1e-15, atol=0)


@pytest.mark.parametrize("n", [1, 2, 7, 4000, 4095, 4096, 4097, 8192, 12289, 100003, 1 << 20])
def test_randn_matches_numpy_stream(u, n):
    g = u.NumPyRNG(7)
    r = np.random.default_rng(7)
    assert_normal_stream_equal(g.randn(1, n).to_numpy(), r.standard_normal(n))
    # the data-dependent number of draws consumed is tracked exactly: both generators continue in step
    assert g.nextLong() & ro.M64 == int(r.bit_generator.random_raw())
    assert_normal_stream_equal(g.randn(33, 3).to_numpy(), r.standard_normal((33, 3)))
    assert g.nextDouble() == r.random()


def test_randn_long_stream_every_segment_boundary(u, gold):
    n = 1 << 22
    z = u.NumPyRNG(1234).randn(1, n).to_numpy().ravel()
    assert np.array_equal(bits(z[::4099]), bits(gold["stride_val_seed1234"]))
    idx = gold["tail_idx_seed1234"]
    assert np.array_equal(np.nonzero(np.abs(z) > R)[0], idx)
    d = np.abs(bits(z[idx]).astype(np.int64) - bits(gold["tail_val_seed1234"]).astype(np.int64))
    assert d.max() <= 1
    want = np.random.default_rng(1234).standard_normal(n)
    ntail = assert_normal_stream_equal(z, want)
    assert 800 < ntail < 1500


def test_randn_64M_draws(u):
    n = 1 << 26
    g = u.NumPyRNG(99)
    r = np.random.default_rng(99)
    assert_normal_stream_equal(g.randn(1, n).to_numpy(), r.standard_normal(n))
    assert g.nextLong() & ro.M64 == int(r.bit_generator.random_raw())


@pytest.mark.parametrize("n", [1, 4096, 4097, 100003, 1 << 22, (1 << 25) + 12345])
def test_randn_lengths_straddling_segments(u, n):
    """Stream lengths on either side of the 4096-draw segment and of the prefix levels: values, a following affine
    fill and the generator state afterwards all equal NumPy's."""
    g = u.NumPyRNG(2024)
    r = np.random.default_rng(2024)
    assert_normal_stream_equal(g.randn(1, n).to_numpy(), r.standard_normal(n))
    got, want = g.normal(1.5, 0.25, 7, 11).to_numpy().ravel(), r.normal(1.5, 0.25, 77)
    ok = np.abs((want - 1.5) / 0.25) <= R - 1e-6
    assert np.array_equal(bits(got[ok]), bits(want[ok])) and np.allclose(got, want, rtol=1e-15, atol=0)
    assert g.nextLong() & ro.M64 == int(r.bit_generator.random_raw())


def test_randn_one_call_of_2_30_draws_equals_chunked_calls(u):
    """Size-independent property at BASELINE scale (32768^2 = 2^30 draws, more than 128 prefix blocks): one fill of
    the whole matrix holds exactly the numbers of eight consecutive fills of its row blocks (each of a size the
    NumPy comparisons above cover), and both generators end in the same state."""
    import ctypes as C
    rows, cols = 32768, 32768
    g1 = u.NumPyRNG(31)
    a = g1.randn(rows, cols)
    g2 = u.NumPyRNG(31)
    b = u.Mat._alloc(rows, cols, u._lib.F64)
    step = rows // 8
    for r0 in range(0, rows, step):
        u._lib.check(u._lib.lib.um_rng_randn(C.byref(g2._g), b._raw_slice(r0, r0 + step, 0, cols)._ref()))
    assert g1.state == g2.state
    assert a.ne(b).any() is False
    del a, b
    a = u.NumPyRNG(32).rand(rows, cols)
    g3, b = u.NumPyRNG(32), u.Mat._alloc(rows, cols, u._lib.F64)
    for r0 in range(0, rows, step):
        u._lib.check(u._lib.lib.um_rng_rand(C.byref(g3._g), b._raw_slice(r0, r0 + step, 0, cols)._ref()))
    assert a.ne(b).any() is False


def test_uniform_and_normal_transforms(u):
    g = u.NumPyRNG(3)
    r = np.random.default_rng(3)
    assert np.array_equal(bits(g.uniform_mat(-7.25, 19.5, 37, 41).to_numpy()), bits(r.uniform(-7.25, 19.5, (37, 41))))
    got = g.normal(-1.5, 2.25, 101, 97).to_numpy()
    want = r.normal(-1.5, 2.25, (101, 97))
    ok = np.abs((want + 1.5) / 2.25) <= R - 1e-6
    assert np.array_equal(bits(got[ok]), bits(want[ok]))
    assert np.allclose(got, want, rtol=1e-15, atol=0)


def test_randint_matches_reference_bounded_int(u):
    g = u.NumPyRNG(11)
    o = ro.NumPyRNG(11)
    a = g.randint(0, 1000, 3, 5)            # 15 ints: 8 draws, the last high half stays cached
    assert a.dtype == u._lib.I32 and np.array_equal(a.to_numpy(), o.randint_mat(0, 1000, 3, 5))
    assert np.array_equal(g.randint(5, 17, 2, 2).to_numpy(), o.randint_mat(5, 17, 2, 2))   # cached half goes first
    assert np.array_equal(g.randint(-50, 50, 1, 1).to_numpy(), o.randint_mat(-50, 50, 1, 1))
    assert np.array_equal(g.randint(-50, 50, 1, 1).to_numpy(), o.randint_mat(-50, 50, 1, 1))
    assert np.array_equal(g.randint(0, 2**31 - 1, 257, 33).to_numpy(), o.randint_mat(0, 2**31 - 1, 257, 33))
    assert g.nextDouble() == o.nextDouble()
    with pytest.raises(ValueError, match="bound must be positive"):
        g.randint(5, 5, 2, 2)


def test_matd_factories_use_the_global_generator(u):
    u.MatD.setSeed(42)
    r = np.random.default_rng(42)
    assert_normal_stream_equal(u.MatD.randn(3, 4).to_numpy(), r.standard_normal((3, 4)))
    v = u.MatD.randn(5)
    assert v.shape == (5, 1)
    assert_normal_stream_equal(v.to_numpy(), r.standard_normal((5, 1)))
    assert np.array_equal(bits(u.MatD.rand(4, 6).to_numpy()), bits(r.random((4, 6))))
    assert u.MatD.nextRandDouble() == r.random()
    assert np.array_equal(bits(u.MatD.uniform(1.0, 2.0, 2, 3).to_numpy()), bits(r.uniform(1.0, 2.0, (2, 3))))
    k = u.MatD.randint(0, 10)
    assert 0 <= k < 10
    # empty and degenerate shapes
    assert u.MatD.randn(0, 5).shape == (0, 5)
    assert u.MatD.rand(7, 0).shape == (7, 0)


def test_fills_into_contiguous_row_blocks_of_a_larger_matrix(u):
    """The C ABI fills any contiguous view: consecutive row blocks of one matrix filled by consecutive calls hold the
    same numbers as one fill of the whole matrix (how bench.py builds the config-4 panel in 1 GiB chunks)."""
    import ctypes as C
    rows, cols = 37, 501
    whole = u.NumPyRNG(12).randn(rows, cols).to_numpy()
    g = u.NumPyRNG(12)
    m = u.MatD.zeros(rows, cols)
    for r0, r1 in ((0, 5), (5, 6), (6, 30), (30, 37)):
        u._lib.check(u._lib.lib.um_rng_randn(C.byref(g._g), m._raw_slice(r0, r1, 0, cols)._ref()))
    assert np.array_equal(bits(m.to_numpy()), bits(whole))
    g2, m2 = u.NumPyRNG(12), u.MatD.zeros(rows, cols)
    for r0, r1 in ((0, 20), (20, 37)):
        u._lib.check(u._lib.lib.um_rng_rand(C.byref(g2._g), m2._raw_slice(r0, r1, 0, cols)._ref()))
    assert np.array_equal(bits(m2.to_numpy()), bits(np.random.default_rng(12).random((rows, cols))))


def test_fills_reject_strided_destinations(u):
    import ctypes as C
    g = u.NumPyRNG(0)
    m = u.MatD.zeros(8, 8)
    with pytest.raises(ValueError, match="contiguous"):
        u._lib.check(u._lib.lib.um_rng_rand(C.byref(g._g), m.T._ref()))
    with pytest.raises(ValueError, match="dtype"):
        u._lib.check(u._lib.lib.um_rng_randint(C.byref(g._g), 0, 5, m._ref()))
