"""The C++ restatement of the reference's pure-JVM kernels (oracle/jvm_restated.cpp, a CPU BASELINE for bench.py) against
the NumPy oracle that is pinned to the reference's fixtures: same association orders, so sums are bit-identical."""
import numpy as np
import pytest

from oracle import jvm_restated as jr
from oracle import mat_oracle as mo


@pytest.mark.parametrize("shape", [(1, 1), (3, 5), (17, 33), (64, 64), (130, 70), (1031, 517)])
def test_restated_kernels_match_the_oracle(shape):
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    a = rng.standard_normal(shape)
    om = mo.from2d(a)
    for th in (1, 4):
        assert jr.total(a, th) == mo.mat_sum(om)                                          # sumD: bit-exact association
        assert np.array_equal(jr.col_sums(a, th), mo.sum_axis(om, 0).to2d().ravel())
        assert np.array_equal(jr.row_sums(a, th), mo.sum_axis(om, 1).to2d().ravel())
        v = rng.standard_normal(shape[1])
        assert np.array_equal(jr.sub_rowvec(a, v, np.empty(shape), th), a - v)
        assert np.array_equal(jr.add_same(a, a[::-1].copy(), np.empty(shape), th), a + a[::-1])
        assert np.allclose(jr.sigmoid(a, np.empty(shape), th), mo.sigmoid(om).to2d(), rtol=1e-15, atol=0)
    assert np.allclose(jr.softmax_rows(a, np.empty(shape)), mo.softmax(om, 1).to2d(), rtol=1e-14, atol=0)


@pytest.mark.parametrize("mkn", [(4, 16, 4), (16, 16, 16), (33, 40, 21), (70, 130, 50), (257, 64, 129)])
def test_restated_matmul_is_the_sequential_k_contract(mkn):
    """matmulPure's contract: c_ij = sum_k a_ik b_kj accumulated from 0.0 in increasing k (MatmulPure.scala:9-15)."""
    m, k, n = mkn
    rng = np.random.default_rng(m + k + n)
    a, b = rng.standard_normal((m, k)), rng.standard_normal((k, n))
    want = np.zeros((m, n))
    for kk in range(k):
        want += a[:, kk:kk + 1] * b[kk:kk + 1, :]
    for th in (1, 3):
        assert np.array_equal(jr.matmul_pure(a, b, th), want)
