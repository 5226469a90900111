"""GPU parity tests for the Mat operator path: CUDA (through the C ABI / uni_b200 mirror) vs the
CPU oracle and the reference's bit-pattern fixture.  Bit-exact for masks, arg*/idx*, min/max and
exact elementwise ops; FP sums within rel 1e-12 scaled by sum|x|."""
import math
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import mat_oracle as mo
from tests import parity_corpus as pc


@pytest.fixture(scope="module")
def u():
    import uni_b200
    if uni_b200.device_count() == 0:
        pytest.fail("no CUDA device: the product path has no CPU fallback")
    uni_b200.init(0)
    return uni_b200


@pytest.fixture(scope="module")
def ref(golden_dir):
    return pc.load_reference(os.path.join(golden_dir, "mat-parity", "scala-reference.txt"))


def dev(u, a):
    return u.Mat.from_numpy(np.ascontiguousarray(a))


def omat(a):
    return mo.from2d(np.ascontiguousarray(a))


def sum_tol(a, scale=None):
    s = np.abs(np.asarray(a, dtype=np.float64)).sum() if scale is None else scale
    return 1e-12 * max(float(s), 1e-300)


# ---------------------------------------------------------------- fixture rows, bit for bit
@pytest.mark.parametrize("shape", pc.SHAPES)
def test_fixture_exact_rows(u, ref, shape):
    r, c = shape
    lab = f"{r}x{c}"
    a = pc.corpus()[: r * c].reshape(r, c)
    m = dev(u, a)
    mean0 = dev(u, mo.mean_axis(omat(a), 0).to2d())   # reference's own mean vectors: the broadcast op is exact
    mean1 = dev(u, mo.mean_axis(omat(a), 1).to2d())
    am, an = m.argmax(), m.argmin()
    got = {
        "negfnv": mo.fnv((-m).to_numpy()),
        "divfnv": mo.fnv((m / (m.abs() + 1.0)).to_numpy()),
        "bcast0fnv": mo.fnv((m - mean0).to_numpy()),
        "bcast1fnv": mo.fnv((m - mean1).to_numpy()),
        "tfnv": mo.fnv(m.T.to_numpy()),
        "slicefnv": mo.fnv(m.slice(0, r, 1, c).to_numpy()),
        "min0fnv": mo.fnv(m.min(0).to_numpy()),
        "max1fnv": mo.fnv(m.max(1).to_numpy()),
        "argmax": am[0] * c + am[1],
        "argmin": an[0] * c + an[1],
        "pd.idxmin0": mo.fnv_words(m.idxmin(0).to_numpy().reshape(-1).tolist()),
        "pd.idxmax1": mo.fnv_words(m.idxmax(1).to_numpy().reshape(-1).tolist()),
    }
    me = dev(u, pc.corpus_exp()[: r * c].reshape(r, c))
    got["math.relu"] = mo.fnv(me.relu().to_numpy())
    bad = {k: (hex(v), hex(ref[(lab, k)])) for k, v in got.items() if v != ref[(lab, k)]}
    assert not bad, bad
    # transcendental rows on their quantised grids: allow the grid-straddle the fixture documents
    sig = me.sigmoid().to_numpy()
    np.testing.assert_allclose(sig, mo.sigmoid(omat(pc.corpus_exp()[: r * c].reshape(r, c))).to2d(), rtol=1e-12, atol=0)


@pytest.mark.parametrize("shape", pc.SHAPES)
def test_fixture_sum_rows_tolerance(u, ref, shape):
    r, c = shape
    lab = f"{r}x{c}"
    a = pc.corpus()[: r * c].reshape(r, c)
    m = dev(u, a)
    tol = sum_tol(a)
    f64 = lambda w: np.array([w], dtype=np.uint64).view(np.float64)[0]
    assert abs(m.T.sum() - f64(ref[(lab, "tsum")])) <= tol
    assert abs(m.sum() - f64(ref[(lab, "tsum")])) <= tol
    assert abs(m.T.mean() - f64(ref[(lab, "tmean")])) <= tol / (r * c)
    for got, key in ((m.std(), "std"), (m.T.std(), "tstd"), (m.variance(), "var")):
        want = f64(ref[(lab, key)])
        assert abs(got - want) <= 1e-12 * max(abs(want), 1e-300), key
    assert abs(m.slice(0, r, 1, c).sum() - f64(ref[(lab, "slicesum")])) <= tol


@pytest.mark.parametrize("n", pc.SIZES)
def test_fixture_sizes(u, ref, n):
    a = pc.corpus()[:n]
    f64 = lambda w: np.array([w], dtype=np.uint64).view(np.float64)[0]
    if n == 0:
        m = u.MatD.zeros(0, 1)
        assert m.sum() == 0.0 and m.mean() == 0.0          # Mat.scala:3470
        with pytest.raises(u.EmptyMatrixError):
            m.min()
        return
    m = dev(u, a.reshape(-1, 1))
    assert abs(m.sum() - f64(ref[(str(n), "sum")])) <= sum_tol(a)
    assert abs(m.mean() - f64(ref[(str(n), "mean")])) <= sum_tol(a) / n
    me = dev(u, pc.corpus_exp()[:n].reshape(-1, 1))
    assert mo.bits(me.min()) == ref[(str(n), "min")]
    assert mo.bits(me.max()) == ref[(str(n), "max")]


@pytest.mark.parametrize("name", sorted(pc.ADVERSARIAL))
def test_adversarial_rows_bit_exact(u, ref, name):
    arr = np.array(pc.ADVERSARIAL[name], dtype=np.float64)
    mats = {"col": dev(u, arr.reshape(-1, 1)), "row": dev(u, arr.reshape(1, -1)), "colT": dev(u, arr.reshape(-1, 1)).T}
    for orient, m in mats.items():
        lab = f"adv/{name}/{orient}"
        pos, fin = m.gt(0.0), m.isfinite()
        an, ax = m.argmin(), m.argmax()
        fm = lambda x: mo.fnv_mask(x.to_numpy())
        got = {
            "min": mo.canon(m.min()), "max": mo.canon(m.max()),
            "argmin": an[0] * m.cols + an[1], "argmax": ax[0] * m.cols + ax[1],
            "min0fnv": mo.fnv_canon(m.min(0).to_numpy()), "max1fnv": mo.fnv_canon(m.max(1).to_numpy()),
            "mask.gt0": fm(pos), "mask.lt0": fm(m.lt(0.0)), "mask.gte0": fm(m.gte(0.0)), "mask.lte0": fm(m.lte(0.0)),
            "mask.eq0": fm(m.eq(0.0)), "mask.ne0": fm(m.ne(0.0)), "mask.eqnan": fm(m.eq(np.nan)),
            "mask.nenan": fm(m.ne(np.nan)), "mask.gtinf": fm(m.gt(np.inf)), "mask.gteinf": fm(m.gte(np.inf)),
            "mask.ltninf": fm(m.lt(-np.inf)), "mask.isnan": fm(m.isnan()), "mask.isinf": fm(m.isinf()),
            "mask.isfinite": fm(fin), "mask.and": fm(pos & fin), "mask.or": fm(pos | fin), "mask.not": fm(~pos),
            "mask.all": int(pos.all()), "mask.any": int(pos.any()),
            "mask.all0": fm(pos.all(0)), "mask.any0": fm(pos.any(0)), "mask.all1": fm(pos.all(1)), "mask.any1": fm(pos.any(1)),
        }
        s = m.sum()
        got["sum"] = mo.canon(s)
        bad = {k: (hex(v), hex(ref[(lab, k)])) for k, v in got.items() if v != ref[(lab, k)]}
        assert not bad, (lab, bad)


# ---------------------------------------------------------------- running folds, m(mask), Mat.where
@pytest.mark.parametrize("n", [k for k in pc.SIZES if k])
def test_cumulative_fixture_sizes_bit_exact(u, ref, n):
    """cumlast / cummaxfnv rows (MatParityGen.scala:199,220): the flat cumsum is the reference's sequential fold."""
    m = dev(u, pc.corpus()[:n].reshape(-1, 1))
    flat = m.cumsum()
    assert flat.shape == (1, n)
    assert mo.bits(float(flat.to_numpy()[0, -1])) == ref[(str(n), "cumlast")]
    me = dev(u, pc.corpus_exp()[:n].reshape(-1, 1))
    assert mo.fnv(me.cummax(0).to_numpy().reshape(-1)) == ref[(str(n), "cummaxfnv")]


@pytest.mark.parametrize("shape", pc.SHAPES)
def test_cumulative_fixture_2d_bit_exact(u, ref, shape):
    r, c = shape
    lab = f"{r}x{c}"
    a = pc.corpus()[: r * c].reshape(r, c)
    m = dev(u, a)
    f = lambda x: mo.fnv(x.to_numpy().reshape(-1))
    got = {"cum0fnv": f(m.cumsum(0)), "cum1fnv": f(m.cumsum(1)), "cmax0fnv": f(m.cummax(0)), "cmin1fnv": f(m.cummin(1))}
    bad = {k: (hex(v), hex(ref[(lab, k)])) for k, v in got.items() if v != ref[(lab, k)]}
    assert not bad, bad
    # the same folds through a transposed view (strided reads, contiguous result): bit-equal to the oracle
    o = omat(a)
    for ax in (0, 1):
        assert np.array_equal(m.T.cumsum(ax).to_numpy(), mo.cumsum(o.T, ax).to2d())
        assert np.array_equal(m.T.cummin(ax).to_numpy(), mo.cummin(o.T, ax).to2d())
    assert np.array_equal(m.T.cumsum().to_numpy(), mo.cumsum(o.T).to2d())


@pytest.mark.parametrize("name", sorted(pc.ADVERSARIAL))
def test_adversarial_cumulative_and_selection_bit_exact(u, ref, name):
    arr = np.array(pc.ADVERSARIAL[name], dtype=np.float64)
    mats = {"col": dev(u, arr.reshape(-1, 1)), "row": dev(u, arr.reshape(1, -1)), "colT": dev(u, arr.reshape(-1, 1)).T}
    fc = lambda x: mo.fnv_canon(x.to_numpy().reshape(-1))
    for orient, m in mats.items():
        lab = f"adv/{name}/{orient}"
        pos = m.gt(0.0)
        sel = m[pos]
        got = {
            "cmax0fnv": fc(m.cummax(0)), "cmax1fnv": fc(m.cummax(1)), "cmin0fnv": fc(m.cummin(0)), "cmin1fnv": fc(m.cummin(1)),
            "mask.selfnv": fc(sel), "mask.selshape": sel.rows * 1000 + sel.cols,
            "mask.wherefnv": fc(u.Mat.where(pos, m, m * -1.0)), "mask.wheresfnv": fc(u.Mat.where(pos, 1.0, -1.0)),
        }
        bad = {k: (hex(v), hex(ref[(lab, k)])) for k, v in got.items() if v != ref[(lab, k)]}
        assert not bad, (lab, bad)


@pytest.mark.parametrize("shape", [(1, 1), (1, 70), (70, 1), (33, 31), (129, 257), (1000, 37), (37, 1000), (3, 5000), (2100, 130)])
def test_cumulative_vs_oracle(u, shape):
    r, c = shape
    rng = np.random.default_rng(r * 31 + c)
    a = rng.standard_normal((r, c))
    a[rng.random((r, c)) < 0.02] = np.nan
    a[rng.random((r, c)) < 0.02] = -0.0
    a[rng.random((r, c)) < 0.02] = 0.0
    a[rng.random((r, c)) < 0.01] = np.inf
    m, o = dev(u, a), omat(a)
    same = lambda x, y: np.array_equal(x.view(np.uint64), y.view(np.uint64))
    for ax in (0, 1):
        assert same(m.cummax(ax).to_numpy(), mo.cummax(o, ax).to2d())
        assert same(m.cummin(ax).to_numpy(), mo.cummin(o, ax).to2d())
    b = rng.standard_normal((r, c))
    mb, ob = dev(u, b), omat(b)
    for ax in (0, 1, None):
        assert same(mb.cumsum(ax).to_numpy(), mo.cumsum(ob, ax).to2d())
    if r > 2 and c > 3:   # offset slice: the pitch differs from the width
        assert same(mb.slice(1, r, 2, c).cumsum(0).to_numpy(), mo.cumsum(mo.slice_(ob, 1, r, 2, c), 0).to2d())
        assert same(mb.slice(1, r, 2, c).cumsum(1).to_numpy(), mo.cumsum(mo.slice_(ob, 1, r, 2, c), 1).to2d())
        assert same(mb.slice(1, r, 2, c).cumsum().to_numpy(), mo.cumsum(mo.slice_(ob, 1, r, 2, c)).to2d())
    f = b.astype(np.float32)
    mf = dev(u, f)
    assert mf.cumsum(1).dtype == 1
    assert np.array_equal(mf.cumsum(1).to_numpy(), np.cumsum(f, axis=1, dtype=np.float32))
    assert np.array_equal(mf.cumsum(0).to_numpy(), np.cumsum(f, axis=0, dtype=np.float32))
    assert np.array_equal(mf.cummax(0).to_numpy(), np.maximum.accumulate(f, axis=0))


def test_cumulative_edge_cases(u):
    with pytest.raises(ValueError):
        u.MatD.zeros(3, 3).cumsum(2)                       # require(axis == 0 || axis == 1), Mat.scala:3636
    with pytest.raises(ValueError):
        u.MatD.zeros(0, 3).cummax(0)                       # m(0, j) on an empty lane, Mat.scala:3694-3699
    assert u.MatD.zeros(0, 3).cummax(1).shape == (0, 3)
    assert u.MatD.zeros(0, 3).cumsum(0).shape == (0, 3)
    assert u.MatD.zeros(0, 0).cumsum().shape == (1, 0)
    z = dev(u, np.array([[0.0, -0.0, np.nan, 1.0, np.nan, -np.inf]]))
    assert [mo.bits(x) for x in z.cummin(1).to_numpy()[0]] == [mo.bits(x) for x in [0.0, -0.0, -0.0, -0.0, -0.0, -np.inf]]
    bc = dev(u, np.array([[1.0, 2.0, 3.0]])).broadcastTo(4, 3)       # stride-0 rows
    assert np.array_equal(bc.cumsum(0).to_numpy(), np.cumsum(np.tile([[1.0, 2.0, 3.0]], (4, 1)), axis=0))


@pytest.mark.parametrize("shape", [(1, 1), (7, 9), (64, 64), (65, 63), (300, 41), (3, 9000), (5000, 3)])
def test_mask_select_and_where_vs_oracle(u, shape):
    r, c = shape
    rng = np.random.default_rng(r + 1000 * c)
    a = rng.standard_normal((r, c)); b = rng.standard_normal((r, c))
    m, mb = dev(u, a), dev(u, b)
    for thr in (-10.0, -0.5, 0.0, 1.0, 10.0):
        mask = m.gt(thr)
        sel = m[mask]
        want = a[a > thr]
        assert sel.shape == (1, want.size) and np.array_equal(sel.to_numpy().reshape(-1), want)
        # mask and operand through transposed views: row-major order of the VIEW
        selt = m.T[m.T.gt(thr)]
        assert np.array_equal(selt.to_numpy().reshape(-1), a.T[a.T > thr])
        assert np.array_equal(u.Mat.where(mask, m, mb).to_numpy(), np.where(a > thr, a, b))
        assert np.array_equal(u.Mat.where(mask.T, m.T, mb.T).to_numpy(), np.where(a.T > thr, a.T, b.T))
        assert np.array_equal(u.Mat.where(mask, 2.5, -1.0).to_numpy(), np.where(a > thr, 2.5, -1.0))
    f = a.astype(np.float32)
    mf = dev(u, f)
    assert mf[mf.lt(0.0)].dtype == 1 and np.array_equal(mf[mf.lt(0.0)].to_numpy().reshape(-1), f[f < 0])
    other = mb.gt(0.0)
    assert np.array_equal(u.Mat.where(m.gt(0.0), other, ~other).to_numpy(), np.where(a > 0, b > 0, ~(b > 0)))
    with pytest.raises(ValueError):
        m[dev(u, np.zeros((r + 1, c))).gt(0.0)]              # Mat.scala:4010-4011
    with pytest.raises(ValueError):
        u.Mat.where(m.gt(0.0), m, dev(u, np.zeros((r, c + 1))))   # Mat.scala:1306-1310


def test_mask_select_large_is_stable(u):
    """4M elements over many blocks: positions come from an exact integer scan, so the order is row-major."""
    m = u.MatD.randn(2048, 2048, seed=5)
    a = m.to_numpy()
    sel = m[m.gt(0.25)]
    want = a[a > 0.25]
    assert sel.shape == (1, want.size) and np.array_equal(sel.to_numpy().reshape(-1), want)
    c0 = m.cumsum(0).to_numpy()
    assert np.array_equal(c0, np.cumsum(a, axis=0))
    assert np.array_equal(m.cummax(1).to_numpy(), np.maximum.accumulate(a, axis=1))
    # > 8192 blocks of 4096 elements: the block-count scan takes several rounds with a carry
    big = u.MatD.randn(7000, 6000, seed=11)
    ab = big.to_numpy()
    for thr in (1.5, -0.3):
        selb = big[big.gt(thr)]
        wantb = ab[ab > thr]
        assert selb.shape == (1, wantb.size) and np.array_equal(selb.to_numpy().reshape(-1), wantb)


# ---------------------------------------------------------------- elementwise / broadcasting / views
SHAPES_EW = [(1, 1), (1, 7), (7, 1), (3, 5), (64, 64), (65, 63), (257, 129), (1000, 33), (33, 1000), (2, 4099)]


@pytest.mark.parametrize("shape", SHAPES_EW)
@pytest.mark.parametrize("op", ["add", "sub", "mul", "div"])
def test_binary_broadcast_exact(u, shape, op):
    rng = np.random.default_rng(hash((shape, op)) % 2**32)
    r, c = shape
    a = rng.standard_normal((r, c)); b = rng.standard_normal((r, c))
    row = rng.standard_normal((1, c)); col = rng.standard_normal((r, 1))
    f = {"add": lambda x, y: x + y, "sub": lambda x, y: x - y, "mul": lambda x, y: x * y, "div": lambda x, y: x / y}[op]
    A, B, R, Cc = dev(u, a), dev(u, b), dev(u, row), dev(u, col)
    for X, Y, x, y in ((A, B, a, b), (A, R, a, row), (A, Cc, a, col), (R, A, row, a), (Cc, R, col, row), (A, 1.5, a, 1.5)):
        got = f(X, Y).to_numpy()
        want = mo.binary(op, omat(x), omat(y)).to2d() if not np.isscalar(y) else mo.binary_scalar(op, omat(x), y).to2d()
        assert np.array_equal(got, want), (shape, op)
    got = f(2.5, A).to_numpy()
    assert np.array_equal(got, mo.binary_scalar(op, omat(a), 2.5, scalar_left=True).to2d())
    # views: transposed left operand, offset slice, and both at once
    At = dev(u, a.T.copy()).T
    assert np.array_equal(f(At, B).to_numpy(), f(a, b))
    if r > 2 and c > 2:
        big = rng.standard_normal((r + 2, c + 3))
        V = dev(u, big).slice(1, r + 1, 2, c + 2)
        assert np.array_equal(f(V, B).to_numpy(), f(big[1 : r + 1, 2 : c + 2], b))
        assert np.array_equal(f(B, V).to_numpy(), f(b, big[1 : r + 1, 2 : c + 2]))


def test_broadcast_shape_errors(u):
    a = u.MatD.zeros(3, 4); b = u.MatD.zeros(2, 4)
    with pytest.raises(ValueError):
        a + b
    with pytest.raises(ValueError):
        a @ b
    with pytest.raises(ValueError):
        a.sum(2)
    c = u.MatD.zeros(3, 1)
    with pytest.raises(ValueError):  # in-place Mat form requires EQUAL shapes (Mat.scala:4565)
        a.iadd(c)


@pytest.mark.parametrize("op,sym", [("add", "iadd"), ("sub", "isub"), ("mul", "imul"), ("div", "idiv")])
def test_inplace_family_scalar_and_mat_forms(u, op, sym):
    """`:+= :-= :*= :/=` (Mat.scala:4513-4588), each with a scalar and with an equal-shape Mat, on a contiguous matrix,
    a transposed view and a row-slice view: bit-identical to the oracle, the receiver is returned, division by zero
    and by -0.0 give the IEEE results."""
    rng = np.random.default_rng(11)
    a = rng.standard_normal((33, 18)); b = rng.standard_normal((33, 18))
    b[4, 5] = 0.0; b[6, 7] = -0.0
    for make_view, make_ref in ((lambda m: m, lambda r: r), (lambda m: m.T, lambda r: r.T),
                                (lambda m: m.slice(3, 29, 0, 18), lambda r: mo.slice_(r, 3, 29, 0, 18))):
        m, ref_ = dev(u, a), mo.from2d(a)
        v, rv = make_view(m), make_ref(ref_)
        other = b if v.shape == b.shape else (b.T.copy() if v.shape == b.T.shape else b[3:29].copy())
        with np.errstate(all="ignore"):
            assert getattr(v, sym)(1.75) is v
            mo.inplace_scalar(op, rv, 1.75)
            assert getattr(v, sym)(dev(u, other)) is v
            mo.inplace_mat(op, rv, mo.from2d(other))
        assert np.array_equal(m.to_numpy(), ref_.to2d(), equal_nan=True)


def test_inplace_through_views(u):
    rng = np.random.default_rng(5)
    a = rng.standard_normal((40, 30))
    m = dev(u, a)
    ref_ = mo.from2d(a)
    v = m.slice(5, 25, 0, 30)            # row slice: stays a view (tall enough)
    assert v._buf is m._buf
    v += 2.0
    mo.inplace_scalar("add", mo.slice_(ref_, 5, 25, 0, 30), 2.0)
    t = m.T
    t *= 0.5
    mo.inplace_scalar("mul", ref_.T, 0.5)
    other = rng.standard_normal((30, 40))
    t -= dev(u, other)
    mo.inplace_mat("sub", ref_.T, mo.from2d(other))
    assert np.array_equal(m.to_numpy(), ref_.to2d())
    # a column slice of a WIDE matrix is materialised: writes do not reach the parent
    w = dev(u, a[:3, :])
    s = w.slice(0, 3, 1, 30)
    assert s._buf is not w._buf
    s += 1.0
    assert np.array_equal(w.to_numpy(), a[:3, :])
    # element get/set through a view
    t[2, 3] = 7.25
    assert m[3, 2] == 7.25 and t[2, 3] == 7.25


@pytest.mark.parametrize("shape", [(1, 1), (5, 3), (64, 64), (257, 256), (300, 400), (17, 5000)])
def test_unary_and_activations(u, shape):
    rng = np.random.default_rng(11)
    a = rng.standard_normal(shape) * 3
    a.flat[0] = 0.0
    if a.size > 3:
        a.flat[1] = -0.0; a.flat[2] = np.nan; a.flat[3] = np.inf
    m, o = dev(u, a), omat(a)
    assert np.array_equal((-m).to_numpy(), -a, equal_nan=True)
    assert np.array_equal(np.signbit(m.abs().to_numpy()), np.signbit(mo.abs_(o).to2d()))
    assert np.array_equal(m.abs().to_numpy(), mo.abs_(o).to2d(), equal_nan=True)
    got = m.relu().to_numpy(); want = mo.relu(o).to2d()
    assert np.array_equal(got, want, equal_nan=True) and np.array_equal(np.signbit(got), np.signbit(want))
    np.testing.assert_allclose(m.sigmoid().to_numpy(), mo.sigmoid(o).to2d(), rtol=1e-12, atol=0, equal_nan=True)
    np.testing.assert_allclose(m.exp().to_numpy(), np.exp(a), rtol=1e-12, equal_nan=True)
    # views
    np.testing.assert_allclose(m.T.sigmoid().to_numpy(), mo.sigmoid(o.T).to2d(), rtol=1e-12, atol=0, equal_nan=True)


@pytest.mark.parametrize("shape", [(1, 1), (3, 5), (5, 3), (64, 64), (65, 63), (300, 400), (7, 2049), (3, 9000), (2, 20000), (2, 70000), (4097, 3),
                                   (5, 16384), (3, 32768), (3, 32766), (2, 16385)])   # rows that exactly fill one / two CTAs of the on-chip kernel, and near misses
@pytest.mark.parametrize("axis", [0, 1])
def test_softmax(u, shape, axis):
    rng = np.random.default_rng(3)
    a = rng.uniform(-5.0, 0.5, size=shape)
    m = dev(u, a)
    got = m.softmax(axis).to_numpy()
    want = mo.softmax(omat(a), axis).to2d()
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=0)
    np.testing.assert_allclose(m.T.softmax(1 - axis).to_numpy(), want.T, rtol=1e-12, atol=0)


def test_softmax_special_lanes(u):
    a = np.array([[1.0, 2.0, 3.0], [np.nan, 1.0, 2.0], [1.0, np.nan, 2.0], [-np.inf, 0.0, 1.0],
                  [-np.inf, -np.inf, -np.inf], [np.inf, 1.0, 2.0], [0.0, 0.0, 0.0]])
    for axis, arr in ((1, a), (0, a.T.copy())):
        got = dev(u, arr).softmax(axis).to_numpy()
        want = mo.softmax(omat(arr), axis).to2d()
        np.testing.assert_allclose(got, want, rtol=1e-12, atol=0, equal_nan=True)
    mf = u.Mat.from_numpy(a[:1].astype(np.float32))
    assert mf.softmax().dtype == 0 and mf.sigmoid().dtype == 0   # Mat[Double] even for MatF


STRIP_SHAPES = [(1, 2), (5, 16), (127, 18), (128, 32), (129, 34), (300, 46), (2048, 64), (2049, 50), (4097, 130),
                (8192, 256), (16385, 34), (20000, 66), (32768, 48), (1000, 3000)]   # 1, 2, 4, 8, 16-CTA clusters; ragged rows and strips


@pytest.mark.parametrize("shape", STRIP_SHAPES)
def test_softmax_axis0_strip_kernel(u, shape):
    """softmax(0) through the single-read strip kernel (forced): against the oracle, against the two-read kernels,
    and on a view whose row pitch differs from its width; every cluster size, boxes hanging over the last row and
    strips hanging over the last column."""
    T, N = shape
    rng = np.random.default_rng(T * 31 + N)
    a = rng.standard_normal((T, N)) * 2.0 + 0.25
    want = mo.softmax(omat(a), 0).to2d()
    m = dev(u, a)
    try:
        u.set_softmax_axis0_path(2)
        got = m.softmax(0).to_numpy()
        again = m.softmax(0).to_numpy()
        big = rng.standard_normal((T + 3, N + 6))
        got_v = dev(u, big)._raw_slice(2, T + 2, 4, N + 4).softmax(0).to_numpy()
        u.set_softmax_axis0_path(1)
        two = m.softmax(0).to_numpy()
    finally:
        u.set_softmax_axis0_path(0)
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=0)
    np.testing.assert_allclose(got, two, rtol=1e-12, atol=0)
    assert np.array_equal(got, again), "run-to-run reproducible"
    np.testing.assert_allclose(got_v, mo.softmax(omat(big[2 : T + 2, 4 : N + 4].copy()), 0).to2d(), rtol=1e-12, atol=0)
    np.testing.assert_allclose(got.sum(axis=0), 1.0, rtol=0, atol=1e-12)


def test_softmax_axis0_strip_kernel_special_lanes(u):
    """Lanes the shifted form must hand to the reference's arithmetic (wide spread, +-Inf, a non-finite first row) next
    to ordinary ones in the same strips, NaN lanes, and a strip where only the LAST CTA of the cluster sees the value."""
    T, N = 5000, 40
    rng = np.random.default_rng(5)
    a = rng.standard_normal((T, N))
    a[17, 1] = np.nan
    a[0, 2] = np.nan
    a[4999, 3] = np.inf
    a[0, 4] = np.inf
    a[123, 5] = -np.inf
    a[0, 6] = -np.inf
    a[:, 7] = -np.inf
    a[2500, 8] = 900.0               # max - k > 300
    a[4000, 9] = -900.0              # k - min > 300
    a[:, 10] = np.linspace(-400.0, 400.0, T)
    a[0, 17] = 1e308; a[1, 17] = -1e308
    a[4999, 20] = 350.0
    a[:, 21] = 0.0
    a[3, 33] = -np.inf; a[4998, 33] = np.nan
    want = mo.softmax(omat(a), 0).to2d()
    try:
        u.set_softmax_axis0_path(2)
        got = dev(u, a).softmax(0).to_numpy()
        u.set_softmax_axis0_path(1)
        two = dev(u, a).softmax(0).to_numpy()
    finally:
        u.set_softmax_axis0_path(0)
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=0, equal_nan=True)
    np.testing.assert_allclose(two, want, rtol=1e-12, atol=0, equal_nan=True)
    assert np.array_equal(np.isnan(got), np.isnan(want))


def test_softmax_axis0_strip_kernel_refuses_what_it_cannot_take(u):
    a = np.random.default_rng(1).standard_normal((64, 33))
    try:
        u.set_softmax_axis0_path(2)
        with pytest.raises(ValueError):
            dev(u, a).softmax(0)                        # odd column count
        with pytest.raises(ValueError):
            dev(u, a.astype(np.float32)).softmax(0)     # FP32 input
    finally:
        u.set_softmax_axis0_path(0)
    with pytest.raises(ValueError):
        u.set_softmax_axis0_path(3)


# ---------------------------------------------------------------- reductions
RED_SHAPES = [(1, 1), (1, 9), (9, 1), (3, 5), (5, 3), (64, 64), (65, 63), (257, 256), (300, 400), (4097, 33), (33, 4097), (2, 70001), (70001, 2)]


@pytest.mark.parametrize("shape", RED_SHAPES)
def test_axis_and_whole_reductions(u, shape):
    rng = np.random.default_rng(17)
    a = rng.standard_normal(shape) * 10 + 3
    for M, O in ((dev(u, a), omat(a)), (dev(u, a).T, omat(a).T)):
        v = O.to2d()
        tol = sum_tol(v)
        assert abs(M.sum() - float(mo.mat_sum(O))) <= tol
        assert abs(M.mean() - float(mo.mat_mean(O))) <= tol / v.size
        np.testing.assert_allclose(M.variance(), float(mo.mat_variance(O)), rtol=1e-12)
        np.testing.assert_allclose(M.std(), float(mo.mat_std(O)), rtol=1e-12)
        for axis in (0, 1):
            lane_tol = 1e-12 * np.abs(v).sum(axis=axis, keepdims=True)
            assert np.all(np.abs(M.sum(axis).to_numpy() - mo.sum_axis(O, axis).to2d()) <= lane_tol)
            n = v.shape[axis]
            assert np.all(np.abs(M.mean(axis).to_numpy() - mo.mean_axis(O, axis).to2d()) <= lane_tol / n)
            np.testing.assert_allclose(M.var(axis).to_numpy(), mo.var_axis(O, axis).to2d(), rtol=1e-12, atol=1e-300)
            np.testing.assert_allclose(M.std(axis).to_numpy(), mo.std_axis(O, axis).to2d(), rtol=1e-12, atol=1e-300)
            assert np.array_equal(M.min(axis).to_numpy(), mo.min_axis(O, axis).to2d())
            assert np.array_equal(M.max(axis).to_numpy(), mo.max_axis(O, axis).to2d())
            assert np.array_equal(M.idxmin(axis).to_numpy(), mo.idxmin(O, axis).to2d())
            assert np.array_equal(M.idxmax(axis).to_numpy(), mo.idxmax(O, axis).to2d())
        assert M.argmin() == mo.argmin(O) and M.argmax() == mo.argmax(O)
        assert M.min() == mo.mat_min(O) and M.max() == mo.mat_max(O)


def test_reductions_with_ties_and_specials(u):
    a = np.array([[3.0, 1.0, 1.0, 3.0], [np.nan, 1.0, -0.0, 0.0], [0.0, -0.0, np.inf, np.nan], [-np.inf, 2.0, 2.0, -np.inf]])
    for M, O in ((dev(u, a), omat(a)), (dev(u, a).T, omat(a).T)):
        assert M.argmin() == mo.argmin(O) and M.argmax() == mo.argmax(O)
        for axis in (0, 1):
            assert np.array_equal(M.idxmin(axis).to_numpy(), mo.idxmin(O, axis).to2d())
            assert np.array_equal(M.idxmax(axis).to_numpy(), mo.idxmax(O, axis).to2d())
            g, w = M.min(axis).to_numpy(), mo.min_axis(O, axis).to2d()
            assert np.array_equal(g, w, equal_nan=True) and np.array_equal(np.signbit(g), np.signbit(w))
    big = np.zeros((3, 100000)); big[1, 70000] = -1.0; big[1, 99999] = -1.0; big[2, 5] = 5.0; big[0, 99998] = 5.0
    B = dev(u, big)
    assert B.argmin() == (1, 70000) and B.argmax() == (0, 99998)
    assert B.idxmax(1).to_numpy().ravel().tolist() == [99998, 0, 5]


def test_variance_is_shift_robust(u):
    rng = np.random.default_rng(1)
    a = rng.standard_normal((2000, 50)) + 1e6
    m = dev(u, a)
    np.testing.assert_allclose(m.var(0).to_numpy(), mo.var_axis(omat(a), 0).to2d(), rtol=1e-9)
    np.testing.assert_allclose(m.variance(), float(mo.mat_variance(omat(a))), rtol=1e-9)
    const = dev(u, np.full((300, 7), 4.25))
    assert np.array_equal(const.std(0).to_numpy(), np.zeros((1, 7)))


def test_reduction_reproducible(u):
    a = np.random.default_rng(2).standard_normal((3001, 517))
    m = dev(u, a)
    s0 = [m.sum(), m.sum(0).to_numpy().tobytes(), m.var(1).to_numpy().tobytes()]
    for _ in range(3):
        assert [m.sum(), m.sum(0).to_numpy().tobytes(), m.var(1).to_numpy().tobytes()] == s0


# ---------------------------------------------------------------- MatF
def test_matf_ops(u):
    rng = np.random.default_rng(9)
    a = rng.uniform(-5, 0.5, size=(65, 63)).astype(np.float32)
    m, o = dev(u, a), omat(a)
    assert np.array_equal((m * m).to_numpy(), a * a)
    assert np.array_equal((m / (m + 10.0)).to_numpy(), a / (a + np.float32(10.0)))
    assert np.array_equal((m - 0.5).to_numpy(), a - np.float32(0.5))
    assert np.array_equal(m.gt(-1.0).to_numpy(), a > -1.0)
    assert m.argmin() == mo.argmin(o) and m.argmax() == mo.argmax(o)
    np.testing.assert_allclose(m.sum(), float(mo.mat_sum(o)), rtol=1e-5)
    np.testing.assert_allclose(m.sum(0).to_numpy(), mo.sum_axis(o, 0).to2d(), rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(m.std(0).to_numpy(), mo.std_axis(o, 0).to2d(), rtol=1e-5)
    np.testing.assert_allclose((m @ m.T).to_numpy(), mo.matmul_pure(o, o.T).to2d(), rtol=1e-4, atol=1e-4)
    assert m.relu().dtype == 1 and np.array_equal(m.relu().to_numpy(), np.maximum(a, 0))


MATF_MM_SHAPES = [(128, 32, 128), (128, 160, 256), (256, 512, 384), (384, 4096, 128), (1024, 1056, 512),
                  (200, 300, 132), (129, 1000, 128), (1000, 36, 64), (130, 132, 260), (1, 4096, 512)]   # partial tiles in M, N and K


@pytest.mark.parametrize("mkn", MATF_MM_SHAPES)
def test_matf_matmul_tensor_core_path(u, mkn):
    """FP32 `*@` on the tcgen05 tensor cores (3xTF32, promoted to FP32 registers every 128 k) against float64 and
    against the FFMA kernel.  The reference pins FP32 matmul to tolerance only (MatTest.scala MatF rows; sgemm vs
    matmulPure, Mat.scala:3358-3377).  Bounds: element-wise the textbook sgemm bound k * 2^-24 * (|a| @ |b|), and for
    this N(0,1) data max |err| <= 6e-6 * sqrt(k) (OpenBLAS sgemm measures ~3e-6 * sqrt(k); an accumulator left in
    tensor memory over all of k measured 1.7e-4 * sqrt(k) at k = 4096, which this catches)."""
    M, K, N = mkn
    rng = np.random.default_rng(M + 3 * K + 7 * N)
    a = rng.standard_normal((M, K)).astype(np.float32); b = rng.standard_normal((K, N)).astype(np.float32)
    want = a.astype(np.float64) @ b.astype(np.float64)
    bound = K * 2.0 ** -24 * (np.abs(a).astype(np.float64) @ np.abs(b).astype(np.float64))
    da, db = dev(u, a), dev(u, b)
    try:
        u.set_matf_matmul_path(2)
        got_t = (da @ db).to_numpy()
        # a view with a different pitch and a 16-byte aligned offset stays on the tensor path
        big = rng.standard_normal((M + 8, K + 8)).astype(np.float32)
        got_v = (dev(u, big)._raw_slice(8, M + 8, 4, K + 4) @ db).to_numpy()
        want_v = big[8 : M + 8, 4 : K + 4].astype(np.float64) @ b.astype(np.float64)
        # operands it cannot take are refused, never silently rerouted, when the path is forced
        with pytest.raises(ValueError):
            dev(u, a.T.copy()).T @ db
        with pytest.raises(ValueError):
            dev(u, big)._raw_slice(8, M + 8, 3, K + 3) @ db
        # operands split once in global memory (what automatic takes from 2^30 multiply-adds up): the same bits
        u.set_matf_matmul_path(3)
        got_p = (da @ db).to_numpy()
        got_pv = (dev(u, big)._raw_slice(8, M + 8, 4, K + 4) @ db).to_numpy()
        ragged = (dev(u, big)._raw_slice(8, M + 8, 4, K + 1) @ dev(u, b[: K - 3])).to_numpy() if K > 3 else None   # K % 4 != 0
        with pytest.raises(ValueError):
            dev(u, a.T.copy()).T @ db
        u.set_matf_matmul_path(2)
        ragged_t = (dev(u, big)._raw_slice(8, M + 8, 4, K + 1) @ dev(u, b[: K - 3])).to_numpy() if K > 3 else None
        u.set_matf_matmul_path(1)
        got_c = (da @ db).to_numpy()
        u.set_matf_matmul_path(0)
        got_auto = (da @ db).to_numpy()
        got_auto_t = (dev(u, a.T.copy()).T @ db).to_numpy()    # automatic: falls back to the FFMA kernel
    finally:
        u.set_matf_matmul_path(0)
    assert got_t.dtype == np.float32
    assert np.all(np.abs(got_t - want) <= bound)
    assert np.max(np.abs(got_t - want)) <= 6e-6 * np.sqrt(K)
    assert np.max(np.abs(got_v - want_v)) <= 6e-6 * np.sqrt(K)
    assert np.array_equal(got_p, got_t) and np.array_equal(got_pv, got_v), "split in global memory = split in shared memory, bit for bit"
    if ragged is not None:
        assert np.array_equal(ragged, ragged_t)
        assert np.max(np.abs(ragged - big[8 : M + 8, 4 : K + 1].astype(np.float64) @ b[: K - 3].astype(np.float64))) <= 6e-6 * np.sqrt(K)
    assert np.array_equal(got_auto, got_t), "automatic choice = tensor path for tile-multiple row-major operands"
    assert np.array_equal(got_auto_t, got_c)
    np.testing.assert_allclose(got_t, got_c, rtol=1e-4, atol=1e-5 * np.sqrt(K))
    # bit-reproducible from run to run
    assert np.array_equal((da @ db).to_numpy(), got_t)


# ---------------------------------------------------------------- matmul / inverse / solve
MM_SHAPES = [(1, 1, 1), (3, 5, 3), (5, 3, 5), (8, 9, 8), (64, 64, 64), (65, 63, 65), (127, 129, 131), (256, 256, 256),
             (300, 400, 300), (1, 500, 1), (500, 1, 500), (40, 650, 40), (512, 512, 512), (1000, 17, 9)]


@pytest.mark.parametrize("mkn", MM_SHAPES)
def test_matmul_f64(u, mkn):
    M, K, N = mkn
    rng = np.random.default_rng(M * 7 + K * 3 + N)
    a = rng.standard_normal((M, K)); b = rng.standard_normal((K, N))
    want = a @ b
    tol = dict(rtol=1e-9, atol=1e-12 * K)   # MatmulModeSuite.scala:66-83 (BLAS vs pure contract)
    np.testing.assert_allclose((dev(u, a) @ dev(u, b)).to_numpy(), want, **tol)
    # transposed views on either side: no copies (Mat.scala:2815-2881; mmfnv / tmmfnv operands)
    np.testing.assert_allclose((dev(u, a.T.copy()).T @ dev(u, b)).to_numpy(), want, **tol)
    np.testing.assert_allclose((dev(u, a) @ dev(u, b.T.copy()).T).to_numpy(), want, **tol)
    np.testing.assert_allclose((dev(u, a.T.copy()).T @ dev(u, b.T.copy()).T).to_numpy(), want, **tol)
    if M > 4 and K > 4:
        big = rng.standard_normal((M + 3, K + 5))
        np.testing.assert_allclose((dev(u, big).slice(1, M + 1, 3, K + 3) @ dev(u, b)).to_numpy(), big[1 : M + 1, 3 : K + 3] @ b, **tol)


@pytest.mark.parametrize("mkn", [(33, 20000, 17), (64, 65536, 64), (1, 300000, 9), (9, 4096, 1), (200, 50000, 130), (8, 2047, 8)])
def test_matmul_f64_long_contractions_are_split_along_k(u, mkn):
    """A' A of a tall matrix, normal equations, `row *@ tall`: few output tiles and a long K.  The product is split along K
    over the grid and the partial C's are added in a fixed tree (matmul.cu launch_dgemm): element-wise the textbook bound
    K eps (|a| @ |b|) with room to spare, every operand orientation, run-to-run bit-reproducible."""
    M, K, N = mkn
    rng = np.random.default_rng(M + 13 * N + K)
    a = rng.standard_normal((M, K)) + 0.5; b = rng.standard_normal((K, N)) - 0.25
    want = a @ b
    bound = 1e-14 * np.sqrt(K) * (np.abs(a) @ np.abs(b)) + 1e-300
    da, db, dat, dbt = dev(u, a), dev(u, b), dev(u, a.T.copy()).T, dev(u, b.T.copy()).T
    got = (da @ db).to_numpy()
    assert np.all(np.abs(got - want) <= bound)
    for x, yv in ((dat, db), (da, dbt), (dat, dbt)):
        assert np.all(np.abs((x @ yv).to_numpy() - want) <= bound)
    assert np.array_equal((da @ db).to_numpy(), got)
    # the Gram of a tall matrix through a transposed view, as the reference's OLS spells it (`X.T *@ X`)
    g = (db.T @ db).to_numpy()
    assert np.all(np.abs(g - b.T @ b) <= 1e-14 * np.sqrt(K) * (np.abs(b).T @ np.abs(b)))


@pytest.mark.parametrize("mkn", [(33, 20000, 17), (1, 300000, 9), (64, 65536, 64), (130, 8192, 200)])
def test_matf_long_contractions(u, mkn):
    """MatF `*@` with few output tiles and a long K (A' A of a tall matrix) is promoted to the split-K FP64 path and
    rounded to FP32 once: within half an ulp of the float64 product, every orientation, bit-reproducible."""
    M, K, N = mkn
    rng = np.random.default_rng(M + 17 * N + K)
    a = (rng.standard_normal((M, K)) + 0.5).astype(np.float32); b = (rng.standard_normal((K, N)) - 0.25).astype(np.float32)
    want = a.astype(np.float64) @ b.astype(np.float64)
    da, db, dat, dbt = dev(u, a), dev(u, b), dev(u, a.T.copy()).T, dev(u, b.T.copy()).T
    got = (da @ db).to_numpy()
    assert got.dtype == np.float32
    tol = 2.0 ** -23 * np.abs(want) + 1e-13 * np.sqrt(K) * (np.abs(a).astype(np.float64) @ np.abs(b).astype(np.float64))
    assert np.all(np.abs(got - want) <= tol)
    for x, yv in ((dat, db), (da, dbt), (dat, dbt)):
        assert np.all(np.abs((x @ yv).to_numpy() - want) <= tol)
    assert np.array_equal((da @ db).to_numpy(), got)


def test_matmul_matches_pinned_oracle(u):
    a = pc.corpus()[: 65 * 63].reshape(65, 63)
    m = dev(u, a)
    want = mo.matmul_pure(omat(a), omat(a).T).to2d()
    got = (m @ m.T).to_numpy()
    scale = np.abs(a) @ np.abs(a).T
    assert np.all(np.abs(got - want) <= 1e-12 * scale)


@pytest.mark.parametrize("n", [1, 2, 3, 5, 9, 17, 64, 100])
def test_inverse_and_solve(u, n):
    rng = np.random.default_rng(n)
    a = rng.standard_normal((n, n)) + np.eye(n) * 0.1
    b = rng.standard_normal((n, 3))
    inv = dev(u, a).inverse().to_numpy()
    want = mo.inverse(omat(a)).to2d()
    np.testing.assert_allclose(inv, want, rtol=1e-10, atol=1e-13)
    assert np.array_equal(inv, want), "same pivoting and operation order as luDecomposeD -> bit-identical"
    x = dev(u, a).solve(dev(u, b)).to_numpy()
    assert np.array_equal(x, mo.solve(omat(a), omat(b)).to2d())
    np.testing.assert_allclose(dev(u, a.T.copy()).T.inverse().to_numpy(), want, rtol=1e-10, atol=1e-13)


def test_singular_raises(u):
    a = np.array([[1.0, 2.0], [2.0, 4.0]])
    with pytest.raises(u.SingularMatrixError):
        dev(u, a).inverse()
    with pytest.raises(ValueError):
        u.MatD.zeros(2, 3).inverse()


def test_launch_counter_moves(u):
    n0 = u.launch_count()
    m = u.MatD.ones(8, 8)
    (m + m).sum()
    assert u.launch_count() > n0


def test_mask_select_staged_kernel_and_offset_reuse(u):
    """m(mask) through the shared-memory staged scatter (contiguous f64 / f32 operands) and the reuse of the count's
    block offsets by the fill that follows it: same result as NumPy for ragged sizes, and a mask that is CHANGED
    between um_mask_count and um_mask_select (another library call in between) must not see stale offsets."""
    import ctypes as C
    rng = np.random.default_rng(77)
    for shape in [(1, 1), (3, 5), (1, 4095), (1, 4096), (1, 4097), (257, 129), (1000, 1037), (2048, 2048)]:
        a = rng.standard_normal(shape)
        for frac in (0.0, 0.03, 0.5, 1.0):
            mk = rng.random(shape) < frac
            for arr in (a, a.astype(np.float32)):
                got = dev(u, arr)[dev(u, mk)].to_numpy()
                assert np.array_equal(got.ravel(), arr[mk]), (shape, frac, arr.dtype)
    # stale offsets: count on one mask, flip the mask in place, then fill
    a = rng.standard_normal((300, 70)); mk = rng.random((300, 70)) < 0.5
    da, dm = dev(u, a), dev(u, mk)
    k = C.c_int64(0)
    u._lib.check(u._lib.lib.um_mask_count(dm._ref(), C.byref(k)))
    assert k.value == int(mk.sum())
    flipped = ~dm                                    # a library call in between ...
    u._lib.check(u._lib.lib.um_copy(flipped._ref(), dm._ref()))   # ... that rewrites the mask buffer in place
    k2 = int((~mk).sum())
    res = u.Mat._alloc(1, k2, u._lib.F64)
    u._lib.check(u._lib.lib.um_mask_select(da._ref(), dm._ref(), res._ref()))
    assert np.array_equal(res.to_numpy().ravel(), a[~mk])


@pytest.mark.parametrize("tile", [64, 128])
@pytest.mark.parametrize("mkn", [(1, 1, 1), (65, 63, 65), (127, 129, 131), (300, 400, 300), (1000, 17, 9), (1024, 512, 1024)])
def test_matmul_f64_both_cta_tiles(u, mkn, tile):
    """The 64 x 64 and the 128 x 128 CTA tile of the FP64 `*@` kernel, each forced: same contract, all four operand
    orientations, partial tiles in every dimension; the automatic choice equals one of them bit for bit."""
    M, K, N = mkn
    rng = np.random.default_rng(M + 5 * K + 11 * N)
    a = rng.standard_normal((M, K)); b = rng.standard_normal((K, N))
    want = a @ b
    tol = dict(rtol=1e-9, atol=1e-12 * K)
    da, db, dat, dbt = dev(u, a), dev(u, b), dev(u, a.T.copy()).T, dev(u, b.T.copy()).T
    try:
        u.set_matd_matmul_tile(tile)
        got = (da @ db).to_numpy()
        np.testing.assert_allclose(got, want, **tol)
        for x, yv in ((dat, db), (da, dbt), (dat, dbt)):
            np.testing.assert_allclose((x @ yv).to_numpy(), want, **tol)
        assert np.array_equal((da @ db).to_numpy(), got), "run-to-run reproducible"
        u.set_matd_matmul_tile(0)
        auto = (da @ db).to_numpy()
        u.set_matd_matmul_tile(192 - tile)
        other = (da @ db).to_numpy()
    finally:
        u.set_matd_matmul_tile(0)
    assert np.array_equal(auto, got) or np.array_equal(auto, other)
    with pytest.raises(ValueError):
        u.set_matd_matmul_tile(96)


# ---------------------------------------------------------------- BASELINE config sizes (configs 2 and 3)
@pytest.mark.parametrize("n,dtype", [(4096, np.float64), (8192, np.float64), (4096, np.float32), (8192, np.float32)])
def test_matmul_at_config3_sizes(u, n, dtype):
    """`a *@ b` at the sizes of BASELINE config 3 (the sweep 1024^2 .. 16384^2): sampled rows and columns of the
    n x n product against NumPy in float64.  FP64: the BLAS-vs-pure contract of MatmulModeSuite.scala:66-83;
    FP32: the sgemm bound k * 2^-24 * (|a| @ |b|) and 6e-6 * sqrt(k) on N(0,1) data (as the small-shape tests)."""
    f64 = dtype is np.float64
    a = u.MatD.randn(n, n, seed=31 + n) if f64 else u.MatF.randn(n, n, seed=31 + n)
    b = u.MatD.randn(n, n, seed=77 + n) if f64 else u.MatF.randn(n, n, seed=77 + n)
    c = a @ b
    ha, hb, hc = a.to_numpy(), b.to_numpy(), c.to_numpy()
    assert hc.dtype == dtype and hc.shape == (n, n)
    rows = np.array([0, 1, 127, 128, 129, n // 2 - 1, n // 2, n - 129, n - 2, n - 1])
    cols = np.array([0, 5, 127, 128, n // 3, n - 128, n - 1])
    want_r = ha[rows].astype(np.float64) @ hb.astype(np.float64)
    want_c = ha.astype(np.float64) @ hb[:, cols].astype(np.float64)
    if f64:
        tol = dict(rtol=1e-9, atol=1e-12 * n)
        np.testing.assert_allclose(hc[rows], want_r, **tol)
        np.testing.assert_allclose(hc[:, cols], want_c, **tol)
        # a transposed left operand is read in place (mmfnv / tmmfnv operands): (a.T).T-free check against the same rows
        ct = (a.T @ b).to_numpy()
        np.testing.assert_allclose(ct[rows], ha[:, rows].T @ hb, **tol)
    else:
        bound_r = n * 2.0 ** -24 * (np.abs(ha[rows]).astype(np.float64) @ np.abs(hb).astype(np.float64))
        assert np.all(np.abs(hc[rows] - want_r) <= bound_r)
        assert np.max(np.abs(hc[rows] - want_r)) <= 6e-6 * np.sqrt(n)
        assert np.max(np.abs(hc[:, cols] - want_c)) <= 6e-6 * np.sqrt(n)


def test_config2_ops_at_32768_squared(u):
    """BASELINE config 2 at full size (32768 x 32768 FP64, 8 GiB): every op of the config is checked against a second
    method on sampled lanes -- NumPy on the lanes copied back -- and through size-independent identities."""
    n = 32768
    m = u.MatD.randn(n, n, seed=2024)
    lanes = [0, 1, 63, 64, 4095, n // 2 + 3, n - 2, n - 1]
    cols = {j: m.slice(0, n, j, j + 1).to_numpy().ravel() for j in lanes}
    rows = {i: m.slice(i, i + 1, 0, n).to_numpy().ravel() for i in lanes}
    s0, s1 = m.sum(0).to_numpy().ravel(), m.sum(1).to_numpy().ravel()
    mean0, mean1 = m.mean(0), m.mean(1)
    v0, v1 = m.var(0).to_numpy().ravel(), m.var(1).to_numpy().ravel()
    sd0 = m.std(0).to_numpy().ravel()
    for j in lanes:
        c, r = cols[j], rows[j]
        assert abs(s0[j] - math.fsum(c)) <= 1e-12 * np.abs(c).sum() and abs(s1[j] - math.fsum(r)) <= 1e-12 * np.abs(r).sum()
        np.testing.assert_allclose(v0[j], c.var(), rtol=1e-12)
        np.testing.assert_allclose(v1[j], r.var(), rtol=1e-12)
        np.testing.assert_allclose(sd0[j], c.std(), rtol=1e-12)
    # a sum of the axis sums is the whole-matrix sum either way (a checksum of checksums)
    tot = m.sum()
    scale = 1e-12 * n * n
    assert abs(math.fsum(s0) - tot) <= scale and abs(math.fsum(s1) - tot) <= scale
    np.testing.assert_allclose(mean0.to_numpy().ravel(), s0 / n, rtol=1e-15)
    # m - m.mean(0): the centred matrix has column sums ~ 0 and the sampled lanes equal NumPy's bit for bit
    cen = m - mean0
    mu0 = mean0.to_numpy().ravel()
    for j in lanes:
        assert np.array_equal(cen.slice(0, n, j, j + 1).to_numpy().ravel(), cols[j] - mu0[j])
    assert np.max(np.abs(cen.sum(0).to_numpy())) <= 1e-12 * n * 8
    del cen
    cen1 = m - mean1
    mu1 = mean1.to_numpy().ravel()
    for i in lanes:
        assert np.array_equal(cen1.slice(i, i + 1, 0, n).to_numpy().ravel(), rows[i] - mu1[i])
    del cen1
    sg = m.sigmoid()
    for i in lanes:
        np.testing.assert_allclose(sg.slice(i, i + 1, 0, n).to_numpy().ravel(), 1.0 / (1.0 + np.exp(-rows[i])), rtol=1e-12)
    del sg
    for ax, lanes_of in ((0, cols), (1, rows)):
        sm = m.softmax(ax)
        for j in lanes:
            x = lanes_of[j]
            got = (sm.slice(0, n, j, j + 1) if ax == 0 else sm.slice(j, j + 1, 0, n)).to_numpy().ravel()
            e = np.exp(x - x.max())
            np.testing.assert_allclose(got, e / e.sum(), rtol=1e-12, atol=0)
        # every lane of a softmax sums to one
        lane_sums = sm.sum(ax).to_numpy().ravel()
        assert np.max(np.abs(lane_sums - 1.0)) <= 1e-12
        del sm


# ---------------------------------------------------------------- allocator / device rules (include/unimat.h)
def test_free_from_another_host_thread_and_one_device_per_process(u):
    """um_free may run on a thread that never touched the library (a JVM Cleaner, Python's GC): the block is released
    on the device and stream it was allocated on and no context is opened; a pointer that is not a live block is
    refused; a worker thread without um_init inherits the process's device; a second device is refused."""
    import ctypes as C
    import threading
    from uni_b200._lib import lib
    a = np.random.default_rng(0).standard_normal((257, 129))
    m = dev(u, a)
    s = (m + m).to_numpy()
    ptrs = []
    for _ in range(4):
        p = C.c_void_p()
        assert lib.um_malloc(1 << 20, C.byref(p)) == 0
        ptrs.append(p.value)
    out = {}

    def worker():
        out["free"] = [lib.um_free(C.c_void_p(p)) for p in ptrs]
        out["twice"] = lib.um_free(C.c_void_p(ptrs[0]))           # no longer a live block
        out["foreign"] = lib.um_free(C.c_void_p(0x1000))
        w = dev(u, a)                                              # first library call of this thread: inherits the device
        out["sum"] = (w + w).to_numpy()
        del w

    t = threading.Thread(target=worker)
    t.start(); t.join()
    assert out["free"] == [0, 0, 0, 0] and out["twice"] != 0 and out["foreign"] != 0
    assert np.array_equal(out["sum"], s)
    # the freed blocks are back in the pool and the owning thread goes on working
    assert np.array_equal((m + m).to_numpy(), s)
    other = 1 if u.device_count() > 1 else 7
    assert lib.um_init(other) != 0          # another device (or one that does not exist): refused
    assert lib.um_init(0) == 0
