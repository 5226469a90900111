"""CPU oracle for the uni `Mat[T]` hot path -- TEST INFRASTRUCTURE ONLY.

A NumPy restatement of the reference's Scala semantics (association orders, ordering
rules, layout guard) for the operators SURVEY.md section 8a lists.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import this module; the product path (`uni_b200`, `libunimat.so`) never does.

Parity is PINNED: `tests/test_oracle_mat_parity.py` checks every in-scope row of the
reference's own bit-pattern fixture `test-data/mat-parity/scala-reference.txt` (copied to
`tests/golden/mat-parity/`) against this file, bit for bit.

Every function cites the reference file:line it follows (paths relative to
/root/reference/src/main/scala/uni/data/ unless stated).
"""
from __future__ import annotations

import math
import numpy as np

# ---------------------------------------------------------------------------------------
# constants (Mat.scala:356,366,374,383,391,400)
# ---------------------------------------------------------------------------------------
PARALLEL_THRESHOLD = 4096
MAX_SUM_CHUNKS = 16
SCAN_PARALLEL_THRESHOLD = 1 << 16
STRIDED_SUM_PARALLEL_THRESHOLD = 1 << 16

FNV_OFFSET = 0xCBF29CE484222325
FNV_PRIME = 0x100000001B3
MASK64 = (1 << 64) - 1


# ---------------------------------------------------------------------------------------
# storage + views  (Mat.scala:125-133 MatData, :170-189 create, :198-199 transposeView)
# ---------------------------------------------------------------------------------------
class OMat:
    """array + (rows, cols, transposed, offset, rs, cs): Mat.scala:125-133."""

    __slots__ = ("data", "rows", "cols", "transposed", "offset", "rs", "cs")

    def __init__(self, data, rows, cols, transposed=False, offset=0, rs=-1, cs=-1):
        self.data = data
        self.rows = int(rows)
        self.cols = int(cols)
        self.transposed = bool(transposed)
        self.offset = int(offset)
        self.rs = int(rs)
        self.cs = int(cs)

    # -- layout predicates --------------------------------------------------------------
    def is_standard_contiguous(self):  # Mat.scala:1878-1880
        return (self.rs == self.cols and self.cs == 1 and not self.transposed) or (
            self.rs == 1 and self.cs == self.rows and self.transposed
        )

    def is_contiguous(self):  # Mat.scala:49-50 (row-major standard strides)
        return self.rs == self.cols and self.cs == 1

    def is_weird_layout(self):  # Mat.scala:1865-1876
        if self.is_standard_contiguous():
            return False
        if self.rs == 0 or self.cs == 0:
            return False
        leading = self.cols if self.transposed else self.rows
        major = self.cs if self.transposed else self.rs
        return major > max(leading, 1)

    def fast_d(self):  # Mat.scala:403-405
        return (
            self.data.dtype == np.float64
            and self.is_contiguous()
            and self.offset == 0
            and self.data.size == self.rows * self.cols
        )

    # -- element access -----------------------------------------------------------------
    def index_grid(self):
        r = np.arange(self.rows, dtype=np.int64)[:, None]
        c = np.arange(self.cols, dtype=np.int64)[None, :]
        return self.offset + r * self.rs + c * self.cs

    def to2d(self):
        """Logical rows x cols values (always a fresh array)."""
        if self.rows == 0 or self.cols == 0:
            return np.zeros((self.rows, self.cols), dtype=self.data.dtype)
        return self.data[self.index_grid()]

    def to_array(self):
        return self.to2d().reshape(-1)

    @property
    def T(self):  # transposeView, Mat.scala:198-199
        return OMat(self.data, self.cols, self.rows, not self.transposed, self.offset, self.cs, self.rs)

    @property
    def shape(self):
        return (self.rows, self.cols)


def create(data, rows, cols, transposed=False, offset=0, rs=-1, cs=-1) -> OMat:
    """Internal.create: derive default strides, materialise fragmented views (Mat.scala:170-189)."""
    a_rs = rs if rs >= 0 else (1 if transposed else cols)
    a_cs = cs if cs >= 0 else (rows if transposed else 1)
    m = OMat(data, rows, cols, transposed, offset, a_rs, a_cs)
    if m.is_weird_layout():
        return mat_copy(m)
    return m


def mat_copy(m: OMat) -> OMat:  # Mat.scala:1886-1899
    return create(m.to_array().copy(), m.rows, m.cols)


def from2d(a) -> OMat:
    a = np.ascontiguousarray(a)
    return create(a.reshape(-1).copy(), a.shape[0], a.shape[1])


def slice_(m: OMat, r0, r1, c0, c1) -> OMat:
    """m.slice(r0 until r1, c0 until c1): Mat.scala:1905-1927."""
    if not (0 <= r0 <= r1 <= m.rows and 0 <= c0 <= c1 <= m.cols):
        raise ValueError("slice out of bounds")
    off = m.offset + r0 * m.rs + c0 * m.cs
    return create(m.data, r1 - r0, c1 - c0, m.transposed, off, m.rs, m.cs)


def broadcast_to(m: OMat, rows, cols) -> OMat:  # Mat.scala:1933-1953
    if m.rows == rows and m.cols == cols:
        return m
    if not ((m.rows == rows or m.rows == 1) and (m.cols == cols or m.cols == 1)):
        raise ValueError(f"Cannot broadcast shape {m.shape} to ({rows}, {cols})")
    nrs = 0 if (m.rows == 1 and rows > 1) else m.rs
    ncs = 0 if (m.cols == 1 and cols > 1) else m.cs
    return create(m.data, rows, cols, m.transposed, m.offset, nrs, ncs)


# ---------------------------------------------------------------------------------------
# FNV digests / quantisers  (apps/MatParityGen.scala:102-111,150,158,165)
# ---------------------------------------------------------------------------------------
def fnv_words(words) -> int:
    d = FNV_OFFSET
    for w in words:
        d = ((d ^ (int(w) & MASK64)) * FNV_PRIME) & MASK64
    return d


def fnv(xs) -> int:
    return fnv_words(np.ascontiguousarray(xs, dtype=np.float64).reshape(-1).view(np.uint64).tolist())


def fnv_f32(xs) -> int:
    """fnvF: one word per float, the raw 32 bits zero-extended (MatParityGen fnvF)."""
    return fnv_words(np.ascontiguousarray(xs, dtype=np.float32).reshape(-1).view(np.uint32).astype(np.uint64).tolist())


def fnv_mask(mask) -> int:
    return fnv_words(np.asarray(mask).astype(np.uint8).reshape(-1).tolist())


def bits(x: float) -> int:
    return int(np.float64(x).view(np.uint64))


def canon(x: float) -> int:
    return 0x7FF8000000000000 if math.isnan(x) else bits(x)


def fnv_canon(xs) -> int:
    return fnv_words([canon(float(v)) for v in np.asarray(xs, dtype=np.float64).reshape(-1)])


def quantize(xs, scale) -> list:
    """floor(x*scale + 0.5).toLong as two's-complement words."""
    q = np.floor(np.asarray(xs, dtype=np.float64) * scale + 0.5)
    return [int(v) & MASK64 for v in q.astype(np.int64).reshape(-1)]


# ---------------------------------------------------------------------------------------
# whole-matrix sum: association order (Mat.scala:1042-1150)
# ---------------------------------------------------------------------------------------
def _sum_range(a: np.ndarray) -> float:
    """sumRange: 8 strided accumulators + sequential tail (Mat.scala:1042-1055)."""
    n = a.size
    n8 = (n // 8) * 8
    if n8:
        # accumulator k folds a[k], a[k+8], ... sequentially from 0.0
        blk = a[:n8].reshape(-1, 8)
        acc = np.zeros(8, dtype=a.dtype)
        for row in blk:
            acc = acc + row
        s = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]))
    else:
        s = a.dtype.type(0)
    for v in a[n8:]:
        s = s + v
    return s


def _sum_range_fast(a: np.ndarray) -> float:
    """Same order as _sum_range, vectorised: cumulative fold down the 8 accumulator lanes."""
    n = a.size
    n8 = (n // 8) * 8
    if n8:
        blk = a[:n8].reshape(-1, 8)
        # np.add.accumulate along axis 0 is a strict sequential fold per column
        acc = np.add.accumulate(blk, axis=0)[-1]
        # first add is 0.0 + a[k] == a[k] exactly (and -0.0 + 0.0 -> 0.0 matches 0.0 + -0.0 = 0.0)
        acc = acc + a.dtype.type(0)
        s = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]))
    else:
        s = a.dtype.type(0)
    for v in a[n8:]:
        s = s + v
    return s


def sum_d(a: np.ndarray) -> float:
    """sumD: chunked (<=16 chunks) 8-accumulator sum (Mat.scala:1064-1081)."""
    n = a.size
    if n < PARALLEL_THRESHOLD:
        return _sum_range_fast(a)
    chunks = min(MAX_SUM_CHUNKS, max(n // PARALLEL_THRESHOLD, 1))
    step = (n + chunks - 1) // chunks
    s = a.dtype.type(0)
    for c in range(chunks):
        lo = c * step
        hi = min(lo + step, n)
        p = _sum_range_fast(a[lo:hi]) if lo < hi else a.dtype.type(0)
        s = s + p
    return s


def _sum_row_block(m2d: np.ndarray) -> float:
    """sumRowBlock: 8 accumulators carried across rows + a carried tail (Mat.scala:1121-1150)."""
    rows, cols = m2d.shape
    c8 = (cols // 8) * 8
    z = m2d.dtype.type(0)
    if c8 and rows:
        lanes = m2d[:, :c8].reshape(rows * (c8 // 8), 8)
        acc = np.add.accumulate(lanes, axis=0)[-1] + z
        s8 = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]))
    else:
        s8 = ((z + z) + (z + z)) + ((z + z) + (z + z))
    tail = z
    if cols > c8 and rows:
        t = m2d[:, c8:].reshape(-1)
        tail = np.add.accumulate(np.concatenate([[z], t]))[-1]
    return s8 + tail


def sum_strided(m: OMat) -> float:
    """sumStrided over the logical row-major sequence (Mat.scala:1100-1118)."""
    v = m.to2d()
    rows, cols = v.shape
    if rows * cols < STRIDED_SUM_PARALLEL_THRESHOLD or rows < 8:
        return _sum_row_block(v)
    chunks = min(MAX_SUM_CHUNKS, max(rows // 8, 1))
    step = (rows + chunks - 1) // chunks
    total = v.dtype.type(0)
    for c in range(chunks):
        lo = c * step
        hi = min(lo + step, rows)
        p = _sum_row_block(v[lo:hi]) if lo < hi else v.dtype.type(0)
        total = total + p
    return total


def _generic_fold(vals: np.ndarray):
    """Boxed general path: one accumulator, row-major, from zero (Mat.scala:2270-2279)."""
    z = vals.dtype.type(0)
    if vals.size == 0:
        return z
    return np.add.accumulate(np.concatenate([[z], vals.reshape(-1)]))[-1]


def mat_sum(m: OMat):
    """m.sum (Mat.scala:2251-2279)."""
    if m.data.dtype == np.float64:
        if m.is_contiguous() and m.offset == 0 and m.data.size == m.rows * m.cols:
            return sum_d(m.data)
        return sum_strided(m)
    return _generic_fold(m.to_array())


def mat_mean(m: OMat):
    """m.mean (Mat.scala:3469-3485): 0 when empty, else sum / n."""
    n = m.rows * m.cols
    if n == 0:
        return m.data.dtype.type(0)
    return mat_sum(m) / m.data.dtype.type(n)


def mat_variance(m: OMat):
    """varianceD: mean from `sum`, squared deviations a plain row-major fold (Mat.scala:904-919,4784-4805)."""
    n = m.rows * m.cols
    t = m.data.dtype.type
    mu = mat_sum(m) / t(n)
    d = m.to_array() - mu
    return _generic_fold(d * d) / t(n)


def mat_std(m: OMat):
    """m.std = sqrt(variance) (Mat.scala:4699-4717)."""
    return np.sqrt(mat_variance(m))


# ---------------------------------------------------------------------------------------
# axis reductions (Mat.scala:586-739, 3414-3492, 4720-4739)
# ---------------------------------------------------------------------------------------
def _fold_axis(v: np.ndarray, axis: int) -> np.ndarray:
    """Sequential fold in increasing index order per lane, from 0 (colSumsD/rowSumsD)."""
    z = np.zeros_like(v[0:1, :] if axis == 0 else v[:, 0:1])
    if v.shape[axis] == 0:
        return z.reshape(-1)
    acc = np.add.accumulate(np.concatenate([z, v], axis=axis), axis=axis)
    return (acc[-1, :] if axis == 0 else acc[:, -1]).reshape(-1)


def sum_axis(m: OMat, axis: int) -> OMat:
    """m.sum(axis) -> 1 x cols (axis 0) or rows x 1 (axis 1) (Mat.scala:3414-3467)."""
    if axis not in (0, 1):
        raise ValueError("axis must be 0 or 1")
    r = _fold_axis(m.to2d(), axis)
    return create(r, 1, m.cols) if axis == 0 else create(r, m.rows, 1)


def mean_axis(m: OMat, axis: int) -> OMat:
    """m.mean(axis) = sum(axis) / n (Mat.scala:3487-3492)."""
    s = sum_axis(m, axis)
    n = m.rows if axis == 0 else m.cols
    return create(s.data / s.data.dtype.type(n), s.rows, s.cols)


def var_axis(m: OMat, axis: int) -> OMat:
    """Pre-sqrt quantity of stdAxisD: sum((x-mu)^2)/n, mu = mean(axis) (Mat.scala:719-739)."""
    v = m.to2d()
    n = m.rows if axis == 0 else m.cols
    t = v.dtype.type
    mu = _fold_axis(v, axis) / t(n)
    d = v - (mu[None, :] if axis == 0 else mu[:, None])
    ss = _fold_axis(d * d, axis)
    r = ss / t(n)
    return create(r, 1, m.cols) if axis == 0 else create(r, m.rows, 1)


def std_axis(m: OMat, axis: int) -> OMat:
    """m.std(axis): population std per lane (Mat.scala:4720-4739 -> stdAxisD :719-739)."""
    r = var_axis(m, axis)
    return create(np.sqrt(r.data), r.rows, r.cols)


# ---------------------------------------------------------------------------------------
# ordering: java.lang.Double.compare (Mat.scala:418-440,892-895)
# ---------------------------------------------------------------------------------------
def total_key(v: np.ndarray) -> np.ndarray:
    """Map doubles/floats to int64 keys ordered like java.lang.Double.compare:
    -Inf < ... < -0.0 < +0.0 < ... < +Inf < NaN, every NaN equal."""
    v = np.asarray(v)
    if v.dtype == np.float32:
        b = v.view(np.int32).astype(np.int64)
        nan_key = np.int64(0x7FC00000)
        flip = np.int64(0x7FFFFFFF)
    else:
        b = v.astype(np.float64).view(np.int64).copy()
        nan_key = np.int64(0x7FF8000000000000)
        flip = np.int64(0x7FFFFFFFFFFFFFFF)
    b = np.where(np.isnan(v), nan_key, b)
    return np.where(b < 0, b ^ flip, b)


def _arg_first(keys: np.ndarray, want_max: bool) -> int:
    """First occurrence of the extremum in scan order ('replace only on strictly better')."""
    return int(np.argmax(keys) if want_max else np.argmin(keys))


def mat_min(m: OMat):
    if m.rows * m.cols == 0:
        raise ArithmeticError("empty")  # UnsupportedOperationException, Mat.scala:2216
    a = m.to_array()
    return a[_arg_first(total_key(a), False)]


def mat_max(m: OMat):
    if m.rows * m.cols == 0:
        raise ArithmeticError("empty")
    a = m.to_array()
    return a[_arg_first(total_key(a), True)]


def argmin(m: OMat):
    """(row, col) of the first minimum in row-major scan order (Mat.scala:2281-2334,470-548)."""
    if m.rows * m.cols == 0:
        raise ArithmeticError("empty")
    i = _arg_first(total_key(m.to_array()), False)
    return (i // m.cols, i % m.cols)


def argmax(m: OMat):
    if m.rows * m.cols == 0:
        raise ArithmeticError("empty")
    i = _arg_first(total_key(m.to_array()), True)
    return (i // m.cols, i % m.cols)


def _lane_arg(m: OMat, axis: int, want_max: bool) -> np.ndarray:
    k = total_key(m.to2d())
    return (np.argmax(k, axis=axis) if want_max else np.argmin(k, axis=axis)).astype(np.int32)


def idxmin(m: OMat, axis: int) -> OMat:
    """MatPandasOps.scala:20-41: per-lane first-occurrence argmin -> Mat[Int]."""
    r = _lane_arg(m, axis, False)
    return create(r, 1, m.cols) if axis == 0 else create(r, m.rows, 1)


def idxmax(m: OMat, axis: int) -> OMat:
    """MatPandasOps.scala:43-64."""
    r = _lane_arg(m, axis, True)
    return create(r, 1, m.cols) if axis == 0 else create(r, m.rows, 1)


def min_axis(m: OMat, axis: int) -> OMat:
    """m.min(axis) under the total order (Mat.scala:3503-3556,748-831)."""
    v = m.to2d()
    idx = _lane_arg(m, axis, False).astype(np.int64)
    r = v[idx, np.arange(m.cols)] if axis == 0 else v[np.arange(m.rows), idx]
    return create(r, 1, m.cols) if axis == 0 else create(r, m.rows, 1)


def max_axis(m: OMat, axis: int) -> OMat:
    v = m.to2d()
    idx = _lane_arg(m, axis, True).astype(np.int64)
    r = v[idx, np.arange(m.cols)] if axis == 0 else v[np.arange(m.rows), idx]
    return create(r, 1, m.cols) if axis == 0 else create(r, m.rows, 1)


# ---------------------------------------------------------------------------------------
# running folds (Mat.scala:3635-3745; cumExtremumD :929-964)
# ---------------------------------------------------------------------------------------
def cumsum(m: OMat, axis=None) -> OMat:
    """Sequential running total per lane, seeded with 0 (Mat.scala:3635-3690).  np.cumsum on a strided
    1-D lane is the same left fold; float32 stays float32 (Numeric[Float].plus)."""
    v = m.to2d()
    if axis is None:  # flatten (row-major), 1 x n
        return create(np.cumsum(v.reshape(-1), dtype=v.dtype), 1, m.rows * m.cols)
    if axis not in (0, 1):
        raise ValueError(f"axis must be 0 or 1, got {axis}")
    return create(np.cumsum(v, axis=axis, dtype=v.dtype).reshape(-1), m.rows, m.cols)


def _cum_extremum(m: OMat, axis: int, want_max: bool) -> OMat:
    """acc = lane[0]; every cell (the first included) written after `if better(v, acc) acc = v`
    under java.lang.Double.compare (Mat.scala:3692-3744, 929-964).  The running extremum of the
    order keys finds where the accumulator is replaced (strictly better than everything before it); each
    cell then copies the element of the most recent replacement, so NaN payloads and signed zeros are the
    reference's."""
    if axis not in (0, 1):
        raise ValueError(f"axis must be 0 or 1, got {axis}")
    v = m.to2d()
    if m.rows * m.cols == 0:
        if (axis == 0 and m.rows == 0 and m.cols > 0) or (axis == 1 and m.cols == 0 and m.rows > 0):
            raise ValueError("cummax/cummin of an empty lane")  # apply's require, Mat.scala:3694-3699
        return create(np.zeros(0, dtype=v.dtype), m.rows, m.cols)
    k = total_key(v)
    n = v.shape[axis]
    idx = np.arange(n).reshape((n, 1) if axis == 0 else (1, n))
    run = np.maximum.accumulate(k, axis=axis) if want_max else np.minimum.accumulate(k, axis=axis)
    # an element starts a new run when its key is strictly better than everything before it
    if axis == 0:
        prev = np.vstack([np.full((1, v.shape[1]), np.iinfo(np.int64).min if want_max else np.iinfo(np.int64).max), run[:-1]])
    else:
        prev = np.hstack([np.full((v.shape[0], 1), np.iinfo(np.int64).min if want_max else np.iinfo(np.int64).max), run[:, :-1]])
    starts = (k > prev) if want_max else (k < prev)
    src = np.maximum.accumulate(np.where(starts, idx, 0), axis=axis)
    out = np.take_along_axis(v, src, axis=axis)
    return create(out.reshape(-1), m.rows, m.cols)


def cummax(m: OMat, axis: int) -> OMat:
    return _cum_extremum(m, axis, True)


def cummin(m: OMat, axis: int) -> OMat:
    return _cum_extremum(m, axis, False)


# ---------------------------------------------------------------------------------------
# mask-driven selection (Mat.scala:4008-4021 m(mask); :1305-1336 Mat.where)
# ---------------------------------------------------------------------------------------
def mask_select(m: OMat, mask: OMat) -> OMat:
    if (mask.rows, mask.cols) != (m.rows, m.cols):
        raise ValueError(f"Mask shape {mask.shape} must match matrix shape {m.shape}")
    sel = m.to2d()[mask.to2d().astype(bool)]  # boolean indexing walks row-major, like the reference's double loop
    return create(sel.copy(), 1, sel.size)


def where(cond: OMat, x, y) -> OMat:
    c = cond.to2d().astype(bool)
    if isinstance(x, OMat):
        if not ((cond.rows, cond.cols) == (x.rows, x.cols) == (y.rows, y.cols)):
            raise ValueError("where: shape mismatch")
        return create(np.where(c, x.to2d(), y.to2d()).reshape(-1), cond.rows, cond.cols)
    return create(np.where(c, np.float64(x), np.float64(y)).reshape(-1), cond.rows, cond.cols)


def mask_all(mask: OMat, axis=None):
    """Mat.scala:4907-5001."""
    v = mask.to2d().astype(bool)
    if axis is None:
        return bool(v.all())
    r = v.all(axis=axis)
    return create(r.reshape(-1), 1, mask.cols) if axis == 0 else create(r.reshape(-1), mask.rows, 1)


def mask_any(mask: OMat, axis=None):
    v = mask.to2d().astype(bool)
    if axis is None:
        return bool(v.any())
    r = v.any(axis=axis)
    return create(r.reshape(-1), 1, mask.cols) if axis == 0 else create(r.reshape(-1), mask.rows, 1)


# ---------------------------------------------------------------------------------------
# elementwise + broadcasting (Mat.scala:1989-2102,2154-2209,2514-2531,4513-4588)
# ---------------------------------------------------------------------------------------
_BIN = {
    "add": np.add,
    "sub": np.subtract,
    "mul": np.multiply,
    "div": np.divide,
}


def binary(op: str, a: OMat, b: OMat) -> OMat:
    """`a op b` with NumPy broadcasting: target = elementwise max of the shapes, each
    operand dimension equal to the target or 1 (Mat.scala:1989-2005, 2013-2082, 4820)."""
    rows, cols = max(a.rows, b.rows), max(a.cols, b.cols)
    av = broadcast_to(a, rows, cols).to2d()
    bv = broadcast_to(b, rows, cols).to2d()
    with np.errstate(all="ignore"):
        r = _BIN[op](av, bv)
    return create(r.reshape(-1), rows, cols)


def binary_scalar(op: str, a: OMat, s, scalar_left=False) -> OMat:
    """Mat op scalar (Mat.scala:2154-2209,2514-2531) / scalar op Mat (:1416-1419)."""
    v = a.to2d()
    s = v.dtype.type(s)
    with np.errstate(all="ignore"):
        r = _BIN[op](s, v) if scalar_left else _BIN[op](v, s)
    return create(r.reshape(-1), a.rows, a.cols)


def neg(a: OMat) -> OMat:  # unary_-  Mat.scala:2086-2102
    return create((-a.to2d()).reshape(-1), a.rows, a.cols)


def abs_(a: OMat) -> OMat:
    """abs = x < 0 ? -x : x  -- keeps -0.0 (Mat.scala:3557-3560)."""
    v = a.to2d()
    return create(np.where(v < 0, -v, v).reshape(-1), a.rows, a.cols)


def inplace_scalar(op: str, a: OMat, s) -> OMat:
    """`a :op= s`: mutate through the strides, return the receiver (Mat.scala:4513-4560)."""
    idx = a.index_grid()
    with np.errstate(all="ignore"):
        a.data[idx] = _BIN[op](a.data[idx], a.data.dtype.type(s))
    return a


def inplace_mat(op: str, a: OMat, b: OMat) -> OMat:
    """`a :+= b` / `a :-= b`: equal shapes required, no broadcast (Mat.scala:4562-4588)."""
    if a.shape != b.shape:
        raise ValueError("in-place Mat op requires equal shapes")
    idx = a.index_grid()
    a.data[idx] = _BIN[op](a.data[idx], b.to2d())
    return a


def java_exp(x):
    return np.exp(x)


def sigmoid(a: OMat) -> OMat:
    """x >= 0 ? 1/(1+exp(-x)) : (e = exp(x); e/(1+e)); always Mat[Double] (MatMathOps.scala:97-122)."""
    x = a.to2d().astype(np.float64)
    with np.errstate(all="ignore"):
        pos = 1.0 / (1.0 + np.exp(-x))
        e = np.exp(x)
        negv = e / (1.0 + e)
    return create(np.where(x >= 0, pos, negv).reshape(-1), a.rows, a.cols)


def relu(a: OMat) -> OMat:
    """Double fast path: Math.max(x, 0.0) -- NaN survives, -0.0 -> +0.0 (MatMathOps.scala:125-131);
    other element types: `if x > 0 then x else 0` (:133-146)."""
    x = a.to2d()
    if a.fast_d():
        r = np.where(np.isnan(x), x, np.where(x > 0, x, 0.0))
    else:
        r = np.where(x > 0, x, x.dtype.type(0))
    return create(r.reshape(-1), a.rows, a.cols)


def softmax(a: OMat, axis: int = 1) -> OMat:
    """Per lane: mx = x0; v > mx adopts; e = exp(x - mx); s = left fold; e / s
    (MatMathOps.scala:156-226).  Always Mat[Double]."""
    if axis not in (0, 1):
        raise ValueError("axis must be 0 (columns) or 1 (rows)")
    x = a.to2d().astype(np.float64)
    if axis == 0:
        x = x.T
    out = np.empty_like(x)
    for i in range(x.shape[0]):
        lane = x[i]
        mx = lane[0]
        # running max with `v > mx`: a NaN is adopted only if it is element 0
        if math.isnan(mx):
            pass
        else:
            with np.errstate(invalid="ignore"):
                cand = lane[~np.isnan(lane)]
            mx = cand.max() if cand.size else mx
        with np.errstate(all="ignore"):
            e = np.exp(lane - mx)
            s = np.add.accumulate(np.concatenate([[0.0], e]))[-1]
            out[i] = e / s
    if axis == 0:
        out = out.T
    return create(np.ascontiguousarray(out).reshape(-1), a.rows, a.cols)


# ---------------------------------------------------------------------------------------
# masks: IEEE semantics (Mat.scala:2564-2582,3948-4003,4214-4225,5157-5170)
# ---------------------------------------------------------------------------------------
_CMP = {
    "gt": np.greater,
    "lt": np.less,
    "gte": np.greater_equal,
    "lte": np.less_equal,
    "eq": np.equal,
    "ne": np.not_equal,
}


def compare_scalar(op: str, a: OMat, s) -> OMat:
    v = a.to2d().astype(np.float64)
    with np.errstate(invalid="ignore"):
        r = _CMP[op](v, np.float64(s))
    return create(r.reshape(-1), a.rows, a.cols)


def isnan(a):
    return create(np.isnan(a.to2d()).reshape(-1), a.rows, a.cols)


def isinf(a):
    return create(np.isinf(a.to2d()).reshape(-1), a.rows, a.cols)


def isfinite(a):
    return create(np.isfinite(a.to2d()).reshape(-1), a.rows, a.cols)


def mask_and(a, b):
    return create((a.to2d() & b.to2d()).reshape(-1), a.rows, a.cols)


def mask_or(a, b):
    return create((a.to2d() | b.to2d()).reshape(-1), a.rows, a.cols)


def mask_not(a):
    return create((~a.to2d()).reshape(-1), a.rows, a.cols)


# ---------------------------------------------------------------------------------------
# matmul: pinned semantics (MatmulPure.scala:9-15,78-110; Mat.scala:3236-3268 for Float)
# ---------------------------------------------------------------------------------------
def matmul_pure(a: OMat, b: OMat) -> OMat:
    """Every C cell = sequential k-sum from 0.0, multiply and add rounded separately
    (no FMA), in the element type (MatmulPure.scala:9-15; Float: Mat.scala:3236-3268)."""
    if a.cols != b.rows:
        raise ValueError("shape mismatch")
    av, bv = a.to2d(), b.to2d()
    c = np.zeros((a.rows, b.cols), dtype=av.dtype)
    with np.errstate(all="ignore"):
        for k in range(a.cols):
            c = c + av[:, k : k + 1] * bv[k : k + 1, :]
    return create(c.reshape(-1), a.rows, b.cols)


def matmul_blas(a: OMat, b: OMat) -> OMat:
    """The default `*@` route above the BLAS threshold: an unpinned dgemm/sgemm
    (Mat.scala:3300-3380); tolerance contract MatmulModeSuite.scala:66-83."""
    return from2d(a.to2d() @ b.to2d())


# ---------------------------------------------------------------------------------------
# LU / inverse / solve (Mat.scala:970-1001, 2979-3006, 3598-3634)
# ---------------------------------------------------------------------------------------
class SingularMatrix(ArithmeticError):
    pass


def lu_decompose(a2d: np.ndarray):
    n = a2d.shape[0]
    lu = np.array(a2d, dtype=a2d.dtype, copy=True)
    piv = list(range(n))
    for i in range(n):
        max_row, max_abs = i, abs(lu[i, i])
        for k in range(i + 1, n):
            v = abs(lu[k, i])
            if v > max_abs:
                max_abs, max_row = v, k
        if max_row != i:
            lu[[i, max_row]] = lu[[max_row, i]]
            piv[i], piv[max_row] = piv[max_row], piv[i]
        pivot = lu[i, i]
        if pivot == 0.0:
            raise SingularMatrix("Matrix is singular or nearly singular")
        for r in range(i + 1, n):
            f = lu[r, i] / pivot
            lu[r, i] = f
            lu[r, i + 1 :] = lu[r, i + 1 :] - f * lu[i, i + 1 :]
    return lu, piv


def _lu_solve_cols(lu, rhs):
    n = lu.shape[0]
    x = np.array(rhs, copy=True)
    for i in range(1, n):
        for k in range(i):
            x[i] = x[i] - lu[i, k] * x[k]
    for i in range(n - 1, -1, -1):
        for k in range(i + 1, n):
            x[i] = x[i] - lu[i, k] * x[k]
        x[i] = x[i] / lu[i, i]
    return x


def inverse(a: OMat) -> OMat:
    if a.rows != a.cols:
        raise ValueError("inverse requires square matrix")
    n = a.rows
    lu, piv = lu_decompose(a.to2d())
    rhs = np.zeros((n, n), dtype=lu.dtype)
    for i in range(n):
        rhs[i, piv[i]] = 1.0
    return from2d(_lu_solve_cols(lu, rhs))


def solve(a: OMat, b: OMat) -> OMat:
    if a.rows != a.cols:
        raise ValueError("solve requires square matrix")
    bc = b.T if (b.rows == 1 and b.cols == a.rows) else b
    if bc.rows != a.rows:
        raise ValueError("rhs rows must match")
    lu, piv = lu_decompose(a.to2d())
    rhs = bc.to2d()[piv, :]
    return from2d(_lu_solve_cols(lu, rhs))
