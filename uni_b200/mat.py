"""Host-side mirror of uni's `Mat[T]` operator API over the libunimat C ABI.

The reference's host language is Scala 3 (no JVM in the build image), so this module plays
the role the Panama binding module plays there: it owns device buffers, carries the
`(offset, rows, cols, rs, cs, transposed)` descriptor of `MatData`
(src/main/scala/uni/data/Mat.scala:125-133), applies the reference's layout guard and
broadcasting / shape rules, and forwards every operator to one C-ABI call.  Names follow the
reference (`matmul`/`*@` -> `@`, `:+=` -> `+=`/`iadd`, `gt`/`lt`/..., `idxmin`, `softmax`).
Nothing here computes on the CPU: without the CUDA library + device every operator raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np

from . import _lib as L
from ._lib import UmView, check, lib

_NP = {L.F64: np.float64, L.F32: np.float32, L.BOOL: np.bool_, L.I32: np.int32}
_SIZE = {L.F64: 8, L.F32: 4, L.BOOL: 1, L.I32: 4}
_NAME = {L.F64: "Double", L.F32: "Float", L.BOOL: "Boolean", L.I32: "Int"}


def _dtype_of(np_dtype) -> int:
    d = np.dtype(np_dtype)
    if d == np.float64:
        return L.F64
    if d == np.float32:
        return L.F32
    if d == np.bool_:
        return L.BOOL
    if d == np.int32:
        return L.I32
    raise ValueError(f"unsupported element type {d}")


class DeviceBuffer:
    """Owner of one device allocation; views share it (views retain their parent)."""

    __slots__ = ("ptr", "nbytes", "__weakref__")

    def __init__(self, nbytes: int):
        p = C.c_void_p()
        check(lib.um_malloc(max(int(nbytes), 1), C.byref(p)))
        self.ptr = p.value
        self.nbytes = int(nbytes)

    def __del__(self):
        try:
            if self.ptr:
                lib.um_free(C.c_void_p(self.ptr))
                self.ptr = None
        except Exception:  # interpreter shutdown
            pass


class Mat:
    """`Mat[T]`: device buffer + stride descriptor.  T in {Double, Float, Boolean, Int}."""

    __slots__ = ("_buf", "_v")
    __array_priority__ = 1000

    # ------------------------------------------------------------------ construction
    def __init__(self, buf: DeviceBuffer, view: UmView):
        self._buf = buf
        self._v = view

    @staticmethod
    def _alloc(rows: int, cols: int, dtype: int) -> "Mat":
        if rows < 0 or cols < 0:
            raise ValueError("negative shape")
        buf = DeviceBuffer(rows * cols * _SIZE[dtype])
        v = UmView()
        check(lib.um_view_init(C.byref(v), C.c_void_p(buf.ptr), rows, cols, dtype))
        return Mat(buf, v)

    @staticmethod
    def _create(buf, rows, cols, dtype, transposed=False, offset=0, rs=-1, cs=-1) -> "Mat":
        """Internal.create (Mat.scala:170-189): default strides, then the layout guard."""
        v = UmView()
        v.data = buf.ptr
        v.offset = offset
        v.rows, v.cols = rows, cols
        v.rs = rs if rs >= 0 else (1 if transposed else cols)
        v.cs = cs if cs >= 0 else (rows if transposed else 1)
        v.dtype = dtype
        v.transposed = 1 if transposed else 0
        m = Mat(buf, v)
        w = C.c_int32(0)
        check(lib.um_view_is_weird(C.byref(v), C.byref(w)))
        return m.copy() if w.value else m

    @classmethod
    def from_numpy(cls, a, dtype=None) -> "Mat":
        a = np.asarray(a)
        if a.ndim == 1:
            a = a.reshape(-1, 1)  # MatD(array) is a column vector (MatFacades.scala:91)
        if a.ndim != 2:
            raise ValueError("need a 1-D or 2-D array")
        a = np.ascontiguousarray(a, dtype=_NP[dtype] if dtype is not None else a.dtype)
        dt = _dtype_of(a.dtype)
        m = cls._alloc(a.shape[0], a.shape[1], dt)
        if a.size:
            check(lib.um_h2d(C.c_void_p(m._buf.ptr), a.ctypes.data_as(C.c_void_p), a.nbytes))
        return m

    def to_numpy(self) -> np.ndarray:
        src = self if self.is_dense() else self.copy()
        out = np.empty((self.rows, self.cols), dtype=_NP[self.dtype])
        if out.size:
            p = src._buf.ptr + src._v.offset * _SIZE[self.dtype]
            check(lib.um_d2h(out.ctypes.data_as(C.c_void_p), C.c_void_p(p), out.nbytes))
        return out

    toArray = to_numpy

    # ------------------------------------------------------------------ descriptor
    @property
    def rows(self) -> int:
        return self._v.rows

    @property
    def cols(self) -> int:
        return self._v.cols

    @property
    def shape(self) -> Tuple[int, int]:
        return (self._v.rows, self._v.cols)

    @property
    def size(self) -> int:
        return self._v.rows * self._v.cols

    @property
    def dtype(self) -> int:
        return self._v.dtype

    @property
    def rs(self) -> int:
        return self._v.rs

    @property
    def cs(self) -> int:
        return self._v.cs

    @property
    def offset(self) -> int:
        return self._v.offset

    @property
    def transposed(self) -> bool:
        return bool(self._v.transposed)

    @property
    def isEmpty(self) -> bool:
        return self.size == 0

    def isContiguous(self) -> bool:  # Mat.scala:49-50
        return self._v.rs == self._v.cols and self._v.cs == 1

    def is_dense(self) -> bool:
        return self._v.cs == 1 and (self._v.rs == self._v.cols or self._v.rows <= 1)

    def _ref(self):
        return C.byref(self._v)

    def __repr__(self):
        return (f"Mat[{_NAME[self.dtype]}]({self.rows}x{self.cols}, offset={self.offset}, rs={self.rs}, "
                f"cs={self.cs}{', T' if self.transposed else ''})")

    # ------------------------------------------------------------------ views (zero copy)
    @property
    def T(self) -> "Mat":
        """O(1) transpose view (Mat.scala:198-199, 2107-2109)."""
        v = UmView()
        check(lib.um_view_transpose(self._ref(), C.byref(v)))
        return Mat(self._buf, v)

    transpose = T

    def slice(self, r0: int, r1: int, c0: int, c1: int) -> "Mat":
        """m.slice(r0 until r1, c0 until c1) (Mat.scala:1905-1927): a view, unless the layout
        guard (Mat.scala:1865-1876) classifies it as fragmented, then a contiguous copy."""
        v = UmView()
        needs = C.c_int32(0)
        check(lib.um_view_slice(self._ref(), r0, r1, c0, c1, C.byref(v), C.byref(needs)))
        m = Mat(self._buf, v)
        return m.copy() if needs.value else m

    def _raw_slice(self, r0, r1, c0, c1) -> "Mat":
        """Sub-view WITHOUT the layout guard: the write target of range updates / hstack."""
        v = UmView()
        check(lib.um_view_slice(self._ref(), r0, r1, c0, c1, C.byref(v), None))
        return Mat(self._buf, v)

    def broadcastTo(self, rows: int, cols: int) -> "Mat":  # Mat.scala:1933-1953
        v = UmView()
        check(lib.um_view_broadcast(self._ref(), rows, cols, C.byref(v)))
        return Mat(self._buf, v)

    def copy(self) -> "Mat":
        """matCopy (Mat.scala:1886-1899): fresh contiguous row-major copy."""
        out = Mat._alloc(self.rows, self.cols, self.dtype)
        check(lib.um_copy(self._ref(), out._ref()))
        return out

    matCopy = copy

    def astype(self, dtype: int) -> "Mat":
        out = Mat._alloc(self.rows, self.cols, dtype)
        check(lib.um_copy(self._ref(), out._ref()))
        return out

    def toDouble(self) -> "Mat":
        return self.astype(L.F64)

    def toFloat(self) -> "Mat":
        return self.astype(L.F32)

    def __getitem__(self, key):
        """m(r, c) -> element; m(rows, cols) with ranges / `::` -> view; m(mask) -> selected elements."""
        if isinstance(key, Mat):
            return self.select(key)
        if not isinstance(key, tuple) or len(key) != 2:
            raise ValueError("Mat indexing needs (row, col)")
        r, c = key
        if isinstance(r, (int, np.integer)) and isinstance(c, (int, np.integer)):
            r = r + self.rows if r < 0 else r
            c = c + self.cols if c < 0 else c
            out = C.c_double(0)
            check(lib.um_get(self._ref(), int(r), int(c), C.byref(out)))
            return bool(out.value) if self.dtype == L.BOOL else (int(out.value) if self.dtype == L.I32 else out.value)

        def rng(k, n):
            if isinstance(k, (int, np.integer)):
                k = k + n if k < 0 else k
                return int(k), int(k) + 1
            if isinstance(k, slice):
                if k.step not in (None, 1):
                    raise ValueError("strided ranges are not views in the reference")
                a, b, _ = k.indices(n)
                return a, max(a, b)
            raise ValueError("bad index")

        r0, r1 = rng(r, self.rows)
        c0, c1 = rng(c, self.cols)
        return self.slice(r0, r1, c0, c1)

    def __setitem__(self, key, value):
        r, c = key
        if isinstance(r, (int, np.integer)) and isinstance(c, (int, np.integer)):
            check(lib.um_set(self._ref(), int(r), int(c), float(value)))
            return
        def rng(k, n):
            if isinstance(k, (int, np.integer)):
                k = k + n if k < 0 else k
                return int(k), int(k) + 1
            a, b, _ = k.indices(n)
            return a, max(a, b)

        r0, r1 = rng(r, self.rows)
        c0, c1 = rng(c, self.cols)
        dst = self._raw_slice(r0, r1, c0, c1)  # range update writes through (Mat.update)
        if isinstance(value, Mat):
            check(lib.um_copy(value.broadcastTo(dst.rows, dst.cols)._ref(), dst._ref()))
        else:
            check(lib.um_fill(dst._ref(), float(value)))

    # ------------------------------------------------------------------ elementwise
    def _binary(self, op: int, other, reverse=False, out: Optional["Mat"] = None) -> "Mat":
        if isinstance(other, Mat):
            a, b = (other, self) if reverse else (self, other)
            if a.dtype != b.dtype:
                raise ValueError("operands must share an element type")
            rows, cols = C.c_int64(0), C.c_int64(0)
            check(lib.um_broadcast_shape(a._ref(), b._ref(), C.byref(rows), C.byref(cols)))
            res = out if out is not None else Mat._alloc(rows.value, cols.value, a.dtype)
            check(lib.um_binary(op, a._ref(), b._ref(), res._ref()))
            return res
        s = float(other)
        res = out if out is not None else Mat._alloc(self.rows, self.cols, self.dtype)
        check(lib.um_binary_scalar(op, self._ref(), s, 1 if reverse else 0, res._ref()))
        return res

    def __add__(self, o): return self._binary(L.ADD, o)
    def __radd__(self, o): return self._binary(L.ADD, o, reverse=True)
    def __sub__(self, o): return self._binary(L.SUB, o)
    def __rsub__(self, o): return self._binary(L.SUB, o, reverse=True)
    def __mul__(self, o): return self._binary(L.MUL, o)      # `*` / `*:*` are elementwise (Mat.scala:4808)
    def __rmul__(self, o): return self._binary(L.MUL, o, reverse=True)
    def __truediv__(self, o): return self._binary(L.DIV, o)  # Mat.scala:4820-4822, 2514-2531
    def __rtruediv__(self, o): return self._binary(L.DIV, o, reverse=True)
    def hadamard(self, o): return self._binary(L.MUL, o)
    def maximum(self, o): return self._binary(L.MAXIMUM, o)
    def minimum(self, o): return self._binary(L.MINIMUM, o)

    def __pow__(self, p):  # `m ~^ p` = math.pow (Mat.scala:4456-4472)
        return self._binary(L.POW, p)

    def power(self, n: int) -> "Mat":  # repeated multiply (Mat.scala:4477-4496)
        if n == 2:
            return self._unary(L.SQUARE)
        r = self.copy()
        for _ in range(int(n) - 1):
            r = r * self
        return r

    def _inplace(self, op: int, other) -> "Mat":
        """`:+= :-= :*= :/=`: mutate the receiver through its strides and return it
        (Mat.scala:4513-4588).  The Mat form requires EQUAL shapes -- no broadcast (:4565)."""
        if isinstance(other, Mat):
            if other.shape != self.shape:
                raise ValueError(f"in-place op requires equal shapes, got {self.shape} and {other.shape}")
            return self._binary(op, other, out=self)
        return self._binary(op, other, out=self)

    def __iadd__(self, o): return self._inplace(L.ADD, o)
    def __isub__(self, o): return self._inplace(L.SUB, o)
    def __imul__(self, o): return self._inplace(L.MUL, o)
    def __itruediv__(self, o): return self._inplace(L.DIV, o)
    iadd, isub, imul, idiv = __iadd__, __isub__, __imul__, __itruediv__

    def _unary(self, op: int, out_dtype: Optional[int] = None) -> "Mat":
        res = Mat._alloc(self.rows, self.cols, self.dtype if out_dtype is None else out_dtype)
        check(lib.um_unary(op, self._ref(), res._ref()))
        return res

    def __neg__(self): return self._unary(L.NEG)
    def abs(self): return self._unary(L.ABS)
    def exp(self): return self._unary(L.EXP, L.F64)
    def sqrt(self): return self._unary(L.SQRT)
    def log(self): return self._unary(L.LOG, L.F64)
    def tanh(self): return self._unary(L.TANH, L.F64)

    def sigmoid(self) -> "Mat":
        """MatMathOps.scala:97-122 -- always Mat[Double]."""
        return self._unary(L.SIGMOID, L.F64)

    def relu(self) -> "Mat":
        """MatMathOps.scala:125-146 -- keeps T."""
        return self._unary(L.RELU)

    def softmax(self, axis: int = 1) -> "Mat":
        """MatMathOps.scala:156-226 -- default axis 1, always Mat[Double]."""
        if axis not in (0, 1):
            raise ValueError("axis must be 0 (columns) or 1 (rows)")
        res = Mat._alloc(self.rows, self.cols, L.F64)
        check(lib.um_softmax(self._ref(), axis, res._ref()))
        return res

    # ------------------------------------------------------------------ masks
    def _cmp(self, op: int, other) -> "Mat":
        if isinstance(other, Mat):
            rows, cols = C.c_int64(0), C.c_int64(0)
            check(lib.um_broadcast_shape(self._ref(), other._ref(), C.byref(rows), C.byref(cols)))
            res = Mat._alloc(rows.value, cols.value, L.BOOL)
            check(lib.um_compare(op, self._ref(), other._ref(), res._ref()))
            return res
        res = Mat._alloc(self.rows, self.cols, L.BOOL)
        check(lib.um_compare_scalar(op, self._ref(), float(other), res._ref()))
        return res

    def gt(self, o): return self._cmp(L.GT, o)
    def lt(self, o): return self._cmp(L.LT, o)
    def gte(self, o): return self._cmp(L.GTE, o)
    def lte(self, o): return self._cmp(L.LTE, o)
    def eq(self, o): return self._cmp(L.EQ, o)      # `:==`
    def ne(self, o): return self._cmp(L.NE, o)      # `:!=`
    __gt__, __lt__, __ge__, __le__ = gt, lt, gte, lte

    def _classify(self, op):
        res = Mat._alloc(self.rows, self.cols, L.BOOL)
        check(lib.um_classify(op, self._ref(), res._ref()))
        return res

    def isnan(self): return self._classify(L.ISNAN)
    def isinf(self): return self._classify(L.ISINF)
    def isfinite(self): return self._classify(L.ISFINITE)

    def _logic(self, op, other):
        res = Mat._alloc(self.rows, self.cols, L.BOOL)
        check(lib.um_mask_logic(op, self._ref(), other._ref() if other is not None else None, res._ref()))
        return res

    def __and__(self, o): return self._logic(L.AND, o)   # `&&`
    def __or__(self, o): return self._logic(L.OR, o)     # `||`
    def __invert__(self): return self._logic(L.NOT, None)  # `!`

    def _mask_reduce(self, want_all: int, axis: Optional[int]):
        if axis is None:
            out = C.c_int32(0)
            check(lib.um_mask_reduce(want_all, self._ref(), -1, None, C.byref(out)))
            return bool(out.value)
        res = Mat._alloc(1 if axis == 0 else self.rows, self.cols if axis == 0 else 1, L.BOOL)
        check(lib.um_mask_reduce(want_all, self._ref(), axis, res._ref(), None))
        return res

    def all(self, axis: Optional[int] = None): return self._mask_reduce(1, axis)
    def any(self, axis: Optional[int] = None): return self._mask_reduce(0, axis)

    # ------------------------------------------------------------------ mask-driven selection
    def select(self, mask: "Mat") -> "Mat":
        """m(mask) (Mat.scala:4008-4021): the elements where `mask` is true, row-major order, as a 1 x k row."""
        if not isinstance(mask, Mat) or mask.dtype != L.BOOL:
            raise ValueError("mask select needs a Mat[Boolean]")
        if mask.shape != self.shape:
            raise ValueError(f"Mask shape {mask.shape} must match matrix shape {self.shape}")
        k = C.c_int64(0)
        check(lib.um_mask_count(mask._ref(), C.byref(k)))
        res = Mat._alloc(1, k.value, self.dtype)
        check(lib.um_mask_select(self._ref(), mask._ref(), res._ref()))
        return res

    @staticmethod
    def where(condition: "Mat", x, y) -> "Mat":
        """Mat.where(condition, x, y) (Mat.scala:1305-1336): x and y both matrices of the condition's shape,
        or both scalars (result Mat[Double])."""
        if not isinstance(condition, Mat) or condition.dtype != L.BOOL:
            raise ValueError("where: condition must be a Mat[Boolean]")
        if isinstance(x, Mat) and isinstance(y, Mat):
            if not (condition.shape == x.shape == y.shape):
                raise ValueError(f"where: shape mismatch: condition={condition.shape} x={x.shape} y={y.shape}")
            if x.dtype != y.dtype:
                raise ValueError("where: x and y must have the same element type")
            res = Mat._alloc(x.rows, x.cols, x.dtype)
            check(lib.um_where(condition._ref(), x._ref(), y._ref(), res._ref()))
            return res
        if isinstance(x, Mat) or isinstance(y, Mat):
            raise ValueError("where: x and y must both be matrices or both be scalars")
        res = Mat._alloc(condition.rows, condition.cols, L.F64)
        check(lib.um_where_scalar(condition._ref(), float(x), float(y), res._ref()))
        return res

    # ------------------------------------------------------------------ running folds
    def _cumulative(self, kind: int, axis: Optional[int]) -> "Mat":
        if axis is None:
            res = Mat._alloc(1, self.rows * self.cols, self.dtype)
            check(lib.um_cumulative(kind, self._ref(), -1, res._ref()))
            return res
        if axis not in (0, 1):
            raise ValueError(f"axis must be 0 or 1, got {axis}")
        res = Mat._alloc(self.rows, self.cols, self.dtype)
        check(lib.um_cumulative(kind, self._ref(), axis, res._ref()))
        return res

    def cumsum(self, axis: Optional[int] = None): return self._cumulative(L.CUMSUM, axis)   # Mat.scala:3635-3690
    def cummax(self, axis: int): return self._cumulative(L.CUMMAX, axis)                    # Mat.scala:3692-3717
    def cummin(self, axis: int): return self._cumulative(L.CUMMIN, axis)                    # Mat.scala:3719-3744

    # ------------------------------------------------------------------ reductions
    def _reduce(self, kind: int, axis: Optional[int]):
        if axis is None:
            out = C.c_double(0)
            check(lib.um_reduce_all(kind, self._ref(), C.byref(out)))
            return out.value
        if axis not in (0, 1):
            raise ValueError(f"axis must be 0 or 1, got {axis}")
        res = Mat._alloc(1 if axis == 0 else self.rows, self.cols if axis == 0 else 1, self.dtype)
        check(lib.um_reduce_axis(kind, self._ref(), axis, res._ref()))
        return res

    def sum(self, axis: Optional[int] = None): return self._reduce(L.SUM, axis)
    def mean(self, axis: Optional[int] = None): return self._reduce(L.MEAN, axis)
    def std(self, axis: Optional[int] = None): return self._reduce(L.STD, axis)       # population
    def var(self, axis: Optional[int] = None): return self._reduce(L.VAR, axis)       # population
    def variance(self): return self._reduce(L.VAR, None)
    def min(self, axis: Optional[int] = None): return self._reduce(L.MIN, axis)
    def max(self, axis: Optional[int] = None): return self._reduce(L.MAX, axis)
    def rowSums(self): return self.sum(1)
    def colSums(self): return self.sum(0)
    def rowMeans(self): return self.mean(1)
    def colMeans(self): return self.mean(0)

    def _arg(self, want_max: int) -> Tuple[int, int]:
        r, c = C.c_int64(0), C.c_int64(0)
        check(lib.um_arg_all(want_max, self._ref(), C.byref(r), C.byref(c)))
        return (r.value, c.value)

    def argmin(self): return self._arg(0)   # (row, col), Mat.scala:2281-2334
    def argmax(self): return self._arg(1)

    def _idx(self, want_max: int, axis: int) -> "Mat":
        if axis not in (0, 1):
            raise ValueError(f"axis must be 0 or 1, got {axis}")
        res = Mat._alloc(1 if axis == 0 else self.rows, self.cols if axis == 0 else 1, L.I32)
        check(lib.um_idx_axis(want_max, self._ref(), axis, res._ref()))
        return res

    def idxmin(self, axis: int): return self._idx(0, axis)   # MatPandasOps.scala:20-41
    def idxmax(self, axis: int): return self._idx(1, axis)   # MatPandasOps.scala:43-64

    # ------------------------------------------------------------------ linear algebra
    def matmul(self, other: "Mat") -> "Mat":
        """`a *@ b` (Mat.scala:2815-2881): shape mismatch -> ValueError (IllegalArgumentException)."""
        if not isinstance(other, Mat):
            raise ValueError("matmul needs a Mat")
        if self.cols != other.rows:
            raise ValueError(f"matmul shape mismatch: {self.shape} *@ {other.shape}")
        res = Mat._alloc(self.rows, other.cols, self.dtype)
        check(lib.um_matmul(self._ref(), other._ref(), res._ref()))
        return res

    __matmul__ = matmul
    dot = matmul
    matmulPure = matmul   # the device has one route; both escape hatches land on it
    matmulBlas = matmul

    def saveCSV(self, path, sep: str = ",", nanAs: str = "NaN") -> None:  # Mat.scala:90-116
        from . import io
        io.saveCSV(self, path, sep, nanAs)

    writeCsv = saveCSV

    def inverse(self) -> "Mat":  # Mat.scala:2979-3006
        if self.rows != self.cols:
            raise ValueError(f"inverse requires square matrix, got {self.shape}")
        src = self if self.dtype == L.F64 else self.astype(L.F64)
        res = Mat._alloc(self.rows, self.cols, L.F64)
        check(lib.um_inverse(src._ref(), res._ref()))
        return res if self.dtype == L.F64 else res.astype(self.dtype)

    def solve(self, b: "Mat") -> "Mat":  # Mat.scala:3598-3634
        if self.rows != self.cols:
            raise ValueError(f"solve requires square matrix, got {self.shape}")
        bc = b.T if (b.rows == 1 and b.cols == self.rows) else b
        if bc.rows != self.rows:
            raise ValueError(f"bCol.rows {bc.rows} must match matrix rows {self.rows}")
        res = Mat._alloc(bc.rows, bc.cols, L.F64)
        check(lib.um_solve(self._ref(), bc._ref(), res._ref()))
        return res


class _Facade:
    """MatFacade factories (MatFacades.scala:41-147): zeros / ones / full / eye / randn / ..."""

    def __init__(self, dtype: int):
        self.dtype = dtype

    def __call__(self, data, rows: Optional[int] = None, cols: Optional[int] = None) -> Mat:
        a = np.asarray(data, dtype=_NP[self.dtype])
        if rows is not None:
            a = a.reshape(rows, cols)
        return Mat.from_numpy(a)

    def zeros(self, rows, cols):
        return self.full(rows, cols, 0.0)

    def ones(self, rows, cols):
        return self.full(rows, cols, 1.0)

    def full(self, rows, cols, value):
        m = Mat._alloc(rows, cols, self.dtype)
        check(lib.um_fill(m._ref(), float(value)))
        return m

    def eye(self, n):
        m = Mat._alloc(n, n, self.dtype)
        check(lib.um_eye(m._ref()))
        return m

    # -- random factories (Mat.scala:1461-1530).  Without `seed` they draw from the global NumPy-compatible PCG64
    #    generator (`MatD.setSeed`), element by element in row-major order like the reference.  With an explicit
    #    `seed=` they produce a synthetic counter-based panel instead (bench / test inputs; NOT the PCG64 stream). --
    def setSeed(self, seed: int) -> None:
        from . import rng
        rng.set_seed(seed)

    def nextRandLong(self) -> int:
        from . import rng
        return rng.global_rng().nextLong()

    def nextRandInt(self) -> int:  # globalRNG.nextInt() & 0xFFFFFFFFL: the LOW 32 bits (Mat.scala:1469, NumPyRNG.scala:64-66)
        from . import rng
        return rng.global_rng().nextLong() & 0xFFFFFFFF

    def nextRandDouble(self) -> float:
        from . import rng
        return rng.global_rng().nextDouble()

    def randn(self, rows, cols=None, seed: Optional[int] = None):
        if seed is not None:
            m = Mat._alloc(rows, cols, self.dtype)
            check(lib.um_randn(m._ref(), seed))
            return m
        from . import rng
        m = rng.global_rng().randn(rows, 1 if cols is None else cols)  # randn(n) is an n x 1 column (Mat.scala:1480-1482)
        return m if self.dtype == L.F64 else m.astype(self.dtype)

    rnorm = randn

    def rand(self, rows, cols, seed: Optional[int] = None):
        if seed is not None:
            m = Mat._alloc(rows, cols, self.dtype)
            check(lib.um_rand(m._ref(), seed))
            return m
        from . import rng
        m = rng.global_rng().rand(rows, cols)
        return m if self.dtype == L.F64 else m.astype(self.dtype)

    def normal(self, mean, std, rows, cols):
        from . import rng
        return rng.global_rng().normal(mean, std, rows, cols)

    def uniform(self, low, high, rows, cols):
        from . import rng
        return rng.global_rng().uniform_mat(low, high, rows, cols)

    def randint(self, low, high, rows=None, cols=None):
        from . import rng
        if rows is None:  # randint(low, high): one Int (Mat.scala:1520-1522)
            return int(rng.global_rng().randint(low, high, 1, 1).to_numpy()[0, 0])
        return rng.global_rng().randint(low, high, rows, cols)

    def hstack(self, *mats):
        rows = mats[0].rows
        out = Mat._alloc(rows, sum(m.cols for m in mats), self.dtype)
        c = 0
        for m in mats:
            check(lib.um_copy(m._ref(), out._raw_slice(0, rows, c, c + m.cols)._ref()))
            c += m.cols
        return out

    def vstack(self, *mats):
        cols = mats[0].cols
        out = Mat._alloc(sum(m.rows for m in mats), cols, self.dtype)
        r = 0
        for m in mats:
            check(lib.um_copy(m._ref(), out._raw_slice(r, r + m.rows, 0, cols)._ref()))
            r += m.rows
        return out


MatD = _Facade(L.F64)
MatF = _Facade(L.F32)


MATF_PATH_AUTO, MATF_PATH_CUDA_CORES, MATF_PATH_TENSOR = 0, 1, 2


def set_matf_matmul_path(path: int) -> None:
    """FP32 `@` kernel choice (um_matmul_set_f32_path; the counterpart of the reference's uni.mat.blas
    switch, Mat.scala:299-335): 0 = automatic, 1 = FFMA kernel only, 2 = tcgen05 3xTF32 kernel with the hi / lo split in shared memory or ValueError,
    3 = the same kernel with the operands split once in global memory (what automatic picks from 2^30 multiply-adds up)."""
    check(lib.um_matmul_set_f32_path(int(path)))


def set_softmax_axis0_path(path: int) -> None:
    """softmax(0) kernel choice (um_softmax_set_axis0_path): 0 = automatic, 1 = two-read kernels, 2 = single-read strip
    kernel or ValueError."""
    check(lib.um_softmax_set_axis0_path(int(path)))


def set_matd_matmul_tile(tile: int) -> None:
    """FP64 `@` CTA tile (um_matmul_set_f64_tile): 0 = automatic, 64 or 128 forced (tests, profiling)."""
    check(lib.um_matmul_set_f64_tile(int(tile)))


def init(device: int = 0) -> None:
    check(lib.um_init(device))


def sync() -> None:
    check(lib.um_sync())
