"""Host mirror of the reference's NumPy-compatible generator (`uni.data.NumPyRNG`, NumPyRNG.scala:22-160) and of
the global-generator entry points of `MatD` (Mat.scala:1461-1530: setSeed, nextRandLong / nextRandInt /
nextRandDouble, rand, randn, normal, uniform, randint).  The state is a POD held on the host (`um_rng`); all
matrix fills run on the device in libunimat.so (csrc/rng.cu) and consume the PCG64 stream exactly as the
reference's sequential loops do, so a seeded sequence of calls yields the values NumPy yields."""
from __future__ import annotations

import ctypes as C

from . import _lib as L
from ._lib import check, lib


class UmRng(C.Structure):
    _fields_ = [
        ("state_hi", C.c_uint64),
        ("state_lo", C.c_uint64),
        ("inc_hi", C.c_uint64),
        ("inc_lo", C.c_uint64),
        ("has_uint32", C.c_int32),
        ("uinteger", C.c_uint32),
    ]


class NumPyRNG:
    """`new NumPyRNG(seed)`; seed < 0 raises ValueError (IllegalArgumentException, NumPyRNG.scala:770-771)."""

    def __init__(self, seed: int = 0):
        self._g = UmRng()
        check(lib.um_rng_seed(C.byref(self._g), int(seed)))

    # -- scalar draws: host arithmetic on the same state (NumPyRNG.scala:58-74,82-110) --
    def nextLong(self) -> int:
        out = C.c_uint64(0)
        check(lib.um_rng_next_u64(C.byref(self._g), C.byref(out)))
        v = out.value
        return v - (1 << 64) if v >= (1 << 63) else v  # a JVM Long

    def nextDouble(self) -> float:
        return float((self.nextLong() & 0xFFFFFFFFFFFFFFFF) >> 11) / 9007199254740992.0

    def uniform(self, low: float, high: float) -> float:
        return low + self.nextDouble() * (high - low)

    def advance(self, draws: int) -> None:
        check(lib.um_rng_advance(C.byref(self._g), int(draws)))

    @property
    def state(self):
        """(state, inc) as Python ints -- comparable with numpy.random.PCG64(seed).state['state']."""
        g = self._g
        return ((g.state_hi << 64) | g.state_lo, (g.inc_hi << 64) | g.inc_lo)

    # -- device fills --
    def _fill(self, fn, m, *args):
        check(fn(C.byref(self._g), *args, m._ref()))
        return m

    def rand(self, rows: int, cols: int):
        from .mat import Mat
        return self._fill(lib.um_rng_rand, Mat._alloc(rows, cols, L.F64))

    def uniform_mat(self, low: float, high: float, rows: int, cols: int):
        from .mat import Mat
        return self._fill(lib.um_rng_uniform, Mat._alloc(rows, cols, L.F64), float(low), float(high))

    def randn(self, rows: int, cols: int = 1):
        from .mat import Mat
        return self._fill(lib.um_rng_randn, Mat._alloc(rows, cols, L.F64))

    def normal(self, mean: float, std: float, rows: int, cols: int):
        from .mat import Mat
        return self._fill(lib.um_rng_normal, Mat._alloc(rows, cols, L.F64), float(mean), float(std))

    def randint(self, low: int, high: int, rows: int, cols: int):
        from .mat import Mat
        return self._fill(lib.um_rng_randint, Mat._alloc(rows, cols, L.I32), int(low), int(high))


# `private[data] lazy val defaultRNG = new NumPyRNG(0)`, `@volatile var globalRNG` (Mat.scala:1460-1461)
_global: NumPyRNG | None = None


def global_rng() -> NumPyRNG:
    global _global
    if _global is None:
        _global = NumPyRNG(0)
    return _global


def set_seed(seed: int) -> None:
    """MatD.setSeed (Mat.scala:1465-1466)."""
    global _global
    _global = NumPyRNG(seed)
