"""CPU restatement of the reference's NumPy-compatible generator -- TEST INFRASTRUCTURE ONLY (imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline leg; the product path is csrc/rng.cu).

Follows /root/reference/src/main/scala/uni/data/NumPyRNG.scala: SeedSequence :557-570,:635-718, PCG64 seeding
:758-786, step/output :45-56, nextDouble :68-71, uniform :73-74, nextBoundedInt :82-110, ziggurat randn :124-160;
and the Mat factories src/main/scala/uni/data/Mat.scala:1473-1530 (element order = row-major, one draw sequence).
Pure-Python integers: exact, slow (about 1 us per draw) -- sized for tests.

Pinned by tests/test_oracle_rng.py against numpy.random.PCG64 / default_rng outputs committed under
tests/golden/rng_numpy/ (tests/golden/gen_rng_numpy.py) -- the library NumPyRNG.scala itself is pinned to.
The ziggurat tables are NumPy's (tools/gen_ziggurat_tables.py reads them out of the installed binary and checks them)."""
import math
import os
import re

import numpy as np

M64 = (1 << 64) - 1
M128 = (1 << 128) - 1
PCG_MULT = (0x2360ED051FC65DA4 << 64) | 0x4385DF649FCCF645
ZIG_R = 3.6541528853610088
ZIG_INV_R = 0.27366123732975827

_TABLES = None


def ziggurat_tables():
    """(ki, wi, fi) parsed from the generated header (exact bit patterns)."""
    global _TABLES
    if _TABLES is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "uni_b200", "csrc", "ziggurat_tables.cuh")
        text = open(path).read()
        out = []
        for name in ("h_zig_ki", "h_zig_wi_bits", "h_zig_fi_bits"):
            body = text[text.index(name + "[256]"):]
            body = body[body.index("{") + 1:body.index("}")]
            vals = [int(h, 16) for h in re.findall(r"0x([0-9A-Fa-f]{16})ULL", body)]
            assert len(vals) == 256
            out.append(np.array(vals, dtype=np.uint64))
        _TABLES = ([int(v) for v in out[0]], [float(v) for v in out[1].view(np.float64)], [float(v) for v in out[2].view(np.float64)])
    return _TABLES


def _u32(x):
    return x & 0xFFFFFFFF


def seed_sequence_words(seed: int):
    """SeedSequence(seed).generate_state(4, uint64) -- NumPyRNG.scala:484-501,557-570,635-718."""
    if seed < 0:
        raise ValueError(f"seed [{seed}] must be non-negative")
    ent = []
    v = seed
    while True:
        ent.append(_u32(v))
        v >>= 32
        if v == 0:
            break
    hc = [0x43B0D7E5]

    def hashmix(value):
        value = _u32(value ^ hc[0])
        hc[0] = _u32(hc[0] * 0x931E8875)
        value = _u32(value * hc[0])
        return _u32(value ^ (value >> 16))

    def mix(x, y):
        r = _u32(0xCA01F9DD * x - 0x4973F715 * y)
        return _u32(r ^ (r >> 16))

    pool = [hashmix(ent[i] if i < len(ent) else 0) for i in range(4)]
    for src in range(4):
        for dst in range(4):
            if src != dst:
                pool[dst] = mix(pool[dst], hashmix(pool[src]))
    for i in range(4, len(ent)):  # entropy longer than the pool (seeds >= 2^128: not reachable from a Long)
        for dst in range(4):
            pool[dst] = mix(pool[dst], hashmix(ent[i]))
    hb = 0x8B51F9DD
    w32 = []
    for i in range(8):
        d = pool[i % 4]
        d = _u32(d ^ hb)
        hb = _u32(hb * 0x58F38DED)
        d = _u32(d * hb)
        w32.append(_u32(d ^ (d >> 16)))
    return [w32[2 * i] | (w32[2 * i + 1] << 32) for i in range(4)]


class NumPyRNG:
    """class NumPyRNG(seed) -- NumPyRNG.scala:22-160."""

    def __init__(self, seed: int = 0):
        w = seed_sequence_words(seed)
        initstate = (w[0] << 64) | w[1]
        initseq = (w[2] << 64) | w[3]
        self.inc = ((initseq << 1) | 1) & M128          # pcgInit, NumPyRNG.scala:761-767
        s = self.inc                                    # step(0)
        s = (s + initstate) & M128
        self.state = (s * PCG_MULT + self.inc) & M128
        self.use_upper = False
        self.cached = 0

    def nextLong(self) -> int:                          # :45-61 (returned unsigned)
        self.state = (self.state * PCG_MULT + self.inc) & M128
        hi, lo = self.state >> 64, self.state & M64
        x = hi ^ lo
        r = hi >> 58
        return ((x >> r) | (x << (64 - r))) & M64 if r else x

    def nextDouble(self) -> float:                      # :68-71
        return (self.nextLong() >> 11) / 9007199254740992.0

    def uniform(self, low: float, high: float) -> float:  # :73-74
        return low + self.nextDouble() * (high - low)

    def nextBoundedInt(self, bound: int) -> int:        # :82-110
        if bound <= 0:
            raise ValueError("bound must be positive")
        if self.use_upper:
            raw = self.cached
            bits = (raw >> 32) & 0xFFFFFFFF
        else:
            raw = self.nextLong()
            self.cached = raw
            bits = raw & 0xFFFFFFFF
        self.use_upper = not self.use_upper
        return ((bits * bound) >> 32) & 0xFFFFFFFF

    def randn(self) -> float:                           # :124-160
        ki, wi, fi = ziggurat_tables()
        while True:
            r = self.nextLong()
            idx = r & 0xFF
            r >>= 8
            sign = -1.0 if (r & 1) else 1.0
            rabs = (r >> 1) & 0x000FFFFFFFFFFFFF
            x = float(rabs) * wi[idx]
            if rabs < ki[idx]:
                return sign * x
            if idx == 0:
                xx = -ZIG_INV_R * math.log1p(-self.nextDouble())
                yy = -math.log1p(-self.nextDouble())
                while yy + yy <= xx * xx:
                    xx = -ZIG_INV_R * math.log1p(-self.nextDouble())
                    yy = -math.log1p(-self.nextDouble())
                return (-1.0 if ((rabs >> 8) & 1) else 1.0) * (ZIG_R + xx)
            if (fi[idx - 1] - fi[idx]) * self.nextDouble() + fi[idx] < math.exp(-0.5 * x * x):
                return sign * x

    # Mat factories (Mat.scala:1473-1530): row-major fill from one draw sequence
    def rand_mat(self, rows, cols):
        return np.array([self.nextDouble() for _ in range(rows * cols)], dtype=np.float64).reshape(rows, cols)

    def randn_mat(self, rows, cols):
        return np.array([self.randn() for _ in range(rows * cols)], dtype=np.float64).reshape(rows, cols)

    def normal_mat(self, mean, std, rows, cols):
        return np.array([mean + std * self.randn() for _ in range(rows * cols)], dtype=np.float64).reshape(rows, cols)

    def uniform_mat(self, low, high, rows, cols):
        return np.array([self.uniform(low, high) for _ in range(rows * cols)], dtype=np.float64).reshape(rows, cols)

    def randint_mat(self, low, high, rows, cols):
        def wrap(v):  # JVM Int arithmetic
            v &= 0xFFFFFFFF
            return v - (1 << 32) if v >= (1 << 31) else v
        bound = wrap(high - low)
        return np.array([wrap(low + wrap(self.nextBoundedInt(bound))) for _ in range(rows * cols)], dtype=np.int32).reshape(rows, cols)
