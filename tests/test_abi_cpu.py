"""CPU-side checks of the C-ABI boundary: the library loads, exports every symbol the header
declares, the host-side view arithmetic follows the reference's layout rules, and the product
path fails loudly without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import mat_oracle as mo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "unimat.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(um_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from uni_b200 import _lib
    names = header_symbols()
    assert len(names) >= 50
    for n in names:
        assert hasattr(_lib.lib, n), f"{n} declared in include/unimat.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == names, "binding and header out of sync"
    assert _lib.lib.um_version() == 1


def _c_layout(t):
    """C parameter type -> the Panama ValueLayout the Scala binding must name for it."""
    if "*" in t or "[" in t:   # arrays decay to pointers
        return "ADDRESS"
    t = re.sub(r"\bconst\b", "", t).strip()
    if t.startswith(("long long", "unsigned long long")):
        return "JAVA_LONG"
    return {"int32_t": "JAVA_INT", "uint32_t": "JAVA_INT", "int": "JAVA_INT", "int64_t": "JAVA_LONG", "uint64_t": "JAVA_LONG",
            "size_t": "JAVA_LONG", "double": "JAVA_DOUBLE", "float": "JAVA_FLOAT"}[t.split()[0]]


def test_scala_binding_matches_header():
    """scala/uni/cuda/UniMatLib.scala cannot be compiled here (no JVM), so it is parsed: every prototype of the header
    has exactly one downcall handle, with the same number of arguments and the matching Panama layout for each."""
    src = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "unimat.h")).read(), flags=re.S)
    protos = {}
    for ret, name, args in re.findall(r"([A-Za-z_0-9\* ]+?)\b(um_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", src):
        a = [] if args.strip() in ("", "void") else [_c_layout(x.strip()) for x in args.split(",")]
        protos[name] = ("ADDRESS" if "*" in ret else "JAVA_INT", a)
    assert sorted(protos) == header_symbols()
    sc = open(os.path.join(ROOT, "scala", "uni", "cuda", "UniMatLib.scala")).read()
    handles = re.findall(r'val\s+(um_[a-z0-9_]+)\s*=\s*h\("(um_[a-z0-9_]+)",\s*(I\(([^)]*)\)|FunctionDescriptor\.of\(([^)]*)\))\)', sc)
    seen = {}
    for val, name, _, iargs, fargs in handles:
        assert val == name, f"handle {val} binds {name}"
        assert name not in seen, f"{name} bound twice"
        if fargs:
            parts = [x.strip() for x in fargs.split(",") if x.strip()]
            seen[name] = (parts[0], parts[1:])
        else:
            seen[name] = ("JAVA_INT", [x.strip() for x in iargs.split(",") if x.strip()])
    assert sorted(seen) == sorted(protos), (sorted(set(protos) - set(seen)), sorted(set(seen) - set(protos)))
    for name, sig in protos.items():
        assert seen[name] == sig, f"{name}: header {sig} vs Scala {seen[name]}"


def test_struct_layout_matches_header():
    from uni_b200._lib import UmTprfOut, UmView
    assert C.sizeof(UmView) == 56 and UmView.rows.offset == 16 and UmView.dtype.offset == 48
    assert C.sizeof(UmTprfOut) == 80 and UmTprfOut.r_squared.offset == 64


def _view(rows, cols):
    from uni_b200._lib import UmView, check, lib
    v = UmView()
    check(lib.um_view_init(C.byref(v), None, rows, cols, 0))
    return v


@pytest.mark.parametrize("shape", [(3, 5), (5, 3), (8, 9), (9, 8), (64, 64), (1, 7), (7, 1)])
def test_view_arithmetic_matches_reference_rules(shape):
    from uni_b200._lib import UmView, check, lib
    r, c = shape
    rng = np.random.default_rng(r * 100 + c)
    o = mo.create(np.arange(r * c, dtype=np.float64), r, c)
    v = _view(r, c)
    for _ in range(40):
        r0 = int(rng.integers(0, r)); r1 = int(rng.integers(r0, r + 1))
        c0 = int(rng.integers(0, c)); c1 = int(rng.integers(c0, c + 1))
        for transposed in (False, True):
            src_o, src_v = (o.T, UmView()) if transposed else (o, v)
            if transposed:
                check(lib.um_view_transpose(C.byref(v), C.byref(src_v)))
                a0, a1, b0, b1 = c0, c1, r0, r1
            else:
                a0, a1, b0, b1 = r0, r1, c0, c1
            s = UmView(); need = C.c_int32(0)
            check(lib.um_view_slice(C.byref(src_v), a0, a1, b0, b1, C.byref(s), C.byref(need)))
            want = mo.slice_(src_o, a0, a1, b0, b1)
            copied = want.data is not src_o.data
            assert bool(need.value) == copied
            if not copied:
                assert (s.offset, s.rows, s.cols, s.rs, s.cs, bool(s.transposed)) == (
                    want.offset, want.rows, want.cols, want.rs, want.cs, want.transposed)
    b = UmView()
    if r == 1 or c == 1:
        check(lib.um_view_broadcast(C.byref(v), 7 if r == 1 else r, 7 if c == 1 else c, C.byref(b)))
        assert (b.rs == 0) == (r == 1) and (b.cs == 0) == (c == 1)
    else:
        assert lib.um_view_broadcast(C.byref(v), r + 1, c, C.byref(b)) == 1  # UM_ERR_ARG -> IllegalArgumentException
    assert lib.um_view_slice(C.byref(v), 0, r + 1, 0, c, C.byref(b), None) == 1


def test_broadcast_shape_rule():
    from uni_b200._lib import check, lib
    rows, cols = C.c_int64(), C.c_int64()
    check(lib.um_broadcast_shape(C.byref(_view(5, 1)), C.byref(_view(1, 4)), C.byref(rows), C.byref(cols)))
    assert (rows.value, cols.value) == (5, 4)       # N x 1 (+) 1 x M -> N x M (Mat.scala:1989-2005)
    assert lib.um_broadcast_shape(C.byref(_view(5, 3)), C.byref(_view(4, 3)), C.byref(rows), C.byref(cols)) == 1


def test_no_silent_cpu_fallback():
    import uni_b200
    if uni_b200.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(uni_b200.UnimatError):
        uni_b200.MatD.zeros(2, 2)
    with pytest.raises(uni_b200.UnimatError):
        uni_b200.Mat.from_numpy(np.ones((2, 2)))


def test_shard_ranges_cover_and_align():
    from uni_b200.dist import packed_len, shard_range
    for n in (1, 63, 64, 65, 1000, 262144, 2048):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
                assert a1 == b0 and a0 <= a1
            assert all(a0 % 64 == 0 for a0, _ in spans if a0 < n)
    assert shard_range(262144, 3, 8) == (98304, 131072)
    assert packed_len(8192) == 73809
