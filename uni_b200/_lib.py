"""ctypes binding of libunimat.so (the C ABI in include/unimat.h).

This is the Python stand-in for the Scala 3 Panama binding: same symbols,
same struct layout.  There is NO CPU fallback -- if the shared library is missing the import
fails, and without a CUDA device every compute call raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# UNIMAT_LIB: development override (A/B timing of kernel variants built side by side); the product path is the in-tree library
LIB_PATH = os.environ.get("UNIMAT_LIB") or os.path.join(_HERE, "libunimat.so")

UM_OK, UM_ERR_ARG, UM_ERR_SINGULAR, UM_ERR_EMPTY, UM_ERR_CUDA, UM_ERR_NCCL, UM_ERR_OOM = range(7)
F64, F32, BOOL, I32 = 0, 1, 2, 3
ADD, SUB, MUL, DIV, MAXIMUM, MINIMUM, POW = range(7)
NEG, ABS, EXP, SQRT, LOG, SIGMOID, RELU, TANH, SQUARE, RELU_GENERIC = range(10)
GT, LT, GTE, LTE, EQ, NE = range(6)
ISNAN, ISINF, ISFINITE = range(3)
AND, OR, NOT, XOR = range(4)
SUM, MEAN, VAR, STD, MIN, MAX = range(6)
CUMSUM, CUMMAX, CUMMIN = range(3)
TPRF_CLOSED, TPRF_ITER, TPRF_IS_FULL, TPRF_WINDOW = range(4)


class UmView(C.Structure):
    _fields_ = [
        ("data", C.c_void_p),
        ("offset", C.c_int64),
        ("rows", C.c_int64),
        ("cols", C.c_int64),
        ("rs", C.c_int64),
        ("cs", C.c_int64),
        ("dtype", C.c_int32),
        ("transposed", C.c_int32),
    ]


class UmTprfOut(C.Structure):
    _fields_ = [
        ("forecasts", C.c_void_p),
        ("residuals", C.c_void_p),
        ("alpha", C.c_void_p),
        ("phi", C.c_void_p),
        ("phi_hat", C.c_void_p),
        ("sigma", C.c_void_p),
        ("beta3", C.c_void_p),
        ("xstd", C.c_void_p),
        ("r_squared", C.c_double),
        ("singular", C.c_int32),
    ]


class UnimatError(RuntimeError):
    """CUDA / NCCL / OOM failure inside libunimat (no reference analogue)."""


class EmptyMatrixError(ArithmeticError):
    """UnsupportedOperationException in the reference (Mat.scala:2216)."""


class SingularMatrixError(ArithmeticError):
    """ArithmeticException("Matrix is singular or nearly singular"), Mat.scala:990."""


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build it with `make -C uni_b200/csrc` (or __graft_entry__.build()). "
        "uni_b200 has no CPU fallback."
    )
lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)

VP = C.POINTER(UmView)
_i32, _i64, _f64, _vp, _sz = C.c_int32, C.c_int64, C.c_double, C.c_void_p, C.c_size_t

# every symbol include/unimat.h declares: (name, restype, argtypes)
SIGNATURES = {
    "um_version": (_i32, []),
    "um_last_error": (C.c_char_p, []),
    "um_device_count": (_i32, [C.POINTER(_i32)]),
    "um_init": (_i32, [_i32]),
    "um_shutdown": (_i32, []),
    "um_sync": (_i32, []),
    "um_device_props": (_i32, [C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i64)]),
    "um_malloc": (_i32, [_sz, C.POINTER(_vp)]),
    "um_free": (_i32, [_vp]),
    "um_malloc_host": (_i32, [_sz, C.POINTER(_vp)]),
    "um_free_host": (_i32, [_vp]),
    "um_h2d": (_i32, [_vp, _vp, _sz]),
    "um_d2h": (_i32, [_vp, _vp, _sz]),
    "um_d2d": (_i32, [_vp, _vp, _sz]),
    "um_memset": (_i32, [_vp, _i32, _sz]),
    "um_h2d_2d": (_i32, [_vp, _sz, _vp, _sz, _sz, _sz]),
    "um_launch_count": (_i32, [C.POINTER(_i64)]),
    "um_timer_start": (_i32, []),
    "um_timer_stop": (_i32, [C.POINTER(C.c_float)]),
    "um_prof_enable": (_i32, [_i32]),
    "um_prof_reset": (_i32, []),
    "um_prof_read": (_i32, [_i32, C.POINTER(_f64), C.POINTER(_i64)]),
    "um_view_init": (_i32, [VP, _vp, _i64, _i64, _i32]),
    "um_view_is_weird": (_i32, [VP, C.POINTER(_i32)]),
    "um_view_is_contiguous": (_i32, [VP, C.POINTER(_i32)]),
    "um_view_transpose": (_i32, [VP, VP]),
    "um_view_slice": (_i32, [VP, _i64, _i64, _i64, _i64, VP, C.POINTER(_i32)]),
    "um_view_broadcast": (_i32, [VP, _i64, _i64, VP]),
    "um_broadcast_shape": (_i32, [VP, VP, C.POINTER(_i64), C.POINTER(_i64)]),
    "um_copy": (_i32, [VP, VP]),
    "um_fill": (_i32, [VP, _f64]),
    "um_eye": (_i32, [VP]),
    "um_randn": (_i32, [VP, C.c_uint64]),
    "um_rand": (_i32, [VP, C.c_uint64]),
    "um_rng_seed": (_i32, [_vp, _i64]),
    "um_rng_advance": (_i32, [_vp, C.c_uint64]),
    "um_rng_next_u64": (_i32, [_vp, C.POINTER(C.c_uint64)]),
    "um_rng_rand": (_i32, [_vp, VP]),
    "um_rng_uniform": (_i32, [_vp, _f64, _f64, VP]),
    "um_rng_randn": (_i32, [_vp, VP]),
    "um_rng_normal": (_i32, [_vp, _f64, _f64, VP]),
    "um_rng_randint": (_i32, [_vp, _i32, _i32, VP]),
    "um_csv_open": (_i32, [C.c_char_p, _i32, C.POINTER(C.c_uint64), C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i32)]),
    "um_csv_header": (_i32, [C.c_uint64, _i64, C.c_char_p, _i64]),
    "um_csv_read_host": (_i32, [C.c_uint64, _vp, _i64]),
    "um_csv_upload": (_i32, [C.c_uint64, VP]),
    "um_csv_close": (_i32, [C.c_uint64]),
    "um_csv_save": (_i32, [VP, C.c_char_p, C.c_char_p, C.c_char_p]),
    "um_csv_format_f64": (_i32, [_f64, C.c_char_p, C.c_char_p, _i64]),
    "um_csv_format_f32": (_i32, [C.c_float, C.c_char_p, C.c_char_p, _i64]),
    "um_csv_parse_cell": (_i32, [C.c_char_p, C.POINTER(_f64)]),
    "um_get": (_i32, [VP, _i64, _i64, C.POINTER(_f64)]),
    "um_set": (_i32, [VP, _i64, _i64, _f64]),
    "um_binary": (_i32, [_i32, VP, VP, VP]),
    "um_binary_scalar": (_i32, [_i32, VP, _f64, _i32, VP]),
    "um_unary": (_i32, [_i32, VP, VP]),
    "um_softmax": (_i32, [VP, _i32, VP]),
    "um_softmax_set_axis0_path": (_i32, [_i32]),
    "um_compare_scalar": (_i32, [_i32, VP, _f64, VP]),
    "um_compare": (_i32, [_i32, VP, VP, VP]),
    "um_classify": (_i32, [_i32, VP, VP]),
    "um_mask_logic": (_i32, [_i32, VP, VP, VP]),
    "um_mask_reduce": (_i32, [_i32, VP, _i32, VP, C.POINTER(_i32)]),
    "um_cumulative": (_i32, [_i32, VP, _i32, VP]),
    "um_where": (_i32, [VP, VP, VP, VP]),
    "um_where_scalar": (_i32, [VP, _f64, _f64, VP]),
    "um_mask_count": (_i32, [VP, C.POINTER(_i64)]),
    "um_mask_select": (_i32, [VP, VP, VP]),
    "um_reduce_all": (_i32, [_i32, VP, C.POINTER(_f64)]),
    "um_reduce_axis": (_i32, [_i32, VP, _i32, VP]),
    "um_arg_all": (_i32, [_i32, VP, C.POINTER(_i64), C.POINTER(_i64)]),
    "um_idx_axis": (_i32, [_i32, VP, _i32, VP]),
    "um_matmul": (_i32, [VP, VP, VP]),
    "um_matmul_set_f32_path": (_i32, [_i32]),
    "um_matmul_set_f64_tile": (_i32, [_i32]),
    "um_inverse": (_i32, [VP, VP]),
    "um_solve": (_i32, [VP, VP, VP]),
    "um_inverse_batched": (_i32, [VP, _i64, VP, _i64, _i64]),
    "um_tprf_fit": (_i32, [_i32, VP, VP, VP, _i64, C.POINTER(UmTprfOut)]),
    "um_tprf_fit_host": (_i32, [_i32, _vp, _vp, _i64, _vp, _i64, _i64, _i64, _i64, _vp, _vp, C.POINTER(_f64)]),
    "um_tprf_fit_batched": (_i32, [_i32, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _vp, _vp, _vp]),
    "um_tprf_oos_forecast": (_i32, [VP, VP, VP, VP, C.POINTER(_f64)]),
    "um_tprf_oos_windows": (_i32, [VP, VP, VP, _i32, _i32, C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32), _i64, _i32, _i32,
                            C.POINTER(_f64), C.POINTER(_f64)]),
    "um_tprf_work": (_i32, [_i64, _i64, _i64, C.POINTER(_f64), C.POINTER(_f64)]),
    "um_tprf_set_path": (_i32, [_i32]),
    "um_tprf_debug_stamps": (_i32, [_vp]),
    "um_tprf_fused_info": (_i32, [_i64, C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32)]),
    "um_comm_unique_id": (_i32, [C.POINTER(C.c_uint8)]),
    "um_comm_init": (_i32, [_i32, _i32, C.POINTER(C.c_uint8)]),
    "um_comm_destroy": (_i32, []),
    "um_comm_info": (_i32, [C.POINTER(_i32), C.POINTER(_i32)]),
    "um_comm_transport": (_i32, [C.POINTER(_i32)]),
    "um_comm_allreduce_sum_f64": (_i32, [_vp, _i64]),
    "um_comm_barrier": (_i32, []),
}
for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)  # AttributeError here = header and library out of sync
    _fn.restype = _res
    _fn.argtypes = _args


def last_error() -> str:
    return (lib.um_last_error() or b"").decode("utf-8", "replace")


def check(status: int) -> None:
    """Status -> the exception class the reference throws for the same condition."""
    if status == UM_OK:
        return
    msg = last_error()
    if status == UM_ERR_ARG:
        raise ValueError(msg)  # IllegalArgumentException (`require`)
    if status == UM_ERR_SINGULAR:
        raise SingularMatrixError(msg)
    if status == UM_ERR_EMPTY:
        raise EmptyMatrixError(msg)
    if status == UM_ERR_OOM:
        raise MemoryError(msg)
    raise UnimatError(f"libunimat status {status}: {msg}")


def device_count() -> int:
    n = _i32(0)
    check(lib.um_device_count(C.byref(n)))
    return n.value


def launch_count() -> int:
    n = _i64(0)
    check(lib.um_launch_count(C.byref(n)))
    return n.value
