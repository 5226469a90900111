"""ctypes loader of oracle/libjvm_restated.so (the C++ restatement of the reference's pure-JVM Mat kernels,
oracle/jvm_restated.cpp).  TEST / BENCH INFRASTRUCTURE: imported by tests/ and by bench.py's cpu legs only."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "libjvm_restated.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_PATH):
            raise ImportError(f"{_PATH} missing: run `make -C oracle` (or __graft_entry__.build())")
        _lib = C.CDLL(_PATH)
        _lib.jr_sum.restype = C.c_double
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def matmul_pure(a, b, threads):
    a = np.ascontiguousarray(a, dtype=np.float64); b = np.ascontiguousarray(b, dtype=np.float64)
    out = np.empty((a.shape[0], b.shape[1]))
    lib().jr_matmul_pure(_p(a), _p(b), _p(out), C.c_long(a.shape[0]), C.c_long(a.shape[1]), C.c_long(b.shape[1]), C.c_int(threads))
    return out


def total(a, threads):
    return lib().jr_sum(_p(a), C.c_long(a.size), C.c_int(threads))


def col_sums(a, threads):
    out = np.empty(a.shape[1])
    lib().jr_col_sums(_p(a), C.c_long(a.shape[0]), C.c_long(a.shape[1]), _p(out), C.c_int(threads))
    return out


def row_sums(a, threads):
    out = np.empty(a.shape[0])
    lib().jr_row_sums(_p(a), C.c_long(a.shape[0]), C.c_long(a.shape[1]), _p(out), C.c_int(threads))
    return out


def sub_rowvec(a, v, out, threads):
    lib().jr_sub_rowvec(_p(a), _p(v), _p(out), C.c_long(a.shape[0]), C.c_long(a.shape[1]), C.c_int(threads))
    return out


def add_same(a, b, out, threads):
    lib().jr_add_same(_p(a), _p(b), _p(out), C.c_long(a.size), C.c_int(threads))
    return out


def sigmoid(a, out, threads):
    lib().jr_sigmoid(_p(a), _p(out), C.c_long(a.size), C.c_int(threads))
    return out


def softmax_rows(a, out):
    lib().jr_softmax_rows(_p(a), _p(out), C.c_long(a.shape[0]), C.c_long(a.shape[1]))
    return out
