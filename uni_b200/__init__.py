"""uni_b200 -- B200-native device backend for uni's `MatD`/`MatF` + 3PRF hot path.

Everything computes in libunimat.so (hand-written sm_100a CUDA behind the C ABI of
include/unimat.h).  Importing this package loads that library and fails if it is missing.
"""
from . import _lib
from ._lib import (EmptyMatrixError, SingularMatrixError, UnimatError, device_count, launch_count)
from .mat import Mat, MatD, MatF, init, set_matd_matmul_tile, set_matf_matmul_path, set_softmax_axis0_path, sync
from . import tprf3
from . import rng
from . import io
from .io import MatResult, loadMatD, loadMatF, loadSmartD, saveCSV
from .rng import NumPyRNG
from .tprf3 import AvarEstimates, Pls3prfModel, Tprf3Result, estimate3prf, forecast3prf, pls1Fit, plsClosedForm, t3prf, tprfClosedForm

__all__ = ["Mat", "MatD", "MatF", "init", "sync", "set_matf_matmul_path", "set_matd_matmul_tile", "set_softmax_axis0_path", "tprf3", "Tprf3Result", "AvarEstimates", "estimate3prf", "forecast3prf",
           "t3prf", "tprfClosedForm", "plsClosedForm", "pls1Fit", "Pls3prfModel", "NumPyRNG", "rng", "io", "MatResult", "loadMatD", "loadMatF", "loadSmartD", "saveCSV", "device_count", "launch_count", "EmptyMatrixError",
           "SingularMatrixError", "UnimatError"]
