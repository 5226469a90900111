"""Host mirror of the reference's matrix file I/O: `path.loadMatD / loadMatF / loadSmartD / readCsv`
(ext/PathExts.scala:583-594, io/FileOps.scala:139-188) and `m.saveCSV / writeCsv` (data/Mat.scala:90-119).
Parsing and formatting run on the host inside libunimat.so (csrc/csv.cu); the matrix lives on the device."""
from __future__ import annotations

import ctypes as C
import os
from typing import List, NamedTuple, Optional

import numpy as np

from . import _lib as L
from ._lib import check, lib
from .mat import Mat


class MatResult(NamedTuple):
    """`MatResult(headers, mat)` (io/FileOps.scala:95-100)."""
    headers: List[str]
    mat: Mat

    @property
    def columnIndex(self):
        return {h: i for i, h in enumerate(self.headers)}


class _Table:
    def __init__(self, path, delimiter: Optional[str]):
        self.h = C.c_uint64(0)
        self.rows, self.cols, self.hdr = L._i64(0), L._i64(0), L._i32(0)
        d = ord(delimiter) if delimiter else 0
        check(lib.um_csv_open(os.fsencode(path), d, C.byref(self.h), C.byref(self.rows), C.byref(self.cols), C.byref(self.hdr)))

    def headers(self) -> List[str]:
        if not self.hdr.value:
            return []
        out = []
        buf = C.create_string_buffer(4096)
        for c in range(self.cols.value):
            check(lib.um_csv_header(self.h, c, buf, len(buf)))
            out.append(buf.value.decode("utf-8", "replace"))
        return out

    def close(self):
        check(lib.um_csv_close(self.h))


def loadSmart(path, dtype: int = L.F64, delimiter: Optional[str] = None) -> MatResult:
    t = _Table(path, delimiter)
    try:
        m = Mat._alloc(t.rows.value, t.cols.value, dtype)
        check(lib.um_csv_upload(t.h, m._ref()))
        return MatResult(t.headers(), m)
    finally:
        t.close()


def loadSmartD(path, delimiter: Optional[str] = None) -> MatResult:
    return loadSmart(path, L.F64, delimiter)


def loadMatD(path, delimiter: Optional[str] = None) -> Mat:
    return loadSmart(path, L.F64, delimiter).mat


def loadMatF(path, delimiter: Optional[str] = None) -> Mat:
    return loadSmart(path, L.F32, delimiter).mat


readCsv = loadMatD
readCsvF = loadMatF


def parse_host(path, delimiter: Optional[str] = None):
    """The parsed table without a device: (headers, float64 ndarray).  Host arithmetic only (tests, tooling)."""
    t = _Table(path, delimiter)
    try:
        a = np.empty((t.rows.value, t.cols.value), dtype=np.float64)
        check(lib.um_csv_read_host(t.h, a.ctypes.data_as(C.c_void_p), a.size))
        return t.headers(), a
    finally:
        t.close()


def saveCSV(m: Mat, path, sep: str = ",", nanAs: str = "NaN") -> None:
    """`m.saveCSV(filePath, sep, nanAs)` (Mat.scala:90-116); views are materialised first."""
    src = m if m.isContiguous() or m.size == 0 else m.copy()
    check(lib.um_csv_save(src._ref(), os.fsencode(path), sep.encode(), nanAs.encode()))


writeCsv = saveCSV


def format_cell(v, nanAs: str = "NaN", f32: bool = False) -> str:
    buf = C.create_string_buffer(64)
    if f32:
        check(lib.um_csv_format_f32(C.c_float(v), nanAs.encode(), buf, len(buf)))
    else:
        check(lib.um_csv_format_f64(float(v), nanAs.encode(), buf, len(buf)))
    return buf.value.decode()


def parse_cell(text: str) -> float:
    out = L._f64(0)
    check(lib.um_csv_parse_cell(text.encode("utf-8"), C.byref(out)))
    return out.value
