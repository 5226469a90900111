"""CPU restatement of the reference's matrix file I/O -- TEST INFRASTRUCTURE ONLY (the product path is csrc/csv.cu).

Follows /root/reference/src/main/scala/uni/io/FileOps.scala:117-121 (namedHeaders), :139-188 (loadSmart),
src/main/scala/uni/data/Big.scala:59-64 (`big(str)`), src/main/scala/uni/io/FastCsv.scala:21 (isContentRow),
:35-41 (rectangular), src/main/scala/uni/data/Mat.scala:90-116 (saveCSV; cells via String.valueOf, i.e.
java.lang.Double.toString -- JDK 19+ shortest-digits form).  Row splitting is Python's csv module (RFC 4180 quotes),
the delimiter is given or the most frequent of , TAB ; | on the first content line (same scope as csv.cu; the
reference's statistical detector io/Delimiter.scala is not restated).

Pinned by tests/test_oracle_csv.py: the reference's own CSV fixtures (t3prf-validation/*.csv written by NumPy's
savetxt, tprf3-parity inputs written with %.17e) must load to exactly the arrays NumPy reads, and Java's documented
Double.toString examples must format as documented."""
import csv
import io
import math
import re
from decimal import Decimal

import numpy as np

_BIGDEC = re.compile(r"^[+-]?(\d+\.?\d*|\.\d+)([eE][+-]?\d{1,10})?$")


def big_to_double(s: str) -> float:
    """big(str).toDouble (Big.scala:59-64)."""
    t = s.strip(" \t\r\n\x0b\x0c" + "".join(chr(c) for c in range(0, 33)))
    t = t.replace(",", "").replace("$", "")
    percent = t.endswith("%")
    if percent:
        t = t[:-1]
    if not _BIGDEC.match(t):
        return math.nan
    d = Decimal(t)
    if percent:
        d = d / Decimal(100)  # BigDecimal / 100 (exact: a decimal shift)
    v = float(d)              # correctly rounded, like BigDecimal.doubleValue
    return 0.0 if v == 0 else v


def _detect(text: str) -> str:
    for line in text.split("\n"):
        if line.strip():
            best, bestn = ",", 0
            for cand in ",\t;|":
                cnt, q = 0, False
                for ch in line:
                    if ch == '"':
                        q = not q
                    elif not q and ch == cand:
                        cnt += 1
                if cnt > bestn:
                    best, bestn = cand, cnt
            return best
    return ","


def load_smart(path, delimiter=None):
    """loadSmart(p, _.toDouble) -> (headers, float64 2-D array) (FileOps.scala:139-188)."""
    raw = open(path, "rb").read()
    if raw[:3] == b"\xef\xbb\xbf":
        raw = raw[3:]
    text = raw.decode("utf-8", "replace")
    delim = delimiter or _detect(text)
    rows = [r for r in csv.reader(io.StringIO(text, newline=""), delimiter=delim, quotechar='"', doublequote=True)
            if any(f.strip() for f in r)]                                  # FastCsv.isContentRow
    if not rows:
        return [], np.empty((0, 0))
    width = max(len(r) for r in rows)
    rows = [r + [""] * (width - len(r)) for r in rows]                    # FastCsv.rectangular
    header = 0
    if len(rows) > 1:
        row0_text = all(math.isnan(big_to_double(s)) for s in rows[0])
        row1_data = any(not math.isnan(big_to_double(s)) for s in rows[1])
        if row0_text and row1_data:
            header = 1
    headers = [(h.strip() or f"col{i + 1}") for i, h in enumerate(rows[0])] if header else []
    data = rows[header:]
    out = np.empty((len(data), width))
    for r, row in enumerate(data):
        for c in range(width):
            out[r, c] = big_to_double(row[c])
    return headers, out


def java_double_to_string(v, f32=False) -> str:
    """java.lang.Double.toString / Float.toString (JDK 19+)."""
    v = float(np.float32(v)) if f32 else float(v)
    if v != v:
        return "NaN"
    if math.isinf(v):
        return "Infinity" if v > 0 else "-Infinity"
    if v == 0:
        return "-0.0" if math.copysign(1.0, v) < 0 else "0.0"
    sci = np.format_float_scientific(np.float32(v) if f32 else np.float64(v), unique=True, trim="-", exp_digits=1)
    mant, exp = sci.split("e")
    sign = "-" if mant.startswith("-") else ""
    digits = mant.lstrip("-").replace(".", "")
    e = int(exp)
    if len(digits) == 1:
        # at least two digits, and the closest two-digit decimal that reads back (Double.MIN_VALUE -> "4.9E-324")
        two = "%.1e" % abs(v)
        back = float(np.float32(two)) if f32 else float(two)
        if back == abs(v):
            m2, e2 = two.split("e")
            digits, e = m2.replace(".", "").rstrip("0") or "0", int(e2)
            if len(digits) < 1:
                digits = "0"
    if -3 <= e < 7:
        if e >= 0:
            ip = (digits + "0" * (e + 1))[: e + 1]
            fp = digits[e + 1:] or "0"
            return f"{sign}{ip}.{fp}"
        return f"{sign}0.{'0' * (-e - 1)}{digits}"
    return f"{sign}{digits[0]}.{digits[1:] or '0'}E{e}"


def format_cell(v, nan_as="NaN", kind="f64") -> str:
    """one cell of saveCSV (Mat.scala:95-110)."""
    if kind in ("f64", "f32"):
        if v != v:
            return nan_as
        if math.isinf(v):
            return "Inf" if v > 0 else "-Inf"
        return java_double_to_string(v, f32=(kind == "f32"))
    if kind == "bool":
        return "true" if v else "false"
    return str(int(v))


def save_csv_text(a, sep=",", nan_as="NaN") -> str:
    a = np.asarray(a)
    kind = {"float64": "f64", "float32": "f32", "bool": "bool"}.get(a.dtype.name, "i32")
    return "".join(sep.join(format_cell(v, nan_as, kind) for v in row) + "\n" for row in a)
