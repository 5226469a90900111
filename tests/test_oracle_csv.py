"""CPU: the CSV oracle (oracle/csv_oracle.py) against the reference's own CSV fixtures and Java's documented
Double.toString behaviour, and the host half of the C ABI (um_csv_open / um_csv_read_host / um_csv_header /
um_csv_format_* / um_csv_parse_cell -- no device involved) against the oracle."""
import glob
import math
import os

import numpy as np
import pytest

from oracle import csv_oracle as co

# Java SE documentation / well-known values of Double.toString and Float.toString
JAVA_DOUBLE = [(1.0, "1.0"), (100.0, "100.0"), (0.001, "0.001"), (1.0e-4, "1.0E-4"), (1234567.0, "1234567.0"), (1.0e7, "1.0E7"),
               (12345678.9, "1.23456789E7"), (0.1, "0.1"), (1.0 / 3.0, "0.3333333333333333"), (4.9e-324, "4.9E-324"),
               (1.7976931348623157e308, "1.7976931348623157E308"), (-2.5e-7, "-2.5E-7"), (2.0e-3, "0.002"), (1.0e23, "1.0E23"),
               (9.999e-4, "9.999E-4"), (-0.0, "-0.0"), (0.0, "0.0"), (123456.789, "123456.789"), (2.0 ** 63, "9.223372036854776E18")]
JAVA_FLOAT = [(0.1, "0.1"), (16777216.0, "1.6777216E7"), (3.4028235e38, "3.4028235E38"), (1.0e-3, "0.001"), (1.4e-45, "1.4E-45"), (0.33333334, "0.33333334")]


def test_oracle_java_number_layout():
    for v, want in JAVA_DOUBLE:
        assert co.java_double_to_string(v) == want
    for v, want in JAVA_FLOAT:
        assert co.java_double_to_string(v, f32=True) == want
    assert co.format_cell(math.nan, "NA") == "NA" and co.format_cell(math.inf) == "Inf" and co.format_cell(-math.inf) == "-Inf"


CELLS = ["1", "-1.5", " 2.5 ", "1,234.5", "$3.50", "12%", "12.5%", "1e3", "1E-2", ".5", "5.", "+7", "abc", "NaN", "Inf", "", "1e", "--1",
         "1.2.3", "0x10", "1e400", "1e-400", "-0", "1 2", "-1,000,000", "$-5", "1e+05", "00012", "1.7976931348623157e308",
         "2.2250738585072011e-308", "0.1e-1%", "%", "e5", ".", "+", "9007199254740993", "0.30000000000000004"]


def test_oracle_cell_grammar_known_answers():
    want = {"1": 1.0, "1,234.5": 1234.5, "$3.50": 3.5, "12%": 0.12, ".5": 0.5, "5.": 5.0, "+7": 7.0, "1e400": math.inf, "1e-400": 0.0,
            "-0": 0.0, "9007199254740993": 9007199254740992.0, "0.1e-1%": 1e-4}
    for s, v in want.items():
        assert co.big_to_double(s) == v, s
    for s in ["abc", "NaN", "Inf", "", "1e", "--1", "1.2.3", "0x10", "1 2", "%", "e5", ".", "+"]:
        assert math.isnan(co.big_to_double(s)), s


def test_oracle_loads_the_reference_fixtures_exactly(golden_dir):
    files = sorted(glob.glob(os.path.join(golden_dir, "t3prf-validation", "*.csv")) + glob.glob(os.path.join(golden_dir, "tprf3-parity", "*.csv")))
    assert len(files) >= 15
    for f in files:
        headers, a = co.load_smart(f)
        want = np.loadtxt(f, delimiter=",", ndmin=2)
        assert headers == [] and a.shape == want.shape and np.array_equal(a.view(np.uint64), want.view(np.uint64)), f


@pytest.fixture()
def nasty(tmp_path):
    p = tmp_path / "nasty.csv"
    p.write_bytes(b"\xef\xbb\xbfdate, price ,,\"vol, k\"\r\n"
                  b"1,\"1,234.50\",$7,12%\r\n"
                  b"\r\n"
                  b"   ,  \n"
                  b"2,abc,3e2\n"
                  b"3,\"say \"\"hi\"\"\",-0.5,4,99\n"
                  b"4,1e400,.25,5.")
    return str(p)


def test_oracle_header_padding_and_blank_rows(nasty):
    headers, a = co.load_smart(nasty)
    assert headers == ["date", "price", "col3", "vol, k", "col5"]
    want = np.array([[1, 1234.5, 7, 0.12, np.nan], [2, np.nan, 300, np.nan, np.nan], [3, np.nan, -0.5, 4, 99], [4, np.inf, 0.25, 5, np.nan]])
    assert a.shape == (4, 5) and np.array_equal(a, want, equal_nan=True)


# ---- host half of the C ABI ----
def test_abi_cell_format_and_parse_match_oracle():
    from uni_b200 import io
    for v, want in JAVA_DOUBLE:
        assert io.format_cell(v) == want
    for v, want in JAVA_FLOAT:
        assert io.format_cell(v, f32=True) == want
    rng = np.random.default_rng(5)
    vals = np.concatenate([rng.standard_normal(2000), rng.standard_normal(2000) * 10.0 ** rng.integers(-300, 300, 2000), [math.nan, math.inf, -math.inf]])
    for v in vals:
        assert io.format_cell(float(v), "n/a") == co.format_cell(float(v), "n/a"), v
    for s in CELLS:
        a, b = io.parse_cell(s), co.big_to_double(s)
        assert (math.isnan(a) and math.isnan(b)) or (a == b and math.copysign(1, a) == math.copysign(1, b)), s
    for v in vals[:4000]:  # every formatted finite cell parses back to the same double
        assert io.parse_cell(io.format_cell(float(v))) == float(v)


def test_abi_parses_files_like_the_oracle(golden_dir, nasty, tmp_path):
    from uni_b200 import io
    files = sorted(glob.glob(os.path.join(golden_dir, "t3prf-validation", "*.csv")) + glob.glob(os.path.join(golden_dir, "tprf3-parity", "*.csv")))
    for f in files + [nasty]:
        h, a = io.parse_host(f)
        ho, ao = co.load_smart(f)
        assert h == ho and a.shape == ao.shape and np.array_equal(a.view(np.uint64), ao.view(np.uint64)), f
    # tab / semicolon detection, a single column, an empty file, no trailing newline, a large table (threaded path)
    (tmp_path / "t.tsv").write_text("a\tb\n1\t2\n3\t4")
    (tmp_path / "s.csv").write_text("1;2;3\n4;5;6\n")
    (tmp_path / "one.csv").write_text("x\n1\n2\n\n3\n")
    (tmp_path / "empty.csv").write_text("")
    big = np.random.default_rng(9).standard_normal((3000, 7))
    np.savetxt(tmp_path / "big.csv", big, delimiter=",", fmt="%.17e")
    for name in ["t.tsv", "s.csv", "one.csv", "empty.csv", "big.csv"]:
        h, a = io.parse_host(str(tmp_path / name))
        ho, ao = co.load_smart(str(tmp_path / name))
        assert h == ho and a.shape == ao.shape and np.array_equal(a, ao, equal_nan=True), name
    assert np.array_equal(io.parse_host(str(tmp_path / "big.csv"))[1], big)
    with pytest.raises(ValueError, match="cannot read"):
        io.parse_host(str(tmp_path / "missing.csv"))


def test_abi_readers_agree_with_oracle_on_random_ragged_files(tmp_path):
    """Differential test of both readers of csv.cu -- the parallel quote-free one and the sequential quoted one (forced by
    appending one quoted field) -- against the oracle: ragged rows, blank and delimiter-only rows, \\n / \\r\\n / \\r line
    ends, a BOM, all four delimiters, with and without a header row."""
    import random
    from uni_b200 import io
    random.seed(3)
    cells = ["1", "-2.5", "3e2", "abc", "", "  7 ", "1,5", "$4", "12%", "NaN", "1e400", "-0", ".5", "5.", "+3", "x y", "0x1", "1e-400", "--1", "+-2"]

    def gen(delim, eol, bom, header, nrows):
        rows = []
        if header:
            rows.append(delim.join(random.choice(["id", "name", "", " px ", "v"]) for _ in range(random.randint(1, 5))))
        for _ in range(nrows):
            k = random.randint(0, 6)
            if k == 0:
                rows.append(random.choice(["", "   ", "\t" if delim != "\t" else " ", delim, delim + " " + delim]))
            else:
                cs = [random.choice(cells) for _ in range(k)]
                if delim == ",":
                    cs = [c for c in cs if "," not in c] or ["1"]
                rows.append(delim.join(cs))
        txt = eol.join(rows) + (eol if random.random() < 0.5 else "")
        return (b"\xef\xbb\xbf" if bom else b"") + txt.encode()

    p = tmp_path / "r.csv"
    for _ in range(400):
        delim = random.choice([",", ";", "\t", "|"])
        data = gen(delim, random.choice(["\n", "\r\n", "\r"]), random.random() < 0.2, random.random() < 0.5, random.randint(0, 12))
        for blob in (data, data + (b'\n"9"' if data and not data.endswith((b"\n", b"\r")) else b'"9"')):
            p.write_bytes(blob)
            h, a = io.parse_host(str(p), delim)
            ho, ao = co.load_smart(str(p), delim)
            assert h == ho and a.shape == ao.shape and np.array_equal(a, ao, equal_nan=True), blob
