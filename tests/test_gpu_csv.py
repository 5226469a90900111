"""GPU: matrix file I/O through the C ABI (csrc/csv.cu): loadMatD / loadMatF / loadSmartD upload what the oracle
parses, saveCSV writes the text the oracle's restatement of Mat.saveCSV writes (bit-exact: text), and the
reference's own fixture files drive a 3PRF fit end to end (file -> device -> fit -> kp_yhat)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import csv_oracle as co


@pytest.fixture(scope="module")
def u():
    import uni_b200
    if uni_b200.device_count() == 0:
        pytest.fail("no CUDA device: the product path has no CPU fallback")
    uni_b200.init(0)
    return uni_b200


def test_load_fixture_files_and_fit(u, golden_dir):
    d = os.path.join(golden_dir, "t3prf-validation")
    X, Z, y = (u.loadMatD(os.path.join(d, f"ref_{n}.csv")) for n in "XZy")
    assert X.shape == (50, 6) and Z.shape == (50, 2) and y.shape == (50, 1)
    for m, n in ((X, "X"), (Z, "Z"), (y, "y")):
        assert np.array_equal(m.to_numpy(), co.load_smart(os.path.join(d, f"ref_{n}.csv"))[1])
    r = u.tprfClosedForm(y, X, Z)
    kp = np.loadtxt(os.path.join(d, "kp_yhat.csv"), delimiter=",").reshape(-1, 1)
    assert np.allclose(r.forecasts.to_numpy(), kp, rtol=1e-9, atol=1e-12)   # the north-star 3PRF tolerance


def test_load_headers_padding_f32(u, tmp_path):
    p = tmp_path / "h.csv"
    p.write_text("id,price,,\"vol, k\"\n1,\"1,234.50\",$7,12%\n\n2,abc,3e2\n3,x,-0.5,4,99\n")
    res = u.loadSmartD(str(p))
    ho, ao = co.load_smart(str(p))
    assert res.headers == ho == ["id", "price", "col3", "vol, k", "col5"]
    assert np.array_equal(res.mat.to_numpy(), ao, equal_nan=True)
    assert res.columnIndex["price"] == 1
    f = u.loadMatF(str(p))
    assert f.dtype == u._lib.F32 and np.array_equal(f.to_numpy(), ao.astype(np.float32), equal_nan=True)
    (tmp_path / "empty.csv").write_text("")
    assert u.loadMatD(str(tmp_path / "empty.csv")).shape == (0, 0)
    with pytest.raises(ValueError, match="cannot read"):
        u.loadMatD(str(tmp_path / "nope.csv"))


def test_save_matches_reference_text_and_round_trips(u, tmp_path):
    rng = np.random.default_rng(3)
    a = rng.standard_normal((700, 9)) * 10.0 ** rng.integers(-12, 12, (700, 9))
    a[0, 0], a[1, 1], a[2, 2], a[3, 3], a[4, 4] = np.nan, np.inf, -np.inf, 0.0, -0.0
    m = u.Mat.from_numpy(a)
    p = tmp_path / "out.csv"
    m.saveCSV(str(p))
    assert p.read_text() == co.save_csv_text(a)
    back = u.loadMatD(str(p)).to_numpy()
    fin = np.isfinite(a)
    assert np.array_equal(back[fin], a[fin]) and np.all(np.isnan(back[~fin]))   # "Inf" reads back as NaN, as in the reference
    # separators, nanAs, a transposed view (materialised first), f32 / bool / int matrices
    m.T.saveCSV(str(p), sep=";", nanAs="NA")
    assert p.read_text() == co.save_csv_text(a.T, ";", "NA")
    f = u.Mat.from_numpy(a.astype(np.float32)[:50])
    f.saveCSV(str(p))
    assert p.read_text() == co.save_csv_text(a.astype(np.float32)[:50])
    mask = m.gt(0.0)
    mask.saveCSV(str(p))
    assert p.read_text() == co.save_csv_text(a > 0.0)
    idx = m.idxmax(0)
    idx.saveCSV(str(p), sep="\t")
    assert p.read_text() == co.save_csv_text(idx.to_numpy(), "\t")


def test_generated_panel_to_file_and_back(u, tmp_path):
    u.MatD.setSeed(8)
    X = u.MatD.randn(300, 40)
    p = tmp_path / "panel.csv"
    X.saveCSV(str(p))
    assert np.array_equal(u.loadMatD(str(p)).to_numpy(), X.to_numpy())   # shortest-digit text reads back bit for bit
    want = np.random.default_rng(8).standard_normal((300, 40))
    assert np.allclose(X.to_numpy(), want, rtol=3e-16, atol=0)
