"""Top stall sites of an ncu `--page source --csv` dump:  python tools/ncu_hot.py src.csv [n]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
top = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]] or 0))[:n]
print(f"total samples {tot}")
for i in sorted(top):
    r = data[i]
    s = int(r[ix["# Samples"]] or 0)
    st = sorted(((int(r[ix[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:3]
    print(f"{i:5d} {s:6d} {100.0 * s / tot:5.1f}%  {r[ix['Source']][:70]:70s} " + " ".join(f"{c}:{v}" for v, c in st if v))
