"""profiles/r2_tprf_fused_traffic.json (what bench.py reports as roofline.traffic) from the raw page of an
`ncu --set full` capture of the two sweep kernels of a config-4 fit:  python tools/traffic_json.py raw.csv out.json"""
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}


def val(r, name):
    v = float(r[ix[name]].replace(",", ""))
    u = units[ix[name]].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1)


per, rd, wr = {}, 0.0, 0.0
for r in rows[2:]:
    name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("fused::", "").replace("gridk::", "")
    if name in per:
        continue   # one launch per kernel
    a, b = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
    per[name] = {"dram_bytes_read": a, "dram_bytes_write": b, "gpu_time_ms_under_ncu": round(val(r, "gpu__time_duration.sum"), 4)}
    rd += a
    wr += b
out = {"kernel": " + ".join(per) + " (split sweep: hardware clusters + virtual groups, concurrent in a normal run)",
       "T": 8192, "N_local": 262144, "L": 8, "dram_bytes_read": rd, "dram_bytes_write": wr, "per_kernel": per,
       "source": "ncu --set full --clock-control none -k regex:k_tprf_fused|k_tprf_grid -s 6 -c 2 python bench.py --steps 2 --warmup 1 --no-e2e --skip-ops; "
                 "ncu serialises the two kernels, in a normal run they overlap"}
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out)[:400])
