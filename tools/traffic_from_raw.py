"""profiles/r02_traffic.json from a `*_ncu_full_raw.csv` written by tools/ncu_export.py (rows = metrics, columns = launches):
per-launch DRAM bytes and pipe utilisation of the named kernels; bench.py reads `dram_bytes` and the tensor-pipe figures
from it for `roofline.traffic`.
usage: python tools/traffic_from_raw.py profiles/<name>_ncu_full_raw.csv "<how the capture was taken>" label=launch_index ..."""
import csv
import json
import os
import sys

UNIT_SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3}
raw, how = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(raw)))
names = rows[1][2:]
out = {}
for spec in sys.argv[3:]:
    label, idx = spec.rsplit("=", 1)
    idx = int(idx)
    e = {"source": f"profiles/{os.path.basename(raw)} ({how}), launch {idx}: {names[idx]}"}
    for r in rows[2:]:
        if len(r) <= 2 + idx or r[2 + idx] in ("", "n/a"):
            continue
        try:
            e[r[0]] = float(r[2 + idx].replace(",", "")) * UNIT_SCALE.get(r[1], 1)
        except ValueError:
            pass
    if "dram__bytes_read.sum" in e:
        e["dram_bytes"] = e["dram__bytes_read.sum"] + e["dram__bytes_write.sum"]
    out[label] = e
dst = os.path.join(os.path.dirname(os.path.abspath(raw)), "r02_traffic.json")
json.dump(out, open(dst, "w"), indent=1)
for k, v in out.items():
    print(k, "|", v["source"][-90:], "|", v.get("gpu__time_duration.sum"), "us", v.get("dram_bytes"))
