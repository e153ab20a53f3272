#!/usr/bin/env python3
"""DRAM traffic per kernel of one bench step, from an ncu launch list (run on the GPU box):

  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 600 --csv \
      --log-file gpurun_out/traffic.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline
  python tools/ncu_traffic.py gpurun_out/traffic.csv profiles/ncu_traffic.json

bench.py reads the JSON for roofline.traffic (bytes per launch of the dominant kernel).  Numbers taken under ncu are
cold-cache and serialised; they are traffic counts, never timings."""
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[h]
ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}
per = {}
for r in rows[h + 1:]:
    if len(r) <= vi:
        continue
    name = r[ki].split("(")[0].split("<")[0].split("::")[-1].strip()
    d = per.setdefault(r[0], {"kernel": name})
    try:
        v = float(r[vi].replace(",", "")) * scale.get(r[ui], 1.0)
    except ValueError:
        continue
    d[r[mi]] = v
out = {}
for d in per.values():
    k = out.setdefault(d["kernel"], {"launches": 0, "dram_bytes": 0.0, "us": 0.0})
    k["launches"] += 1
    k["dram_bytes"] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    k["us"] += d.get("gpu__time_duration.sum", 0.0)
for k in out.values():
    k["dram_bytes_per_launch"] = k["dram_bytes"] / max(1, k["launches"])
json.dump({"source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over one bench.py step (configs[1])", "kernels": out},
          open(sys.argv[2], "w"), indent=1, sort_keys=True)
for n, k in sorted(out.items(), key=lambda x: -x[1]["us"]):
    print("%-32s launches %4d  dram %10.2f MB  %9.1f us" % (n, k["launches"], k["dram_bytes"] / 1e6, k["us"]))
