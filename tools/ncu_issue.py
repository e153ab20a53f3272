#!/usr/bin/env python3
"""Issue-slot utilisation of a kernel group from `ncu --set full` captures, merged into profiles/ncu_traffic.json (bench.py
quotes it beside the byte roofline of the recursion kernels, which are instruction / shared-memory bound by construction):

  python tools/ncu_issue.py profiles/ncu_traffic.json "<bench.py roofline kernel label>" a.ncu-rep [b.ncu-rep ...]

value = smsp__issue_active.avg.pct_of_peak_sustained_active of the captured launches, weighted by their duration."""
import csv
import io
import json
import subprocess
import sys

out, label, reps = sys.argv[1], sys.argv[2], sys.argv[3:]
tot_t = tot = 0.0
parts = []
for rep in reps:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    r = list(csv.reader(io.StringIO(raw)))
    hdr = r[0]
    ki, ti, ii = hdr.index("Kernel Name"), hdr.index("gpu__time_duration.sum"), hdr.index("smsp__issue_active.avg.pct_of_peak_sustained_active")
    for row in r[2:]:
        t, v = float(row[ti].replace(",", "")), float(row[ii].replace(",", ""))
        tot_t += t
        tot += t * v
        parts.append({"kernel": row[ki].split("(")[0][-60:], "duration": t, "issue_active_pct": v})
d = json.load(open(out))
d.setdefault("issue_active_pct", {})[label] = {"value": tot / tot_t if tot_t else None, "launches": parts,
                                                "source": "ncu --set full (%s), duration-weighted" % ", ".join(x.split("/")[-1] for x in reps)}
json.dump(d, open(out, "w"), indent=1, sort_keys=True)
print(label, d["issue_active_pct"][label]["value"])
