#!/usr/bin/env python3
"""hottest CUDA source lines of an .ncu-rep: stall samples and executed instructions per line"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25; which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
blocks = []; cur = None
for r in rows:
    if r and r[0] == "Line No":
        cur = {"hdr": r, "lines": [], "file": lastfile}; blocks.append(cur); continue
    if r and r[0] == "File Path":
        lastfile = r[1]; continue
    if cur is not None and r and r[0].isdigit():
        cur["lines"].append(r)
kernels = {}
for b in blocks:
    kernels.setdefault(id(b), b)
# blocks repeat per (kernel instance, file); print per block
seen = 0
for b in blocks:
    h = b["hdr"]
    cs = h.index("Warp Stall Sampling (All Samples)"); ci = h.index("Instructions Executed"); ct = h.index("Avg. Threads Executed") if "Avg. Threads Executed" in h else None
    def f(x):
        try: return float(x.replace(",", ""))
        except ValueError: return 0.0
    tot_s = sum(f(r[cs]) for r in b["lines"]) or 1; tot_i = sum(f(r[ci]) for r in b["lines"]) or 1
    if tot_s < 50: continue
    if seen != which: seen += 1; continue
    seen += 1
    print("### %s  samples=%d inst=%.3g" % (b["file"].split("/")[-1], tot_s, tot_i))
    for r in sorted(b["lines"], key=lambda r: -(f(r[ci]) if len(sys.argv) > 4 else f(r[cs])))[:top]:
        print("  %5.1f%% smp  %5.1f%% inst  thr=%-5s L%-4s %s" % (100 * f(r[cs]) / tot_s, 100 * f(r[ci]) / tot_i, r[ct] if ct else "", r[0], r[1].strip()[:130]))
    break
