#!/usr/bin/env python3
"""print the key raw metrics and the hottest source lines of an .ncu-rep (run here, no GPU needed)"""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
r = list(csv.reader(io.StringIO(raw)))
hdr, units = r[0], r[1]
want = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum']
for row in r[2:]:
    print("=== ", row[hdr.index('Kernel Name')][:100])
    for w in want:
        if w in hdr:
            print("  %-75s %s %s" % (w, row[hdr.index(w)], units[hdr.index(w)]))
    for i, h in enumerate(hdr):
        if 'issue_stalled' in h and h.endswith('per_issue_active.ratio') and 'not_issued' not in h:
            try:
                v = float(row[i])
            except ValueError:
                continue
            if v > 0.3:
                print("  stall %-60s %.2f" % (h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
# the source page prints one CSV block per kernel
blocks = src.split("\n\"Kernel Name\"")
for bi, b in enumerate(blocks[:2]):
    if bi > 0:
        b = "\"Kernel Name\"" + b
    rows = list(csv.reader(io.StringIO(b)))
    h = None
    for i, row in enumerate(rows):
        if row and row[0] in ('#', 'Source', 'Address') or (row and '# Warp Stall Sampling (All Samples)' in ''.join(row)):
            h = i
            break
    if h is None:
        continue
    head = rows[h]
    try:
        ci_src = head.index('Source')
    except ValueError:
        continue
    def col(name):
        for j, x in enumerate(head):
            if x.strip() == name:
                return j
        return None
    c_samp = col('# Warp Stall Sampling (All Samples)') or col('Warp Stall Sampling (All Samples)')
    c_inst = col('# Instructions Executed') or col('Instructions Executed')
    if c_samp is None:
        continue
    data = []
    for row in rows[h + 1:]:
        if len(row) <= max(c_samp, ci_src):
            continue
        try:
            sv = float(row[c_samp].replace(',', '') or 0)
        except ValueError:
            continue
        iv = row[c_inst] if c_inst is not None else ''
        data.append((sv, iv, row[ci_src].strip()[:140], row[0]))
    tot = sum(d[0] for d in data) or 1
    print("--- hottest source lines (stall samples) block", bi)
    for sv, iv, s, ln in sorted(data, key=lambda x: -x[0])[:top]:
        print("  %6.2f%%  inst=%-12s L%-5s %s" % (100 * sv / tot, iv, ln, s))
