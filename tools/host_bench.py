#!/usr/bin/env python3
"""Host orchestrator alone, timed on the CPU (test infrastructure, no GPU).

  PB200_HOST_THREADS=4 python tools/host_bench.py [L=1000000] [queries=8] [runs=5]

The search goes through oracle/ref_backend.cpp's record/replay table: the first alignment sends every window to csgmum and
records the answer, the following ones get their candidates back at memcpy speed - the situation on a B200, where the search
kernels take a few ms.  What remains is the product's host code (anchors accept, speculation ‖ exact replay, LCB chaining), with
its own phase timers.  Not a benchmark of the product (bench.py is): a tool to see what a host-side change does before a GPU is
at hand."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from parsnp_b200 import api, synth
from oracle import hosttest
if os.environ.get("PB200_HOSTTEST_LIB"):         # A/B of two builds of the host code
    hosttest.LIB = os.environ["PB200_HOSTTEST_LIB"]

L = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 8
runs = int(sys.argv[3]) if len(sys.argv) > 3 else 5
g = synth.g_indep(L, nq, 0.01, 1)
lib = hosttest.load()
lib.pbtest_replay_misses.restype = __import__("ctypes").c_long
t0 = time.time()
res = hosttest.align(g, api.make_params(), backend=2)
print("warm-up (records %d windows through csgmum): %.1f s, %d MUMs" % (lib.pbtest_replay_misses(1), time.time() - t0, len(res["mum_length"])), flush=True)
keys = ["t_total", "t_anchor_search", "t_anchor_host", "t_spec_search", "t_spec_host", "t_replay", "t_replay_search", "t_replay_wait", "t_lcb"]
rows = []
for i in range(runs):
    t0 = time.time()
    res = hosttest.align(g, api.make_params(), backend=2)
    wall = time.time() - t0
    st = res["stats"]
    rows.append([wall] + [st[k] for k in keys])
    print("run %d: wall %.1f ms, misses %d | " % (i, wall * 1e3, lib.pbtest_replay_misses(1)) + "  ".join("%s %.1f" % (k[2:], st[k] * 1e3) for k in keys), flush=True)
walls = sorted(r[0] for r in rows)
print("wall min %.1f / median %.1f ms, replay_wait median %.1f ms (threads %s, slices %d)" % (
    walls[0] * 1e3, walls[len(walls) // 2] * 1e3, sorted(r[1 + keys.index("t_replay_wait")] for r in rows)[len(rows) // 2] * 1e3,
    int(res["stats"]["host_threads"]), int(res["stats"]["spec_slices"])))
