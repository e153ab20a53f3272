#!/usr/bin/env python3
"""GPU campaign for the "final gaps" path (host/replay.cpp) with the REAL device flags: random sets with 8-12 base MUMs,
parallel replay forced, product (CUDA) == sequential loop over csgmum.  python tools/gpu_fuzz_final.py <first seed> <cases>
(tests/test_zz_gpu_fuzz.py::test_cuda_final_gaps_fuzz runs 16 of these cases in the GPU suite)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from parsnp_b200 import api, synth
from oracle import hosttest
from tests.refcmp import result_to_dump, diff_dumps

s0, nc = int(sys.argv[1]), int(sys.argv[2])
bad, tot = 0, dict(replay_gaps=0, replay_final_gaps=0, replay_final_mums=0, replay_foreign_reads=0, replay_foreign_writes=0, replay_restarts=0, replay_misses=0)
t0 = time.time()
for seed in range(s0, s0 + nc):
    rng = np.random.default_rng(seed)
    L = int(rng.choice([30000, 80000, 200000]))
    nq = int(rng.integers(1, 7))
    div = float(rng.choice([0.01, 0.03, 0.05]))
    g = (synth.g_indep if rng.random() < 0.6 else synth.g_pop)(L, nq, div, int(rng.integers(1, 10**6)))
    if rng.random() < 0.3:
        a = int(rng.integers(0, L - 400)); ln = int(rng.integers(30, 300)); a = min(a, L - 2 * ln)
        for x in g:
            x[a + ln:a + 2 * ln] = synth.revcomp(x[a:a + ln])
    if rng.random() < 0.2:
        k = int(rng.integers(1, len(g)))
        g[k] = synth.rearrange(g[k], rng, n_inv=1, inv_len=int(rng.integers(200, 2000)), dels=(), ins=())
    kw = dict(mums=str(rng.choice(["8", "10", "12", "1.1*(Log(S))", "0.7*(Log(S))"])), q=int(rng.choice([10, 30])))
    os.environ["PB200_REPLAY_MODE"] = "seq"
    want = result_to_dump(hosttest.align(g, api.make_params(**kw), backend=1))
    for rep in range(2):
        os.environ["PB200_REPLAY_MODE"] = "par"
        os.environ["PB200_REPLAY_TASK"] = str(int(rng.choice([1, 2, 5, 40])))
        os.environ["PB200_REPLAY_THREADS"] = str(int(rng.choice([2, 4, 8])))
        got = api.align(g, api.make_params(**kw))
        for k in tot:
            tot[k] += got["stats"].get(k, 0)
        d = diff_dumps(result_to_dump(got), want)
        if d:
            bad += 1
            print("MISMATCH seed", seed, "rep", rep, kw, d[:3], flush=True)
    if seed % 10 == 0:
        print("seed", seed, "mums", len(want["mums"]), "%.0fs" % (time.time() - t0), flush=True)
print("done", nc, "cases,", bad, "mismatches; totals", tot)
