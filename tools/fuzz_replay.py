#!/usr/bin/env python3
"""Parallel exact replay (parsnp_b200/csrc/host/replay.cpp) against the sequential loop, on the CPU (test infrastructure).

  python tools/fuzz_replay.py <first seed> <cases>

Random collinear genome sets with short minimum MUM lengths (many chance reverse-strand candidates inside sub-regions = foreign
reads and writes of mumlayout), random task sizes down to one gap per task, 2-8 workers, scheduling jitter.  Every case: the
product's host orchestrator over csgmum (oracle/hosttest.py) once with PB200_REPLAY_MODE=seq and several times with =par; MUM
and LCB lists must be identical.  tools/fuzz_host.py compares the same code with the reference binary itself.

FUZZ_BACKEND=3: the parallel runs go through the CPU emulation of the engine's device-resident discovery
(oracle/discover_emul.cpp), i.e. the replay takes the discovery's accept decisions as final for the gaps where they cannot
depend on the order ("final gaps", replay.cpp); regions of a level in random order (PB200_EMUL_SEED), random largest window."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from parsnp_b200 import api, synth
from oracle import hosttest
from tests.refcmp import result_to_dump, diff_dumps


BACKEND = int(os.environ.get("FUZZ_BACKEND", "2"))


def one_case(seed, verbose=False):
    rng = np.random.default_rng(seed)
    L = int(rng.choice([30000, 80000, 200000]))
    nq = int(rng.integers(1, 7))
    div = float(rng.choice([0.01, 0.03, 0.05]))
    g = (synth.g_indep if rng.random() < 0.6 else synth.g_pop)(L, nq, div, int(rng.integers(1, 10**6)))
    if rng.random() < 0.3:                       # an inverted repeat / palindromic stretch: reverse-strand anchors inside collinear data
        a = int(rng.integers(0, L - 400)); ln = int(rng.integers(30, 300)); a = min(a, L - 2 * ln)
        for x in g:
            x[a + ln:a + 2 * ln] = synth.revcomp(x[a:a + ln])
    if rng.random() < 0.2:                       # one rearranged query: the parallel part must decline (or fall back) and still agree
        k = int(rng.integers(1, len(g)))
        g[k] = synth.rearrange(g[k], rng, n_inv=1, inv_len=int(rng.integers(200, 2000)), dels=(), ins=())
    kw = dict(mums=str(rng.choice(["8", "10", "12", "1.1*(Log(S))", "0.7*(Log(S))"])), q=int(rng.choice([10, 30])))
    if rng.random() < 0.3:
        kw["anchors"] = str(rng.choice(["14", "18", "1.1*(Log(S))"]))
    prm = lambda: api.make_params(**kw)
    os.environ["PB200_REPLAY_MODE"] = "seq"
    os.environ["PB200_HOST_THREADS"] = "4"
    ref = hosttest.align(g, prm(), backend=2)
    want = result_to_dump(ref)
    bad = 0
    tot = dict(replay_gaps=0, replay_final_gaps=0, replay_final_mums=0, replay_misses=0, replay_tasks=0, replay_foreign_reads=0, replay_foreign_writes=0, replay_restarts=0, replay_fallback=0, slow_queue_iters=0)
    for rep in range(4):
        os.environ["PB200_REPLAY_MODE"] = "par"
        os.environ["PB200_REPLAY_OWN_THREADS"] = "1"
        os.environ["PB200_HOST_THREADS"] = str(int(rng.choice([2, 3, 8])))
        os.environ["PB200_REPLAY_TASK"] = str(int(rng.choice([1, 1, 2, 5, 40])))
        os.environ["PB200_REPLAY_JITTER"] = str(int(rng.choice([0, 3, 20])))
        os.environ["PB200_SPEC_SLICES"] = str(int(rng.choice([1, 4])))
        os.environ["PB200_EMUL_SEED"] = str(int(rng.integers(1, 10**6)))
        os.environ["PB200_EMUL_MAXLEN"] = str(int(rng.choice([4096, 4096, 300, 80])))
        got = hosttest.align(g, prm(), backend=BACKEND)
        for k in tot:
            tot[k] += got["stats"][k]
        d = diff_dumps(result_to_dump(got), want)
        if d:
            bad += 1
            print("MISMATCH seed", seed, "rep", rep, {k: os.environ[k] for k in os.environ if k.startswith("PB200_")}, d[:3], flush=True)
    return bad, tot, len(want["mums"])


if __name__ == "__main__":
    s0, nc = int(sys.argv[1]), int(sys.argv[2])
    bad = 0
    agg = {}
    t0 = time.time()
    for it in range(nc):
        b, tot, nm = one_case(s0 + it)
        bad += b
        for k, v in tot.items():
            agg[k] = agg.get(k, 0) + v
        if it % 5 == 0:
            print("seed", s0 + it, "mums", nm, tot, "%.0fs" % (time.time() - t0), flush=True)
    print("done", nc, "cases,", bad, "mismatches; totals", agg)
