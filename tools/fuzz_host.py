#!/usr/bin/env python3
"""CPU fuzz of the host orchestrator against the reference binary (test infrastructure, no GPU).

  GLIBC_TUNABLES=glibc.malloc.tcache_count=0 MALLOC_PERTURB_=255 python tools/fuzz_host.py <first seed> <cases>

Random genome sets (independent / population divergence, repeats and N runs in the reference, inversions, deletions, insertions,
whole-query reverse complements, multi-contig FASTA) x random ini values (c, d, q, diagdiff, p -> several reference windows)
x random speculation slicing / anchor-accept mode (FUZZ_WORLD=N: also the N-rank sharded host path, ranks as threads).  Every case runs oracle/_ref/parsnp_core_ref and the product's host
orchestrator with the reference's own csgmum as search backend (oracle/hosttest.py) and compares MUM and LCB lists bit for bit.
MALLOC_PERTURB_=255 makes glibc zero every allocation of the reference binary (tcache off: its hits bypass the fill): wherever
a reverse-strand match wins, the binary's result otherwise depends on the stale contents of the never-initialised
MasterRC[].UP."""
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from parsnp_b200 import api, synth
from oracle import runner, hosttest
from tests.refcmp import result_to_dump, diff_dumps
seed0 = int(sys.argv[1]); ncases = int(sys.argv[2])
bad = 0
ub = 0
t0 = time.time()
for it in range(ncases):
    rng = np.random.default_rng(seed0 + it)
    L = int(rng.choice([8000, 20000, 50000, 90000]))
    nq = int(rng.integers(1, 6))
    div = float(rng.choice([0.005, 0.01, 0.03, 0.06]))
    if rng.random() < 0.5:
        g = synth.g_indep(L, nq, div, int(rng.integers(1, 10**6)))
    else:
        g = synth.g_pop(L, nq, div, int(rng.integers(1, 10**6)))
    ref = g[0].copy()
    if rng.random() < 0.5:                      # repeats in the reference
        for _ in range(int(rng.integers(1, 4))):
            a, b, ln = (int(x) for x in (rng.integers(0, L - 600), rng.integers(0, L - 600), rng.integers(20, 500)))
            ref[b:b + ln] = ref[a:a + ln]
    if rng.random() < 0.3:
        a = int(rng.integers(0, L - 50)); ref[a:a + int(rng.integers(1, 40))] = ord("N")
    qs = []
    for q in g[1:]:
        if rng.random() < 0.6:
            q = synth.rearrange(q, rng, n_inv=int(rng.integers(0, 3)), inv_len=int(rng.integers(200, 3000)),
                                dels=tuple(int(x) for x in rng.integers(1, 400, int(rng.integers(0, 3)))),
                                ins=tuple(int(x) for x in rng.integers(1, 200, int(rng.integers(0, 2)))))
        if rng.random() < 0.15:
            q = synth.revcomp(q)
        qs.append(q)
    g = [ref] + qs
    contigs = int(rng.choice([1, 1, 2, 4]))
    kw = {}
    if rng.random() < 0.4:
        kw = dict(c=int(rng.choice([10, 21, 60])), d=int(rng.choice([50, 300, 1000])), q=int(rng.choice([10, 30, 100])),
                  diagdiff=float(rng.choice([0.05, 0.12, 0.5, 30.0])))
    if rng.random() < 0.4:
        kw["p"] = int(rng.choice([5000, 17000, 40000]))
    if rng.random() < 0.25:                     # MUM / anchor length expressions (Converter + Calculator) and the LCB filter
        kw["anchors"] = str(rng.choice(["1.1*(Log(S))", "2*(Log(S))", "25", "1.5*(Log(S))+3"]))
        kw["mums"] = str(rng.choice(["1.1*(Log(S))", "0.9*(Log(S))", "14", "1.1*(Log(S))"]))
    if rng.random() < 0.15:
        kw["filter"] = 0
    if rng.random() < 0.12:                     # many queries: mutated copies of the first ones
        for _ in range(int(rng.integers(3, 9))):
            x = g[int(rng.integers(0, len(g)))].copy()
            hit = rng.random(len(x)) < float(rng.choice([0.002, 0.01, 0.03]))
            x[hit] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, int(hit.sum()))]
            g.append(x)
    with tempfile.TemporaryDirectory() as td:
        rf, qf = synth.write_dataset(os.path.join(td, "d"), g, contigs=contigs)
        r = runner.run_ref(rf, qf, os.path.join(td, "r"), **kw)
        gi = [api.ingest_fasta(rf, True, d=kw.get("d", 300))] + [api.ingest_fasta(x, False, d=kw.get("d", 300)) for x in qf]
    os.environ["PB200_SPEC_SLICES"] = str(int(rng.choice([1, 3, 8])))
    os.environ["PB200_PAR_ANCHORS_MIN"] = str(int(rng.choice([1, 10**9])))
    hosttest.runoff_skips()
    res = hosttest.align(gi, api.make_params(**kw), backend=1)
    runoff = hosttest.runoff_skips()
    world = int(os.environ.get("FUZZ_WORLD", "0"))
    if world > 1:                               # N>1 host path (thread ranks): every rank must hold the single-rank result
        mine = result_to_dump(res)
        outs, counters = hosttest.ThreadRanks(world).align(gi, api.make_params(**kw))
        for rk, o in enumerate(outs):
            dd = diff_dumps(result_to_dump(o), mine)
            if dd or o["no_mums"] != res["no_mums"]:
                bad += 1
                print("SHARDED MISMATCH seed", seed0 + it, "world", world, "rank", rk, dd[:3], flush=True)
                break
    if r["dump"] is None:
        ok = res.get("no_mums", False) or len(res["mum_length"]) == 0
        nm = 0
    else:
        dd = diff_dumps(result_to_dump(res), r["dump"])
        ok = dd == []
        nm = len(r["dump"]["mums"])
    if not ok and (runoff or r["returncode"] < 0):
        # Find_UM walked off the end of a query buffer (a window sharing no symbol with a query strand, e.g. inside an N run;
        # src/csgmum/mum.c:193-198 has no bound): the binary's answer then depends on the bytes behind that buffer, changes
        # with the length of the file names, and sometimes is a crash.  Not a parity case.
        ub += 1
        print("reference UB (Find_UM run-off x%d, rc %d), case not comparable: seed" % (runoff, r["returncode"]), seed0 + it, flush=True)
    elif not ok:
        bad += 1
        print("MISMATCH seed", seed0 + it, L, nq, div, contigs, kw, flush=True)
        if os.environ.get("FUZZ_VERBOSE"):
            print("  reference rc", r["returncode"], "dump", r["dump"] is not None, "diffs", dd[:4] if r["dump"] is not None else None, flush=True)
    elif it % 10 == 0:
        print("ok", seed0 + it, L, nq, contigs, kw, "mums", nm, "%.0fs" % (time.time() - t0), flush=True)
print("done", ncases, "cases,", bad, "mismatches,", ub, "not comparable (reference UB)")
