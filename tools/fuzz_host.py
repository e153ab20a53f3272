#!/usr/bin/env python3
"""CPU fuzz of the host orchestrator against the reference binary (test infrastructure, no GPU).

  GLIBC_TUNABLES=glibc.malloc.tcache_count=0 MALLOC_PERTURB_=255 python tools/fuzz_host.py <first seed> <cases>

Random genome sets (independent / population divergence, repeats and N runs in the reference, inversions, deletions, insertions,
whole-query reverse complements, multi-contig FASTA) x random ini values (c, d, q, diagdiff, p -> several reference windows)
x random speculation slicing / anchor-accept mode (FUZZ_WORLD=N: also the N-rank sharded host path, ranks as threads).  Every case runs oracle/_ref/parsnp_core_ref and the product's host
orchestrator with the reference's own csgmum as search backend (oracle/hosttest.py) and compares MUM and LCB lists bit for bit.
MALLOC_PERTURB_=255 makes glibc zero every allocation of the reference binary (tcache off: its hits bypass the fill): wherever
a reverse-strand match wins, the binary's result otherwise depends on the stale contents of the never-initialised
MasterRC[].UP.  FUZZ_BACKEND=3: the same through the CPU emulation of the engine's device-resident
discovery, parallel replay forced (the "final gaps" path of host/replay.cpp)."""
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from parsnp_b200 import api, synth
from oracle import runner, hosttest
from tests.refcmp import result_to_dump, diff_dumps
from tools.fuzz_cases import make_case, params_kw
seed0 = int(sys.argv[1]); ncases = int(sys.argv[2])
bad = 0
ub = 0
t0 = time.time()
for it in range(ncases):
    g, contigs, kw, (L, nq, div), rng = make_case(seed0 + it)
    with tempfile.TemporaryDirectory() as td:
        rf, qf = synth.write_dataset(os.path.join(td, "d"), g, contigs=contigs)
        r = runner.run_ref(rf, qf, os.path.join(td, "r"), **kw)
        gi = [api.ingest_fasta(rf, True, d=kw.get("d", 300))] + [api.ingest_fasta(x, False, d=kw.get("d", 300)) for x in qf]
    os.environ["PB200_SPEC_SLICES"] = str(int(rng.choice([1, 3, 8])))
    os.environ["PB200_PAR_ANCHORS_MIN"] = str(int(rng.choice([1, 10**9])))
    if rng.random() < 0.25:                     # the defaults: slices by the number of initial regions, parallel accept by size
        del os.environ["PB200_SPEC_SLICES"], os.environ["PB200_PAR_ANCHORS_MIN"]
    hosttest.runoff_skips()
    backend = int(os.environ.get("FUZZ_BACKEND", "1"))
    if backend == 3:                            # the engine's discovery emulated on the CPU: the replay's "final gaps" path (oracle/discover_emul.cpp)
        os.environ["PB200_REPLAY_MODE"] = "par"
        os.environ["PB200_REPLAY_OWN_THREADS"] = "1"
        os.environ["PB200_HOST_THREADS"] = str(int(rng.choice([2, 4])))
        os.environ["PB200_REPLAY_TASK"] = str(int(rng.choice([1, 3, 40])))
        os.environ["PB200_EMUL_SEED"] = str(int(rng.integers(1, 10**6)))
        os.environ["PB200_EMUL_MAXLEN"] = str(int(rng.choice([4096, 4096, 200])))
    res = hosttest.align(gi, api.make_params(**params_kw(kw)), backend=backend)
    runoff = hosttest.runoff_skips()
    world = int(os.environ.get("FUZZ_WORLD", "0"))
    if world > 1:                               # N>1 host path (thread ranks): every rank must hold the single-rank result
        mine = result_to_dump(res)
        outs, counters = hosttest.ThreadRanks(world).align(gi, api.make_params(**params_kw(kw)))
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
