#!/usr/bin/env python3
"""CPU fuzz of the brute-force specification (oracle/mumspec.cpp) against the real csgmum (test infrastructure, no GPU).

  python tools/fuzz_spec.py <seed> <windows>

Single windows in regimes far from real genomes as well: homopolymers and two-letter alphabets, N runs, 30 to 1 200 bases, up to
6 queries, minsize 2 to 13.  The GPU window tests lean on this specification (tests/test_gpu_engine.py) and, for the same
generator, directly on csgmum (tests/test_zz_gpu_fuzz.py)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import hosttest
from tests.conftest import random_case, whole_window_task
seed=int(sys.argv[1]); N=int(sys.argv[2])
rng=np.random.default_rng(seed)
bad=0; tot=0; t=time.time()
for it in range(N):
    alphabet=[b"AT", b"ACGT", b"ACGT", b"AACGGT", b"A", b"AC"][int(rng.integers(0,6))]
    with_n = bool(rng.random()<0.5)
    hi = int(rng.choice([60, 160, 400, 1200]))
    g = random_case(rng, 30, hi, 6, alphabet, with_n)
    minsize = int(rng.integers(2, 14))
    w, coords = whole_window_task(g, minsize)
    a = hosttest.search_windows(g, w, coords, backend=0)[0]
    b = hosttest.search_windows(g, w, coords, backend=1)[0]
    if not all(np.array_equal(x,y) for x,y in zip(a,b)):
        bad+=1
        print("MISMATCH it", it, alphabet, with_n, minsize, [len(x) for x in g], flush=True)
        if bad > 5: break
    tot += len(a[0])
print("windows", N, "bad", bad, "cands", tot, "runoff", hosttest.runoff_skips(), "%.0fs"%(time.time()-t))
