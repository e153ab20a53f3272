"""Random cases for the CPU fuzzers (tools/fuzz_host.py, tools/fuzz_xmfa.py): test infrastructure.

make_case(seed) -> (genomes, contigs, ini keywords, description, rng): random genome sets (independent / population divergence,
repeats and N runs in the reference, inversions, deletions, insertions, whole-query reverse complements, up to 13 queries,
multi-contig FASTA) x random ini values (c, d, q, diagdiff, p -> several reference windows, length expressions, filter, anchorsonly)."""
import os

import numpy as np

from parsnp_b200 import synth


def make_case(seed):
    rng = np.random.default_rng(seed)
    L = int(rng.choice([8000, 20000, 50000, 90000])) * int(os.environ.get("FUZZ_SCALE", "1"))    # FUZZ_SCALE=4: up to 360 kbp
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
    if os.environ.get("FUZZ_FILTERS") != "0" and rng.random() < 0.12:
        # ini [MUM] filter > 1 (src/parsnp.cpp:327-425, filterRandom1): drawn from its own stream so that the cases of earlier
        # campaigns keep their seeds
        kw["filter"] = int(np.random.default_rng(seed + 7919).choice([2, 5, 30]))
    if rng.random() < 0.12:                     # many queries: mutated copies of the first ones
        for _ in range(int(rng.integers(3, 9))):
            x = g[int(rng.integers(0, len(g)))].copy()
            hit = rng.random(len(x)) < float(rng.choice([0.002, 0.01, 0.03]))
            x[hit] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, int(hit.sum()))]
            g.append(x)
    if rng.random() < 0.1:
        kw["anchorsonly"] = 1                    # ini [MUM] anchorsonly: no recursion into the regions between the anchors
    return g, contigs, kw, (L, nq, div), rng


def params_kw(kw):
    """ini keywords of a case (oracle/runner.py write_ini) -> keywords of api.make_params"""
    out = {k: v for k, v in kw.items() if k not in ("recombfilter", "unaligned", "anchorsonly", "doalign")}
    if kw.get("anchorsonly"):
        out["anchors_only"] = 1
    return out
