import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def load_golden(name):
    d = json.load(open(os.path.join(GOLDEN, name + ".json")))
    return dict(n=d["n"], mums=[(m[0], m[1], [tuple(x) for x in m[2]]) for m in d["mums"]],
                clusters=[(c[0], c[1], c[2], [tuple(x) for x in c[3]]) for c in d["clusters"]])


def mers_genomes(names):
    from parsnp_b200 import api
    g = [api.ingest_fasta(os.path.join(GOLDEN, "mers", "England1.fna"), True)]
    g += [api.ingest_fasta(os.path.join(GOLDEN, "mers", q + ".fna"), False) for q in names]
    return g


C1A = ["Al-Hasa_1_2013", "Bisha_1_2012", "EMC_2012", "Jordan-N3_2012"]
C1B = ["Al-Hasa_12_2013", "Al-Hasa_15_2013", "Al-Hasa_16_2013", "Al-Hasa_17_2013"]


def golden_case(name):
    """-> (genomes, params kwargs, golden dump)"""
    sys.path.insert(0, GOLDEN)
    import importlib
    mg = importlib.import_module("tests.golden.make_golden")
    if name == "c1a":
        return mers_genomes(C1A), {}, load_golden("c1a")
    if name == "c1b":
        return mers_genomes(C1B), {}, load_golden("c1b")
    if name == "c1c":
        import tarfile
        import tempfile
        from parsnp_b200 import api
        summ = json.load(open(os.path.join(GOLDEN, "summary.json")))
        with tempfile.TemporaryDirectory() as td:
            tarfile.open(os.path.join(GOLDEN, "mers_all.tar.gz")).extractall(td)
            g = [api.ingest_fasta(os.path.join(GOLDEN, "mers", "England1.fna"), True)]
            g += [api.ingest_fasta(os.path.join(td, q), False) for q in summ["c1c"]["order"]]
        return g, {}, load_golden("c1c")
    g, kw = mg.synth_cases()[name]
    return g, kw, load_golden(name)


def random_case(rng, n_lo=40, n_hi=400, nq_hi=4, alphabet=b"ACGT", with_n=False):
    """random reference window + related query regions (mutations, indels, inversions, repeats)"""
    n = int(rng.integers(n_lo, n_hi))
    alpha = np.frombuffer(alphabet, np.uint8)
    ref = alpha[rng.integers(0, len(alpha), n)]
    if rng.random() < 0.5 and n > 60:          # a repeat inside the reference
        a, b, L = int(rng.integers(0, n - 30)), int(rng.integers(0, n - 30)), int(rng.integers(8, 30))
        ref[b:b + L] = ref[a:a + L]
    if with_n and rng.random() < 0.7:
        a = int(rng.integers(0, n - 5))
        ref[a:a + int(rng.integers(1, 12))] = ord("N")
    comp = np.zeros(256, np.uint8)
    for x, y in zip(b"ACGTN", b"TGCAN"):
        comp[x] = y
    nq = int(rng.integers(1, nq_hi + 1))
    qs = []
    for _ in range(nq):
        q = ref.copy()
        mask = rng.random(n) < rng.choice([0.01, 0.03, 0.08])
        q[mask] = alpha[rng.integers(0, len(alpha), int(mask.sum()))]
        if rng.random() < 0.4 and n > 80:      # inversion
            a = int(rng.integers(0, n - 40)); L = int(rng.integers(15, 40))
            q[a:a + L] = comp[q[a:a + L][::-1]]
        if rng.random() < 0.3:                 # deletion
            a = int(rng.integers(0, n - 10)); L = int(rng.integers(1, 10))
            q = np.concatenate([q[:a], q[a + L:]])
        if rng.random() < 0.3:                 # insertion
            a = int(rng.integers(0, len(q))); L = int(rng.integers(1, 12))
            q = np.concatenate([q[:a], alpha[rng.integers(0, len(alpha), L)], q[a:]])
        if rng.random() < 0.15:                # whole query reverse-complemented
            q = comp[q[::-1]]
        if with_n and rng.random() < 0.5:
            a = int(rng.integers(0, len(q) - 3))
            q[a:a + int(rng.integers(1, 8))] = ord("N")
        qs.append(np.ascontiguousarray(q))
    return [np.ascontiguousarray(ref)] + qs


def whole_window_task(genomes, minsize):
    """one window covering all of genome 0 against all of every query: (windows, coords)"""
    nq = len(genomes) - 1
    coords = np.array([0] * nq + [len(g) for g in genomes[1:]], np.int64)
    return [(0, len(genomes[0]), 0, minsize)], coords
