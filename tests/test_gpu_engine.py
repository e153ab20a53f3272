"""GPU parity tests (B200): every test calls the CUDA engine through the C ABI (include/parsnp_b200.h) and compares
bit-exactly with the oracle (CPU spec / real csgmum / real parsnp_core dumps / committed goldens)."""
import ctypes as C
import os
import tempfile

import numpy as np
import pytest

from tests.conftest import ROOT, golden_case, random_case, whole_window_task
from tests.refcmp import result_to_dump, diff_dumps
from parsnp_b200 import api, synth

pytestmark = pytest.mark.gpu


def _codes(text):
    m = np.full(256, 5, np.uint8)
    for i, c in enumerate(b"ACGTN"):
        m[c] = i + 1
    return m[text]


def cpu_sa_lrp(text):
    """suffix array (window end sorts first, A<C<G<T<N) and longest-repeated-prefix by numpy prefix doubling + Kasai"""
    n = len(text)
    c = _codes(text).astype(np.int64)
    rank = c.copy()
    sa = np.argsort(rank, kind="stable")
    h = 1
    while True:
        r2 = np.zeros(n, np.int64)
        r2[:n - h] = rank[h:] if h < n else 0
        key = rank * (n + 7) + r2
        sa = np.argsort(key, kind="stable")
        ks = key[sa]
        newr = np.zeros(n, np.int64)
        newr[sa] = np.concatenate([[1], 1 + np.cumsum(ks[1:] != ks[:-1])])
        rank = newr
        if rank.max() == n:
            break
        h *= 2
    isa = np.zeros(n, np.int64)
    isa[sa] = np.arange(n)
    lcp = np.zeros(n + 1, np.int64)
    k = 0
    t = text
    for i in range(n):
        r = isa[i]
        if r == 0:
            k = 0
            continue
        j = sa[r - 1]
        while i + k < n and j + k < n and t[i + k] == t[j + k]:
            k += 1
        lcp[r] = k
        if k:
            k -= 1
    lrp = np.maximum(lcp[isa], lcp[isa + 1])
    return sa.astype(np.uint32), lrp.astype(np.int32)


def _debug_index(G, n, minsize=25):
    lib = api.load()
    lib.pb200_debug_index.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    sa = np.zeros(n, np.uint32)
    lrp = np.zeros(n, np.int32)
    rc = lib.pb200_debug_index(G.h, 0, n, minsize, sa.ctypes.data, lrp.ctypes.data)
    assert rc == 0, lib.pb200_last_error()
    return sa, lrp


def _texts():
    rng = np.random.default_rng(5)
    A = np.frombuffer(b"ACGT", np.uint8)
    out = {}
    out["random_5k"] = A[rng.integers(0, 4, 5000)]
    t = A[rng.integers(0, 4, 30000)]
    t[20000:23000] = t[1000:4000]            # 3 kb exact repeat -> several doubling rounds
    t[25000:25400] = ord("A")                # homopolymer
    t[26000:26300] = ord("N")                # N run (N is an ordinary symbol)
    out["repeats_30k"] = t
    out["tiny"] = np.frombuffer(b"ACGTACGGTTACGTAACCGGTAC", np.uint8)
    out["tandem"] = np.tile(np.frombuffer(b"ACGTTGCA", np.uint8), 700)
    out["random_200k"] = A[rng.integers(0, 4, 200000)]
    out["ends_in_A"] = np.concatenate([A[rng.integers(0, 4, 3000)], np.full(23, ord("A"), np.uint8)])
    out["all_A"] = np.full(700, ord("A"), np.uint8)
    out["short"] = np.frombuffer(b"ACGTA", np.uint8)
    t = A[rng.integers(0, 4, 30000)]
    t[20000:23000] = t[1000:4000]            # one 3 kb duplication: tied pairs, ranked by direct text comparison
    t[26000:27500] = t[2500:4000]            # third copy of its second half: tied triples
    out["dup_3k"] = t
    t = A[rng.integers(0, 4, 60000)]
    t[35000:55000] = t[5000:25000]           # 20 kb duplication: deeper than the direct comparison goes -> general path
    out["dup_20k"] = t
    return out


# which index path each text must take when nothing is forced: False = tied groups ranked directly, True = prefix doubling
_EXPECT_DOUBLING = {"tiny": False, "random_5k": False, "random_200k": False, "dup_3k": False, "ends_in_A": False, "short": False,
                    "repeats_30k": True, "tandem": True, "all_A": True, "dup_20k": True}


@pytest.mark.parametrize("doubling", [False, True])
@pytest.mark.parametrize("three_bit", [False, True])
@pytest.mark.parametrize("name", ["tiny", "random_5k", "repeats_30k", "tandem", "random_200k", "ends_in_A", "all_A", "short", "dup_3k", "dup_20k"])
def test_suffix_index_matches_cpu(name, three_bit, doubling):
    """suffix array + longest-repeated-prefix; N-free windows take the 16-mer 2-bit key path unless forced to the 21-mer
    3-bit one; tied k-mer groups are ranked by direct text comparison unless the window has large/deep repeats (or the test
    forces it), then by prefix doubling - all four combinations must give the true suffix order (window end sorts first)"""
    text = np.ascontiguousarray(_texts()[name])
    if three_bit:
        os.environ["PB200_FORCE_3BIT_KEYS"] = "1"
    if doubling:
        os.environ["PB200_FORCE_DOUBLING"] = "1"
    try:
        G = api.Genomes([text, text[: max(3, len(text) // 2)].copy()])
        sa, lrp = _debug_index(G, len(text))
        lib = api.load()
        lib.pb200_debug_index_flags.argtypes = [C.c_void_p]
        lib.pb200_debug_index_flags.restype = C.c_int
        flags = lib.pb200_debug_index_flags(G.h)
    finally:
        os.environ.pop("PB200_FORCE_3BIT_KEYS", None)
        os.environ.pop("PB200_FORCE_DOUBLING", None)
    wsa, wlrp = cpu_sa_lrp(text)
    assert np.array_equal(sa, wsa)
    assert np.array_equal(lrp, wlrp)
    assert bool(flags & 2) == (not three_bit and ord("N") not in text)
    if not doubling:
        assert bool(flags & 1) == _EXPECT_DOUBLING[name]
    G.close()


def _compare_windows(g, minsize, backend, force_big):
    from oracle import hosttest
    w, coords = whole_window_task(g, minsize)
    want = hosttest.search_windows(g, w, coords, backend=backend)[0]
    if force_big:
        os.environ["PB200_FORCE_PATH"] = "big"
    try:
        G = api.Genomes(g)
        got = G.search_windows(w, coords)[0]
        G.close()
    finally:
        os.environ.pop("PB200_FORCE_PATH", None)
    for x, y in zip(got, want):
        assert np.array_equal(x, y), (minsize, got, want)
    return len(want[0])


@pytest.mark.parametrize("force_big", [False, True])
@pytest.mark.parametrize("alphabet,with_n", [(b"AT", False), (b"ACGT", False), (b"ACGT", True)])
def test_window_search_random_small(alphabet, with_n, force_big):
    """single windows (both strands, inversions, indels, repeats, N runs) vs the CPU specification; the same inputs
    through the shared-memory path and (forced) through the suffix-array path"""
    rng = np.random.default_rng(100 + len(alphabet) + with_n)
    total = 0
    for it in range(40):
        g = random_case(rng, 30, 260, 4, alphabet, with_n)
        total += _compare_windows(g, int(rng.integers(4, 14)), 0, force_big)
    assert total > 10


def test_window_search_batch_mixed_sizes():
    """one call with many windows of different size classes (sub-regions of the same genomes) == csgmum per window"""
    from oracle import hosttest
    g = synth.g_indep(40000, 3, 0.02, 4)
    rng = np.random.default_rng(9)
    nq = 3
    wins, coords = [], []
    for i in range(60):
        L = int(rng.choice([60, 150, 400, 900, 2500, 6000]))
        s = int(rng.integers(0, 40000 - L - 50))
        off = len(coords)
        qs = [s + int(rng.integers(-20, 20)) for _ in range(nq)]
        qs = [max(0, x) for x in qs]
        ql = [L + int(rng.integers(-15, 15)) for _ in range(nq)]
        coords += qs + ql
        wins.append((s, L, off, api.minsize("1.1*(Log(S))", min([L] + ql))))
    coords = np.array(coords, np.int64)
    want = hosttest.search_windows(g, wins, coords, backend=1)
    G = api.Genomes(g)
    got = G.search_windows(wins, coords)
    G.close()
    ncand = 0
    for a, b in zip(got, want):
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
        ncand += len(a[0])
    assert ncand > 30


def test_window_search_medium_vs_csgmum():
    """a 120 kbp window with rearranged queries through the suffix-array path == real csgmum"""
    rng = np.random.default_rng(3)
    g = synth.g_indep(120000, 3, 0.02, 8)
    g = [g[0]] + [synth.rearrange(x, rng, n_inv=2, inv_len=5000) for x in g[1:]]
    n = _compare_windows(g, api.minsize("1.1*(Log(S))", min(len(x) for x in g)), 1, False)
    assert n > 500


@pytest.mark.parametrize("name", ["c1a", "c1b", "indep_20k", "rearr_60k", "windows_50k", "pop_30k_x12", "c1c"])
def test_align_matches_golden(name):
    """whole MUM+LCB path on the GPU == committed reference dumps (MUM coordinates and LCB boundaries bit-exact)"""
    g, kw, gold = golden_case(name)
    res = api.align(g, api.make_params(**kw))
    assert diff_dumps(result_to_dump(res), gold) == []


def test_align_matches_reference_binary_fresh_input():
    """a fresh synthetic set: GPU == oracle/_ref/parsnp_core_ref run here, incl. the order of searched windows"""
    from oracle import runner
    rng = np.random.default_rng(12)
    g = synth.g_indep(150000, 4, 0.02, 31)
    g = [g[0]] + [synth.rearrange(x, rng, n_inv=1, inv_len=8000) for x in g[1:]]
    with tempfile.TemporaryDirectory() as td:
        ref, qs = synth.write_dataset(os.path.join(td, "d"), g)
        r = runner.run_ref(ref, qs, os.path.join(td, "r"), cands=True)
    G = api.Genomes(g)
    res = G.align(api.make_params(flags=api.FLAG_TRACE_WINDOWS))
    assert diff_dumps(result_to_dump(res), r["dump"]) == []
    assert [tuple(x) for x in res["trace"].tolist()] == [(w["ini0"], w["len0"]) for w in r["cands"]]
    res2 = G.align(api.make_params(flags=api.FLAG_NO_SPECULATION))
    assert diff_dumps(result_to_dump(res2), r["dump"]) == []
    G.close()


def test_c4_shape_matches_reference_binary():
    """BASELINE configs[3] shape at a size the reference finishes in seconds: multi-contig reference (concatenated seamlessly),
    multi-contig queries (310 N's at every contig break - N is an ordinary matching symbol), several reference windows
    (p < reference length, so 3-bit keys for no window but N-runs in every query) == reference binary run here"""
    from oracle import runner
    g = synth.g_pop(240_000, 5, 0.01, 9)
    with tempfile.TemporaryDirectory() as td:
        ref, qs = synth.write_dataset(os.path.join(td, "d"), g, contigs=4)
        r = runner.run_ref(ref, qs, os.path.join(td, "r"), p=70000)
        gi = [api.ingest_fasta(ref, True)] + [api.ingest_fasta(q, False) for q in qs]
    for a, b in zip(gi[1:], g[1:]):
        assert np.array_equal(a, synth.with_contig_padding(b, 4))
    res = api.align(gi, api.make_params(p=70000))
    assert len(res["mum_length"]) > 500
    assert diff_dumps(result_to_dump(res), r["dump"]) == []


def _edge_cases():
    """small inputs at the corners of the path: (genomes, make_params kwargs, run_ref kwargs)"""
    rng = np.random.default_rng(2024)
    A = np.frombuffer(b"ACGT", np.uint8)
    base = synth.g_indep(60000, 3, 0.02, 41)
    cases = {}
    # ragged: one query is a prefix of its genome, one carries 40 % foreign sequence at the end
    cases["ragged_lengths"] = ([base[0], base[1][:25000].copy(), np.concatenate([base[2], A[rng.integers(0, 4, 24000)]]), base[3]], {}, {})
    # a single query
    cases["single_query"] = (base[:2], {}, {})
    # N runs inside the reference (3-bit key path for the anchors) and inside a query
    g = [x.copy() for x in base]
    g[0][10000:10400] = ord("N"); g[0][30000:30007] = ord("N"); g[2][45000:45300] = ord("N")
    cases["n_runs"] = (g, {}, {})
    # a query that is the reverse complement of its genome: every anchor on the reverse strand
    cases["revcomp_query"] = ([base[0], synth.revcomp(base[1]), base[2]], {}, {})
    # identical genomes: one match as long as the genome
    cases["identical"] = ([base[0][:20000].copy(), base[0][:20000].copy(), base[0][:20000].copy()], {}, {})
    # tiny genomes: the anchor search itself runs in the shared-memory kernel
    t = synth.g_indep(700, 3, 0.03, 5)
    cases["tiny"] = (t, {}, {})
    # constant minimum lengths and other ini values; several reference windows
    cases["ini_values"] = (base, dict(c=50, d=100, q=20, p=25000, diagdiff=0.3, anchors="25", mums="15"),
                           dict(c=50, d=100, q=20, p=25000, diagdiff=0.3, anchors="25", mums="15"))
    # diagdiff > 1 takes the absolute-difference branch of setFinalClusters
    cases["diagdiff_abs"] = (base, dict(diagdiff=40.0), dict(diagdiff=40.0))
    return cases


@pytest.mark.parametrize("name", ["ragged_lengths", "single_query", "n_runs", "revcomp_query", "identical", "tiny", "ini_values", "diagdiff_abs"])
def test_edge_cases_match_reference_binary(name):
    """corner inputs: GPU path == reference binary run here (MUM coordinates, LCB records)"""
    from oracle import runner
    g, kw, rkw = _edge_cases()[name]
    with tempfile.TemporaryDirectory() as td:
        ref, qs = synth.write_dataset(os.path.join(td, "d"), g)
        r = runner.run_ref(ref, qs, os.path.join(td, "r"), **rkw)
    res = api.align(g, api.make_params(**kw))
    assert r["dump"] is not None and len(r["dump"]["mums"]) > 0
    assert diff_dumps(result_to_dump(res), r["dump"]) == []


def test_no_mums_found():
    """unrelated genomes: the reference writes NO MUMS FOUND and exits 0 (src/parsnp.cpp:3223-3229); the library returns
    PB200_ERR_NO_MUMS with empty lists"""
    rng = np.random.default_rng(8)
    A = np.frombuffer(b"ACGT", np.uint8)
    g = [A[rng.integers(0, 4, 30000)] for _ in range(3)]
    res = api.align(g, api.make_params())
    assert res["no_mums"] and len(res["mum_length"]) == 0 and len(res["cluster_type"]) == 0


def test_full_size_properties():
    """BASELINE config-2 shape at reduced query count (5 Mbp reference, 2 queries): size-independent properties -
    MUMs are exact matches in every genome, disjoint on the reference, LCB MUM sums consistent."""
    g = synth.g_indep(5_000_000, 2, 0.01, 1)
    res = api.align(g, api.make_params())
    st, ln, fw = res["mum_start"], res["mum_length"], res["mum_fwd"]
    assert len(ln) > 50000
    order = np.argsort(st[:, 0])
    assert np.all(st[order, 0][1:] >= (st[order, 0] + ln[order])[:-1])          # disjoint, sorted on the reference
    rng = np.random.default_rng(0)
    for i in rng.integers(0, len(ln), 3000):
        a = g[0][st[i, 0]:st[i, 0] + ln[i]]
        for k in range(1, 3):
            b = g[k][st[i, k]:st[i, k] + ln[i]]
            if not fw[i, k]:
                b = synth.revcomp(b)
            assert np.array_equal(a, b)
    lcb = res["cluster_type"] == 1
    assert res["cluster_length"][lcb].sum() == ln.sum()


@pytest.mark.parametrize("name", ["c1a", "rearr_60k", "pop_30k_x12", "indep_20k", "ratio_40k"])
def test_mumi_matches_reference(name):
    """calcmumi=1 mode (Aligner::setMumi): the distances parsnp_core writes to all.mumi, to the printed digit"""
    import json
    from tests.conftest import GOLDEN
    gold = json.load(open(os.path.join(GOLDEN, "mumi.json")))[name]
    if name == "ratio_40k":
        g = synth.g_indep(40000, 2, 0.02, 17)
        g[2] = g[2][:20000].copy()
    else:
        g, kw, _ = golden_case(name)
    G = api.Genomes(g)
    got = ["%f" % v for v in G.mumi()]
    G.close()
    assert got == gold
