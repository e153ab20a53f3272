"""CPU tests (no GPU): the oracle restatement against the real reference (golden vectors + csgmum), and the
product's host orchestrator (parsnp_b200/csrc/host) driven by the reference's own csgmum search."""
import ctypes as C
import hashlib
import os
import tempfile

import numpy as np
import pytest

from tests.conftest import ROOT, GOLDEN, golden_case, load_golden, random_case, whole_window_task, C1A
from tests.refcmp import result_to_dump, diff_dumps
from parsnp_b200 import api

REFDIR = os.path.join(ROOT, "oracle", "_ref")
have_ref = os.path.exists(os.path.join(REFDIR, "libpb200_hosttest.so"))
pytestmark = pytest.mark.skipif(not have_ref, reason="oracle/_ref not built (python oracle/build_ref.py && make -C oracle)")


def test_reference_binary_reproduces_golden_c1a():
    """pins the goldens to the reference itself: XMFA md5 of SURVEY.md App. C + the committed MUM/LCB dump"""
    from oracle import runner
    with tempfile.TemporaryDirectory() as td:
        r = runner.run_ref(os.path.join(GOLDEN, "mers", "England1.fna"), [os.path.join(GOLDEN, "mers", q + ".fna") for q in C1A],
                           td, dump_exit=False)
        md5 = hashlib.md5(open(os.path.join(r["outdir"], "parsnpAligner.xmfa"), "rb").read()).hexdigest()
    assert md5 == "5b59e50c5b8c1f79165fc41cfd2a6ac4"
    assert diff_dumps(r["dump"], load_golden("c1a")) == []
    assert len(r["dump"]["mums"]) == 149


def test_known_answer_csgmum():
    """SURVEY.md App. A known-answer vector, straight through the real csg.c/mum.c"""
    lib = C.CDLL(os.path.join(REFDIR, "libcsgmum_ref.so"))
    lib.ref_index_build.restype = C.c_void_p
    lib.ref_index_build.argtypes = [C.c_char_p, C.c_long, C.c_double]
    lib.ref_find_um.argtypes = [C.c_void_p, C.c_char_p, C.c_long, C.c_void_p, C.c_void_p]
    lib.ref_intersect_um.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib.ref_index_stats.argtypes = [C.c_void_p] + [C.c_void_p] * 3
    R = b"ACGTACGGTTACGTAACCGGTAC"
    Q = b"GGTTACGTAACCACGTACGG"
    n = len(R)
    ix = lib.ref_index_build(R, n, 2.0)
    a, b, c = C.c_int(), C.c_int(), C.c_int()
    lib.ref_index_stats(ix, C.byref(a), C.byref(b), C.byref(c))
    assert (a.value, b.value, c.value) == (25, 17, 8)
    pair = np.zeros(2 * n, np.int32)
    sp = np.zeros(n, np.uint64)
    lib.ref_find_um(ix, Q, len(Q), sp.ctypes.data, pair.ctypes.data)
    assert tuple(pair[0:2]) == (5, 8) and sp[0] == 12
    assert tuple(pair[12:14]) == (9, 18) and sp[6] == 0
    assert pair.reshape(-1, 2)[[i for i in range(n) if i not in (0, 6)]].sum() == 0
    master = np.zeros(2 * n, np.int32)
    master[1::2] = n
    lib.ref_intersect_um(ix, master.ctypes.data, pair.ctypes.data, n, sp.ctypes.data)
    m = master.reshape(-1, 2)
    assert all(tuple(m[k]) == (5, 8) and sp[k] == 12 + k for k in range(6))
    assert all(tuple(m[k]) == (9, 18) and sp[k] == k - 6 for k in range(6, n))
    # and the specification's uniqueness floor agrees: u[6] = 9, u[0] = 5
    from oracle import hosttest
    lrp = hosttest.lrp(np.frombuffer(R, np.uint8))
    assert 6 + lrp[6] == 9 and 0 + lrp[0] == 5


@pytest.mark.parametrize("alphabet,with_n", [(b"AT", False), (b"ACGT", False), (b"ACGT", True), (b"AACGGT", True)])
def test_spec_matches_real_csgmum(alphabet, with_n):
    """brute-force specification (oracle/mumspec.cpp) == real csgmum on random windows: candidates, SP, strand flags"""
    from oracle import hosttest
    rng = np.random.default_rng(len(alphabet) * 10 + with_n)
    total = 0
    for it in range(150):
        g = random_case(rng, 30, 160, 4, alphabet, with_n)
        minsize = int(rng.integers(4, 12))
        w, coords = whole_window_task(g, minsize)
        a = hosttest.search_windows(g, w, coords, backend=0)[0]
        b = hosttest.search_windows(g, w, coords, backend=1)[0]
        for x, y in zip(a, b):
            assert np.array_equal(x, y), (it, a, b)
        total += len(a[0])
    assert total > 50


def test_minsize_matches_reference_calculator():
    lib = C.CDLL(os.path.join(REFDIR, "libcsgmum_ref.so"))
    lib.ref_minsize.argtypes = [C.c_char_p, C.c_long]
    lib.ref_postfix.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
    ht = C.CDLL(os.path.join(REFDIR, "libpb200_hosttest.so"))
    ht.pb200_minsize.argtypes = [C.c_char_p, C.c_int64]
    # (expressions such as "1.1*Log(S)" - Log after an operator without parentheses - make the reference's own Converter
    #  pop an empty stack and exit, so only forms the reference survives are compared)
    exprs = [b"1.1*(Log(S))", b"1.2*(Log(S))", b"25", b"Log(S)", b"2*(Log(S))-4", b"(Log(S))/2+10", b"1.5*(Log(S))+2",
             b"(Log(S))*(Log(S))/10"]
    rng = np.random.default_rng(1)
    for e in exprs:
        vals = list(range(1, 3000)) + [int(x) for x in rng.integers(3000, 60_000_000, 4000)] + [2 ** k for k in range(1, 30)] + \
               [2 ** k + d for k in range(4, 28) for d in (-1, 1)]
        for s in vals:
            assert ht.pb200_minsize(e, s) == lib.ref_minsize(e, s), (e, s)
    # SURVEY.md a18 examples
    for s, want in ((31, 6), (100, 8), (30030, 17), (5000000, 25), (15000000, 27)):
        assert ht.pb200_minsize(b"1.1*(Log(S))", s) == want


def test_record_replay_backend_is_transparent():
    """the record/replay search backend of tools/host_bench.py (oracle/ref_backend.cpp): same alignment as csgmum directly, and
    the second run of an input finds every window in the table"""
    import ctypes
    from oracle import hosttest
    g, kw, gold = golden_case("rearr_60k")
    lib = hosttest.load()
    lib.pbtest_replay_misses.restype = ctypes.c_long
    lib.pbtest_replay_clear()
    lib.pbtest_replay_misses(1)
    a = hosttest.align(g, api.make_params(**kw), backend=2)
    assert lib.pbtest_replay_misses(1) > 100
    b = hosttest.align(g, api.make_params(**kw), backend=2)
    assert lib.pbtest_replay_misses(1) == 0
    assert diff_dumps(result_to_dump(a), gold) == [] and diff_dumps(result_to_dump(b), gold) == []
    lib.pbtest_replay_clear()


def test_minsize_fuzz_random_expressions():
    """tools/fuzz_minsize.py: 400 random infix expressions x 14 lengths == the reference's Converter + Calculator wherever the
    reference survives the expression"""
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_minsize.py"), "7", "400"], stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0 and " differ 0 " in r.stdout, r.stdout[-1500:]
    assert int(r.stdout.split(" equal ")[1].split()[0]) > 200


@pytest.mark.parametrize("name", ["c1a", "c1b", "indep_20k", "rearr_60k", "windows_50k", "pop_30k_x12", "c1c"])
def test_host_orchestrator_reproduces_reference(name):
    """the product's host logic (queue order, trim, accept, LCB chaining) fed by the reference's own search == golden"""
    from oracle import hosttest
    g, kw, gold = golden_case(name)
    res = hosttest.align(g, api.make_params(**kw), backend=1)
    assert diff_dumps(result_to_dump(res), gold) == []
    # speculation off must give the same answer (every region searched on demand in reference order)
    if name in ("indep_20k", "rearr_60k"):
        res2 = hosttest.align(g, api.make_params(flags=api.FLAG_NO_SPECULATION, **kw), backend=1)
        assert diff_dumps(result_to_dump(res2), gold) == []


@pytest.mark.parametrize("name", ["c1a", "indep_20k", "rearr_60k", "windows_50k", "pop_30k_x12", "c1c"])
def test_parallel_anchor_accept_reproduces_reference(name, monkeypatch):
    """the anchors' accept pass split into non-overlapping candidates (parallel) and overlapping ones (literal loop) == golden;
    rearr_60k has reverse-strand anchors and out-of-order query coordinates, windows_50k several reference windows"""
    from oracle import hosttest
    monkeypatch.setenv("PB200_PAR_ANCHORS_MIN", "1")
    monkeypatch.setenv("PB200_HOST_THREADS", "4")
    g, kw, gold = golden_case(name)
    res = hosttest.align(g, api.make_params(**kw), backend=1)
    assert diff_dumps(result_to_dump(res), gold) == []


@pytest.mark.parametrize("slices,threads", [(1, 1), (4, 2), (9, 4)])
@pytest.mark.parametrize("name", ["indep_20k", "rearr_60k", "windows_50k"])
def test_pipelined_speculation_reproduces_reference(name, slices, threads, monkeypatch):
    """the speculation runs slice by slice on its own thread while the exact replay consumes the finished slices: any
    slicing gives the golden result and no region is left to an on-demand search"""
    from oracle import hosttest
    monkeypatch.setenv("PB200_SPEC_SLICES", str(slices))
    monkeypatch.setenv("PB200_HOST_THREADS", str(threads))
    g, kw, gold = golden_case(name)
    res = hosttest.align(g, api.make_params(**kw), backend=1)
    assert diff_dumps(result_to_dump(res), gold) == []
    assert res["stats"]["spec_slices"] == slices
    assert res["stats"]["replay_misses"] == 0


def test_parallel_anchor_accept_equals_serial_random(monkeypatch):
    """random rearranged / repeat-carrying sets: parallel anchor accept == the literal serial loop (MUMs, LCBs, window trace)"""
    from oracle import hosttest
    from parsnp_b200 import synth
    monkeypatch.setenv("PB200_HOST_THREADS", "4")
    rng = np.random.default_rng(77)
    for it in range(5):
        g = synth.g_indep(25000, 3, 0.03, 200 + it)
        ref = g[0].copy()
        for _ in range(3):                     # short repeats inside the reference: overlapping / trimmed anchors
            a, b, L = (int(x) for x in (rng.integers(0, 20000), rng.integers(0, 20000), rng.integers(30, 200)))
            ref[b:b + L] = ref[a:a + L]
        g = [ref] + [synth.rearrange(x, rng, n_inv=2, inv_len=1500) for x in g[1:]]
        outs = []
        for pmin in ("1", "1000000000"):
            monkeypatch.setenv("PB200_PAR_ANCHORS_MIN", pmin)
            outs.append(hosttest.align(g, api.make_params(flags=api.FLAG_TRACE_WINDOWS), backend=1))
        assert diff_dumps(result_to_dump(outs[0]), result_to_dump(outs[1])) == []
        assert np.array_equal(outs[0]["trace"], outs[1]["trace"])


def test_small_windows_with_inversions_zero_init_reference(monkeypatch):
    """Several reference windows small enough for the reference's per-window arrays to come from the heap (p = 5000 -> 40 kB)
    + reverse-strand matches: `MasterRC[].UP` is never initialised (src/parsnp.cpp:1591-1597), a later
    window's array re-uses a freed chunk and the binary's MUM count then varies with the heap contents (434 / 439 / 426 on one
    input under MALLOC_PERTURB_ = unset / 255 / 170).  Real windows (15 Mbp -> 120 MB) are mmap'ed, i.e. zero.  The product
    implements the zero-initialised semantics; the oracle for this corner is the reference under MALLOC_PERTURB_=255 (glibc
    fills every allocation with ~255 = 0) with the thread cache off (tcache hits bypass the fill: arrays of <= 129 positions,
    i.e. the small recursion windows, would keep their stale contents)."""
    from oracle import hosttest, runner
    from parsnp_b200 import synth
    monkeypatch.setenv("MALLOC_PERTURB_", "255")
    monkeypatch.setenv("GLIBC_TUNABLES", "glibc.malloc.tcache_count=0")      # (tcache hits bypass the perturbation)
    rng = np.random.default_rng(1032)
    total_rc = 0
    for it in range(3):
        g = synth.g_indep(20000, 3, 0.03, 900 + it)
        g = [g[0]] + [synth.rearrange(x, rng, n_inv=2, inv_len=1500, dels=(30,), ins=(20,)) for x in g[1:]]
        with tempfile.TemporaryDirectory() as td:
            ref, qs = synth.write_dataset(os.path.join(td, "d"), g)
            r = runner.run_ref(ref, qs, os.path.join(td, "r"), p=5000)
        res = hosttest.align(g, api.make_params(p=5000), backend=1)
        assert diff_dumps(result_to_dump(res), r["dump"]) == []
        total_rc += int((res["mum_fwd"] == 0).any(axis=1).sum())
    assert total_rc > 10


def test_host_fuzz_against_reference_binary():
    """tools/fuzz_host.py: 25 random genome sets x ini values x speculation slicings == reference binary (zero-initialising
    allocator, see the test above); includes poly-N windows, on which csgmum's Find_UM runs off the query buffer when a strand
    shares no symbol with the window (the checker skips that call, oracle/ref_backend.cpp).  Where that happened and the binary's
    answer differs (it depends on the bytes behind the buffer and on the length of the file names, or is a crash) the case is
    reported as not comparable, never as equal; at most 2 of the 25 seeds may end there"""
    import subprocess
    import sys
    env = dict(os.environ, MALLOC_PERTURB_="255", GLIBC_TUNABLES="glibc.malloc.tcache_count=0", PB200_HOST_THREADS="4")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_host.py"), "5070", "25"], env=env, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:]
    assert "done 25 cases, 0 mismatches" in r.stdout, r.stdout[-2000:]
    assert int(r.stdout.rsplit("mismatches,", 1)[1].split()[0]) <= 2, r.stdout[-2000:]


@pytest.mark.parametrize("name", ["indep_20k", "rearr_60k", "pop_30k_x12", "windows_50k"])
@pytest.mark.parametrize("task,maxlen", [(1, 4096), (7, 120)])
def test_final_gaps_reproduce_reference(name, task, maxlen, monkeypatch):
    """the parallel replay takes the accept decisions of the engine's discovery as FINAL for the gaps where they cannot depend on
    the order (host/replay.cpp "final gaps"); the discovery here is the CPU emulation of the device pass
    (oracle/discover_emul.cpp: levels, regions of a level in random order, pairs in the reference's order) == golden"""
    from oracle import hosttest
    monkeypatch.setenv("PB200_REPLAY_MODE", "par")
    monkeypatch.setenv("PB200_REPLAY_OWN_THREADS", "1")
    monkeypatch.setenv("PB200_HOST_THREADS", "3")
    monkeypatch.setenv("PB200_REPLAY_TASK", str(task))
    monkeypatch.setenv("PB200_EMUL_MAXLEN", str(maxlen))
    g, kw, gold = golden_case(name)
    nfinal = 0
    for seed in (1, 2, 3):
        monkeypatch.setenv("PB200_EMUL_SEED", str(seed))
        res = hosttest.align(g, api.make_params(**kw), backend=3)
        assert diff_dumps(result_to_dump(res), gold) == []
        nfinal += res["stats"]["replay_final_gaps"]
        monkeypatch.setenv("PB200_NO_DEVICE_FINAL", "1")           # the same discovery with every gap replayed
        res = hosttest.align(g, api.make_params(**kw), backend=3)
        assert diff_dumps(result_to_dump(res), gold) == [] and res["stats"]["replay_final_gaps"] == 0
        monkeypatch.delenv("PB200_NO_DEVICE_FINAL")
    if name == "indep_20k":
        assert nfinal > 0


def test_final_gaps_fuzz():
    """tools/fuzz_replay.py through the emulated discovery (FUZZ_BACKEND=3): sequential loop == parallel replay with final gaps on
    random sets with 8-12 base MUMs (chance reverse-strand candidates: foreign reads and writes that touch final gaps)"""
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_replay.py"), "9100", "12"], env=dict(os.environ, FUZZ_BACKEND="3"),
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:]
    assert "done 12 cases, 0 mismatches" in r.stdout, r.stdout[-2000:]
    assert "'replay_final_gaps': 0.0" not in r.stdout.rsplit("totals", 1)[1]


def test_parallel_literal_sort_equals_std_sort():
    """host/parallel.cpp: literal_std_sort_by_first == std::sort on (key, id) records, position by position, on inputs full of
    ties (the only case where the permutation is not determined by the keys): random, few distinct keys, sorted, reversed,
    ascending runs (the MUM list's shape: anchors, then the recursion's MUMs), constant"""
    from oracle import hosttest
    lib = hosttest.load()
    lib.pbtest_literal_sort_check.argtypes = [C.c_void_p, C.c_int64, C.c_int]
    rng = np.random.default_rng(77)
    for n in (40000, 100003, 262144, 300001):
        shapes = [rng.integers(0, n // 3, n), rng.integers(0, 7, n), np.sort(rng.integers(0, n // 2, n)),
                  np.sort(rng.integers(0, n // 2, n))[::-1].copy(), np.concatenate([np.sort(rng.integers(0, n, n // 4)), np.sort(rng.integers(0, n, n - n // 4))]),
                  np.zeros(n, np.int64), np.arange(n) // 2]
        # + the MUM list's own shape with ONE tied pair at a random place (what the alignments really produce): two ascending
        # runs, a few local descents in the second, distinct keys except one
        for rep in range(6):
            ks = rng.permutation(4 * n)[:n]
            a, b = np.sort(ks[:n // 4]), np.sort(ks[n // 4:])
            for i in range(0, len(b) - 1, 120):
                b[i], b[i + 1] = b[i + 1], b[i]
            keys = np.concatenate([a, b])
            i, j = rng.integers(0, n, 2)
            keys[i] = keys[j]
            shapes.append(keys)
        for keys in shapes:
            keys = np.ascontiguousarray(keys, np.int64)
            for threads in (1, 2, 3, 8):
                assert lib.pbtest_literal_sort_check(keys.ctypes.data, len(keys), threads) == 0


def test_window_order_matches_reference_trace():
    """sequence of (window start, length) searched by the exact replay == the reference's setMums1 call sequence"""
    from oracle import hosttest, runner
    from parsnp_b200 import synth
    g = synth.g_indep(30000, 3, 0.03, 13)
    with tempfile.TemporaryDirectory() as td:
        ref, qs = synth.write_dataset(os.path.join(td, "d"), g)
        r = runner.run_ref(ref, qs, os.path.join(td, "r"), cands=True)
    res = hosttest.align(g, api.make_params(flags=api.FLAG_TRACE_WINDOWS), backend=1)
    want = [(w["ini0"], w["len0"]) for w in r["cands"]]
    assert [tuple(x) for x in res["trace"].tolist()] == want
    assert diff_dumps(result_to_dump(res), r["dump"]) == []


def test_spec_backend_end_to_end_small():
    """host orchestrator + brute-force specification on a tiny set == reference binary"""
    from oracle import hosttest, runner
    from parsnp_b200 import synth
    g = synth.g_indep(3000, 2, 0.04, 3)
    with tempfile.TemporaryDirectory() as td:
        ref, qs = synth.write_dataset(os.path.join(td, "d"), g)
        r = runner.run_ref(ref, qs, os.path.join(td, "r"))
    res = hosttest.align(g, api.make_params(), backend=0)
    assert diff_dumps(result_to_dump(res), r["dump"]) == []
