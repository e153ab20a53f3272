"""CUDA search against the reference's csgmum on random cases (runs last: the fixed parity cases come first).

The host orchestrator is the same on both sides (parsnp_b200/csrc/host); only the search differs - the CUDA engine behind
pb200_align_resident versus the unmodified csg.c + mum.c behind oracle/ref_backend.cpp - so any difference is a kernel
difference.  The csgmum side was itself compared with the reference binary on the same generator (tools/fuzz_host.py, thousands
of cases on the CPU).  Where csgmum's Find_UM would have read past a query buffer the checker skips the call, which is the
product's semantics (no seed), so those cases stay comparable here."""
import os
import tempfile

import pytest

from tests.conftest import ROOT
from tests.refcmp import result_to_dump, diff_dumps

have_ref = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libpb200_hosttest.so"))


@pytest.mark.gpu
@pytest.mark.skipif(not have_ref, reason="oracle/_ref not built")
@pytest.mark.parametrize("seed0", [80000, 80012, 80024, 80036])
def test_cuda_search_fuzz_against_csgmum(seed0):
    from oracle import hosttest
    from parsnp_b200 import api, synth
    from tools.fuzz_cases import make_case, params_kw
    bad = []
    for seed in range(seed0, seed0 + 12):
        g, contigs, kw, desc, _ = make_case(seed)
        with tempfile.TemporaryDirectory() as td:
            rf, qf = synth.write_dataset(os.path.join(td, "d"), g, contigs=contigs)
            gi = [api.ingest_fasta(rf, True, d=kw.get("d", 300))] + [api.ingest_fasta(x, False, d=kw.get("d", 300)) for x in qf]
        want = hosttest.align(gi, api.make_params(**params_kw(kw)), backend=1)
        got = api.align(gi, api.make_params(**params_kw(kw)))
        d = diff_dumps(result_to_dump(got), result_to_dump(want))
        if d or got["no_mums"] != want["no_mums"]:
            bad.append((seed, desc, kw, d[:2]))
    assert bad == []


@pytest.mark.gpu
@pytest.mark.skipif(not have_ref, reason="oracle/_ref not built")
@pytest.mark.parametrize("seed0", [91000, 91008])
def test_cuda_final_gaps_fuzz(seed0, monkeypatch):
    """the device's own accept decisions taken as final by the parallel replay (host/replay.cpp "final gaps") on random sets with
    8-12 base MUMs - chance reverse-strand candidates inside sub-regions, i.e. gaps that must NOT be taken as final, foreign
    writes that touch final gaps, restarts - against the sequential loop over csgmum (tools/fuzz_replay.py's generator; the CPU
    suite runs the same host code over the emulated discovery, oracle/discover_emul.cpp)"""
    import numpy as np
    from oracle import hosttest
    from parsnp_b200 import api, synth
    nfinal, bad = 0, []
    for seed in range(seed0, seed0 + 8):
        rng = np.random.default_rng(seed)
        L = int(rng.choice([30000, 80000, 200000]))
        nq = int(rng.integers(1, 7))
        div = float(rng.choice([0.01, 0.03, 0.05]))
        g = (synth.g_indep if rng.random() < 0.6 else synth.g_pop)(L, nq, div, int(rng.integers(1, 10**6)))
        if rng.random() < 0.3:
            a = int(rng.integers(0, L - 400)); ln = int(rng.integers(30, 300)); a = min(a, L - 2 * ln)
            for x in g:
                x[a + ln:a + 2 * ln] = synth.revcomp(x[a:a + ln])
        kw = dict(mums=str(rng.choice(["8", "10", "12", "1.1*(Log(S))", "0.7*(Log(S))"])), q=int(rng.choice([10, 30])))
        monkeypatch.setenv("PB200_REPLAY_MODE", "seq")
        want = result_to_dump(hosttest.align(g, api.make_params(**kw), backend=1))
        for rep in range(2):
            monkeypatch.setenv("PB200_REPLAY_MODE", "par")
            monkeypatch.setenv("PB200_REPLAY_TASK", str(int(rng.choice([1, 2, 5, 40]))))
            monkeypatch.setenv("PB200_REPLAY_THREADS", str(int(rng.choice([2, 4, 8]))))
            got = api.align(g, api.make_params(**kw))
            nfinal += got["stats"]["replay_final_gaps"]
            d = diff_dumps(result_to_dump(got), want)
            if d:
                bad.append((seed, rep, kw, d[:2]))
    assert bad == []
    assert nfinal > 100


@pytest.mark.gpu
@pytest.mark.skipif(not have_ref, reason="oracle/_ref not built")
@pytest.mark.parametrize("force_big", [False, True])
def test_cuda_window_fuzz_wide_regimes(force_big):
    """single windows far from the comfortable regime - homopolymers and two-letter alphabets (everything repeats), N runs in
    both sequences, 30 to 1 200 bases, up to 6 queries, minsize 2 to 13 (below 4 the engine must leave the shared-memory
    path) - through the shared-memory path and forced through the suffix-array path == real csgmum (oracle/ref_backend.cpp);
    tools/fuzz_spec.py runs the same generator for the CPU specification (23 000 windows, no difference)"""
    import numpy as np
    from oracle import hosttest
    from parsnp_b200 import api
    from tests.conftest import random_case, whole_window_task
    rng = np.random.default_rng(4242 + force_big)
    bad, total = [], 0
    if force_big:
        os.environ["PB200_FORCE_PATH"] = "big"
    try:
        for it in range(160):
            alphabet = [b"AT", b"ACGT", b"ACGT", b"AACGGT", b"A", b"AC"][int(rng.integers(0, 6))]
            with_n = bool(rng.random() < 0.5)
            g = random_case(rng, 30, int(rng.choice([60, 160, 400, 1200])), 6, alphabet, with_n)
            minsize = int(rng.integers(2, 14))
            w, coords = whole_window_task(g, minsize)
            want = hosttest.search_windows(g, w, coords, backend=1)[0]
            G = api.Genomes(g)
            got = G.search_windows(w, coords)[0]
            G.close()
            if not all(np.array_equal(x, y) for x, y in zip(got, want)):
                bad.append((it, alphabet, with_n, minsize, [len(x) for x in g]))
            total += len(want[0])
    finally:
        os.environ.pop("PB200_FORCE_PATH", None)
    assert bad == [] and total > 100


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "parsnp_core_ref")), reason="oracle/_ref not built")
def test_cuda_mumi_fuzz_against_reference_binary():
    """calcmumi=1 on 12 random cases (rearranged queries, contigs, several reference windows): pb200_mumi == the distances the
    reference binary, run here on the zero-filled heap, prints to all.mumi - to the digit"""
    from oracle import runner
    from parsnp_b200 import api, synth
    from tools.fuzz_cases import make_case, params_kw
    bad = []
    for seed in range(81000, 81012):
        g, contigs, kw, desc, _ = make_case(seed)
        with tempfile.TemporaryDirectory() as td:
            rf, qf = synth.write_dataset(os.path.join(td, "d"), g, contigs=contigs)
            want = runner.run_ref_mumi(rf, qf, os.path.join(td, "r"), **kw)
            gi = [api.ingest_fasta(rf, True, d=kw.get("d", 300))] + [api.ingest_fasta(x, False, d=kw.get("d", 300)) for x in qf]
        G = api.Genomes(gi)
        got = ["%f" % v for v in G.mumi(api.make_params(**params_kw(kw)))]
        G.close()
        if got != want:
            bad.append((seed, desc, kw, got, want))
    assert bad == []
