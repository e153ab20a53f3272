"""parsnp_b200_core <ini>: the process boundary of parsnp_core (argv, ini keys, FASTA ingest, output files, exit codes)."""
import json
import os
import subprocess

import pytest

from tests.conftest import ROOT, GOLDEN, C1A, load_golden
from tests.refcmp import diff_dumps

EXE = os.path.join(ROOT, "parsnp_b200", "bin", "parsnp_b200_core")
REF = os.path.join(ROOT, "oracle", "_ref", "parsnp_core_ref")


def _ini(tmp_path, **kw):
    from oracle import runner
    out = tmp_path / "out"
    out.mkdir(exist_ok=True)
    return runner.write_ini(str(tmp_path / "run.ini"), os.path.join(GOLDEN, "mers", "England1.fna"),
                            [os.path.join(GOLDEN, "mers", q + ".fna") for q in C1A], str(out), **kw), out


@pytest.fixture(scope="module", autouse=True)
def _build():
    import __graft_entry__ as ge
    ge.build()


def test_argv_and_exit_codes(tmp_path):
    r = subprocess.run([EXE, "-v"], stdout=subprocess.PIPE, text=True)
    assert r.returncode == 0 and r.stdout.startswith("Parsnp v1.0.1")
    r = subprocess.run([EXE, "-h"], stdout=subprocess.PIPE, text=True)
    assert r.returncode == 0 and "parameter file" in r.stdout
    r = subprocess.run([EXE], stdout=subprocess.PIPE, text=True)
    assert r.returncode == 1 and "No parameter file" in r.stdout
    bad = tmp_path / "bad.ini"
    bad.write_text("[Reference]\nfile=/nonexistent.fna\nreverse=0\n[Query]\n[LCB]\nd=300\n[Output]\noutdir=%s\n" % tmp_path)
    r = subprocess.run([EXE, str(bad)], stdout=subprocess.PIPE, text=True)
    assert r.returncode == 1 and "Cannot open reference file" in r.stdout


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref not built")
def test_ingest_lines_match_reference(tmp_path):
    """`<file>,Len:<n>,GC:<pct>` lines (src/parsnp.cpp:3154): same ini + FASTA handling as the reference binary"""
    ini, out = _ini(tmp_path)
    mine = subprocess.run([EXE, ini], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=str(tmp_path))
    ref = subprocess.run([REF, ini], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=str(tmp_path),
                         env=dict(os.environ, PARSNP_ORACLE_DUMP=str(tmp_path / "d.txt"), PARSNP_ORACLE_DUMP_EXIT="1"))
    want = [ln for ln in ref.stdout.splitlines() if ",Len:" in ln]
    got = [ln for ln in mine.stdout.splitlines() if ",Len:" in ln]
    assert len(want) == 5 and got == want


@pytest.mark.gpu
def test_binary_mum_lcb_and_log(tmp_path):
    from oracle import runner
    ini, out = _ini(tmp_path)
    r = subprocess.run([EXE, ini], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr
    d = runner.parse_dump(str(out / "parsnpAligner.mums"))
    assert diff_dumps(d, load_golden("c1a")) == []
    log = (out / "parsnpAligner.log").read_text()
    assert "Total coverage among all sequences:" in log and "Number of clusters created:   3" in log
    assert "Number of MUM anchors found:   149" in log


@pytest.mark.gpu
def test_binary_mumi_mode(tmp_path):
    ini, out = _ini(tmp_path, calcmumi=1)
    r = subprocess.run([EXE, ini], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr
    got = [ln.strip().split(":")[1] for ln in (out / "all.mumi").read_text().splitlines()]
    assert got == json.load(open(os.path.join(GOLDEN, "mumi.json")))["c1a"]


def _ref_run(tmp_path, ref, queries, **kw):
    from oracle import runner
    r = runner.run_ref(ref, queries, str(tmp_path / "refrun"), dump_exit=False, **kw)
    return r, os.path.join(str(tmp_path / "refrun"), "dump.txt"), os.path.join(r["outdir"], "parsnpAligner.xmfa"), \
        os.path.join(str(tmp_path / "refrun"), "ref.ini")


def _synthetic(tmp_path, kind):
    import numpy as np
    from parsnp_b200 import synth
    if kind == "rearr":
        rng = np.random.default_rng(5)
        g = synth.g_indep(40000, 3, 0.02, 33)
        g = [g[0]] + [synth.rearrange(x, rng, n_inv=2, inv_len=2500) for x in g[1:]]
        return synth.write_dataset(str(tmp_path / "data"), g)
    g = synth.g_indep(30000, 2, 0.03, 8)
    return synth.write_dataset(str(tmp_path / "data"), g, contigs=3)      # multi-contig: N padding + s<k>:p<pos> headers


XTOOL = os.path.join(ROOT, "oracle", "_ref", "xmfa_from_dump")


@pytest.mark.skipif(not (os.path.exists(REF) and os.path.exists(XTOOL)), reason="oracle/_ref tools not built")
@pytest.mark.parametrize("kind", ["c1a", "rearr", "contigs"])
def test_xmfa_writer_matches_reference_bytes(tmp_path, kind):
    """the product's XMFA writer (csrc/main/xmfa.cpp + libMUSCLE) fed with the reference's own MUM/LCB dump == the reference's
    parsnpAligner.xmfa, byte for byte (reverse-strand LCBs, multi-contig coordinates, MUSCLE-aligned gaps)"""
    if kind == "c1a":
        ref, qs = os.path.join(GOLDEN, "mers", "England1.fna"), [os.path.join(GOLDEN, "mers", q + ".fna") for q in C1A]
    else:
        ref, qs = _synthetic(tmp_path, kind)
    r, dump, xmfa, ini = _ref_run(tmp_path, ref, qs)
    mine = str(tmp_path / "mine.xmfa")
    rc = subprocess.run([XTOOL, ini, dump, mine]).returncode
    assert rc == 0
    assert open(mine, "rb").read() == open(xmfa, "rb").read()
    if kind == "c1a":
        import hashlib
        assert hashlib.md5(open(mine, "rb").read()).hexdigest() == "5b59e50c5b8c1f79165fc41cfd2a6ac4"


@pytest.mark.skipif(not (os.path.exists(REF) and os.path.exists(XTOOL)), reason="oracle/_ref tools not built")
def test_writer_fuzz_against_reference_binary():
    """tools/fuzz_xmfa.py, 12 random cases: XMFA, blocks/ and parsnp.unalign written from the product's own MUMs, LCBs and
    cluster -> MUM lists and log counters (host orchestrator over csgmum) == the reference binary's files, parsnpAligner.log included.  Seed 61174 has LCBs that overlap on
    the reference: the header coordinates then come out of the reference's trim loop reading row 0 behind its new end
    (src/parsnp.cpp:941-949), reproduced in csrc/main/xmfa.cpp"""
    import sys
    env = dict(os.environ, MALLOC_PERTURB_="255", GLIBC_TUNABLES="glibc.malloc.tcache_count=0", PB200_HOST_THREADS="4")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_xmfa.py"), "61170", "12"], env=env, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0 and "done 12 cases, 0 mismatches" in r.stdout, r.stdout[-2000:]


def _log_lines(path, own_paths=False):
    """parsnpAligner.log, comparable: elapsed-time values dropped (the reference's have 1 s resolution, ours are the run's own),
    `Sequence i : <path>` reduced to the file name"""
    out = []
    for ln in open(path).read().splitlines():
        if "elapsed time:" in ln or "running time:" in ln:
            ln = ln.split(":")[0]
        if ln.startswith("Sequence ") and " : " in ln:
            ln = ln.split(" : ")[0] + " : " + os.path.basename(ln.split(" : ")[1])
        out.append(ln)
    return out


def _tree(root):
    """{relative path: bytes} of every file below root"""
    out = {}
    for base, _, files in os.walk(root):
        for f in files:
            p = os.path.join(base, f)
            out[os.path.relpath(p, root)] = open(p, "rb").read()
    return out


@pytest.mark.skipif(not (os.path.exists(REF) and os.path.exists(XTOOL)), reason="oracle/_ref tools not built")
@pytest.mark.parametrize("kind", ["c1a", "rearr"])
def test_recombfilter_blocks_match_reference(tmp_path, kind):
    """ini recombfilter=1: <outdir>/blocks/b<k>/seq.fna (one directory per LCB, src/parsnp.cpp:538-543, 605-644, 958-963) from
    the product's writer == the reference's, file for file"""
    if kind == "c1a":
        ref, qs = os.path.join(GOLDEN, "mers", "England1.fna"), [os.path.join(GOLDEN, "mers", q + ".fna") for q in C1A]
    else:
        ref, qs = _synthetic(tmp_path, kind)
    r, dump, xmfa, ini = _ref_run(tmp_path, ref, qs, recombfilter=1)
    want = _tree(os.path.join(r["outdir"], "blocks"))
    assert len(want) >= 2
    mine = tmp_path / "mine"
    mine.mkdir()
    rc = subprocess.run([XTOOL, ini, dump, str(mine / "parsnpAligner.xmfa"), str(mine)]).returncode
    assert rc == 0
    assert (mine / "parsnpAligner.xmfa").read_bytes() == open(xmfa, "rb").read()
    assert _tree(str(mine / "blocks")) == want
    assert sorted(os.listdir(str(mine / "blocks"))) == sorted(os.listdir(os.path.join(r["outdir"], "blocks")))


@pytest.mark.skipif(not (os.path.exists(REF) and os.path.exists(XTOOL)), reason="oracle/_ref tools not built")
@pytest.mark.parametrize("kind", ["c1a", "rearr", "contigs"])
def test_unaligned_regions_match_reference(tmp_path, kind):
    """ini unaligned=1: parsnp.unalign (Aligner::setUnalignableRegions, src/parsnp.cpp:2310-2382) - the records come from the
    product's host orchestrator (final mumlayout; search by the reference's csgmum on the CPU), the file from its writer"""
    from oracle import hosttest
    from parsnp_b200 import api
    if kind == "c1a":
        ref, qs = os.path.join(GOLDEN, "mers", "England1.fna"), [os.path.join(GOLDEN, "mers", q + ".fna") for q in C1A]
    else:
        ref, qs = _synthetic(tmp_path, kind)
    r, dump, xmfa, ini = _ref_run(tmp_path, ref, qs, unaligned=1)
    want = open(os.path.join(r["outdir"], "parsnp.unalign"), "rb").read()
    assert len(want) > 100
    g = [api.ingest_fasta(ref, True)] + [api.ingest_fasta(q, False) for q in qs]
    res = hosttest.align(g, api.make_params(flags=api.FLAG_UNALIGNED), backend=1)
    rec = tmp_path / "unaligned.txt"
    rec.write_text("".join("%d %d %d\n" % tuple(x) for x in res["unaligned"].tolist()))
    mine = tmp_path / "mine"
    mine.mkdir()
    rc = subprocess.run([XTOOL, ini, dump, str(mine / "parsnpAligner.xmfa"), str(mine), str(rec)]).returncode
    assert rc == 0
    assert (mine / "parsnp.unalign").read_bytes() == want


@pytest.mark.gpu
def test_binary_blocks_and_unaligned_end_to_end(tmp_path):
    """parsnp_b200_core with recombfilter=1 and unaligned=1 on the GPU == parsnp_core: XMFA, blocks/ tree, parsnp.unalign"""
    from oracle import runner
    ref, qs = _synthetic(tmp_path, "rearr")
    r, dump, xmfa, _ = _ref_run(tmp_path, ref, qs, recombfilter=1, unaligned=1)
    out = tmp_path / "mine"
    out.mkdir()
    ini = runner.write_ini(str(tmp_path / "mine.ini"), ref, qs, str(out), cores=4, recombfilter=1, unaligned=1)
    p = subprocess.run([EXE, ini], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=str(tmp_path))
    assert p.returncode == 0, p.stderr
    assert (out / "parsnpAligner.xmfa").read_bytes() == open(xmfa, "rb").read()
    # the statistics log the Python driver parses (parsnp:1530-1536), line for line
    assert _tree(str(out / "blocks")) == _tree(os.path.join(r["outdir"], "blocks"))
    assert (out / "parsnp.unalign").read_bytes() == open(os.path.join(r["outdir"], "parsnp.unalign"), "rb").read()
    assert _log_lines(str(out / "parsnpAligner.log")) == _log_lines(os.path.join(r["outdir"], "parsnpAligner.log"))


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["c1a", "rearr", "contigs"])
def test_binary_xmfa_end_to_end(tmp_path, kind):
    """parsnp_b200_core <ini> on the GPU == parsnp_core on the CPU: identical parsnpAligner.xmfa"""
    from oracle import runner
    if kind == "c1a":
        ref, qs = os.path.join(GOLDEN, "mers", "England1.fna"), [os.path.join(GOLDEN, "mers", q + ".fna") for q in C1A]
    else:
        ref, qs = _synthetic(tmp_path, kind)
    r, dump, xmfa, _ = _ref_run(tmp_path, ref, qs)
    out = tmp_path / "mine"
    out.mkdir()
    ini = runner.write_ini(str(tmp_path / "mine.ini"), ref, qs, str(out), cores=4)
    p = subprocess.run([EXE, ini], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=str(tmp_path))
    assert p.returncode == 0, p.stderr
    assert (out / "parsnpAligner.xmfa").read_bytes() == open(xmfa, "rb").read()
    # the statistics log the Python driver parses (parsnp:1530-1536), line for line
    assert _log_lines(str(out / "parsnpAligner.log")) == _log_lines(os.path.join(r["outdir"], "parsnpAligner.log"))


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref not built")
def test_concurrent_partitions_match_reference(tmp_path):
    """BASELINE configs[4] at reduced size, the way the Python driver runs partition mode (parsnp:1553-1615): one genome pool,
    sorted + random.Random(42).shuffle, consecutive slices, one parsnp_core process per partition, min(threads, partitions) of
    them AT THE SAME TIME with nothing but the ini differing - so the binary has to choose its GPU by itself (advisory lock per
    device, else pid % devices; every process shares GPU 0 on a one-GPU box).  Every partition's parsnpAligner.xmfa must equal the
    reference binary's byte for byte, and the cwd must hold the (empty) allmums.out the reference leaves there."""
    import random
    from oracle import runner
    from parsnp_b200 import synth
    NPART, PER, L = 10, 6, 120_000
    g = synth.g_pop(L, NPART * PER, 0.01, 23)
    ref, qs = synth.write_dataset(str(tmp_path / "pool"), g)
    names = sorted(qs)
    random.Random(42).shuffle(names)
    parts = [names[i * PER:(i + 1) * PER] for i in range(NPART)]
    procs, want = [], []
    lockdir = tmp_path / "locks"
    lockdir.mkdir()
    for i, p in enumerate(parts):
        d = tmp_path / ("part%d" % i)
        (d / "out").mkdir(parents=True)
        ini = runner.write_ini(str(d / "run.ini"), ref, p, str(d / "out"), cores=2)
        procs.append((d, subprocess.Popen([EXE, ini], cwd=str(d), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                                          env=dict(os.environ, PB200_LOCK_DIR=str(lockdir), PB200_HOST_THREADS="2"))))
    for i, p in enumerate(parts):                       # the reference meanwhile, on the CPU
        r = runner.run_ref(ref, p, str(tmp_path / ("ref%d" % i)), dump=False, cores=2)
        assert r["returncode"] == 0
        want.append(open(os.path.join(r["outdir"], "parsnpAligner.xmfa"), "rb").read())
    for i, (d, pr) in enumerate(procs):
        out, err = pr.communicate(timeout=600)
        assert pr.returncode == 0, err[-2000:]
        got = (d / "out" / "parsnpAligner.xmfa").read_bytes()
        assert got == want[i], "partition %d: XMFA differs from the reference" % i
        assert (d / "allmums.out").exists() and (d / "allmums.out").stat().st_size == 0
