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
