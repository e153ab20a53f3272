"""CPU tests: the product library builds for sm_100a, loads, exports every symbol of include/parsnp_b200.h, and
refuses to compute without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from tests.conftest import ROOT
from parsnp_b200 import api


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    return api.load()


def header_symbols():
    h = open(os.path.join(ROOT, "include", "parsnp_b200.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    return sorted(set(re.findall(r"\b(pb200_[a-z0-9_]+)\s*\(", h)))


def test_exports_every_declared_symbol(lib):
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), s


def test_minsize_through_product_abi(lib):
    assert api.minsize("1.1*(Log(S))", 5000000) == 25
    assert api.minsize("1.1*(Log(S))", 31) == 6


def test_params_default(lib):
    p = api.CParams()
    lib.pb200_params_default(C.byref(p))
    assert (p.c, p.d, p.q, p.p, p.filter) == (21, 300, 30, 15000000, 1)
    assert abs(p.diagdiff - 0.12) < 1e-7 and p.anchors == b"1.1*(Log(S))"


def test_no_cpu_fallback(lib):
    if api.cuda_available():
        pytest.skip("GPU present")
    g = [np.frombuffer(b"ACGT" * 20, np.uint8)] * 2
    with pytest.raises(api.Pb200Error) as e:
        api.align(g)
    assert "CUDA" in str(e.value)


def test_ingest_rules(tmp_path):
    p = tmp_path / "x.fna"
    p.write_bytes(b">h1 desc\nACGTacgtNnRYKM-\nUu*12\n>contig2\nAAxxCC\r\n")
    ref = api.ingest_fasta(str(p), True, d=300)
    assert ref.tobytes() == b"ACGTACGTNNNNNNNTTAANNCC"
    q = api.ingest_fasta(str(p), False, d=300)
    assert q.tobytes() == b"ACGTACGTNNNNNNNTT" + b"N" * 310 + b"AANNCC"
