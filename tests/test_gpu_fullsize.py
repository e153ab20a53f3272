"""Parity at BASELINE.json's stated sizes: the CUDA path against the REAL reference binary, element by element.

The reference needs 3-20 minutes of one CPU core per config, so it ran once (tests/golden/make_golden_full.py, here, from
/root/reference via oracle/_ref/parsnp_core_ref) and tests/golden/full_size.json keeps a digest of its MUM/LCB dump: counts,
sha256 of the whole dump, sha256 of every block of 4096 MUM lines, sha256 of the LCB lines.  The test renders the product's
result in the same text format and compares all of them - equality of the sha256 is element-wise equality of every MUM
coordinate, strand flag and LCB boundary (src/parsnp.cpp:505 state).  Inputs are regenerated from their seeds."""
import json
import os

import numpy as np
import pytest

from tests.conftest import GOLDEN
from tests.refcmp import dump_lines

pytestmark = pytest.mark.gpu

FULL = os.path.join(GOLDEN, "full_size.json")
_gold = json.load(open(FULL)) if os.path.exists(FULL) else {}


def _genomes(c):
    from parsnp_b200 import synth
    f = synth.g_indep if c["kind"] == "indep" else synth.g_pop
    g = f(c["L"], c["nq"], c["div"], c["seed"])
    if c["contigs"] > 1:
        # multi-contig FASTA as parsnp_core ingests it: the reference genome is concatenated seamlessly, every query gets
        # d+10 N's at each contig break (src/parsnp.cpp:3114-3118) - tests/test_gpu_engine.py::test_c4_shape_* checks that
        # this equals api.ingest_fasta of the written files
        d = int(c["ini"].get("d", 300))
        g = [g[0]] + [synth.with_contig_padding(x, c["contigs"], d) for x in g[1:]]
    return g


@pytest.mark.parametrize("name", sorted(k for k in _gold if not k.endswith("_stock_heap")))
def test_full_size_matches_reference_golden(name):
    from parsnp_b200 import api
    from tests.golden.make_golden_full import digest_lines
    gold = _gold[name]
    c = gold["case"]
    g = _genomes(c)
    res = api.align(g, api.make_params(**c["ini"]))
    del g
    got = digest_lines(dump_lines(res))
    assert (got["n_mums"], got["n_clusters"]) == (gold["n_mums"], gold["n_clusters"])
    bad = [i for i, (a, b) in enumerate(zip(got["mum_blocks"], gold["mum_blocks"])) if a != b]
    assert not bad, "MUM blocks (of 4096) that differ from the reference: %s" % bad[:10]
    assert got["clusters_sha256"] == gold["clusters_sha256"]
    assert got["sha256"] == gold["sha256"]


def test_stock_heap_run_agrees():
    """the goldens come from the reference on a zero-filled heap (oracle/runner.py:run_ref - MasterRC[].UP is never
    initialised, src/parsnp.cpp:1591-1597); where the same config was also run on the stock allocator, both digests agree"""
    pairs = [(k, k[:-len("_stock_heap")]) for k in _gold if k.endswith("_stock_heap")]
    for a, b in pairs:
        if b in _gold:
            assert _gold[a]["sha256"] == _gold[b]["sha256"], (a, b)
