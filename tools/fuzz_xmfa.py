#!/usr/bin/env python3
"""CPU fuzz of the output writers against the reference binary (test infrastructure, no GPU).

  GLIBC_TUNABLES=glibc.malloc.tcache_count=0 MALLOC_PERTURB_=255 python tools/fuzz_xmfa.py <first seed> <cases>

Every case (tools/fuzz_cases.py) runs oracle/_ref/parsnp_core_ref to the end (libMUSCLE, XMFA, recombfilter blocks, parsnp.unalign)
and the product's writer (parsnp_b200/csrc/main/xmfa.cpp through oracle/_ref/xmfa_from_dump) twice: on the reference's own
MUM/LCB dump, and on the result of the product's host orchestrator (csgmum as search, oracle/hosttest.py: MUMs, LCBs, cluster ->
MUM lists, unaligned-region records, log counters) - everything of the product between the search kernels and the files, parsnpAligner.log
included.  Every output file is
compared byte for byte."""
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from parsnp_b200 import api, synth
from oracle import runner, hosttest
from tools.fuzz_cases import make_case, params_kw
from tests.refcmp import write_dump

XTOOL = os.environ.get("PB200_XTOOL") or os.path.join(os.path.dirname(runner.EXE), "xmfa_from_dump")


def tree(root):
    out = {}
    for base, _, files in os.walk(root):
        for f in files:
            p = os.path.join(base, f)
            out[os.path.relpath(p, root)] = open(p, "rb").read()
    return out


def messy_fasta(path, rng, allow_u):
    """rewrites a FASTA file the way real ones look: lower-case lines, IUPAC codes, '-', U, blanks, digits, '*', blank lines,
    CRLF - everything the character loop of src/parsnp.cpp:2999-3133 folds, maps to N or skips"""
    codes = np.frombuffer(b"RYKMSWBDHVNXrykn-" + (b"Uu" if allow_u else b""), np.uint8)
    out = []
    for ln in open(path, "rb").read().split(b"\n"):
        if ln.startswith(b">") or not ln:
            out.append(ln)
            continue
        a = np.frombuffer(ln, np.uint8).copy()
        if rng.random() < 0.3:
            a |= 0x20
        hit = rng.random(len(a)) < 0.003
        a[hit] = codes[rng.integers(0, len(codes), int(hit.sum()))]
        b = a.tobytes()
        if rng.random() < 0.05:
            k = int(rng.integers(0, len(b) + 1))
            b = b[:k] + [b" ", b"\t", b"12", b"*", b"zq", b"."][int(rng.integers(0, 6))] + b[k:]
        out.append(b)
        if rng.random() < 0.01:
            out.append(b"")
    with open(path, "wb") as f:
        f.write((b"\r\n" if rng.random() < 0.2 else b"\n").join(out))


def log_lines(path):
    """parsnpAligner.log without the values of the elapsed-time lines (the reference's have 1 s resolution)"""
    return [ln.split(":")[0] if ("elapsed time:" in ln or "running time:" in ln) else ln for ln in open(path).read().splitlines()]


seed0 = int(sys.argv[1]); ncases = int(sys.argv[2])
bad = skipped = 0
t0 = time.time()
for it in range(ncases):
    g, contigs, kw, (L, nq, div), rng = make_case(seed0 + it)
    if L > 50000:
        g = [x[:50000] for x in g]              # bounds the libMUSCLE time of a case
    kw["recombfilter"] = int(rng.random() < 0.4)
    kw["unaligned"] = int(rng.random() < 0.5)
    if rng.random() < 0.08:
        kw["doalign"] = 0                        # ini [LCB] doalign=0: header and statistics only, no LCB records
    with tempfile.TemporaryDirectory() as td:
        rf, qf = synth.write_dataset(os.path.join(td, "d"), g, contigs=contigs)
        rev = [0] * len(g)
        if rng.random() < 0.3:                   # real-life FASTA text and ini reverse flags (whole genome reverse-complemented at ingest)
            rev = [int(rng.random() < 0.3) for _ in g]
            for k, path in enumerate([rf] + qf):
                messy_fasta(path, rng, allow_u=not rev[k])       # (a U stays T under reverse=1, src/parsnp.cpp:3090: not modelled by the numpy ingest)
        r = runner.run_ref(rf, qf, os.path.join(td, "r"), dump_exit=False, reverse=rev, **kw)
        if r["dump"] is None or r["returncode"] != 0:
            skipped += 1                         # no MUMs, or one of the reference's own crashes
            continue
        hosttest.runoff_skips()
        prm = params_kw(kw)
        gi = [api.ingest_fasta(rf, True, d=kw.get("d", 300))] + [api.ingest_fasta(x, False, d=kw.get("d", 300)) for x in qf]
        gi = [synth.revcomp(x) if rv else x for x, rv in zip(gi, rev)]
        res = hosttest.align(gi, api.make_params(flags=api.FLAG_UNALIGNED if kw["unaligned"] else 0, **prm), backend=1)
        runoff = hosttest.runoff_skips()
        rec = os.path.join(td, "unaligned.txt")
        with open(rec, "w") as f:
            f.write("".join("%d %d %d\n" % tuple(x) for x in res["unaligned"].tolist()))
        write_dump(res, os.path.join(td, "mine.dump"))
        want = {k: v for k, v in tree(r["outdir"]).items() if k.endswith(".xmfa") or k.startswith("blocks") or k.endswith(".unalign")}
        rc, diff = 0, []
        # writer on the reference's MUM/LCB dump, then the whole CPU chain: product's own MUMs, LCBs and cluster -> MUM lists
        for tag, dump in (("refdump", os.path.join(td, "r", "dump.txt")), ("product", os.path.join(td, "mine.dump"))):
            out = os.path.join(td, tag)
            os.makedirs(out)
            args = [XTOOL, os.path.join(td, "r", "ref.ini"), dump, os.path.join(out, "parsnpAligner.xmfa"), out,
                    rec if kw["unaligned"] else "-"]
            if tag == "product":                 # parsnpAligner.log from the product's counters
                st = res["stats"]
                with open(os.path.join(td, "logstats.txt"), "w") as f:
                    f.write("%d %d %d\n" % (st["anchors"], st["mums_filtered"], st["clusters_filtered"]))
                args.append(os.path.join(td, "logstats.txt"))
            leg = subprocess.run(args, stdout=subprocess.PIPE, stderr=subprocess.STDOUT).returncode
            if leg == 6 and tag == "refdump":
                continue                         # overlapping LCBs: the hook dump alone does not say which MUMs an LCB owns
            rc |= leg
            got = {k: v for k, v in tree(out).items() if k in want or k.startswith("blocks")}
            diff += sorted(tag + ":" + k for k in set(want) | set(got) if want.get(k) != got.get(k))
            if tag == "product" and log_lines(os.path.join(out, "parsnpAligner.log")) != log_lines(os.path.join(r["outdir"], "parsnpAligner.log")):
                diff.append("product:parsnpAligner.log")
        if runoff and diff and all(x.startswith("product:") or x.endswith(".unalign") for x in diff):
            skipped += 1                         # the binary's own MUM list is heap dependent there (tools/fuzz_host.py)
            continue
    if rc != 0 or diff:
        bad += 1
        print("MISMATCH seed", seed0 + it, L, nq, div, contigs, kw, "rc", rc, "files", diff[:5], flush=True)
    elif it % 10 == 0:
        print("ok", seed0 + it, L, nq, contigs, kw, "files", len(want), "%.0fs" % (time.time() - t0), flush=True)
print("done", ncases, "cases,", bad, "mismatches,", skipped, "skipped")
