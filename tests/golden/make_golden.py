"""Regenerates the golden MUM/LCB dumps in this directory by running the REAL reference binary
(oracle/_ref/parsnp_core_ref, built from /root/reference by oracle/build_ref.py) on the committed inputs.
Run from the repo root:  python tests/golden/make_golden.py
Cases: C1a, C1b (MERS, template-default ini), C1c (all 45 others), plus small synthetics that
exercise recursion, inversions/indels and multiple reference windows."""
import hashlib
import json
import os
import shutil
import sys
import tarfile
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import runner            # noqa: E402
from parsnp_b200 import synth        # noqa: E402

G = os.path.join(ROOT, "tests", "golden")
C1A = ["Al-Hasa_1_2013", "Bisha_1_2012", "EMC_2012", "Jordan-N3_2012"]
C1B = ["Al-Hasa_12_2013", "Al-Hasa_15_2013", "Al-Hasa_16_2013", "Al-Hasa_17_2013"]


def synth_cases():
    """name -> (genomes, ini overrides)"""
    cases = {}
    cases["indep_20k"] = (synth.g_indep(20000, 3, 0.03, 5), {})
    rng = np.random.default_rng(77)
    g = synth.g_indep(60000, 3, 0.02, 21)
    g = [g[0]] + [synth.rearrange(x, rng, n_inv=2, inv_len=3000) for x in g[1:]]
    cases["rearr_60k"] = (g, {})
    cases["windows_50k"] = (synth.g_indep(50000, 2, 0.02, 9), dict(p=20000))
    cases["pop_30k_x12"] = (synth.g_pop(30000, 12, 0.01, 3), {})
    return cases


def dump_to_json(d):
    return dict(n=d["n"], mums=[[m[0], m[1], [list(x) for x in m[2]]] for m in d["mums"]],
                clusters=[[c[0], c[1], c[2], [list(x) for x in c[3]]] for c in d["clusters"]])


def main():
    summary = {}
    with tempfile.TemporaryDirectory() as td:
        for name, qs in (("c1a", C1A), ("c1b", C1B)):
            r = runner.run_ref(os.path.join(G, "mers", "England1.fna"), [os.path.join(G, "mers", q + ".fna") for q in qs],
                               os.path.join(td, name), dump_exit=False)
            md5 = hashlib.md5(open(os.path.join(r["outdir"], "parsnpAligner.xmfa"), "rb").read()).hexdigest()
            json.dump(dump_to_json(r["dump"]), open(os.path.join(G, name + ".json"), "w"))
            summary[name] = dict(mums=len(r["dump"]["mums"]), clusters=len(r["dump"]["clusters"]), xmfa_md5=md5)
        ex = os.path.join(td, "all")
        os.makedirs(ex)
        tarfile.open(os.path.join(G, "mers_all.tar.gz")).extractall(ex)
        others = sorted(x for x in os.listdir(ex) if x != "England1.fna")
        r = runner.run_ref(os.path.join(G, "mers", "England1.fna"), [os.path.join(ex, q) for q in others], os.path.join(td, "c1c"),
                           dump_exit=False)
        md5 = hashlib.md5(open(os.path.join(r["outdir"], "parsnpAligner.xmfa"), "rb").read()).hexdigest()
        json.dump(dump_to_json(r["dump"]), open(os.path.join(G, "c1c.json"), "w"))
        summary["c1c"] = dict(mums=len(r["dump"]["mums"]), clusters=len(r["dump"]["clusters"]), xmfa_md5=md5, order=others)
        for name, (g, kw) in synth_cases().items():
            ref, qs = synth.write_dataset(os.path.join(td, name + "_d"), g)
            r = runner.run_ref(ref, qs, os.path.join(td, name), cands=True, **kw)
            json.dump(dump_to_json(r["dump"]), open(os.path.join(G, name + ".json"), "w"))
            summary[name] = dict(mums=len(r["dump"]["mums"]), clusters=len(r["dump"]["clusters"]), windows=len(r["cands"]),
                                 rev=sum(1 for m in r["dump"]["mums"] if any(not x[2] for x in m[2])))
    # MUMi mode (calcmumi=1)
    mumi = {}
    with tempfile.TemporaryDirectory() as td:
        mumi["c1a"] = runner.run_ref_mumi(os.path.join(G, "mers", "England1.fna"), [os.path.join(G, "mers", q + ".fna") for q in C1A],
                                          os.path.join(td, "m1"))
        for name in ("rearr_60k", "pop_30k_x12", "indep_20k"):
            g, kw = synth_cases()[name]
            ref, qs = synth.write_dataset(os.path.join(td, name + "_d"), g)
            mumi[name] = runner.run_ref_mumi(ref, qs, os.path.join(td, name))
        # length-ratio rule (src/parsnp.cpp:2074): a query much shorter than the reference gets distance 1
        g = synth.g_indep(40000, 2, 0.02, 17)
        g[2] = g[2][:20000].copy()
        ref, qs = synth.write_dataset(os.path.join(td, "ratio_d"), g)
        mumi["ratio_40k"] = runner.run_ref_mumi(ref, qs, os.path.join(td, "ratio"))
    # the binary's MUMi depends on uninitialised memory when reverse-strand matches win (see oracle/runner.py:mumi_zero_init):
    # keep its raw output for the record, test against the zero-initialised csgmum emulation
    mumi["rearr_60k_binary_uninitialised"] = mumi["rearr_60k"]
    mumi["rearr_60k"] = runner.mumi_zero_init(synth_cases()["rearr_60k"][0])
    for name in ("pop_30k_x12", "indep_20k"):
        assert mumi[name] == runner.mumi_zero_init(synth_cases()[name][0]), name      # no inversions: both agree
    json.dump(mumi, open(os.path.join(G, "mumi.json"), "w"), indent=1)
    summary["mumi"] = {k: len(v) for k, v in mumi.items()}
    json.dump(summary, open(os.path.join(G, "summary.json"), "w"), indent=1)
    print(json.dumps(summary, indent=1)[:2000])


if __name__ == "__main__":
    main()
