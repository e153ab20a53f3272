"""Full-size goldens: the REAL reference binary (oracle/_ref/parsnp_core_ref, built from /root/reference by
oracle/build_ref.py) run once per BASELINE.json config on the CPU; only a digest of its MUM/LCB dump is committed
(tests/golden/full_size.json), because the dumps themselves are 10-100 MB.

    python tests/golden/make_golden_full.py --case c2_indep_5m_8q        # one case (9-15 min of one core)
    python tests/golden/make_golden_full.py --all --jobs 4               # every case, 4 at a time

Digest of a dump (hook H1 of oracle/build_ref.py = this->mums / this->clusters at src/parsnp.cpp:505):
  n_mums, n_clusters, sha256 of the dump text, and the sha256 of every block of 4096 'M' lines (so that a mismatch on the
  GPU says WHERE along the reference it is) and of all 'C' lines.
tests/test_gpu_fullsize.py renders the product's result in the same text format (tests/refcmp.dump_text) and compares
the digests.  The inputs are not committed either: they are the seeded synthetic genomes of SURVEY 8(d)
(parsnp_b200/synth.py), regenerated on the GPU box.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

OUT = os.path.join(ROOT, "tests", "golden", "full_size.json")
BLOCK = 4096


def cases():
    """name -> dict(kind, L, nq, div, seed, contigs, ini overrides). BASELINE.json configs[1..3]."""
    return {
        # configs[1]: the benched workload
        "c2_indep_5m_8q": dict(kind="indep", L=5_000_000, nq=8, div=0.01, seed=1, contigs=1, ini={}),
        # the same shape with seed 3 = the genome set of rank 2 in `bench.py --gpus N` (N >= 3): its final MUM list has two MUMs
        # with the same start[0], i.e. the literal replay of the reference's unstable std::sort decides their order
        "c2_indep_5m_8q_seed3": dict(kind="indep", L=5_000_000, nq=8, div=0.01, seed=3, contigs=1, ini={}),
        # configs[2] shape at 32 and at the full 200 queries
        "c3_pop_5m_32q": dict(kind="pop", L=5_000_000, nq=32, div=0.01, seed=1, contigs=1, ini={}),
        "c3_pop_5m_200q": dict(kind="pop", L=5_000_000, nq=200, div=0.01, seed=1, contigs=1, ini={}),
        # configs[3] shape: 50 Mbp reference in 10 contigs, multi-contig queries (310 N's at every break), 8 queries;
        # default p (4 reference windows) and p = 50 M (one window)
        "c4_pop_50m_8q_p15m": dict(kind="pop", L=50_000_000, nq=8, div=0.01, seed=1, contigs=10, ini=dict(p=15_000_000)),
        "c4_pop_50m_8q_p50m": dict(kind="pop", L=50_000_000, nq=8, div=0.01, seed=1, contigs=10, ini=dict(p=50_000_000)),
    }


def make_genomes(c):
    from parsnp_b200 import synth
    f = synth.g_indep if c["kind"] == "indep" else synth.g_pop
    return f(c["L"], c["nq"], c["div"], c["seed"])


def digest_lines(lines):
    """lines: iterable of dump lines WITHOUT the trailing newline"""
    whole = hashlib.sha256()
    blocks, cur, nm, nc, k = [], hashlib.sha256(), 0, 0, 0
    cl = hashlib.sha256()
    for ln in lines:
        whole.update(ln.encode() + b"\n")
        if ln.startswith("M "):
            cur.update(ln.encode() + b"\n")
            nm += 1
            k += 1
            if k == BLOCK:
                blocks.append(cur.hexdigest()[:16])
                cur, k = hashlib.sha256(), 0
        elif ln.startswith("C "):
            cl.update(ln.encode() + b"\n")
            nc += 1
    if k:
        blocks.append(cur.hexdigest()[:16])
    return dict(n_mums=nm, n_clusters=nc, sha256=whole.hexdigest(), mum_blocks=blocks, clusters_sha256=cl.hexdigest())


def run_case(name, zero_heap):
    from oracle import runner
    from parsnp_b200 import synth
    c = cases()[name]
    g = make_genomes(c)
    with tempfile.TemporaryDirectory(dir=os.environ.get("PB200_GOLDEN_TMP")) as td:
        ref, qs = synth.write_dataset(os.path.join(td, "d"), g, contigs=c["contigs"])
        del g
        t0 = time.time()
        r = runner.run_ref(ref, qs, os.path.join(td, "r"), dump=True, dump_exit=True, zero_heap=zero_heap, **c["ini"])
        assert r["returncode"] == 0, r["stderr"][-2000:]
        with open(os.path.join(td, "r", "dump.txt")) as f:
            d = digest_lines(x.rstrip("\n") for x in f)
        d.update(case=c, reference_mumlcb_seconds=r["mumlcb_seconds"], reference_wall_seconds=round(time.time() - t0, 1),
                 zero_heap=bool(zero_heap), host_cores_used=1)
    return d


def merge(name, d):
    import fcntl
    with open(OUT + ".lock", "w") as lk:                      # cases run concurrently
        fcntl.flock(lk, fcntl.LOCK_EX)
        cur = json.load(open(OUT)) if os.path.exists(OUT) else {}
        cur[name] = d
        tmp = OUT + ".tmp%d" % os.getpid()
        with open(tmp, "w") as f:
            json.dump(cur, f, indent=1, sort_keys=True)
        os.replace(tmp, OUT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", action="append")
    ap.add_argument("--all", action="store_true")
    ap.add_argument("--jobs", type=int, default=1)
    ap.add_argument("--stock-heap", action="store_true", help="run the binary on the stock allocator (default: zero-filled heap, "
                    "see oracle/runner.py:run_ref)")
    a = ap.parse_args()
    names = list(cases()) if a.all else (a.case or [])
    if a.jobs > 1 and len(names) > 1:
        procs = []
        pending = list(names)
        while pending or procs:
            while pending and len(procs) < a.jobs:
                n = pending.pop(0)
                cmd = [sys.executable, os.path.abspath(__file__), "--case", n] + (["--stock-heap"] if a.stock_heap else [])
                procs.append((n, subprocess.Popen(cmd)))
            time.sleep(5)
            for n, p in list(procs):
                if p.poll() is not None:
                    procs.remove((n, p))
                    print(n, "rc", p.returncode, flush=True)
        return
    for n in names:
        d = run_case(n, not a.stock_heap)
        key = n + ("_stock_heap" if a.stock_heap else "")
        merge(key, d)
        print(key, {k: d[k] for k in ("n_mums", "n_clusters", "sha256", "reference_mumlcb_seconds")}, flush=True)


if __name__ == "__main__":
    main()
