"""Drive oracle/_ref/parsnp_core_ref (the REAL reference binary, built by oracle/build_ref.py).

TEST INFRASTRUCTURE ONLY - see oracle/build_ref.py header.  Writes the minimal ini of
SURVEY.md App. D (mirrors template.ini with the driver's defaults) and parses the hook dumps.
"""
import os
import re
import subprocess
import time

HERE = os.path.dirname(os.path.abspath(__file__))
EXE = os.path.join(HERE, "_ref", "parsnp_core_ref")

DEFAULTS = dict(anchors="1.1*(Log(S))", mums="1.1*(Log(S))", filter=1, factor="2.0", extendmums=0,
                anchorsonly=0, calcmumi=0, recombfilter=0, cores=8, diagdiff="0.12", doalign=2,
                c=21, d=300, q=30, p=15000000, unaligned=0)


def write_ini(path, ref, queries, outdir, **kw):
    o = dict(DEFAULTS)
    o.update(kw)
    rev = o.pop("reverse", None) or [0] * (len(queries) + 1)      # per genome (reference first): ini reverse / reverse<i>
    L = ["[Reference]", "file=%s" % ref, "reverse=%d" % rev[0], "[Query]"]
    for i, q in enumerate(queries):
        L += ["file%d=%s" % (i + 1, q), "reverse%d=%d" % (i + 1, rev[i + 1])]
    L += ["[MUM]", "anchors=%s" % o["anchors"], "anchorfile=", "anchorsonly=%d" % o["anchorsonly"],
          "calcmumi=%d" % o["calcmumi"], "mums=%s" % o["mums"], "mumfile=", "filter=%d" % o["filter"],
          "factor=%s" % o["factor"], "extendmums=%d" % o["extendmums"],
          "[LCB]", "recombfilter=%d" % o["recombfilter"], "cores=%d" % o["cores"], "diagdiff=%s" % o["diagdiff"],
          "doalign=%d" % o["doalign"], "c=%d" % o["c"], "d=%d" % o["d"], "q=%d" % o["q"], "p=%d" % o["p"],
          "icr=0", "unaligned=%d" % o["unaligned"], "[Output]", "outdir=%s" % outdir, "prefix=parsnp", "showbps=1"]
    with open(path, "w") as f:
        f.write("\n".join(L) + "\n")
    return path


def parse_dump(path):
    """-> dict(n, mums=[(length, slength, [(start,end,fwd)...])], clusters=[(type, nm, length, [(start,end)...])])"""
    mums, clusters, n = [], [], 0
    with open(path) as f:
        for line in f:
            t = line.split()
            if not t:
                continue
            if t[0] == "N":
                n = int(t[1])
            elif t[0] == "M":
                mums.append((int(t[1]), int(t[2]), [tuple(int(x) for x in s.split(":")) for s in t[3:]]))
            elif t[0] == "C":
                clusters.append((int(t[1]), int(t[2]), int(t[3]), [tuple(int(x) for x in s.split(":")) for s in t[4:]]))
    return dict(n=n, mums=mums, clusters=clusters)


def parse_cands(path):
    """-> list of windows: dict(anchors, minsize, ini0, len0, region=[(start,len)..], cands=[(LON, [(dsp,fwd)..])])"""
    wins = []
    with open(path) as f:
        for line in f:
            t = line.split()
            if not t:
                continue
            if t[0] == "W":
                wins.append(dict(anchors=int(t[1]), minsize=int(t[2]), ini0=int(t[3]), len0=int(t[4]),
                                 region=[tuple(int(x) for x in s.split(":")) for s in t[6:]], cands=[]))
            elif t[0] == "K":
                wins[-1]["cands"].append((int(t[1]), [tuple(int(x) for x in s.split(":")) for s in t[2:]]))
    return wins


def run_ref(ref, queries, workdir, dump=True, cands=False, dump_exit=True, timeout=None, zero_heap=True, **kw):
    """run the reference binary; returns dict(dump=..., cands=..., mumlcb_seconds=..., wall=..., stdout=...).
    zero_heap (default): glibc hands out zero-filled memory (MALLOC_PERTURB_=255, tcache off) - the binary never initialises
    MasterRC[].UP (src/parsnp.cpp:1591-1597), so on small windows its answer otherwise depends on stale heap contents; with the
    fill it is the function of its inputs the product implements.  Timed runs (bench.py) switch it off."""
    if not os.path.exists(EXE):
        raise RuntimeError("oracle/_ref/parsnp_core_ref missing: run python oracle/build_ref.py")
    os.makedirs(workdir, exist_ok=True)
    outdir = os.path.join(workdir, "out")
    os.makedirs(outdir, exist_ok=True)
    ini = write_ini(os.path.join(workdir, "ref.ini"), ref, queries, outdir, **kw)
    env = dict(os.environ)
    if zero_heap:
        env["MALLOC_PERTURB_"] = "255"
        env["GLIBC_TUNABLES"] = "glibc.malloc.tcache_count=0"
    res = {}
    if dump:
        env["PARSNP_ORACLE_DUMP"] = os.path.join(workdir, "dump.txt")
        if dump_exit:
            env["PARSNP_ORACLE_DUMP_EXIT"] = "1"
        if os.path.exists(env["PARSNP_ORACLE_DUMP"]):
            os.remove(env["PARSNP_ORACLE_DUMP"])
    if cands:
        env["PARSNP_ORACLE_CANDS"] = os.path.join(workdir, "cands.txt")
        if os.path.exists(env["PARSNP_ORACLE_CANDS"]):
            os.remove(env["PARSNP_ORACLE_CANDS"])
    t0 = time.time()
    r = subprocess.run([EXE, ini], cwd=workdir, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                       text=True, timeout=timeout)
    res["wall"] = time.time() - t0
    res["returncode"] = r.returncode
    res["stdout"] = r.stdout
    res["stderr"] = r.stderr
    m = re.search(r"ORACLE_MUMLCB_SECONDS=([0-9.]+)", r.stderr)
    res["mumlcb_seconds"] = float(m.group(1)) if m else None
    res["outdir"] = outdir
    if dump and os.path.exists(env["PARSNP_ORACLE_DUMP"]):
        res["dump"] = parse_dump(env["PARSNP_ORACLE_DUMP"])
    else:
        res["dump"] = None
    if cands and os.path.exists(env.get("PARSNP_ORACLE_CANDS", "")):
        res["cands"] = parse_cands(env["PARSNP_ORACLE_CANDS"])
    return res


def run_ref_mumi(ref, queries, workdir, **kw):
    """reference in calcmumi=1 mode (Aligner::setMumi) -> list of '%f' strings in query order (cores=1: deterministic order)"""
    r = run_ref(ref, queries, workdir, dump=False, calcmumi=1, cores=1, **kw)
    vals = {}
    with open(os.path.join(r["outdir"], "all.mumi")) as f:
        for line in f:
            if ":" in line:
                i, v = line.strip().split(":")
                vals[int(i)] = v
    return [vals[i + 1] for i in range(len(queries))]


def mumi_zero_init(genomes, p=15000000):
    """Aligner::setMumi (src/parsnp.cpp:1869-2115) re-run through the REAL csgmum (oracle/_ref/libcsgmum_ref.so) with the
    emission loop restated literally and MasterRC[].UP zero-initialised.  The binary leaves MasterRC[].UP uninitialised
    (src/parsnp.cpp:1996-2001, App. B #1); from the second query on it holds stale heap data, which changes the printed
    value whenever a reverse-strand match wins - so for inputs with inversions the binary's all.mumi is not a function of
    its inputs, and this well-defined variant is the oracle.  -> list of '%f' strings."""
    import ctypes as C
    import numpy as np
    lib = C.CDLL(os.path.join(HERE, "_ref", "libcsgmum_ref.so"))
    lib.ref_index_build.restype = C.c_void_p
    lib.ref_index_build.argtypes = [C.c_char_p, C.c_long, C.c_double]
    lib.ref_index_free.argtypes = [C.c_void_p]
    lib.ref_find_um.argtypes = [C.c_void_p, C.c_char_p, C.c_long, C.c_void_p, C.c_void_p]
    lib.ref_intersect_um.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib.ref_merge_master.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    comp = np.zeros(256, np.uint8)
    for a, b in zip(b"ACGTN", b"TGCAN"):
        comp[a] = b
    L0 = len(genomes[0])
    n = min(p, L0)
    R = genomes[0][:n].tobytes()
    ix = lib.ref_index_build(R, n, 2.0)
    out = []
    for Q in genomes[1:]:
        rc = comp[Q[::-1]]
        M = np.zeros(2 * n, np.int32); M[1::2] = n
        MR = M.copy(); P = np.zeros(2 * n, np.int32); PR = np.zeros(2 * n, np.int32)
        sp = np.zeros(n, np.uint64); tmp = np.zeros(n, np.uint64); fw = np.ones(n, np.int8)
        lib.ref_find_um(ix, Q.tobytes(), len(Q), sp.ctypes.data, P.ctypes.data)
        lib.ref_find_um(ix, rc.tobytes(), len(Q), tmp.ctypes.data, PR.ctypes.data)
        lib.ref_intersect_um(ix, M.ctypes.data, P.ctypes.data, n, sp.ctypes.data)
        lib.ref_intersect_um(ix, MR.ctypes.data, PR.ctypes.data, n, tmp.ctypes.data)
        lib.ref_merge_master(M.ctypes.data, MR.ctypes.data, n, sp.ctypes.data, fw.ctypes.data, tmp.ctypes.data)
        UP = M[0::2].astype(np.int64); EP = M[1::2].astype(np.int64)
        covered = np.zeros(2 * n + 16, np.int8)
        m_ep = 0
        for k in range(n):                                  # src/parsnp.cpp:2031-2066
            if EP[k] > m_ep and UP[k] < EP[k] and EP[k] - k < n:
                m_ep = EP[k]
                if EP[k] - k >= 15:
                    covered[k:EP[k]] = 1
        total = int(covered.sum())
        ratio = np.float32(L0) / np.float32(len(Q))
        if ratio > 1.3 or ratio < 0.7:
            total = 0
        total = min(total, n)
        out.append("%f" % (1.0 - float(np.float32(total) / np.float32(n))))
    lib.ref_index_free(ix)
    return out
