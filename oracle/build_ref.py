#!/usr/bin/env python3
"""Build the REAL reference (marbl/parsnp) as the parity oracle -> oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is imported, linked or executed
by the product path (parsnp_b200/); only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may use it, and only as the
checker / CPU baseline.

What it does (no autotools, no reference sources copied into this repo):
  * copies /root/reference/{src,muscle} to a scratch dir under $TMPDIR,
  * inserts three env-gated, output-neutral hooks into the scratch copy of
    src/parsnp.cpp (anchored by line content, so a different reference fails
    loudly):
      H1  PARSNP_ORACLE_DUMP=<file>      final MUM list + LCB (cluster) list,
          dumped at the top of Aligner::writeOutput (src/parsnp.cpp:505), i.e.
          after setInterClusterRegions (src/parsnp.cpp:3270).  With
          PARSNP_ORACLE_DUMP_EXIT=1 the process exits right after the dump
          (skips libMUSCLE + XMFA; used to time the MUM+LCB path alone).
      H2  PARSNP_ORACLE_CANDS=<file>     per setMums1 window: region, window,
          minsize and the emitted candidate list (src/parsnp.cpp:1633-1695),
          written just before the arrays are freed (src/parsnp.cpp:1698).
      H3  std::chrono timer around src/parsnp.cpp:3187-3273 (anchors + recursion
          + filter + LCBs), printed as 'ORACLE_MUMLCB_SECONDS=<s>' on stderr.
  * compiles libMUSCLE objects and parsnp_core with the reference's own flags
    (src/Makefile.am:1: -fopenmp -O2 -m64 ...), -> oracle/_ref/parsnp_core_ref
  * compiles src/csgmum/{csg.c,mum.c} + src/Converter.cpp unchanged into
    oracle/_ref/libcsgmum_ref.so with a thin extern "C" shim (our code) so that
    tests can call new_CSG/build_CSG/find_leaves/Find_UM/Intersect_UM/
    Merge_Master and Converter/Calculator directly through ctypes.

Usage: python oracle/build_ref.py [--ref /root/reference] [--jobs N] [--force]
"""
import argparse
import os
import re
import shutil
import subprocess
import sys
import tempfile
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")

H1 = r'''
    { /* ---- oracle hook H1 (not part of the reference) ---- */
        const char* _dp = getenv("PARSNP_ORACLE_DUMP");
        if (_dp) {
            FILE* _f = fopen(_dp, "w");
            fprintf(_f, "N %lu\n", (unsigned long)this->n);
            for (size_t _i = 0; _i < this->mums.size(); _i++) {
                TMum& _m = this->mums[_i];
                fprintf(_f, "M %ld %ld", _m.length, _m.slength);
                for (size_t _k = 0; _k < _m.start.size(); _k++)
                    fprintf(_f, " %ld:%ld:%d", _m.start[_k], (_k < _m.end.size() ? _m.end[_k] : -1L), (int)_m.isforward[_k]);
                fprintf(_f, "\n");
            }
            for (size_t _i = 0; _i < this->clusters.size(); _i++) {
                Cluster& _c = this->clusters[_i];
                fprintf(_f, "C %d %lu %ld", _c.type, (unsigned long)_c.mums.size(), _c.length);
                for (size_t _k = 0; _k < _c.start.size(); _k++)
                    fprintf(_f, " %ld:%ld", _c.start[_k], _c.end[_k]);
                fprintf(_f, "\n");
            }
            fclose(_f);
            if (getenv("PARSNP_ORACLE_DUMP_EXIT")) exit(0);
        }
    }
'''

H2 = r'''
        { /* ---- oracle hook H2 (not part of the reference) ---- */
            const char* _cp = getenv("PARSNP_ORACLE_CANDS");
            if (_cp) {
                FILE* _f = fopen(_cp, "a");
                fprintf(_f, "W %d %d %lu %lu %lu", (int)anchors, minsize, (unsigned long)rs[0].ini_region, (unsigned long)rs[0].len_region, (unsigned long)num_mums);
                for (size_t _a = 0; _a < this->n; _a++)
                    fprintf(_f, " %ld:%ld", r1.start.at(_a), r1.length.at(_a));
                fprintf(_f, "\n");
                for (size_t _i = 0; _i < num_mums; _i++) {
                    fprintf(_f, "K %d", list_mums[_i].LON);
                    for (size_t _a = 0; _a < this->n; _a++)
                        fprintf(_f, " %lu:%d", list_mums[_i].DSP[_a], (int)list_mums[_i].forward[_a]);
                    fprintf(_f, "\n");
                }
                fclose(_f);
            }
        }
'''

H3A = r'''
    auto _oracle_t0 = std::chrono::steady_clock::now(); /* oracle hook H3 */
'''
H3B = r'''
    { /* oracle hook H3 */
        double _s = std::chrono::duration<double>(std::chrono::steady_clock::now() - _oracle_t0).count();
        fprintf(stderr, "ORACLE_MUMLCB_SECONDS=%.6f\n", _s);
    }
'''

SHIM = r'''
// extern "C" shim (ours) over the unmodified reference csgmum + Converter.
#include <string>
#include <cstring>
#include <cmath>
extern "C" {
#include "csgmum/csg.c"
#include "csgmum/mum.c"
}
#include "Converter.h"
extern "C" {
// index over text[0..n) (ASCII ACGTN); the shim appends the byte-5 terminator
// like Aligner::setMums1 does (src/parsnp.cpp:1542).
struct RefIndex { CSG* csg; char* seq; long n; };
void* ref_index_build(const char* text, long n, double factor) {
    RefIndex* ix = new RefIndex;
    ix->seq = (char*)calloc(n + 10, 1);
    memcpy(ix->seq, text, n); ix->seq[n] = (char)5; ix->n = n;
    ix->csg = 0;
    ix->csg = new_CSG(ix->csg, (unsigned long)((int)factor * n), ix->seq, n, 0);
    build_CSG(ix->csg, ix->seq, n, 0);
    find_leaves(ix->csg);
    return ix;
}
void ref_index_stats(void* h, int* last_state, int* num_nodes, int* num_leafs) {
    RefIndex* ix = (RefIndex*)h;
    *last_state = ix->csg->last_state; *num_nodes = ix->csg->num_nodes; *num_leafs = ix->csg->num_leafs;
}
void ref_index_free(void* h) { RefIndex* ix = (RefIndex*)h; free_CSG(ix->csg); free(ix->seq); delete ix; }
// Find_UM on one query strand (ASCII, length m); Pair/SP are caller arrays of n.
void ref_find_um(void* h, const char* q, long m, unsigned long* SP, int* pairUPEP) {
    RefIndex* ix = (RefIndex*)h;
    char* s = (char*)calloc(m + 10, 1); memcpy(s, q, m); s[m] = (char)5;
    Find_UM(ix->csg, s, SP, (UM*)pairUPEP);
    free(s);
}
void ref_intersect_um(void* h, int* masterUPEP, int* pairUPEP, int size, unsigned long* SP) {
    RefIndex* ix = (RefIndex*)h;
    Intersect_UM(ix->csg, (UM*)masterUPEP, (UM*)pairUPEP, size, SP);
}
// Merge_Master for a single query (pos = 0): SPF.MSP = fwdSP (in/out), SPF.forward = fwd flags (out)
void ref_merge_master(int* masterUPEP, int* masterRCUPEP, int size, unsigned long* fwdSP, char* fwdflag, unsigned long* rcSP) {
    SP spf; spf.MSP = fwdSP; spf.forward = fwdflag;
    SP spr; unsigned long two[2]; char* dummy = (char*)malloc(size); spr.MSP = two; spr.forward = dummy;
    Merge_Master((UM*)masterUPEP, (UM*)masterRCUPEP, size, 0, &spf, &spr, rcSP, 0);
    free(dummy);
}
// minsize as Aligner::setMums1 computes it (src/parsnp.cpp:1502-1514)
int ref_minsize(const char* expr, long slength) {
    std::string out;
    Converter(std::string(expr), out, 80);
    float limit = Calculator(out, out.length(), slength);
    return int(ceil(limit));
}
int ref_postfix(const char* expr, char* outbuf, int cap) {
    std::string out; Converter(std::string(expr), out, 80);
    strncpy(outbuf, out.c_str(), cap - 1); outbuf[cap - 1] = 0; return (int)out.size();
}
}
'''


def run(cmd, **kw):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, **kw)
    if r.returncode != 0:
        sys.stderr.write("FAILED: %s\n%s\n" % (" ".join(cmd), r.stdout[-4000:]))
        raise SystemExit(1)
    return r.stdout


def insert_after(lines, lineno, expect_re, text, what):
    """insert `text` after 1-based line `lineno`, which must match expect_re"""
    if not re.search(expect_re, lines[lineno - 1]):
        raise SystemExit("oracle hook %s: reference line %d does not match /%s/: %r" % (what, lineno, expect_re, lines[lineno - 1]))
    lines[lineno - 1] = lines[lineno - 1] + text


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--jobs", type=int, default=os.cpu_count() or 4)
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--asan", action="store_true", help="also build oracle/_ref/parsnp_core_ref_asan (AddressSanitizer, -O1) for diagnosing reference UB")
    a = ap.parse_args()
    exe = os.path.join(OUT, "parsnp_core_ref")
    lib = os.path.join(OUT, "libcsgmum_ref.so")
    if os.path.exists(exe) and os.path.exists(lib) and not a.force and not a.asan:
        print("oracle/_ref up to date")
        return 0
    if not os.path.isdir(os.path.join(a.ref, "src")):
        print("reference tree not present at %s; keeping prebuilt oracle/_ref" % a.ref)
        return 0 if os.path.exists(exe) else 1
    os.makedirs(OUT, exist_ok=True)
    W = tempfile.mkdtemp(prefix="parsnp_oracle_")
    try:
        shutil.copytree(os.path.join(a.ref, "src"), os.path.join(W, "src"))
        shutil.copytree(os.path.join(a.ref, "muscle"), os.path.join(W, "muscle"))
        run(["chmod", "-R", "u+w", W])
        # ---- hooks into the scratch copy of parsnp.cpp
        pc = os.path.join(W, "src", "parsnp.cpp")
        with open(pc) as f:
            lines = f.readlines()
        # apply from the bottom up so line numbers stay valid
        insert_after(lines, 3270, r"align\.setInterClusterRegions\(\);", H3B, "H3B")
        insert_after(lines, 3187, r"time \( &start\);", H3A, "H3A")
        insert_after(lines, 1697, r"^\s*$", H2, "H2")
        if not re.search(r"delete\[\] Master;", lines[1697]):
            raise SystemExit("oracle hook H2: anchor mismatch")
        insert_after(lines, 512, r'prefix\.append\("/"\);', H1, "H1")
        insert_after(lines, 60, r"#include <cmath>", "#include <chrono>\n#include <cstdio>\n", "inc")
        with open(pc, "w") as f:
            f.writelines(lines)
        # ---- libMUSCLE objects
        md = os.path.join(W, "muscle", "libMUSCLE")
        od = os.path.join(md, "obj")
        os.makedirs(od)
        srcs = sorted(x for x in os.listdir(md) if x.endswith(".cpp") and x != "main.cpp")

        def cc(x):
            run(["g++", "-O2", "-fopenmp", "-fpermissive", "-w", "-D_LIB", "-DNDEBUG=1", "-I..", "-c", x, "-o", "obj/" + x[:-4] + ".o"], cwd=md)
        with ThreadPoolExecutor(a.jobs) as ex:
            list(ex.map(cc, srcs))
        ar = os.path.join(W, "libMUSCLE-3.7.a")
        run(["ar", "rcs", ar] + [os.path.join(od, x[:-4] + ".o") for x in srcs])
        # ---- parsnp_core with the reference's flags (src/Makefile.am:1)
        sd = os.path.join(W, "src")
        run(["g++", "-fopenmp", "-O2", "-m64", "-funroll-all-loops", "-fomit-frame-pointer", "-ftree-vectorize",
             "-w", "-fpermissive", "-I../muscle", "-o", exe,
             "MuscleInterface.cpp", "parsnp.cpp", "LCB.cpp", "LCR.cpp", "TMum.cpp", "Converter.cpp", "ext/iniFile.cpp",
             ar, "-lpthread"], cwd=sd)
        if a.asan:
            run(["g++", "-fopenmp", "-O1", "-g", "-m64", "-fsanitize=address", "-fsanitize-recover=address", "-fno-omit-frame-pointer", "-w", "-fpermissive",
                 "-I../muscle", "-o", exe + "_asan",
                 "MuscleInterface.cpp", "parsnp.cpp", "LCB.cpp", "LCR.cpp", "TMum.cpp", "Converter.cpp", "ext/iniFile.cpp",
                 ar, "-lpthread"], cwd=sd)
        # ---- csgmum + Converter shared library
        shim = os.path.join(sd, "_oracle_shim.cpp")
        with open(shim, "w") as f:
            f.write(SHIM)
        run(["g++", "-O2", "-m64", "-w", "-fpermissive", "-shared", "-fPIC", "-I.", "-o", lib, "_oracle_shim.cpp", "Converter.cpp"], cwd=sd)
        print("built", exe)
        print("built", lib)
    finally:
        shutil.rmtree(W, ignore_errors=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
