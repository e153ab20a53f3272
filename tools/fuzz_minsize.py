#!/usr/bin/env python3
"""CPU fuzz of the minimum-length expressions (ini anchors= / mums=): parsnp_b200/csrc/host/minsize.cpp against the reference's
Converter + Calculator (src/Converter.cpp through oracle/_ref/libcsgmum_ref.so).  Test infrastructure.

  python tools/fuzz_minsize.py <seed> <expressions>

Random infix expressions over numbers, S, (Log(S)), + - * / and parentheses, evaluated at 14 region lengths.  The reference's
converter exits the process on some well-formed inputs (an operator popped from an empty stack), so every evaluation runs in a
forked child; expressions the reference does not survive are counted and skipped."""
import ctypes as C, os, sys, random
REFDIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")
lib = C.CDLL(os.path.join(REFDIR, "libcsgmum_ref.so")); lib.ref_minsize.argtypes=[C.c_char_p, C.c_long]
ht = C.CDLL(os.path.join(REFDIR, "libpb200_hosttest.so")); ht.pb200_minsize.argtypes=[C.c_char_p, C.c_int64]
rnd = random.Random(int(sys.argv[1])); N=int(sys.argv[2])
def num():
    return rnd.choice(["1.1","2","0.9","25","3","1.5","10","0.5","7","100","1.25","4.75"])
def atom(d):
    r = rnd.random()
    if r < 0.35: return "(Log(S))"
    if r < 0.5: return "S"
    if r < 0.8 or d > 2: return num()
    return "(" + expr(d+1) + ")"
def expr(d=0):
    e = atom(d)
    for _ in range(rnd.randint(0,3)):
        e += rnd.choice("+-*/") + atom(d)
    return e
lens = [1,2,3,7,10,31,100,1000,4097,30030,123456,5000000,15000000,60000000]
bad=0; died=0; ok=0
for i in range(N):
    e = expr().encode()
    r, w = os.pipe()
    pid = os.fork()
    if pid == 0:
        os.close(r)
        try:
            vals = [lib.ref_minsize(e, s) for s in lens]
            os.write(w, (" ".join(map(str, vals))).encode())
        finally:
            os._exit(0)
    os.close(w)
    data = b""
    while True:
        chunk = os.read(r, 4096)
        if not chunk: break
        data += chunk
    os.close(r); _, st = os.waitpid(pid, 0)
    if not data: died += 1; continue
    want = list(map(int, data.split()))
    # product side also in a child (it may throw/abort on forms it rejects)
    r, w = os.pipe(); pid = os.fork()
    if pid == 0:
        os.close(r)
        try:
            vals = [ht.pb200_minsize(e, s) for s in lens]
            os.write(w, (" ".join(map(str, vals))).encode())
        finally:
            os._exit(0)
    os.close(w); data=b""
    while True:
        chunk = os.read(r, 4096)
        if not chunk: break
        data += chunk
    os.close(r); os.waitpid(pid, 0)
    got = list(map(int, data.split())) if data else None
    if got != want:
        bad += 1
        if bad <= 12: print("DIFF", e.decode(), "ref", want, "mine", got, flush=True)
    else: ok += 1
print("exprs", N, "equal", ok, "differ", bad, "reference died", died)
