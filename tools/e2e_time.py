"""wall-clock split of the end-to-end call (create = upload + encode, align, close) on the bench workload"""
import sys, time
sys.path.insert(0, '.')
from parsnp_b200 import api, synth
import numpy as np
g = synth.g_indep(5000000, 8, 0.01, 1)
prm = api.make_params()
for it in range(4):
    t0 = time.time(); G = api.Genomes(g); t1 = time.time(); r = G.align(prm); t2 = time.time(); G.close(); t3 = time.time()
    print("create %.1f ms align %.1f ms (core t_total %.1f) close %.1f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, r['stats']['t_total'] * 1e3, (t3 - t2) * 1e3))
import ctypes as C
lib = api.load()
G = api.Genomes(g)
for it in range(3):
    out = C.c_void_p()
    t0 = time.time(); rc = lib.pb200_align_resident(G.h, C.byref(prm), C.byref(out)); t1 = time.time(); r = api.unpack_result(lib, out); t2 = time.time()
    print("resident C call %.1f ms, python unpack %.1f ms (core t_total %.1f)" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, r['stats']['t_total'] * 1e3))
G.close()
for it in range(3):
    t0 = time.time(); r = api.align(g, prm); t1 = time.time()
    print("api.align %.1f ms (core t_total %.1f)" % ((t1 - t0) * 1e3, r['stats']['t_total'] * 1e3))
