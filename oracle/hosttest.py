"""ctypes access to oracle/_ref/libpb200_hosttest.so (TEST INFRASTRUCTURE ONLY): the product's host orchestrator
linked with a CPU search backend (0 = brute-force spec, 1 = real csgmum)."""
import ctypes as C
import os
import numpy as np
from parsnp_b200 import api

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libpb200_hosttest.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(LIB)
        api._decl_result_api(_lib)
        vp = C.c_void_p
        _lib.pbtest_align.argtypes = [C.c_int, C.c_int, vp, vp, vp, vp]
        _lib.pbtest_search_windows.argtypes = [C.c_int, C.c_int, vp, vp, C.c_int, vp, vp] + [vp] * 5
        _lib.pbtest_lrp.argtypes = [vp, C.c_int64, vp]
        _lib.pbtest_runoff_skips.argtypes = [C.c_int]
        _lib.pbtest_runoff_skips.restype = C.c_long
    return _lib


def align(genomes, params=None, backend=1):
    lib = load()
    keep, ptrs, lens = api._seq_arrays(genomes)
    prm = params or api.make_params()
    out = C.c_void_p()
    rc = lib.pbtest_align(backend, len(keep), ptrs, api._ptr(lens), C.byref(prm), C.byref(out))
    if rc not in (0, -5):
        raise RuntimeError(lib.pb200_last_error().decode())
    res = api.unpack_result(lib, out)
    res["no_mums"] = rc == -5
    return res


def runoff_skips(reset=True):
    """Find_UM calls the csgmum backend skipped since the last reset because the real function would have read past the end
    of the query buffer (oracle/ref_backend.cpp): on such inputs the reference binary's own answer depends on heap contents."""
    return int(load().pbtest_runoff_skips(1 if reset else 0))


def search_windows(genomes, windows, coords, backend=0):
    lib = load()
    keep, ptrs, lens = api._seq_arrays(genomes)
    nt = len(windows)
    arr = (api.CWindow * nt)(*[api.CWindow(int(a), int(b), int(c), int(d), 0) for a, b, c, d in windows])
    coords = np.ascontiguousarray(coords, np.int64)
    off = C.POINTER(C.c_int64)(); k = C.POINTER(C.c_int32)(); lon = C.POINTER(C.c_int32)()
    sp = C.POINTER(C.c_int32)(); fwd = C.POINTER(C.c_uint8)()
    lib.pbtest_search_windows(backend, len(keep), ptrs, api._ptr(lens), nt, arr, api._ptr(coords), C.byref(off), C.byref(k),
                              C.byref(lon), C.byref(sp), C.byref(fwd))
    res = api._unpack_cands(off, k, lon, sp, fwd, nt, len(keep) - 1)
    for p in (off, k, lon, sp, fwd):
        lib.pb200_free_buffer(p)
    return res


def lrp(text):
    lib = load()
    t = np.ascontiguousarray(text, np.uint8)
    out = np.zeros(len(t), np.int32)
    lib.pbtest_lrp(api._ptr(t), len(t), api._ptr(out))
    return out
