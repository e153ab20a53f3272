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
        # RTLD_DEEPBIND: the library holds its own copy of the host orchestrator; in a process that has also loaded the product
        # library (RTLD_GLOBAL, parsnp_b200/api.py) its calls must keep binding to that copy, not to the product's.
        # (Sanitizer runtimes refuse DEEPBIND: not used when one is preloaded.)
        mode = C.DEFAULT_MODE if os.environ.get("LD_PRELOAD") else (os.RTLD_NOW | os.RTLD_LOCAL | os.RTLD_DEEPBIND)
        _lib = C.CDLL(LIB, mode=mode)
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


class ThreadRanks:
    """The collectives of include/parsnp_b200.h (pb200_comm_set) among `world` THREADS of this process: lets a test run the
    N>1 host path (queries / windows sharded over ranks, parsnp_b200/csrc/host/sharded.cpp) without torchrun."""

    def __init__(self, world):
        import threading
        self.world, self.bar, self.slot = world, threading.Barrier(world, timeout=300), [None] * world   # a rank that never arrives breaks the barrier instead of hanging the test
        self.calls = 0

    def _callbacks(self, rank):
        import numpy as np

        def view(ptr, nbytes):
            return np.frombuffer((C.c_uint8 * nbytes).from_address(ptr), dtype=np.uint8)

        def ag(user, send, recv, nbytes, dev):
            if nbytes == 0:
                return 0
            self.slot[rank] = view(send, nbytes).copy()
            self.bar.wait()
            view(recv, nbytes * self.world)[:] = np.concatenate(self.slot)
            self.bar.wait()
            self.calls += rank == 0
            return 0

        def ar(user, buf, count, is_max, dev):
            if count == 0:
                return 0
            v = view(buf, count * 4).view(np.int32)
            self.slot[rank] = v.copy()
            self.bar.wait()
            v[:] = (np.max if is_max else np.min)(np.stack(self.slot), axis=0)
            self.bar.wait()
            self.calls += rank == 0
            return 0

        def bc(user, buf, nbytes, root, dev):
            if nbytes == 0:
                return 0
            if rank == root:
                self.slot[root] = view(buf, nbytes).copy()
            self.bar.wait()
            if rank != root:
                view(buf, nbytes)[:] = self.slot[root]
            self.bar.wait()
            self.calls += rank == 0
            return 0
        def guarded(f):                         # a broken barrier (another rank failed) -> error code, the library throws
            def g(*a):
                try:
                    return f(*a)
                except Exception:  # noqa: BLE001
                    return 1
            return g
        return api.AG_CB(guarded(ag)), api.AR_CB(guarded(ar)), api.BC_CB(guarded(bc))

    def align(self, genomes, params=None, params_of_rank=None):
        """-> per-rank results of pbtest_align_sharded (csgmum backend), counters of rank 0.  params_of_rank: {rank: params}
        overrides (a test of the lock-step check: ranks that search different windows must fail, not mix their results)"""
        import threading
        import numpy as np
        lib = load()
        lib.pbtest_align_sharded.argtypes = [C.c_int, C.c_int, api.AG_CB, api.AR_CB, api.BC_CB, C.c_int] + [C.c_void_p] * 5
        keep, ptrs, lens = api._seq_arrays(genomes)
        prm = params or api.make_params()
        out, err, counters = [None] * self.world, [None] * self.world, np.zeros((self.world, 2), np.int64)

        def run(rank):
            try:
                cbs = self._callbacks(rank)
                h = C.c_void_p()
                mine = (params_of_rank or {}).get(rank, prm)
                rc = lib.pbtest_align_sharded(rank, self.world, *cbs, len(keep), ptrs, api._ptr(lens), C.byref(mine), C.byref(h),
                                              counters[rank].ctypes.data)
                if rc not in (0, -5):
                    raise RuntimeError("rank %d: rc %d: %s" % (rank, rc, lib.pb200_last_error().decode()))
                out[rank] = api.unpack_result(lib, h)
                out[rank]["no_mums"] = rc == -5
            except Exception as e:  # noqa: BLE001 - re-raised on the caller's thread
                err[rank] = e
                self.bar.abort()
        ts = [threading.Thread(target=run, args=(r,)) for r in range(self.world)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        for e in sorted((e for e in err if e is not None), key=lambda e: "allgather failed" in str(e) or "allreduce failed" in str(e)):
            raise e                             # the first cause, not the ranks that were only cut off by it
        return out, counters[0].tolist()


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
