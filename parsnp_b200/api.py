"""ctypes mirror of include/parsnp_b200.h - the Python host side above the C ABI.

Mirrors the reference's operator surface for the MUM + LCB path: genomes in (after the ingest rules of
src/parsnp.cpp:2999-3133), `this->mums` / `this->clusters` out (as at src/parsnp.cpp:505).  The product library is
CUDA-only; importing works anywhere, computing without a B200-class GPU raises Pb200Error (no CPU fallback).
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libparsnp_b200.so")

FLAG_TRACE_WINDOWS = 1
FLAG_NO_SPECULATION = 2
FLAG_UNALIGNED = 4


class Pb200Error(RuntimeError):
    pass


class CParams(C.Structure):
    _fields_ = [("c", C.c_int32), ("d", C.c_int32), ("q", C.c_int32), ("p", C.c_int64), ("diagdiff", C.c_float),
                ("filter", C.c_int32), ("anchors_only", C.c_int32), ("anchors", C.c_char_p), ("mums", C.c_char_p),
                ("flags", C.c_int32), ("reserved", C.c_int32)]


class CWindow(C.Structure):
    _fields_ = [("ref_start", C.c_int64), ("ref_len", C.c_int64), ("coord_off", C.c_int64), ("minsize", C.c_int32),
                ("pad", C.c_int32)]


AG_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int)
AR_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int)
BC_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int)


class _DevPtr:
    """raw device pointer -> zero-copy torch tensor via __cuda_array_interface__"""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


class TorchComm:
    """The collectives the library calls back for (include/parsnp_b200.h, pb200_comm_set), implemented with
    torch.distributed: NCCL over NVLink when the process group is NCCL (device pointers are wrapped zero-copy, host
    buffers are staged through a CUDA tensor), gloo on CPU (tests)."""

    def __init__(self, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.nccl = dist.get_backend(group) == "nccl"
        self.ag = AG_CB(self._allgather)
        self.ar = AR_CB(self._allreduce)
        self.bc = BC_CB(self._bcast)
        self.calls = {"allgather": 0, "allreduce": 0, "bcast": 0, "bytes": 0}

    def _tensor(self, ptr, nbytes, on_device):
        torch = self.torch
        if on_device:
            return torch.as_tensor(_DevPtr(ptr, nbytes), device="cuda"), None
        host = torch.frombuffer((C.c_uint8 * nbytes).from_address(ptr), dtype=torch.uint8)
        if self.nccl:
            return host.cuda(), host
        return host, None

    def _finish(self, t, host):
        if host is not None:
            host.copy_(t.cpu())
        if self.nccl:
            self.torch.cuda.synchronize()

    def _allgather(self, user, send, recv, nbytes, on_device):
        try:
            if nbytes == 0:
                return 0
            st, _ = self._tensor(send, nbytes, on_device)
            rt, rhost = self._tensor(recv, nbytes * self.world, on_device)
            self.dist.all_gather_into_tensor(rt, st, group=self.group)
            self._finish(rt, rhost)
            self.calls["allgather"] += 1
            self.calls["bytes"] += nbytes * self.world
            return 0
        except Exception as e:  # noqa: BLE001 - reported through the C ABI
            import traceback
            traceback.print_exc()
            return 1

    def _allreduce(self, user, buf, count, is_max, on_device):
        try:
            if count == 0:
                return 0
            t, host = self._tensor(buf, count * 4, on_device)
            v = t.view(self.torch.int32)
            self.dist.all_reduce(v, op=self.dist.ReduceOp.MAX if is_max else self.dist.ReduceOp.MIN, group=self.group)
            self._finish(t, host)
            self.calls["allreduce"] += 1
            self.calls["bytes"] += count * 4
            return 0
        except Exception:  # noqa: BLE001
            import traceback
            traceback.print_exc()
            return 1

    def _bcast(self, user, buf, nbytes, root, on_device):
        try:
            if nbytes == 0:
                return 0
            t, host = self._tensor(buf, nbytes, on_device)
            self.dist.broadcast(t, src=self.dist.get_global_rank(self.group, root) if self.group is not None else root, group=self.group)
            self._finish(t, host)
            self.calls["bcast"] += 1
            self.calls["bytes"] += nbytes
            return 0
        except Exception:  # noqa: BLE001
            import traceback
            traceback.print_exc()
            return 1


def make_params(c=21, d=300, q=30, p=15000000, diagdiff=0.12, filter=1, anchors_only=0,
                anchors="1.1*(Log(S))", mums="1.1*(Log(S))", flags=0):
    return CParams(c, d, q, p, diagdiff, filter, anchors_only, anchors.encode(), mums.encode(), flags, 0)


def _decl_result_api(lib):
    vp = C.c_void_p
    lib.pb200_last_error.restype = C.c_char_p
    lib.pb200_result_n.argtypes = [vp]
    lib.pb200_result_num_mums.argtypes = [vp]
    lib.pb200_result_num_mums.restype = C.c_int64
    lib.pb200_result_mums.argtypes = [vp] + [vp] * 5
    lib.pb200_result_mums_view.argtypes = [vp] + [vp] * 5
    lib.pb200_result_free.argtypes = [vp]
    lib.pb200_result_num_clusters.argtypes = [vp]
    lib.pb200_result_num_clusters.restype = C.c_int64
    lib.pb200_result_clusters.argtypes = [vp] + [vp] * 5
    lib.pb200_result_cluster_mums.argtypes = [vp, vp, vp]
    lib.pb200_result_unaligned.argtypes = [vp, vp, vp, vp]
    lib.pb200_result_unaligned.restype = C.c_int64
    lib.pb200_result_num_trace.argtypes = [vp]
    lib.pb200_result_num_trace.restype = C.c_int64
    lib.pb200_result_trace.argtypes = [vp, vp]
    lib.pb200_result_stats.argtypes = [vp, vp, C.c_int]
    lib.pb200_stats_names.restype = C.c_char_p
    lib.pb200_result_free.argtypes = [vp]
    lib.pb200_minsize.argtypes = [C.c_char_p, C.c_int64]
    lib.pb200_free_buffer.argtypes = [vp]


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class _ResultOwner:
    """keeps a pb200_result alive while numpy views of its arrays exist"""

    def __init__(self, lib, h):
        self.lib, self.h = lib, h

    def __del__(self):
        try:
            if self.h:
                self.lib.pb200_result_free(self.h)
                self.h = None
        except Exception:
            pass


def _view(owner, addr, shape, dtype):
    """numpy array over library memory at `addr` (no copy); the array keeps `owner` alive"""
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    if n == 0 or not addr:
        return np.zeros(shape, dtype)
    buf = (C.c_char * n).from_address(addr)
    buf._owner = owner
    return np.frombuffer(buf, dtype=dtype).reshape(shape)


class _Result(dict):
    """result dictionary; "mum_end" (= start + length, TMum::end) is computed when somebody asks for it"""

    def __missing__(self, key):
        if key == "mum_end":
            self[key] = self["mum_start"] + self["mum_length"][:, None]
            return self[key]
        raise KeyError(key)


def unpack_result(lib, h):
    """pb200_result* -> dict of numpy arrays.  The MUM arrays are views of the result's own memory (pb200_result_mums_view);
    the handle is freed when the last of them is garbage collected."""
    n = lib.pb200_result_n(h)
    M = lib.pb200_result_num_mums(h)
    owner = _ResultOwner(lib, h)
    p = [C.c_void_p() for _ in range(5)]
    lib.pb200_result_mums_view(h, C.byref(p[0]), C.byref(p[1]), C.byref(p[2]), None, C.byref(p[4]))
    length = _view(owner, p[0].value, (M,), np.int64); slength = _view(owner, p[1].value, (M,), np.int64)
    start = _view(owner, p[2].value, (M, n), np.int64)
    fwd = _view(owner, p[4].value, (M, n), np.uint8)
    K = lib.pb200_result_num_clusters(h)
    ctype = np.zeros(K, np.int32); cn = np.zeros(K, np.int64); cl = np.zeros(K, np.int64)
    cs = np.zeros((K, n), np.int64); ce = np.zeros((K, n), np.int64)
    lib.pb200_result_clusters(h, _ptr(ctype), _ptr(cn), _ptr(cl), _ptr(cs), _ptr(ce))
    cmo = np.zeros(K + 1, np.int64)
    cmi = np.empty(max(1, lib.pb200_result_cluster_mums(h, None, None)), np.int64)
    nidx = lib.pb200_result_cluster_mums(h, _ptr(cmo), _ptr(cmi))
    T = lib.pb200_result_num_trace(h)
    tr = np.zeros((T, 2), np.int64)
    if T:
        lib.pb200_result_trace(h, _ptr(tr))
    U = lib.pb200_result_unaligned(h, None, None, None)
    ug = np.zeros(U, np.int32); us = np.zeros(U, np.int64); ue = np.zeros(U, np.int64)
    if U:
        lib.pb200_result_unaligned(h, _ptr(ug), _ptr(us), _ptr(ue))
    names = lib.pb200_stats_names().decode().split(",")
    sv = np.zeros(len(names), np.float64)
    lib.pb200_result_stats(h, _ptr(sv), len(names))
    del owner                      # the views hold the remaining references
    return _Result(n=n, mum_length=length, mum_slength=slength, mum_start=start, mum_fwd=fwd,
                cluster_type=ctype, cluster_nmums=cn, cluster_length=cl, cluster_start=cs, cluster_end=ce,
                cluster_mum_off=cmo, cluster_mum_idx=cmi[:nidx], trace=tr, unaligned=np.stack([ug.astype(np.int64), us, ue], axis=1), stats=dict(zip(names, sv.tolist())))


def _seq_arrays(genomes):
    """genomes: list of numpy uint8 ASCII arrays -> (keepalive list, uint8** array, int64 lens)"""
    gs = [np.ascontiguousarray(g, dtype=np.uint8) for g in genomes]
    n = len(gs)
    ptrs = (C.c_void_p * n)(*[g.ctypes.data for g in gs])
    lens = np.array([len(g) for g in gs], np.int64)
    return gs, ptrs, lens


_lib = None


def load():
    """load the product library (built by __graft_entry__.build()); fails loudly if it is missing"""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise Pb200Error("parsnp_b200: %s not built - run `python -c 'import __graft_entry__ as g; g.build()'` "
                         "(the CUDA extension is required; there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    _decl_result_api(lib)
    vp = C.c_void_p
    lib.pb200_version.restype = C.c_char_p
    lib.pb200_genomes_create.argtypes = [C.c_int, C.c_int, vp, vp, vp]
    lib.pb200_genomes_free.argtypes = [vp]
    lib.pb200_search_windows.argtypes = [vp, C.c_int, vp, vp, C.c_int64] + [vp] * 5
    lib.pb200_align_resident.argtypes = [vp, vp, vp]
    lib.pb200_align.argtypes = [C.c_int, C.c_int, vp, vp, vp, vp]
    lib.pb200_mumi.argtypes = [vp, vp, vp]
    lib.pb200_engine_timers.argtypes = [vp, vp, C.c_int]
    lib.pb200_engine_timer_names.restype = C.c_char_p
    lib.pb200_engine_reset_timers.argtypes = [vp]
    lib.pb200_comm_set.argtypes = [vp, C.c_int, C.c_int, AG_CB, AR_CB, BC_CB, vp, C.c_int]
    lib.pb200_comm_clear.argtypes = [vp]
    _lib = lib
    return lib


def _check(lib, rc, allow=()):
    if rc != 0 and rc not in allow:
        raise Pb200Error("parsnp_b200 error %d: %s" % (rc, lib.pb200_last_error().decode()))
    return rc


def cuda_available():
    return bool(load().pb200_cuda_available())


def minsize(expr, slength):
    return load().pb200_minsize(expr.encode(), int(slength))


class Genomes:
    """genome texts resident in HBM (pb200_genomes)"""

    def __init__(self, genomes, device=0):
        self.lib = load()
        self._keep, self._ptrs, self._lens = _seq_arrays(genomes)
        self.n = len(self._keep)
        self.h = C.c_void_p()
        _check(self.lib, self.lib.pb200_genomes_create(device, self.n, self._ptrs, _ptr(self._lens), C.byref(self.h)))

    def close(self):
        if self.h:
            self.lib.pb200_genomes_free(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def align(self, params=None):
        prm = params or make_params()
        out = C.c_void_p()
        rc = _check(self.lib, self.lib.pb200_align_resident(self.h, C.byref(prm), C.byref(out)), allow=(-5,))
        res = unpack_result(self.lib, out)
        res["no_mums"] = rc == -5
        return res

    def search_windows(self, windows, coords):
        """windows: list of (ref_start, ref_len, coord_off, minsize); coords: int64 array.
        -> list per window of (k[], lon[], sp[c, n-1], fwd[c, n-1])"""
        nt = len(windows)
        arr = (CWindow * nt)(*[CWindow(int(a), int(b), int(c), int(d), 0) for a, b, c, d in windows])
        coords = np.ascontiguousarray(coords, np.int64)
        off = C.POINTER(C.c_int64)(); k = C.POINTER(C.c_int32)(); lon = C.POINTER(C.c_int32)()
        sp = C.POINTER(C.c_int32)(); fwd = C.POINTER(C.c_uint8)()
        _check(self.lib, self.lib.pb200_search_windows(self.h, nt, arr, _ptr(coords), len(coords), C.byref(off), C.byref(k),
                                                       C.byref(lon), C.byref(sp), C.byref(fwd)))
        res = _unpack_cands(off, k, lon, sp, fwd, nt, self.n - 1)
        for p in (off, k, lon, sp, fwd):
            self.lib.pb200_free_buffer(p)
        return res

    def mumi(self, params=None):
        """MUMi distance of every query to the reference (parsnp_core's calcmumi=1 mode, all.mumi)"""
        prm = params or make_params()
        out = np.zeros(self.n - 1, np.float64)
        _check(self.lib, self.lib.pb200_mumi(self.h, C.byref(prm), _ptr(out)))
        return out

    def set_comm(self, comm, bcast_index=True):
        """shard the search of align() over the ranks of `comm` (a TorchComm); every rank must hold the same genomes"""
        self._comm = comm           # keep the callbacks alive
        _check(self.lib, self.lib.pb200_comm_set(self.h, comm.rank, comm.world, comm.ag, comm.ar, comm.bc, None, 1 if bcast_index else 0))

    def clear_comm(self):
        self.lib.pb200_comm_clear(self.h)
        self._comm = None

    def engine_timers(self):
        names = self.lib.pb200_engine_timer_names().decode().split(",")
        v = np.zeros(len(names), np.float64)
        self.lib.pb200_engine_timers(self.h, _ptr(v), len(names))
        return dict(zip(names, v.tolist()))

    def reset_timers(self):
        self.lib.pb200_engine_reset_timers(self.h)


def _unpack_cands(off, k, lon, sp, fwd, nt, nq):
    o = np.ctypeslib.as_array(off, shape=(nt + 1,)).copy()
    tot = int(o[-1])
    kk = np.ctypeslib.as_array(k, shape=(max(tot, 1),))[:tot].copy()
    ll = np.ctypeslib.as_array(lon, shape=(max(tot, 1),))[:tot].copy()
    ss = np.ctypeslib.as_array(sp, shape=(max(tot * nq, 1),))[:tot * nq].copy().reshape(tot, nq)
    ff = np.ctypeslib.as_array(fwd, shape=(max(tot * nq, 1),))[:tot * nq].copy().reshape(tot, nq)
    return [(kk[o[t]:o[t + 1]], ll[o[t]:o[t + 1]], ss[o[t]:o[t + 1]], ff[o[t]:o[t + 1]]) for t in range(nt)]


def align(genomes, params=None, device=0):
    """end-to-end call with host buffers: upload + MUM/LCB path (pb200_align)"""
    lib = load()
    keep, ptrs, lens = _seq_arrays(genomes)
    prm = params or make_params()
    out = C.c_void_p()
    rc = _check(lib, lib.pb200_align(device, len(keep), ptrs, _ptr(lens), C.byref(prm), C.byref(out)), allow=(-5,))
    res = unpack_result(lib, out)
    res["no_mums"] = rc == -5
    return res


# ---------------------------------------------------------------- ingest (host, mirrors src/parsnp.cpp:2973-3142)
_INGEST = np.zeros(256, np.uint8)          # 0 = skipped character
for _ch in b"ACGT":
    _INGEST[_ch] = _ch
    _INGEST[_ch + 32] = _ch
for _ch in b"XYSWKHRMVDBN":
    _INGEST[_ch] = ord("N")
    _INGEST[_ch + 32] = ord("N")
_INGEST[ord("U")] = ord("T"); _INGEST[ord("u")] = ord("T")
_INGEST[ord("-")] = ord("N")


def ingest_fasta(path, is_ref, d=300):
    """FASTA -> uint8 ASCII genome exactly as parsnp_core builds `genomes[i]`: the first line is the header whatever
    it contains; A,C,G,T kept (case folded); IUPAC codes, '-' -> N; U -> T; every other character skipped; each
    further '>' line ends a contig and, in queries only, appends d+10 N's (src/parsnp.cpp:3106-3126)."""
    with open(path, "rb") as f:
        data = f.read()
    nl = data.find(b"\n")
    body = data[nl + 1:] if nl >= 0 else b""
    parts = []
    pos = 0
    pad = np.full(d + 10, ord("N"), np.uint8)
    while True:
        gt = body.find(b">", pos)
        chunk = body[pos:gt if gt >= 0 else len(body)]
        a = _INGEST[np.frombuffer(chunk, np.uint8)]
        parts.append(a[a != 0])
        if gt < 0:
            break
        if not is_ref:
            parts.append(pad)
        e = body.find(b"\n", gt)
        if e < 0:
            break
        pos = e + 1
    return np.concatenate(parts) if parts else np.zeros(0, np.uint8)
