"""TEST INFRASTRUCTURE: ctypes view of oracle/_ref/libref_ogs.so — the reference's OWN ogs_t / mesh_t compiled
unchanged for the CPU (oracle/Makefile `ref-host`, oracle/ref_ogs_capi.cpp).  Only tests/, bench.py's CPU arm and
the golden-vector scripts may import this."""
import ctypes
import os
from ctypes import POINTER, byref, c_char_p, c_int, c_longlong, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_ref", "libref_ogs.so")
_lib = None

FLOAT, DOUBLE, INT32, INT64 = 0, 1, 2, 3
ADD, MUL, MAX, MIN = 0, 1, 2, 3
SYM, NOTRANS, TRANS = 0, 1, 2
UNSIGNED, SIGNED, HALO = 0, 1, 2
_TYPES = {np.dtype(np.float32): FLOAT, np.dtype(np.float64): DOUBLE, np.dtype(np.int32): INT32, np.dtype(np.int64): INT64}


def available():
    return os.path.exists(LIB)


def lib():
    global _lib
    if _lib is None:
        l = ctypes.CDLL(LIB)
        l.ref_ogs_create.restype = c_void_p
        l.ref_ogs_create.argtypes = [c_int, c_void_p, c_int, c_int, c_int]
        l.ref_ogs_destroy.argtypes = [c_void_p]
        l.ref_ogs_dims.argtypes = [c_void_p, c_void_p]
        l.ref_ogs_array.restype = POINTER(c_int)
        l.ref_ogs_array.argtypes = [c_void_p, c_char_p, POINTER(c_longlong)]
        l.ref_ogs_gather_scatter.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_int]
        l.ref_ogs_gather_scatter_device.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_int]
        l.ref_ogs_gather.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int]
        l.ref_ogs_scatter.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int]
        l.ref_ogs_time_gather_scatter.restype = ctypes.c_double
        l.ref_ogs_time_gather_scatter.argtypes = [c_void_p, c_int, c_int, c_int, POINTER(c_int)]
        l.ref_halo_create.restype = c_void_p
        l.ref_halo_create.argtypes = [c_int, c_void_p, c_int]
        l.ref_halo_destroy.argtypes = [c_void_p]
        l.ref_halo_nhalo.argtypes = [c_void_p]
        l.ref_halo_apply.argtypes = [c_void_p, c_int, c_void_p, c_int, c_int]
        l.ref_ogs_global_to_local.argtypes = [c_void_p, c_void_p]
        l.ref_mesh_box_hex_ids.restype = c_longlong
        l.ref_mesh_box_hex_ids.argtypes = [c_int, c_int, c_int, c_int, c_void_p, c_longlong]
        l.ref_mesh_box_ids.restype = c_longlong
        l.ref_mesh_box_ids.argtypes = [ctypes.c_char_p, c_int, c_int, c_int, c_int, c_void_p, c_longlong]
        _lib = l
    return _lib


class RefOgs:
    """libp::ogs::ogs_t::Setup(N, ids, comm(1 rank), kind, Pairwise, unique, verbose, platform(Serial))"""

    ARRAYS = ("rowStartsT", "colIdsT", "rowStartsN", "colIdsN", "blockRowStartsT", "blockRowStartsN")

    def __init__(self, ids, kind=SIGNED, unique=True, verbose=False):
        self.ids = np.ascontiguousarray(ids, dtype=np.int64).copy()
        self._h = lib().ref_ogs_create(self.ids.size, self.ids.ctypes.data, int(kind), int(bool(unique)), int(bool(verbose)))
        d = (c_longlong * 16)()
        lib().ref_ogs_dims(self._h, d)
        (self.N, self.Ngather, self.NlocalT, self.NlocalP, self.NhaloT, self.NhaloP, self.nnzT, self.nnzN,
         self.NrowBlocksT, self.NrowBlocksN, self.NgatherGlobal) = list(d)[:11]

    def global_to_local(self):
        out = np.empty(self.N, dtype=np.int32)
        lib().ref_ogs_global_to_local(self._h, out.ctypes.data)
        return out

    def array(self, name):
        n = c_longlong()
        p = lib().ref_ogs_array(self._h, name.encode(), byref(n))
        if not p or n.value == 0:
            return np.zeros(0, dtype=np.int32)
        return np.ctypeslib.as_array(p, shape=(n.value,)).copy()

    def maps(self):
        """the local operator's CSR as the reference built it (trimmed to the lengths it uses)"""
        out = {}
        out["rowStartsT"] = self.array("local_rowStartsT")[: self.NlocalT + 1]
        out["colIdsT"] = self.array("local_colIdsT")[: self.nnzT]
        out["rowStartsN"] = self.array("local_rowStartsN")[: self.NlocalT + 1]
        out["colIdsN"] = self.array("local_colIdsN")[: self.nnzN]
        out["blockRowStartsT"] = self.array("local_blockRowStartsT")[: self.NrowBlocksT + 1]
        out["blockRowStartsN"] = self.array("local_blockRowStartsN")[: self.NrowBlocksN + 1]
        return out

    def gather_scatter(self, v, k=1, op=ADD, trans=SYM, device=False):
        v = np.ascontiguousarray(v).copy()
        f = lib().ref_ogs_gather_scatter_device if device else lib().ref_ogs_gather_scatter
        rc = f(self._h, v.ctypes.data, k, _TYPES[v.dtype], op, trans)
        assert rc == 0
        return v

    def gather(self, v, k=1, op=ADD, trans=TRANS):
        v = np.ascontiguousarray(v)
        gv = np.zeros((self.NlocalT + self.NhaloT) * k, dtype=v.dtype)
        rc = lib().ref_ogs_gather(self._h, gv.ctypes.data, v.ctypes.data, k, _TYPES[v.dtype], op, trans)
        assert rc == 0
        return gv

    def scatter(self, gv, v0, k=1, trans=NOTRANS):
        gvp = np.zeros((self.NlocalT + self.NhaloT) * k, dtype=v0.dtype)
        gvp[: gv.size] = gv
        v = np.ascontiguousarray(v0).copy()
        rc = lib().ref_ogs_scatter(self._h, v.ctypes.data, gvp.ctypes.data, k, _TYPES[v.dtype], trans)
        assert rc == 0
        return v

    def time_gather_scatter(self, path, steps, warmup):
        """seconds per call of the reference's mesh.ogs.GatherScatter(q, 1, Add, Sym): path 0 = host-memory overload
        (OpenMP host operator), path 1 = deviceMemory on the Serial backend (its OKL kernel); returns (seconds, threads)"""
        th = c_int()
        t = lib().ref_ogs_time_gather_scatter(self._h, int(path), int(steps), int(warmup), byref(th))
        return float(t), th.value

    def close(self):
        if self._h:
            lib().ref_ogs_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class RefHalo:
    """libp::ogs::halo_t::Setup(N, ids, comm, Pairwise, verbose, platform(Serial)) and its host-memory Exchange / Combine"""

    def __init__(self, ids):
        ids = np.ascontiguousarray(ids, dtype=np.int64)
        self._h = lib().ref_halo_create(ids.size, ids.ctypes.data, 0)
        self.Nhalo = lib().ref_halo_nhalo(self._h)

    def _apply(self, which, v, k):
        v = np.ascontiguousarray(v).copy()
        assert lib().ref_halo_apply(self._h, which, v.ctypes.data, k, _TYPES[v.dtype]) == 0
        return v

    def exchange(self, v, k=1):
        return self._apply(0, v, k)

    def combine(self, v, k=1):
        return self._apply(1, v, k)

    def close(self):
        if self._h:
            lib().ref_halo_destroy(self._h)
            self._h = None


def mesh_box_hex_ids(nx, ny, nz, p):
    """globalIds of the reference's own box mesh of degree-p hex elements at one rank
    (libs/mesh/meshSetupBoxHex3D.cpp + meshConnectNodes.cpp:44-81)"""
    n = nx * ny * nz * (p + 1) ** 3
    ids = np.empty(n, dtype=np.int64)
    got = lib().ref_mesh_box_hex_ids(nx, ny, nz, p, ids.ctypes.data, n)
    assert got == n, got
    return ids


def mesh_box_ids(element, nx, ny, nz, p):
    """globalIds of the reference's own box mesh of `element` in {"Tri", "Quad", "Tet", "Hex"} at one rank
    (libs/mesh/meshSetupBox*.cpp + meshConnectNodes.cpp:44-81)"""
    n = lib().ref_mesh_box_ids(element.encode(), nx, ny, nz, p, None, 0)
    assert n > 0, n
    ids = np.empty(n, dtype=np.int64)
    got = lib().ref_mesh_box_ids(element.encode(), nx, ny, nz, p, ids.ctypes.data, n)
    assert got == n, got
    return ids
