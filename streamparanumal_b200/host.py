"""Python view of the C++ host layer (libspn_host.so, C binding in include/spn_host.h).

The classes follow the reference's C++ API names: comm_t -> Comm, platform_t -> Platform,
ogs::ogs_t -> Ogs (Setup / Gather / Scatter / GatherScatter / *Start / *Finish,
include/ogs.hpp:216-341).  Device vectors are torch CUDA tensors used as memory handles only.
There is no CPU implementation behind the device calls: Ogs without a Platform holds a host-side
plan for inspection and cannot run anything.
"""
import ctypes
import os
from ctypes import POINTER, byref, c_char_p, c_double, c_int, c_longlong, c_void_p

import numpy as np

from . import _lib

SYM, NOTRANS, TRANS = 0, 1, 2            # ogs::Transpose
UNSIGNED, SIGNED, HALO = 0, 1, 2         # ogs::Kind
ADD, MUL, MAX, MIN = 0, 1, 2, 3          # ogs::Op
FLOAT, DOUBLE, INT32, INT64 = 0, 1, 2, 3  # ogs::Type

_HOST_SIGNATURES = {
    "spn_host_comm_create": [POINTER(c_void_p), c_int, c_int, c_char_p],
    "spn_host_comm_create_world": [POINTER(c_void_p)],
    "spn_host_comm_destroy": [c_void_p],
    "spn_host_comm_rank": [c_void_p, POINTER(c_int)],
    "spn_host_comm_size": [c_void_p, POINTER(c_int)],
    "spn_host_comm_barrier": [c_void_p],
    "spn_host_comm_allreduce_i64": [c_void_p, c_void_p, c_int, c_int],
    "spn_host_comm_allreduce_f64": [c_void_p, c_void_p, c_int, c_int],
    "spn_host_comm_allreduce_device_f64": [c_void_p, c_void_p, c_int, c_void_p],
    "spn_host_platform_create": [POINTER(c_void_p), c_void_p, c_int],
    "spn_host_platform_destroy": [c_void_p],
    "spn_host_platform_set_stream": [c_void_p, c_void_p],
    "spn_host_platform_finish": [c_void_p],
    "spn_host_box_hex_ids": [c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, POINTER(c_longlong)],
    "spn_host_rank_grid": [c_int, c_int, c_void_p],
    "spn_host_ogs_setup": [POINTER(c_void_p), c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p],
    "spn_host_ogs_free": [c_void_p],
    "spn_host_ogs_info": [c_void_p, c_void_p],
    "spn_host_ogs_exchange_info": [c_void_p, c_int, c_void_p],
    "spn_host_ogs_array": [c_void_p, c_int, c_int, POINTER(c_void_p), POINTER(c_longlong)],
    "spn_host_ogs_halo_labels": [c_void_p, POINTER(c_void_p), POINTER(c_longlong)],
    "spn_host_ogs_global_to_local": [c_void_p, c_void_p],
    "spn_host_ogs_gather_scatter": [c_void_p, c_void_p, c_int, c_int, c_int, c_int],
    "spn_host_ogs_gather_scatter_start": [c_void_p, c_void_p, c_int, c_int, c_int, c_int],
    "spn_host_ogs_gather_scatter_finish": [c_void_p, c_void_p, c_int, c_int, c_int, c_int],
    "spn_host_ogs_gather_scatter_host_start": [c_void_p, c_void_p, c_int, c_int, c_int, c_int],
    "spn_host_ogs_gather_scatter_host_finish": [c_void_p, c_void_p],
    "spn_host_debug_set": [c_int, c_int],
    "spn_host_halo_exchange": [c_void_p, c_void_p, c_int, c_int],
    "spn_host_halo_exchange_start": [c_void_p, c_void_p, c_int, c_int],
    "spn_host_halo_exchange_finish": [c_void_p, c_void_p, c_int, c_int],
    "spn_host_halo_combine": [c_void_p, c_void_p, c_int, c_int],
    "spn_host_halo_combine_start": [c_void_p, c_void_p, c_int, c_int],
    "spn_host_halo_combine_finish": [c_void_p, c_void_p, c_int, c_int],
    "spn_host_ogs_gather": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int],
    "spn_host_ogs_scatter": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int],
}

_host = None


def host_lib():
    """libspn_host.so; it links libspn_b200.so and libspn_b200_comm.so (NCCL)."""
    global _host
    if _host is None:
        if not os.path.exists(_lib.HOST_LIB):
            raise ImportError(f"{_lib.HOST_LIB} is missing: build it with `make`.")
        _lib.comm_lib()  # loads the kernel library and binds torch's libnccl.so.2 first
        lib = ctypes.CDLL(_lib.HOST_LIB, mode=ctypes.RTLD_GLOBAL)
        _lib._bind(lib, _HOST_SIGNATURES)
        lib.spn_host_last_error.restype = c_char_p
        lib.spn_host_last_error.argtypes = []
        _host = lib
    return _host


def _check(rc, where):
    if rc == 0:
        return
    if rc == -16:
        raise _lib.SpnError(rc, f"{where}: {host_lib().spn_host_last_error().decode()}")
    raise _lib.SpnError(rc, where)


class Comm:
    """comm_t: host collectives over local sockets, device buffers over NCCL."""

    def __init__(self, rank=None, size=None, rendezvous=None):
        self._h = c_void_p()
        if rank is None:
            _check(host_lib().spn_host_comm_create_world(byref(self._h)), "comm_t::World")
        else:
            _check(host_lib().spn_host_comm_create(byref(self._h), int(rank), int(size),
                                                   None if rendezvous is None else os.fsencode(rendezvous)), "comm_t")
        r, s = c_int(), c_int()
        host_lib().spn_host_comm_rank(self._h, byref(r))
        host_lib().spn_host_comm_size(self._h, byref(s))
        self.rank, self.size = r.value, s.value

    def barrier(self):
        _check(host_lib().spn_host_comm_barrier(self._h), "Barrier")

    def allreduce(self, value, op="sum"):
        code = {"sum": 0, "max": 1, "min": 2}[op]
        if isinstance(value, float):
            buf = (c_double * 1)(value)
            _check(host_lib().spn_host_comm_allreduce_f64(self._h, buf, 1, code), "Allreduce")
        else:
            buf = (c_longlong * 1)(int(value))
            _check(host_lib().spn_host_comm_allreduce_i64(self._h, buf, 1, code), "Allreduce")
        return buf[0]

    def allreduce_device(self, t, stream=None):
        """in-place NCCL sum of a float64 CUDA tensor (dot / norm / rdotr across ranks)"""
        import torch
        st = torch.cuda.current_stream().cuda_stream if stream is None else stream
        _check(host_lib().spn_host_comm_allreduce_device_f64(self._h, t.data_ptr(), t.numel(), st), "AllreduceDevice")

    def close(self):
        if self._h:
            host_lib().spn_host_comm_destroy(self._h)
            self._h = c_void_p()


class Platform:
    """platform_t on one CUDA device (device_id < 0: LOCAL_RANK)."""

    def __init__(self, comm, device_id=-1):
        self.comm = comm
        self._h = c_void_p()
        _check(host_lib().spn_host_platform_create(byref(self._h), comm._h, int(device_id)), "platform_t")

    def set_stream(self, cuda_stream):
        _check(host_lib().spn_host_platform_set_stream(self._h, cuda_stream), "setStream")

    def finish(self):
        _check(host_lib().spn_host_platform_finish(self._h), "finish")

    def close(self):
        if self._h:
            host_lib().spn_host_platform_destroy(self._h)
            self._h = c_void_p()


def box_hex_ids(rank, size, nx, ny, nz, p):
    """labels of the nodes of this rank's brick; returns (ids int64, number of distinct labels over all ranks)"""
    ids = np.empty(nx * ny * nz * (p + 1) ** 3, dtype=np.int64)
    ng = c_longlong()
    _check(host_lib().spn_host_box_hex_ids(rank, size, nx, ny, nz, p, ids.ctypes.data, byref(ng)), "box_hex_ids")
    return ids, ng.value


def rank_grid(rank, size):
    """((size_x, size_y, size_z), (rank_x, rank_y, rank_z)) of the box's process grid"""
    out = (c_int * 6)()
    _check(host_lib().spn_host_rank_grid(int(rank), int(size), out), "rank_grid")
    return tuple(out[:3]), tuple(out[3:])


_ARRAYS = {"local_rowStartsT": 0, "local_colIdsT": 1, "local_rowStartsN": 2, "local_colIdsN": 3,
           "halo_rowStartsT": 4, "halo_colIdsT": 5, "halo_rowStartsN": 6, "halo_colIdsN": 7,
           "sendIds": 8, "sendRanks": 9, "sendCounts": 10, "sendOffsets": 11, "recvRanks": 12, "recvCounts": 13,
           "recvOffsets": 14, "cmbStarts": 15, "cmbIds": 16, "recvRows": 17, "rowSendStarts": 18, "rowSendDst": 19, "slotDst": 20}


class Ogs:
    """ogs::ogs_t.  Setup is collective over comm; ids (int64 numpy) gets its signs rewritten when unique."""

    def __init__(self, ids, comm, kind=SIGNED, unique=True, verbose=False, platform=None):
        ids = np.ascontiguousarray(ids, dtype=np.int64)
        self.ids = ids
        self.kind = int(kind)
        self.comm, self.platform = comm, platform
        self._h = c_void_p()
        _check(host_lib().spn_host_ogs_setup(byref(self._h), ids.size, ids.ctypes.data, comm._h, int(kind), int(bool(unique)),
                                             int(bool(verbose)), None if platform is None else platform._h), "ogs_t::Setup")
        info = (c_longlong * 16)()
        _check(host_lib().spn_host_ogs_info(self._h, info), "ogs info")
        (self.N, self.Ngather, self.NgatherGlobal, self.NlocalT, self.NlocalP, self.NhaloT, self.NhaloP) = list(info)[:7]
        self.gather_defined = bool(info[7])
        self.local_nnzT, self.local_nnzN, self.halo_nnzT, self.halo_nnzN, self.local_tiles = list(info)[8:13]
        self.peer_exchange = bool(list(info)[13])

    def exchange_info(self, mode):
        info = (c_longlong * 8)()
        _check(host_lib().spn_host_ogs_exchange_info(self._h, mode, info), "exchange info")
        return dict(zip(("Nsend", "Nrecv", "NrowsP", "Nrows", "NranksSend", "NranksRecv"), list(info)[:6]))

    def array(self, name, mode=SYM):
        p, n = c_void_p(), c_longlong()
        _check(host_lib().spn_host_ogs_array(self._h, _ARRAYS[name], mode, byref(p), byref(n)), f"ogs array {name}")
        if n.value == 0:
            return np.zeros(0, dtype=np.int32)
        return np.ctypeslib.as_array(ctypes.cast(p, POINTER(c_int)), shape=(n.value,)).copy()

    def halo_labels(self):
        p, n = c_void_p(), c_longlong()
        _check(host_lib().spn_host_ogs_halo_labels(self._h, byref(p), byref(n)), "halo labels")
        if n.value == 0:
            return np.zeros(0, dtype=np.int64)
        return np.ctypeslib.as_array(ctypes.cast(p, POINTER(c_longlong)), shape=(n.value,)).copy()

    def global_to_local(self):
        """ogs_t::SetupGlobalToLocalMapping: row of every node in a gathered vector, -1 for masked nodes"""
        out = np.empty(self.N, dtype=np.int32)
        _check(host_lib().spn_host_ogs_global_to_local(self._h, out.ctypes.data), "global to local")
        return out

    # ---- device operations (torch CUDA tensors as memory handles) ----
    def _sync_stream(self):
        import torch
        self.platform.set_stream(torch.cuda.current_stream().cuda_stream)

    @staticmethod
    def _type(t):
        from .kernels import ogs_type
        return ogs_type(t)

    def GatherScatter(self, v, k=1, op=ADD, trans=SYM):
        self._sync_stream()
        _check(host_lib().spn_host_ogs_gather_scatter(self._h, v.data_ptr(), k, self._type(v), op, trans), "GatherScatter")

    def GatherScatterStart(self, v, k=1, op=ADD, trans=SYM):
        self._sync_stream()
        _check(host_lib().spn_host_ogs_gather_scatter_start(self._h, v.data_ptr(), k, self._type(v), op, trans), "GatherScatterStart")

    def GatherScatterFinish(self, v, k=1, op=ADD, trans=SYM):
        self._sync_stream()
        _check(host_lib().spn_host_ogs_gather_scatter_finish(self._h, v.data_ptr(), k, self._type(v), op, trans), "GatherScatterFinish")

    def HostGatherScatterStart(self, h_ptr, type, k=1, op=ADD, trans=SYM):
        """h_ptr: address of a PINNED host vector (spn_host_malloc); enqueues H2D -> GatherScatter -> D2H and returns"""
        self._sync_stream()
        _check(host_lib().spn_host_ogs_gather_scatter_host_start(self._h, h_ptr, k, type, op, trans), "HostGatherScatterStart")

    def HostGatherScatterFinish(self, h_ptr):
        _check(host_lib().spn_host_ogs_gather_scatter_host_finish(self._h, h_ptr), "HostGatherScatterFinish")

    # ---- halo_t operations (handles set up with kind=HALO) ----
    def HaloExchange(self, v, k=1):
        self._sync_stream()
        _check(host_lib().spn_host_halo_exchange(self._h, v.data_ptr(), k, self._type(v)), "halo_t::Exchange")

    def HaloCombine(self, v, k=1):
        self._sync_stream()
        _check(host_lib().spn_host_halo_combine(self._h, v.data_ptr(), k, self._type(v)), "halo_t::Combine")

    def Gather(self, gv, v, k=1, op=ADD, trans=TRANS):
        self._sync_stream()
        _check(host_lib().spn_host_ogs_gather(self._h, gv.data_ptr(), v.data_ptr(), k, self._type(v), op, trans), "Gather")

    def Scatter(self, v, gv, k=1, trans=NOTRANS):
        self._sync_stream()
        _check(host_lib().spn_host_ogs_scatter(self._h, v.data_ptr(), gv.data_ptr(), k, self._type(v), trans), "Scatter")

    def close(self):
        if self._h:
            host_lib().spn_host_ogs_free(self._h)
            self._h = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
