"""Kernel-level Python surface over the C ABI (include/spn_b200.h).

Arguments follow the reference kernel_t call sites; device arrays
are torch CUDA tensors used purely as memory handles (data_ptr + current stream).
Every call goes through libspn_b200.so — there is no torch or CPU implementation
behind these names.
"""
import torch

from . import _lib
from ._lib import check

# ogs::Type / ogs::Op numeric values (reference include/ogs.hpp:178-191)
FLOAT, DOUBLE, INT32, INT64 = 0, 1, 2, 3
ADD, MUL, MAX, MIN = 0, 1, 2, 3
OPS = {"add": ADD, "mul": MUL, "max": MAX, "min": MIN}
_TYPE_OF = {torch.float32: FLOAT, torch.float64: DOUBLE, torch.int32: INT32, torch.int64: INT64}


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _dev(t, dtype=None, name="tensor"):
    if t is None:
        return None
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TypeError(f"{name} must be a CUDA tensor (no CPU path exists)")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    return t.data_ptr()


def ogs_type(t):
    try:
        return _TYPE_OF[t.dtype]
    except KeyError:
        raise TypeError(f"ogs supports float32/float64/int32/int64, got {t.dtype}") from None


def nodes_per_block():
    return _lib.kernel_lib().spn_ogs_nodes_per_block()


def reduce_workspace_bytes():
    return _lib.kernel_lib().spn_reduce_workspace_bytes()


# ---- BS0-BS5 -------------------------------------------------------------------
def bs0(N, a, b):
    check(_lib.kernel_lib().spn_bs0(int(N), _dev(a, torch.float64, "a"), _dev(b, torch.float64, "b"), _stream()), "bs0")


def copy(n, a, b):
    check(_lib.kernel_lib().spn_copy_f64(int(n), _dev(a, torch.float64, "a"), _dev(b, torch.float64, "b"), _stream()), "bs1 copy")


def axpby(n, alpha, x, beta, y):
    check(_lib.kernel_lib().spn_axpby_f64(int(n), float(alpha), _dev(x, torch.float64, "x"), float(beta),
                                          _dev(y, torch.float64, "y"), _stream()), "bs2 axpby")


def norm2(n, a, out, workspace=None):
    check(_lib.kernel_lib().spn_norm2_f64(int(n), _dev(a, torch.float64, "a"), _dev(out, torch.float64, "out"),
                                          _dev(workspace, None, "workspace"), _stream()), "bs3 norm2")


def dot(n, a, b, out, workspace=None):
    check(_lib.kernel_lib().spn_dot_f64(int(n), _dev(a, torch.float64, "a"), _dev(b, torch.float64, "b"),
                                        _dev(out, torch.float64, "out"), _dev(workspace, None, "workspace"),
                                        _stream()), "bs4 dot")


def cg_update(n, Ap, alpha, r, rdotr, workspace=None):
    check(_lib.kernel_lib().spn_cg_update_f64(int(n), _dev(Ap, torch.float64, "Ap"), float(alpha),
                                              _dev(r, torch.float64, "r"), _dev(rdotr, torch.float64, "rdotr"),
                                              _dev(workspace, None, "workspace"), _stream()), "bs5 cg_update")


# two-launch forms with the reference kernels' argument lists
def bs3_1(Nblocks, N, a, tmp):
    check(_lib.kernel_lib().spn_bs3_1(int(Nblocks), int(N), _dev(a, torch.float64), _dev(tmp, torch.float64), _stream()), "bs3_1")


def bs3_2(Nblocks, tmp, out):
    check(_lib.kernel_lib().spn_bs3_2(int(Nblocks), _dev(tmp, torch.float64), _dev(out, torch.float64), _stream()), "bs3_2")


def bs4_1(Nblocks, N, a, b, tmp):
    check(_lib.kernel_lib().spn_bs4_1(int(Nblocks), int(N), _dev(a, torch.float64), _dev(b, torch.float64),
                                      _dev(tmp, torch.float64), _stream()), "bs4_1")


def bs4_2(Nblocks, tmp, out):
    check(_lib.kernel_lib().spn_bs4_2(int(Nblocks), _dev(tmp, torch.float64), _dev(out, torch.float64), _stream()), "bs4_2")


def bs5_1(Nblocks, N, Ap, alpha, r, tmp):
    check(_lib.kernel_lib().spn_bs5_1(int(Nblocks), int(N), _dev(Ap, torch.float64), float(alpha),
                                      _dev(r, torch.float64), _dev(tmp, torch.float64), _stream()), "bs5_1")


def bs5_2(Nblocks, tmp, out):
    check(_lib.kernel_lib().spn_bs5_2(int(Nblocks), _dev(tmp, torch.float64), _dev(out, torch.float64), _stream()), "bs5_2")


# ---- ogs kernels ---------------------------------------------------------------
def ogs_gather(NrowBlocks, K, blockRowStarts, rowStarts, colIds, v, gv, op=ADD):
    if NrowBlocks == 0:
        return
    if gv.dtype != v.dtype:
        raise TypeError("v and gv must have the same dtype")
    check(_lib.kernel_lib().spn_ogs_gather(int(NrowBlocks), int(K), _dev(blockRowStarts, torch.int32),
                                           _dev(rowStarts, torch.int32), _dev(colIds, torch.int32),
                                           _dev(v), _dev(gv), ogs_type(v), int(op), _stream()), "ogs gather")


def ogs_scatter(NrowBlocks, K, blockRowStarts, rowStarts, colIds, gv, v):
    if NrowBlocks == 0:
        return
    if gv.dtype != v.dtype:
        raise TypeError("v and gv must have the same dtype")
    check(_lib.kernel_lib().spn_ogs_scatter(int(NrowBlocks), int(K), _dev(blockRowStarts, torch.int32),
                                            _dev(rowStarts, torch.int32), _dev(colIds, torch.int32),
                                            _dev(gv), _dev(v), ogs_type(v), _stream()), "ogs scatter")


def ogs_gather_scatter(NrowBlocks, K, blockRowStarts, gRowStarts, gColIds, sRowStarts, sColIds, v, op=ADD):
    if NrowBlocks == 0:
        return
    check(_lib.kernel_lib().spn_ogs_gather_scatter(
        int(NrowBlocks), int(K), _dev(blockRowStarts, torch.int32), _dev(gRowStarts, torch.int32),
        _dev(gColIds, torch.int32), _dev(sRowStarts, torch.int32), _dev(sColIds, torch.int32),
        _dev(v), ogs_type(v), int(op), _stream()), "ogs gatherScatter")


def ogs_extract(N, K, ids, q, out):
    if N == 0:
        return
    check(_lib.kernel_lib().spn_ogs_extract(int(N), int(K), _dev(ids, torch.int32), _dev(q), _dev(out),
                                            ogs_type(q), _stream()), "ogs extract")


# ---- ogs operator plan (warp-tile form of ogsOperator_t) --------------------------
class OgsPlan:
    """Device operator built once from the host rowStarts (reference: ogsOperator_t ctor,
    libs/ogs/ogsOperator.cpp:597-665).  rowStarts is a host int32 array; d_rowStarts /
    d_colIds are the device copies (kept alive by this object)."""

    def __init__(self, rowStarts, d_rowStarts, d_colIds, Ncols, colIds=None):
        import ctypes
        import numpy as np
        rs = np.ascontiguousarray(rowStarts, dtype=np.int32)
        self.Nrows = rs.size - 1
        self._keep = (d_rowStarts, d_colIds)
        self._h = ctypes.c_void_p()
        self.Ncols = int(Ncols)
        ci = None if colIds is None else np.ascontiguousarray(colIds, dtype=np.int32)
        check(_lib.kernel_lib().spn_ogs_plan_create(ctypes.byref(self._h), self.Nrows, self.Ncols, rs.ctypes.data,
                                                    None if ci is None else ci.ctypes.data,
                                                    _dev(d_rowStarts, torch.int32), _dev(d_colIds, torch.int32)),
              "ogs plan create")
        nt, nirr, nbytes = ctypes.c_int(), ctypes.c_int(), ctypes.c_size_t()
        check(_lib.kernel_lib().spn_ogs_plan_info(self._h, ctypes.byref(nt), ctypes.byref(nirr), ctypes.byref(nbytes)))
        self.Ntiles, self.Nirregular, self.plan_bytes = nt.value, nirr.value, nbytes.value

    def gather(self, K, v, gv, op=ADD):
        if gv.dtype != v.dtype:
            raise TypeError("v and gv must have the same dtype")
        check(_lib.kernel_lib().spn_ogs_plan_gather(self._h, int(K), _dev(v), _dev(gv), ogs_type(v), int(op), _stream()),
              "ogs plan gather")

    def scatter(self, K, gv, v):
        if gv.dtype != v.dtype:
            raise TypeError("v and gv must have the same dtype")
        check(_lib.kernel_lib().spn_ogs_plan_scatter(self._h, int(K), _dev(gv), _dev(v), ogs_type(v), _stream()),
              "ogs plan scatter")

    def gather_scatter(self, K, v, op=ADD):
        check(_lib.kernel_lib().spn_ogs_plan_gather_scatter(self._h, int(K), _dev(v), ogs_type(v), int(op), _stream()),
              "ogs plan gatherScatter")

    def gather_scatter_to(self, K, v, w, op=ADD):
        check(_lib.kernel_lib().spn_ogs_plan_gather_scatter_to(self._h, int(K), _dev(v), _dev(w), ogs_type(v), int(op),
                                                               _stream()), "ogs plan gatherScatter (out of place)")

    def close(self):
        if self._h:
            _lib.kernel_lib().spn_ogs_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
