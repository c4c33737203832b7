"""numpy/ctypes front-end of the CPU oracle — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; nothing under streamparanumal_b200/ does.

Two libraries:
  liboracle.so          our C restatement (oracle/stream_oracle.c, oracle/ogs_oracle.c)
  _ref/libref_okl.so    the reference's own OKL kernels compiled for the host from
                        /root/reference by oracle/Makefile (Serial semantics)
"""
import ctypes
import os
import subprocess
from ctypes import c_double, c_int, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_LIB = os.path.join(_HERE, "liboracle.so")
REF_LIB = os.path.join(_HERE, "_ref", "libref_okl.so")

TYPE_TAG = {np.dtype(np.float32): "float", np.dtype(np.float64): "double",
            np.dtype(np.int32): "int32", np.dtype(np.int64): "int64"}
OP_TAG = ["add", "mul", "max", "min"]  # ogs::Op order

_oracle = None
_ref = None


def build(ref=True):
    """(Re)build liboracle.so and, when the reference tree is present, _ref/."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"] + (["ref"] if ref else []))


def oracle_lib():
    global _oracle
    if _oracle is None:
        if not os.path.exists(ORACLE_LIB):
            build(ref=False)
        _oracle = ctypes.CDLL(ORACLE_LIB)
        for name in ("oracle_norm2", "oracle_dot", "oracle_cg_update"):
            getattr(_oracle, name).restype = c_double
    return _oracle


def have_ref():
    return os.path.exists(REF_LIB)


def ref_lib():
    global _ref
    if _ref is None:
        if not have_ref():
            raise FileNotFoundError(f"{REF_LIB} not built (needs /root/reference; run make -C oracle)")
        _ref = ctypes.CDLL(REF_LIB)
    return _ref


def _p(a):
    return a.ctypes.data_as(c_void_p)


def _f64(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a


# ---- BS1-BS5: oracle ---------------------------------------------------------------
def bs1(a):
    a = _f64(a)
    b = np.empty_like(a)
    oracle_lib().oracle_bs1(c_int(a.size), _p(a), _p(b))
    return b


def bs2(alpha, x, beta, y):
    x = _f64(x)
    y = _f64(y).copy()
    oracle_lib().oracle_bs2(c_int(x.size), c_double(alpha), _p(x), c_double(beta), _p(y))
    return y


def norm2(a, bs=256):
    a = _f64(a)
    return oracle_lib().oracle_norm2(c_int(bs), c_int(a.size), _p(a))


def dot(a, b, bs=256):
    a, b = _f64(a), _f64(b)
    return oracle_lib().oracle_dot(c_int(bs), c_int(a.size), _p(a), _p(b))


def cg_update(Ap, alpha, r, bs=256):
    """returns (r_new, rdotr)"""
    Ap = _f64(Ap)
    r = _f64(r).copy()
    rdotr = oracle_lib().oracle_cg_update(c_int(bs), c_int(Ap.size), _p(Ap), c_double(alpha), _p(r))
    return r, rdotr


# ---- BS1-BS5: reference OKL kernels (driver geometry of bs3Run.cpp:80-81 etc.) -----
def _nblock(N, bs):
    return min((N + bs - 1) // bs, bs)


def ref_bs1(a):
    a = _f64(a)
    b = np.empty_like(a)
    ref_lib().bs1(c_int(a.size), _p(a), _p(b))
    return b


def ref_bs2(alpha, x, beta, y):
    x = _f64(x)
    y = _f64(y).copy()
    ref_lib().bs2(c_int(x.size), c_double(alpha), _p(x), c_double(beta), _p(y))
    return y


def ref_norm2(a):
    a = _f64(a)
    nb = _nblock(a.size, 256)
    tmp = np.zeros(256)
    out = np.zeros(1)
    ref_lib().bs3_1(c_int(nb), c_int(a.size), _p(a), _p(tmp))
    ref_lib().bs3_2(c_int(nb), _p(tmp), _p(out))
    return float(out[0])


def ref_dot(a, b):
    a, b = _f64(a), _f64(b)
    nb = _nblock(a.size, 256)
    tmp = np.zeros(256)
    out = np.zeros(1)
    ref_lib().bs4_1(c_int(nb), c_int(a.size), _p(a), _p(b), _p(tmp))
    ref_lib().bs4_2(c_int(nb), _p(tmp), _p(out))
    return float(out[0])


def ref_cg_update(Ap, alpha, r, bs=256):
    Ap = _f64(Ap)
    r = _f64(r).copy()
    nb = _nblock(Ap.size, bs)
    tmp = np.zeros(bs)
    out = np.zeros(1)
    k1, k2 = ("bs5_1", "bs5_2") if bs == 256 else ("bs5_1_b1024", "bs5_2_b1024")
    getattr(ref_lib(), k1)(c_int(nb), c_int(Ap.size), _p(Ap), c_double(alpha), _p(r), _p(tmp))
    getattr(ref_lib(), k2)(c_int(nb), _p(tmp), _p(out))
    return r, float(out[0])


# ---- ogs kernels: oracle --------------------------------------------------------------
def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def ogs_gather(rowStarts, colIds, v, K=1, op=0, threads=1):
    rowStarts, colIds = _i32(rowStarts), _i32(colIds)
    v = np.ascontiguousarray(v)
    nrows = rowStarts.size - 1
    gv = np.empty(nrows * K, dtype=v.dtype)
    fn = getattr(oracle_lib(), "oracle_ogs_gather_" + TYPE_TAG[v.dtype])
    fn(c_int(nrows), c_int(K), _p(rowStarts), _p(colIds), _p(v), _p(gv), c_int(op), c_int(threads))
    return gv


def ogs_scatter(rowStarts, colIds, gv, v, K=1, threads=1):
    """returns a copy of v with the scattered entries overwritten"""
    rowStarts, colIds = _i32(rowStarts), _i32(colIds)
    gv = np.ascontiguousarray(gv)
    v = np.ascontiguousarray(v).copy()
    nrows = rowStarts.size - 1
    fn = getattr(oracle_lib(), "oracle_ogs_scatter_" + TYPE_TAG[v.dtype])
    fn(c_int(nrows), c_int(K), _p(rowStarts), _p(colIds), _p(gv), _p(v), c_int(threads))
    return v


def ogs_gather_scatter(gRowStarts, gColIds, sRowStarts, sColIds, v, K=1, op=0, threads=1):
    gRowStarts, gColIds = _i32(gRowStarts), _i32(gColIds)
    sRowStarts, sColIds = _i32(sRowStarts), _i32(sColIds)
    v = np.ascontiguousarray(v).copy()
    nrows = gRowStarts.size - 1
    fn = getattr(oracle_lib(), "oracle_ogs_gather_scatter_" + TYPE_TAG[v.dtype])
    fn(c_int(nrows), c_int(K), _p(gRowStarts), _p(gColIds), _p(sRowStarts), _p(sColIds), _p(v),
       c_int(op), c_int(threads))
    return v


def ogs_extract(ids, q, K=1):
    ids = _i32(ids)
    q = np.ascontiguousarray(q)
    out = np.empty(ids.size * K, dtype=q.dtype)
    fn = getattr(oracle_lib(), "oracle_ogs_extract_" + TYPE_TAG[q.dtype])
    fn(c_int(ids.size), c_int(K), _p(ids), _p(q), _p(out))
    return out


# ---- ogs kernels: reference OKL (needs the reference's 512-nnz row blocks) ---------------
def ref_ogs_gather(blockStarts, rowStarts, colIds, v, K=1, op=0):
    blockStarts, rowStarts, colIds = _i32(blockStarts), _i32(rowStarts), _i32(colIds)
    v = np.ascontiguousarray(v)
    gv = np.empty((rowStarts.size - 1) * K, dtype=v.dtype)
    fn = getattr(ref_lib(), f"ref_gather_{TYPE_TAG[v.dtype]}_{OP_TAG[op]}")
    fn(c_int(blockStarts.size - 1), c_int(K), _p(blockStarts), _p(rowStarts), _p(colIds), _p(v), _p(gv))
    return gv


def ref_ogs_scatter(blockStarts, rowStarts, colIds, gv, v, K=1):
    blockStarts, rowStarts, colIds = _i32(blockStarts), _i32(rowStarts), _i32(colIds)
    gv = np.ascontiguousarray(gv)
    v = np.ascontiguousarray(v).copy()
    fn = getattr(ref_lib(), f"ref_scatter_{TYPE_TAG[v.dtype]}_add")
    fn(c_int(blockStarts.size - 1), c_int(K), _p(blockStarts), _p(rowStarts), _p(colIds), _p(gv), _p(v))
    return v


def ref_ogs_gather_scatter(blockStarts, gRowStarts, gColIds, sRowStarts, sColIds, v, K=1, op=0):
    blockStarts = _i32(blockStarts)
    gRowStarts, gColIds = _i32(gRowStarts), _i32(gColIds)
    sRowStarts, sColIds = _i32(sRowStarts), _i32(sColIds)
    v = np.ascontiguousarray(v).copy()
    fn = getattr(ref_lib(), f"ref_gatherScatter_{TYPE_TAG[v.dtype]}_{OP_TAG[op]}")
    fn(c_int(blockStarts.size - 1), c_int(K), _p(blockStarts), _p(gRowStarts), _p(gColIds),
       _p(sRowStarts), _p(sColIds), _p(v))
    return v


def ref_ogs_extract(ids, q, K=1):
    ids = _i32(ids)
    q = np.ascontiguousarray(q)
    out = np.empty(ids.size * K, dtype=q.dtype)
    fn = getattr(ref_lib(), f"ref_extract_{TYPE_TAG[q.dtype]}_add")
    fn(c_int(ids.size), c_int(K), _p(ids), _p(q), _p(out))
    return out


# ---- row blocking (libs/ogs/ogsOperator.cpp:522-594) ---------------------------------------
def block_rows(rowStarts, nodes_per_block=512):
    """Recursive bisection of the row list until every block holds <= nodes_per_block
    non-zeros; returns blockRowStarts.  Restates ogsOperator.cpp:522-594 (upperBound at
    :497-520): split at the first row whose start is >= start + ceil(size/2), keeping at
    least one row on the right."""
    rowStarts = np.asarray(rowStarts, dtype=np.int64)
    nrows = rowStarts.size - 1
    if nrows == 0:
        return np.zeros(1, dtype=np.int32)
    if np.max(np.diff(rowStarts)) > nodes_per_block:
        raise ValueError("Multiplicity of a global node is too large")
    blocks = [(0, nrows)]
    while True:
        sizes = [rowStarts[e] - rowStarts[s] for s, e in blocks]
        if max(sizes) <= nodes_per_block:
            break
        nxt = []
        for (s, e), size in zip(blocks, sizes):
            if size > nodes_per_block:
                mid_size = rowStarts[s] + (size + 1) // 2
                mid = s + int(np.searchsorted(rowStarts[s:e], mid_size, side="left"))
                if mid == e:
                    mid -= 1
                nxt += [(s, mid), (mid, e)]
            else:
                nxt += [(s, e), (e, e)]
        blocks = nxt
    starts = [s for s, _ in blocks] + [nrows]
    return np.unique(np.asarray(starts, dtype=np.int32))
