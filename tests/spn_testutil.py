"""Shared test helpers: seeded inputs and small CSR builders."""
import numpy as np


def exact_a(n):
    """a[i] = (i mod 7) - 3 : integer-valued, every partial sum exact in FP64"""
    return ((np.arange(n, dtype=np.int64) % 7) - 3).astype(np.float64)


def exact_b(n):
    return ((np.arange(n, dtype=np.int64) % 5) - 2).astype(np.float64)


from streamparanumal_b200.synth import splitmix64  # noqa: F401  (one generator for the tests and the benchmark)


def hex_ids(nx, ny, nz, p):
    """Labels of the (p+1)^3 nodes of an nx*ny*nz box of hex elements; coincident nodes
    share a label.  Node n=i+j*Nq+k*Nq^2 of element e=ex+ey*nx+ez*nx*ny (reference
    numbering: libs/mesh/meshBasisHex3D.cpp:41-47, meshSetupBoxHex3D.cpp:88-96).  At one
    rank the CSR depends only on which nodes coincide, not on the label values."""
    nq = p + 1
    e = np.arange(nx * ny * nz, dtype=np.int64)
    ex, ey, ez = e % nx, (e // nx) % ny, e // (nx * ny)
    n = np.arange(nq ** 3, dtype=np.int64)
    i, j, k = n % nq, (n // nq) % nq, n // (nq * nq)
    gi = ex[:, None] * p + i[None, :]
    gj = ey[:, None] * p + j[None, :]
    gk = ez[:, None] * p + k[None, :]
    sx, sy = nx * p + 1, ny * p + 1
    return (1 + gi + gj * sx + gk * sx * sy).reshape(-1)


def csr_from_ids(ids):
    """Gather CSR the way the reference builds it at one rank: row r is the r-th distinct
    label in first-appearance order (libs/ogs/ogsSetup.cpp:531-546), columns ascending
    within a row (stable SortByRow, :561-579); zero labels do not participate (:120-137)."""
    ids = np.asarray(ids, dtype=np.int64)
    cols_all = np.flatnonzero(ids != 0)
    lab = np.abs(ids[cols_all])
    uniq, first_idx, inv = np.unique(lab, return_index=True, return_inverse=True)
    order = np.argsort(first_idx, kind="stable")
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    rows = rank[inv]
    perm = np.argsort(rows, kind="stable")
    col_ids = cols_all[perm].astype(np.int32)
    counts = np.bincount(rows, minlength=uniq.size)
    row_starts = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    return row_starts, col_ids


def greedy_blocks(row_starts, nodes_per_block):
    """Row partition with <= nodes_per_block nnz per block (a single longer row gets a
    block of its own, which exercises the kernels' oversized-block path)."""
    row_starts = np.asarray(row_starts, dtype=np.int64)
    nrows = row_starts.size - 1
    starts = [0]
    while starts[-1] < nrows:
        s = starts[-1]
        e = int(np.searchsorted(row_starts, row_starts[s] + nodes_per_block, side="right")) - 1
        starts.append(min(max(e, s + 1), nrows))
    return np.asarray(starts, dtype=np.int32)
