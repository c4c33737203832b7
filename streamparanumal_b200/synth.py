"""Seeded synthetic inputs for the benchmarks and tests (the reference supplies none — its arrays are
uninitialised).  Pure numpy; no compute path lives here."""
import numpy as np


def splitmix64(n, seed):
    """counter-based generator -> uniform(-1,1) doubles with 53-bit mantissas"""
    z = (np.arange(n, dtype=np.uint64) + np.uint64(seed)) * np.uint64(0x9E3779B97F4A7C15)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (2.0 / (1 << 53)) - 1.0


def box_multiplicity(rank_xyz, grid_xyz, n_xyz, p):
    """Number of coincident copies of every node of this rank's brick in the GLOBAL box of degree-p hex elements
    (element-major, node n = i + j*Nq + k*Nq^2): along each axis a node on an element boundary that is not on the
    boundary of the whole box belongs to two elements.  Computed from the lattice position alone — independent of
    any ogs setup — so GatherScatter(Add, Sym) of all-ones must return exactly this (int32 array)."""
    nq = p + 1
    per_axis = []
    for axis in range(3):
        n_el, r, g = n_xyz[axis], rank_xyz[axis], grid_xyz[axis]
        e = np.arange(n_el)[:, None] + r * n_el          # global element coordinate
        i = np.arange(nq)[None, :]
        total = g * n_el
        shared = ((i == 0) & (e > 0)) | ((i == p) & (e < total - 1))
        per_axis.append(np.where(shared, 2, 1).astype(np.int32))      # [n_el, nq]
    mx, my, mz = per_axis
    nx, ny, nz = n_xyz
    # element-major layout: e = ex + ey*nx + ez*nx*ny, n = i + j*nq + k*nq^2
    full = (mz[:, None, None, :, None, None] * my[None, :, None, None, :, None] * mx[None, None, :, None, None, :])
    return np.ascontiguousarray(full.reshape(nz * ny * nx, nq * nq * nq)).reshape(-1)
