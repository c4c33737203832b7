"""Pins the CPU oracle (oracle/*.c) against the reference's own OKL kernels compiled
for the host (oracle/_ref/libref_okl.so, built by oracle/Makefile from /root/reference).
The reference ships no golden vectors for this path, so its Serial
kernels run here are the pin; tests/golden/ holds their outputs for boxes without the
reference tree."""
import numpy as np
import pytest

import spn_testutil as util

pyoracle = pytest.importorskip("oracle.pyoracle")
needs_ref = pytest.mark.skipif(not pyoracle.have_ref(), reason="oracle/_ref not built (no /root/reference)")

SIZES = [1, 2, 3, 255, 256, 257, 1000, 65536, 65537, 300001]


@needs_ref
@pytest.mark.parametrize("n", SIZES)
def test_bs1_bs2_bit_exact(n):
    a, b = util.splitmix64(n, 1), util.splitmix64(n, 2)
    assert np.array_equal(pyoracle.bs1(a), pyoracle.ref_bs1(a))
    for alpha, beta in [(1.0, 1.0), (0.37, -1.25)]:
        assert np.array_equal(pyoracle.bs2(alpha, a, beta, b), pyoracle.ref_bs2(alpha, a, beta, b))


@needs_ref
@pytest.mark.parametrize("n", SIZES)
def test_reductions_bit_exact(n):
    # same block/thread/tree order as the OKL -> identical bits, not just 1e-12
    a, b = util.splitmix64(n, 3), util.splitmix64(n, 4)
    assert pyoracle.norm2(a) == pyoracle.ref_norm2(a)
    assert pyoracle.dot(a, b) == pyoracle.ref_dot(a, b)
    for alpha in (1.0, 0.37):
        for bs in (256, 1024):
            r1, d1 = pyoracle.cg_update(a, alpha, b, bs=bs)
            r2, d2 = pyoracle.ref_cg_update(a, alpha, b, bs=bs)
            assert np.array_equal(r1, r2)
            assert d1 == d2


@needs_ref
def test_reductions_exact_inputs():
    n = 100003
    a, b = util.exact_a(n), util.exact_b(n)
    assert pyoracle.ref_norm2(a) == float(np.sum(a.astype(np.int64) ** 2))
    assert pyoracle.norm2(a) == pyoracle.ref_norm2(a)
    assert pyoracle.dot(a, b) == float(np.sum(a.astype(np.int64) * b.astype(np.int64)))


def _values(n, dtype, seed):
    if np.issubdtype(dtype, np.floating):
        return util.splitmix64(n, seed).astype(dtype)
    return ((np.arange(n, dtype=np.int64) * 7919 + seed) % 23 - 11).astype(dtype)


MESHES = [(2, 2, 2, 1), (3, 2, 2, 3), (2, 2, 2, 7), (1, 1, 1, 15), (4, 4, 4, 2)]


@needs_ref
@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int32, np.int64])
@pytest.mark.parametrize("op", [0, 1, 2, 3])
@pytest.mark.parametrize("K", [1, 3])
def test_ogs_kernels_bit_exact(dtype, op, K):
    dtype = np.dtype(dtype)
    for m, mesh in enumerate(MESHES):
        ids = util.hex_ids(*mesh)
        rs, ci = util.csr_from_ids(ids)
        bs = pyoracle.block_rows(rs, 512)
        v = _values(ids.size * K, dtype, 10 + m)
        if op == 1:  # keep products finite/exact enough to compare bits
            v = np.where(v == 0, 1, np.sign(v)).astype(dtype) if not np.issubdtype(dtype, np.floating) else v
        g_o = pyoracle.ogs_gather(rs, ci, v, K=K, op=op)
        g_r = pyoracle.ref_ogs_gather(bs, rs, ci, v, K=K, op=op)
        assert np.array_equal(g_o, g_r)
        gs_o = pyoracle.ogs_gather_scatter(rs, ci, rs, ci, v, K=K, op=op)
        gs_r = pyoracle.ref_ogs_gather_scatter(bs, rs, ci, rs, ci, v, K=K, op=op)
        assert np.array_equal(gs_o, gs_r)
        if op == 0:
            s_o = pyoracle.ogs_scatter(rs, ci, g_o, np.zeros_like(v), K=K)
            s_r = pyoracle.ref_ogs_scatter(bs, rs, ci, g_o, np.zeros_like(v), K=K)
            assert np.array_equal(s_o, s_r)
            pick = (np.arange(97) * 31) % ids.size
            assert np.array_equal(pyoracle.ogs_extract(pick, v, K=K), pyoracle.ref_ogs_extract(pick, v, K=K))


@needs_ref
def test_ogs_invariants_from_survey():
    # GatherScatter(Sym) of ones = multiplicity in {1,2,4,8};
    # Scatter∘Gather == GatherScatter; sum(gather(x)) == sum(x) on integer-valued data
    ids = util.hex_ids(3, 3, 3, 4)
    rs, ci = util.csr_from_ids(ids)
    ones = np.ones(ids.size)
    mult = pyoracle.ogs_gather_scatter(rs, ci, rs, ci, ones)
    assert set(np.unique(mult)) == {1.0, 2.0, 4.0, 8.0}
    x = ((np.arange(ids.size) % 11) - 5).astype(np.float64)
    g = pyoracle.ogs_gather(rs, ci, x)
    assert g.sum() == x.sum()
    assert np.array_equal(pyoracle.ogs_scatter(rs, ci, g, np.zeros_like(x)),
                          pyoracle.ogs_gather_scatter(rs, ci, rs, ci, x))
    assert rs.size - 1 == (3 * 4 + 1) ** 3
