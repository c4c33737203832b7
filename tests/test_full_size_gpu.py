"""Parity at BASELINE.json's FULL sizes, through properties that do not need a CPU walk of the whole
problem: exact integer-valued data (sums exact in any order => bit-exact whatever
the summation tree), the oracle where it is fast enough (BS5 at N = 2^26 takes the C oracle about a
second), all-ones -> node multiplicity, Scatter o Gather == GatherScatter, sum conservation.

  C1  BS5, N = 2^26 FP64                                   (BS/BS5/src/bs5Run.cpp)
  C2  BS1-BS4 at the top of the sweep, 4 GiB moved          (BS/BS1..4/src/bs*Run.cpp)
  C3/C4  BS6 / BS7 / BS8 on the 64^3 degree-7 hex box through ogs_t (BS/BS6..8/src/bs*Run.cpp)
"""
import os

import numpy as np
import pytest
import torch

import spn_testutil as util

pytestmark = pytest.mark.gpu
REL_TOL = 1e-12


@pytest.fixture(scope="module")
def K():
    from streamparanumal_b200 import kernels
    assert torch.cuda.is_available()
    return kernels


def _exact_vec(n, mod, shift):
    """((i mod `mod`) - shift) as FP64 on the device: the 'exact set' of SURVEY.md section 8d"""
    return (torch.arange(n, dtype=torch.int64, device="cuda") % mod - shift).to(torch.float64)


def _exact_sum_of_products(n, ma, sa, mb, sb):
    """sum_i ((i%ma)-sa)*((i%mb)-sb) in exact integer arithmetic (period ma*mb)"""
    period = ma * mb
    i = np.arange(period, dtype=np.int64)
    prod = ((i % ma) - sa) * ((i % mb) - sb)
    full, rem = divmod(n, period)
    return int(full * prod.sum() + prod[:rem].sum())


def test_c1_bs5_full_size_against_the_oracle(K, oracle):
    n = 1 << 26
    Ap, r = util.splitmix64(n, 41), util.splitmix64(n, 42)
    dAp, dr = torch.from_numpy(Ap).cuda(), torch.from_numpy(r).cuda()
    out = torch.zeros(1, dtype=torch.float64, device="cuda")
    K.cg_update(n, dAp, 0.37, dr, out)
    r_ref, d_ref = oracle.cg_update(Ap, 0.37, r)
    assert np.array_equal(dr.cpu().numpy(), r_ref)            # r: bit-exact
    assert abs(out.item() - d_ref) <= REL_TOL * abs(d_ref)    # r.r: 1e-12 relative
    # exact set: every partial sum is an integer below 2^53, so ANY summation order gives the same bits
    a, b = _exact_vec(n, 7, 3), _exact_vec(n, 5, 2)
    K.cg_update(n, a, 1.0, b, out)                             # b <- b + a ; out = sum b^2
    i = np.arange(35, dtype=np.int64)
    sq = (((i % 7) - 3) + ((i % 5) - 2)) ** 2
    full, rem = divmod(n, 35)
    assert out.item() == float(full * sq.sum() + sq[:rem].sum())


def test_c2_bs1_to_bs4_at_the_top_of_the_sweep(K):
    B = 1 << 32                                                # 4 GiB moved
    n1 = B // 16
    a = _exact_vec(n1, 7, 3)
    b = torch.empty_like(a)
    K.copy(n1, a, b)
    assert torch.equal(a, b)                                   # BS1: bit-exact
    y = _exact_vec(n1, 5, 2)
    expect = a + y                                             # alpha = beta = 1.0 as bs2Run.cpp:52-53
    K.axpby(n1, 1.0, a, 1.0, y)
    assert torch.equal(y, expect)                              # BS2: bit-exact
    del expect, y
    out = torch.zeros(1, dtype=torch.float64, device="cuda")
    K.dot(n1, a, b, out)                                       # BS4 at N = B/16 (b == a)
    assert out.item() == float(_exact_sum_of_products(n1, 7, 3, 7, 3))
    del b
    n3 = B // 8                                                # BS3 at N = B/8: 2^29 entries
    a3 = torch.cat([a, a])
    del a
    K.norm2(n3, a3, out)
    assert out.item() == float(2 * _exact_sum_of_products(n1, 7, 3, 7, 3))


def test_c3_c4_ogs_full_size_properties():
    from streamparanumal_b200 import host
    nx, p = 64, 7
    comm = host.Comm(0, 1)
    platform = host.Platform(comm, 0)
    ids, nglobal = host.box_hex_ids(0, 1, nx, nx, nx, p)
    ogs = host.Ogs(ids, comm, kind=host.SIGNED, unique=True, platform=platform)
    try:
        n = ids.size
        assert n == 1 << 27 and ogs.NgatherGlobal == nglobal == (nx * p + 1) ** 3 == ogs.Ngather
        rs = ogs.array("local_rowStartsT")
        lens = np.diff(rs).astype(np.int64)
        assert set(np.unique(lens)) == {1, 2, 4, 8}
        # BS8 on all-ones: every node ends up with its multiplicity; the total is sum over rows of len^2
        q = torch.ones(n, dtype=torch.float64, device="cuda")
        ogs.GatherScatter(q, 1, host.ADD, host.SYM)
        vals, counts = torch.unique(q, return_counts=True)
        assert vals.cpu().tolist() == [1.0, 2.0, 4.0, 8.0]
        assert counts.cpu().tolist() == [int((lens == m).sum()) * m for m in (1, 2, 4, 8)]
        mult = q
        # integer-valued data: exact in any order
        x = (torch.arange(n, dtype=torch.int64, device="cuda") % 11 - 5).to(torch.float64)
        g = torch.empty(ogs.Ngather, dtype=torch.float64, device="cuda")
        ogs.Gather(g, x, 1, host.ADD, host.TRANS)             # BS6
        assert g.sum().item() == x.sum().item()                # conservation
        s = torch.zeros_like(x)
        ogs.Scatter(s, g, 1, host.NOTRANS)                     # BS7
        gs = x.clone()
        ogs.GatherScatter(gs, 1, host.ADD, host.SYM)           # BS8
        assert torch.equal(s, gs)                              # Scatter o Gather == GatherScatter, bit for bit
        gs2 = gs.clone()
        ogs.GatherScatter(gs2, 1, host.ADD, host.SYM)
        assert torch.equal(gs2, gs * mult)                     # idempotent up to the multiplicity
        # uniform(-1,1) data: BS8 must equal BS7 o BS6 bit for bit (same CSR-ordered sums)
        xr = torch.from_numpy(util.splitmix64(n, 43)).cuda()
        ogs.Gather(g, xr, 1, host.ADD, host.TRANS)
        ogs.Scatter(s, g, 1, host.NOTRANS)
        ogs.GatherScatter(xr, 1, host.ADD, host.SYM)
        assert torch.equal(s, xr)
        # ... and against the CPU oracle on the plan's own CSR at the FULL size (the oracle walks 2^27 non-zeros in
        # about a second per call with the host threads): BS6, BS7 and BS8 bit for bit
        from oracle import pyoracle
        ci = ogs.array("local_colIdsT")
        threads = os.cpu_count() or 1
        x0 = util.splitmix64(n, 43)
        assert np.array_equal(xr.cpu().numpy().view(np.int64),
                              pyoracle.ogs_gather_scatter(rs, ci, rs, ci, x0, threads=threads).view(np.int64))   # BS8
        assert np.array_equal(g.cpu().numpy().view(np.int64), pyoracle.ogs_gather(rs, ci, x0, threads=threads).view(np.int64))  # BS6
        gq = util.splitmix64(ogs.Ngather, 44)
        d_gq = torch.from_numpy(gq).cuda()
        ogs.Scatter(s, d_gq, 1, host.NOTRANS)
        assert np.array_equal(s.cpu().numpy().view(np.int64),
                              pyoracle.ogs_scatter(rs, ci, gq, np.zeros(n), threads=threads).view(np.int64))      # BS7
    finally:
        ogs.close()
        platform.close()
        comm.close()
