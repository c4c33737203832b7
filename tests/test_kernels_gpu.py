"""GPU parity: every kernel through the C ABI (libspn_b200.so) against the CPU oracle
on the same seeded inputs.  Bars (BASELINE.json north_star): bit-exact for copy,
axpby, r, scatter, extract and CSR-ordered gather sums; 1e-12 relative for norm, dot,
r.r.  Run on the B200 box: pytest -m gpu."""
import numpy as np
import pytest
import torch

import spn_testutil as util

pytestmark = pytest.mark.gpu
REL_TOL = 1e-12  # north_star tolerance for FP64 reductions


@pytest.fixture(scope="module")
def K():
    from streamparanumal_b200 import kernels
    assert torch.cuda.is_available()
    return kernels


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


SIZES = [1, 2, 3, 5, 257, 1023, 4096, 65537, 1000003, (1 << 22) + 1]


@pytest.mark.parametrize("n", SIZES)
def test_copy_axpby_bit_exact(K, oracle, n):
    a, b = util.splitmix64(n, 1), util.splitmix64(n, 2)
    da, db = dev(a), torch.empty(n, dtype=torch.float64, device="cuda")
    K.copy(n, da, db)
    assert np.array_equal(db.cpu().numpy(), oracle.bs1(a))
    for alpha, beta in [(1.0, 1.0), (0.37, -1.25)]:
        dy = dev(b)
        K.axpby(n, alpha, da, beta, dy)
        assert np.array_equal(dy.cpu().numpy(), oracle.bs2(alpha, a, beta, b))


@pytest.mark.parametrize("off_a,off_b", [(1, 0), (0, 1), (1, 1), (2, 0), (3, 2)])
def test_misaligned_pointers(K, oracle, off_a, off_b):
    # deviceMemory<T>::operator+ hands the kernels arbitrary element offsets (include/memory.hpp:319-324)
    n = 10007
    a, b = util.splitmix64(n + 8, 5), util.splitmix64(n + 8, 6)
    da, db = dev(a), dev(b)
    out = torch.zeros(n + 8, dtype=torch.float64, device="cuda")
    K.copy(n, da[off_a:], out[off_b:])
    assert np.array_equal(out.cpu().numpy()[off_b:off_b + n], a[off_a:off_a + n])
    assert out[off_b + n:].abs().sum().item() == 0 and out[:off_b].abs().sum().item() == 0
    dy = db.clone()
    K.axpby(n, 0.5, da[off_a:], 2.0, dy[off_b:])
    assert np.array_equal(dy.cpu().numpy()[off_b:off_b + n], oracle.bs2(0.5, a[off_a:off_a + n], 2.0, b[off_b:off_b + n]))
    res = torch.zeros(1, dtype=torch.float64, device="cuda")
    K.dot(n, da[off_a:], db[off_b:], res)
    ref = oracle.dot(a[off_a:off_a + n], b[off_b:off_b + n])
    assert abs(res.item() - ref) <= REL_TOL * abs(ref)


@pytest.mark.parametrize("n", SIZES)
def test_reductions(K, oracle, n):
    a, b = util.splitmix64(n, 3), util.splitmix64(n, 4)
    da, db = dev(a), dev(b)
    out = torch.zeros(1, dtype=torch.float64, device="cuda")
    K.norm2(n, da, out)
    ref = oracle.norm2(a)
    assert abs(out.item() - ref) <= REL_TOL * abs(ref)
    K.dot(n, da, db, out)
    ref = oracle.dot(a, b)
    scale = float(np.sum(np.abs(a * b)))
    assert abs(out.item() - ref) <= REL_TOL * scale
    # exact inputs: any summation order gives the same bits
    ea, eb = util.exact_a(n), util.exact_b(n)
    K.norm2(n, dev(ea), out)
    assert out.item() == oracle.norm2(ea)
    K.dot(n, dev(ea), dev(eb), out)
    assert out.item() == oracle.dot(ea, eb)


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("alpha", [1.0, 0.37])
def test_cg_update(K, oracle, n, alpha):
    Ap, r = util.splitmix64(n, 7), util.splitmix64(n, 8)
    dAp, dr = dev(Ap), dev(r)
    out = torch.zeros(1, dtype=torch.float64, device="cuda")
    K.cg_update(n, dAp, alpha, dr, out)
    r1, d1 = oracle.cg_update(Ap, alpha, r)
    assert np.array_equal(dr.cpu().numpy(), r1)  # r bit-exact (unfused mul+add, as the Serial reference)
    assert abs(out.item() - d1) <= REL_TOL * abs(d1)


def test_cg_update_25_cumulative_calls(K, oracle):
    # the reference driver mutates r cumulatively: 5 warm-ups + 20 timed (BS/BS5/src/bs5Run.cpp:58-64,83-89)
    n = 100003
    Ap, r = util.splitmix64(n, 9) * 1e-2, util.splitmix64(n, 10)
    dAp, dr = dev(Ap), dev(r)
    out = torch.zeros(1, dtype=torch.float64, device="cuda")
    rr = r
    for _ in range(25):
        K.cg_update(n, dAp, 1.0, dr, out)
        rr, d = oracle.cg_update(Ap, 1.0, rr)
    assert np.array_equal(dr.cpu().numpy(), rr)
    assert abs(out.item() - d) <= REL_TOL * abs(d)


def test_two_launch_entry_points(K, oracle):
    # kernel1/kernel2 argument lists of bs3/bs4/bs5 (bs3Run.cpp:83-84 etc.)
    n = 300001
    a, b = util.splitmix64(n, 11), util.splitmix64(n, 12)
    da, db = dev(a), dev(b)
    for bs in (256, 1024):
        nb = min((n + bs - 1) // bs, bs)
        tmp = torch.full((bs,), 7.0, dtype=torch.float64, device="cuda")
        out = torch.zeros(1, dtype=torch.float64, device="cuda")
        K.bs3_1(nb, n, da, tmp)
        K.bs3_2(nb, tmp, out)
        ref = oracle.norm2(a)
        assert abs(out.item() - ref) <= REL_TOL * abs(ref)
        K.bs4_1(nb, n, da, db, tmp)
        K.bs4_2(nb, tmp, out)
        ref = oracle.dot(a, b)
        assert abs(out.item() - ref) <= REL_TOL * float(np.sum(np.abs(a * b)))
        dr = dev(b)
        K.bs5_1(nb, n, da, 0.5, dr, tmp)
        K.bs5_2(nb, tmp, out)
        r1, d1 = oracle.cg_update(a, 0.5, b)
        assert np.array_equal(dr.cpu().numpy(), r1)
        assert abs(out.item() - d1) <= REL_TOL * abs(d1)


def test_bs0_and_zero_sizes(K):
    a = torch.zeros(1, dtype=torch.float64, device="cuda")
    b = torch.ones(1, dtype=torch.float64, device="cuda")
    for _ in range(10):
        K.bs0(1, a, b)
    torch.cuda.synchronize()
    assert b.item() == 1.0
    K.copy(0, a, b)
    K.axpby(0, 1.0, a, 1.0, b)
    out = torch.full((1,), 3.0, dtype=torch.float64, device="cuda")
    K.norm2(0, a, out)
    assert out.item() == 0.0


def test_user_workspace_and_streams(K, oracle):
    n = 200003
    a = util.splitmix64(n, 13)
    da = dev(a)
    ws = torch.zeros(K.reduce_workspace_bytes() // 8 + 1, dtype=torch.float64, device="cuda")
    out = torch.zeros(1, dtype=torch.float64, device="cuda")
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            K.norm2(n, da, out, workspace=ws)
    s.synchronize()
    ref = oracle.norm2(a)
    assert abs(out.item() - ref) <= REL_TOL * abs(ref)


# ---------------------------------------------------------------------------- ogs
def _values(n, dtype, seed):
    if np.issubdtype(dtype, np.floating):
        return util.splitmix64(n, seed).astype(dtype)
    return ((np.arange(n, dtype=np.int64) * 7919 + seed) % 23 - 11).astype(dtype)


MESHES = [(2, 2, 2, 1), (3, 2, 2, 3), (4, 4, 4, 7), (1, 1, 1, 15), (2, 1, 1, 15), (8, 8, 8, 2)]


def _ops_for(dtype):
    return [0, 1, 2, 3]


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int32, np.int64])
@pytest.mark.parametrize("Kv", [1, 3])
@pytest.mark.parametrize("npb", ["native", 512])
def test_ogs_kernels_bit_exact(K, oracle, dtype, Kv, npb):
    dtype = np.dtype(dtype)
    limit = K.nodes_per_block() if npb == "native" else npb
    for m, mesh in enumerate(MESHES):
        ids = util.hex_ids(*mesh)
        rs, ci = util.csr_from_ids(ids)
        bs = oracle.block_rows(rs, limit) if npb == 512 else util.greedy_blocks(rs, limit)
        nb = bs.size - 1
        d_bs, d_rs, d_ci = dev(bs), dev(rs), dev(ci)
        plan = K.OgsPlan(rs, d_rs, d_ci, ids.size, colIds=ci)
        v = _values(ids.size * Kv, dtype, 20 + m)
        for op in _ops_for(dtype):
            if op == 1 and not np.issubdtype(dtype, np.floating):
                vv = np.where(v == 0, 1, np.sign(v)).astype(dtype)
            else:
                vv = v
            dv = dev(vv)
            dgv = torch.empty((rs.size - 1) * Kv, dtype=dv.dtype, device="cuda")
            K.ogs_gather(nb, Kv, d_bs, d_rs, d_ci, dv, dgv, op=op)
            g_ref = oracle.ogs_gather(rs, ci, vv, K=Kv, op=op)
            assert np.array_equal(dgv.cpu().numpy(), g_ref), (mesh, op)
            dq = dev(vv)
            K.ogs_gather_scatter(nb, Kv, d_bs, d_rs, d_ci, d_rs, d_ci, dq, op=op)
            assert np.array_equal(dq.cpu().numpy(), oracle.ogs_gather_scatter(rs, ci, rs, ci, vv, K=Kv, op=op)), (mesh, op)
            # the same three operations through the warp-tile plan
            dgv.zero_()
            plan.gather(Kv, dv, dgv, op=op)
            assert np.array_equal(dgv.cpu().numpy(), g_ref), ("plan", mesh, op)
            dq = dev(vv)
            plan.gather_scatter(Kv, dq, op=op)
            assert np.array_equal(dq.cpu().numpy(), oracle.ogs_gather_scatter(rs, ci, rs, ci, vv, K=Kv, op=op)), ("plan", mesh, op)
        dv = dev(v)
        g = oracle.ogs_gather(rs, ci, v, K=Kv, op=0)
        dout = torch.zeros_like(dv)
        K.ogs_scatter(nb, Kv, d_bs, d_rs, d_ci, dev(g), dout)
        assert np.array_equal(dout.cpu().numpy(), oracle.ogs_scatter(rs, ci, g, np.zeros_like(v), K=Kv))
        dout.zero_()
        plan.scatter(Kv, dev(g), dout)
        assert np.array_equal(dout.cpu().numpy(), oracle.ogs_scatter(rs, ci, g, np.zeros_like(v), K=Kv))
        pick = ((np.arange(1 + ids.size // 3) * 31) % ids.size).astype(np.int32)
        dext = torch.empty(pick.size * Kv, dtype=dv.dtype, device="cuda")
        K.ogs_extract(pick.size, Kv, dev(pick), dv, dext)
        assert np.array_equal(dext.cpu().numpy(), oracle.ogs_extract(pick, v, K=Kv))


def test_ogs_signed_variant_two_csrs(K, oracle):
    # Trans / NoTrans on a Signed operator: gather over one CSR, scatter over another with
    # the same rows (libs/ogs/ogsOperator.cpp:450-459,470-479).  The "N" CSR keeps one
    # column per row (the positive copy).
    ids = util.hex_ids(4, 3, 2, 5)
    rsT, ciT = util.csr_from_ids(ids)
    pick = (np.arange(rsT.size - 1) * 2654435761) % np.diff(rsT)
    ciN = ciT[rsT[:-1] + pick].astype(np.int32)
    rsN = np.arange(rsT.size, dtype=np.int32)
    bs = util.greedy_blocks(rsT, K.nodes_per_block())
    nb = bs.size - 1
    v = util.splitmix64(ids.size, 31)
    for (grs, gci, srs, sci) in [(rsT, ciT, rsN, ciN), (rsN, ciN, rsT, ciT)]:
        dq = dev(v)
        K.ogs_gather_scatter(nb, 1, dev(bs), dev(grs), dev(gci), dev(srs), dev(sci), dq)
        assert np.array_equal(dq.cpu().numpy(), oracle.ogs_gather_scatter(grs, gci, srs, sci, v))


def test_ogs_high_multiplicity_and_masked_nodes(K, oracle):
    # ragged rows incl. one longer than a block (oversized-block path), zero labels
    # (non-participating nodes), and the 512-nnz block edge the reference aborts beyond
    rng = np.random.default_rng(7)
    n = 60000
    ids = rng.integers(1, 9000, size=n).astype(np.int64)
    ids[rng.integers(0, n, size=5000)] = 0
    ids[rng.integers(0, n, size=3000)] = 4242        # ~3000-long row > nodes_per_block
    ids[1000:1512] = 7777                            # exactly >=512-long row
    rs, ci = util.csr_from_ids(ids)
    assert np.diff(rs).max() > K.nodes_per_block()
    bs = util.greedy_blocks(rs, K.nodes_per_block())
    nb = bs.size - 1
    v = ((np.arange(n) % 11) - 5).astype(np.float64)
    dv = dev(v)
    dgv = torch.empty(rs.size - 1, dtype=torch.float64, device="cuda")
    K.ogs_gather(nb, 1, dev(bs), dev(rs), dev(ci), dv, dgv)
    assert np.array_equal(dgv.cpu().numpy(), oracle.ogs_gather(rs, ci, v))
    dq = dev(v)
    K.ogs_gather_scatter(nb, 1, dev(bs), dev(rs), dev(ci), dev(rs), dev(ci), dq)
    ref = oracle.ogs_gather_scatter(rs, ci, rs, ci, v)
    assert np.array_equal(dq.cpu().numpy(), ref)
    assert np.array_equal(ref[ids == 0], v[ids == 0])  # masked nodes untouched
    dout = torch.full((n,), -1.0, dtype=torch.float64, device="cuda")
    K.ogs_scatter(nb, 1, dev(bs), dev(rs), dev(ci), dgv, dout)
    assert np.array_equal(dout.cpu().numpy(), oracle.ogs_scatter(rs, ci, dgv.cpu().numpy(), np.full(n, -1.0)))
    # warp-tile plan: the long rows (> 32) go through the irregular-row kernel
    d_rs, d_ci = dev(rs), dev(ci)
    plan = K.OgsPlan(rs, d_rs, d_ci, ids.size, colIds=ci)
    assert plan.Nirregular >= 2
    g2 = torch.zeros_like(dgv)
    plan.gather(1, dv, g2)
    assert np.array_equal(g2.cpu().numpy(), oracle.ogs_gather(rs, ci, v))
    dq = dev(v)
    plan.gather_scatter(1, dq)
    assert np.array_equal(dq.cpu().numpy(), ref)
    dout = torch.full((n,), -1.0, dtype=torch.float64, device="cuda")
    plan.scatter(1, g2, dout)
    assert np.array_equal(dout.cpu().numpy(), oracle.ogs_scatter(rs, ci, g2.cpu().numpy(), np.full(n, -1.0)))


def test_ogs_properties_at_scale(K):
    # size-independent properties on a mesh too big for the CPU oracle to be the checker:
    # degree 7, 32^3 elements (16.8M nodes)
    p, nx = 7, 32
    ids = util.hex_ids(nx, nx, nx, p)
    rs, ci = util.csr_from_ids(ids)
    assert rs.size - 1 == (nx * p + 1) ** 3
    bs = util.greedy_blocks(rs, K.nodes_per_block())
    nb = bs.size - 1
    d_bs, d_rs, d_ci = dev(bs), dev(rs), dev(ci)
    n = ids.size
    ones = torch.ones(n, dtype=torch.float64, device="cuda")
    K.ogs_gather_scatter(nb, 1, d_bs, d_rs, d_ci, d_rs, d_ci, ones)
    mult = ones.cpu().numpy()
    assert set(np.unique(mult)) == {1.0, 2.0, 4.0, 8.0}
    x = torch.from_numpy(((np.arange(n) % 11) - 5).astype(np.float64)).cuda()
    g = torch.empty(rs.size - 1, dtype=torch.float64, device="cuda")
    K.ogs_gather(nb, 1, d_bs, d_rs, d_ci, x, g)
    assert g.sum().item() == x.sum().item()  # integer-valued: exact in any order
    s = torch.zeros_like(x)
    K.ogs_scatter(nb, 1, d_bs, d_rs, d_ci, g, s)
    gs = x.clone()
    K.ogs_gather_scatter(nb, 1, d_bs, d_rs, d_ci, d_rs, d_ci, gs)
    assert torch.equal(s, gs)  # Scatter∘Gather == GatherScatter, bit for bit
    # idempotence up to multiplicity: gs(gs(x)) == mult * gs(x) exactly on integer data
    gs2 = gs.clone()
    K.ogs_gather_scatter(nb, 1, d_bs, d_rs, d_ci, d_rs, d_ci, gs2)
    assert torch.equal(gs2, gs * torch.from_numpy(mult).cuda())


def test_ogs_signed_zeros_and_infinities_bit_patterns(K, oracle):
    # 0.0 + (-0.0) = +0.0 is what the reference computes for every row (fold starts from the identity), and the
    # select-free FP64 fold rounds of the tile kernels (row heads add +0.0) rely on a prefix never being -0.0:
    # compare BIT PATTERNS, not values, on data full of -0.0, +0.0 and infinities
    ids = util.hex_ids(3, 3, 2, 7)
    rs, ci = util.csr_from_ids(ids)
    n = ids.size
    v = util.splitmix64(n, 77)
    pick = np.arange(n)
    v[pick % 3 == 0] = -0.0
    v[pick % 7 == 1] = 0.0
    v[pick % 101 == 5] = np.inf
    v[pick % 103 == 6] = -np.inf   # +inf + -inf rows give NaN on both sides: compared as bits only where not NaN
    d_rs, d_ci = dev(rs), dev(ci)
    plan = K.OgsPlan(rs, d_rs, d_ci, n, colIds=ci)
    bs = util.greedy_blocks(rs, K.nodes_per_block())
    d_bs = dev(bs)
    g_ref = oracle.ogs_gather(rs, ci, v)
    gs_ref = oracle.ogs_gather_scatter(rs, ci, rs, ci, v)

    def same_bits(got, ref):
        got, ref = np.asarray(got), np.asarray(ref)
        nan = np.isnan(ref)
        assert np.array_equal(np.isnan(got), nan)
        return np.array_equal(got[~nan].view(np.int64), ref[~nan].view(np.int64))

    assert (g_ref.view(np.int64) == np.float64(-0.0).view(np.int64)).sum() == 0   # no row total is -0.0
    dg = torch.empty(rs.size - 1, dtype=torch.float64, device="cuda")
    plan.gather(1, dev(v), dg)
    assert same_bits(dg.cpu().numpy(), g_ref)
    K.ogs_gather(bs.size - 1, 1, d_bs, d_rs, d_ci, dev(v), dg)
    assert same_bits(dg.cpu().numpy(), g_ref)
    dq = dev(v)
    plan.gather_scatter(1, dq)
    assert same_bits(dq.cpu().numpy(), gs_ref)
    dq = dev(v)
    K.ogs_gather_scatter(bs.size - 1, 1, d_bs, d_rs, d_ci, d_rs, d_ci, dq)
    assert same_bits(dq.cpu().numpy(), gs_ref)
    plan.close()
