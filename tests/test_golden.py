"""Golden vectors of the reference's own kernels (tests/golden/ref_okl_golden.npz, written by
tests/golden/make_golden.py from oracle/_ref = the OKL sources of /root/reference compiled for the host).
CPU: the oracle restatement reproduces them bit for bit (the pin that works without the reference tree).
GPU: the CUDA path through the C ABI reproduces them — bit-exact for copy, axpby, r, every ogs kernel; 1e-12
relative for norm, dot, r.r (BASELINE.json north_star)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import cases  # noqa: E402

GOLD = np.load(os.path.join(HERE, "golden", "ref_okl_golden.npz"))
REL_TOL = 1e-12


def test_fixture_is_complete():
    keys = set(GOLD.files)
    for c in cases.stream_cases():
        for k in ("bs1", "bs2", "bs3", "bs4", "bs5_r", "bs5_rdotr"):
            assert cases.stream_key(c) + "/" + k in keys
    for c in cases.ogs_cases():
        assert cases.ogs_key(c) + "/gather" in keys and cases.ogs_key(c) + "/gatherScatter" in keys


@pytest.mark.parametrize("c", cases.stream_cases(), ids=cases.stream_key)
def test_oracle_stream_kernels_match_the_reference_outputs(oracle, c):
    a, b = cases.stream_inputs(c)
    key = cases.stream_key(c)
    assert np.array_equal(oracle.bs1(a), GOLD[key + "/bs1"])
    assert np.array_equal(oracle.bs2(c["alpha"], a, c["beta"], b), GOLD[key + "/bs2"])
    assert oracle.norm2(a) == float(GOLD[key + "/bs3"])          # same block/tree order -> same bits
    assert oracle.dot(a, b) == float(GOLD[key + "/bs4"])
    r, d = oracle.cg_update(a, c["alpha"], b, bs=c["bs"])
    assert np.array_equal(r, GOLD[key + "/bs5_r"]) and d == float(GOLD[key + "/bs5_rdotr"])


@pytest.mark.parametrize("c", cases.ogs_cases(), ids=cases.ogs_key)
def test_oracle_ogs_kernels_match_the_reference_outputs(oracle, c):
    rs, ci, v = cases.ogs_inputs(c)
    key = cases.ogs_key(c)
    g = oracle.ogs_gather(rs, ci, v, K=c["K"], op=c["op"])
    assert np.array_equal(g, GOLD[key + "/gather"])
    assert np.array_equal(oracle.ogs_gather_scatter(rs, ci, rs, ci, v, K=c["K"], op=c["op"]), GOLD[key + "/gatherScatter"])
    if c["op"] == 0:
        assert np.array_equal(oracle.ogs_scatter(rs, ci, g, np.zeros_like(v), K=c["K"]), GOLD[key + "/scatter"])
        assert np.array_equal(oracle.ogs_extract(cases.extract_ids(rs, ci, v, c), v, K=c["K"]), GOLD[key + "/extract"])


# ------------------------------------------------------------------------------------------------ GPU
def _dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.gpu
@pytest.mark.parametrize("c", cases.stream_cases(), ids=cases.stream_key)
def test_cuda_stream_kernels_match_the_reference_outputs(c):
    import torch
    from streamparanumal_b200 import kernels as K
    a, b = cases.stream_inputs(c)
    key, n = cases.stream_key(c), c["n"]
    da, db = _dev(a), _dev(b)
    out = torch.empty_like(da)
    K.copy(n, da, out)
    assert np.array_equal(out.cpu().numpy(), GOLD[key + "/bs1"])
    y = db.clone()
    K.axpby(n, c["alpha"], da, c["beta"], y)
    assert np.array_equal(y.cpu().numpy(), GOLD[key + "/bs2"])
    s = torch.zeros(1, dtype=torch.float64, device="cuda")
    K.norm2(n, da, s)
    assert abs(s.item() - float(GOLD[key + "/bs3"])) <= REL_TOL * abs(float(GOLD[key + "/bs3"]))
    K.dot(n, da, db, s)
    assert abs(s.item() - float(GOLD[key + "/bs4"])) <= REL_TOL * abs(float(GOLD[key + "/bs4"]))
    r = db.clone()
    K.cg_update(n, da, c["alpha"], r, s)
    assert np.array_equal(r.cpu().numpy(), GOLD[key + "/bs5_r"])
    assert abs(s.item() - float(GOLD[key + "/bs5_rdotr"])) <= REL_TOL * abs(float(GOLD[key + "/bs5_rdotr"]))


@pytest.mark.gpu
@pytest.mark.parametrize("c", cases.ogs_cases(), ids=cases.ogs_key)
def test_cuda_ogs_kernels_match_the_reference_outputs(c):
    import torch
    import spn_testutil as util
    from streamparanumal_b200 import kernels as K
    rs, ci, v = cases.ogs_inputs(c)
    key, Kv, op = cases.ogs_key(c), c["K"], c["op"]
    bs = util.greedy_blocks(rs, K.nodes_per_block())
    nb = bs.size - 1
    d_bs, d_rs, d_ci = _dev(bs), _dev(rs), _dev(ci)
    plan = K.OgsPlan(rs, d_rs, d_ci, v.size // Kv, colIds=ci)
    dv = _dev(v)
    for gather in (lambda out: K.ogs_gather(nb, Kv, d_bs, d_rs, d_ci, dv, out, op=op), lambda out: plan.gather(Kv, dv, out, op=op)):
        dg = torch.empty((rs.size - 1) * Kv, dtype=dv.dtype, device="cuda")
        gather(dg)
        assert np.array_equal(dg.cpu().numpy(), GOLD[key + "/gather"])
    for gs in (lambda q: K.ogs_gather_scatter(nb, Kv, d_bs, d_rs, d_ci, d_rs, d_ci, q, op=op), lambda q: plan.gather_scatter(Kv, q, op=op)):
        dq = _dev(v)
        gs(dq)
        assert np.array_equal(dq.cpu().numpy(), GOLD[key + "/gatherScatter"])
    if op == 0:
        dg = _dev(GOLD[key + "/gather"])
        for scatter in (lambda out: K.ogs_scatter(nb, Kv, d_bs, d_rs, d_ci, dg, out), lambda out: plan.scatter(Kv, dg, out)):
            ds = torch.zeros_like(dv)
            scatter(ds)
            assert np.array_equal(ds.cpu().numpy(), GOLD[key + "/scatter"])
        pick = cases.extract_ids(rs, ci, v, c)
        de = torch.empty(pick.size * Kv, dtype=dv.dtype, device="cuda")
        K.ogs_extract(pick.size, Kv, _dev(pick), dv, de)
        assert np.array_equal(de.cpu().numpy(), GOLD[key + "/extract"])
    plan.close()
