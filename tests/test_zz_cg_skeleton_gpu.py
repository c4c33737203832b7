"""the streaming skeleton of one CG iteration — BS1 (stand-in for the local operator), BS8 on Ap,
BS4 p.Ap, BS2 x += alpha p, BS5 r += alpha Ap with r.r, BS2 p = r + beta p — captured ONCE through the C ABI's graph calls
(spn_graph_begin_capture / end_capture / launch: include/spn_b200.h) and replayed.

  * the first iteration, launched kernel by kernel, equals the CPU oracle (bit-exact vectors, 1e-12 scalars);
  * three replays of the captured iteration equal three iterations launched one by one, bit for bit, scalars included
    (same kernels, same launch shapes, same order);
  * the replay is timed next to the direct launches at a launch-bound size (printed; `-s` shows it).
(The file sorts last on purpose: the graph calls are exercised after everything else.)"""
import ctypes
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

pytestmark = pytest.mark.gpu

ALPHA, BETA = 0.37, 0.61


def _iteration(K, plan, n, Ap, p, x, r, pAp, rr):
    K.copy(n, p, Ap)                    # BS1: Ap <- p (where a solver applies its local operator)
    plan.gather_scatter(1, Ap)          # BS8
    K.dot(n, p, Ap, pAp)                # BS4
    K.axpby(n, ALPHA, p, 1.0, x)        # BS2: x <- alpha p + x
    K.cg_update(n, Ap, ALPHA, r, rr)    # BS5: r <- r + alpha Ap, rr <- r.r
    K.axpby(n, 1.0, r, BETA, p)         # BS2: p <- r + beta p


def _state(n, dev):
    import spn_testutil as util
    host = {k: util.splitmix64(n, 40 + i) for i, k in enumerate(("p", "x", "r"))}
    d = {k: torch.from_numpy(v.copy()).to(dev) for k, v in host.items()}
    d["Ap"] = torch.zeros(n, dtype=torch.float64, device=dev)
    d["pAp"] = torch.zeros(1, dtype=torch.float64, device=dev)
    d["rr"] = torch.zeros(1, dtype=torch.float64, device=dev)
    return host, d


def test_cg_iteration_skeleton_replayed_from_a_graph():
    import spn_testutil as util
    from oracle import pyoracle
    from streamparanumal_b200 import _lib
    from streamparanumal_b200 import kernels as K
    lib = _lib.kernel_lib()
    dev = "cuda:0"
    torch.cuda.set_device(0)
    ids = util.hex_ids(4, 3, 3, 7)
    n = ids.size
    rs, ci = util.csr_from_ids(ids)
    d_rs, d_ci = torch.from_numpy(rs).to(dev), torch.from_numpy(ci).to(dev)
    plan = K.OgsPlan(rs, d_rs, d_ci, n)
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        host, a = _state(n, dev)
        _, b = _state(n, dev)
        args = lambda s: (K, plan, n, s["Ap"], s["p"], s["x"], s["r"], s["pAp"], s["rr"])

        # one iteration, kernel by kernel, against the oracle
        _iteration(*args(a))
        stream.synchronize()
        Ap = pyoracle.ogs_gather_scatter(rs, ci, rs, ci, host["p"].copy())
        assert np.array_equal(a["Ap"].cpu().numpy(), Ap), "BS8 inside the iteration"
        pAp = pyoracle.dot(host["p"], Ap)
        assert abs(a["pAp"].item() - pAp) <= 1e-12 * abs(pAp)
        x1 = pyoracle.bs2(ALPHA, host["p"], 1.0, host["x"])
        assert np.array_equal(a["x"].cpu().numpy(), x1)
        r1, rr = pyoracle.cg_update(Ap, ALPHA, host["r"])
        assert np.array_equal(a["r"].cpu().numpy(), r1)
        assert abs(a["rr"].item() - rr) <= 1e-12 * abs(rr)
        assert np.array_equal(a["p"].cpu().numpy(), pyoracle.bs2(1.0, r1, BETA, host["p"]))
        for _ in range(2):
            _iteration(*args(a))

        # the same iteration recorded once (nothing runs during the capture), replayed three times on the second state
        graph = ctypes.c_void_p()
        K.check(lib.spn_graph_begin_capture(stream.cuda_stream), "begin capture")
        try:
            _iteration(*args(b))
        finally:
            rc = lib.spn_graph_end_capture(stream.cuda_stream, ctypes.byref(graph))
        K.check(rc, "end capture")
        assert graph.value
        stream.synchronize()
        assert np.array_equal(b["Ap"].cpu().numpy(), np.zeros(n)), "a capture must not execute anything"
        for _ in range(3):
            K.check(lib.spn_graph_launch(graph, stream.cuda_stream), "graph launch")
        stream.synchronize()
        for k in ("Ap", "p", "x", "r", "pAp", "rr"):
            assert np.array_equal(a[k].cpu().numpy(), b[k].cpu().numpy()), f"{k}: replayed iteration differs from direct launches"

        # launch-bound size: what the replay saves per iteration
        reps = 200
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        for _ in range(10):
            _iteration(*args(a))
        ev[0].record(stream)
        for _ in range(reps):
            _iteration(*args(a))
        ev[1].record(stream)
        for _ in range(10):
            lib.spn_graph_launch(graph, stream.cuda_stream)
        ev[2].record(stream)
        for _ in range(reps):
            lib.spn_graph_launch(graph, stream.cuda_stream)
        ev[3].record(stream)
        stream.synchronize()
        direct_us = ev[0].elapsed_time(ev[1]) * 1e3 / reps
        replay_us = ev[2].elapsed_time(ev[3]) * 1e3 / reps
        print(f"\nCG iteration skeleton (6 launches, n = {n}): {direct_us:.1f} us kernel by kernel, {replay_us:.1f} us replayed from the graph")
        K.check(lib.spn_graph_destroy(graph), "graph destroy")
