"""The index maps of the native host layer pinned to the reference's OWN setup code.

tests/golden/ref_ogs_maps.npz holds what libp::ogs::ogs_t::Setup (libs/ogs/ogsSetup.cpp:87-290) and the ogsOperator_t
constructor (libs/ogs/ogsOperator.cpp:597-665) produce for the label sets of tests/golden/ref_map_cases.py — rowStartsT /
colIdsT / rowStartsN / colIdsN, the signs written back into ids (the rank-seeded LCG pick, ogsSetup.cpp:355-382), the row
counts — and the results of the reference's host-memory Gather / Scatter / GatherScatter on seeded vectors; it was
written by tests/golden/make_ref_maps.py from oracle/_ref/libref_ogs.so, the reference sources compiled unchanged
against include/shim.  Here:
  * CPU: spn::ogs::ogs_t::SetupHost == the reference maps, bit for bit (order of a label's copies = the serial
    reference build's, SetPositivePickOrder(1)); the default (stable) order gives the same T maps, row counts and a
    valid pick; tests/spn_testutil.csr_from_ids == the reference T maps; the closed-form box-hex labels == the reference
    mesh's converged globalIds; the CPU oracle on those maps == the reference's own operators;
  * when oracle/_ref/libref_ogs.so is present, the fixture is re-derived live and must not have drifted;
  * GPU (-m gpu): device Gather / Scatter / GatherScatter of spn::ogs::ogs_t == the reference's results, bit for bit."""
import os

import numpy as np
import pytest

import spn_testutil as util
from golden.ref_map_cases import CASES, case_ids, case_values

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "ref_ogs_maps.npz")


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


@pytest.fixture(scope="module")
def comm():
    from streamparanumal_b200 import host
    return host.Comm(0, 1, None)


def _native(name, comm, order, platform=None):
    from streamparanumal_b200 import host
    ids, kind, unique = case_ids(name)
    host.host_lib().spn_host_debug_set(1, order)
    try:
        return host.Ogs(ids.copy(), comm, kind=kind, unique=unique, platform=platform)
    finally:
        host.host_lib().spn_host_debug_set(1, 0)


@pytest.mark.parametrize("name", CASES)
def test_native_maps_equal_reference_setup(name, golden, comm):
    o = _native(name, comm, order=1)
    dims = golden[f"{name}/dims"]
    assert (o.N, o.Ngather, o.NlocalT, o.NlocalP, o.NhaloT, o.NhaloP, o.local_nnzT, o.local_nnzN) == tuple(int(x) for x in dims)
    for k in ("rowStartsT", "colIdsT", "rowStartsN", "colIdsN"):
        ref = golden[f"{name}/{k}"]
        got = o.array("local_" + k)[: ref.size]
        assert np.array_equal(got, ref), f"{name}: {k} differs from the reference's ogs_t::Setup"
    assert np.array_equal(o.ids, golden[f"{name}/ids_out"]), f"{name}: signs written back differ from the reference's pick"


@pytest.mark.parametrize("threads", ["1", "3", "13"])
def test_setup_does_not_depend_on_the_thread_count(threads, golden, comm, monkeypatch):
    """the setup's order-preserving parallel passes (chunked selections, the skip-ahead generator of the positive pick, the
    radix sort) give the reference's maps and picks for any number of threads — including more chunks than label groups"""
    monkeypatch.setenv("SPN_SETUP_THREADS", threads)
    for name in ("hex_p3_nx4", "hex_p7_nx2", "hex_p7_masked", "random_labels", "long_rows"):
        o = _native(name, comm, order=1)
        for k in ("rowStartsT", "colIdsT", "rowStartsN", "colIdsN"):
            ref = golden[f"{name}/{k}"]
            assert np.array_equal(o.array("local_" + k)[: ref.size], ref), f"{name}: {k} with {threads} threads"
        assert np.array_equal(o.ids, golden[f"{name}/ids_out"]), f"{name}: pick with {threads} threads"


@pytest.mark.parametrize("name", CASES)
def test_stable_pick_order_changes_only_the_positive_copy(name, golden, comm):
    """default order of a label's copies: same T maps and counts; N maps name one (possibly other) copy of each row"""
    o = _native(name, comm, order=0)
    dims = golden[f"{name}/dims"]
    assert (o.N, o.Ngather, o.NlocalT, o.NlocalP) == tuple(int(x) for x in dims[:4])
    rsT, ciT = golden[f"{name}/rowStartsT"], golden[f"{name}/colIdsT"]
    assert np.array_equal(o.array("local_rowStartsT")[: rsT.size], rsT)
    assert np.array_equal(o.array("local_colIdsT")[: ciT.size], ciT)
    rsN, ciN = o.array("local_rowStartsN")[: rsT.size], o.array("local_colIdsN")
    assert np.array_equal(rsN, golden[f"{name}/rowStartsN"])  # one positive per row either way
    for r in range(0, rsT.size - 1, max(1, (rsT.size - 1) // 257)):
        assert set(ciN[rsN[r]:rsN[r + 1]]) <= set(ciT[rsT[r]:rsT[r + 1]])


@pytest.mark.parametrize("name", [c for c in CASES if c.startswith("hex") or c == "random_labels"])
def test_test_helper_csr_equals_reference(name, golden):
    """tests/spn_testutil.csr_from_ids (used all over the kernel tests) is the reference's T map"""
    ids, _, _ = case_ids(name)
    rs, ci = util.csr_from_ids(ids)
    assert np.array_equal(rs, golden[f"{name}/rowStartsT"]) and np.array_equal(ci, golden[f"{name}/colIdsT"])


@pytest.mark.parametrize("name", CASES)
def test_oracle_operators_equal_reference_operators(name, golden, oracle):
    """the CPU oracle (restatement of libs/ogs/ogsOperator.cpp:61-114,223-267,340-409) on the reference's maps gives the
    reference's own host-memory results"""
    ids, kind, _ = case_ids(name)
    v = case_values(name, ids.size)
    T = (golden[f"{name}/rowStartsT"], golden[f"{name}/colIdsT"])
    N = (golden[f"{name}/rowStartsN"], golden[f"{name}/colIdsN"])
    got = oracle.ogs_gather_scatter(T[0], T[1], T[0], T[1], v)
    assert np.array_equal(got.view(np.int64), golden[f"{name}/gs_sym"].view(np.int64))
    if f"{name}/gs_trans" in golden.files and kind != 0:
        # Trans: gather over T, scatter over N; NoTrans: gather over N, scatter over T (ogsOperator.cpp:450-480)
        nrows = T[0].size - 1
        rsN = np.concatenate([N[0], np.full(nrows + 1 - N[0].size, N[0][-1], dtype=np.int32)]) if N[0].size < nrows + 1 else N[0]
        got = oracle.ogs_gather_scatter(T[0], T[1], rsN, N[1], v)
        assert np.array_equal(got.view(np.int64), golden[f"{name}/gs_trans"].view(np.int64))
        got = oracle.ogs_gather_scatter(rsN, N[1], T[0], T[1], v)
        assert np.array_equal(got.view(np.int64), golden[f"{name}/gs_notrans"].view(np.int64))
    if f"{name}/gather_trans" in golden.files:
        gv = oracle.ogs_gather(T[0], T[1], v)
        ref = golden[f"{name}/gather_trans"]
        assert np.array_equal(gv[: ref.size].view(np.int64), ref.view(np.int64))


@pytest.mark.parametrize("dims", [(2, 2, 2, 1), (3, 2, 2, 3), (2, 3, 4, 7)])
def test_box_hex_labels_equal_reference_mesh(dims, golden, comm):
    """BoxHexGlobalIds (closed form) == |globalIds| of the reference's mesh_t after ConnectNodes
    (libs/mesh/meshConnectNodes.cpp:44-81: min over coincident copies of 1 + e*Np + n)"""
    from streamparanumal_b200 import host
    nx, ny, nz, p = dims
    ref = golden[f"mesh_{nx}_{ny}_{nz}_{p}"]
    ids, nglobal = host.box_hex_ids(0, 1, nx, ny, nz, p)
    assert np.array_equal(ids, np.abs(ref))
    assert nglobal == (nx * p + 1) * (ny * p + 1) * (nz * p + 1) == np.unique(np.abs(ref)).size
    # and the reference's signs are what ogs_t::Setup(unique) writes back for these labels
    o = host.Ogs(ids.copy(), comm, kind=host.SIGNED, unique=True, platform=None)
    host.host_lib().spn_host_debug_set(1, 1)
    try:
        o = host.Ogs(ids.copy(), comm, kind=host.SIGNED, unique=True, platform=None)
    finally:
        host.host_lib().spn_host_debug_set(1, 0)
    assert np.array_equal(o.ids, ref)


def test_fixture_matches_the_live_reference_build(golden):
    from oracle import refogs
    if not refogs.available():
        pytest.skip("oracle/_ref/libref_ogs.so not built here (needs /root/reference); the committed fixture stands")
    for name in ("hex_p3_nx2", "hex_p7_masked", "random_labels", "long_rows"):
        ids, kind, unique = case_ids(name)
        r = refogs.RefOgs(ids, kind=kind, unique=unique)
        for k, v in r.maps().items():
            assert np.array_equal(v, golden[f"{name}/{k}"]), (name, k)
        assert np.array_equal(r.ids, golden[f"{name}/ids_out"])
        got = r.gather_scatter(case_values(name, ids.size), 1, refogs.ADD, refogs.SYM)
        assert np.array_equal(got.view(np.int64), golden[f"{name}/gs_sym"].view(np.int64))
        r.close()
    assert np.array_equal(refogs.mesh_box_hex_ids(3, 2, 2, 3), golden["mesh_3_2_2_3"])


# ---- GPU: the device operations of the native ogs_t against the reference's results ------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_device_operations_equal_reference_results(name, golden):
    import torch
    from streamparanumal_b200 import host
    comm = host.Comm(0, 1, None)
    platform = host.Platform(comm, 0)
    o = _native(name, comm, order=1, platform=platform)
    ids, kind, _ = case_ids(name)
    v = case_values(name, ids.size)
    for tname, trans in (("sym", host.SYM), ("notrans", host.NOTRANS), ("trans", host.TRANS)):
        key = f"{name}/gs_{tname}"
        if key not in golden.files:
            continue
        d_v = torch.from_numpy(v.copy()).cuda()
        o.GatherScatter(d_v, 1, host.ADD, trans)
        assert np.array_equal(d_v.cpu().numpy().view(np.int64), golden[key].view(np.int64)), key
    if f"{name}/gather_trans" in golden.files:
        ref = golden[f"{name}/gather_trans"]
        d_v = torch.from_numpy(v.copy()).cuda()
        d_gv = torch.zeros(max(o.NlocalT + o.NhaloT, 1), dtype=torch.float64, device="cuda")
        o.Gather(d_gv, d_v, 1, host.ADD, host.TRANS)
        assert np.array_equal(d_gv.cpu().numpy()[: ref.size].view(np.int64), ref.view(np.int64))
        gq = case_values(name + "g", o.NlocalT + o.NhaloT)
        d_gq = torch.from_numpy(gq.copy()).cuda()
        d_out = torch.zeros(ids.size, dtype=torch.float64, device="cuda")
        o.Scatter(d_out, d_gq, 1, host.NOTRANS)
        assert np.array_equal(d_out.cpu().numpy().view(np.int64), golden[f"{name}/scatter_notrans"].view(np.int64))
    o.close()
    platform.close()


def test_random_label_sets_against_the_live_reference(comm):
    """120 random label sets — sizes 0..5000, few or many copies per label, labels above 2^33, zeros, negative inputs with
    unique = true, unsigned, caller-signed with unique = false — through the reference's ogs_t::Setup (oracle/_ref) and the
    native setup: same counts, same signs written back, same four maps.  (1200 such sets were run when this was written.)"""
    from oracle import refogs
    from streamparanumal_b200 import host
    if not refogs.available():
        pytest.skip("oracle/_ref/libref_ogs.so not built here (needs /root/reference)")
    rng = np.random.RandomState(20261018)
    for it in range(120):
        n = int(rng.choice([0, 1, 2, 7, 33, 100, 1000, 5000]))
        nl = max(1, int(n * rng.choice([0.05, 0.3, 0.9, 2.0])))
        ids = rng.randint(1, nl + 1, size=n).astype(np.int64)
        if rng.rand() < 0.3:
            ids = ids * (1 << 33) + rng.randint(0, 3, size=n)
        mode = rng.randint(0, 4)
        kind, unique = host.SIGNED, True
        if mode == 1:
            ids[rng.rand(n) < 0.2] = 0
            neg = rng.rand(n) < 0.3
            ids[neg] = -ids[neg]
        elif mode == 2:
            kind, unique = host.UNSIGNED, False
            ids[rng.rand(n) < 0.1] = 0
        elif mode == 3:
            unique = False   # the caller marks the first copy of every label positive
            seen = set()
            for i, v in enumerate(ids.tolist()):
                if v in seen:
                    ids[i] = -v
                seen.add(v)
            ids[rng.rand(n) < 0.1] = 0
        r = refogs.RefOgs(ids.copy(), kind=kind, unique=unique)
        host.host_lib().spn_host_debug_set(1, 1)
        try:
            o = host.Ogs(ids.copy(), comm, kind=kind, unique=unique, platform=None)
        finally:
            host.host_lib().spn_host_debug_set(1, 0)
        what = f"set {it}: n={n}, labels<={nl}, mode {mode}"
        assert (o.N, o.Ngather, o.NlocalT, o.NlocalP) == (r.N, r.Ngather, r.NlocalT, r.NlocalP), what
        assert np.array_equal(o.ids, r.ids), what + ": signs written back"
        for k, v in r.maps().items():
            if not k.startswith("block"):
                assert np.array_equal(o.array("local_" + k)[: v.size], v), what + f": {k}"
        if o.NgatherGlobal:   # ogs_t::SetupGlobalToLocalMapping (refused, as by the reference, when no label takes part)
            assert np.array_equal(o.global_to_local(), r.global_to_local()), what + ": GlobalToLocal"
        r.close()
        o.close()
