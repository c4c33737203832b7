"""Multi-rank parity against the reference's OWN code on the CPU.

oracle/_ref/libref_ogs.so is the reference's libs/core + primitives + ogs compiled unchanged; with SPN_RANK / SPN_SIZE /
SPN_RENDEZVOUS set it runs as one rank of P over the MPI stand-in (include/shim/mpi.h), i.e. the reference's ogs_t::Setup
(its directory rounds, LCG pick, pairwise exchange plan) and its host-memory Gather / Scatter / GatherScatter with real
messages between P processes.  Each worker builds BOTH the reference's ogs_t and the native host plan on the same labels:
  * the signs written back into ids (which copy is positive) and every count (Ngather, NlocalT, NlocalP, NhaloT, NhaloP) agree
    (native pick order 1 = the serial reference build's);
  * the local operator's maps agree bit for bit;
  * GatherScatter in all three modes, Gather(Trans) and Scatter(NoTrans) of the native plan — walked by the oracle kernels with
    a gloo exchange (tests/ogs_sim.py), the same arithmetic the device kernels run — equal the reference's results bit for bit
    on uniform(-1,1) data: same partials, same fold order across ranks.
Skipped where the library is absent (no /root/reference at build time)."""
import os
import sys
import tempfile

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, tmpdir, nx, ny, nz, p, element="Hex"):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), SPN_RANK=str(rank), SPN_SIZE=str(world),
                      SPN_RENDEZVOUS=tmpdir, OMP_NUM_THREADS="2")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import ogs_sim
        import spn_testutil as util
        from oracle import refogs
        from streamparanumal_b200 import host
        comm = host.Comm(rank, world, rendezvous=tmpdir)
        if element == "Hex":
            ids, nglobal = host.box_hex_ids(rank, world, nx, ny, nz, p)
        elif element.startswith("Random"):
            # no mesh: nx nodes per rank (ragged: rank r gets nx * (r + 1) / 2) over a common pool of ny labels, so labels have
            # several copies on one rank AND on several ranks; every 11th node masked; seed p
            rng = np.random.RandomState(1000 * p + rank)
            ids = rng.randint(1, ny + 1, size=nx * (rank + 1) // 2).astype(np.int64)
            ids[::11] = 0
            mine = [None] * world
            dist.all_gather_object(mine, np.unique(ids[ids != 0]))
            nglobal = np.unique(np.concatenate(mine)).size
        else:
            # the reference's own partitioned Tri / Quad / Tet box mesh, built by its mesh library across the P ranks
            ids = np.abs(refogs.mesh_box_ids(element, nx, ny, nz, p))
            mine = [None] * world
            dist.all_gather_object(mine, np.unique(ids))
            nglobal = np.unique(np.concatenate(mine)).size
        # the reference, as rank `rank` of `world`
        ref = refogs.RefOgs(ids, kind=refogs.SIGNED, unique=True)
        # the native plan, copies of a label in the order the serial reference build leaves them
        host.host_lib().spn_host_debug_set(1, 1)
        ogs = host.Ogs(ids.copy(), comm, kind=host.SIGNED, unique=True, platform=None)
        host.host_lib().spn_host_debug_set(1, 0)
        assert np.array_equal(ogs.ids, ref.ids), "positive copies differ from the reference's at P > 1"
        assert (ogs.N, ogs.Ngather, ogs.NlocalT, ogs.NlocalP, ogs.NhaloT, ogs.NhaloP) == \
               (ref.N, ref.Ngather, ref.NlocalT, ref.NlocalP, ref.NhaloT, ref.NhaloP)
        assert ogs.NgatherGlobal == ref.NgatherGlobal == nglobal
        for k, v in ref.maps().items():
            if k.startswith("block"):
                continue
            assert np.array_equal(ogs.array("local_" + k)[: v.size], v), f"local {k} differs from the reference's"
        assert np.array_equal(ogs.global_to_local(), ref.global_to_local()), "SetupGlobalToLocalMapping differs from the reference's"
        plan = ogs_sim.Plan(ogs)
        v = util.splitmix64(ids.size, 300 + rank)
        for tname, trans in (("Sym", host.SYM), ("NoTrans", host.NOTRANS), ("Trans", host.TRANS)):
            exp = ref.gather_scatter(v, 1, refogs.ADD, trans)          # MPI messages between the P reference ranks
            got = ogs_sim.gather_scatter(plan, v, K=1, op=host.ADD, trans=trans)
            assert np.array_equal(got.view(np.int64), exp.view(np.int64)), f"GatherScatter({tname}) differs from the reference's"
        # vector-valued fields, the other element types and reduction ops
        for K, dt, op in ((3, np.float32, refogs.MAX), (2, np.int64, refogs.ADD), (2, np.int32, refogs.MIN), (3, np.float64, refogs.MUL)):
            if np.issubdtype(dt, np.floating):
                w = util.splitmix64(ids.size * K, 900 + rank).astype(dt)
            else:
                w = ((np.arange(ids.size * K, dtype=np.int64) * 7919 + rank) % 23 - 11).astype(dt)
            for trans in (host.SYM, host.TRANS):
                exp = ref.gather_scatter(w, K, op, trans)
                got = ogs_sim.gather_scatter(plan, w, K=K, op=op, trans=trans)
                assert np.array_equal(got.view(np.uint8), exp.view(np.uint8)), f"GatherScatter K={K} {dt.__name__} op={op} trans={trans}"
        gv_ref = ref.gather(v, 1, refogs.ADD, refogs.TRANS)[: ref.Ngather]
        gv = ogs_sim.gather(plan, v, K=1, op=host.ADD, trans=host.TRANS)
        assert np.array_equal(gv[: ogs.Ngather].view(np.int64), gv_ref.view(np.int64)), "Gather(Trans) differs from the reference's"
        gq = util.splitmix64(ogs.Ngather, 400 + rank)
        back_ref = ref.scatter(gq, np.zeros(ids.size), 1, refogs.NOTRANS)
        back = ogs_sim.scatter(plan, gq, np.zeros(ids.size), K=1, trans=host.NOTRANS)
        assert np.array_equal(back.view(np.int64), back_ref.view(np.int64)), "Scatter(NoTrans) differs from the reference's"
        dist.barrier()
        ref.close()
        ogs.close()
        comm.close()
    finally:
        dist.destroy_process_group()


def _halo_worker(rank, world, port, tmpdir, K, dtype_name):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), SPN_RANK=str(rank), SPN_SIZE=str(world),
                      SPN_RENDEZVOUS=tmpdir, OMP_NUM_THREADS="2")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import ogs_sim
        import test_halo_cpu
        import spn_testutil as util
        from oracle import refogs
        from streamparanumal_b200 import host
        dtype = np.dtype(dtype_name)
        comm = host.Comm(rank, world, rendezvous=tmpdir)
        ids = test_halo_cpu._ids(rank, world)
        ref = refogs.RefHalo(ids)                     # the reference's halo_t as rank `rank` of `world`
        halo = host.Ogs(ids.copy(), comm, kind=host.HALO, unique=False)
        assert halo.NhaloT - halo.NhaloP == ref.Nhalo
        plan = ogs_sim.Plan(halo)
        if np.issubdtype(dtype, np.floating):
            v = util.splitmix64(ids.size * K, 700 + rank).astype(dtype)
        else:
            v = ((np.arange(ids.size * K, dtype=np.int64) * 7919 + rank) % 23 - 11).astype(dtype)
        got = ogs_sim.gather_scatter(plan, v, K=K, op=host.ADD, trans=host.NOTRANS)
        assert np.array_equal(got.view(np.uint8), ref.exchange(v, K).view(np.uint8)), "halo_t::Exchange differs from the reference's"
        got = ogs_sim.gather_scatter(plan, v, K=K, op=host.ADD, trans=host.TRANS)
        assert np.array_equal(got.view(np.uint8), ref.combine(v, K).view(np.uint8)), "halo_t::Combine differs from the reference's"
        dist.barrier()
        ref.close()
        halo.close()
        comm.close()
    finally:
        dist.destroy_process_group()


def _free_port():
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,nx,ny,nz,p", [(2, 3, 2, 2, 3), (4, 2, 2, 2, 2), (8, 2, 2, 2, 2), (3, 2, 1, 2, 4)])
def test_native_plan_equals_reference_ogs_at_p_ranks(world, nx, ny, nz, p):
    from oracle import refogs
    if not refogs.available():
        pytest.skip("oracle/_ref/libref_ogs.so not built here (needs /root/reference)")
    with tempfile.TemporaryDirectory(prefix="spn_rdv_") as tmpdir:
        mp.spawn(_worker, args=(world, _free_port(), tmpdir, nx, ny, nz, p), nprocs=world, join=True)


@pytest.mark.parametrize("world,element,nx,ny,nz,p", [(2, "Tet", 2, 2, 2, 3), (4, "Tri", 3, 3, 1, 4), (2, "Quad", 3, 2, 1, 5)])
def test_native_plan_equals_reference_ogs_on_its_other_meshes(world, element, nx, ny, nz, p):
    """same pins on the labels of the reference's partitioned Tri / Quad / Tet box meshes (its mesh library running on the P ranks)"""
    from oracle import refogs
    if not refogs.available():
        pytest.skip("oracle/_ref/libref_ogs.so not built here (needs /root/reference)")
    with tempfile.TemporaryDirectory(prefix="spn_rdv_") as tmpdir:
        mp.spawn(_worker, args=(world, _free_port(), tmpdir, nx, ny, nz, p, element), nprocs=world, join=True)


@pytest.mark.parametrize("world,nodes,labels,seed", [(3, 400, 150, 1), (5, 300, 700, 2), (2, 50, 400, 3), (4, 1000, 60, 4)])
def test_native_plan_equals_reference_ogs_on_random_labels(world, nodes, labels, seed):
    """unstructured sharing: labels with several copies per rank on several ranks, ranks of different sizes, labels held by
    one rank only, masked nodes — 2 to 5 ranks (the Pairwise method the capi asks for works at any rank count)"""
    from oracle import refogs
    if not refogs.available():
        pytest.skip("oracle/_ref/libref_ogs.so not built here (needs /root/reference)")
    with tempfile.TemporaryDirectory(prefix="spn_rdv_") as tmpdir:
        mp.spawn(_worker, args=(world, _free_port(), tmpdir, nodes, labels, 0, seed, "Random"), nprocs=world, join=True)


@pytest.mark.parametrize("world,K,dtype", [(2, 1, "float64"), (3, 2, "float64"), (4, 1, "int32"), (3, 3, "float32")])
def test_native_halo_equals_reference_halo_t(world, K, dtype):
    """halo_t::Exchange / Combine (libs/ogs/ogsHalo.cpp) of the reference, P real processes, against the native Kind-Halo plan"""
    from oracle import refogs
    if not refogs.available():
        pytest.skip("oracle/_ref/libref_ogs.so not built here (needs /root/reference)")
    with tempfile.TemporaryDirectory(prefix="spn_rdv_") as tmpdir:
        mp.spawn(_halo_worker, args=(world, _free_port(), tmpdir, K, dtype), nprocs=world, join=True)
