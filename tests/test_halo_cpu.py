"""halo_t on CPU (world_size 1-4 over gloo): the Kind-Halo setup of the C++ host layer and the Exchange / Combine
semantics of libs/ogs/ogsHalo.cpp:39-400, walked over the host plan with the oracle's kernels (tests/ogs_sim.py) and
compared with expectations computed independently from the global ids:
  Exchange: every halo copy (id < 0) receives identity (+) the owner's value; the owner of a shared node keeps
            identity (+) its own value; nodes nobody shares are untouched (the reference builds no local operator
            for a halo_t: libs/ogs/ogsSetup.cpp:592-603);
  Combine : the owner receives its own value, then the copies' values by ascending rank
            (libs/ogs/ogsPairwise.cpp:214-232); copies are untouched."""
import os
import sys
import tempfile

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NP_NODES = 6   # nodes per element
ELEMS = 5      # owned elements per rank


def _ids(rank, world):
    """owned elements first (unique positive ids), then copies of the neighbours' boundary elements (negative ids)"""
    own = 1 + (rank * ELEMS + np.arange(ELEMS)[:, None]) * NP_NODES + np.arange(NP_NODES)[None, :]
    halo = []
    if rank > 0:            # last element of the left neighbour
        halo.append(-(1 + ((rank - 1) * ELEMS + ELEMS - 1) * NP_NODES + np.arange(NP_NODES)))
    if rank < world - 1:    # first element of the right neighbour
        halo.append(-(1 + ((rank + 1) * ELEMS) * NP_NODES + np.arange(NP_NODES)))
    if world > 2 and rank == 0:   # a second copy of rank 1's first element's node 0 lives on rank 0 twice
        halo.append(-(1 + (1 * ELEMS) * NP_NODES + np.arange(1)))
    return np.concatenate([own.reshape(-1)] + halo).astype(np.int64)


def _worker(rank, world, port, tmpdir, K, dtype_name):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import ogs_sim
        from streamparanumal_b200 import host
        dtype = np.dtype(dtype_name)
        comm = host.Comm(rank, world, rendezvous=tmpdir)
        ids = _ids(rank, world)
        halo = host.Ogs(ids.copy(), comm, kind=host.HALO, unique=False)
        assert np.array_equal(halo.ids, ids)              # signs are the caller's: nothing is rewritten
        plan = ogs_sim.Plan(halo)
        n = ids.size
        if np.issubdtype(dtype, np.floating):
            import spn_testutil as util
            v = util.splitmix64(n * K, 500 + rank).astype(dtype)
            v[::7] = -0.0
        else:
            v = ((np.arange(n * K, dtype=np.int64) * 7919 + rank) % 23 - 11).astype(dtype)

        everyone = [None] * world
        dist.all_gather_object(everyone, (ids, v))
        owner_value, copies = {}, {}
        for r, (rids, rv) in enumerate(everyone):
            for i, lab in enumerate(rids):
                val = rv[i * K:(i + 1) * K]
                if lab > 0:
                    owner_value[int(lab)] = (r, val)
                else:
                    copies.setdefault(int(-lab), []).append((r, i, val))
        shared = {lab for lab, c in copies.items() if any(r != owner_value[lab][0] for r, _, _ in c)}
        zero = np.zeros(K, dtype=dtype)

        # ---- Exchange ----------------------------------------------------------------------------
        got = ogs_sim.gather_scatter(plan, v, K=K, op=host.ADD, trans=host.NOTRANS)
        exp = v.copy()
        for i, lab in enumerate(ids):
            a = abs(int(lab))
            if a in shared:
                exp[i * K:(i + 1) * K] = zero + (zero + owner_value[a][1])   # pack: identity (+) x; combine: identity (+) that
        assert np.array_equal(got.view(np.uint8), exp.view(np.uint8)), "Exchange"

        # ---- Combine -----------------------------------------------------------------------------
        got = ogs_sim.gather_scatter(plan, v, K=K, op=host.ADD, trans=host.TRANS)
        exp = v.copy()
        for i, lab in enumerate(ids):
            if lab > 0 and int(lab) in shared:
                # own rank's partial first: CSR-ordered fold of the local copies (ascending column), then remote
                # partials (each a CSR-ordered fold on its rank) by ascending rank
                by_rank = {}
                for r, col, val in [(rank, i, v[i * K:(i + 1) * K])] + copies[int(lab)]:
                    by_rank.setdefault(r, []).append((col, val))
                def partial(lst):
                    acc = zero.copy()
                    for _, val in sorted(lst, key=lambda t: t[0]):
                        acc = acc + val
                    return acc
                acc = zero + partial(by_rank.pop(rank))
                for r in sorted(by_rank):
                    acc = acc + partial(by_rank[r])
                exp[i * K:(i + 1) * K] = acc
        assert np.array_equal(got.view(np.uint8), exp.view(np.uint8)), "Combine"
        assert halo.NhaloT - halo.NhaloP == len({abs(int(l)) for l in ids if l < 0 and abs(int(l)) in shared} -
                                                {int(l) for l in ids if l > 0})
        dist.barrier()
        halo.close()
        comm.close()
    finally:
        dist.destroy_process_group()


def _free_port():
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,K,dtype", [(2, 1, "float64"), (3, 2, "float64"), (4, 1, "int32"), (3, 1, "float32")])
def test_halo_exchange_and_combine(world, K, dtype):
    with tempfile.TemporaryDirectory(prefix="spn_rdv_") as tmpdir:
        mp.spawn(_worker, args=(world, _free_port(), tmpdir, K, dtype), nprocs=world, join=True)


def test_halo_needs_exactly_one_owner_per_node():
    sys.path.insert(0, ROOT)
    from streamparanumal_b200 import host
    with tempfile.TemporaryDirectory(prefix="spn_rdv_") as tmpdir:
        comm = host.Comm(0, 1, rendezvous=tmpdir)
        with pytest.raises(RuntimeError, match="unique positive"):
            host.Ogs(np.array([1, 2, -3, -3], dtype=np.int64), comm, kind=host.HALO, unique=False)   # node 3 has no owner
        with pytest.raises(RuntimeError, match="unique positive"):
            host.Ogs(np.array([1, 1, -1], dtype=np.int64), comm, kind=host.HALO, unique=False)        # node 1 has two owners
        # one rank: nothing is shared between ranks, so the halo operator is empty (copies on the owner's own rank are
        # not touched, as in the reference)
        h = host.Ogs(np.array([1, 2, -2, 3], dtype=np.int64), comm, kind=host.HALO, unique=False)
        assert h.NhaloT == 0
        h.close()
        comm.close()
