"""world_size>1 on CPU: the C++ host layer's ogs Setup (libspn_host.so: directory round, row
numbering, pairwise exchange plan) run by 2/4/8 real processes, its plan walked with the oracle's
kernels and a gloo exchange (tests/ogs_sim.py), and compared with expectations computed
independently from the global labels:
  * GatherScatter(Sym): bit-exact against "own CSR-ordered partial first, then the other ranks'
    partials in ascending rank" (reference order: libs/ogs/ogsPairwise.cpp:214-232);
  * every label has exactly one positive copy and one owner; sum over ranks of Ngather equals
    the number of distinct labels ((Px*nx*p+1)(Py*ny*p+1)(Pz*nz*p+1));
  * Gather(Trans) puts the fold at the owner, Scatter(NoTrans) returns identity(+)value to every copy;
  * on integer-valued data every result equals the single-rank oracle on the assembled global mesh."""
import os
import sys
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, tmpdir, nx, ny, nz, p, K, dtype_name, kind_unique):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import ogs_sim
        import spn_testutil as util
        from streamparanumal_b200 import host
        dtype = np.dtype(dtype_name)
        comm = host.Comm(rank, world, rendezvous=tmpdir)
        ids, nglobal = host.box_hex_ids(rank, world, nx, ny, nz, p)
        kind, unique = kind_unique
        if kind == host.UNSIGNED:
            ogs = host.Ogs(ids, comm, kind=host.UNSIGNED, unique=False)
        else:
            ogs = host.Ogs(ids, comm, kind=host.SIGNED, unique=unique)
        ids = ogs.ids  # signs rewritten when unique
        plan = ogs_sim.Plan(ogs)
        n = ids.size

        # ---- structural checks -----------------------------------------------------------------
        assert ogs.gather_defined == (kind == host.SIGNED)  # Unsigned: every copy positive -> gather not unique
        tot = comm.allreduce(int(ogs.Ngather))
        if kind == host.SIGNED:
            assert tot == nglobal == ogs.NgatherGlobal
            assert comm.allreduce(int((ids > 0).sum())) == nglobal  # exactly one positive copy per label
            if unique:
                # ... and it is the copy the reference's rule picks (rank-seeded LCG on the directory rank)
                everyone = [None] * world
                dist.all_gather_object(everyone, np.abs(ids))
                assert np.array_equal(ids > 0, reference_positive_pick(everyone, world)[rank]), "positive copies differ from the reference's rule"
        sym = plan.mode[host.SYM]
        assert sym["Nsend"] == sym["Nrecv"]  # Sym: every sharer sends to and hears from every other sharer
        # peer-memory pack tables: the send slots listed per halo row are exactly the slots whose row it is
        for m in (host.SYM, host.NOTRANS, host.TRANS):
            d = plan.mode[m]
            rss, rsd, sd = ogs.array("rowSendStarts", m), ogs.array("rowSendDst", m), ogs.array("slotDst", m)
            assert sd.size == d["Nsend"] and rsd.size == d["Nsend"] and (ogs.NhaloT == 0 or rss.size == ogs.NhaloT + 1)
            for p_, (c, o) in enumerate(zip(d["sendCounts"], d["sendOffsets"])):
                assert np.array_equal(sd[o:o + c], (p_ << 24) | np.arange(c))
            for row in range(ogs.NhaloT):
                mine = sorted(rsd[rss[row]:rss[row + 1]].tolist())
                assert mine == sorted(sd[np.flatnonzero(d["sendIds"] == row)].tolist())

        # ---- values --------------------------------------------------------------------------------
        if np.issubdtype(dtype, np.floating):
            v = util.splitmix64(n * K, 100 + rank).astype(dtype)
        else:
            v = ((np.arange(n * K, dtype=np.int64) * 7919 + rank) % 23 - 11).astype(dtype)
        vi = (((np.arange(n * K, dtype=np.int64) + 3 * rank) % 11) - 5).astype(dtype)  # integer-valued: order-free sums

        # everyone learns every rank's (labels, values) to build the expectations
        gathered = [None] * world
        dist.all_gather_object(gathered, (np.abs(ids), ids > 0, v, vi))

        def partial(r, label_cols, vals):  # CSR-ordered partial of rank r for one label, per k
            acc = np.zeros(K, dtype=dtype)
            for c in label_cols:
                acc = acc + vals[c * K:(c + 1) * K]
            return acc

        cols_of = []  # per rank: label -> ascending columns
        for r in range(world):
            lab = gathered[r][0]
            order = np.argsort(lab, kind="stable")
            d = {}
            for c in order:
                d.setdefault(int(lab[c]), []).append(int(c))
            cols_of.append(d)

        def expected_gs(vals_by_rank, me):
            out = vals_by_rank[me].copy()
            for label, cols in cols_of[me].items():
                acc = np.zeros(K, dtype=dtype) + partial(me, cols, vals_by_rank[me])
                for r in range(world):
                    if r != me and label in cols_of[r]:
                        acc = acc + partial(r, cols_of[r][label], vals_by_rank[r])
                for c in cols:
                    out[c * K:(c + 1) * K] = acc
            return out

        got = ogs_sim.gather_scatter(plan, v, K=K, op=0, trans=host.SYM)
        exp = expected_gs([g[2] for g in gathered], rank)
        assert np.array_equal(got, exp), "GatherScatter(Sym) differs from own-first/ascending-rank fold"

        if kind == host.SIGNED:
            # Gather(Trans) then Scatter(NoTrans) on integer-valued data == GatherScatter(Sym) (exact sums)
            gv = ogs_sim.gather(plan, vi, K=K, op=0, trans=host.TRANS)
            assert gv.size == ogs.Ngather * K
            back = ogs_sim.scatter(plan, gv, np.zeros_like(vi), K=K, trans=host.NOTRANS)
            exp_i = expected_gs([g[3] for g in gathered], rank)
            assert np.array_equal(back, exp_i), "Scatter(NoTrans) o Gather(Trans) != GatherScatter(Sym)"
            assert np.array_equal(ogs_sim.gather_scatter(plan, vi, K=K, op=0, trans=host.SYM), exp_i)
            # sum of all gathered entries over ranks == sum of all inputs
            s_in = comm.allreduce(int(vi.astype(np.int64).sum()))
            s_g = comm.allreduce(int(gv.astype(np.int64).sum()))
            assert s_in == s_g
            # Trans / NoTrans gatherScatter variants: positive copies only on one side
            pos = np.repeat(gathered[rank][1], K)
            gs_n = ogs_sim.gather_scatter(plan, vi, K=K, op=0, trans=host.NOTRANS)  # gather positives -> write all
            exp_n = expected_gs([np.where(np.repeat(g[1], K), g[3], 0).astype(dtype) for g in gathered], rank)
            assert np.array_equal(gs_n, exp_n)
            gs_t = ogs_sim.gather_scatter(plan, vi, K=K, op=0, trans=host.TRANS)    # gather all -> write positives
            assert np.array_equal(gs_t[pos], exp_i[pos]) and np.array_equal(gs_t[~pos], vi[~pos])
        dist.barrier()
    finally:
        dist.destroy_process_group()


def reference_positive_pick(all_abs_ids, world):
    """The reference's choice of the positive copy of every label, restated from libs/ogs/ogsSetup.cpp:292-382 and
    libs/primitives/random.cpp:45-164: every non-zero label travels to directory rank |label| % P, which receives them
    in (source rank, source order), groups them by label in ascending label order and keeps, for its n-th group, member
    rands[n] % size — rands[n] = |int32(X_n)|, X_0 = a * (~0 ^ rank) + c, X_(n+1) = a * X_n + c (mod 2^64) — with the
    members of a group in arrival order (the stable instance of the reference's unstable sort).
    Returns, per rank, a boolean array: entry is the positive copy."""
    A, C, M = 6364136223846793005, 1, (1 << 64) - 1
    positive = [np.zeros(len(a), dtype=bool) for a in all_abs_ids]
    for d in range(world):
        lab, src, pos = [], [], []
        for r in range(world):
            idx = np.flatnonzero((all_abs_ids[r] != 0) & (all_abs_ids[r] % world == d))
            lab.append(all_abs_ids[r][idx]); src.append(np.full(idx.size, r)); pos.append(idx)
        lab, src, pos = np.concatenate(lab), np.concatenate(src), np.concatenate(pos)
        order = np.argsort(lab, kind="stable")
        lab, src, pos = lab[order], src[order], pos[order]
        starts = np.flatnonzero(np.r_[True, lab[1:] != lab[:-1]]) if lab.size else np.zeros(0, dtype=np.int64)
        ends = np.r_[starts[1:], lab.size]
        x = (A * (M ^ d) + C) & M
        for s, e in zip(starts, ends):
            v = x & 0xffffffff
            v = v - (1 << 32) if v >= (1 << 31) else v   # int32(x)
            m = abs(v) % (e - s)
            positive[src[s + m]][pos[s + m]] = True
            x = (A * x + C) & M
    return positive


def _free_port():
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


CASES = [
    # world, nx, ny, nz, p, K, dtype, (kind, unique)
    (2, 2, 2, 2, 3, 1, "float64", (1, True)),
    (2, 3, 2, 1, 2, 3, "float32", (1, True)),
    (4, 2, 2, 2, 2, 1, "float64", (1, True)),
    (8, 2, 2, 2, 2, 1, "float64", (1, True)),
    (8, 1, 2, 1, 3, 2, "int64", (1, True)),
    (3, 2, 1, 2, 1, 1, "int32", (1, True)),
    (2, 2, 2, 2, 2, 1, "float64", (0, False)),
]


@pytest.mark.parametrize("world,nx,ny,nz,p,K,dtype,kind_unique", CASES)
def test_multirank_plan(world, nx, ny, nz, p, K, dtype, kind_unique):
    from streamparanumal_b200 import _lib
    if not os.path.exists(_lib.HOST_LIB):
        pytest.fail("libspn_host.so is missing: run `make`")
    with tempfile.TemporaryDirectory(prefix="spn_rdv_") as tmpdir:
        mp.spawn(_worker, args=(world, _free_port(), tmpdir, nx, ny, nz, p, K, dtype, kind_unique), nprocs=world, join=True)
