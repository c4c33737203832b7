"""CPU walk-through of the multi-rank ogs operations over the HOST plan of spn::ogs::ogs_t
(test infrastructure).  Every step is the oracle's restatement of one device launch
(oracle/pyoracle.py) or a torch.distributed (gloo) point-to-point exchange in place of NCCL, in
the order streamparanumal_b200/host/ogs_ops.cpp issues them.  Used by the world_size>1 CPU
tests to check the partition / exchange plan; the GPU multi-rank tests run the real thing."""
import numpy as np
import torch
import torch.distributed as dist

from oracle import pyoracle
from streamparanumal_b200 import host


class Plan:
    def __init__(self, ogs):
        self.ogs = ogs
        a = ogs.array
        self.local = {"T": (a("local_rowStartsT"), a("local_colIdsT")), "N": (a("local_rowStartsN"), a("local_colIdsN"))}
        self.halo = {"T": (a("halo_rowStartsT"), a("halo_colIdsT")), "N": (a("halo_rowStartsN"), a("halo_colIdsN"))}
        self.mode = {}
        for m in (host.SYM, host.NOTRANS, host.TRANS):
            d = {k: a(k, m) for k in ("sendIds", "sendRanks", "sendCounts", "sendOffsets", "recvRanks", "recvCounts",
                                      "recvOffsets", "cmbStarts", "cmbIds", "recvRows")}
            d.update(ogs.exchange_info(m))
            self.mode[m] = d


def exchange(d, sendbuf, K=1):
    """pairwise exchange of the packed send buffer (gloo isend/irecv in place of the NCCL group)"""
    recvbuf = np.zeros(d["Nrecv"] * K, dtype=sendbuf.dtype)
    reqs, keep = [], []
    for r, c, o in zip(d["recvRanks"], d["recvCounts"], d["recvOffsets"]):
        t = torch.zeros(int(c) * K, dtype=torch.from_numpy(recvbuf[:0]).dtype)
        keep.append((t, int(o) * K))
        reqs.append(dist.irecv(t, src=int(r)))
    for r, c, o in zip(d["sendRanks"], d["sendCounts"], d["sendOffsets"]):
        t = torch.from_numpy(np.ascontiguousarray(sendbuf[int(o) * K:(int(o) + int(c)) * K]))
        keep.append((t, None))
        reqs.append(dist.isend(t, dst=int(r)))
    for q in reqs:
        q.wait()
    for t, o in keep:
        if o is not None:
            recvbuf[o:o + t.numel()] = t.numpy()
    return recvbuf


def _rows(csr, nrows=None):
    rs, ci = csr
    return rs if nrows is None else rs[:nrows + 1], ci


def combine(d, own, recv, K, op):
    """spn_ogs_halo_combine_scatter's fold: identity (+) own (row < NrowsP) (+) receive slots in order"""
    nrows, nrowsP = d["Nrows"], d["NrowsP"]
    # as a gather over the concatenated [own rows < NrowsP | recv slots] space
    starts = np.zeros(nrows + 1, dtype=np.int32)
    cols = []
    cs, ci = d["cmbStarts"], d["cmbIds"]
    for r in range(nrows):
        row = ([r] if r < nrowsP else []) + [nrowsP + int(j) for j in ci[cs[r]:cs[r + 1]]]
        cols += row
        starts[r + 1] = starts[r] + len(row)
    space = np.concatenate([own[:nrowsP * K], recv])
    return pyoracle.ogs_gather(starts, np.asarray(cols, dtype=np.int32), space, K=K, op=op)


def gather_scatter(plan, v, K=1, op=0, trans=host.SYM):
    o, d = plan.ogs, plan.mode[trans]
    g = "N" if trans == host.NOTRANS else "T"
    s = "N" if trans == host.TRANS else "T"
    own = pyoracle.ogs_gather(*plan.halo[g], v, K=K, op=op) if o.NhaloT else np.zeros(0, dtype=v.dtype)
    send = pyoracle.ogs_extract(d["sendIds"], own, K=K) if d["Nsend"] else np.zeros(0, dtype=v.dtype)
    recv = exchange(d, send, K)
    out = v.copy()
    if o.NlocalT and getattr(o, "kind", host.SIGNED) != host.HALO:  # a halo_t has no interior operator
        out = pyoracle.ogs_gather_scatter(*plan.local[g], *plan.local[s], out, K=K, op=op)
    if o.NhaloT:
        folded = combine(d, own, recv, K, op)
        rs, ci = plan.halo[s]
        out = pyoracle.ogs_scatter(rs[:d["Nrows"] + 1], ci, folded, out, K=K)
    return out


def gather(plan, v, K=1, op=0, trans=host.TRANS):
    o = plan.ogs
    gv = np.zeros((o.NlocalT + (o.NhaloP if trans == host.TRANS or trans == host.NOTRANS else o.NhaloT)) * K, dtype=v.dtype)
    variant = "N" if trans == host.NOTRANS else "T"
    if trans == host.TRANS:
        d = plan.mode[host.TRANS]
        own = pyoracle.ogs_gather(*plan.halo["T"], v, K=K, op=op) if o.NhaloT else np.zeros(0, dtype=v.dtype)
        send = pyoracle.ogs_extract(d["sendIds"], own, K=K) if d["Nsend"] else np.zeros(0, dtype=v.dtype)
        recv = exchange(d, send, K)
        if d["Nrows"]:
            gv[o.NlocalT * K:] = combine(d, own, recv, K, op)
    elif o.NhaloT:
        rs, ci = plan.halo[variant]
        nrows = o.NhaloP if variant == "N" else o.NhaloT
        gv[o.NlocalT * K:(o.NlocalT + nrows) * K] = pyoracle.ogs_gather(rs[:nrows + 1], ci, v, K=K, op=op)
    if o.NlocalT:
        rs, ci = plan.local[variant]
        nrows = o.NlocalP if variant == "N" else o.NlocalT
        gv[:nrows * K] = pyoracle.ogs_gather(rs[:nrows + 1], ci, v, K=K, op=op)
    return gv


def scatter(plan, gv, v, K=1, trans=host.NOTRANS):
    o = plan.ogs
    out = v.copy()
    variant = "N" if trans == host.TRANS else "T"
    if o.NlocalT:
        rs, ci = plan.local[variant]
        nrows = o.NlocalP if variant == "N" else o.NlocalT
        out = pyoracle.ogs_scatter(rs[:nrows + 1], ci, gv[:nrows * K], out, K=K)
    halo_gv = gv[o.NlocalT * K:]
    if trans == host.NOTRANS:
        d = plan.mode[host.NOTRANS]
        send = pyoracle.ogs_extract(d["sendIds"], halo_gv, K=K) if d["Nsend"] else np.zeros(0, dtype=v.dtype)
        recv = exchange(d, send, K)
        if o.NhaloT:
            folded = combine(d, halo_gv, recv, K, 0)
            out = pyoracle.ogs_scatter(*plan.halo["T"], folded, out, K=K)
    elif o.NhaloT:
        rs, ci = plan.halo[variant]
        nrows = o.NhaloP if variant == "N" else o.NhaloT
        out = pyoracle.ogs_scatter(rs[:nrows + 1], ci, halo_gv[:nrows * K], out, K=K)
    return out
