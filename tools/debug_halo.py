"""3-rank halo_t Exchange / Combine on devices with mismatch details (debug aid; run on a >= 3-GPU box)."""
import os, sys, tempfile
import numpy as np
import torch
import torch.multiprocessing as mp
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def worker(rank, world, port, tmpdir, K, dtype_name, exchange):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LOCAL_RANK=str(rank), SPN_OGS_EXCHANGE=exchange)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ogs_sim, test_halo_cpu, test_host_ogs_gpu as T
    from streamparanumal_b200 import host
    torch.cuda.set_device(rank)
    dev = f"cuda:{rank}"
    comm = host.Comm(rank, world, rendezvous=tmpdir)
    platform = host.Platform(comm, rank)
    ids = test_halo_cpu._ids(rank, world)
    halo = host.Ogs(ids.copy(), comm, kind=host.HALO, unique=False, platform=platform)
    plan = ogs_sim.Plan(halo)
    v = T._values(ids.size * K, T.NP_OF[dtype_name], 600 + rank)
    for it in range(3):
        for name, fn, trans in (("Exchange", halo.HaloExchange, host.NOTRANS), ("Combine", halo.HaloCombine, host.TRANS)):
            d_v = torch.from_numpy(v.copy()).to(dev)
            fn(d_v, K)
            got = d_v.cpu().numpy()
            exp = ogs_sim.gather_scatter(plan, v, K=K, op=host.ADD, trans=trans)
            bad = np.flatnonzero(got != exp)
            if bad.size:
                print(f"[{exchange} K={K} {dtype_name}] rank {rank} iter {it} {name}: {bad.size} mismatches at entries {bad[:12].tolist()} "
                      f"(ids {ids[bad[:12] // K].tolist()}) got {got[bad[:6]].tolist()} exp {exp[bad[:6]].tolist()} in {v[bad[:6]].tolist()} "
                      f"peer={halo.peer_exchange} NhaloT={halo.NhaloT} NhaloP={halo.NhaloP}", flush=True)
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        print(f"[{exchange} K={K} {dtype_name}] done", flush=True)
    halo.close(); platform.close(); comm.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    import socket
    for exchange in ("p2p", "nccl"):
        for K, dt in ((1, "float64"), (2, "int32"), (2, "float64"), (1, "int32")):
            with socket.socket() as s:
                s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
            with tempfile.TemporaryDirectory(prefix="spn_rdv_") as tmpdir:
                mp.spawn(worker, args=(3, port, tmpdir, K, dt, exchange), nprocs=3, join=True)
