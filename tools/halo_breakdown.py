"""Where does a multi-rank BS8 step spend its time?  Run under torch.distributed.run with N ranks:
times ogs_t::GatherScatter with phases switched off (spn_host_debug_set: 1 pack/send, 2 interior, 4 combine).
Needs a host library built with the probes compiled in:  make -B CXXFLAGS+=-DSPN_TIMING_PROBES streamparanumal_b200/lib/libspn_host.so
(the shipped library compiles them away: results are wrong with a phase skipped)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import spn_testutil as util  # noqa: E402
from streamparanumal_b200 import host  # noqa: E402


def main():
    nx = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    world, rank, lr = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    comm = host.Comm()
    platform = host.Platform(comm, lr)
    ids, _ = host.box_hex_ids(rank, world, nx, nx, nx, 7)
    ogs = host.Ogs(ids, comm, kind=host.SIGNED, unique=True, platform=platform)
    q = torch.from_numpy(util.splitmix64(ids.size, 3 + rank)).cuda()
    L = host.host_lib()

    def timed(mask, steps=50):
        L.spn_host_debug_set(0, mask)
        for _ in range(5):
            ogs.GatherScatter(q, 1, host.ADD, host.SYM)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            ogs.GatherScatter(q, 1, host.ADD, host.SYM)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        L.spn_host_debug_set(0, 0)
        return float(t.item())

    names = {0: "full step", 5: "interior only", 6: "pack(+send) only", 3: "combine only (no wait target)", 2: "pack + combine (no interior)",
             4: "pack + interior", 1: "interior + combine"}
    from streamparanumal_b200 import _lib
    KL = _lib.kernel_lib()
    for iters in [int(x) for x in os.environ.get("SPN_CHUNKS", "0,4,8,16,32,64").split(",")]:
        KL.spn_debug_set(4, iters)
        t = timed(0)
        t5 = timed(5)
        if rank == 0:
            print(f"chunk iters {iters:3d}: full step {t*1e3:8.1f} us | interior only {t5*1e3:8.1f} us", flush=True)
    KL.spn_debug_set(4, 8)
    for mask in (0, 5, 6, 2, 4, 1, 0):
        if mask == 3:
            continue
        t = timed(mask)
        if rank == 0:
            print(f"mask {mask} [{names[mask]:32s}] {t*1e3:8.1f} us  (peer_exchange={ogs.peer_exchange}, halo rows {ogs.NhaloT})", flush=True)
    if world > 1:
        dist.barrier()
    ogs.close(); platform.close(); comm.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
