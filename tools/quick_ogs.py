"""Quick device timing of the ogs kernels on a degree-p hex box (run on the GPU box)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import spn_testutil as util  # noqa: E402
from streamparanumal_b200 import kernels as K  # noqa: E402


def timeit(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def main():
    nx = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    p = int(sys.argv[2]) if len(sys.argv) > 2 else 7
    limits = [K.nodes_per_block() if x == "native" else int(x) for x in sys.argv[3].split(",") if x != "none"] if len(sys.argv) > 3 else [K.nodes_per_block(), 1024, 512]
    t0 = time.time()
    ids = util.hex_ids(nx, nx, nx, p)
    rs, ci = util.csr_from_ids(ids)
    Nt, Ng = ids.size, rs.size - 1
    print(f"mesh nx={nx} p={p}: Nt={Nt} Ng={Ng} host setup {time.time() - t0:.1f}s", flush=True)
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    d_rs, d_ci = d(rs), d(ci)
    v = torch.rand(Nt, dtype=torch.float64, device="cuda")
    gv = torch.empty(Ng, dtype=torch.float64, device="cuda")
    b6 = 4 * (Ng + 1) + 4 * Nt + 8 * Nt + 8 * Ng
    b8 = 4 * (Ng + 1) + 4 * Nt + 16 * Nt
    t0 = time.time()
    plan = K.OgsPlan(rs, d_rs, d_ci, Nt, colIds=None if os.environ.get("OGS_NO_GS_ORDER") else ci)
    print(f"plan: {plan.Ntiles} tiles ({Nt / max(plan.Ntiles, 1):.1f} nnz/tile), {plan.Nirregular} irregular rows, "
          f"{plan.plan_bytes / 1e6:.1f} MB, built in {time.time() - t0:.2f}s", flush=True)
    if os.environ.get("OGS_CHUNK"):
        from streamparanumal_b200 import _lib as _l
        _l.kernel_lib().spn_debug_set(4, int(os.environ["OGS_CHUNK"]))
    tg = timeit(lambda: plan.gather(1, v, gv))
    ts = timeit(lambda: plan.scatter(1, gv, v))
    tgs = timeit(lambda: plan.gather_scatter(1, v), iters=50)
    print(f"tile plan: gather {tg*1e3:.3f} ms {b6/tg*1e-9:7.1f} GB/s | scatter {ts*1e3:.3f} ms {b6/ts*1e-9:7.1f} GB/s | "
          f"gatherScatter {tgs*1e3:.3f} ms {b8/tgs*1e-9:7.1f} GB/s {Ng/tgs*1e-9:6.1f} GDOF/s", flush=True)
    for lim in limits:
        bs = util.greedy_blocks(rs, lim)
        nb = bs.size - 1
        d_bs = d(bs)
        tg = timeit(lambda: K.ogs_gather(nb, 1, d_bs, d_rs, d_ci, v, gv))
        ts = timeit(lambda: K.ogs_scatter(nb, 1, d_bs, d_rs, d_ci, gv, v))
        tgs = timeit(lambda: K.ogs_gather_scatter(nb, 1, d_bs, d_rs, d_ci, d_rs, d_ci, v), iters=50)
        print(f"blocks<= {lim:5d} ({nb} blocks): gather {tg*1e3:.3f} ms {b6/tg*1e-9:7.1f} GB/s | "
              f"scatter {ts*1e3:.3f} ms {b6/ts*1e-9:7.1f} GB/s | gatherScatter {tgs*1e3:.3f} ms {b8/tgs*1e-9:7.1f} GB/s "
              f"{Ng/tgs*1e-9:6.1f} GDOF/s", flush=True)


if __name__ == "__main__":
    main()
