"""Window-staged gather+scatter against the warp-tile kernel on a degree-p hex box: bit comparison on seeded data
(in place and out of place, four types) and device timing of both (run on the GPU box).
usage: quick_win.py nx [p]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import spn_testutil as util  # noqa: E402
from streamparanumal_b200 import _lib, kernels as K  # noqa: E402


def timeit(fn, iters=50, warm=5):
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
    lib = _lib.kernel_lib()
    t0 = time.time()
    ids = util.hex_ids(nx, nx, nx, p)
    rs, ci = util.csr_from_ids(ids)
    Nt, Ng = ids.size, rs.size - 1
    print(f"mesh nx={nx} p={p}: Nt={Nt} Ng={Ng} host setup {time.time() - t0:.1f}s", flush=True)
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    d_rs, d_ci = d(rs), d(ci)
    plan = K.OgsPlan(rs, d_rs, d_ci, Nt)
    print(f"plan: {plan.Ntiles} tiles, {plan.Nchunks} chunks, far {plan.Nfar} ({plan.Nfar / Nt:.3f} of nnz, max {plan.maxFar} per chunk), "
          f"window columns {plan.NwindowCols} ({plan.NwindowCols / Nt:.3f} of Ncols), {plan.plan_bytes / 1e6:.1f} MB", flush=True)
    b8 = 4 * (Ng + 1) + 4 * Nt + 16 * Nt
    ok = True
    cases = ((torch.float64, K.ADD), (torch.float64, K.MAX), (torch.float32, K.ADD), (torch.int32, K.MIN),
             (torch.int64, K.ADD), (torch.float64, K.MUL))
    if os.environ.get("QUICK_FAST"):
        cases = cases[:1]
    for dtype, op in cases:
        if dtype.is_floating_point:
            v0 = (torch.rand(Nt, dtype=torch.float64, device="cuda") * 2 - 1).to(dtype)
            v0[::97] = -0.0
        else:
            v0 = torch.randint(-1000, 1000, (Nt,), dtype=dtype, device="cuda")
        lib.spn_debug_set(6, 0)
        a = v0.clone(); plan.gather_scatter(1, a, op)
        wa = torch.full_like(v0, 7); plan.gather_scatter_to(1, v0, wa, op)
        lib.spn_debug_set(6, 1)
        b = v0.clone(); plan.gather_scatter(1, b, op)
        wb = torch.full_like(v0, 7); plan.gather_scatter_to(1, v0, wb, op)
        torch.cuda.synchronize()
        view = (lambda t: t.view(torch.int64)) if dtype in (torch.float64, torch.int64) else (lambda t: t.view(torch.int32))
        e1, e2 = bool(torch.equal(view(a), view(b))), bool(torch.equal(view(wa), view(wb)))
        print(f"  {dtype} op {op}: in place {'bit-equal' if e1 else 'DIFFERENT'}, out of place {'bit-equal' if e2 else 'DIFFERENT'}", flush=True)
        if not e1:
            bad = (view(a) != view(b)).nonzero().flatten()
            print("    first differing columns:", bad[:10].tolist(), "count", bad.numel())
        ok = ok and e1 and e2
    v = torch.rand(Nt, dtype=torch.float64, device="cuda")
    for rep in range(2):
        lib.spn_debug_set(6, 0)
        tt = timeit(lambda: plan.gather_scatter(1, v))
        lib.spn_debug_set(6, 1)
        tw = timeit(lambda: plan.gather_scatter(1, v))
        print(f"gatherScatter: tiles {tt*1e3:.4f} ms {b8/tt*1e-9:7.1f} GB/s {Ng/tt*1e-9:6.1f} GDOF/s | "
              f"window {tw*1e3:.4f} ms {b8/tw*1e-9:7.1f} GB/s {Ng/tw*1e-9:6.1f} GDOF/s ({b8/tw*1e-9/8000:.3f} of 8 TB/s)", flush=True)
    print("PARITY", "OK" if ok else "FAILED")


if __name__ == "__main__":
    main()
