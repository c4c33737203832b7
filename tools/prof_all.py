"""Launch every hot-path kernel a few times at a roofline-relevant size (run under ncu on the
GPU box: tools/gpu_pass_1gpu.sh).  Sizes: streaming kernels N=2^26 FP64 (BS5 config 1), ogs on a
degree-7 hex box of nx^3 elements (default 48; the headline 64^3 mesh under bench.py)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import spn_testutil as util  # noqa: E402
from streamparanumal_b200 import kernels as K  # noqa: E402


def main():
    nx = int(sys.argv[1]) if len(sys.argv) > 1 else 48
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    n = 1 << 26
    a = torch.rand(n, dtype=torch.float64, device="cuda")
    b = torch.rand(n, dtype=torch.float64, device="cuda")
    o = torch.zeros(1, dtype=torch.float64, device="cuda")
    for _ in range(reps):
        K.copy(n, a, b)
        K.axpby(n, 1.0, a, 1.0, b)
        K.norm2(n, a, o)
        K.dot(n, a, b, o)
        K.cg_update(n, a, 1.0, b, o)
    torch.cuda.synchronize()
    del a, b
    ids = util.hex_ids(nx, nx, nx, 7)
    rs, ci = util.csr_from_ids(ids)
    Nt, Ng = ids.size, rs.size - 1
    d_rs, d_ci = torch.from_numpy(rs).cuda(), torch.from_numpy(ci).cuda()
    plan = K.OgsPlan(rs, d_rs, d_ci, Nt, colIds=ci)
    v = torch.rand(Nt, dtype=torch.float64, device="cuda")
    gv = torch.empty(Ng, dtype=torch.float64, device="cuda")
    for _ in range(reps):
        plan.gather(1, v, gv)
        plan.scatter(1, gv, v)
        plan.gather_scatter(1, v)
    torch.cuda.synchronize()
    print(f"prof_all done: n={n} nx={nx} Nt={Nt} Ng={Ng}")


if __name__ == "__main__":
    main()
