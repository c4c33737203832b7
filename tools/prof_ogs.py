"""Launch the plan gather / gather+scatter once at nx^3 degree-7 elements (run under ncu)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import spn_testutil as util
from streamparanumal_b200 import kernels as K
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 64
ids = util.hex_ids(nx, nx, nx, 7)
rs, ci = util.csr_from_ids(ids)
Nt, Ng = ids.size, rs.size - 1
d_rs, d_ci = torch.from_numpy(rs).cuda(), torch.from_numpy(ci).cuda()
plan = K.OgsPlan(rs, d_rs, d_ci, Nt, colIds=None)
v = torch.rand(Nt, dtype=torch.float64, device="cuda")
gv = torch.empty(Ng, dtype=torch.float64, device="cuda")
for _ in range(2):
    plan.gather(1, v, gv)
    plan.gather_scatter(1, v)
torch.cuda.synchronize()
