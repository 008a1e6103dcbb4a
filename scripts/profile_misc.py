"""One invocation of the sequential, merge and grid kernels for ncu (profiles/): python scripts/profile_misc.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import __graft_entry__ as g
g.import_package()
from pssgp_b200 import kernels as PK, ops
dev = torch.device("cuda", 0)
n = 512
t, y = bench.make_series(n)
for k in (PK.Matern52(1., 1.), PK.RBF(1., 1., order=6, balancing_iter=5)):
    with torch.no_grad():
        ssm = k.get_ssm(torch.as_tensor(t[:, None]).to(dev), torch.tensor([[0.1]], dtype=torch.float64, device=dev))
    H, R = ssm.H.reshape(-1).contiguous(), ssm.R.reshape(-1).contiguous()
    B = 8192
    ys = torch.as_tensor(y).to(dev)[None, :].repeat(B, 1) + 0.01 * torch.randn(B, n, dtype=torch.float64, device=dev)
    o = ops.kf(ssm.P0, ssm.Fs, ssm.Qs, H, R, ys, want_predicted=True)
    ops.ks(ssm.Fs, o[0], o[1], o[3], o[4])
    del o, ys
N = 1_000_000
t, y = bench.make_series(N)
ops.merge_queries(torch.as_tensor(t).to(dev), torch.as_tensor(y).to(dev), torch.as_tensor(t + 0.002).to(dev))
torch.cuda.synchronize()
