"""One fused step at d > 4 for ncu (tuning aid): python scripts/profile_mid.py <case> [N]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import __graft_entry__ as g
g.import_package()
from pssgp_b200 import kernels as PK, ops
name = sys.argv[1] if len(sys.argv) > 1 else "rbf6"
N = int(float(sys.argv[2])) if len(sys.argv) > 2 else 1_000_000
dev = torch.device("cuda", 0)
t, y = bench.make_series(N)
cases = {"rbf6": lambda: PK.RBF(1.0, 1.0, order=6, balancing_iter=5),
         "m52rbf6": lambda: PK.Matern52(1.0, 1.0) + PK.RBF(1.0, 1.0, order=6, balancing_iter=5),
         "qp3": lambda: PK.Periodic(PK.SquaredExponential(5.0, 1.0), period=1.0, order=3) * PK.Matern32(0.1, 50.0),
         "qp5": lambda: PK.Periodic(PK.SquaredExponential(5.0, 1.0), period=1.0, order=5) * PK.Matern32(0.1, 50.0)}
with torch.no_grad():
    sde = cases[name]().get_sde()
F, Pinf, H = sde.F.to(dev).contiguous(), sde.P0.to(dev).contiguous(), sde.H.to(dev).reshape(-1).contiguous()
R = torch.tensor([0.1], dtype=torch.float64, device=dev)
td = torch.as_tensor(t).to(dev)
dts = td - torch.cat([torch.zeros(1, dtype=torch.float64, device=dev), td[:-1]])
yd = torch.as_tensor(y).to(dev)
g1 = torch.ones(1, dtype=torch.float64, device=dev)
Fs, Qs = ops.discretise(F, Pinf, dts)
for _ in range(2):
    ops.pkfs_grad(Pinf, Fs, Qs, H, R, yd, g1)
torch.cuda.synchronize()
