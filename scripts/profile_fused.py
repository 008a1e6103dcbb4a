"""One device-resident fused filter+smoother+grad step on the bench workload (for ncu)."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import __graft_entry__ as g
g.import_package()
from pssgp_b200 import kernels, ops, _lib

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda", 0)
t, y = bench.make_series(n)
with torch.no_grad():
    sde = kernels.Matern52(1.0, 1.0).get_sde()
F, Pinf, H = sde.F.to(dev).contiguous(), sde.P0.to(dev).contiguous(), sde.H.to(dev).reshape(-1).contiguous()
R = torch.tensor([0.1], dtype=torch.float64, device=dev)
td = torch.as_tensor(t).to(dev)
dts = td - torch.cat([torch.zeros(1, dtype=torch.float64, device=dev), td[:-1]])
yd = torch.as_tensor(y).to(dev)
g1 = torch.ones(1, dtype=torch.float64, device=dev)
Fs, Qs = ops.discretise(F, Pinf, dts)
for _ in range(reps):
    out = ops.pkfs_grad(Pinf, Fs, Qs, H, R, yd, g1)
torch.cuda.synchronize()
print("ll", float(out[0][2]))
