"""One discretise forward + backward per case for ncu (tuning aid): python scripts/profile_disc.py [N]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
g.import_package()
from pssgp_b200 import kernels as PK, ops
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 200_000
dev = torch.device("cuda", 0)
rng = np.random.RandomState(0)
dts = torch.as_tensor(0.004 * rng.uniform(0.5, 1.5, N)).to(dev)
for k in (PK.RBF(1., 1., order=6, balancing_iter=5), PK.Matern52(1., 1.) + PK.RBF(1., 1., order=6, balancing_iter=5),
          PK.Periodic(PK.SquaredExponential(5., 1.), period=1., order=3) * PK.Matern32(.1, 50.),
          PK.Periodic(PK.SquaredExponential(5., 1.), period=1., order=5) * PK.Matern32(.1, 50.)):
    with torch.no_grad():
        sde = k.get_sde()
    F, P = sde.F.to(dev).contiguous(), sde.P0.to(dev).contiguous()
    Fs, Qs = ops.discretise(F, P, dts)
    ops.discretise_backward(F, P, dts, Fs, torch.randn_like(Fs), torch.randn_like(Qs))
torch.cuda.synchronize()
