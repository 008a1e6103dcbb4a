"""Sweep of the warps-per-CTA cap of the fragment kernels (tuning aid): python scripts/sweep_mid_warps.py case N w1 w2 ..."""
import os, sys, subprocess
case, N = sys.argv[1], sys.argv[2]
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import __graft_entry__ as g
g.import_package()
from pssgp_b200 import _lib, kernels as PK, ops
dev = torch.device("cuda", 0)
n = int(float(N))
t, y = bench.make_series(n)
cases = {"rbf6": lambda: PK.RBF(1.0, 1.0, order=6, balancing_iter=5),
         "m52rbf6": lambda: PK.Matern52(1.0, 1.0) + PK.RBF(1.0, 1.0, order=6, balancing_iter=5),
         "qp3": lambda: PK.Periodic(PK.SquaredExponential(5.0, 1.0), period=1.0, order=3) * PK.Matern32(0.1, 50.0)}
with torch.no_grad():
    sde = cases[case]().get_sde()
F, Pinf, H = sde.F.to(dev).contiguous(), sde.P0.to(dev).contiguous(), sde.H.to(dev).reshape(-1).contiguous()
R = torch.tensor([0.1], dtype=torch.float64, device=dev)
td = torch.as_tensor(t).to(dev)
dts = td - torch.cat([torch.zeros(1, dtype=torch.float64, device=dev), td[:-1]])
yd = torch.as_tensor(y).to(dev)
g1 = torch.ones(1, dtype=torch.float64, device=dev)
Fs, Qs = ops.discretise(F, Pinf, dts)
h = _lib.handle(0)
for w in [int(x) for x in sys.argv[3:]]:
    h.set_option("mid_warps", w)
    for _ in range(2): ops.pkfs_grad(Pinf, Fs, Qs, H, R, yd, g1)
    h.set_option("timing", 1); h.timing_report()
    for _ in range(3): ops.pkfs_grad(Pinf, Fs, Qs, H, R, yd, g1)
    rep = h.timing_report(); h.set_option("timing", 0)
    tot = sum(v[1] for v in rep.values()) / 3
    print(f"{case} N={n} warps<={w}: {tot:.3f} ms  " + " ".join(f"{k}={v[1] / 3 * 1e3:.0f}" for k, v in rep.items() if k.startswith("mid_")), flush=True)
