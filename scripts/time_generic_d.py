"""Throughput of the generic state-dimension path (d > 4) on synthetic data (tuning aid):
python scripts/time_generic_d.py [N]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import __graft_entry__ as g
g.import_package()
from pssgp_b200 import kernels as PK, ops
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 200_000
dev = torch.device("cuda", 0)
t, y = bench.make_series(N)
cases = [("matern52 d=3", lambda: PK.Matern52(1.0, 1.0)),
         ("rbf6 d=6", lambda: PK.RBF(1.0, 1.0, order=6, balancing_iter=5)),
         ("m52+rbf6 d=9", lambda: PK.Matern52(1.0, 1.0) + PK.RBF(1.0, 1.0, order=6, balancing_iter=5)),
         ("qp3 d=16", lambda: PK.Periodic(PK.SquaredExponential(5.0, 1.0), period=1.0, order=3) * PK.Matern32(0.1, 50.0))]
for name, mk in cases:
    with torch.no_grad():
        sde = mk().get_sde()
    d = sde.F.shape[0]
    F, Pinf, H = sde.F.to(dev).contiguous(), sde.P0.to(dev).contiguous(), sde.H.to(dev).reshape(-1).contiguous()
    R = torch.tensor([0.1], dtype=torch.float64, device=dev)
    td = torch.as_tensor(t).to(dev)
    dts = td - torch.cat([torch.zeros(1, dtype=torch.float64, device=dev), td[:-1]])
    yd = torch.as_tensor(y).to(dev)
    g1 = torch.ones(1, dtype=torch.float64, device=dev)
    Fs, Qs = ops.discretise(F, Pinf, dts)
    def step():
        return ops.pkfs_grad(Pinf, Fs, Qs, H, R, yd, g1)
    for _ in range(2): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    alg = 8 * (12 * d * d + 4 * d + 2)
    print(f"{name:16s} N={N}: {ms:8.3f} ms/step  {N/ms/1e3:8.1f} M steps/s  {alg*N/ms/1e6:7.0f} GB/s algorithmic ({100*alg*N/ms/1e6/6547.8:4.1f}% of HBM)")
