"""Throughput and per-kernel times of the fused step for d > 4 (warp-level DMMA path) on synthetic data (tuning aid):
python scripts/time_mid.py [N] [case ...]   cases: rbf6 m52rbf6 qp3 qp5 (default: all)"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import __graft_entry__ as g
g.import_package()
from pssgp_b200 import _lib, kernels as PK, ops
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
want = sys.argv[2:] or ["rbf6", "m52rbf6", "qp3", "qp5"]
dev = torch.device("cuda", 0)
t, y = bench.make_series(N)
cases = {"rbf6": lambda: PK.RBF(1.0, 1.0, order=6, balancing_iter=5),
         "m52rbf6": lambda: PK.Matern52(1.0, 1.0) + PK.RBF(1.0, 1.0, order=6, balancing_iter=5),
         "qp3": lambda: PK.Periodic(PK.SquaredExponential(5.0, 1.0), period=1.0, order=3) * PK.Matern32(0.1, 50.0),
         "qp5": lambda: PK.Periodic(PK.SquaredExponential(5.0, 1.0), period=1.0, order=5) * PK.Matern32(0.1, 50.0)}
h = _lib.handle(0)
for name in want:
    with torch.no_grad():
        sde = cases[name]().get_sde()
    d = sde.F.shape[0]
    n = N if d <= 16 else min(N, 400_000)
    F, Pinf, H = sde.F.to(dev).contiguous(), sde.P0.to(dev).contiguous(), sde.H.to(dev).reshape(-1).contiguous()
    R = torch.tensor([0.1], dtype=torch.float64, device=dev)
    td = torch.as_tensor(t[:n]).to(dev)
    dts = td - torch.cat([torch.zeros(1, dtype=torch.float64, device=dev), td[:-1]])
    yd = torch.as_tensor(y[:n]).to(dev)
    g1 = torch.ones(1, dtype=torch.float64, device=dev)
    Fs, Qs = ops.discretise(F, Pinf, dts)
    def step():
        return ops.pkfs_grad(Pinf, Fs, Qs, H, R, yd, g1)
    for _ in range(2): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record()
    for _ in range(reps): step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    alg = 8 * (12 * d * d + 4 * d + 2)
    print(f"{name:10s} d={d:2d} N={n}: {ms:8.3f} ms/step  {n/ms/1e3:8.1f} M steps/s  {alg*n/ms/1e6:7.0f} GB/s algorithmic "
          f"({100*alg*n/ms/1e6/6547.8:4.1f}% of HBM)", flush=True)
    h.set_option("timing", 1)
    step(); h.timing_report()
    for _ in range(reps): step()
    rep = h.timing_report()
    h.set_option("timing", 0)
    for k, (cnt, tot) in rep.items():
        print(f"      {k:22s} x{cnt // reps:3d}  {tot / reps * 1e3:9.1f} us")
    del Fs, Qs
    torch.cuda.empty_cache()
