import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import __graft_entry__ as g
g.import_package()
from pssgp_b200 import kernels, ops, _lib
n = 1_000_000
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
fms, fPs, ll, s_sm, s_ad = ops.pkf_with_summaries(Pinf, Fs, Qs, H, R, yd, last_special=False, Fnext=Fs[0].contiguous(), Qnext=Qs[0].contiguous())
def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
for count in (0, 1, 3, 7):
    gathered = torch.cat([s_sm, s_ad]).repeat(8, 1).contiguous()
    na = s_sm.numel()
    def sm():
        if count: ops.set_fold(1, gathered[1:1 + count, :na], count)
        init = None if count else torch.zeros(9, dtype=torch.float64, device=dev)
        ops.pks(Fs, Qs, fms, fPs, last_special=False, Fnext=Fs[0].contiguous(), Qnext=Qs[0].contiguous(), init=init)
    def ad():
        if count: ops.set_fold(2, gathered[1:1 + count, na:], count)
        ops.pkf_backward(Pinf, Fs, Qs, H, R, yd, fms, fPs, g1, first_special=True)
    print(f"count={count}: pks {timeit(sm):.1f} us, pkf_backward {timeit(ad):.1f} us")
