"""Wall-clock breakdown of StateSpaceGP.predict_f (tuning aid)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import __graft_entry__ as g
g.import_package()
from pssgp_b200 import kernels, ops, _arrays as A
from pssgp_b200.model import StateSpaceGP, _merge_sorted
n = 1_000_000
t, y = bench.make_series(n)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
t_pin, y_pin, q_pin = pin(t[:, None]), pin(y[:, None]), pin((t + 0.002)[:, None])
mean_pin, var_pin = torch.empty((n, 1), dtype=torch.float64).pin_memory(), torch.empty((n, 1), dtype=torch.float64).pin_memory()
model = StateSpaceGP((t_pin, y_pin), kernels.Matern52(1.0, 1.0), noise_variance=0.1, parallel=True, max_parallel=2 * n)
acc = {}
for it in range(8):
    marks = []
    def mark(name):
        torch.cuda.synchronize(); marks.append((name, time.perf_counter()))
    self = model
    mark("start")
    ts, ys = self._data
    dtype, dev = ts.dtype, ts.device
    Xd = A.to_device(q_pin, dtype, dev, "Xnew").reshape(-1); mark("H2D queries")
    K = Xd.shape[0]
    nan_ys = torch.full((K, ys.shape[1]), float("nan"), dtype=dtype, device=dev)
    all_ts, all_ys, all_flags = _merge_sorted(ts.reshape(-1), Xd, (ys, nan_ys),
                                              (torch.zeros(ts.shape[0], dtype=torch.bool, device=dev),
                                               torch.ones(K, dtype=torch.bool, device=dev))); mark("merge")
    with torch.no_grad():
        ssm = self._make_model(all_ts[:, None]); mark("make_model (get_sde + discretise)")
        Hd, Rd = ssm.H.reshape(-1).contiguous(), ssm.R.reshape(-1).contiguous()
        yv = all_ys.reshape(-1).contiguous()
        proj = ops.pkfs(ssm.P0, ssm.Fs, ssm.Qs, Hd, Rd, yv, project=True)[3]; mark("pkfs")
        sel = proj[all_flags]
        mean, var = sel[:, 0:1].contiguous(), sel[:, 1:2].contiguous(); mark("select")
    A.to_host_into(mean, mean_pin); A.to_host_into(var, var_pin); mark("D2H")
    if it >= 3:
        for (a, ta), (b, tb) in zip(marks[:-1], marks[1:]):
            acc[b] = acc.get(b, 0.0) + (tb - ta) / 5
print({k: round(v * 1e3, 3) for k, v in acc.items()}, "ms; total", round(sum(acc.values()) * 1e3, 3))
