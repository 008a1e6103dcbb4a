"""Tuning aid: phase durations of K1 (filter reduce) from a -DPSSGP_PHASES build (lib_var/phases.so)."""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["PSSGP_B200_LIB"] = os.path.join(ROOT, "parallel-gps_b200", "lib_var", "phases.so")
sys.path.insert(0, ROOT)
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
Fs, Qs = ops.discretise(F, Pinf, dts)
for _ in range(3):
    ops.pkf(Pinf, Fs, Qs, H, R, yd)
torch.cuda.synchronize()
lib = _lib.lib()
buf = (ctypes.c_ulonglong * (148 * 8))()
lib.pssgp_debug_phases.argtypes = [ctypes.c_void_p, ctypes.c_int]
assert lib.pssgp_debug_phases(buf, 148 * 8) == 0
a = np.array(buf, dtype=np.int64).reshape(148, 8)
t0 = a[:, 0].min()
for s, name in enumerate(["start", "stream end", "cta scan end", "ticket", "mid: loaded", "mid: warp scan", "mid: cross-warp", "mid end"]):
    col = a[:, s]
    col = col[col >= t0] - t0
    if len(col) == 0:
        print(f"{name:14s} (no stamp)")
        continue
    print(f"{name:14s} min {col.min()/1e3:8.2f} us  median {np.median(col)/1e3:8.2f} us  max {col.max()/1e3:8.2f} us  (n={len(col)})")
