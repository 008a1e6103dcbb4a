"""Grid log-likelihood timing (configs[4]b): native batched path vs the per-setting Python path, d = 9, N = 1e5."""
import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry
entry.import_package()
from pssgp_b200 import batch, kernels, _lib
import bench
dev = torch.device("cuda", 0)
n = 100_000
t_host, y_host = bench.make_series(n)
ls = np.logspace(-1, 1, 32)
settings = [(a, b) for a in ls for b in ls]
mk = lambda a, b: kernels.Matern52(1.0, float(a)) + kernels.RBF(1.0, float(b), order=6, balancing_iter=5)
data = (torch.as_tensor(t_host[:, None]).to(dev), torch.as_tensor(y_host[:, None]).to(dev))
h = _lib.handle(0)
for lanes, native in ((4, True), (6, True), (8, True)):
    h.set_option("grid_lanes", lanes)
    batch.grid_log_likelihood(mk, settings[:16], data, 0.1, device=dev, native=native)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ll = batch.grid_log_likelihood(mk, settings, data, 0.1, device=dev, native=native)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"native={native} lanes={lanes}: {len(settings) / dt:.0f} settings/s ({dt * 1e3 / len(settings):.3f} ms per setting, {len(settings) * n / dt / 1e6:.0f} M steps/s)  ll_max {float(ll.max()):.6f}", flush=True)
# per-kernel times of the native path
h.set_option("timing", 1); h.timing_report()
batch.grid_log_likelihood(mk, settings[:64], data, 0.1, device=dev)
rep = h.timing_report(); h.set_option("timing", 0)
tot = sum(v[1] for v in rep.values())
for k, (c, ms) in sorted(rep.items(), key=lambda kv: -kv[1][1]):
    print(f"   {k:28s} {c:5d} launches  {ms / 64 * 1e3:8.1f} us per setting")
print(f"   total kernel time per setting {tot / 64 * 1e3:.1f} us")
