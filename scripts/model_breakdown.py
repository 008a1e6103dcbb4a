"""Per-kernel and host-side breakdown of one StateSpaceGP training + prediction step for a d > 4 kernel (tuning aid)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import __graft_entry__ as g
g.import_package()
from pssgp_b200 import kernels, _lib
from pssgp_b200.model import StateSpaceGP
n = 1_000_000
t, y = bench.make_series(n)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
t_pin, y_pin, q_pin = pin(t[:, None]), pin(y[:, None]), pin((t + 0.002)[:, None])
mean_pin, var_pin = torch.empty((n, 1), dtype=torch.float64).pin_memory(), torch.empty((n, 1), dtype=torch.float64).pin_memory()
which = sys.argv[1] if len(sys.argv) > 1 else "d9"
k = kernels.RBF(1.0, 1.0, order=6, balancing_iter=5) if which == "d6" else kernels.Matern52(1.0, 1.0) + kernels.RBF(1.0, 1.0, order=6, balancing_iter=5)
model = StateSpaceGP((t_pin, y_pin), k, noise_variance=0.1, parallel=True, max_parallel=2 * n)
h = _lib.handle(0)
def sync(): torch.cuda.synchronize()
acc = {}
for it in range(6):
    if it == 3:
        h.set_option("timing", 1); h.timing_report()
    marks = []
    def mark(name):
        sync(); marks.append((name, time.perf_counter()))
    mark("start")
    model.data = (t_pin, y_pin); mark("data H2D")
    ll = model.maximum_log_likelihood_objective(); mark("ll forward")
    grads = torch.autograd.grad(ll, model.trainable_variables); mark("ll backward")
    mean, var = model.predict_f(q_pin, out=(mean_pin, var_pin)); mark("predict_f")
    if it >= 3:
        for (a, ta), (b, tb) in zip(marks[:-1], marks[1:]):
            acc[b] = acc.get(b, 0.0) + (tb - ta) / 3
rep = h.timing_report(); h.set_option("timing", 0)
print({k: round(v * 1e3, 3) for k, v in acc.items()}, "ms; total", round(sum(acc.values()) * 1e3, 3))
for kname, (cnt, ms) in sorted(rep.items(), key=lambda kv: -kv[1][1]):
    print(f"   {kname:28s} {cnt:5d} launches  {ms / 3:8.3f} ms per step")
print("   kernels total", round(sum(v[1] for v in rep.values()) / 3, 3), "ms per step")
