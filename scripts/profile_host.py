"""cProfile of the host side of one training step (tuning aid)."""
import os, sys, cProfile, pstats, io
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import __graft_entry__ as g
g.import_package()
from pssgp_b200 import kernels
from pssgp_b200.model import StateSpaceGP
n = 1_000_000
t, y = bench.make_series(n)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
model = StateSpaceGP((pin(t[:, None]), pin(y[:, None])), kernels.Matern52(1.0, 1.0), noise_variance=0.1, parallel=True)
def step():
    ll = model.maximum_log_likelihood_objective()
    grads = torch.autograd.grad(ll, model.trainable_variables)
    return float(ll)
for _ in range(5): step()
pr = cProfile.Profile(); pr.enable()
for _ in range(50): step()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45); print(s.getvalue()[:9000])
