"""Timing of pssgp_discretise / pssgp_discretise_backward for d > 4 (tuning aid)."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry
entry.import_package()
from pssgp_b200 import ops, kernels as PK
dev = torch.device("cuda", 0)
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
n = int(os.environ.get("N", 200000))
rng = np.random.RandomState(0)
dts = torch.as_tensor(0.004 * rng.uniform(0.5, 1.5, n)).to(dev)
for name, k in (("rbf6 d=6", PK.RBF(1., 1., order=6, balancing_iter=5)), ("m52+rbf6 d=9", PK.Matern52(1., 1.) + PK.RBF(1., 1., order=6, balancing_iter=5)),
                ("qp3 d=16", PK.Periodic(PK.SquaredExponential(5., 1.), period=1., order=3) * PK.Matern32(.1, 50.)),
                ("qp5 d=24", PK.Periodic(PK.SquaredExponential(5., 1.), period=1., order=5) * PK.Matern32(.1, 50.))):
    with torch.no_grad():
        sde = k.get_sde()
    F, P = sde.F.to(dev).contiguous(), sde.P0.to(dev).contiguous()
    d = F.shape[0]
    t_f = timed(lambda: ops.discretise(F, P, dts))
    Fs, Qs = ops.discretise(F, P, dts)
    g1, g2 = torch.randn_like(Fs), torch.randn_like(Qs)
    t_b = timed(lambda: ops.discretise_backward(F, P, dts, Fs, g1, g2))
    bytes_f, bytes_b = 8 * (1 + 2 * d * d) * n, 8 * (1 + 3 * d * d) * n
    print(f"{name}: n={n} fwd {t_f*1e3:.0f} us ({bytes_f/t_f/1e6:.0f} GB/s), bwd {t_b*1e3:.0f} us ({bytes_b/t_b/1e6:.0f} GB/s)", flush=True)
