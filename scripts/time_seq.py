"""Timing of the sequential kf / ks kernels (C ABI pssgp_kf / pssgp_ks): one long series (the reference's
parallel=False comparator) and batches of short series (the throughput use), against pkf/pks on the same data."""
import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry
entry.import_package()
from pssgp_b200 import ops, kernels as PK

dev = torch.device("cuda", 0)


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def lgssm(kernel, n, dt=0.004):
    rng = np.random.RandomState(0)
    t = np.cumsum(dt * rng.uniform(0.5, 1.5, n))
    ssm = kernel.get_ssm(torch.as_tensor(t[:, None]).to(dev), torch.tensor([[0.1]], dtype=torch.float64, device=dev))
    y = torch.as_tensor(np.sin(t) + 0.3 * rng.randn(n)).to(dev)
    return ssm, y


for name, k, n in (("matern52 d=3", PK.Matern52(1., 1.), 1_000_000), ("rbf6 d=6", PK.RBF(1., 1., order=6, balancing_iter=5), 100_000),
                   ("m52+rbf6 d=9", PK.Matern52(1., 1.) + PK.RBF(1., 1., order=6, balancing_iter=5), 100_000)):
    with torch.no_grad():
        ssm, y = lgssm(k, n)
    H, R = ssm.H.reshape(-1).contiguous(), ssm.R.reshape(-1).contiguous()
    t_kf = timed(lambda: ops.kf(ssm.P0, ssm.Fs, ssm.Qs, H, R, y, want_predicted=True), 1)
    fms, fPs, ll, mps, Pps = ops.kf(ssm.P0, ssm.Fs, ssm.Qs, H, R, y, want_predicted=True)
    t_ks = timed(lambda: ops.ks(ssm.Fs, fms, fPs, mps, Pps), 1)
    t_p = timed(lambda: ops.pkfs(ssm.P0, ssm.Fs, ssm.Qs, H, R, y, want_ll=True))
    print(f"{name}: single series n={n}: kf {t_kf:.1f} ms ({t_kf * 1e3 / n:.3f} us/step), ks {t_ks:.1f} ms "
          f"({t_ks * 1e3 / n:.3f} us/step); parallel pkfs {t_p:.2f} ms", flush=True)
    # batches of short series, shared LGSSM
    for B, T in ((65536, 64), (8192, 512)):
        if ssm.Fs.shape[1] > 4 and B > 8192:
            B = 8192
        Fs, Qs = ssm.Fs[:T].contiguous(), ssm.Qs[:T].contiguous()
        ys = y[:T][None, :].repeat(B, 1) + 0.01 * torch.randn(B, T, dtype=torch.float64, device=dev)
        t_kf = timed(lambda: ops.kf(ssm.P0, Fs, Qs, H, R, ys, want_predicted=True))
        o = ops.kf(ssm.P0, Fs, Qs, H, R, ys, want_predicted=True)
        t_ks = timed(lambda: ops.ks(Fs, o[0], o[1], o[3], o[4]))
        print(f"    batch {B} x {T}: kf {t_kf:.2f} ms = {B * T / t_kf / 1e3:.1f} M steps/s; ks {t_ks:.2f} ms = "
              f"{B * T / t_ks / 1e3:.1f} M steps/s", flush=True)
        del o, ys
