"""Shared helpers for the test-suite (oracle access, problem generators, error metric)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import __graft_entry__ as entry  # noqa: E402

O = entry.import_oracle()


def pkg():
    return entry.import_package()


def rel_err(a, b):
    """max |a-b| / max |b|  (the parity metric: relative to ||reference||_inf)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b.detach().cpu() if isinstance(b, torch.Tensor) else b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-300))


def make_kernel(name):
    if name == "matern12":
        return O.Matern12(1.0, 0.5)
    if name == "matern32":
        return O.Matern32(1.0, 0.5)
    if name == "matern52":
        return O.Matern52(1.0, 0.5)
    if name == "m32xm32":
        return O.Matern32(1.0, 0.5) * O.Matern32(0.7, 1.3)
    if name == "rbf6":
        return O.RBF(1.0, 1.0, order=6, balancing_iter=5)
    if name == "m52+rbf6":
        return O.Matern52(1.0, 1.0) + O.RBF(1.0, 1.0, order=6, balancing_iter=5)
    if name == "m32+m52":
        return O.Matern32(1.0, 0.5) + O.Matern52(1.0, 0.5)
    if name == "m32xm52":
        return O.Matern32(1.0, 0.5) * O.Matern52(1.0, 0.5)
    if name == "qp3":
        return O.Periodic(O.SquaredExponential(5.0, 1.0), period=1.0, order=3) * O.Matern32(0.1, 50.0)
    if name == "qp5":
        return O.Periodic(O.SquaredExponential(5.0, 1.0), period=1.0, order=5) * O.Matern32(0.1, 50.0)
    if name == "qp6":
        return O.Periodic(O.SquaredExponential(5.0, 1.0), period=1.0, order=6) * O.Matern32(0.1, 50.0)
    if name == "periodic2":
        return O.Periodic(O.SquaredExponential(1.0, 0.5), period=0.5, order=2)
    raise KeyError(name)


def make_problem(name, T, seed=0, nan_frac=0.05, span=None, noise=0.1, irregular=True):
    """Synthetic series in the shape of SURVEY.md §8(d): irregular sampling, a fraction of NaNs."""
    rng = np.random.RandomState(seed)
    span = 4.0 if span is None else span
    if irregular:
        dts = (span / T) * rng.uniform(0.5, 1.5, size=T)
        t = np.cumsum(dts)
    else:
        t = np.linspace(0, span, T)
    y = O.obs_noise(O.sinu(t), noise, seed)
    if nan_frac > 0 and T > 2:
        idx = rng.choice(T, size=max(1, int(nan_frac * T)), replace=False)
        y[idx] = np.nan
    cov = make_kernel(name)
    with torch.no_grad():
        ssm = cov.get_ssm(t[:, None], torch.tensor([[noise]], dtype=torch.float64))
    return t, y, cov, ssm


def ssm_numpy(ssm, dtype=np.float64):
    return tuple(np.ascontiguousarray(x.detach().numpy(), dtype=dtype) for x in ssm)
