"""Batches of independent problems: log-likelihood over a grid of hyper-parameter settings (BASELINE configs[4]:
"batch of 1,024 hyperparameter settings for grid log-lik").

The reference has no batch axis (``get_lssm_spec`` fixes rank-3 ``[T,d,d]``, pssgp/kernels/base.py:18-26; a grid
search there is a Python loop over models).  Independent settings shard trivially: every rank evaluates a strided
slice of the grid on its own GPU with no data-path collective, and one all-gather of the scalars at the end puts the
whole grid on every rank.  Each evaluation is the model's own path: ``get_sde`` -> discretise -> pkf (log-likelihood
only), i.e. ``StateSpaceGP.maximum_log_likelihood_objective`` without the autograd graph.
"""
import torch

from . import _arrays as A
from . import config as pssgp_config
from . import ops
from .kernels.base import time_steps


def shard_indices(n_settings, rank=0, world=1):
    """Settings evaluated by `rank`: rank, rank + world, rank + 2 world, ... (balanced to within one)."""
    return list(range(int(rank), int(n_settings), int(world)))


def grid_log_likelihood(make_kernel, settings, data, noise_variance, rank=0, world=1, dist=None, group=None,
                        device=None):
    """ll[i] = log p(y | kernel = make_kernel(*settings[i]), noise_variance) for every i.

    ``make_kernel(*setting)`` returns an SDE kernel (pssgp_b200.kernels); ``noise_variance`` is a float or a
    callable of the setting.  ``data = (ts[T,1], ys[T,1])`` host or device.  With ``world > 1`` (one process per
    GPU) the grid is split across ranks and ``dist.all_gather_into_tensor`` assembles the result on every rank.
    Returns a float64 CPU tensor of shape [len(settings)].
    """
    A.require_cuda()
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    dtype = pssgp_config.default_float()
    ts = A.to_device(data[0], dtype, device, "grid_ts").reshape(-1)
    ys = A.to_device(data[1], dtype, device, "grid_ys").reshape(-1).contiguous()
    dts = time_steps(ts, 0., dtype, device)
    n = len(settings)
    per = (n + world - 1) // world
    mine = shard_indices(n, rank, world)
    local = torch.full((per,), float("nan"), dtype=torch.float64, device=device)
    with torch.no_grad():
        for slot, i in enumerate(mine):
            setting = settings[i]
            setting = setting if isinstance(setting, (tuple, list)) else (setting,)
            sde = make_kernel(*setting).get_sde()
            nv = noise_variance(*setting) if callable(noise_variance) else noise_variance
            F = A.to_device(sde.F, dtype, device, "grid_F")
            Pinf = A.to_device(sde.P0, dtype, device, "grid_P")
            H = A.to_device(sde.H, dtype, device, "grid_H").reshape(-1)
            R = torch.full((1,), float(nv), dtype=dtype, device=device)
            Fs, Qs = ops.discretise(F, Pinf, dts)
            ll = ops.pkf(Pinf, Fs, Qs, H, R, ys, want_ll=True)[2]
            local[slot] = ll[0].to(torch.float64)
    if world > 1:
        gathered = torch.empty((world * per,), dtype=torch.float64, device=device)
        dist.all_gather_into_tensor(gathered, local, group=group)
        gathered = gathered.reshape(world, per)
    else:
        gathered = local.reshape(1, per)
    out = torch.empty((n,), dtype=torch.float64)
    g = gathered.cpu()
    for r in range(world):
        idx = shard_indices(n, r, world)
        out[idx] = g[r, :len(idx)]
    return out
