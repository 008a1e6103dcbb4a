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
from . import _lib
from . import config as pssgp_config
from . import ops
from .kernels.base import time_steps


def shard_indices(n_settings, rank=0, world=1):
    """Settings evaluated by `rank`: rank, rank + world, rank + 2 world, ... (balanced to within one)."""
    return list(range(int(rank), int(n_settings), int(world)))


def _native_grid(make_kernel, settings, noise_variance, dts, ys, dtype, device, with_grad=False):
    """Settings whose kernels share one structure inside the native grammar (kernels/native.py): ONE host call builds
    every SDE (C ABI pssgp_sde_batch, all host threads), one copy uploads them, ONE call enqueues discretise + filter
    + log-likelihood of every setting (pssgp_grid_loglik).  None when the structure is outside the grammar.
    with_grad: also the gradient w.r.t. the constrained kernel hyper-parameters (in the order of
    kernels.native.native_spec) and the noise variance -> [B, 1 + P + 1] = (ll | d ll / d params | d ll / d noise)."""
    import numpy as np
    from .kernels import native
    spec0, rows, nvs = None, [], []
    for setting in settings:
        setting = setting if isinstance(setting, (tuple, list)) else (setting,)
        r = native.native_spec(make_kernel(*setting))
        if r is None or (spec0 is not None and r[0] != spec0):
            return None
        spec0 = r[0]
        rows.append(r[1])
        nvs.append(float(noise_variance(*setting)) if callable(noise_variance) else float(noise_variance))
    if with_grad:
        F, Pinf, H, jF, jP, jH = native.sde_batch_jac(spec0, np.asarray(rows))
    else:
        F, Pinf, H = native.sde_batch(spec0, np.asarray(rows))
    B, d = F.shape[0], F.shape[1]
    packed = np.concatenate([F.reshape(-1), Pinf.reshape(-1), H.reshape(-1), np.asarray(nvs)])
    pd = A.to_device(packed, dtype, device, "grid_sde")
    Fd, Pd = pd[:B * d * d], pd[B * d * d:2 * B * d * d]
    Hd, Rd = pd[2 * B * d * d:2 * B * d * d + B * d], pd[2 * B * d * d + B * d:]
    ll = torch.empty((B,), dtype=dtype, device=device)
    h = _lib.handle(device.index)
    if not with_grad:
        _lib.check(_lib.lib().pssgp_grid_loglik(h.ptr, A.dtype_code(ll), B, dts.numel(), d, A.ptr(Fd), A.ptr(Pd), A.ptr(Hd),
                                               A.ptr(Rd), A.ptr(dts), A.ptr(ys), A.ptr(ll), A.stream_ptr(device)))
        return ll
    dF = torch.empty((B, d, d), dtype=dtype, device=device)
    dPinf, dP0 = torch.empty_like(dF), torch.empty_like(dF)
    dH = torch.empty((B, d), dtype=dtype, device=device)
    dR = torch.empty((B,), dtype=dtype, device=device)
    _lib.check(_lib.lib().pssgp_grid_loglik_grad(h.ptr, A.dtype_code(ll), B, dts.numel(), d, A.ptr(Fd), A.ptr(Pd), A.ptr(Hd),
                                                A.ptr(Rd), A.ptr(dts), A.ptr(ys), A.ptr(ll), A.ptr(dF), A.ptr(dPinf),
                                                A.ptr(dP0), A.ptr(dH), A.ptr(dR), A.stream_ptr(device)))
    # chain rule with the Jacobians of the native builder (tiny: on the device in float64)
    jFd = torch.as_tensor(jF, device=device)
    jPd = torch.as_tensor(jP, device=device)
    jHd = torch.as_tensor(jH, device=device)
    gPtot = (dPinf + dP0).to(torch.float64)
    dparams = (torch.einsum("bij,bqij->bq", dF.to(torch.float64), jFd) + torch.einsum("bij,bqij->bq", gPtot, jPd)
               + torch.einsum("bi,bqi->bq", dH.to(torch.float64), jHd))
    return torch.cat([ll.to(torch.float64)[:, None], dparams, dR.to(torch.float64)[:, None]], dim=1)


def grid_log_likelihood(make_kernel, settings, data, noise_variance, rank=0, world=1, dist=None, group=None,
                        device=None, native=True, with_grad=False):
    """ll[i] = log p(y | kernel = make_kernel(*settings[i]), noise_variance) for every i.

    ``make_kernel(*setting)`` returns an SDE kernel (pssgp_b200.kernels); ``noise_variance`` is a float or a
    callable of the setting.  ``data = (ts[T,1], ys[T,1])`` host or device.  With ``world > 1`` (one process per
    GPU) the grid is split across ranks and ``dist.all_gather_into_tensor`` assembles the result on every rank.
    ``native=True`` (default): kernels inside the native grammar (sums of products of Matern / RBF / Periodic) take
    the batched path — SDEs of all settings from one C++ call, all GPU work enqueued from one C call; other kernels,
    or ``native=False``, build each SDE with the Python host layer and make one discretise + pkf call per setting.
    Returns a float64 CPU tensor of shape [len(settings)].
    ``with_grad=True`` (native path only): returns (ll [n], dparams [n, P], dnoise [n]) — the gradient of every
    log-likelihood w.r.t. the constrained kernel hyper-parameters (order of ``kernels.native.native_spec``) and the
    noise variance: gradient-based search over many starting points, or many MCMC chains, from one call per rank.
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
    local = None
    if native and mine:
        res = _native_grid(make_kernel, [settings[i] for i in mine], noise_variance, dts, ys, dtype, device, with_grad)
        if res is not None:
            res = res.to(torch.float64).reshape(len(mine), -1)
            local = torch.full((per, res.shape[1]), float("nan"), dtype=torch.float64, device=device)
            local[:len(mine)] = res
            mine = []
    if local is None:
        if with_grad:
            raise ValueError("grid_log_likelihood(with_grad=True) needs kernels inside the native grammar and native=True")
        local = torch.full((per, 1), float("nan"), dtype=torch.float64, device=device)
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
            local[slot, 0] = ll[0].to(torch.float64)
    width = local.shape[1]
    if world > 1:
        gathered = torch.empty((world * per * width,), dtype=torch.float64, device=device)
        dist.all_gather_into_tensor(gathered, local.reshape(-1).contiguous(), group=group)
        gathered = gathered.reshape(world, per, width)
    else:
        gathered = local.reshape(1, per, width)
    out = torch.empty((n, width), dtype=torch.float64)
    g = gathered.cpu()
    for r in range(world):
        idx = shard_indices(n, r, world)
        out[idx] = g[r, :len(idx)]
    if with_grad:
        return out[:, 0], out[:, 1:-1], out[:, -1]
    return out[:, 0]
