"""Native batched SDE construction (C ABI ``pssgp_sde_batch``, host C++): the SDE of one covariance structure for many
hyper-parameter settings in one call — what the reference computes one setting at a time in Python/TF with a numba
round trip (kernels/*.py get_sde, math_utils.py balance_ss / solve_lyap_vec, kernels/base.py SDESum / SDEProduct).

``native_spec(kernel)`` translates a kernel object (a base kernel, a product of base kernels, or a sum of those) into
the integer spec of include/pssgp_b200.h plus its hyper-parameter vector; ``sde_batch(spec, params)`` runs the batch.
"""
import ctypes

import numpy as np

from .. import _lib
from .. import config as pssgp_config
from .base import SDEProduct, SDESum
from .matern import Matern12, Matern32, Matern52
from .periodic import Periodic
from .rbf import RBF

MATERN12, MATERN32, MATERN52, RBF_T, PERIODIC = range(5)


def _base(kernel):
    """-> ([type, order, balancing_iter], [hyper-parameters]) of a base kernel, None for anything else."""
    f = lambda p: float(p.value.detach())
    if isinstance(kernel, Matern12):
        return [MATERN12, 0, 0], [f(kernel.variance), f(kernel.lengthscales)]
    if isinstance(kernel, Matern32):
        return [MATERN32, 0, 0], [f(kernel.variance), f(kernel.lengthscales)]
    if isinstance(kernel, Matern52):
        return [MATERN52, 0, int(kernel._balancing_iter)], [f(kernel.variance), f(kernel.lengthscales)]
    if isinstance(kernel, RBF):
        return [RBF_T, int(kernel._order), int(kernel._balancing_iter)], [f(kernel.variance), f(kernel.lengthscales)]
    if isinstance(kernel, Periodic):
        b = kernel.base_kernel
        return [PERIODIC, int(kernel._order), 0], [f(b.variance), f(b.lengthscales), f(kernel.period)]
    return None


def _term(kernel):
    factors = kernel.kernels if isinstance(kernel, SDEProduct) else [kernel]
    spec, params = [len(factors)], []
    for k in factors:
        b = _base(k)
        if b is None:
            return None
        spec += b[0]
        params += b[1]
    return spec, params


def native_spec(kernel):
    """(spec: list of int, params: list of float) of ``kernel`` or None when its structure is outside the native
    grammar (sum of products of base kernels)."""
    terms = kernel.kernels if isinstance(kernel, SDESum) else [kernel]
    spec, params = [int(pssgp_config.NUMBER_OF_BALANCING_STEPS), len(terms)], []
    for t in terms:
        r = _term(t)
        if r is None:
            return None
        spec += r[0]
        params += r[1]
    return spec, params


def sde_dim(spec):
    arr = np.ascontiguousarray(spec, dtype=np.int32)
    d, npar = ctypes.c_int(0), ctypes.c_int(0)
    _lib.check(_lib.lib().pssgp_sde_dim(arr.ctypes.data_as(ctypes.c_void_p), int(arr.size), ctypes.byref(d),
                                       ctypes.byref(npar)))
    return d.value, npar.value


def sde_batch(spec, params, nthreads=0):
    """params [B, P] float64 (per factor in spec order: variance, lengthscale[, period]).
    -> F [B,d,d], Pinf [B,d,d], H [B,d] (numpy float64): balanced drift, stationary covariance, measurement row."""
    arr = np.ascontiguousarray(spec, dtype=np.int32)
    d, npar = sde_dim(arr)
    params = np.ascontiguousarray(np.atleast_2d(np.asarray(params, dtype=np.float64)))
    if params.shape[1] != npar:
        raise ValueError(f"params must be [B,{npar}] for this spec, got {params.shape}")
    B = params.shape[0]
    F, Pinf, H = np.empty((B, d, d)), np.empty((B, d, d)), np.empty((B, d))
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    _lib.check(_lib.lib().pssgp_sde_batch(vp(arr), int(arr.size), B, vp(params), npar, vp(F), vp(Pinf), vp(H),
                                         int(nthreads)))
    return F, Pinf, H
