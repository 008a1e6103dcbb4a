"""Native batched SDE construction (C ABI ``pssgp_sde_batch``, host C++): the SDE of one covariance structure for many
hyper-parameter settings in one call — what the reference computes one setting at a time in Python/TF with a numba
round trip (kernels/*.py get_sde, math_utils.py balance_ss / solve_lyap_vec, kernels/base.py SDESum / SDEProduct).

``native_spec(kernel)`` translates a kernel object (a base kernel, a product of base kernels, or a sum of those) into
the integer spec of include/pssgp_b200.h plus its hyper-parameter vector; ``sde_batch(spec, params)`` runs the batch.
"""
import ctypes

import numpy as np
import torch

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
    # exact types only: a subclass may override get_sde, and then the native construction would not be its SDE
    if type(kernel) is Matern12:
        return [MATERN12, 0, 0], [f(kernel.variance), f(kernel.lengthscales)]
    if type(kernel) is Matern32:
        return [MATERN32, 0, 0], [f(kernel.variance), f(kernel.lengthscales)]
    if type(kernel) is Matern52:
        return [MATERN52, 0, int(kernel._balancing_iter)], [f(kernel.variance), f(kernel.lengthscales)]
    if type(kernel) is RBF:
        return [RBF_T, int(kernel._order), int(kernel._balancing_iter)], [f(kernel.variance), f(kernel.lengthscales)]
    if type(kernel) is Periodic:
        b = kernel.base_kernel
        return [PERIODIC, int(kernel._order), 0], [f(b.variance), f(b.lengthscales), f(kernel.period)]
    return None


def _base_parameters(kernel):
    """The Parameter objects of a base kernel in the order of its hyper-parameter row."""
    if type(kernel) is Periodic:
        return [kernel.base_kernel.variance, kernel.base_kernel.lengthscales, kernel.period]
    return [kernel.variance, kernel.lengthscales]


def native_parameters(kernel):
    """Parameter objects in the order of native_spec(kernel)[1] (None outside the native grammar)."""
    if native_spec(kernel) is None:
        return None
    out = []
    for t in (kernel.kernels if type(kernel) is SDESum else [kernel]):
        for f in (t.kernels if type(t) is SDEProduct else [t]):
            out += _base_parameters(f)
    return out


def _term(kernel):
    factors = kernel.kernels if type(kernel) is SDEProduct else [kernel]
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
    terms = kernel.kernels if type(kernel) is SDESum else [kernel]
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


def sde_batch_jac(spec, params, nthreads=0):
    """sde_batch plus the Jacobians w.r.t. the (constrained) hyper-parameters: C ABI pssgp_sde_batch_jac.
    -> F, Pinf [B,d,d], H [B,d], dF, dPinf [B,P,d,d], dH [B,P,d]."""
    arr = np.ascontiguousarray(spec, dtype=np.int32)
    d, npar = sde_dim(arr)
    params = np.ascontiguousarray(np.atleast_2d(np.asarray(params, dtype=np.float64)))
    if params.shape[1] != npar:
        raise ValueError(f"params must be [B,{npar}] for this spec, got {params.shape}")
    B = params.shape[0]
    F, Pinf, H = np.empty((B, d, d)), np.empty((B, d, d)), np.empty((B, d))
    dF, dPinf, dH = np.empty((B, npar, d, d)), np.empty((B, npar, d, d)), np.empty((B, npar, d))
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    _lib.check(_lib.lib().pssgp_sde_batch_jac(vp(arr), int(arr.size), B, vp(params), npar, vp(F), vp(Pinf), vp(H), vp(dF),
                                             vp(dPinf), vp(dH), int(nthreads)))
    return F, Pinf, H, dF, dPinf, dH


class NativeSDE(torch.autograd.Function):
    """(F, Pinf, H) of a kernel structure as a differentiable function of its constrained hyper-parameters: forward and
    Jacobian come from ONE call of the native builder (forward-mode duals through balancing, Lyapunov solves and
    Kronecker products); backward contracts the upstream gradients with the stored Jacobians.  Replaces torch autograd
    through the Python get_sde on the training path (``config.NATIVE_SDE``)."""

    @staticmethod
    def forward(ctx, spec, *params):
        row = [float(p.detach()) for p in params]
        F, Pinf, H, dF, dPinf, dH = sde_batch_jac(list(spec), [row])
        ctx.jac = (dF[0], dPinf[0], dH[0])
        dt = params[0].dtype
        tt = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=dt)
        return tt(F[0]), tt(Pinf[0]), tt(H[0]).reshape(1, -1)

    @staticmethod
    def backward(ctx, gF, gP, gH):
        dF, dPinf, dH = ctx.jac
        z = lambda g, like: np.zeros(like.shape[1:]) if g is None else g.detach().cpu().numpy().reshape(like.shape[1:])
        gF, gP, gH = z(gF, dF), z(gP, dPinf), z(gH, dH)
        grads = (dF * gF).sum(axis=(1, 2)) + (dPinf * gP).sum(axis=(1, 2)) + (dH * gH).sum(axis=1)
        return (None,) + tuple(torch.tensor(float(g), dtype=torch.float64) for g in grads)


def native_sde(kernel):
    """(F, Pinf, H) through NativeSDE, or None when the kernel is outside the native grammar."""
    r = native_spec(kernel)
    if r is None:
        return None
    return NativeSDE.apply(tuple(r[0]), *[p.value for p in native_parameters(kernel)])
