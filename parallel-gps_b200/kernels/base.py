"""Mirror of the reference's pssgp/kernels/base.py for the hot path.

* ``ContinuousDiscreteModel`` (:15), ``get_lssm_spec`` (:18-26)
* ``_get_ssm`` (:29-47): the per-time-step discretisation runs in CUDA (C ABI ``pssgp_discretise``):
  ``Fs = expm(dt F)``, ``Qs = Pinf - Fs Pinf Fs^T``.
* ``SDEKernelMixin`` (:50-107): ``get_sde`` / ``get_ssm`` / ``get_spec`` / ``+`` / ``*``
* ``SDESum`` (:129-183), ``SDEProduct`` (:186-244)

The d x d SDE construction (once per hyper-parameter setting) is host logic on torch float64 so
that autograd reaches the kernel hyper-parameters; all per-time-step work is on the GPU.
"""
import abc
from collections import namedtuple
from functools import reduce

import torch

from .. import _arrays as A
from .. import config as pssgp_config
from .. import ops
from ..kalman.base import LGSSM
from .math_utils import balance_ss, solve_lyap_vec

ContinuousDiscreteModel = namedtuple("ContinuousDiscreteModel", ["P0", "F", "L", "H", "Q"])

TensorSpec = namedtuple("TensorSpec", ["shape", "dtype"])


def get_lssm_spec(dim, T):
    """kernels/base.py:18-26 (shapes only; there is no tf.function tracing here)."""
    dtype = pssgp_config.default_float()
    return LGSSM(TensorSpec((dim, dim), dtype), TensorSpec((T, dim, dim), dtype), TensorSpec((T, dim, dim), dtype),
                 TensorSpec((1, dim), dtype), TensorSpec((1, 1), dtype))


def time_steps(ts, t0, dtype, device):
    """dts_k = t_k - t_{k-1} with t_{-1} = t0 (kernels/base.py:34-35), computed on the device."""
    tsd = A.to_device(ts, dtype, device, "ts").reshape(-1)
    prev = torch.cat([torch.full((1,), float(t0), dtype=dtype, device=device), tsd[:-1]])
    return tsd - prev


def _get_ssm(sde, ts, R, t0=0.):
    """kernels/base.py:29-47.  Returns an LGSSM of CUDA tensors."""
    device = A.pick_device(ts)
    dtype = pssgp_config.default_float()
    dts = time_steps(ts, t0, dtype, device)
    F = A.to_device(sde.F, dtype, device, "F")
    Pinf = A.to_device(sde.P0, dtype, device, "Pinf")
    Fs, Qs = ops.discretise(F, Pinf, dts)
    H = A.to_device(sde.H, dtype, device, "H")
    Rd = A.to_device(R, dtype, device, "R").reshape(1, 1)
    return LGSSM(Pinf, Fs, Qs, H, Rd)


class SDEKernelMixin(metaclass=abc.ABCMeta):
    """kernels/base.py:50-107."""

    def __init__(self, t0=0., **_kwargs):
        self.t0 = t0

    @abc.abstractmethod
    def get_sde(self):
        """LTI SDE (P0, F, L, H, Q) of the stationary kernel (torch float64, differentiable)."""

    def get_ssm(self, ts, R, t0=0.):
        return _get_ssm(self.get_sde(), ts, R, t0)

    def __add__(self, other):
        return SDESum([self, other])

    def __mul__(self, other):
        return SDEProduct([self, other])

    @abc.abstractmethod
    def get_spec(self, T):
        return None

    # --- parameter plumbing (gpflow.Module stand-in) ---
    @property
    def parameters(self):
        return []

    @property
    def trainable_variables(self):
        return [p.unconstrained_variable for p in self.parameters if p.trainable]

    def K(self, X, X2=None):
        raise NotImplementedError


class _Combination(SDEKernelMixin):
    def __init__(self, kernels, **kw):
        if not all(isinstance(k, SDEKernelMixin) for k in kernels):
            raise TypeError("can only combine SDE Kernel instances")
        super().__init__(**kw)
        flat = []
        for k in kernels:  # gpflow.kernels.Combination flattens nested combinations of the same type
            flat.extend(k.kernels if type(k) is type(self) else [k])
        self.kernels = flat

    @property
    def parameters(self):
        out = []
        for k in self.kernels:
            for p in k.parameters:
                if not any(p is q for q in out):
                    out.append(p)
        return out


class SDESum(_Combination):
    """kernels/base.py:129-183."""

    def get_spec(self, T):
        dim = 0
        for kernel in self.kernels:
            spec = kernel.get_spec(T)
            if spec is None:
                return None
            dim += spec.P0.shape[-1]
        return get_lssm_spec(dim, T)

    def K(self, X, X2=None):
        return reduce(lambda a, b: a + b, [k.K(X, X2) for k in self.kernels])

    def get_sde(self):
        parts = [k.get_sde() for k in self.kernels]
        Fsum = torch.block_diag(*[p.F for p in parts])
        Lsum = torch.block_diag(*[p.L for p in parts])
        Hsum = torch.cat([p.H for p in parts], dim=1)
        Qsum = torch.block_diag(*[p.Q for p in parts])
        Fb, Lb, Hb, Qb = balance_ss(Fsum, Lsum, Hsum, Qsum, pssgp_config.NUMBER_OF_BALANCING_STEPS)
        Pinf = solve_lyap_vec(Fb, Lb, Qb)
        return ContinuousDiscreteModel(Pinf, Fb, Lb, Hb, Qb)


class SDEProduct(_Combination):
    """kernels/base.py:186-244."""

    def get_spec(self, T):
        dim = 1
        for kernel in self.kernels:
            spec = kernel.get_spec(T)
            if spec is None:
                return None
            dim *= spec.P0.shape[-1]
        return get_lssm_spec(dim, T)

    def K(self, X, X2=None):
        return reduce(lambda a, b: a * b, [k.K(X, X2) for k in self.kernels])

    @staticmethod
    def _combine(s1, s2):
        """Kronecker-sum drift (:199-207), product diffusion (:209-220) and stationary covariance."""
        I1 = torch.eye(s1.F.shape[0], dtype=s1.F.dtype)
        I2 = torch.eye(s2.F.shape[0], dtype=s2.F.dtype)
        F = torch.kron(s1.F, I2) + torch.kron(I1, s2.F)
        g1 = s1.L @ s1.Q @ s1.L.T
        g2 = s2.L @ s2.Q @ s2.L.T
        Q = torch.kron(g1, s2.P0) + torch.kron(s1.P0, g2)
        H = torch.kron(s1.H, s2.H)
        P0 = torch.kron(s1.P0, s2.P0)
        L = torch.eye(Q.shape[0], dtype=Q.dtype)
        return ContinuousDiscreteModel(P0, F, L, H, Q)

    def get_sde(self):
        sdes = [k.get_sde() for k in self.kernels]
        comb = reduce(self._combine, sdes)  # the reference's reduce (:237) only type-checks for two factors
        Fb, Lb, Hb, Qb = balance_ss(comb.F, comb.L, comb.H, comb.Q, pssgp_config.NUMBER_OF_BALANCING_STEPS)
        Pinf = solve_lyap_vec(Fb, Lb, Qb)
        return ContinuousDiscreteModel(Pinf, Fb, Lb, Hb, Qb)
