"""Matern-1/2, 3/2, 5/2 kernels with SDE forms: mirrors pssgp/kernels/matern/{common,matern12,matern32,matern52}.py."""
import math

import torch

from .. import config as pssgp_config
from ..params import Parameter
from .base import ContinuousDiscreteModel, SDEKernelMixin, get_lssm_spec
from .math_utils import balance_ss, solve_lyap_vec

DT = torch.float64


def get_matern_sde(variance, lengthscales, d):
    """matern/common.py:26-52: companion-form drift with lambda = sqrt(2d-1)/ell and white-noise density q."""
    lam = math.sqrt(2 * d - 1) / lengthscales
    binom = torch.tensor([math.comb(d, k) for k in range(d)], dtype=DT)
    powers = torch.stack([lam ** (d - k) for k in range(d)])
    last_row = -(binom * powers)
    F = torch.zeros(d, d, dtype=DT)
    if d > 1:
        F = F + torch.diag(torch.ones(d - 1, dtype=DT), diagonal=1)
    F = torch.cat([F[:-1], last_row.reshape(1, d)], dim=0)
    L = torch.zeros(d, 1, dtype=DT)
    L[d - 1, 0] = 1.
    H = torch.zeros(1, d, dtype=DT)
    H[0, 0] = 1.
    q = (2 * lam) ** (2 * d - 1) * variance * math.factorial(d - 1) ** 2 / math.factorial(2 * d - 2)
    return F, L, H, q.reshape(1, 1)


class _Stationary(SDEKernelMixin):
    state_dim = None

    def __init__(self, variance=1.0, lengthscales=1.0, **kwargs):
        super().__init__(**kwargs)
        self.variance = Parameter(variance, name="variance")
        self.lengthscales = Parameter(lengthscales, name="lengthscales")

    @property
    def parameters(self):
        return [self.lengthscales, self.variance]  # gpflow orders variables by attribute name

    def get_spec(self, T):
        return get_lssm_spec(self.state_dim, T)

    def _scaled_dist(self, X, X2):
        X = torch.as_tensor(X, dtype=DT).reshape(-1)
        X2 = X if X2 is None else torch.as_tensor(X2, dtype=DT).reshape(-1)
        return torch.abs(X[:, None] - X2[None, :]) / self.lengthscales.value


class Matern12(_Stationary):
    state_dim = 1

    def K(self, X, X2=None):
        return self.variance.value * torch.exp(-self._scaled_dist(X, X2))

    def get_sde(self):
        """matern12.py:18-23."""
        F, L, H, Q = get_matern_sde(self.variance.value, self.lengthscales.value, 1)
        return ContinuousDiscreteModel(self.variance.value.reshape(1, 1), F, L, H, Q)


class Matern32(_Stationary):
    state_dim = 2

    def K(self, X, X2=None):
        r = math.sqrt(3.) * self._scaled_dist(X, X2)
        return self.variance.value * (1. + r) * torch.exp(-r)

    def get_sde(self):
        """matern32.py:20-28 (closed-form stationary covariance)."""
        v, ell = self.variance.value, self.lengthscales.value
        F, L, H, Q = get_matern_sde(v, ell, 2)
        lam = math.sqrt(3) / ell
        return ContinuousDiscreteModel(torch.diag(torch.stack([v, lam ** 2 * v])), F, L, H, Q)


class _Matern52SDE(torch.autograd.Function):
    """Closed-form fast path of Matern52.get_sde (same values as the generic path below to rounding; tested against
    it): every entry of the balanced SDE is a monomial c * variance^a * lambda^p with lambda = sqrt(5) / lengthscale once
    the balancing vector d is fixed — and d IS a constant w.r.t. differentiation in the reference (it crosses
    tf.numpy_function, math_utils.py:68).  So the forward is a handful of numpy operations and the backward two inner
    products, instead of ~60 torch operations with their autograd graph (0.4 ms + 0.6 ms per training step on the host)."""

    @staticmethod
    def forward(ctx, variance, lengthscales, n_iter):
        import numpy as np
        from .math_utils import _balance_d
        v, ell = float(variance), float(lengthscales)
        lam = math.sqrt(5.0) / ell
        F = np.array([[0.0, 1.0, 0.0], [0.0, 0.0, 1.0], [-lam ** 3, -3.0 * lam ** 2, -3.0 * lam]])
        pF = np.array([[0, 0, 0], [0, 0, 0], [3, 2, 1]], dtype=np.float64)           # power of lambda per entry of F
        q = (2.0 * lam) ** 5 * v * 4.0 / 24.0                                            # (2 lam)^5 v (2!)^2 / 4!
        P = v * np.array([[1.0, 0.0, -lam ** 2 / 3.0], [0.0, lam ** 2 / 3.0, 0.0], [-lam ** 2 / 3.0, 0.0, lam ** 4]])
        pP = np.array([[0, 0, 2], [0, 2, 0], [2, 0, 4]], dtype=np.float64)
        d = _balance_d(torch.as_tensor(F), int(n_iter))                                  # constant w.r.t. autodiff
        Fb = F * d[None, :] / d[:, None]
        L1 = np.array([[0.0], [0.0], [1.0]]) / d[:, None]
        H1 = np.array([[1.0, 0.0, 0.0]]) * d[None, :]
        t3 = np.max(np.abs(L1))
        t4 = np.max(np.abs(H1))
        Lb, Hb = L1 / t3, H1 / t4
        qb = (t3 * t4) ** 2 * q
        Pb = t4 ** 2 * P / np.outer(d, d)
        Pb = 0.5 * (Pb + Pb.T)
        ctx.consts = (v, ell, Fb, pF, Pb, pP, qb)
        tt = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=DT)
        return tt(Pb), tt(Fb), tt(Lb), tt(Hb), tt(np.array([[qb]]))

    @staticmethod
    def backward(ctx, gP, gF, gL, gH, gQ):
        v, ell, Fb, pF, Pb, pP, qb = ctx.consts
        z = lambda g, like: torch.zeros(like.shape, dtype=DT) if g is None else g
        gP, gF, gQ = z(gP, torch.as_tensor(Pb)).numpy(), z(gF, torch.as_tensor(Fb)).numpy(), z(gQ, torch.zeros(1, 1)).numpy()
        # d entry / d variance = entry / variance (P, q); d entry / d ell = -(p / ell) * entry (lambda = sqrt(5) / ell)
        g_v = float((gP * Pb).sum() / v + gQ[0, 0] * qb / v)
        g_ell = float(-((gF * Fb * pF).sum() + (gP * Pb * pP).sum() + 5.0 * gQ[0, 0] * qb) / ell)
        return torch.tensor(g_v, dtype=DT), torch.tensor(g_ell, dtype=DT), None


class Matern52(_Stationary):
    state_dim = 3

    def __init__(self, variance=1.0, lengthscales=1.0, **kwargs):
        self._balancing_iter = kwargs.pop("balancing_iter", pssgp_config.NUMBER_OF_BALANCING_STEPS)
        super().__init__(variance, lengthscales, **kwargs)

    def K(self, X, X2=None):
        r = math.sqrt(5.) * self._scaled_dist(X, X2)
        return self.variance.value * (1. + r + r ** 2 / 3.) * torch.exp(-r)

    def get_sde(self):
        """matern52.py:21-25 (balanced, Lyapunov-solved)."""
        if pssgp_config.FAST_MATERN_SDE:
            return ContinuousDiscreteModel(*_Matern52SDE.apply(self.variance.value, self.lengthscales.value,
                                                               self._balancing_iter))
        F, L, H, q = get_matern_sde(self.variance.value, self.lengthscales.value, 3)
        Fb, Lb, Hb, Qb = balance_ss(F, L, H, q, n_iter=self._balancing_iter)
        return ContinuousDiscreteModel(solve_lyap_vec(Fb, Lb, Qb), Fb, Lb, Hb, Qb)
