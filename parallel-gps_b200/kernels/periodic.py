"""Periodic kernel in state-space form: mirrors pssgp/kernels/periodic.py."""
import math

import numpy as np
import torch
from scipy.special import comb, factorial

from ..params import Parameter
from .base import ContinuousDiscreteModel, SDEKernelMixin, get_lssm_spec
from .matern import DT, _Stationary


class SquaredExponential(_Stationary):
    """Parameter carrier for Periodic (gpflow.kernels.SquaredExponential in the reference)."""
    state_dim = None

    def K(self, X, X2=None):
        return self.variance.value * torch.exp(-0.5 * self._scaled_dist(X, X2) ** 2)

    def get_sde(self):
        raise NotImplementedError("use pssgp_b200.kernels.RBF for a state-space squared-exponential kernel")


def _get_offline_coeffs(N):
    """periodic.py:18-38: hyper-parameter independent coefficients b(K,J), K and 1/K!."""
    r = np.arange(0, N + 1)
    J, K = np.meshgrid(r, r)
    div_facto_K = 1 / factorial(K)
    b = 2 * comb(K, np.floor((K - J) / 2) * (J <= K)) / (1 + (J == 0)) * (J <= K) * (np.mod(K - J, 2) == 0)
    return b, K, div_facto_K


class Periodic(SDEKernelMixin):
    def __init__(self, base_kernel, period=1.0, **kwargs):
        assert isinstance(base_kernel, SquaredExponential), "Only SquaredExponential is supported at the moment"
        self._order = kwargs.pop("order", 6)
        super().__init__(**kwargs)
        self.base_kernel = base_kernel
        self.period = Parameter(period, name="period")

    @property
    def parameters(self):
        return self.base_kernel.parameters + [self.period]

    def get_spec(self, T):
        return get_lssm_spec(2 * (self._order + 1), T)

    def K(self, X, X2=None):
        X = torch.as_tensor(X, dtype=DT).reshape(-1)
        X2 = X if X2 is None else torch.as_tensor(X2, dtype=DT).reshape(-1)
        r = math.pi * torch.abs(X[:, None] - X2[None, :]) / self.period.value
        return self.base_kernel.variance.value * torch.exp(-0.5 * (torch.sin(r) / self.base_kernel.lengthscales.value) ** 2)

    def get_sde(self):
        """periodic.py:53-81: harmonic oscillators j*w0 with stationary variances q_j^2, no diffusion."""
        N = self._order
        w0 = 2 * math.pi / self.period.value
        ell = self.base_kernel.lengthscales.value * 2.
        b, K, div_facto_K = (torch.as_tensor(x, dtype=DT) for x in _get_offline_coeffs(N))
        rot = torch.stack([torch.stack([torch.zeros((), dtype=DT), -w0]), torch.stack([w0, torch.zeros((), dtype=DT)])])
        F = torch.kron(torch.diag(torch.arange(0, N + 1, dtype=DT)), rot)
        dim = 2 * (N + 1)
        L = torch.eye(dim, dtype=DT)
        Q = torch.zeros((dim, dim), dtype=DT)
        q2 = b * ell ** (-2 * K) * div_facto_K * torch.exp(-ell ** (-2)) * 2 ** (-K) * self.base_kernel.variance.value
        Pinf = torch.kron(torch.diag(torch.sum(q2, dim=0)), torch.eye(2, dtype=DT))
        H = torch.kron(torch.ones((1, N + 1), dtype=DT), torch.tensor([[1., 0.]], dtype=DT))
        return ContinuousDiscreteModel(Pinf, F, L, H, Q)
