"""RBF kernel in state-space form: mirrors pssgp/kernels/rbf.py."""
import functools
import math

import numpy as np
import torch

from .. import config as pssgp_config
from .base import ContinuousDiscreteModel, get_lssm_spec
from .matern import DT, _Stationary
from .math_utils import balance_ss, solve_lyap_vec


@functools.lru_cache(maxsize=None)
def _get_unscaled_rbf_sde(order=6):
    """rbf.py:14-61: spectral factorisation of the order-`order` Taylor expansion of exp(w^2/2):
    the stable roots of the denominator polynomial give the companion-form drift."""
    coeffs = np.zeros(2 * order + 1)
    coeffs[0::2] = [0.5 ** k / math.factorial(k) for k in range(order, -1, -1)]
    q = math.sqrt(2 * math.pi) / coeffs[-1]
    signed = np.real(coeffs / (1j ** np.arange(coeffs.size - 1, -1, -1, dtype=np.float64)))
    roots = np.roots(signed)
    denom = np.real(np.poly(roots[np.real(roots) < 0]))
    denom = denom / denom[-1]
    gain = 1.0 / denom[0]
    denom = denom / denom[0]
    n = denom.size - 1
    F = np.zeros((n, n))
    F[-1, :] = -denom[:0:-1]
    F[:-1, 1:] = np.eye(n - 1)
    L = np.zeros((n, 1))
    L[-1, 0] = 1.
    H = np.zeros((1, n))
    H[0, 0] = gain
    return F, L, H, q


class RBF(_Stationary):
    def __init__(self, variance=1.0, lengthscales=1.0, **kwargs):
        self._order = kwargs.pop("order", 3)
        self._balancing_iter = kwargs.pop("balancing_iter", pssgp_config.NUMBER_OF_BALANCING_STEPS)
        super().__init__(variance, lengthscales, **kwargs)

    @property
    def state_dim(self):
        return self._order

    def get_spec(self, T):
        return get_lssm_spec(self._order, T)

    def K(self, X, X2=None):
        return self.variance.value * torch.exp(-0.5 * self._scaled_dist(X, X2) ** 2)

    def get_sde(self):
        """rbf.py:78-101."""
        F_, L_, H_, q_ = _get_unscaled_rbf_sde(self._order)
        F = torch.as_tensor(F_, dtype=DT)
        L = torch.as_tensor(L_, dtype=DT)
        H = torch.as_tensor(H_, dtype=DT)
        ell, v = self.lengthscales.value, self.variance.value
        dim = F.shape[0]
        ell_vec = ell ** torch.arange(dim, 0, -1, dtype=DT)
        F = torch.cat([F[:-1], (F[-1, :] / ell_vec).reshape(1, dim)], dim=0)
        H = H / (ell ** dim)
        Q = (v * ell * q_).reshape(1, 1)
        Fb, Lb, Hb, Qb = balance_ss(F, L, H, Q, n_iter=self._balancing_iter)
        Pinf = solve_lyap_vec(Fb, Lb, Qb)
        return ContinuousDiscreteModel(Pinf, Fb, Lb, Hb, Qb.reshape(1, 1))
