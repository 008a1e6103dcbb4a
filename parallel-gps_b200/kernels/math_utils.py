"""Mirror of the reference's pssgp/kernels/math_utils.py (balance_ss :32-81, solve_lyap_vec :84-120).

``balance_ss`` keeps the reference's contract: the diagonal scaling ``d`` is a constant w.r.t.
autodiff (in the reference it crosses tf.numpy_function, math_utils.py:68).  The scaling itself is
computed by a compiled host routine in the C-ABI library (``pssgp_balance_ss``; the reference's is a
numba-JIT function, math_utils.py:10-29).
"""
import ctypes

import numpy as np
import torch

from .. import _lib


def _balance_d(F, n_iter):
    Fh = np.ascontiguousarray(F.detach().cpu().numpy(), dtype=np.float64)
    d = np.empty((Fh.shape[0],), dtype=np.float64)
    _lib.check(_lib.lib().pssgp_balance_ss(Fh.ctypes.data_as(ctypes.c_void_p), int(Fh.shape[0]), int(n_iter),
                                          d.ctypes.data_as(ctypes.c_void_p)))
    return d


def balance_ss(F, L, H, q, n_iter=5):
    """math_utils.py:32-81."""
    d = torch.as_tensor(_balance_d(F, n_iter), dtype=F.dtype)
    F = F * d[None, :] / d[:, None]
    L = L / d[:, None]
    H = H * d[None, :]
    tmp3 = torch.max(torch.abs(L))
    L = L / tmp3
    q = (tmp3 ** 2) * q
    tmp4 = torch.max(torch.abs(H))
    H = H / tmp4
    q = (tmp4 ** 2) * q
    return F, L, H, q


def _lyap(F, G):
    """X with F X + X F^T = G: compiled host routine (C ABI pssgp_lyap_solve), numpy float64 in and out."""
    Fh = np.ascontiguousarray(F, dtype=np.float64)
    Gh = np.ascontiguousarray(G, dtype=np.float64)
    X = np.empty_like(Fh)
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    _lib.check(_lib.lib().pssgp_lyap_solve(vp(Fh), vp(Gh), int(Fh.shape[0]), vp(X)))
    return X


class _LyapSolve(torch.autograd.Function):
    """X = Lyap(F, G), F X + X F^T = G.  The adjoint is the same solve with F^T: with S from F^T S + S F = gX,
    gG = S and gF = -(S X^T + S^T X) — instead of differentiating through a d^2 x d^2 dense solve and two Kronecker
    products."""

    @staticmethod
    def forward(ctx, F, G):
        X = torch.as_tensor(_lyap(F.detach().numpy(), G.detach().numpy()), dtype=F.dtype)
        ctx.save_for_backward(F.detach(), X)
        return X

    @staticmethod
    def backward(ctx, gX):
        F, X = ctx.saved_tensors
        S = torch.as_tensor(_lyap(F.numpy().T, gX.detach().numpy()), dtype=F.dtype)
        return -(S @ X.T + S.T @ X), S


def solve_lyap_vec(F, L, Q):
    """math_utils.py:84-120:  F P + P F^T + L Q L^T = 0  through the d^2 x d^2 Kronecker system, P = -sym(X)."""
    X = _LyapSolve.apply(F, L @ (Q @ L.T))
    return -0.5 * (X + X.T)
