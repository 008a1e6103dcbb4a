"""GPU parity of the hand-written adjoint scan and of the discretisation (+ its adjoint) against
torch autograd through the CPU oracle (which differentiates the *parallel* computation like TF does)."""
import numpy as np
import pytest
import torch

from util import O, make_kernel, make_problem, pkg, rel_err

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)


def _ops():
    pkg()
    from pssgp_b200 import ops
    return ops


def sym(X):
    return 0.5 * (X + X.transpose(-1, -2))


@pytest.mark.parametrize("name", ["matern12", "matern32", "matern52", "m32xm32"])
@pytest.mark.parametrize("T", [1, 2, 40, 129, 3001])
def test_pkf_backward_vs_autograd(name, T):
    ops = _ops()
    t, y, cov, ssm = make_problem(name, T, seed=T + 1)
    P0, Fs, Qs, H, R = [x.clone().requires_grad_(True) for x in ssm]
    fm, fP, ll = O.pkf((P0, Fs, Qs, H, R), y[:, None], True)
    g = 0.7
    gP0, gFs, gQs, gH, gR = torch.autograd.grad(g * ll, (P0, Fs, Qs, H, R))
    d = lambda x: x.detach().to(DEV).contiguous()
    yd = torch.as_tensor(y).to(DEV)
    fms, fPs, lld, _ = ops.pkf(d(P0), d(Fs), d(Qs), d(H).reshape(-1), d(R).reshape(-1), yd)
    dP0, dFs, dQs, dH, dR = ops.pkf_backward(d(P0), d(Fs), d(Qs), d(H).reshape(-1), d(R).reshape(-1), yd, fms, fPs,
                                             torch.tensor([g], dtype=torch.float64, device=DEV))
    tol = 1e-9
    assert rel_err(dFs.cpu(), gFs) < tol
    assert rel_err(dQs.cpu(), sym(gQs)) < tol
    assert rel_err(dP0.cpu(), sym(gP0)) < tol
    scale = max(float(gH.abs().max()), 1e-300)
    assert float((dH.cpu() - gH.reshape(-1)).abs().max()) / scale < tol
    assert abs(float(dR) - float(gR)) <= tol * abs(float(gR))


@pytest.mark.parametrize("name,span", [("matern32", 4.0), ("matern52", 4.0), ("matern52", 400.0), ("m32xm32", 50.0),
                                       ("matern12", 4.0)])
def test_discretise_forward(name, span):
    """Fs against torch.linalg.matrix_exp; Qs against the stationary form AND the reference's
    matrix-fraction form (pssgp/kernels/base.py:39-46) — they agree to rounding in absolute terms."""
    ops = _ops()
    T = 2000
    t, y, cov, ssm = make_problem(name, T, seed=3, span=span)
    with torch.no_grad():
        sde = cov.get_sde()
        ssm_stat = O.get_ssm_stationary(sde, t[:, None], ssm.R)
    dts = torch.as_tensor(np.diff(np.concatenate([[0.0], t]))).to(DEV)
    Fs, Qs = ops.discretise(sde.F.to(DEV).contiguous(), sde.P0.to(DEV).contiguous(), dts)
    assert rel_err(Fs.cpu(), ssm.Fs) < 1e-13
    pnorm = float(sde.P0.abs().max())
    assert float((Qs.cpu() - ssm_stat.Qs).abs().max()) < 1e-14 * pnorm
    assert float((Qs.cpu() - ssm.Qs).abs().max()) < 1e-13 * pnorm


@pytest.mark.parametrize("name,span", [("matern32", 4.0), ("matern52", 4.0), ("matern52", 4000.0), ("m32xm32", 80.0)])
def test_discretise_backward_vs_autograd(name, span):
    ops = _ops()
    T = 1500
    t, y, cov, ssm = make_problem(name, T, seed=5, span=span)
    with torch.no_grad():
        sde = cov.get_sde()
    F = sde.F.clone().requires_grad_(True)
    Pinf = sde.P0.clone().requires_grad_(True)
    dts = torch.as_tensor(np.diff(np.concatenate([[0.0], t])))
    Fs_ref = torch.linalg.matrix_exp(dts.reshape(-1, 1, 1) * F.unsqueeze(0))
    Qs_ref = Pinf.unsqueeze(0) - Fs_ref @ Pinf.unsqueeze(0) @ Fs_ref.transpose(1, 2)
    Qs_ref = sym(Qs_ref)
    gen = torch.Generator().manual_seed(0)
    dFs = torch.randn(Fs_ref.shape, dtype=torch.float64, generator=gen)
    dQs = sym(torch.randn(Qs_ref.shape, dtype=torch.float64, generator=gen))
    gF, gP = torch.autograd.grad((Fs_ref * dFs).sum() + (Qs_ref * dQs).sum(), (F, Pinf))
    Fd, Pd, dtd = F.detach().to(DEV).contiguous(), Pinf.detach().to(DEV).contiguous(), dts.to(DEV)
    Fs, Qs = ops.discretise(Fd, Pd, dtd)
    dF, dP = ops.discretise_backward(Fd, Pd, dtd, Fs, dFs.to(DEV), dQs.to(DEV))
    assert rel_err(dF.cpu(), gF) < 1e-10
    assert rel_err(dP.cpu(), sym(gP)) < 1e-10
