"""GPU parity of the warp-level (DMMA) path for 5 <= d <= 32 against the oracle: fused filter + log-likelihood +
smoother + gradient (C ABI pssgp_pkfs_grad), filter + smoother (pssgp_pkfs), and the chunk-length invariance of both.

The smoother of this path runs in modified Bryson-Frazier form (no d x d solve per step); the oracle is the
reference's RTS element form (pssgp/kalman/parallel.py:155-196) — equal in exact arithmetic, compared at
1e-9 relative to ||oracle output||_inf (1e-7 for the quasi-periodic kernels, cond(Pinf) 2.6e5 .. 3.4e8,
SURVEY.md App. B.6)."""
import numpy as np
import pytest
import torch

from util import O, make_problem, pkg, rel_err

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)

CASES = [("m32+m52", 1e-9), ("rbf6", 1e-9), ("m32xm52", 1e-9), ("m52+rbf6", 1e-9), ("periodic2", 1e-9), ("qp3", 1e-7),
         ("qp5", 1e-7)]


def sym(X):
    return 0.5 * (X + X.transpose(-1, -2))


def oracle_all(ssm, y, T, g):
    P0, Fs, Qs, H, R = [x.clone().requires_grad_(True) for x in ssm]
    fm, fP, ll = O.pkf((P0, Fs, Qs, H, R), y[:, None], True, max_parallel=max(T, 2))
    grads = torch.autograd.grad(g * ll, (P0, Fs, Qs, H, R))
    with torch.no_grad():
        sm, sP = O.pks(ssm, fm.detach(), fP.detach(), max_parallel=max(T, 2))
    return fm.detach(), fP.detach(), ll.detach(), sm, sP, grads


def check(name, tol, T, seed, chunk=0, smem=0):
    pkg()
    from pssgp_b200 import _lib, ops
    span = 40.0 if name.startswith("qp") else 4.0
    t, y, cov, ssm = make_problem(name, T, seed=seed, span=span)
    # the matrix-fraction Q of the reference (kernels/base.py:39-46) is symmetric only to rounding RELATIVE TO ||expm||
    # (8e-7 relative to a tiny Q_0 for the product kernels): both sides get the symmetrised Q so that dH, the only
    # output that sees the antisymmetric part, is compared on the same inputs
    ssm = ssm._replace(Qs=sym(ssm.Qs))
    g = 0.9
    rfm, rfP, rll, rsm, rsP, (gP0, gFs, gQs, gH, gR) = oracle_all(ssm, y, T, g)
    to = lambda x: x.detach().to(DEV).contiguous()
    P0, Fs, Qs, H, R = to(ssm.P0), to(ssm.Fs), to(ssm.Qs), to(ssm.H).reshape(-1), to(ssm.R).reshape(-1)
    yd = torch.as_tensor(y).to(DEV)
    h = _lib.handle(0)
    h.set_option("chunk", chunk)
    h.set_option("mid_smem", smem)   # 1: shared-memory tile kernels also for d <= 16 (default there: fragment-resident)
    try:
        (fms, fPs, ll), (sms, sPs), (dP0, dFs, dQs, dH, dR) = ops.pkfs_grad(
            P0, Fs, Qs, H, R, yd, torch.tensor([g], dtype=torch.float64, device=DEV))
        f2, fP2, ll2, s2, sP2 = ops.pkfs(P0, Fs, Qs, H, R, yd, want_ll=True)
    finally:
        h.set_option("chunk", 0)
        h.set_option("mid_smem", 0)
    assert rel_err(fms.cpu(), rfm) < tol and rel_err(fPs.cpu(), rfP) < tol
    assert abs(float(ll) - float(rll)) <= tol * max(1.0, abs(float(rll)))
    assert rel_err(sms.cpu(), rsm) < tol and rel_err(sPs.cpu(), rsP) < tol
    assert rel_err(dFs.cpu(), gFs) < tol
    assert rel_err(dQs.cpu(), sym(gQs)) < tol
    assert rel_err(dP0.cpu(), sym(gP0)) < tol
    assert float((dH.cpu() - gH.reshape(-1)).abs().max()) <= tol * max(float(gH.abs().max()), 1e-3 * float(gFs.abs().max()))
    assert abs(float(dR) - float(gR)) <= tol * abs(float(gR))
    # pkfs = the same without the gradient
    assert rel_err(f2.cpu(), rfm) < tol and rel_err(fP2.cpu(), rfP) < tol
    assert abs(float(ll2) - float(rll)) <= tol * max(1.0, abs(float(rll)))
    assert rel_err(s2.cpu(), rsm) < tol and rel_err(sP2.cpu(), rsP) < tol


@pytest.mark.parametrize("smem", [0, 1])
@pytest.mark.parametrize("name,tol", CASES)
@pytest.mark.parametrize("T", [1, 2, 17, 300, 2051])
def test_fused_step_vs_oracle(name, tol, T, smem):
    if smem and name == "qp5":
        pytest.skip("d > 16 always runs the shared-memory family")
    check(name, tol, T, seed=T, smem=smem)


@pytest.mark.parametrize("smem", [0, 1])
@pytest.mark.parametrize("name,tol", CASES)
@pytest.mark.parametrize("chunk", [1, 3, 16, 700])
def test_fused_step_chunk_invariance(name, tol, chunk, smem):
    if smem and name == "qp5":
        pytest.skip("d > 16 always runs the shared-memory family")
    check(name, tol, 1200, seed=9, chunk=chunk, smem=smem)


@pytest.mark.parametrize("name,tol", [("rbf6", 2e-6), ("m52+rbf6", 2e-6), ("qp3", 2e-5), ("qp5", 2e-4)])
def test_fp32_storage_mode(name, tol):
    """FP32 opt-in mode at 5 <= d <= 32 = FP32 storage at the C ABI, FP64 arithmetic (csrc/mid_f32.cu): on inputs that
    are exactly representable in FP32 the results equal the oracle's up to the FP32 rounding of the outputs."""
    pkg()
    from pssgp_b200 import ops
    T = 1500
    span = 40.0 if name.startswith("qp") else 4.0
    t, y, cov, ssm = make_problem(name, T, seed=21, span=span)
    ssm = ssm._replace(Qs=sym(ssm.Qs))
    ssm32 = type(ssm)(*[x.float().double() for x in ssm])
    ssm32 = ssm32._replace(Qs=sym(ssm32.Qs), P0=sym(ssm32.P0))
    y32 = torch.as_tensor(y).float().double().numpy()
    g = 1.0
    rfm, rfP, rll, rsm, rsP, (gP0, gFs, gQs, gH, gR) = oracle_all(ssm32, y32, T, g)
    to = lambda x: x.detach().to(device=DEV, dtype=torch.float32).contiguous()
    P0, Fs, Qs, H, R = to(ssm32.P0), to(ssm32.Fs), to(ssm32.Qs), to(ssm32.H).reshape(-1), to(ssm32.R).reshape(-1)
    yd = torch.as_tensor(y32).to(device=DEV, dtype=torch.float32)
    (fms, fPs, ll), (sms, sPs), (dP0, dFs, dQs, dH, dR) = ops.pkfs_grad(P0, Fs, Qs, H, R, yd,
                                                                     torch.tensor([g], dtype=torch.float32, device=DEV))
    assert fms.dtype == torch.float32 and dFs.dtype == torch.float32
    c = lambda x: x.double().cpu()
    assert rel_err(c(fms), rfm) < tol and rel_err(c(fPs), rfP) < tol
    assert abs(float(ll) - float(rll)) <= tol * max(1.0, abs(float(rll)))
    assert rel_err(c(sms), rsm) < tol and rel_err(c(sPs), rsP) < tol
    assert rel_err(c(dFs), gFs) < tol and rel_err(c(dQs), sym(gQs)) < tol
    f2, fP2, ll2, s2, sP2 = ops.pkfs(P0, Fs, Qs, H, R, yd, want_ll=True)
    assert rel_err(c(s2), rsm) < tol and rel_err(c(sP2), rsP) < tol
    f3, fP3, ll3, _ = ops.pkf(P0, Fs, Qs, H, R, yd)
    assert rel_err(c(f3), rfm) < tol and abs(float(ll3) - float(rll)) <= tol * max(1.0, abs(float(rll)))


def _random_lgssm(d, T, seed):
    """A stable random LGSSM of state dimension d (not tied to a covariance function): covers every compiled
    instantiation of the warp-level path, including exact multiples of the 8 x 8 tile and the shared-memory family."""
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(d, d, dtype=torch.float64, generator=g) / (2.0 * d ** 0.5)
    F0 = torch.linalg.matrix_exp(A - 0.6 * torch.eye(d, dtype=torch.float64))
    dts = 0.5 + torch.rand(T, dtype=torch.float64, generator=g)
    Fs = torch.stack([torch.linalg.matrix_power(F0, 1) * (0.9 + 0.1 * float(x)) for x in dts])
    B = torch.randn(T, d, d, dtype=torch.float64, generator=g) / d ** 0.5
    Qs = 0.05 * (B @ B.transpose(-1, -2)) + 0.01 * torch.eye(d, dtype=torch.float64)
    B0 = torch.randn(d, d, dtype=torch.float64, generator=g) / d ** 0.5
    P0 = B0 @ B0.T + 0.1 * torch.eye(d, dtype=torch.float64)
    H = torch.randn(d, dtype=torch.float64, generator=g)
    R = torch.tensor([0.3], dtype=torch.float64)
    y = torch.randn(T, dtype=torch.float64, generator=g)
    y[torch.rand(T, generator=g) < 0.1] = float("nan")
    return P0, Fs, Qs, H, R, y


@pytest.mark.parametrize("d", [5, 7, 8, 10, 12, 13, 15, 16, 17, 20, 23, 24, 25, 28, 31, 32])
def test_every_dimension_against_the_cta_cooperative_path(d):
    """Warp-level DMMA path (fragment family d <= 24, shared-memory family above, bordered form at d = 9 is covered by
    the kernel cases) against the independent CTA-cooperative kernels of generic.cu (option force_generic: RTS element
    smoother, separate adjoint scan), which the oracle tests validate: 1e-9 on every output of the fused step."""
    pkg()
    from pssgp_b200 import _lib, ops
    T = 257
    P0, Fs, Qs, H, R, y = (x.to(DEV).contiguous() for x in _random_lgssm(d, T, seed=100 + d))
    g = torch.tensor([1.2], dtype=torch.float64, device=DEV)
    h = _lib.handle(0)
    h.set_option("force_generic", 1)
    try:
        (rfm, rfP, rll), (rsm, rsP), (rdP0, rdF, rdQ, rdH, rdR) = ops.pkfs_grad(P0, Fs, Qs, H, R, y, g)
    finally:
        h.set_option("force_generic", 0)
    (fm, fP, ll), (sm, sP), (dP0, dF, dQ, dH, dR) = ops.pkfs_grad(P0, Fs, Qs, H, R, y, g)
    tol = 1e-9
    for a, b in ((fm, rfm), (fP, rfP), (ll, rll), (sm, rsm), (sP, rsP), (dP0, rdP0), (dF, rdF), (dQ, rdQ), (dH, rdH), (dR, rdR)):
        assert rel_err(a.cpu(), b.cpu()) < tol


@pytest.mark.parametrize("name", ["m32+m52", "rbf6", "m52+rbf6", "qp3", "qp5"])
def test_projected_smoother_output(name):
    """pssgp_pkfs with proj (what predict_f keeps, pssgp/model.py:107-111): (H sm_k, H sP_k H^T) straight from the
    reverse kernel equals the projection of the full smoothed moments and of the oracle's."""
    pkg()
    from pssgp_b200 import ops
    T = 777
    span = 40.0 if name.startswith("qp") else 4.0
    t, y, cov, ssm = make_problem(name, T, seed=21, span=span)
    to = lambda x: x.detach().to(DEV).contiguous()
    P0, Fs, Qs, H, R = to(ssm.P0), to(ssm.Fs), to(sym(ssm.Qs)), to(ssm.H).reshape(-1), to(ssm.R).reshape(-1)
    yd = torch.as_tensor(y).to(DEV)
    d = Fs.shape[1]
    assert ops.has_projection(d, torch.float64)
    fms, fPs, ll, proj = ops.pkfs(P0, Fs, Qs, H, R, yd, want_ll=True, project=True)
    f2, fP2, ll2, sms, sPs = ops.pkfs(P0, Fs, Qs, H, R, yd, want_ll=True)
    mean = sms @ H
    var = torch.einsum("i,kij,j->k", H, sPs, H)
    assert torch.equal(fms, f2) and torch.equal(fPs, fP2) and float(ll) == float(ll2)
    assert rel_err(proj[:, 0].cpu(), mean.cpu()) < 1e-11 and rel_err(proj[:, 1].cpu(), var.cpu()) < 1e-9
    with torch.no_grad():
        rfm, rfP = O.pkf(ssm, y[:, None], max_parallel=T)
        rsm, rsP = O.pks(ssm, rfm, rfP, max_parallel=T)
        Ho = ssm.H.reshape(-1)
        tol = 1e-7 if name.startswith("qp") else 1e-9
        assert rel_err(proj[:, 0].cpu(), rsm @ Ho) < tol
        assert rel_err(proj[:, 1].cpu(), torch.einsum("i,kij,j->k", Ho, rsP, Ho)) < tol
