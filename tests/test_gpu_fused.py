"""GPU parity of the fused filter + smoother + gradient step (C ABI pssgp_pkfs_grad) against the CPU oracle
(pkf, pks and torch autograd through pkf) — FP64 <= 1e-9 relative to ||reference||_inf, and against the
three separate entry points at the benchmark's size."""
import numpy as np
import pytest
import torch

from util import O, make_problem, pkg, rel_err

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)
TOL = 1e-9


def _ops():
    pkg()
    from pssgp_b200 import ops
    return ops


def sym(X):
    return 0.5 * (X + X.transpose(-1, -2))


def _check_against_oracle(name, T, seed, nan_frac=0.05, chunk=0, fused_reverse=0):
    ops = _ops()
    from pssgp_b200 import _lib
    t, y, cov, ssm = make_problem(name, T, seed=seed, nan_frac=nan_frac)
    P0, Fs, Qs, H, R = [x.clone().requires_grad_(True) for x in ssm]
    fm, fP, ll = O.pkf((P0, Fs, Qs, H, R), y[:, None], True, max_parallel=max(T, 10000))
    g = 0.7
    gP0, gFs, gQs, gH, gR = torch.autograd.grad(g * ll, (P0, Fs, Qs, H, R))
    with torch.no_grad():
        rsm, rsP = O.pks(ssm, fm.detach(), fP.detach(), max_parallel=max(T, 10000))
    d = lambda x: x.detach().to(DEV).contiguous()
    yd = torch.as_tensor(y).to(DEV)
    h = _lib.handle(torch.cuda.current_device())
    h.set_option("chunk", chunk)
    h.set_option("fused_reverse", fused_reverse)
    try:
        (fms, fPs, lld), (sms, sPs), (dP0, dFs, dQs, dH, dR) = ops.pkfs_grad(
            d(P0), d(Fs), d(Qs), d(H).reshape(-1), d(R).reshape(-1), yd, torch.tensor([g], dtype=torch.float64, device=DEV))
    finally:
        h.set_option("chunk", 0)
        h.set_option("fused_reverse", 0)
    assert rel_err(fms.cpu(), fm) < TOL and rel_err(fPs.cpu(), fP) < TOL
    assert abs(float(lld) - float(ll)) <= TOL * max(1.0, abs(float(ll)))
    assert rel_err(sms.cpu(), rsm) < TOL and rel_err(sPs.cpu(), rsP) < TOL
    assert rel_err(dFs.cpu(), gFs) < TOL
    assert rel_err(dQs.cpu(), sym(gQs)) < TOL
    assert rel_err(dP0.cpu(), sym(gP0)) < TOL
    scale = max(float(gH.abs().max()), 1e-300)
    assert float((dH.cpu() - gH.reshape(-1)).abs().max()) / scale < TOL
    assert abs(float(dR) - float(gR)) <= TOL * abs(float(gR))


@pytest.mark.parametrize("name", ["matern12", "matern32", "matern52", "m32xm32", "rbf6"])
@pytest.mark.parametrize("T", [1, 2, 3, 33, 129, 1000, 4097, 20011, 400003])
def test_fused_step_vs_oracle(name, T):
    if T > 100000 and name != "matern52":
        pytest.skip("one large case (the two-length main partition) is enough")
    if name == "rbf6" and T > 1000:
        pytest.skip("generic-d path is covered by test_gpu_generic_d.py; here only the fall-through of the fused entry")
    _check_against_oracle(name, T, seed=T + 2)


@pytest.mark.parametrize("chunk", [2, 4, 8, 64, 258])
def test_fused_step_chunk_invariance(chunk):
    _check_against_oracle("matern52", 5003, seed=11, chunk=chunk)


@pytest.mark.parametrize("name", ["matern12", "matern32", "matern52", "m32xm32"])
@pytest.mark.parametrize("T", [1, 2, 33, 1000, 20011, 400003])
def test_fused_reverse_kernel_vs_oracle(name, T):
    """option fused_reverse = 1: smoother + adjoint recursions in one kernel (per-row output staging)."""
    if T > 100000 and name != "matern52":
        pytest.skip("one large case is enough")
    _check_against_oracle(name, T, seed=T + 5, fused_reverse=1)


@pytest.mark.parametrize("variant", ["first", "last", "all", "none"])
def test_fused_step_missing_observations(variant):
    ops = _ops()
    t, y, cov, ssm = make_problem("matern32", 700, seed=1, nan_frac=0.0)
    yy = y.copy()
    if variant == "first":
        yy[0] = np.nan
    elif variant == "last":
        yy[-1] = np.nan
    elif variant == "all":
        yy[:] = np.nan
    P0, Fs, Qs, H, R = [x.clone().requires_grad_(True) for x in ssm]
    fm, fP, ll = O.pkf((P0, Fs, Qs, H, R), yy[:, None], True)
    with torch.no_grad():
        rsm, rsP = O.pks(ssm, fm.detach(), fP.detach())
    d = lambda x: x.detach().to(DEV).contiguous()
    (fms, fPs, lld), (sms, sPs), grads = ops.pkfs_grad(d(P0), d(Fs), d(Qs), d(H).reshape(-1), d(R).reshape(-1),
                                                       torch.as_tensor(yy).to(DEV),
                                                       torch.ones(1, dtype=torch.float64, device=DEV))
    assert rel_err(fPs.cpu(), fP) < TOL and rel_err(sPs.cpu(), rsP) < TOL
    assert float((sms.cpu() - rsm).abs().max()) < 1e-9
    assert abs(float(lld) - float(ll)) <= TOL * max(1.0, abs(float(ll)))
    if variant != "all":
        gFs, gQs = torch.autograd.grad(ll, (Fs, Qs))
        assert rel_err(grads[1].cpu(), gFs) < TOL and rel_err(grads[2].cpu(), sym(gQs)) < TOL
    else:
        assert float(grads[1].abs().max()) == 0.0 and float(grads[2].abs().max()) == 0.0


def test_fused_step_matches_separate_calls_at_bench_size():
    """N = 1e6 (BASELINE configs[1]): the oracle is too slow, so the fused step is checked against the
    three separate scans (themselves oracle-checked at small sizes) and through the identity
    'smoothed == filtered at the last step'."""
    ops = _ops()
    import bench
    from pssgp_b200 import kernels
    n = 1_000_000
    t, y = bench.make_series(n)
    with torch.no_grad():
        sde = kernels.Matern52(1.0, 1.0).get_sde()
    F, Pinf, H = sde.F.to(DEV).contiguous(), sde.P0.to(DEV).contiguous(), sde.H.to(DEV).reshape(-1).contiguous()
    R = torch.tensor([0.1], dtype=torch.float64, device=DEV)
    td = torch.as_tensor(t).to(DEV)
    dts = td - torch.cat([torch.zeros(1, dtype=torch.float64, device=DEV), td[:-1]])
    yd = torch.as_tensor(y).to(DEV)
    g1 = torch.ones(1, dtype=torch.float64, device=DEV)
    Fs, Qs = ops.discretise(F, Pinf, dts)
    fms, fPs, ll, _ = ops.pkf(Pinf, Fs, Qs, H, R, yd)
    sms, sPs, _ = ops.pks(Fs, Qs, fms, fPs)
    ref_g = ops.pkf_backward(Pinf, Fs, Qs, H, R, yd, fms, fPs, g1)
    (fms2, fPs2, ll2), (sms2, sPs2), g2 = ops.pkfs_grad(Pinf, Fs, Qs, H, R, yd, g1)
    assert rel_err(fms2.cpu(), fms.cpu()) < 1e-12 and rel_err(fPs2.cpu(), fPs.cpu()) < 1e-12
    assert abs(float(ll2) - float(ll)) <= 1e-12 * abs(float(ll))
    assert rel_err(sms2.cpu(), sms.cpu()) < TOL and rel_err(sPs2.cpu(), sPs.cpu()) < TOL
    for a, b in zip(g2, ref_g):
        assert rel_err(a.cpu(), b.cpu()) < TOL
    assert torch.equal(sms2[-1], fms2[-1])


@pytest.mark.parametrize("name", ["matern12", "matern32", "matern52", "m32xm32"])
@pytest.mark.parametrize("T", [1, 2, 33, 1000, 20011])
def test_fused_pkfs_and_projection(name, T):
    """C ABI pssgp_pkfs: filter + smoother in one call, full outputs and the (H m, H P H^T) projection."""
    ops = _ops()
    t, y, cov, ssm = make_problem(name, T, seed=T + 9)
    with torch.no_grad():
        fm, fP, ll = O.pkf(ssm, y[:, None], True, max_parallel=max(T, 10000))
        rsm, rsP = O.pks(ssm, fm, fP, max_parallel=max(T, 10000))
    d = lambda x: x.detach().to(DEV).contiguous()
    P0, Fs, Qs, H, R = d(ssm.P0), d(ssm.Fs), d(ssm.Qs), d(ssm.H).reshape(-1), d(ssm.R).reshape(-1)
    yd = torch.as_tensor(y).to(DEV)
    fms, fPs, lld, sms, sPs = ops.pkfs(P0, Fs, Qs, H, R, yd, want_ll=True)
    assert rel_err(fms.cpu(), fm) < TOL and rel_err(fPs.cpu(), fP) < TOL
    assert abs(float(lld) - float(ll)) <= TOL * max(1.0, abs(float(ll)))
    assert rel_err(sms.cpu(), rsm) < TOL and rel_err(sPs.cpu(), rsP) < TOL
    if name == "m32xm32":
        # d = 4 in FP64: the kernels of the fused path do not share one partition of the time axis (shared-memory
        # footprints differ), so the projected output is refused loudly and predict_f uses sms / sPs
        from pssgp_b200 import _lib
        with pytest.raises(_lib.PssgpError):
            ops.pkfs(P0, Fs, Qs, H, R, yd, project=True)
        return
    proj = ops.pkfs(P0, Fs, Qs, H, R, yd, project=True)[3]
    h = ssm.H.reshape(-1)
    ref_mean = rsm @ h
    ref_var = torch.einsum("i,kij,j->k", h, rsP, h)
    assert rel_err(proj[:, 0].cpu(), ref_mean) < TOL or float(ref_mean.abs().max()) == 0.0
    assert rel_err(proj[:, 1].cpu(), ref_var) < TOL


@pytest.mark.parametrize("name,T", [("matern32", 4000), ("matern52", 20011), ("matern12", 777)])
@pytest.mark.parametrize("fused_reverse", [0, 1])
def test_fused_step_fp32(name, T, fused_reverse):
    """FP32 opt-in mode of the fused step (4 rows per 16-byte-aligned segment instead of 2): tolerance 2e-3
    relative to ||reference||_inf for the moments and the log-likelihood, 2e-2 for the gradients (float32 eps
    1.2e-7 amplified by the T-step recursions; the adjoint sums T terms of mixed sign)."""
    ops = _ops()
    from pssgp_b200 import _lib
    t, y, cov, ssm = make_problem(name, T, seed=T + 3)
    P0, Fs, Qs, H, R = [x.clone().requires_grad_(True) for x in ssm]
    fm, fP, ll = O.pkf((P0, Fs, Qs, H, R), y[:, None], True, max_parallel=max(T, 10000))
    gP0, gFs, gQs, gH, gR = torch.autograd.grad(ll, (P0, Fs, Qs, H, R))
    with torch.no_grad():
        rsm, rsP = O.pks(ssm, fm.detach(), fP.detach(), max_parallel=max(T, 10000))
    d = lambda x: x.detach().to(device=DEV, dtype=torch.float32).contiguous()
    h = _lib.handle(torch.cuda.current_device())
    h.set_option("fused_reverse", fused_reverse)
    try:
        (fms, fPs, lld), (sms, sPs), (dP0, dFs, dQs, dH, dR) = ops.pkfs_grad(
            d(P0), d(Fs), d(Qs), d(H).reshape(-1), d(R).reshape(-1), torch.as_tensor(y, dtype=torch.float32).to(DEV),
            torch.ones(1, dtype=torch.float32, device=DEV))
    finally:
        h.set_option("fused_reverse", 0)
    assert fms.dtype == torch.float32
    assert rel_err(fms.cpu(), fm) < 2e-3 and rel_err(fPs.cpu(), fP) < 2e-3
    assert abs(float(lld) - float(ll)) <= 2e-3 * abs(float(ll))
    assert rel_err(sms.cpu(), rsm) < 2e-3 and rel_err(sPs.cpu(), rsP) < 2e-3
    assert rel_err(dFs.cpu(), gFs) < 2e-2 and rel_err(dQs.cpu(), sym(gQs)) < 2e-2
    assert abs(float(dR) - float(gR)) <= 2e-2 * abs(float(gR))


def test_fused_step_beyond_2gb_arrays():
    """N = 3.2e7 at d = 3: Fs, Qs, fPs, sPs, dFs, dQs are 2.3 GB each, so every byte offset past the first 2^31
    exercises the 64-bit address arithmetic of the streaming kernels.  Checked against the three separate scans
    (same kernels, different partition bookkeeping) and through size-independent identities: the smoothed state
    equals the filtered one at the last step, smoothed variances never exceed filtered ones, and the
    log-likelihood of the whole equals the sum over two halves seeded with the filtered state at the cut."""
    ops = _ops()
    if torch.cuda.get_device_properties(0).total_memory < 60e9:
        pytest.skip("needs ~35 GB of device memory")
    from pssgp_b200 import kernels
    n = 32_000_000
    rng = np.random.RandomState(12)
    dts = torch.as_tensor(0.004 * rng.uniform(0.5, 1.5, size=n)).to(DEV)
    t = torch.cumsum(dts, 0)
    y = torch.sin(np.pi * t) + 0.3 * torch.as_tensor(rng.standard_normal(n)).to(DEV)
    y[::97] = float("nan")
    with torch.no_grad():
        sde = kernels.Matern52(1.0, 1.0).get_sde()
    F, Pinf, H = sde.F.to(DEV).contiguous(), sde.P0.to(DEV).contiguous(), sde.H.to(DEV).reshape(-1).contiguous()
    R = torch.tensor([0.1], dtype=torch.float64, device=DEV)
    g1 = torch.ones(1, dtype=torch.float64, device=DEV)
    Fs, Qs = ops.discretise(F, Pinf, dts)
    (fms, fPs, ll), (sms, sPs), grads = ops.pkfs_grad(Pinf, Fs, Qs, H, R, y, g1)
    assert torch.isfinite(ll).all() and torch.isfinite(sms).all() and torch.isfinite(grads[1]).all()
    assert torch.equal(sms[-1], fms[-1])
    assert bool((sPs[:, 0, 0] <= fPs[:, 0, 0] * (1 + 1e-12) + 1e-15).all())
    # separate scans on the same data
    fms2, fPs2, ll2, _ = ops.pkf(Pinf, Fs, Qs, H, R, y)
    assert float((fms2 - fms).abs().max()) <= 1e-12 * float(fms.abs().max())
    assert abs(float(ll2) - float(ll)) <= 1e-12 * abs(float(ll))
    del fms2, fPs2
    sms2, sPs2, _ = ops.pks(Fs, Qs, fms, fPs)
    assert float((sms2 - sms).abs().max()) <= 1e-9 * float(sms.abs().max())
    assert float((sPs2 - sPs).abs().max()) <= 1e-9 * float(sPs.abs().max())
    del sms2, sPs2
    ref_g = ops.pkf_backward(Pinf, Fs, Qs, H, R, y, fms, fPs, g1)
    for a, b in zip(grads, ref_g):
        assert float((a - b).abs().max()) <= 1e-9 * max(float(b.abs().max()), 1e-300)
    del ref_g
    # split at the middle: second half seeded with the filtered state at the cut
    h = n // 2
    _, _, ll_a, fin = ops.pkf(Pinf, Fs[:h], Qs[:h], H, R, y[:h], want_final=True)
    d = 3
    _, _, ll_b, _ = ops.pkf(fin[d:].reshape(d, d).contiguous(), Fs[h:], Qs[h:], H, R, y[h:], m0=fin[:d].contiguous(),
                            first_special=False)
    assert abs(float(ll_a) + float(ll_b) - float(ll)) <= 1e-10 * abs(float(ll))
