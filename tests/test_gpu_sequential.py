"""GPU parity of the sequential Kalman filter / RTS smoother (C ABI pssgp_kf / pssgp_ks, drop-in
pssgp_b200.kalman.sequential) against the oracle's restatement of pssgp/kalman/sequential.py:11-73, for the
thread-per-series kernels (d <= 4) and the warp-per-series kernels (5 <= d <= 32), single series and batches, FP64
(1e-9 relative to ||oracle output||_inf; 1e-6 for the quasi-periodic kernels, cond(Pinf) up to 3.4e8) and FP32
(1e-3), and the StateSpaceGP(parallel=False) model of the reference (model.py:73-77)."""
import numpy as np
import pytest
import torch

from util import O, make_problem, pkg, rel_err

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)

CASES = [("matern12", 1e-9), ("matern32", 1e-9), ("matern52", 1e-9), ("m32xm32", 1e-9), ("m32+m52", 1e-9),
         ("rbf6", 1e-9), ("m52+rbf6", 1e-9), ("qp3", 1e-6), ("qp5", 1e-6)]


def _oracle(ssm, y):
    with torch.no_grad():
        fm, fP, ll, mp, Pp = O.kf(ssm, y[:, None], True, True)
        sm, sP = O.ks(ssm, fm, fP, mp, Pp)
    return fm, fP, ll, mp, Pp, sm, sP


@pytest.mark.parametrize("name,tol", CASES)
@pytest.mark.parametrize("T", [1, 2, 37, 300])
def test_kf_ks_match_oracle(name, tol, T):
    pkg()
    from pssgp_b200.kalman.sequential import kf, kfs, ks
    span = 40.0 if name.startswith("qp") else 4.0
    t, y, cov, ssm = make_problem(name, T, seed=3 + T, span=span)
    rfm, rfP, rll, rmp, rPp, rsm, rsP = _oracle(ssm, y)
    lg = tuple(x.detach().numpy() for x in ssm)
    fm, fP, ll, mp, Pp = kf(lg, y[:, None], return_loglikelihood=True, return_predicted=True)
    assert isinstance(fm, np.ndarray) and fm.shape == (T, ssm.P0.shape[0])
    assert rel_err(fm, rfm) < tol and rel_err(fP, rfP) < tol
    assert rel_err(mp, rmp) < tol and rel_err(Pp, rPp) < tol
    assert abs(float(ll) - float(rll)) <= tol * max(1.0, abs(float(rll)))
    # filtered covariances are exactly symmetric (sequential.py:39)
    assert np.array_equal(fP, fP.transpose(0, 2, 1))
    sm, sP = ks(lg, fm, fP, mp, Pp)
    assert rel_err(sm, rsm) < 10 * tol and rel_err(sP, rsP) < 10 * tol
    sm2, sP2 = kfs(lg, y[:, None])
    assert np.array_equal(sm, sm2) and np.array_equal(sP, sP2)
    # the reference's return conventions (sequential.py:45-47)
    assert len(kf(lg, y[:, None])) == 2 and len(kf(lg, y[:, None], return_predicted=True)) == 4


@pytest.mark.parametrize("name", ["matern52", "rbf6"])
def test_kf_equals_pkf(name):
    """kf and pkf compute the same filter."""
    pkg()
    from pssgp_b200.kalman.parallel import pkf, pkfs
    from pssgp_b200.kalman.sequential import kf, kfs
    T = 5000
    t, y, cov, ssm = make_problem(name, T, seed=11)
    lg = tuple(x.detach().to(DEV) for x in ssm)
    yd = torch.as_tensor(y[:, None]).to(DEV)
    fm, fP, ll = kf(lg, yd, return_loglikelihood=True)
    pm, pP, pll = pkf(lg, yd, return_loglikelihood=True)
    assert fm.is_cuda
    assert rel_err(fm.cpu(), pm.cpu()) < 1e-9 and rel_err(fP.cpu(), pP.cpu()) < 1e-9
    assert abs(float(ll) - float(pll)) < 1e-9 * abs(float(pll))
    sm, sP = kfs(lg, yd)
    qm, qP = pkfs(lg, yd)
    assert rel_err(sm.cpu(), qm.cpu()) < 1e-8 and rel_err(sP.cpu(), qP.cpu()) < 1e-8


@pytest.mark.parametrize("name", ["matern32", "m52+rbf6"])
@pytest.mark.parametrize("shared", [True, False])
def test_batched_series(name, shared):
    """B independent series in one call: shared LGSSM (same sampling times) or one LGSSM per series."""
    pkg()
    from pssgp_b200.kalman.sequential import kf, kfs
    B, T = 70, 40
    probs = [make_problem(name, T, seed=100 + (0 if shared else b)) for b in range(B)]
    rng = np.random.RandomState(5)
    ys = np.stack([p[1] if not shared else p[1] + 0.1 * rng.randn(T) for p in probs])
    if shared:
        lg = tuple(x.detach().numpy() for x in probs[0][3])
    else:
        lg = tuple(np.stack([p[3][i].detach().numpy() for p in probs]) for i in range(5))
    fm, fP, ll = kf(lg, ys[:, :, None], return_loglikelihood=True)
    sm, sP = kfs(lg, ys[:, :, None])
    assert fm.shape[:2] == (B, T) and ll.shape == (B,)
    for b in (0, 1, 33, B - 1):
        ssm = probs[0 if shared else b][3]
        rfm, rfP, rll, _, _, rsm, rsP = _oracle(ssm, ys[b])
        assert rel_err(fm[b], rfm) < 1e-9 and rel_err(fP[b], rfP) < 1e-9
        assert abs(ll[b] - float(rll)) < 1e-9 * max(1.0, abs(float(rll)))
        assert rel_err(sm[b], rsm) < 1e-8 and rel_err(sP[b], rsP) < 1e-8


@pytest.mark.parametrize("name", ["matern52", "rbf6"])
def test_fp32(name):
    pkg()
    from pssgp_b200.kalman.sequential import kf, kfs
    T = 200
    t, y, cov, ssm = make_problem(name, T, seed=2)
    rfm, rfP, rll, _, _, rsm, rsP = _oracle(ssm, y)
    lg = tuple(x.detach().numpy().astype(np.float32) for x in ssm)
    y32 = y[:, None].astype(np.float32)
    fm, fP, ll = kf(lg, y32, return_loglikelihood=True)
    assert fm.dtype == np.float32
    assert rel_err(fm, rfm) < 1e-3 and rel_err(fP, rfP) < 1e-3 and abs(float(ll) - float(rll)) < 1e-3 * abs(float(rll))
    sm, sP = kfs(lg, y32)
    assert rel_err(sm, rsm) < 5e-3 and rel_err(sP, rsP) < 5e-3


def test_argument_errors():
    pkg()
    from pssgp_b200 import _lib
    from pssgp_b200.kalman.sequential import kf
    t, y, cov, ssm = make_problem("matern32", 10, seed=0)
    lg = tuple(x.detach().numpy() for x in ssm)
    with pytest.raises(ValueError):
        kf(lg, y[:5, None])
    h = _lib.handle(0)
    z = torch.zeros(40 * 40 * 2, dtype=torch.float64, device=DEV)
    rc = _lib.lib().pssgp_kf(h.ptr, 0, 1, 1, 40, 0, z.data_ptr(), z.data_ptr(), z.data_ptr(), z.data_ptr(),
                             z.data_ptr(), z.data_ptr(), z.data_ptr(), z.data_ptr(), None, None, None, None)
    assert rc == 3  # PSSGP_ERR_UNSUPPORTED: d > 32
    rc = _lib.lib().pssgp_kf(h.ptr, 0, 1, 1, 2, 0, z.data_ptr(), z.data_ptr(), z.data_ptr(), z.data_ptr(),
                             z.data_ptr(), z.data_ptr(), z.data_ptr(), z.data_ptr(), z.data_ptr(), None, None, None)
    assert rc == 1  # mps without Pps


@pytest.mark.parametrize("kname", ["matern32", "m52+rbf6"])
def test_model_sequential_mode(kname):
    """StateSpaceGP(parallel=False) (the reference's default, model.py:73-77): ll, gradient and predict_f equal the
    oracle's sequential model and this package's parallel model."""
    pkg()
    from pssgp_b200 import kernels as PK
    from pssgp_b200.model import StateSpaceGP
    rng = np.random.RandomState(31415926)
    T, K = 200, 50
    t = np.sort(rng.rand(T))
    y = O.obs_noise(O.sinu(t), 0.1, 1)
    q = np.sort(rng.rand(K, 1), 0)
    if kname == "matern32":
        ocov, mk = O.Matern32(1., .5), lambda: PK.Matern32(1., .5)
    else:
        ocov = O.Matern52(1., 1.) + O.RBF(1., 1., order=6, balancing_iter=5)
        mk = lambda: PK.Matern52(1., 1.) + PK.RBF(1., 1., order=6, balancing_iter=5)
    oss = O.StateSpaceGP((t, y), ocov, 0.1, parallel=False, max_parallel=T + K)
    oll = oss.maximum_log_likelihood_objective()
    ograd = torch.autograd.grad(oll, ocov.trainable_variables)
    with torch.no_grad():
        om, ov = oss.predict_f(q)
    res = {}
    for par in (False, True):
        ss = StateSpaceGP((t[:, None], y[:, None]), mk(), 0.1, parallel=par, max_parallel=T + K)
        ll = ss.maximum_log_likelihood_objective()
        grad = torch.autograd.grad(ll, ss.kernel.trainable_variables)
        m, v = ss.predict_f(q)
        res[par] = (float(ll), [float(g) for g in grad], m, v)
    ll, grad, m, v = res[False]
    assert abs(ll - float(oll)) < 1e-9 * abs(float(oll))
    for g, og in zip(grad, ograd):
        assert abs(g - float(og)) < 1e-7 * max(1.0, abs(float(og)))
    assert rel_err(m, om) < 1e-8 and rel_err(v, ov) < 1e-8
    assert abs(ll - res[True][0]) < 1e-9 * abs(ll)
    assert rel_err(m, res[True][2]) < 1e-8 and rel_err(v, res[True][3]) < 1e-8


def test_long_series_against_numba_port():
    """One series of 2e5 steps (the comparator use of kf / ks): against the numba restatement of sequential.py
    (oracle/seq_numba.py, itself pinned to the torch restatement in tests/test_cpu_oracle.py)."""
    pkg()
    import seq_numba
    from pssgp_b200.kalman.sequential import kf, kfs
    T = 200_000
    t, y, cov, ssm = make_problem("matern52", T, seed=8, span=800.0)
    P0, Fs, Qs, H, R = [np.ascontiguousarray(x.detach().numpy()) for x in ssm]
    rfm, rfP, rmp, rPp, rll = seq_numba.kf(P0, Fs, Qs, H, R, y)
    rsm, rsP = seq_numba.ks(Fs, rfm, rfP, rmp, rPp)
    fm, fP, ll = kf((P0, Fs, Qs, H, R), y[:, None], return_loglikelihood=True)
    assert rel_err(fm, rfm) < 1e-9 and rel_err(fP, rfP) < 1e-9 and abs(float(ll) - float(rll)) < 1e-9 * abs(float(rll))
    sm, sP = kfs((P0, Fs, Qs, H, R), y[:, None])
    assert rel_err(sm, rsm) < 1e-8 and rel_err(sP, rsP) < 1e-8
