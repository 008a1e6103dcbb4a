"""CPU: the oracle against every golden vector / known-answer test the reference holds for the path
(tests/test_rbf.py:27-57, tests/test_periodic.py:29-61) and against its GP-equivalence contract
(tests/test_gp_vs_kfs.py:45-99), plus the committed fixtures generated from the importable part of the
reference (scripts/make_golden.py)."""
import os

import numpy as np
import numpy.testing as npt
import pytest
import torch

from util import O, ROOT


def test_rbf_sde_coefficients_known_answer():
    """reference tests/test_rbf.py:26-47 (decimal=8)."""
    F_expected = np.array([[0, 14.520676967550859, 0], [0, 0, 32.857489440296360],
                           [-14.5210953665873, -29.4746060478111, -50.3678777987092]])
    Pinf_expected = np.array([[1.04502531824891, -1.41636387123970e-17, -0.301281550265743],
                              [-1.41636387123970e-17, 0.681741999944955, -1.70331397804495e-17],
                              [-0.301281550265743, -1.70331397804495e-17, 0.611552410634913]])
    with torch.no_grad():
        Pinf, F, L, H, Q = O.RBF(1., 0.1, order=3, balancing_iter=5).get_sde()
    npt.assert_array_almost_equal(F, F_expected, decimal=8)
    npt.assert_array_almost_equal(L, np.array([0., 0., 1.]).reshape(3, 1), decimal=8)
    npt.assert_array_almost_equal(H, np.array([1., 0., 0.]).reshape(1, 3), decimal=8)
    npt.assert_array_almost_equal(Q, 52.8553179255264, decimal=8)
    npt.assert_array_almost_equal(Pinf, Pinf_expected, decimal=8)


def test_rbf_coefficients_convergence():
    """reference tests/test_rbf.py:49-57 (balancing 5 vs 15 agree to 3 decimals)."""
    with torch.no_grad():
        a = O.RBF(1., 0.1, order=3, balancing_iter=5).get_sde()
        b = O.RBF(1., 0.1, order=3, balancing_iter=15).get_sde()
    for x, y in zip(a, b):
        npt.assert_array_almost_equal(x, y, decimal=3)


def test_periodic_offline_coeffs_known_answer():
    """reference tests/test_periodic.py:29-40."""
    b, K, div_facto_K = O._get_offline_coeffs(2)
    npt.assert_almost_equal(b, np.array([[1, 0, 0], [0, 2, 0], [2, 0, 2]]), decimal=8)
    npt.assert_almost_equal(K, np.array([[0, 0, 0], [1, 1, 1], [2, 2, 2]]), decimal=8)
    npt.assert_almost_equal(div_facto_K, np.array([[1, 1, 1], [1, 1, 1], [0.5, 0.5, 0.5]]), decimal=8)


def test_periodic_sde_coeff_known_answer():
    """reference tests/test_periodic.py:42-61 (default decimal=7; the Pinf check is vacuous there too)."""
    F_expected = np.zeros((6, 6))
    F_expected[2, 3] = -6.283185307179586
    F_expected[4, 5] = -12.5663706143592
    F_expected = F_expected - F_expected.T
    Pinf_expected = np.diag([1.20739740482544e-19, 1.20739740482544e-19, 9.64374923981979e-21, 9.64374923981979e-21,
                             1.20546865497747e-19, 1.20546865497747e-19])
    with torch.no_grad():
        Pinf, F, L, H, Q = O.Periodic(O.SquaredExponential(1., 0.1), period=1., order=2).get_sde()
    npt.assert_almost_equal(F, F_expected)
    npt.assert_almost_equal(L, np.eye(6))
    npt.assert_almost_equal(H, np.array([[1, 0, 1, 0, 1, 0]]))
    npt.assert_almost_equal(Q, np.zeros((6, 6)))
    npt.assert_almost_equal(Pinf, Pinf_expected)


def _covs():
    base = O.SquaredExponential(1., 0.5)
    c = [(O.Matern12(1., 0.5), 1e-6, 1e-2), (O.Matern32(1., 0.5), 1e-6, 1e-2), (O.Matern52(1., 0.5), 1e-6, 1e-2),
         (O.RBF(1., 0.5, order=15, balancing_iter=10), 1e-2, 1e-2), (O.Periodic(base, period=0.5, order=10), 1e-3, 1e-3)]
    c.append((c[1][0] + c[2][0], 1e-6, 1e-2))
    c.append((c[1][0] * c[2][0], 1e-6, 1e-1))
    return c


@pytest.mark.parametrize("idx", range(7))
def test_gp_equivalence_contract(idx):
    """reference tests/test_gp_vs_kfs.py:45-99: ll, gradient and posterior of StateSpaceGP(parallel in {False, True})
    equal the dense GP within the reference's per-kernel tolerances."""
    rng = np.random.RandomState(31415926)
    T, K = 200, 50
    t = np.sort(rng.rand(T))
    y = O.obs_noise(O.sinu(t), 0.1, 1)
    q = np.sort(rng.rand(K, 1), 0)
    cov, val_tol, grad_tol = _covs()[idx]
    gp = O.GPR((t, y), cov, 0.1)
    gp_ll = gp.maximum_log_likelihood_objective()
    gp_grad = torch.autograd.grad(gp_ll, cov.trainable_variables)
    gp_mean, gp_var = gp.predict_f(q)
    for parallel in (False, True):
        ss = O.StateSpaceGP((t, y), cov, 0.1, parallel=parallel, max_parallel=T + K)
        ll = ss.maximum_log_likelihood_objective()
        grad = torch.autograd.grad(ll, cov.trainable_variables)
        npt.assert_allclose(gp_ll.item(), ll.item(), atol=val_tol, rtol=val_tol)
        for a, b in zip(gp_grad, grad):
            npt.assert_allclose(a.item(), b.item(), atol=grad_tol, rtol=grad_tol)
        mean, var = ss.predict_f(q)
        npt.assert_allclose(gp_mean.detach().numpy().reshape(-1), mean.detach().numpy().reshape(-1), atol=val_tol, rtol=val_tol)
        npt.assert_allclose(gp_var.detach().numpy().reshape(-1), var.detach().numpy().reshape(-1), atol=val_tol, rtol=val_tol)


def test_pkf_equals_kf_and_pks_equals_ks_with_missing_data():
    """identities the reference relies on but never tests directly: kf == pkf, ks == pks."""
    rng = np.random.RandomState(3)
    T = 257
    t = np.sort(rng.rand(T)) * 3
    y = O.obs_noise(O.sinu(t), 0.1, 2)
    y[[0, 5, 77, 256]] = np.nan
    for cov in (O.Matern32(1., .5), O.Matern52(1.3, .7), O.RBF(1., 1., order=6, balancing_iter=5)):
        with torch.no_grad():
            ssm = cov.get_ssm(t[:, None], torch.tensor([[0.1]], dtype=torch.float64))
            fm, fP, ll = O.kf(ssm, y[:, None], return_loglikelihood=True)
            pfm, pfP, pll = O.pkf(ssm, y[:, None], return_loglikelihood=True, max_parallel=T)
            sm, sP = O.kfs(ssm, y[:, None])
            psm, psP = O.pkfs(ssm, y[:, None], max_parallel=T)
        assert float((fm - pfm).abs().max()) < 1e-10 and float((fP - pfP).abs().max()) < 1e-10
        assert abs(float(ll - pll)) < 1e-9
        assert float((sm - psm).abs().max()) < 1e-9 and float((sP - psP).abs().max()) < 1e-9


def test_golden_toy_fixture_matches_oracle_restatement():
    """tests/golden/toy_sinusoid_n1000.npz was produced by the reference's own pssgp.toymodels
    (scripts/make_golden.py); the oracle's restated sinu / obs_noise (incl. the mean-x quirk) must match bit for bit."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "toy_sinusoid_n1000.npz"))
    t = g["t"]
    npt.assert_array_equal(O.sinu(t), g["ft"])
    npt.assert_array_equal(O.obs_noise(O.sinu(t), 0.1, 0), g["y"])
    npt.assert_array_equal(O.obs_noise(O.sinu(t), 0.5, 666), g["y_seed666"])


def test_expm_restatement_against_scipy():
    """the oracle's restated tf.linalg.expm (Higham 2005) vs scipy.linalg.expm on the SDE drifts used here."""
    from scipy.linalg import expm
    with torch.no_grad():
        for cov in (O.Matern52(1., 1.), O.RBF(1., 1., order=6, balancing_iter=5),
                    O.Periodic(O.SquaredExponential(5., 1.), period=1., order=3) * O.Matern32(0.1, 50.)):
            F = cov.get_sde().F
            for dt in (1e-5, 0.004, 0.08, 1.0, 7.0):
                a = O.expm((F * dt).unsqueeze(0))[0].numpy()
                b = expm(F.numpy() * dt)
                assert np.max(np.abs(a - b)) <= 2e-13 * max(1.0, np.max(np.abs(b)))


def test_expm_tf_floor_variant_is_within_1e8():
    """Documents the size of TF's (recalled) floor-based under-scaling: <= 1e-8 absolute on the drifts used here."""
    with torch.no_grad():
        F = (O.Periodic(O.SquaredExponential(5., 1.), period=1., order=3) * O.Matern32(0.1, 50.)).get_sde().F
        A = torch.stack([F * dt for dt in (0.3, 0.5, 1.0, 2.0, 7.0)])
        exact = O.expm(A)
        O.EXPM_TF_FLOOR = True
        try:
            floor = O.expm(A)
        finally:
            O.EXPM_TF_FLOOR = False
    diff = float((exact - floor).abs().max())
    assert 1e-12 < diff < 1e-8


def test_stationary_q_equals_matrix_fraction_q():
    """Q = Pinf - A Pinf A^T (north star) vs the reference's matrix-fraction form, moderate steps."""
    t = np.cumsum(np.random.RandomState(0).uniform(0.002, 0.3, size=300))
    with torch.no_grad():
        for cov in (O.Matern32(1., .5), O.Matern52(1., 1.), O.RBF(1., 1., order=6, balancing_iter=5),
                    O.Matern52(1., 1.) + O.RBF(1., 1., order=6, balancing_iter=5)):
            sde = cov.get_sde()
            R = torch.tensor([[0.1]], dtype=torch.float64)
            a = O.get_ssm(sde, t[:, None], R)
            b = O.get_ssm_stationary(sde, t[:, None], R)
            assert float((a.Qs - b.Qs).abs().max()) <= 1e-12 * float(sde.P0.abs().max())
            assert float((a.Fs - b.Fs).abs().max()) == 0.0


def test_numba_sequential_kf_ks_matches_the_parallel_restatement():
    """oracle/seq_numba.py (numba restatement of pssgp/kalman/sequential.py:11-73, the 1-core CPU baseline of
    bench.py) against the torch restatement of the parallel path: kf == pkf, ks == pks."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import seq_numba
    from util import make_problem, rel_err
    for name, T in (("matern52", 300), ("m32+m52", 120)):
        t, y, cov, ssm = make_problem(name, T, seed=3)
        P0, Fs, Qs, H, R = [np.ascontiguousarray(x.numpy()) for x in ssm]
        fms, fPs, sms, sPs, ll = seq_numba.kfs(P0, Fs, Qs, H, R, y)
        with torch.no_grad():
            rfm, rfP, rll = O.pkf(ssm, y[:, None], True, max_parallel=T)
            rsm, rsP = O.pks(ssm, rfm, rfP, max_parallel=T)
        assert rel_err(fms, rfm) < 1e-11 and rel_err(fPs, rfP) < 1e-11
        assert abs(ll - float(rll)) < 1e-11 * abs(float(rll))
        assert rel_err(sms, rsm) < 1e-10 and rel_err(sPs, rsP) < 1e-10
