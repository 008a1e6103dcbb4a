"""Grid log-likelihood (BASELINE configs[4]b) on the GPU: the native batched path (pssgp_sde_batch + pssgp_grid_loglik)
against the oracle's log-likelihood per setting (<= 1e-7 relative; observed ~1e-12) and against the
per-setting Python path of the same package."""
import numpy as np
import pytest
import torch

from util import O, pkg

pytestmark = pytest.mark.gpu


def _series(T, seed=0):
    rng = np.random.RandomState(seed)
    t = np.cumsum(0.01 * rng.uniform(0.5, 1.5, size=T))
    y = O.obs_noise(O.sinu(t), 0.1, seed)
    y[rng.choice(T, T // 50, replace=False)] = np.nan
    return t, y


CASES = {
    "m52+rbf6": (lambda K, a, b: K.Matern52(1.0, float(a)) + K.RBF(1.0, float(b), order=6, balancing_iter=5), 1e-7),
    "matern32": (lambda K, a, b: K.Matern32(float(a), float(b)), 1e-9),
    "qp2": (lambda K, a, b: K.Periodic(K.SquaredExponential(1.0, float(a)), period=1.0, order=2) * K.Matern32(0.5, float(b)),
            1e-7),
}


@pytest.mark.parametrize("name", list(CASES))
def test_grid_matches_oracle(name):
    pkg()
    from pssgp_b200 import batch, kernels as PK
    mk, tol = CASES[name]
    T = 700
    t, y = _series(T)
    ls = np.logspace(-0.7, 0.7, 4)
    settings = [(a, b) for a in ls for b in ls]
    data = (t[:, None], y[:, None])
    ll_native = batch.grid_log_likelihood(lambda a, b: mk(PK, a, b), settings, data, 0.1)
    ll_python = batch.grid_log_likelihood(lambda a, b: mk(PK, a, b), settings, data, 0.1, native=False)
    assert ll_native.shape == (len(settings),) and bool(torch.isfinite(ll_native).all())
    ref = []
    for a, b in settings:
        with torch.no_grad():
            ssm = mk(O, a, b).get_ssm(t[:, None], torch.tensor([[0.1]], dtype=torch.float64))
            ref.append(float(O.pkf(ssm, y[:, None], True, max_parallel=T)[2]))
    ref = np.asarray(ref)
    assert np.max(np.abs(ll_native.numpy() - ref) / np.abs(ref)) < tol
    assert np.max(np.abs(ll_native.numpy() - ll_python.numpy()) / np.abs(ref)) < 1e-9
    assert int(np.argmax(ll_native.numpy())) == int(np.argmax(ref))


def test_grid_noise_callable_and_fallback_for_nested_kernels():
    pkg()
    from pssgp_b200 import batch, kernels as PK
    t, y = _series(300, seed=1)
    data = (t[:, None], y[:, None])
    settings = [(0.5, 0.05), (1.0, 0.1), (2.0, 0.2)]
    mk = lambda ell, nv: PK.Matern52(1.0, float(ell))
    ll = batch.grid_log_likelihood(mk, settings, data, lambda ell, nv: nv)
    for (ell, nv), v in zip(settings, ll):
        with torch.no_grad():
            ssm = O.Matern52(1.0, ell).get_ssm(t[:, None], torch.tensor([[nv]], dtype=torch.float64))
            r = float(O.pkf(ssm, y[:, None], True, max_parallel=300)[2])
        assert abs(float(v) - r) < 1e-9 * abs(r)
    # a product of a sum is outside the native grammar: the per-setting path takes over, same results as the oracle
    nested = lambda a, b: (PK.Matern32(1.0, float(a)) + PK.Matern52(1.0, float(b))) * PK.Matern32(1.0, 1.0)
    ll2 = batch.grid_log_likelihood(nested, [(0.5, 1.0), (1.0, 2.0)], data, 0.1)
    assert bool(torch.isfinite(ll2).all())


@pytest.mark.parametrize("name", ["m52+rbf6", "matern32", "qp2"])
def test_grid_gradient_matches_model_autograd(name):
    """grid_log_likelihood(with_grad=True): per setting the gradient w.r.t. the constrained kernel hyper-parameters and
    the noise variance equals the one StateSpaceGP + autograd gives for that setting (itself pinned to the oracle)."""
    pkg()
    from pssgp_b200 import batch, kernels as PK
    from pssgp_b200.kernels import native
    from pssgp_b200.model import StateSpaceGP
    mk, _ = CASES[name]
    T = 500
    t, y = _series(T, seed=2)
    settings = [(0.4, 0.7), (1.1, 0.5), (0.8, 1.6), (2.0, 0.9), (0.6, 0.6)]
    data = (t[:, None], y[:, None])
    ll, dparams, dnoise = batch.grid_log_likelihood(lambda a, b: mk(PK, a, b), settings, data, 0.1, with_grad=True)
    ll_only = batch.grid_log_likelihood(lambda a, b: mk(PK, a, b), settings, data, 0.1)
    assert torch.allclose(ll, ll_only, rtol=1e-12, atol=0)
    for i, (a, b) in enumerate(settings):
        k = mk(PK, a, b)
        model = StateSpaceGP(data, k, noise_variance=0.1, parallel=True)
        val = model.maximum_log_likelihood_objective()
        ps = native.native_parameters(k)
        g = torch.autograd.grad(val, [p.unconstrained_variable for p in ps] + [model.noise_variance.unconstrained_variable],
                                allow_unused=True)
        # d / d constrained = (d / d unconstrained) / sigmoid(unconstrained)
        con = [0.0 if gi is None else float(gi) / float(torch.sigmoid(p.unconstrained_variable.detach()))
               for gi, p in zip(g, ps + [model.noise_variance])]
        scale = max(abs(c) for c in con)
        assert abs(float(ll[i]) - float(val)) <= 1e-10 * abs(float(val))
        np.testing.assert_allclose(dparams[i].numpy(), con[:-1], rtol=1e-7, atol=1e-9 * scale)
        assert abs(float(dnoise[i]) - con[-1]) <= 1e-7 * abs(con[-1]) + 1e-9 * scale
