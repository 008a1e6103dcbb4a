"""MAP fit through the gpflow.optimizers.Scipy stand-in (the reference's MAP experiments, sunspot/map.py:74-82): every
evaluation is a fused filter + adjoint step on the GPU; the optimum equals the one found on the oracle's objective."""
import numpy as np
import pytest
import torch

from util import O, pkg

pytestmark = pytest.mark.gpu


def test_map_fit_matches_oracle_optimum():
    pkg()
    from pssgp_b200 import kernels as PK
    from pssgp_b200.model import StateSpaceGP
    from pssgp_b200.optimizers import Scipy
    rng = np.random.RandomState(3)
    T = 400
    t = np.sort(rng.uniform(0, 6, T))
    y = O.obs_noise(O.sinu(t), 0.1, 4)
    model = StateSpaceGP((t[:, None], y[:, None]), PK.Matern32(2.0, 0.3), noise_variance=0.5, parallel=True)
    opt = Scipy()
    f = opt.eval_func(model.training_loss, model.trainable_variables)
    x0 = opt.initial_parameters(model.trainable_variables)
    l0, g0 = f(x0)
    # the gradient is the one of the oracle's objective at the same point
    ocov = O.Matern32(2.0, 0.3)
    oss = O.StateSpaceGP((t, y), ocov, 0.5, parallel=True, max_parallel=T)
    oll = -oss.maximum_log_likelihood_objective()
    og = torch.autograd.grad(oll, ocov.trainable_variables + [oss.noise_variance_p.unconstrained])
    assert abs(l0 - float(oll)) < 1e-9 * abs(float(oll))
    np.testing.assert_allclose(g0, [float(v) for v in og], rtol=1e-7, atol=1e-9)
    res = opt.minimize(model.training_loss, model.trainable_variables, options=dict(maxiter=60))
    assert res.fun < l0 - 1.0 and np.linalg.norm(res.jac) < 1e-2 * max(1.0, np.linalg.norm(g0))
    # same optimisation on the oracle's objective
    import scipy.optimize
    ovars = ocov.trainable_variables + [oss.noise_variance_p.unconstrained]

    def of(x):
        with torch.no_grad():
            for v, xi in zip(ovars, x):
                v.fill_(float(xi))
        loss = -oss.maximum_log_likelihood_objective()
        g = torch.autograd.grad(loss, ovars)
        return float(loss), np.array([float(v) for v in g])

    ores = scipy.optimize.minimize(of, x0, jac=True, method="L-BFGS-B", options=dict(maxiter=60))
    assert abs(res.fun - ores.fun) < 1e-6 * abs(ores.fun)
    np.testing.assert_allclose(res.x, ores.x, rtol=1e-3, atol=1e-3)
    # the model holds the optimum
    assert abs(float(model.training_loss()) - res.fun) < 1e-9 * abs(res.fun)
