"""Native batched SDE construction (C ABI pssgp_sde_batch — HOST C++, so it runs without a GPU) against the oracle's
restatement of the reference's get_sde / balance_ss / solve_lyap_vec / SDESum / SDEProduct
(pssgp/kernels/*.py, math_utils.py:10-120, kernels/base.py:151-244), per hyper-parameter setting.
Tolerance 1e-12 relative to ||oracle matrix||_inf; 1e-9 for RBF order 15 (companion drift with entries ~1e9)."""
import numpy as np
import pytest
import torch

from util import O, pkg


def _rel(a, b):
    b = b.detach().numpy()
    assert np.all(np.isfinite(b)) and np.all(np.isfinite(a))
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


def _cases():
    pkg()
    from pssgp_b200 import kernels as PK
    per = lambda K, v, l, p, n: K.Periodic(K.SquaredExponential(v, l), period=p, order=n)
    return {
        "m12": (lambda K, v, l: K.Matern12(v, l), 2, 1e-12),
        "m32": (lambda K, v, l: K.Matern32(v, l), 2, 1e-12),
        "m52": (lambda K, v, l: K.Matern52(v, l), 2, 1e-12),
        "rbf3": (lambda K, v, l: K.RBF(v, l), 2, 1e-12),
        "rbf6": (lambda K, v, l: K.RBF(v, l, order=6, balancing_iter=5), 2, 1e-12),
        "rbf15": (lambda K, v, l: K.RBF(v, l, order=15, balancing_iter=10), 2, 1e-9),
        "periodic4": (lambda K, v, l, p: per(K, v, l, p, 4), 3, 1e-12),
        "m52+rbf6": (lambda K, v, l, v2, l2: K.Matern52(v, l) + K.RBF(v2, l2, order=6, balancing_iter=5), 4, 1e-12),
        "m32xm52": (lambda K, v, l, v2, l2: K.Matern32(v, l) * K.Matern52(v2, l2), 4, 1e-12),
        "qp3": (lambda K, v, l, p, v2, l2: per(K, v, l, p, 3) * K.Matern32(v2, l2), 5, 1e-12),
        "m32+m32xm32+qp2": (lambda K, a, b, c, d, e, f, g, h, i, j, k: K.Matern32(a, b) + K.Matern32(c, d) * K.Matern32(e, f)
                            + per(K, g, h, i, 2) * K.Matern32(j, k), 11, 1e-12),
    }, PK


@pytest.mark.parametrize("name", ["m12", "m32", "m52", "rbf3", "rbf6", "rbf15", "periodic4", "m52+rbf6", "m32xm52", "qp3",
                                  "m32+m32xm32+qp2"])
def test_native_sde_matches_oracle(name):
    cases, PK = _cases()
    from pssgp_b200.kernels import native
    mk, npar, tol = cases[name]
    rng = np.random.RandomState(len(name))
    B = 6
    P = rng.uniform(0.3, 2.5, size=(B, npar))
    specs, rows = zip(*(native.native_spec(mk(PK, *P[b])) for b in range(B)))
    assert all(s == specs[0] for s in specs)
    d, npar_spec = native.sde_dim(specs[0])
    assert npar_spec == npar
    for nthreads in (1, 0):
        F, Pinf, H = native.sde_batch(specs[0], np.asarray(rows), nthreads=nthreads)
        assert F.shape == (B, d, d) and Pinf.shape == (B, d, d) and H.shape == (B, d)
        for b in range(B):
            with torch.no_grad():
                s = mk(O, *P[b]).get_sde()
            assert _rel(F[b], s.F) < tol and _rel(Pinf[b], s.P0) < tol and _rel(H[b], s.H.reshape(-1)) < tol


def test_native_sde_equals_python_host_layer():
    """The package's own per-setting get_sde (torch, differentiable) and the native batch agree."""
    cases, PK = _cases()
    from pssgp_b200.kernels import native
    k = PK.Matern52(1.3, 0.7) + PK.RBF(0.9, 1.9, order=6, balancing_iter=5)
    spec, row = native.native_spec(k)
    F, Pinf, H = native.sde_batch(spec, [row])
    with torch.no_grad():
        s = k.get_sde()
    assert _rel(F[0], s.F) < 1e-12 and _rel(Pinf[0], s.P0) < 1e-12 and _rel(H[0], s.H.reshape(-1)) < 1e-12


def test_native_spec_grammar_and_errors():
    cases, PK = _cases()
    from pssgp_b200 import _lib
    from pssgp_b200.kernels import native
    # outside the grammar: a product of a sum
    nested = (PK.Matern32(1., 1.) + PK.Matern52(1., 1.)) * PK.Matern32(1., 1.)
    assert native.native_spec(nested) is None
    spec, row = native.native_spec(PK.Matern32(1., 2.))
    assert spec[1:] == [1, 1, native.MATERN32, 0, 0] and row == [1., 2.]
    with pytest.raises(ValueError):
        native.sde_batch(spec, [[1., 2., 3.]])
    with pytest.raises(_lib.PssgpError):
        native.sde_dim([10, 1, 1, 99, 0, 0])      # unknown kernel type
    with pytest.raises(_lib.PssgpError):
        native.sde_dim([10, 2, 1, 1, 0, 0])       # truncated
    # a bare Periodic inside a sum has a singular Lyapunov system (zero-frequency block): reported, like the
    # reference's tf.linalg.solve failure
    bad = PK.Matern32(1., 1.) + PK.Periodic(PK.SquaredExponential(1., 1.), period=1., order=2)
    spec, row = native.native_spec(bad)
    with pytest.raises(_lib.PssgpError):
        native.sde_batch(spec, [row])


@pytest.mark.parametrize("d", [1, 3, 6, 9, 16])
def test_native_lyapunov_solve_and_adjoint(d):
    """solve_lyap_vec of the package (native pssgp_lyap_solve forward, the same routine with F^T as analytic adjoint)
    against the oracle's Kronecker solve differentiated by torch autograd (math_utils.py:84-120)."""
    pkg()
    from pssgp_b200.kernels.math_utils import _lyap, solve_lyap_vec
    gen = torch.Generator().manual_seed(d)
    F = (-2 * torch.eye(d, dtype=torch.float64) + 0.5 * torch.randn(d, d, dtype=torch.float64, generator=gen) / d ** 0.5)
    F.requires_grad_(True)
    L = torch.randn(d, 2, dtype=torch.float64, generator=gen).requires_grad_(True)
    Q = torch.tensor([[1.0, 0.2], [0.2, 2.0]], dtype=torch.float64).requires_grad_(True)
    W = torch.randn(d, d, dtype=torch.float64, generator=gen)
    P = solve_lyap_vec(F, L, Q)
    Po = O.solve_lyap_vec(F, L, Q)
    assert float((P - Po).abs().max() / Po.abs().max()) < 1e-12
    g = torch.autograd.grad((P * W).sum(), (F, L, Q))
    go = torch.autograd.grad((Po * W).sum(), (F, L, Q))
    for a, b in zip(g, go):
        assert float((a - b).abs().max() / b.abs().max()) < 1e-11
    # general (non-symmetric) right-hand side: the full Kronecker system
    G = torch.randn(d, d, dtype=torch.float64, generator=gen).numpy()
    X = _lyap(F.detach().numpy(), G)
    Fn = F.detach().numpy()
    assert np.max(np.abs(Fn @ X + X @ Fn.T - G)) < 1e-12 * max(1.0, np.max(np.abs(X)))


@pytest.mark.parametrize("name", ["m12", "m32", "m52", "rbf6", "periodic4", "m52+rbf6", "m32xm52", "qp3", "m32+m32xm32+qp2"])
def test_native_sde_jacobian_matches_autograd(name):
    """pssgp_sde_batch_jac (forward-mode duals in C++) and the differentiable wrapper kernels.native.NativeSDE against
    torch autograd through the package's Python get_sde (itself pinned to the oracle in test_cpu_host.py)."""
    cases, PK = _cases()
    from pssgp_b200 import config as cfg
    from pssgp_b200.kernels import native
    mk, npar, _ = cases[name]
    rng = np.random.RandomState(3 + len(name))
    P = rng.uniform(0.4, 2.0, size=npar)
    old = cfg.FAST_MATERN_SDE
    cfg.FAST_MATERN_SDE = False
    try:
        k = mk(PK, *P)
        sde = k.get_sde()
        gen = torch.Generator().manual_seed(1)
        W1 = torch.randn(sde.F.shape, dtype=torch.float64, generator=gen)
        W2 = torch.randn(sde.P0.shape, dtype=torch.float64, generator=gen)
        W3 = torch.randn(sde.H.shape, dtype=torch.float64, generator=gen)
        ps = native.native_parameters(k)
        ref = torch.autograd.grad((sde.F * W1).sum() + (sde.P0 * W2).sum() + (sde.H * W3).sum(),
                                  [p.unconstrained_variable for p in ps], allow_unused=True)
        F, Pinf, H = native.native_sde(k)
        assert _rel(F.detach().numpy(), sde.F.detach()) < 1e-12 and _rel(Pinf.detach().numpy(), sde.P0.detach()) < 1e-12
        got = torch.autograd.grad((F * W1).sum() + (Pinf * W2).sum() + (H * W3).sum(),
                                  [p.unconstrained_variable for p in ps], allow_unused=True)
    finally:
        cfg.FAST_MATERN_SDE = old
    scale = max(abs(float(g)) for g in ref if g is not None)
    for a, b in zip(got, ref):
        a = 0.0 if a is None else float(a)
        b = 0.0 if b is None else float(b)
        assert abs(a - b) <= 1e-10 * scale


def test_native_spec_ignores_subclasses():
    """A user subclass may override get_sde: only the exact kernel classes take the native construction."""
    cases, PK = _cases()
    from pssgp_b200.kernels import native

    class MyMatern(PK.Matern32):
        pass

    assert native.native_spec(MyMatern(1., 1.)) is None
    assert native.native_spec(PK.Matern32(1., 1.) + MyMatern(1., 1.)) is None
    assert native.native_sde(MyMatern(1., 1.)) is None
