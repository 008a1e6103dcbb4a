"""CPU: host logic of the product package and the C-ABI surface (no GPU compute)."""
import ctypes
import os
import re

import numpy as np
import numpy.testing as npt
import pytest
import torch

from util import O, ROOT, pkg


def _pk():
    pkg()
    from pssgp_b200 import kernels
    return kernels


def test_library_loads_and_exports_every_declared_symbol():
    pkg()
    from pssgp_b200 import _lib
    header = open(os.path.join(ROOT, "include", "pssgp_b200.h")).read()
    declared = set(re.findall(r"\b(pssgp_[a-z0-9_]+)\s*\(", header))
    declared.discard("pssgp_handle")
    assert len(declared) >= 18
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"libpssgp_b200.so does not export {name}"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert _lib.load_library().pssgp_version() >= 100


def test_missing_library_fails_loudly(tmp_path):
    pkg()
    from pssgp_b200 import _lib
    with pytest.raises(ImportError):
        _lib.load_library(str(tmp_path / "nope.so"))


def test_no_cuda_fails_loudly():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    pkg()
    from pssgp_b200.kalman.parallel import pkf
    lg = (np.eye(2), np.zeros((3, 2, 2)), np.zeros((3, 2, 2)), np.ones((1, 2)), np.ones((1, 1)))
    with pytest.raises(RuntimeError):
        pkf(lg, np.zeros((3, 1)))


def test_product_package_never_imports_oracle():
    for root, _, files in os.walk(os.path.join(ROOT, "parallel-gps_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(root, f)).read()
                assert "pssgp_oracle" not in src and "import_oracle" not in src, f


def test_balance_ss_host_routine_matches_reference_algorithm():
    """pssgp_balance_ss (compiled) vs the restated numba routine (math_utils.py:10-29)."""
    pkg()
    from pssgp_b200.kernels.math_utils import _balance_d
    rng = np.random.RandomState(0)
    for d in (2, 3, 6, 15):
        F = rng.randn(d, d) * np.exp(rng.randn(d, d) * 2)
        for it in (0, 1, 5, 10):
            npt.assert_allclose(_balance_d(torch.as_tensor(F), it), O._balance_ss_d(F, it), rtol=1e-13)


KERNELS = ["matern12", "matern32", "matern52", "rbf3", "rbf6", "periodic2", "periodic6", "sum", "prod", "qp"]


def _both(name):
    PK = _pk()
    if name == "matern12":
        return O.Matern12(1.3, .7), PK.Matern12(1.3, .7)
    if name == "matern32":
        return O.Matern32(1.3, .7), PK.Matern32(1.3, .7)
    if name == "matern52":
        return O.Matern52(1.3, .7), PK.Matern52(1.3, .7)
    if name == "rbf3":
        return O.RBF(1., .1, order=3, balancing_iter=5), PK.RBF(1., .1, order=3, balancing_iter=5)
    if name == "rbf6":
        return O.RBF(2., 1.5, order=6, balancing_iter=5), PK.RBF(2., 1.5, order=6, balancing_iter=5)
    if name == "periodic2":
        return (O.Periodic(O.SquaredExponential(1., .1), period=1., order=2),
                PK.Periodic(PK.SquaredExponential(1., .1), period=1., order=2))
    if name == "periodic6":
        return (O.Periodic(O.SquaredExponential(5., 1.), period=.7),
                PK.Periodic(PK.SquaredExponential(5., 1.), period=.7))
    if name == "sum":
        return O.Matern32(1., .5) + O.Matern52(1., .5), PK.Matern32(1., .5) + PK.Matern52(1., .5)
    if name == "prod":
        return O.Matern32(1., .5) * O.Matern52(1., .5), PK.Matern32(1., .5) * PK.Matern52(1., .5)
    if name == "qp":
        return (O.Periodic(O.SquaredExponential(5., 1.), period=1., order=3) * O.Matern32(.1, 50.) + O.Matern32(1., 100.),
                PK.Periodic(PK.SquaredExponential(5., 1.), period=1., order=3) * PK.Matern32(.1, 50.) + PK.Matern32(1., 100.))
    raise KeyError(name)


@pytest.mark.parametrize("name", KERNELS)
def test_get_sde_matches_oracle_and_is_differentiable(name):
    ocov, pcov = _both(name)
    osde, psde = ocov.get_sde(), pcov.get_sde()
    for a, b in zip(osde, psde):
        assert a.shape == b.shape
        npt.assert_allclose(b.detach().numpy(), a.detach().numpy(), rtol=1e-12, atol=1e-14 * float(a.abs().max() + 1))
    # same hyper-parameter gradient of a scalar functional of the SDE
    f = lambda s: (s.P0 * s.P0).sum() + (s.F * s.F).sum() * 1e-3 + s.H.sum() + s.Q.sum()
    go = torch.autograd.grad(f(osde), ocov.trainable_variables, allow_unused=True)
    gp = torch.autograd.grad(f(psde), pcov.trainable_variables, allow_unused=True)
    assert len(go) == len(gp)
    for a, b in zip(go, gp):
        if a is None:
            assert b is None
        else:
            # "qp": nested Lyapunov systems with cond 2e5 each (lengthscales 50 and 100); the oracle differentiates
            # through LAPACK's LU of the Kronecker matrix, the package uses the analytic adjoint (a second native
            # solve with F^T): two roundings of an ill-conditioned gradient, 4e-8 apart
            npt.assert_allclose(b.item(), a.item(), rtol=1e-6 if name == "qp" else 1e-10, atol=1e-12)
    spec = pcov.get_spec(17)
    assert spec.Fs.shape == (17, psde.F.shape[0], psde.F.shape[0]) and spec.H.shape == (1, psde.F.shape[0])


def test_known_answers_hold_for_product_kernels():
    """reference tests/test_rbf.py:27-47 and tests/test_periodic.py:43-50 against the PRODUCT host code."""
    PK = _pk()
    with torch.no_grad():
        Pinf, F, L, H, Q = PK.RBF(1., 0.1, order=3, balancing_iter=5).get_sde()
    npt.assert_array_almost_equal(F, np.array([[0, 14.520676967550859, 0], [0, 0, 32.857489440296360],
                                               [-14.5210953665873, -29.4746060478111, -50.3678777987092]]), decimal=8)
    npt.assert_array_almost_equal(Q, 52.8553179255264, decimal=8)
    npt.assert_array_almost_equal(np.diag(Pinf), [1.04502531824891, 0.681741999944955, 0.611552410634913], decimal=8)
    with torch.no_grad():
        Pinf, F, L, H, Q = PK.Periodic(PK.SquaredExponential(1., 0.1), period=1., order=2).get_sde()
    assert F[2, 3].item() == pytest.approx(-6.283185307179586) and F[5, 4].item() == pytest.approx(12.5663706143592)
    npt.assert_almost_equal(H, np.array([[1, 0, 1, 0, 1, 0]]))


def test_sum_with_uncoupled_state_is_finite():
    """the reference's balancing divides 0/0 for a 1x1 block inside a Sum; the compiled routine skips it."""
    PK = _pk()
    with torch.no_grad():
        sde = (PK.Matern12(1., .5) + PK.Matern32(1., .5)).get_sde()
    assert torch.isfinite(sde.P0).all() and torch.isfinite(sde.F).all()
    npt.assert_allclose(sde.P0[0, 0].item(), 1.0, rtol=1e-12)


def test_parameter_softplus_and_prior():
    pkg()
    from pssgp_b200.params import Parameter, set_trainable
    p = Parameter(0.37)
    assert p.value.item() == pytest.approx(0.37, rel=1e-14)
    p.assign(2.5)
    assert p.numpy() == pytest.approx(2.5, rel=1e-14)
    p.prior = lambda x: -0.5 * (x - 1.0) ** 2
    lp = p.log_prior_density()
    expect = -0.5 * 1.5 ** 2 + torch.nn.functional.logsigmoid(p.unconstrained_variable).item()
    assert lp.item() == pytest.approx(expect, rel=1e-12)
    set_trainable(p, False)
    assert not p.unconstrained_variable.requires_grad


def test_merge_sorted_matches_argsort():
    """model.py:15-55 semantics: stable merge of two sorted arrays plus companions, ties put queries first
    (searchsorted side='left')."""
    pkg()
    from pssgp_b200.model import _merge_sorted
    rng = np.random.RandomState(0)
    for na, nb in ((10, 3), (3, 10), (7, 7), (1, 5), (100, 1)):
        a = torch.as_tensor(np.sort(rng.rand(na)))
        b = torch.as_tensor(np.sort(rng.rand(nb)))
        ya, yb = torch.arange(na, dtype=torch.float64)[:, None], -torch.arange(1, nb + 1, dtype=torch.float64)[:, None]
        t, y, f = _merge_sorted(a, b, (ya, yb), (torch.zeros(na, dtype=torch.bool), torch.ones(nb, dtype=torch.bool)))
        ref_t, ref_y, ref_f = O.merge_sorted(a, b, (ya, yb), (torch.zeros(na, dtype=torch.bool), torch.ones(nb, dtype=torch.bool)))
        assert torch.equal(t, ref_t) and torch.equal(y, ref_y) and torch.equal(f, ref_f)
        assert torch.all(t[1:] >= t[:-1]) and int(f.sum()) == nb
    # duplicates: the toy configuration predicts at the training times themselves
    a = torch.linspace(0, 1, 5, dtype=torch.float64)
    t, f = _merge_sorted(a, a.clone(), (torch.zeros(5, dtype=torch.bool), torch.ones(5, dtype=torch.bool)))
    assert torch.all(t[1:] >= t[:-1]) and int(f.sum()) == 5


def test_merge_sorted_idx_positions_and_batch_sharding():
    """The sync-free merge returns where each input element went (used by predict_f to pick the query rows without a
    boolean mask, pssgp/model.py:107-108), for both argument orders; grid settings shard round-robin."""
    pkg()
    from pssgp_b200.batch import shard_indices
    from pssgp_b200.model import _merge_sorted_idx
    rng = np.random.RandomState(3)
    for na, nb in ((12, 5), (5, 12), (6, 6), (1, 4)):
        a = torch.as_tensor(np.sort(rng.rand(na)))
        b = torch.as_tensor(np.sort(rng.rand(nb)))
        (t,), (ia, ib) = _merge_sorted_idx(a, b)
        assert torch.equal(t[ia], a) and torch.equal(t[ib], b)
        assert sorted(ia.tolist() + ib.tolist()) == list(range(na + nb))
        assert torch.all(t[1:] >= t[:-1])
    # ties: the shorter array goes first (searchsorted side='left' in the reference's scatter)
    a = torch.tensor([0.0, 1.0, 2.0, 3.0], dtype=torch.float64)
    b = torch.tensor([1.0, 3.0], dtype=torch.float64)
    (t,), (ia, ib) = _merge_sorted_idx(a, b)
    assert ib.tolist() == [1, 4] and ia.tolist() == [0, 2, 3, 5]
    assert [shard_indices(10, r, 4) for r in range(4)] == [[0, 4, 8], [1, 5, 9], [2, 6], [3, 7]]


def test_dlpack_producers_are_recognised_on_the_host_side():
    """_arrays.from_dlpack / the DLPack branch of the host API accept any __dlpack__ producer and raw capsules
    (CPU producers here; the zero-copy device case is tests/test_gpu_dlpack.py)."""
    import numpy as np
    import torch
    from torch.utils.dlpack import to_dlpack
    pkg()
    from pssgp_b200 import _arrays as A
    from pssgp_b200.kalman.parallel import _dl

    class Producer:
        def __init__(self, t):
            self._t = t

        def __dlpack__(self, stream=None):
            return self._t.__dlpack__()

        def __dlpack_device__(self):
            return self._t.__dlpack_device__()

    x = torch.arange(12, dtype=torch.float64).reshape(3, 4)
    for obj in (Producer(x), to_dlpack(x)):
        v = _dl(obj)
        assert isinstance(v, torch.Tensor) and v.data_ptr() == x.data_ptr() and torch.equal(v, x)
    a = np.arange(5.0)
    assert _dl(a) is a and _dl(x) is x          # numpy arrays and torch tensors pass through untouched
    assert torch.equal(A.from_dlpack(Producer(x)), x)


def test_matern52_closed_form_sde_equals_generic_path():
    """The closed-form Matern52.get_sde (analytic gradient) against the generic mirror of matern52.py:21-25 (balance +
    Lyapunov solve under torch autograd): values to 1e-13, vector-Jacobian products to 1e-11."""
    import torch
    pkg()
    from pssgp_b200 import config as C, kernels as PK
    for var, ell in ((1.0, 1.0), (0.37, 2.9), (5.0, 0.11)):
        outs = {}
        for fast in (True, False):
            C.FAST_MATERN_SDE = fast
            try:
                k = PK.Matern52(var, ell)
                sde = k.get_sde()
                gen = torch.Generator().manual_seed(7)
                w = [torch.randn(x.shape, dtype=torch.float64, generator=gen) for x in (sde.P0, sde.F, sde.L, sde.H, sde.Q)]
                s = sum((a * b).sum() for a, b in zip(w, (sde.P0, sde.F, sde.L, sde.H, sde.Q)))
                g = torch.autograd.grad(s, [p.unconstrained_variable for p in k.parameters])
                outs[fast] = ([x.detach() for x in (sde.P0, sde.F, sde.L, sde.H, sde.Q)], [x.detach() for x in g])
            finally:
                C.FAST_MATERN_SDE = True
        for a, b in zip(outs[True][0], outs[False][0]):
            assert float((a - b).abs().max()) <= 1e-13 * max(1.0, float(b.abs().max()))
        for a, b in zip(outs[True][1], outs[False][1]):
            assert float((a - b).abs().max()) <= 1e-11 * max(1.0, float(b.abs().max()))


def test_toymodels_match_reference_vectors():
    """pssgp_b200.toymodels against vectors generated from the reference's own pssgp.toymodels (scripts/make_golden.py)."""
    import os
    pkg()
    from pssgp_b200 import toymodels as TM
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "toy_sinusoid_n1000.npz"))
    t = g["t"]
    assert np.array_equal(TM.sinu(t), g["ft"]) and np.array_equal(TM.comp_sinu(t), g["comp"]) and np.array_equal(TM.rect(t), g["rect"])
    assert np.array_equal(TM.obs_noise(TM.sinu(t), 0.1, 0), g["y"])
    assert np.array_equal(TM.obs_noise(TM.sinu(t), 0.5, 666), g["y_seed666"])
