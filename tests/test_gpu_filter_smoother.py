"""GPU parity: pkf / pks / pkfs through the C ABI against the CPU oracle (FP64: <= 1e-9 relative to
||reference||_inf; FP32: stated per test)."""
import numpy as np
import pytest
import torch

from util import O, make_problem, pkg, rel_err, ssm_numpy

pytestmark = pytest.mark.gpu

TOL64 = 1e-9


def _api():
    pkg()
    from pssgp_b200.kalman import parallel
    return parallel


@pytest.mark.parametrize("name", ["matern12", "matern32", "matern52", "m32xm32"])
@pytest.mark.parametrize("T", [1, 2, 3, 31, 32, 33, 127, 129, 1000, 4097, 20011])
def test_pkf_pks_parity_fp64(name, T):
    api = _api()
    t, y, cov, ssm = make_problem(name, T, seed=T)
    with torch.no_grad():
        rfm, rfP, rll = O.pkf(ssm, y[:, None], return_loglikelihood=True)
        rsm, rsP = O.pks(ssm, rfm, rfP)
    lg = ssm_numpy(ssm)
    fm, fP, ll = api.pkf(lg, y[:, None], return_loglikelihood=True)
    assert fm.shape == (T, lg[0].shape[0]) and fP.shape == (T,) + lg[0].shape
    assert rel_err(fm, rfm) < TOL64
    assert rel_err(fP, rfP) < TOL64
    assert abs(ll - float(rll)) <= TOL64 * max(1.0, abs(float(rll)))
    sm, sP = api.pks(lg, rfm.numpy(), rfP.numpy())
    assert rel_err(sm, rsm) < TOL64
    assert rel_err(sP, rsP) < TOL64
    sm2, sP2 = api.pkfs(lg, y[:, None])
    assert rel_err(sm2, rsm) < TOL64
    assert rel_err(sP2, rsP) < TOL64


@pytest.mark.parametrize("chunk", [1, 2, 5, 8, 64, 257])
def test_chunk_length_invariance(chunk):
    """The result must not depend on how the time axis is cut into chunks."""
    api = _api()
    from pssgp_b200 import _lib
    t, y, cov, ssm = make_problem("matern52", 5003, seed=3)
    with torch.no_grad():
        rfm, rfP, rll = O.pkf(ssm, y[:, None], return_loglikelihood=True)
        rsm, rsP = O.pks(ssm, rfm, rfP)
    h = _lib.handle(torch.cuda.current_device())
    h.set_option("chunk", chunk)
    try:
        lg = ssm_numpy(ssm)
        fm, fP, ll = api.pkf(lg, y[:, None], return_loglikelihood=True)
        sm, sP = api.pks(lg, fm, fP)
    finally:
        h.set_option("chunk", 0)
    assert rel_err(fm, rfm) < TOL64 and rel_err(fP, rfP) < TOL64
    assert abs(ll - float(rll)) <= TOL64 * abs(float(rll))
    assert rel_err(sm, rsm) < TOL64 and rel_err(sP, rsP) < TOL64


def test_all_nan_and_first_nan():
    api = _api()
    t, y, cov, ssm = make_problem("matern32", 300, seed=1, nan_frac=0.0)
    for variant in ("first", "last", "all"):
        yy = y.copy()
        if variant == "first":
            yy[0] = np.nan
        elif variant == "last":
            yy[-1] = np.nan
        else:
            yy[:] = np.nan
        with torch.no_grad():
            rfm, rfP, rll = O.pkf(ssm, yy[:, None], return_loglikelihood=True)
            rsm, rsP = O.pks(ssm, rfm, rfP)
        fm, fP, ll = api.pkf(ssm_numpy(ssm), yy[:, None], return_loglikelihood=True)
        sm, sP = api.pks(ssm_numpy(ssm), fm, fP)
        assert rel_err(fm, rfm) < TOL64 or float(np.max(np.abs(rfm.numpy()))) == 0.0
        assert rel_err(fP, rfP) < TOL64
        assert abs(ll - float(rll)) <= TOL64 * max(1.0, abs(float(rll)))
        assert rel_err(sP, rsP) < TOL64
        assert np.max(np.abs(sm - rsm.numpy())) < 1e-9


def test_device_tensors_stay_on_device():
    api = _api()
    t, y, cov, ssm = make_problem("matern52", 2000, seed=5)
    dev = torch.device("cuda", 0)
    lg = tuple(x.to(dev) for x in ssm)
    yd = torch.as_tensor(y[:, None]).to(dev)
    fm, fP, ll = api.pkf(lg, yd, return_loglikelihood=True)
    assert fm.is_cuda and fP.is_cuda and ll.is_cuda
    with torch.no_grad():
        rfm, rfP, rll = O.pkf(ssm, y[:, None], return_loglikelihood=True)
    assert rel_err(fm.cpu().numpy(), rfm) < TOL64
    assert abs(float(ll) - float(rll)) <= TOL64 * abs(float(rll))


def test_fp32_parity():
    """FP32 opt-in mode: tolerance 2e-3 relative to ||reference||_inf (float32 eps 1.2e-7 amplified by the
    T=4000-step recursion and the I + C J solves)."""
    api = _api()
    t, y, cov, ssm = make_problem("matern32", 4000, seed=7)
    with torch.no_grad():
        rfm, rfP, rll = O.pkf(ssm, y[:, None], return_loglikelihood=True)
        rsm, rsP = O.pks(ssm, rfm, rfP)
    lg = ssm_numpy(ssm, np.float32)
    fm, fP, ll = api.pkf(lg, y[:, None].astype(np.float32), return_loglikelihood=True)
    assert fm.dtype == np.float32
    sm, sP = api.pkfs(lg, y[:, None].astype(np.float32))
    assert rel_err(fm, rfm) < 2e-3 and rel_err(fP, rfP) < 2e-3
    assert abs(ll - float(rll)) <= 2e-3 * abs(float(rll))
    assert rel_err(sm, rsm) < 2e-3 and rel_err(sP, rsP) < 2e-3


def test_errors_are_loud():
    api = _api()
    t, y, cov, ssm = make_problem("matern32", 10, seed=1)
    lg = list(ssm_numpy(ssm))
    with pytest.raises(ValueError):
        api.pkf(lg, y[:5, None])
    bad = list(lg)
    bad[3] = np.ones((2, 2))
    with pytest.raises(ValueError):
        api.pkf(bad, y[:, None])
