"""DLPack ingestion at the boundary (north star: "zero-copy DLPack views of TF tensors"): any ``__dlpack__`` producer
or raw capsule is accepted by the host API, device memory is not copied, and the capsule-level binding a TensorFlow
maintainer wraps (pssgp_b200.dlpack_binding) reproduces the oracle."""
import numpy as np
import pytest
import torch
from torch.utils.dlpack import from_dlpack, to_dlpack

from util import O, make_problem, pkg, rel_err

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)


class Producer:
    """A non-torch object that only speaks the DLPack protocol (stands in for tf.Tensor / cupy.ndarray)."""

    def __init__(self, t):
        self._t = t

    def __dlpack__(self, stream=None):
        return self._t.__dlpack__(stream=stream) if stream is not None else self._t.__dlpack__()

    def __dlpack_device__(self):
        return self._t.__dlpack_device__()


def test_to_device_is_zero_copy_for_device_dlpack_producers():
    pkg()
    from pssgp_b200 import _arrays as A
    x = torch.arange(64, dtype=torch.float64, device=DEV)
    for obj in (Producer(x), to_dlpack(x)):
        v = A.to_device(obj, torch.float64, DEV)
        assert v.data_ptr() == x.data_ptr() and v.is_cuda
    h = np.arange(16.0)
    v = A.to_device(Producer(torch.from_numpy(h)), torch.float64, DEV)  # host producer: staged H2D copy
    assert v.is_cuda and np.array_equal(v.cpu().numpy(), h)


def test_host_api_accepts_dlpack_producers():
    pkg()
    from pssgp_b200.kalman.parallel import pkf, pks
    t, y, cov, ssm = make_problem("matern52", 700, seed=3)
    with torch.no_grad():
        rfm, rfP, rll = O.pkf(ssm, y[:, None], True)
        rsm, rsP = O.pks(ssm, rfm, rfP)
    lg = tuple(Producer(x.detach().to(DEV).contiguous()) for x in ssm)
    fm, fP, ll = pkf(lg, Producer(torch.as_tensor(y[:, None]).to(DEV)), return_loglikelihood=True)
    assert isinstance(fm, torch.Tensor) and fm.is_cuda           # device producers in -> device results out
    assert rel_err(fm.cpu(), rfm) < 1e-9 and rel_err(fP.cpu(), rfP) < 1e-9 and abs(float(ll) - float(rll)) < 1e-9 * abs(float(rll))
    sm, sP = pks(lg, Producer(fm), Producer(fP))
    assert rel_err(sm.cpu(), rsm) < 1e-9 and rel_err(sP.cpu(), rsP) < 1e-9


@pytest.mark.parametrize("name", ["matern52", "rbf6"])
def test_capsule_level_binding_matches_oracle(name):
    """The functions make_tf_ops() wraps in tf.custom_gradient, driven with raw capsules."""
    pkg()
    from pssgp_b200 import dlpack_binding as B
    t, y, cov, ssm = make_problem(name, 900, seed=5)
    P0, Fs, Qs, H, R = [x.clone().requires_grad_(True) for x in ssm]
    rfm, rfP, rll = O.pkf((P0, Fs, Qs, H, R), y[:, None], True)
    g = 0.7
    gP0, gFs, gQs, gH, gR = torch.autograd.grad(g * rll, (P0, Fs, Qs, H, R))
    with torch.no_grad():
        rsm, rsP = O.pks(ssm, rfm.detach(), rfP.detach())
    dev = [x.detach().to(DEV).contiguous() for x in ssm]
    yd = torch.as_tensor(y[:, None]).to(DEV)
    caps = lambda: [to_dlpack(x) for x in dev] + [to_dlpack(yd)]
    (c_fm, c_fP, c_ll), ctx = B.pkf_ll(*caps())
    fm, fP, ll = from_dlpack(c_fm), from_dlpack(c_fP), from_dlpack(c_ll)
    assert rel_err(fm.cpu(), rfm.detach()) < 1e-9 and rel_err(fP.cpu(), rfP.detach()) < 1e-9
    assert abs(float(ll) - float(rll)) < 1e-9 * abs(float(rll))
    grads = [from_dlpack(c) for c in B.pkf_ll_grad(ctx, to_dlpack(torch.tensor([g], dtype=torch.float64, device=DEV)))]
    sym = lambda X: 0.5 * (X + X.transpose(-1, -2))
    assert rel_err(grads[1].cpu(), gFs) < 1e-9 and rel_err(grads[2].cpu(), sym(gQs)) < 1e-9
    assert rel_err(grads[0].cpu(), sym(gP0)) < 1e-9
    assert tuple(grads[3].shape) == tuple(ssm.H.shape) and rel_err(grads[3].cpu(), gH) < 1e-8
    assert abs(float(grads[4]) - float(gR)) < 1e-9 * abs(float(gR))
    sm, sP = (from_dlpack(c) for c in B.pkfs(*caps()))
    assert rel_err(sm.cpu(), rsm) < 1e-9 and rel_err(sP.cpu(), rsP) < 1e-9
    # the sequential filter / smoother through the same boundary (StateSpaceGP(parallel=False) of the reference)
    (k_fm, k_fP, k_ll), kctx = B.kf_ll(*caps())
    assert rel_err(from_dlpack(k_fm).cpu(), rfm.detach()) < 1e-9 and abs(float(from_dlpack(k_ll)) - float(rll)) < 1e-9 * abs(float(rll))
    kgrads = [from_dlpack(c) for c in B.pkf_ll_grad(kctx, to_dlpack(torch.tensor([g], dtype=torch.float64, device=DEV)))]
    assert rel_err(kgrads[1].cpu(), gFs) < 1e-8 and rel_err(kgrads[2].cpu(), sym(gQs)) < 1e-8
    ksm, ksP = (from_dlpack(c) for c in B.kfs(*caps()))
    assert rel_err(ksm.cpu(), rsm) < 1e-8 and rel_err(ksP.cpu(), rsP) < 1e-8
    # discretisation through the same boundary
    with torch.no_grad():
        sde = cov.get_sde()
    dts = torch.as_tensor(np.diff(np.concatenate([[0.0], t]))).to(DEV)
    cFs, cQs = B.get_ssm(to_dlpack(sde.F.to(DEV).contiguous()), to_dlpack(sde.P0.to(DEV).contiguous()), to_dlpack(dts))
    assert rel_err(from_dlpack(cFs).cpu(), ssm.Fs) < 1e-11
