"""Drop-in for the reference's pssgp/kalman/parallel.py: same names, argument meaning and returns.

    pkf(lgssm, observations, return_loglikelihood=False, max_parallel=10000)   parallel.py:121-152
    pks(lgssm, ms, Ps, max_parallel=10000)                                     parallel.py:187-196
    pkfs(model, observations, max_parallel=10000)                              parallel.py:199-201

``lgssm`` is any 5-tuple (P0[d,d], Fs[T,d,d], Qs[T,d,d], H[1,d], R[1,1]) (unpacked positionally like
parallel.py:123), ``observations`` is [T,1] with NaN marking a missing observation.  Inputs may be
numpy arrays (results come back as numpy; H2D/D2H through pinned staging) or torch CUDA tensors
(results stay on the device, zero-copy).  ``max_parallel`` is accepted for compatibility; the
chunked single-sweep scan does not need a recursion-depth bound.

All arithmetic runs in the CUDA library (C ABI ``pssgp_pkf`` / ``pssgp_pks``); no CPU fallback.
"""
import torch

from .. import _arrays as A
from .. import _lib

__all__ = ["pkf", "pks", "pkfs"]


def _dl(x):
    """DLPack producers (TF / CuPy / JAX tensors, raw capsules) become zero-copy torch views; the rest is untouched."""
    import numpy as np
    if not isinstance(x, (torch.Tensor, np.ndarray)) and (hasattr(x, "__dlpack__") or type(x).__name__ == "PyCapsule"):
        return A.from_dlpack(x)
    return x


def _prep_lgssm(lgssm, extra=()):
    P0, Fs, Qs, H, R = (_dl(v) for v in lgssm)
    device = A.pick_device(P0, Fs, Qs, H, R, *extra)
    dtype = A.torch_dtype(Fs)
    P0d = A.to_device(P0, dtype, device, "P0")
    Fsd = A.to_device(Fs, dtype, device, "Fs")
    Qsd = A.to_device(Qs, dtype, device, "Qs")
    Hd = A.to_device(H, dtype, device, "H").reshape(-1)
    Rd = A.to_device(R, dtype, device, "R").reshape(-1)
    if Fsd.dim() != 3 or Fsd.shape[1] != Fsd.shape[2]:
        raise ValueError(f"Fs must be [T,d,d], got {tuple(Fsd.shape)}")
    n, d = Fsd.shape[0], Fsd.shape[1]
    if tuple(Qsd.shape) != (n, d, d):
        raise ValueError(f"Qs must be {(n, d, d)}, got {tuple(Qsd.shape)}")
    if tuple(P0d.shape) != (d, d):
        raise ValueError(f"P0 must be {(d, d)}, got {tuple(P0d.shape)}")
    if Hd.numel() != d:
        raise ValueError(f"H must be [1,{d}] (single-output models only, cf. pssgp/kernels/base.py:23), got {tuple(H.shape)}")
    if Rd.numel() != 1:
        raise ValueError("R must be [1,1] (single-output models only, cf. pssgp/kernels/base.py:24)")
    return device, dtype, n, d, P0d, Fsd, Qsd, Hd, Rd


def _wants_numpy(*xs):
    return not any(A.is_device_tensor(x) for x in xs)


def pkf(lgssm, observations, return_loglikelihood=False, max_parallel=10000):
    """Parallel Kalman filter (parallel.py:121-152). Returns (fms[T,d], fPs[T,d,d][, ll])."""
    lgssm, observations = tuple(_dl(v) for v in lgssm), _dl(observations)
    device, dtype, n, d, P0, Fs, Qs, H, R = _prep_lgssm(lgssm, (observations,))
    y = A.to_device(observations, dtype, device, "y").reshape(-1)
    if y.numel() != n:
        raise ValueError(f"observations must be [{n},1], got {tuple(observations.shape)}")
    fms = torch.empty((n, d), dtype=dtype, device=device)
    fPs = torch.empty((n, d, d), dtype=dtype, device=device)
    ll = torch.empty((1,), dtype=dtype, device=device) if return_loglikelihood else None
    h = _lib.handle(device.index)
    _lib.check(_lib.lib().pssgp_pkf(h.ptr, A.dtype_code(Fs), n, d, A.ptr(P0), A.ptr(Fs), A.ptr(Qs), A.ptr(H),
                                   A.ptr(R), A.ptr(y), None, 1, A.ptr(fms), A.ptr(fPs), A.ptr(ll), None,
                                   A.stream_ptr(device)))
    if _wants_numpy(*lgssm, observations):
        out = (A.to_host(fms, "fms"), A.to_host(fPs, "fPs"))
        if return_loglikelihood:
            out = out + (A.to_host(ll, "ll")[0],)
        return out
    if return_loglikelihood:
        return fms, fPs, ll[0]
    return fms, fPs


def pks(lgssm, ms, Ps, max_parallel=10000):
    """Parallel RTS smoother (parallel.py:187-196). Returns (sms[T,d], sPs[T,d,d])."""
    lgssm, ms, Ps = tuple(_dl(v) for v in lgssm), _dl(ms), _dl(Ps)
    device, dtype, n, d, P0, Fs, Qs, H, R = _prep_lgssm(lgssm, (ms, Ps))
    msd = A.to_device(ms, dtype, device, "ms")
    Psd = A.to_device(Ps, dtype, device, "Ps")
    if tuple(msd.shape) != (n, d) or tuple(Psd.shape) != (n, d, d):
        raise ValueError("ms/Ps must be [T,d] / [T,d,d]")
    sms = torch.empty((n, d), dtype=dtype, device=device)
    sPs = torch.empty((n, d, d), dtype=dtype, device=device)
    h = _lib.handle(device.index)
    _lib.check(_lib.lib().pssgp_pks(h.ptr, A.dtype_code(Fs), n, d, A.ptr(Fs), A.ptr(Qs), A.ptr(msd), A.ptr(Psd),
                                   1, None, None, None, A.ptr(sms), A.ptr(sPs), None, A.stream_ptr(device)))
    if _wants_numpy(*lgssm, ms, Ps):
        return A.to_host(sms, "sms"), A.to_host(sPs, "sPs")
    return sms, sPs


def pkfs(model, observations, max_parallel=10000):
    """Filter then smoother (parallel.py:199-201); the filtered moments never leave the device."""
    model, observations = tuple(_dl(v) for v in model), _dl(observations)
    device, dtype, n, d, P0, Fs, Qs, H, R = _prep_lgssm(model, (observations,))
    y = A.to_device(observations, dtype, device, "y").reshape(-1)
    if y.numel() != n:
        raise ValueError(f"observations must be [{n},1], got {tuple(observations.shape)}")
    from .. import ops
    _, _, _, sms, sPs = ops.pkfs(P0, Fs, Qs, H, R, y)  # C ABI pssgp_pkfs: one fused call
    if _wants_numpy(*model, observations):
        return A.to_host(sms, "sms"), A.to_host(sPs, "sPs")
    return sms, sPs
