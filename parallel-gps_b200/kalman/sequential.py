"""Drop-in for the reference's pssgp/kalman/sequential.py: same names, argument meaning and returns.

    kf(lgssm, observations, return_loglikelihood=False, return_predicted=False)   sequential.py:11-47
    ks(lgssm, ms, Ps, mps, Pps)                                                   sequential.py:50-68
    kfs(model, observations)                                                      sequential.py:71-73

The recursion is serial in time (C ABI ``pssgp_kf`` / ``pssgp_ks``): for one long series it is the comparator the
reference keeps beside ``pkf`` / ``pks``; as an extension, ``observations`` of shape [B,T,1] (or [B,T]) runs B
independent series at once — one thread (d <= 4) or one warp per series — with a shared LGSSM (Fs [T,d,d]) or one
per series (Fs [B,T,d,d], P0 [B,d,d], H [B,1,d], R [B,1,1]).  No CPU fallback.
"""
from .. import _arrays as A
from .. import ops
from .parallel import _dl, _wants_numpy

__all__ = ["kf", "ks", "kfs"]


def _prep(lgssm, extra):
    P0, Fs, Qs, H, R = (_dl(v) for v in lgssm)
    device = A.pick_device(P0, Fs, Qs, H, R, *extra)
    dtype = A.torch_dtype(Fs)
    P0d, Fsd, Qsd = (A.to_device(v, dtype, device, nm) for v, nm in ((P0, "P0"), (Fs, "Fs"), (Qs, "Qs")))
    if Fsd.dim() not in (3, 4) or Fsd.shape[-1] != Fsd.shape[-2] or Qsd.shape != Fsd.shape:
        raise ValueError(f"Fs and Qs must be [T,d,d] (or [B,T,d,d]), got {tuple(Fsd.shape)} / {tuple(Qsd.shape)}")
    d = Fsd.shape[-1]
    lead = tuple(Fsd.shape[:-3])
    Hd = A.to_device(H, dtype, device, "H").reshape(lead + (d,))
    Rd = A.to_device(R, dtype, device, "R").reshape(lead + (1,) if lead else (1,))
    if tuple(P0d.shape) != lead + (d, d):
        raise ValueError(f"P0 must be {lead + (d, d)}, got {tuple(P0d.shape)}")
    return device, dtype, P0d, Fsd, Qsd, Hd, Rd


def _obs(observations, dtype, device, Fs):
    y = A.to_device(observations, dtype, device, "y")
    n = Fs.shape[-3]
    if y.dim() >= 2 and y.shape[-1] == 1 and y.shape[-2] == n:
        y = y.reshape(y.shape[:-1])
    if y.shape[-1] != n or y.dim() > 2 or (Fs.dim() == 4 and (y.dim() != 2 or y.shape[0] != Fs.shape[0])):
        raise ValueError(f"observations must be [{n},1] (or [B,{n},1]), got {tuple(observations.shape)}")
    return y.contiguous()


def kf(lgssm, observations, return_loglikelihood=False, return_predicted=False):
    """Sequential Kalman filter (sequential.py:11-47).
    Returns (fms, fPs) + (ll,) if return_loglikelihood + (mps, Pps) if return_predicted — the reference's order."""
    lgssm, observations = tuple(_dl(v) for v in lgssm), _dl(observations)
    device, dtype, P0, Fs, Qs, H, R = _prep(lgssm, (observations,))
    y = _obs(observations, dtype, device, Fs)
    fms, fPs, ll, mps, Pps = ops.kf(P0, Fs, Qs, H, R, y, want_ll=return_loglikelihood,
                                    want_predicted=return_predicted)
    out = (fms, fPs)
    if return_loglikelihood:
        out = out + ((ll[0] if y.dim() == 1 else ll),)
    if return_predicted:
        out = out + (mps, Pps)
    if _wants_numpy(*lgssm, observations):
        out = tuple(A.to_host(v, "kf") for v in out)
        if return_loglikelihood and y.dim() == 1:
            out = out[:2] + (out[2][()],) + out[3:]
    return out


def ks(lgssm, ms, Ps, mps, Pps):
    """Sequential RTS smoother (sequential.py:50-68). Returns (sms, sPs)."""
    lgssm = tuple(_dl(v) for v in lgssm)
    ms, Ps, mps, Pps = (_dl(v) for v in (ms, Ps, mps, Pps))
    device, dtype, P0, Fs, Qs, H, R = _prep(lgssm, (ms, Ps, mps, Pps))
    msd, Psd, mpd, Ppd = (A.to_device(v, dtype, device, nm).contiguous()
                          for v, nm in ((ms, "ms"), (Ps, "Ps"), (mps, "mps"), (Pps, "Pps")))
    n, d = Fs.shape[-3], Fs.shape[-1]
    if msd.shape[-2:] != (n, d) or Psd.shape[-3:] != (n, d, d) or mpd.shape != msd.shape or Ppd.shape != Psd.shape:
        raise ValueError("ms/mps must be [T,d] and Ps/Pps [T,d,d] (optionally with a leading batch axis)")
    if Fs.dim() == 4 and (msd.dim() != 3 or msd.shape[0] != Fs.shape[0]):
        raise ValueError("a batched LGSSM needs ms of shape [B,T,d] with the same B")
    sms, sPs = ops.ks(Fs, msd, Psd, mpd, Ppd)
    if _wants_numpy(*lgssm, ms, Ps, mps, Pps):
        return A.to_host(sms, "sms"), A.to_host(sPs, "sPs")
    return sms, sPs


def kfs(model, observations):
    """Filter then smoother (sequential.py:71-73); the filtered and predicted moments never leave the device."""
    model, observations = tuple(_dl(v) for v in model), _dl(observations)
    device, dtype, P0, Fs, Qs, H, R = _prep(model, (observations,))
    y = _obs(observations, dtype, device, Fs)
    fms, fPs, _, mps, Pps = ops.kf(P0, Fs, Qs, H, R, y, want_ll=False, want_predicted=True)
    sms, sPs = ops.ks(Fs, fms, fPs, mps, Pps)
    if _wants_numpy(*model, observations):
        return A.to_host(sms, "sms"), A.to_host(sPs, "sPs")
    return sms, sPs
