"""Thin device-tensor wrappers over the C ABI (one function per entry point of include/pssgp_b200.h).

Inputs are contiguous CUDA tensors (float64 or float32); outputs are freshly allocated CUDA tensors.
Everything is enqueued on the current torch stream of the tensors' device.
"""
import torch

from . import _arrays as A
from . import _lib


def _h(t):
    return _lib.handle(t.device.index)


def nstate(d):
    return d + d * (d + 1) // 2


def discretise(F, Pinf, dts):
    """(F[d,d], Pinf[d,d], dts[n]) -> (Fs[n,d,d], Qs[n,d,d]);  C ABI: pssgp_discretise."""
    n, d = dts.numel(), F.shape[0]
    Fs = torch.empty((n, d, d), dtype=F.dtype, device=F.device)
    Qs = torch.empty((n, d, d), dtype=F.dtype, device=F.device)
    _lib.check(_lib.lib().pssgp_discretise(_h(F).ptr, A.dtype_code(F), n, d, A.ptr(F), A.ptr(Pinf), A.ptr(dts),
                                          A.ptr(Fs), A.ptr(Qs), A.stream_ptr(F.device)))
    return Fs, Qs


def discretise_backward(F, Pinf, dts, Fs, dFs, dQs):
    """Adjoint of discretise: -> (dF[d,d], dPinf[d,d]);  C ABI: pssgp_discretise_backward."""
    n, d = dts.numel(), F.shape[0]
    dF = torch.empty((d, d), dtype=F.dtype, device=F.device)
    dPinf = torch.empty((d, d), dtype=F.dtype, device=F.device)
    _lib.check(_lib.lib().pssgp_discretise_backward(_h(F).ptr, A.dtype_code(F), n, d, A.ptr(F), A.ptr(Pinf),
                                                   A.ptr(dts), A.ptr(Fs), A.ptr(dFs), A.ptr(dQs), A.ptr(dF),
                                                   A.ptr(dPinf), A.stream_ptr(F.device)))
    return dF, dPinf


def pkf(P0, Fs, Qs, H, R, y, m0=None, first_special=True, want_ll=True, want_final=False):
    """C ABI: pssgp_pkf.  H[d], R[1], y[n] flat.  -> fms, fPs, ll(1-elem tensor or None), final_state or None."""
    n, d = Fs.shape[0], Fs.shape[1]
    fms = torch.empty((n, d), dtype=Fs.dtype, device=Fs.device)
    fPs = torch.empty((n, d, d), dtype=Fs.dtype, device=Fs.device)
    ll = torch.empty((1,), dtype=Fs.dtype, device=Fs.device) if want_ll else None
    fin = torch.empty((nstate(d),), dtype=Fs.dtype, device=Fs.device) if want_final else None
    _lib.check(_lib.lib().pssgp_pkf(_h(Fs).ptr, A.dtype_code(Fs), n, d, A.ptr(P0), A.ptr(Fs), A.ptr(Qs), A.ptr(H),
                                   A.ptr(R), A.ptr(y), A.ptr(m0), 1 if first_special else 0, A.ptr(fms), A.ptr(fPs),
                                   A.ptr(ll), A.ptr(fin), A.stream_ptr(Fs.device)))
    return fms, fPs, ll, fin


def pks(Fs, Qs, fms, fPs, last_special=True, Fnext=None, Qnext=None, init=None, want_first=False):
    """C ABI: pssgp_pks. -> sms, sPs, first_state or None."""
    n, d = Fs.shape[0], Fs.shape[1]
    sms = torch.empty((n, d), dtype=Fs.dtype, device=Fs.device)
    sPs = torch.empty((n, d, d), dtype=Fs.dtype, device=Fs.device)
    first = torch.empty((nstate(d),), dtype=Fs.dtype, device=Fs.device) if want_first else None
    _lib.check(_lib.lib().pssgp_pks(_h(Fs).ptr, A.dtype_code(Fs), n, d, A.ptr(Fs), A.ptr(Qs), A.ptr(fms), A.ptr(fPs),
                                   1 if last_special else 0, A.ptr(Fnext), A.ptr(Qnext), A.ptr(init), A.ptr(sms),
                                   A.ptr(sPs), A.ptr(first), A.stream_ptr(Fs.device)))
    return sms, sPs, first


def pkf_backward(P0, Fs, Qs, H, R, y, fms, fPs, g_ll):
    """C ABI: pssgp_pkf_backward. g_ll: 1-elem device tensor. -> dP0, dFs, dQs, dH, dR."""
    n, d = Fs.shape[0], Fs.shape[1]
    kw = dict(dtype=Fs.dtype, device=Fs.device)
    dP0 = torch.empty((d, d), **kw)
    dFs = torch.empty((n, d, d), **kw)
    dQs = torch.empty((n, d, d), **kw)
    dH = torch.empty((d,), **kw)
    dR = torch.empty((1,), **kw)
    _lib.check(_lib.lib().pssgp_pkf_backward(_h(Fs).ptr, A.dtype_code(Fs), n, d, A.ptr(P0), A.ptr(Fs), A.ptr(Qs),
                                            A.ptr(H), A.ptr(R), A.ptr(y), A.ptr(fms), A.ptr(fPs), A.ptr(g_ll),
                                            A.ptr(dP0), A.ptr(dFs), A.ptr(dQs), A.ptr(dH), A.ptr(dR),
                                            A.stream_ptr(Fs.device)))
    return dP0, dFs, dQs, dH, dR
