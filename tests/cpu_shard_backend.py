"""CPU stand-in for pssgp_b200.ops used ONLY by the gloo test of the time-sharding protocol: the
per-shard operations are evaluated with the oracle's elements and operators (float64, sequential)."""
import torch

from util import O


def _pack_f(e):
    return torch.cat([x.reshape(-1) for x in e])


def _unpack_f(v, d):
    A = v[:d * d].reshape(1, d, d)
    b = v[d * d:d * d + d].reshape(1, d)
    C = v[d * d + d:2 * d * d + d].reshape(1, d, d)
    J = v[2 * d * d + d:3 * d * d + d].reshape(1, d, d)
    eta = v[3 * d * d + d:].reshape(1, d)
    return A, b, C, J, eta


def nagg_filter(d):
    return 3 * d * d + 2 * d


def _filter_elements(P0, Fs, Qs, H, R, y, first_special):
    Hm, Rm = H.reshape(1, -1), R.reshape(1, 1)
    yy = y.reshape(-1, 1)
    if first_special:
        return O.make_associative_filtering_elements(torch.zeros(P0.shape[0], dtype=P0.dtype), P0, Fs, Qs, Hm, Rm, yy)
    nan = torch.isnan(yy).reshape(-1)
    ok = O._generic_filtering_element(Fs, Qs, Hm, Rm, torch.where(torch.isnan(yy), torch.zeros_like(yy), yy))
    na = O._generic_filtering_element_nan(Fs, Qs)
    return tuple(torch.where(nan.reshape((-1,) + (1,) * (a.dim() - 1)), a, b) for a, b in zip(na, ok))


def _reduce(op, elems):
    acc = tuple(e[0:1] for e in elems)
    for i in range(1, elems[0].shape[0]):
        acc = op(acc, tuple(e[i:i + 1] for e in elems))
    return acc


def pkf_summary(P0, Fs, Qs, H, R, y, first_special):
    return _pack_f(_reduce(O.filtering_operator, _filter_elements(P0, Fs, Qs, H, R, y, first_special)))


def _state_elem(m, P):
    d = P.shape[0]
    z = torch.zeros(1, d, d, dtype=P.dtype)
    return z, m.reshape(1, d), P.reshape(1, d, d), z.clone(), torch.zeros(1, d, dtype=P.dtype)


def filter_fold(P0, m0, summaries, count):
    d = P0.shape[0]
    acc = _state_elem(torch.zeros(d, dtype=P0.dtype) if m0 is None else m0, P0)
    for i in range(count):
        acc = O.filtering_operator(acc, _unpack_f(summaries[i], d))
    return torch.cat([acc[1].reshape(-1), acc[2].reshape(-1)])


def pkf(P_in, Fs, Qs, H, R, y, m0=None, first_special=True, want_ll=True, want_final=False):
    d = Fs.shape[1]
    elems = _filter_elements(P_in, Fs, Qs, H, R, y, first_special)
    if not first_special:
        pre = _state_elem(m0, P_in)
        elems = tuple(torch.cat([p, e]) for p, e in zip(pre, elems))
    res = O.scan_associative(O.filtering_operator, elems)
    fms, fPs = res[1], res[2]
    if not first_special:
        fms, fPs = fms[1:], fPs[1:]
    # log-likelihood of the shard (parallel.py:135-151) with the entering state as "previous filtered"
    m_prev = torch.zeros(d, dtype=Fs.dtype) if m0 is None else m0
    pm = torch.cat([m_prev.reshape(1, d), fms[:-1]])
    pP = torch.cat([P_in.reshape(1, d, d), fPs[:-1]])
    mp = O.mv(Fs, pm)
    Pp = Fs @ pP @ O.tr(Fs) + Qs
    h = H.reshape(-1)
    S = torch.einsum("i,kij,j->k", h, Pp, h) + R.reshape(())
    r = y.reshape(-1) - mp @ h
    lp = -0.5 * (r * r / S + torch.log(2 * torch.pi * S))
    ll = torch.where(torch.isnan(lp), torch.zeros_like(lp), lp).sum().reshape(1)
    return fms, fPs, ll, None


def nagg_smoother(d):
    return 2 * d * d + d


def _smoother_elements(Fs, Qs, fms, fPs, last_special, Fnext, Qnext):
    if last_special:
        return O.make_associative_smoothing_elements(Fs, Qs, fms, fPs)
    Fx = torch.cat([Fs[1:], Fnext.reshape(1, *Fnext.shape)])
    Qx = torch.cat([Qs[1:], Qnext.reshape(1, *Qnext.shape)])
    return O.generic_smoothing_element(Fx, Qx, fms, fPs)


def pks_summary(Fs, Qs, fms, fPs, last_special, Fnext=None, Qnext=None):
    elems = tuple(torch.flip(e, dims=[0]) for e in _smoother_elements(Fs, Qs, fms, fPs, last_special, Fnext, Qnext))
    return torch.cat([x.reshape(-1) for x in _reduce(O.smoothing_operator, elems)])


def _unpack_s(v, d):
    return v[:d * d].reshape(1, d, d), v[d * d:d * d + d].reshape(1, d), v[d * d + d:].reshape(1, d, d)


def smoother_fold(summaries, count, d):
    acc = _unpack_s(summaries[count - 1], d)
    for i in range(count - 2, -1, -1):
        acc = O.smoothing_operator(acc, _unpack_s(summaries[i], d))
    return torch.cat([acc[1].reshape(-1), acc[2].reshape(-1)])  # mean | FULL covariance (CPU backend convention)


def pks(Fs, Qs, fms, fPs, last_special=True, Fnext=None, Qnext=None, init=None, want_first=False):
    d = Fs.shape[1]
    elems = tuple(torch.flip(e, dims=[0]) for e in _smoother_elements(Fs, Qs, fms, fPs, last_special, Fnext, Qnext))
    if not last_special:
        pre = (torch.zeros(1, d, d, dtype=Fs.dtype), init[:d].reshape(1, d), init[d:].reshape(1, d, d))
        elems = tuple(torch.cat([p, e]) for p, e in zip(pre, elems))
    res = O.scan_associative(O.smoothing_operator, elems)
    sms, sPs = torch.flip(res[1], dims=[0]), torch.flip(res[2], dims=[0])
    if not last_special:
        sms, sPs = sms[:-1], sPs[:-1]
    return sms, sPs, None
