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


def _al(t):
    """The streaming kernels move 16-byte pieces: a sliced view may start off a 16-byte boundary."""
    return t if t is None or t.data_ptr() % 16 == 0 else t.clone()


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
    """C ABI: pssgp_pkf.  H[d], R[1], y[n] flat.  -> fms, fPs, ll(1-elem tensor or None), final_state or None
    (final_state = m[d] | P[d,d])."""
    Fs, Qs, y = _al(Fs), _al(Qs), _al(y)
    n, d = Fs.shape[0], Fs.shape[1]
    fms = torch.empty((n, d), dtype=Fs.dtype, device=Fs.device)
    fPs = torch.empty((n, d, d), dtype=Fs.dtype, device=Fs.device)
    ll = torch.empty((1,), dtype=Fs.dtype, device=Fs.device) if want_ll else None
    fin = torch.empty((d + d * d,), dtype=Fs.dtype, device=Fs.device) if want_final else None
    _lib.check(_lib.lib().pssgp_pkf(_h(Fs).ptr, A.dtype_code(Fs), n, d, A.ptr(P0), A.ptr(Fs), A.ptr(Qs), A.ptr(H),
                                   A.ptr(R), A.ptr(y), A.ptr(m0), 1 if first_special else 0, A.ptr(fms), A.ptr(fPs),
                                   A.ptr(ll), A.ptr(fin), A.stream_ptr(Fs.device)))
    return fms, fPs, ll, fin


def pks(Fs, Qs, fms, fPs, last_special=True, Fnext=None, Qnext=None, init=None, want_first=False):
    """C ABI: pssgp_pks. -> sms, sPs, first_state or None."""
    Fs, Qs, fms, fPs = _al(Fs), _al(Qs), _al(fms), _al(fPs)
    n, d = Fs.shape[0], Fs.shape[1]
    sms = torch.empty((n, d), dtype=Fs.dtype, device=Fs.device)
    sPs = torch.empty((n, d, d), dtype=Fs.dtype, device=Fs.device)
    first = torch.empty((nstate(d),), dtype=Fs.dtype, device=Fs.device) if want_first else None
    _lib.check(_lib.lib().pssgp_pks(_h(Fs).ptr, A.dtype_code(Fs), n, d, A.ptr(Fs), A.ptr(Qs), A.ptr(fms), A.ptr(fPs),
                                   1 if last_special else 0, A.ptr(Fnext), A.ptr(Qnext), A.ptr(init), A.ptr(sms),
                                   A.ptr(sPs), A.ptr(first), A.stream_ptr(Fs.device)))
    return sms, sPs, first


def pkf_backward(P0, Fs, Qs, H, R, y, fms, fPs, g_ll, m0=None, first_special=True, adj_init=None, want_first=False):
    """C ABI: pssgp_pkf_backward. g_ll: 1-elem device tensor. -> dP0, dFs, dQs, dH, dR[, adj_first]."""
    Fs, Qs, y, fms, fPs = _al(Fs), _al(Qs), _al(y), _al(fms), _al(fPs)
    n, d = Fs.shape[0], Fs.shape[1]
    kw = dict(dtype=Fs.dtype, device=Fs.device)
    dP0 = torch.zeros((d, d), **kw)
    dFs = torch.empty((n, d, d), **kw)
    dQs = torch.empty((n, d, d), **kw)
    dH = torch.empty((d,), **kw)
    dR = torch.empty((1,), **kw)
    first = torch.empty((nstate(d),), **kw) if want_first else None
    _lib.check(_lib.lib().pssgp_pkf_backward(_h(Fs).ptr, A.dtype_code(Fs), n, d, A.ptr(P0), A.ptr(m0), A.ptr(Fs),
                                            A.ptr(Qs), A.ptr(H), A.ptr(R), A.ptr(y), A.ptr(fms), A.ptr(fPs),
                                            A.ptr(g_ll), 1 if first_special else 0, A.ptr(adj_init), A.ptr(dP0),
                                            A.ptr(dFs), A.ptr(dQs), A.ptr(dH), A.ptr(dR), A.ptr(first),
                                            A.stream_ptr(Fs.device)))
    if want_first:
        return dP0, dFs, dQs, dH, dR, first
    return dP0, dFs, dQs, dH, dR


def pkfs(P0, Fs, Qs, H, R, y, want_ll=False, project=False):
    """C ABI: pssgp_pkfs (filter + smoother of one whole series, fused for d <= 4).
    -> fms, fPs, ll (or None), then (sms, sPs), or with project=True proj[n,2] = (H m_k, H P_k H^T) of the smoothed
    states (d <= 4 only; raises PssgpError otherwise)."""
    Fs, Qs, y = _al(Fs), _al(Qs), _al(y)
    n, d = Fs.shape[0], Fs.shape[1]
    kw = dict(dtype=Fs.dtype, device=Fs.device)
    fms, fPs = torch.empty((n, d), **kw), torch.empty((n, d, d), **kw)
    ll = torch.empty((1,), **kw) if want_ll else None
    sms = sPs = proj = None
    if project:
        proj = torch.empty((n, 2), **kw)
    else:
        sms, sPs = torch.empty((n, d), **kw), torch.empty((n, d, d), **kw)
    _lib.check(_lib.lib().pssgp_pkfs(_h(Fs).ptr, A.dtype_code(Fs), n, d, A.ptr(P0), A.ptr(Fs), A.ptr(Qs), A.ptr(H),
                                    A.ptr(R), A.ptr(y), A.ptr(fms), A.ptr(fPs), A.ptr(ll), A.ptr(sms), A.ptr(sPs),
                                    A.ptr(proj), A.stream_ptr(Fs.device)))
    if project:
        return fms, fPs, ll, proj
    return fms, fPs, ll, sms, sPs


def merge_queries(ts, ys, q, t0=0.0):
    """C ABI: pssgp_merge_queries.  ts, ys [n], q [K] sorted device vectors.
    -> t_all, y_all, dts [n+K], q_idx [K] (int64 rows of the queries in the merged arrays)."""
    n, K = ts.numel(), q.numel()
    kw = dict(dtype=ts.dtype, device=ts.device)
    t_all, y_all, dts = torch.empty(n + K, **kw), torch.empty(n + K, **kw), torch.empty(n + K, **kw)
    q_idx = torch.empty(K, dtype=torch.int64, device=ts.device)
    _lib.check(_lib.lib().pssgp_merge_queries(_h(ts).ptr, A.dtype_code(ts), n, K, A.ptr(ts), A.ptr(ys), A.ptr(q), float(t0),
                                             A.ptr(t_all), A.ptr(y_all), A.ptr(dts), A.ptr(q_idx),
                                             A.stream_ptr(ts.device)))
    return t_all, y_all, dts, q_idx


def kf(P0, Fs, Qs, H, R, y, want_ll=True, want_predicted=False):
    """C ABI: pssgp_kf (sequential Kalman filter).  y [n] or [batch,n]; the LGSSM is shared by the series when Fs is
    [n,d,d] and per series when it is [batch,n,d,d].  -> fms, fPs, ll[batch] or None, mps, Pps (or None, None)."""
    batched_y = y.dim() == 2
    batch = y.shape[0] if batched_y else 1
    lg = Fs.dim() == 4
    n, d = Fs.shape[-3], Fs.shape[-1]
    kw = dict(dtype=Fs.dtype, device=Fs.device)
    lead = (batch,) if batched_y else ()
    fms, fPs = torch.empty(lead + (n, d), **kw), torch.empty(lead + (n, d, d), **kw)
    mps = torch.empty(lead + (n, d), **kw) if want_predicted else None
    Pps = torch.empty(lead + (n, d, d), **kw) if want_predicted else None
    ll = torch.empty((batch,), **kw) if want_ll else None
    _lib.check(_lib.lib().pssgp_kf(_h(Fs).ptr, A.dtype_code(Fs), batch, n, d, 1 if lg else 0, A.ptr(P0), A.ptr(Fs),
                                  A.ptr(Qs), A.ptr(H), A.ptr(R), A.ptr(y), A.ptr(fms), A.ptr(fPs), A.ptr(mps),
                                  A.ptr(Pps), A.ptr(ll), A.stream_ptr(Fs.device)))
    return fms, fPs, ll, mps, Pps


def ks(Fs, fms, fPs, mps, Pps):
    """C ABI: pssgp_ks (sequential RTS smoother) on the outputs of kf(want_predicted=True). -> sms, sPs."""
    batched = fms.dim() == 3
    batch = fms.shape[0] if batched else 1
    n, d = fms.shape[-2], fms.shape[-1]
    sms, sPs = torch.empty_like(fms), torch.empty_like(fPs)
    _lib.check(_lib.lib().pssgp_ks(_h(Fs).ptr, A.dtype_code(Fs), batch, n, d, 1 if Fs.dim() == 4 else 0, A.ptr(Fs),
                                  A.ptr(fms), A.ptr(fPs), A.ptr(mps), A.ptr(Pps), A.ptr(sms), A.ptr(sPs),
                                  A.stream_ptr(Fs.device)))
    return sms, sPs


def has_projection(d, dtype):
    """pssgp_pkfs can emit (H m, H P H^T) of the smoothed states directly: the fused d <= 4 kernels with a common
    partition (every d <= 3, and d = 4 in FP32) and the FP64 fragment-resident kernels (5 <= d <= 24)."""
    return d <= 3 or (d == 4 and dtype == torch.float32) or (5 <= d <= 24 and dtype == torch.float64)


def pkfs_grad(P0, Fs, Qs, H, R, y, g_ll, want_smoother=True):
    """C ABI: pssgp_pkfs_grad (filter + log-likelihood + smoother + gradient of one whole series, fused).
    -> (fms, fPs, ll), (sms, sPs) or None, (dP0, dFs, dQs, dH, dR)."""
    Fs, Qs, y = _al(Fs), _al(Qs), _al(y)
    n, d = Fs.shape[0], Fs.shape[1]
    kw = dict(dtype=Fs.dtype, device=Fs.device)
    fms = torch.empty((n, d), **kw)
    fPs, dFs, dQs = (torch.empty((n, d, d), **kw) for _ in range(3))
    sms = torch.empty((n, d), **kw) if want_smoother else None
    sPs = torch.empty((n, d, d), **kw) if want_smoother else None
    ll, dR = torch.empty((1,), **kw), torch.empty((1,), **kw)
    dP0 = torch.zeros((d, d), **kw)
    dH = torch.empty((d,), **kw)
    _lib.check(_lib.lib().pssgp_pkfs_grad(_h(Fs).ptr, A.dtype_code(Fs), n, d, A.ptr(P0), A.ptr(Fs), A.ptr(Qs),
                                         A.ptr(H), A.ptr(R), A.ptr(y), A.ptr(g_ll), A.ptr(fms), A.ptr(fPs), A.ptr(ll),
                                         A.ptr(sms), A.ptr(sPs), A.ptr(dP0), A.ptr(dFs), A.ptr(dQs), A.ptr(dH),
                                         A.ptr(dR), A.stream_ptr(Fs.device)))
    return (fms, fPs, ll), ((sms, sPs) if want_smoother else None), (dP0, dFs, dQs, dH, dR)


# ---- time sharding (one contiguous shard per GPU) --------------------------------------------------------
SMALL_D = 4  # d <= SMALL_D: register-resident kernels with packed symmetric aggregates; above: full matrices


def nagg_filter(d):
    """length of a filter shard summary (A,b,C,J,eta)."""
    return d * d + 2 * d + d * (d + 1) if d <= SMALL_D else 3 * d * d + 2 * d


def nagg_smoother(d):
    """length of a smoother (E,g,L) or adjoint (Abar,a,B) shard summary."""
    return d * d + d + d * (d + 1) // 2 if d <= SMALL_D else 2 * d * d + d


def pkf_summary(P0, Fs, Qs, H, R, y, first_special):
    n, d = Fs.shape[0], Fs.shape[1]
    out = torch.empty((nagg_filter(d),), dtype=Fs.dtype, device=Fs.device)
    _lib.check(_lib.lib().pssgp_pkf_summary(_h(Fs).ptr, A.dtype_code(Fs), n, d, A.ptr(P0), A.ptr(Fs), A.ptr(Qs),
                                           A.ptr(H), A.ptr(R), A.ptr(y), 1 if first_special else 0, A.ptr(out),
                                           A.stream_ptr(Fs.device)))
    return out


def pkf_with_summaries(P0, Fs, Qs, H, R, y, m0=None, first_special=True, last_special=True, Fnext=None, Qnext=None):
    """C ABI: pssgp_pkf_with_summaries (pkf on one shard + shard summaries of the smoother and adjoint scans).
    -> fms, fPs, ll, smoother_summary, adjoint_summary."""
    Fs, Qs, y = _al(Fs), _al(Qs), _al(y)
    n, d = Fs.shape[0], Fs.shape[1]
    kw = dict(dtype=Fs.dtype, device=Fs.device)
    fms = torch.empty((n, d), **kw)
    fPs = torch.empty((n, d, d), **kw)
    ll = torch.empty((1,), **kw)
    s_sm = torch.empty((nagg_smoother(d),), **kw)
    s_ad = torch.empty((nagg_smoother(d),), **kw)
    _lib.check(_lib.lib().pssgp_pkf_with_summaries(_h(Fs).ptr, A.dtype_code(Fs), n, d, A.ptr(P0), A.ptr(Fs), A.ptr(Qs),
                                                  A.ptr(H), A.ptr(R), A.ptr(y), A.ptr(m0), 1 if first_special else 0,
                                                  1 if last_special else 0, A.ptr(Fnext), A.ptr(Qnext), A.ptr(fms),
                                                  A.ptr(fPs), A.ptr(ll), A.ptr(s_sm), A.ptr(s_ad),
                                                  A.stream_ptr(Fs.device)))
    return fms, fPs, ll, s_sm, s_ad


def filter_fold(P0, m0, summaries, count):
    """summaries: [>=count, NAGG] in rank order. -> m[d] | P[d,d]."""
    d = P0.shape[0]
    out = torch.empty((d + d * d,), dtype=P0.dtype, device=P0.device)
    _lib.check(_lib.lib().pssgp_filter_fold(_h(P0).ptr, A.dtype_code(P0), d, int(count), A.ptr(P0), A.ptr(m0),
                                           A.ptr(summaries), A.ptr(out), A.stream_ptr(P0.device)))
    return out


def pks_summary(Fs, Qs, fms, fPs, last_special, Fnext=None, Qnext=None):
    n, d = Fs.shape[0], Fs.shape[1]
    out = torch.empty((nagg_smoother(d),), dtype=Fs.dtype, device=Fs.device)
    _lib.check(_lib.lib().pssgp_pks_summary(_h(Fs).ptr, A.dtype_code(Fs), n, d, A.ptr(Fs), A.ptr(Qs), A.ptr(fms),
                                           A.ptr(fPs), 1 if last_special else 0, A.ptr(Fnext), A.ptr(Qnext),
                                           A.ptr(out), A.stream_ptr(Fs.device)))
    return out


def set_fold(kind, summaries, count, state_out=None):
    """C ABI: pssgp_set_fold.  kind: 0 = filter (summaries of the shards before this one, consumed by
    pkf_with_summaries; state_out [d + d*d] receives the folded state entering the shard), 1 = smoother, 2 = adjoint
    (summaries of the shards after this one, consumed by pks / pkf_backward).  summaries: [count, NAGG] view in rank
    order (rows may be strided, e.g. a column block of the all-gather buffer).  d <= 4 only."""
    if count > 0 and (summaries.dim() != 2 or summaries.stride(1) != 1):
        raise ValueError("summaries must be a 2-d view with unit stride along the last axis")
    _lib.check(_lib.lib().pssgp_set_fold(_h(summaries).ptr, int(kind), A.ptr(summaries) if count > 0 else None, int(count),
                                         int(summaries.stride(0)) if count > 0 else 0, A.ptr(state_out)))


def smoother_fold(summaries, count, d):
    """summaries: [count, NAGG] of the FOLLOWING shards in rank order. -> packed state."""
    out = torch.empty((nstate(d),), dtype=summaries.dtype, device=summaries.device)
    _lib.check(_lib.lib().pssgp_smoother_fold(_h(summaries).ptr, A.dtype_code(summaries), d, int(count),
                                             A.ptr(summaries), A.ptr(out), A.stream_ptr(summaries.device)))
    return out


def pkf_backward_summary(P0, m0, Fs, Qs, H, R, y, fms, fPs, first_special):
    n, d = Fs.shape[0], Fs.shape[1]
    out = torch.empty((nagg_smoother(d),), dtype=Fs.dtype, device=Fs.device)
    _lib.check(_lib.lib().pssgp_pkf_backward_summary(_h(Fs).ptr, A.dtype_code(Fs), n, d, A.ptr(P0), A.ptr(m0),
                                                    A.ptr(Fs), A.ptr(Qs), A.ptr(H), A.ptr(R), A.ptr(y), A.ptr(fms),
                                                    A.ptr(fPs), 1 if first_special else 0, A.ptr(out),
                                                    A.stream_ptr(Fs.device)))
    return out


def adjoint_fold(summaries, count, d):
    out = torch.empty((nstate(d),), dtype=summaries.dtype, device=summaries.device)
    _lib.check(_lib.lib().pssgp_adjoint_fold(_h(summaries).ptr, A.dtype_code(summaries), d, int(count),
                                            A.ptr(summaries), A.ptr(out), A.stream_ptr(summaries.device)))
    return out


# ---- time sharding for 5 <= d <= 32: combined reverse scan (smoother in MBF form + log-likelihood adjoint) ----
MID_D_MAX = 32


def has_combined_reverse(d, dtype=torch.float64):
    """True where pssgp_shard_forward / pssgp_rev_fold / pssgp_shard_reverse are implemented."""
    return SMALL_D < d <= MID_D_MAX and dtype == torch.float64


def nagg_rev(d):
    """length of a combined reverse-scan shard summary  Abar | Ba | Bm | a."""
    return 3 * d * d + d


def nstate_rev(d):
    """length of a combined reverse-scan state  dm | lam | dP | Lam."""
    return 2 * d * d + 2 * d


def shard_forward(P0, Fs, Qs, H, R, y, m0=None, first_special=True):
    """C ABI: pssgp_shard_forward. -> fms, fPs, ll, rev_summary."""
    Fs, Qs, y = _al(Fs), _al(Qs), _al(y)
    n, d = Fs.shape[0], Fs.shape[1]
    kw = dict(dtype=Fs.dtype, device=Fs.device)
    fms, fPs = torch.empty((n, d), **kw), torch.empty((n, d, d), **kw)
    ll = torch.empty((1,), **kw)
    summ = torch.empty((nagg_rev(d),), **kw)
    _lib.check(_lib.lib().pssgp_shard_forward(_h(Fs).ptr, A.dtype_code(Fs), n, d, A.ptr(P0), A.ptr(Fs), A.ptr(Qs), A.ptr(H),
                                             A.ptr(R), A.ptr(y), A.ptr(m0), 1 if first_special else 0, A.ptr(fms),
                                             A.ptr(fPs), A.ptr(ll), A.ptr(summ), A.stream_ptr(Fs.device)))
    return fms, fPs, ll, summ


def rev_fold(summaries, count, d):
    """summaries: [count, nagg_rev(d)] view of the FOLLOWING shards in rank order (rows may be strided).
    -> state entering this shard from above (dm | lam | dP | Lam)."""
    if summaries.dim() != 2 or summaries.stride(1) != 1:
        raise ValueError("summaries must be a 2-d view with unit stride along the last axis")
    out = torch.empty((nstate_rev(d),), dtype=summaries.dtype, device=summaries.device)
    _lib.check(_lib.lib().pssgp_rev_fold(_h(summaries).ptr, A.dtype_code(summaries), d, int(count), A.ptr(summaries),
                                        int(summaries.stride(0)), A.ptr(out), A.stream_ptr(summaries.device)))
    return out


def shard_reverse(P0, Fs, Qs, H, R, y, fms, fPs, g_ll=None, m0=None, first_special=True, rev_init=None,
                  want_smoother=True, want_grad=True):
    """C ABI: pssgp_shard_reverse. -> (sms, sPs) or None, (dP0, dFs, dQs, dH, dR) or None."""
    Fs, Qs, y, fms, fPs = _al(Fs), _al(Qs), _al(y), _al(fms), _al(fPs)
    n, d = Fs.shape[0], Fs.shape[1]
    kw = dict(dtype=Fs.dtype, device=Fs.device)
    sms = sPs = dP0 = dFs = dQs = dH = dR = None
    if want_smoother:
        sms, sPs = torch.empty((n, d), **kw), torch.empty((n, d, d), **kw)
    if want_grad:
        dP0 = torch.zeros((d, d), **kw)
        dFs, dQs = torch.empty((n, d, d), **kw), torch.empty((n, d, d), **kw)
        dH, dR = torch.empty((d,), **kw), torch.empty((1,), **kw)
    _lib.check(_lib.lib().pssgp_shard_reverse(_h(Fs).ptr, A.dtype_code(Fs), n, d, A.ptr(P0), A.ptr(m0), A.ptr(Fs), A.ptr(Qs),
                                             A.ptr(H), A.ptr(R), A.ptr(y), A.ptr(fms), A.ptr(fPs), A.ptr(g_ll),
                                             1 if first_special else 0, A.ptr(rev_init), A.ptr(sms), A.ptr(sPs),
                                             A.ptr(dP0), A.ptr(dFs), A.ptr(dQs), A.ptr(dH), A.ptr(dR),
                                             A.stream_ptr(Fs.device)))
    return ((sms, sPs) if want_smoother else None), ((dP0, dFs, dQs, dH, dR) if want_grad else None)
