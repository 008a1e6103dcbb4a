"""TEST / BENCHMARK INFRASTRUCTURE ONLY (never imported by the product package).

numba-JIT restatement of the reference's SEQUENTIAL Kalman filter and RTS smoother
(pssgp/kalman/sequential.py:11-47 `kf`, :50-68 `ks`, :71-73 `kfs`) for a scalar observation model
(H [1, d], R [1, 1]; the only case the reference supports, pssgp/model.py:72).  One host core: this is the
"numba sequential kf/ks" CPU baseline of BASELINE.md section 4; `tests/test_cpu_oracle.py` pins it to the
torch restatement in pssgp_oracle.py.
"""
import math

import numba
import numpy as np


@numba.njit(cache=True, fastmath=False)
def kf(P0, Fs, Qs, H, R, y):
    """sequential.py:11-47.  Returns fms [T,d], fPs [T,d,d], mps [T,d], Pps [T,d,d], ll."""
    T, d = Fs.shape[0], Fs.shape[1]
    fms = np.zeros((T, d))
    fPs = np.zeros((T, d, d))
    mps = np.zeros((T, d))
    Pps = np.zeros((T, d, d))
    m = np.zeros(d)
    P = P0.copy()
    h = H.reshape(-1)
    r = R.reshape(-1)[0]
    ll = 0.0
    for k in range(T):
        F = Fs[k]
        mp = F @ m
        Pp = F @ P @ F.T + Qs[k]
        Pp = 0.5 * (Pp + Pp.T)                      # sequential.py:21
        yk = y[k]
        if math.isnan(yk):                          # sequential.py:38 (tf.cond on NaN)
            m = mp
            P = Pp
        else:
            u = Pp @ h
            S = h @ u + r
            e = yk - h @ mp
            ll += -0.5 * (math.log(2.0 * math.pi * S) + e * e / S)   # :27-28 MultivariateNormalTriL.log_prob
            K = u / S
            m = mp + K * e
            P = Pp - np.outer(K, K) * S
            P = 0.5 * (P + P.T)                     # sequential.py:39
        fms[k] = m
        fPs[k] = P
        mps[k] = mp
        Pps[k] = Pp
    return fms, fPs, mps, Pps, ll


@numba.njit(cache=True, fastmath=False)
def ks(Fs, fms, fPs, mps, Pps):
    """sequential.py:50-68 (reverse tf.scan)."""
    T, d = Fs.shape[0], Fs.shape[1]
    sms = np.zeros((T, d))
    sPs = np.zeros((T, d, d))
    sm = fms[T - 1].copy()
    sP = fPs[T - 1].copy()
    sms[T - 1] = sm
    sPs[T - 1] = sP
    for k in range(T - 2, -1, -1):
        F = Fs[k + 1]
        Pp = Pps[k + 1]
        # E = P_k F^T Pp^-1 through a Cholesky solve (sequential.py:57-58)
        C = np.linalg.cholesky(Pp)
        X = np.linalg.solve(C, F @ fPs[k])
        Et = np.linalg.solve(C.T, X)
        E = Et.T
        sm = fms[k] + E @ (sm - mps[k + 1])
        sP = fPs[k] + E @ (sP - Pp) @ E.T
        sms[k] = sm
        sPs[k] = sP
    return sms, sPs


def kfs(P0, Fs, Qs, H, R, y):
    """sequential.py:71-73."""
    fms, fPs, mps, Pps, ll = kf(P0, Fs, Qs, H, R, y)
    sms, sPs = ks(Fs, fms, fPs, mps, Pps)
    return fms, fPs, sms, sPs, ll
