// RTS smoothing algebra for D <= 4 (one thread per chunk), reverse-time scan done by index
// reversal (logical index j  <->  time k = n-1-j), never by physically reversing arrays.
//
// Replaces, for the reference's pssgp/kalman/parallel.py:
//   last_smoothing_element     :155-156
//   generic_smoothing_element  :159-166
//   smoothing_operator         :176-184
//   pks                        :187-196 (tf.reverse x5 + scan_associative)
//
// Aggregate layout: E[D*D] | g[D] | L[NS] ; state: sm[D] | sP[NS]
#pragma once
#include "smalld.cuh"
#include "workspace.h"

namespace pssgp {

template <typename T, int D>
struct SmootherAlg {
    using scalar = T;
    static constexpr int KIND = KIND_SMOOTHER;
    static const char* name_reduce() { return "pks_reduce"; }
    static const char* name_mid() { return "pks_mid"; }
    static const char* name_apply() { return "pks_apply"; }
    static constexpr int NS = nsym(D);
    static constexpr int oE = 0, og = D * D, oL = og + D;
    static constexpr int NAGG = oL + NS;
    static constexpr int NSTATE = D + NS;
    static constexpr int NACC = 0;

    struct Params {
        const T* Fs;     // [n, D, D]
        const T* Qs;     // [n, D, D]
        const T* fms;    // [n, D]
        const T* fPs;    // [n, D, D]
        T* sms;          // [n, D]
        T* sPs;          // [n, D, D]
        long n;
        int last_special;   // 1: time n-1 is the global last step (element (0, m, P))
        const T* Fnext;     // [D, D] F of time n (halo) when !last_special
        const T* Qnext;     // [D, D]
        const T* init;      // [NSTATE] smoothed state at time n when !last_special
    };

    PSSGP_DEV static void identity(T* a) {
#pragma unroll
        for (int e = 0; e < NAGG; ++e) a[e] = T(0);
#pragma unroll
        for (int i = 0; i < D; ++i) a[oE + i * D + i] = T(1);
    }

    // element (E, g, L) of time k (parallel.py:159-166)
    PSSGP_DEV static void element(const Params& p, long k, T* E, T* g, T* L) {
        const T* f = (k + 1 < p.n) ? p.Fs + (k + 1) * (D * D) : p.Fnext;
        const T* q = (k + 1 < p.n) ? p.Qs + (k + 1) * (D * D) : p.Qnext;
        T F[D * D], Pp[NS], m[D], P[NS];
#pragma unroll
        for (int e = 0; e < D * D; ++e) F[e] = __ldg(f + e);
        {
            T qf[D * D];
#pragma unroll
            for (int e = 0; e < D * D; ++e) qf[e] = __ldg(q + e);
#pragma unroll
            for (int i = 0; i < D; ++i)
#pragma unroll
                for (int j = 0; j <= i; ++j) Pp[sidx(i, j)] = T(0.5) * (qf[i * D + j] + qf[j * D + i]);
        }
        {
            const T* pm = p.fms + k * D;
            const T* pP = p.fPs + k * (D * D);
#pragma unroll
            for (int i = 0; i < D; ++i) m[i] = __ldg(pm + i);
            T pf[D * D];
#pragma unroll
            for (int e = 0; e < D * D; ++e) pf[e] = __ldg(pP + e);
#pragma unroll
            for (int i = 0; i < D; ++i)
#pragma unroll
                for (int j = 0; j <= i; ++j) P[sidx(i, j)] = T(0.5) * (pf[i * D + j] + pf[j * D + i]);
        }
        T FP[D * D];
        mm_fs<T, D>(F, P, FP);
        sym_xat_plus<T, D>(FP, F, Pp, Pp);  // Pp = FP F^T + Q
        chol_packed<T, D>(Pp);
        T X[D * D];
#pragma unroll
        for (int e = 0; e < D * D; ++e) X[e] = FP[e];
        chol_solve<T, D, D>(Pp, X);  // X = Pp^{-1} F P ; E = X^T
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j < D; ++j) E[i * D + j] = X[j * D + i];
        // g = m - E F m
        T Fm[D], EFm[D];
        mv_f<T, D>(F, m, Fm);
        mv_f<T, D>(E, Fm, EFm);
#pragma unroll
        for (int i = 0; i < D; ++i) g[i] = m[i] - EFm[i];
        // L = sym(P - E Pp E^T) with E Pp E^T = E (F P)
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) {
                T a1 = T(0), a2 = T(0);
#pragma unroll
                for (int kk = 0; kk < D; ++kk) {
                    a1 = fma(E[i * D + kk], FP[kk * D + j], a1);
                    a2 = fma(E[j * D + kk], FP[kk * D + i], a2);
                }
                L[sidx(i, j)] = P[sidx(i, j)] - T(0.5) * (a1 + a2);
            }
    }

    PSSGP_DEV static void load_filtered(const Params& p, long k, T* m, T* P) {
        const T* pm = p.fms + k * D;
        const T* pP = p.fPs + k * (D * D);
#pragma unroll
        for (int i = 0; i < D; ++i) m[i] = __ldg(pm + i);
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) P[sidx(i, j)] = T(0.5) * (__ldg(pP + i * D + j) + __ldg(pP + j * D + i));
    }

    PSSGP_DEV static void append(T* a, long j, const Params& p) {
        const long k = p.n - 1 - j;
        if (j == 0 && p.last_special) {
#pragma unroll
            for (int e = 0; e < D * D; ++e) a[oE + e] = T(0);
            load_filtered(p, k, a + og, a + oL);
            return;
        }
        T E[D * D], g[D], L[NS];
        element(p, k, E, g, L);
        // new = elem_k o agg : E = E_k E_a ; g = E_k g_a + g_k ; L = E_k L_a E_k^T + L_k
        T En[D * D], gn[D], X[D * D], Ln[NS];
        mm_ff<T, D>(E, a + oE, En);
        mv_f<T, D>(E, a + og, gn);
        mm_fs<T, D>(E, a + oL, X);
        sym_xat_plus<T, D>(X, E, L, Ln);
#pragma unroll
        for (int e = 0; e < D * D; ++e) a[oE + e] = En[e];
#pragma unroll
        for (int e = 0; e < D; ++e) a[og + e] = gn[e] + g[e];
#pragma unroll
        for (int e = 0; e < NS; ++e) a[oL + e] = Ln[e];
    }

    // a1 comes first in the reversed sequence (later in time): parallel.py:176-184
    PSSGP_DEV static void combine(const T* a1, const T* a2, T* r) {
        mm_ff<T, D>(a2 + oE, a1 + oE, r + oE);
        T t[D];
        mv_f<T, D>(a2 + oE, a1 + og, t);
#pragma unroll
        for (int i = 0; i < D; ++i) r[og + i] = t[i] + a2[og + i];
        T X[D * D];
        mm_fs<T, D>(a2 + oE, a1 + oL, X);
        sym_xat_plus<T, D>(X, a2 + oE, a2 + oL, r + oL);
    }

    PSSGP_DEV static void apply(const T* s, const T* a, T* s2) {
        T t[D];
        mv_f<T, D>(a + oE, s, t);
#pragma unroll
        for (int i = 0; i < D; ++i) s2[i] = t[i] + a[og + i];
        T X[D * D];
        mm_fs<T, D>(a + oE, s + D, X);
        sym_xat_plus<T, D>(X, a + oE, a + oL, s2 + D);
    }

    PSSGP_DEV static void load_init(const Params& p, T* s) {
#pragma unroll
        for (int e = 0; e < NSTATE; ++e) s[e] = p.init ? p.init[e] : T(0);
    }

    PSSGP_DEV static void step(T* s, long j, const Params& p, T*) {
        const long k = p.n - 1 - j;
        if (j == 0 && p.last_special) {
            load_filtered(p, k, s, s + D);
        } else {
            T a[NAGG], s2[NSTATE];
            element(p, k, a + oE, a + og, a + oL);
            apply(s, a, s2);
#pragma unroll
            for (int e = 0; e < NSTATE; ++e) s[e] = s2[e];
        }
        T* om = p.sms + k * D;
        T* oP = p.sPs + k * (D * D);
#pragma unroll
        for (int i = 0; i < D; ++i) om[i] = s[i];
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int jj = 0; jj < D; ++jj) oP[i * D + jj] = s[D + sidx(i, jj)];
    }

    PSSGP_DEV static void expand_state(const T* s, T* out) {
#pragma unroll
        for (int e = 0; e < NSTATE; ++e) out[e] = s[e];
    }

    PSSGP_DEV static void finish(const Params&, int, T, T*) {}
};

}  // namespace pssgp
