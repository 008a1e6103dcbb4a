// Filtering algebra for D <= 4 (one thread per chunk).
//
// Replaces, for the reference's pssgp/kalman/parallel.py:
//   first_filtering_element           :13-43    (step 0: update on (m0,P0) without prediction)
//   _generic_filtering_element[_nan]  :46-72    (element (A,b,C,J,eta) of one step)
//   filtering_operator                :100-118  (combine)
//   pkf's log-likelihood block        :135-151  (step(): one-step-ahead predictive density)
//
// Aggregate layout (packed): A[D*D] | b[D] | C[NS] | J[NS] | eta[D]   (C, J symmetric packed)
// State layout: m[D] | P[NS]
#pragma once
#include "smalld.cuh"
#include "workspace.h"

namespace pssgp {

template <typename T, int D>
struct FilterAlg {
    using scalar = T;
    static constexpr int KIND = KIND_FILTER;
    static const char* name_reduce() { return "pkf_reduce"; }
    static const char* name_mid() { return "pkf_mid"; }
    static const char* name_apply() { return "pkf_apply"; }
    static constexpr int NS = nsym(D);
    static constexpr int oA = 0, ob = D * D, oC = ob + D, oJ = oC + NS, oE = oJ + NS;
    static constexpr int NAGG = oE + D;
    static constexpr int NSTATE = D + NS;
    static constexpr int NACC = 1;
    // streaming tables (scan_stream.cuh): inputs F, Q, y at row k; outputs fms, fPs at row k
    static constexpr bool REVERSE = false;
    static constexpr int NIN = 3, NOUT = 2, WMAX = D * D;
    __host__ __device__ static constexpr int in_w(int a) { return a < 2 ? D * D : 1; }
    static constexpr int OUT_SHIFT = 0;
    static constexpr bool FLUSH = false;
    static constexpr bool HAS_DONE = true;
    static constexpr bool HAS_SIDE = false;  // fused_small.cuh: extra per-chunk aggregates built by K3  // step_done folds the chunk's log-likelihood pieces into acc
    static constexpr bool OUT8 = false;      // scan_stream.cuh: per-row output staging
    __host__ __device__ static constexpr int out_shift(int) { return OUT_SHIFT; }
    __host__ __device__ static constexpr int out_w(int a) { return a == 0 ? D : D * D; }

    struct Params {
        const T* Fs;   // [n, D, D]
        const T* Qs;   // [n, D, D]
        const T* y;    // [n]
        const T* H;    // [D]
        const T* R;    // [1]
        const T* P0;   // [D, D]
        const T* m0;   // [D] or null (zeros)
        T* fms;        // [n, D]
        T* fPs;        // [n, D, D]
        int first_special;  // 1: logical step 0 is the global first step (parallel.py:13-43 semantics)
        // time sharding: summaries (A, b, C, J, eta) of the shards BEFORE this one, in rank order, fold_stride scalars
        // apart; folded (first to last) onto (m0, P0) by load_init instead of by pssgp_filter_fold
        const T* fold = nullptr;
        int fold_count = 0;
        long fold_stride = 0;
    };

    PSSGP_DEV static void identity(T* a) {
#pragma unroll
        for (int e = 0; e < NAGG; ++e) a[e] = T(0);
#pragma unroll
        for (int i = 0; i < D; ++i) a[oA + i * D + i] = T(1);
    }

    __host__ __device__ __forceinline__ static const T* in_ptr(const Params& p, int a) { return a == 0 ? p.Fs : (a == 1 ? p.Qs : p.y); }
    __host__ __device__ __forceinline__ static T* out_ptr(const Params& p, int a) { return a == 0 ? p.fms : p.fPs; }

    struct Ctx {
        T h[D];
        T R;
    };
    PSSGP_DEV static void load_ctx(const Params& p, Ctx& c) {
#pragma unroll
        for (int i = 0; i < D; ++i) c.h[i] = __ldg(p.H + i);
        c.R = __ldg(p.R);
    }

    // Q of one row, symmetrised and packed
    PSSGP_DEV static void sym_pack(const T* qf, T* Q) {
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) Q[sidx(i, j)] = T(0.5) * (qf[i * D + j] + qf[j * D + i]);
    }

    // Append time step k (row r of the staged tile) to the aggregate: conditional Kalman recursion given
    // the chunk-entry state.
    // per-chunk pieces of the log-likelihood (K3): sum of log innovation variances, quadratic term, count
    struct Carry {
        LogSum<T> ls;
        T quad;
        int nobs;
    };
    PSSGP_DEV static void carry_init(Carry& c, const Ctx&, long, long, const Params&) {
        c.ls.init();
        c.quad = T(0);
        c.nobs = 0;
    }
    PSSGP_DEV static void step_done(T* acc, const Carry& c) {
        acc[0] = T(-0.5) * (c.ls.value() + T(c.nobs) * T(1.8378770664093454835606594728112) + c.quad);
    }

    PSSGP_DEV static void append_row(T* a, const Ctx& cx, const T (&in)[NIN][WMAX], long k, const Params& p, Carry&) {
        const T* F = in[0];
        const T* h = cx.h;
        const T R = cx.R;
        const T yk = in[2][0];
        T Ap[D * D], bp[D], Cp[NS];
        if (k == 0 && p.first_special) {
            // no prediction at the global first step (parallel.py:24-30)
#pragma unroll
            for (int e = 0; e < D * D; ++e) Ap[e] = a[oA + e];
#pragma unroll
            for (int e = 0; e < D; ++e) bp[e] = a[ob + e];
#pragma unroll
            for (int e = 0; e < NS; ++e) Cp[e] = a[oC + e];
        } else {
            T Q[NS], FC[D * D];
            sym_pack(in[1], Q);
            mm_ff<T, D>(F, a + oA, Ap);
            mv_f<T, D>(F, a + ob, bp);
            mm_fs<T, D>(F, a + oC, FC);
            sym_xat_plus<T, D>(FC, F, Q, Cp);
        }
        if (t_isnan(yk)) {
#pragma unroll
            for (int e = 0; e < D * D; ++e) a[oA + e] = Ap[e];
#pragma unroll
            for (int e = 0; e < D; ++e) a[ob + e] = bp[e];
#pragma unroll
            for (int e = 0; e < NS; ++e) a[oC + e] = Cp[e];
            return;
        }
        T u[D], w[D];
        mv_s<T, D>(Cp, h, u);
        const T s = dot<T, D>(h, u) + R;
        mv_t<T, D>(Ap, h, w);  // (H A')^T
        const T e0 = yk - dot<T, D>(h, bp);
        const T is = t_rcp(s);
#pragma unroll
        for (int i = 0; i < D; ++i) {
            a[oE + i] = fma(w[i], e0 * is, a[oE + i]);
            a[ob + i] = fma(u[i], e0 * is, bp[i]);
        }
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) {
                a[oJ + sidx(i, j)] = fma(w[i] * is, w[j], a[oJ + sidx(i, j)]);
                a[oC + sidx(i, j)] = fma(-u[i] * is, u[j], Cp[sidx(i, j)]);
            }
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j < D; ++j) a[oA + i * D + j] = fma(-u[i] * is, w[j], Ap[i * D + j]);
    }

    // r = a1 (earlier) o a2 (later): parallel.py:100-118 with one LU of M = I + C1 J2 serving both solves.
    PSSGP_DEV static void combine(const T* a1, const T* a2, T* r) {
        constexpr int NR = 2 * D + 1;
        T M[D * D], B[D * NR];
        // M = I + C1 J2
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j < D; ++j) {
                T acc = (i == j) ? T(1) : T(0);
#pragma unroll
                for (int k = 0; k < D; ++k) acc = fma(a1[oC + sidx(i, k)], a2[oJ + sidx(k, j)], acc);
                M[i * D + j] = acc;
            }
        // B = [A1 | b1 + C1 eta2 | C1 A2^T]
        T c1e[D];
        mv_s<T, D>(a1 + oC, a2 + oE, c1e);
#pragma unroll
        for (int i = 0; i < D; ++i) {
#pragma unroll
            for (int j = 0; j < D; ++j) B[i * NR + j] = a1[oA + i * D + j];
            B[i * NR + D] = a1[ob + i] + c1e[i];
#pragma unroll
            for (int j = 0; j < D; ++j) {
                T acc = T(0);
#pragma unroll
                for (int k = 0; k < D; ++k) acc = fma(a1[oC + sidx(i, k)], a2[oA + j * D + k], acc);
                B[i * NR + D + 1 + j] = acc;
            }
        }
        lu_solve<T, D, NR>(M, B);
        T ZA[D * D], zb[D], ZC[D * D];
#pragma unroll
        for (int i = 0; i < D; ++i) {
#pragma unroll
            for (int j = 0; j < D; ++j) {
                ZA[i * D + j] = B[i * NR + j];
                ZC[i * D + j] = B[i * NR + D + 1 + j];
            }
            zb[i] = B[i * NR + D];
        }
        // A = A2 ZA ; b = A2 zb + b2 ; C = sym(A2 ZC) + C2
        mm_ff<T, D>(a2 + oA, ZA, r + oA);
        T t1[D];
        mv_f<T, D>(a2 + oA, zb, t1);
#pragma unroll
        for (int i = 0; i < D; ++i) r[ob + i] = t1[i] + a2[ob + i];
        sym_half_ab_plus<T, D>(a2 + oA, ZC, a2 + oC, r + oC);
        // eta = A1^T (eta2 - J2 zb) + eta1 ; J = sym(A1^T J2 ZA) + J1
        T j2z[D], t2[D];
        mv_s<T, D>(a2 + oJ, zb, j2z);
#pragma unroll
        for (int i = 0; i < D; ++i) t2[i] = a2[oE + i] - j2z[i];
        mv_t<T, D>(a1 + oA, t2, t1);
#pragma unroll
        for (int i = 0; i < D; ++i) r[oE + i] = t1[i] + a1[oE + i];
        T J2ZA[D * D];
        mm_sf<T, D>(a2 + oJ, ZA, J2ZA);
        sym_half_atb_plus<T, D>(a1 + oA, J2ZA, a1 + oJ, r + oJ);
    }

    // s2 = state o aggregate  (a prefix that starts at the origin has A = 0: only (b, C) = (m, P) matter)
    PSSGP_DEV static void apply(const T* s, const T* a, T* s2) {
        constexpr int NR = D + 1;
        T M[D * D], B[D * NR];
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j < D; ++j) {
                T acc = (i == j) ? T(1) : T(0);
#pragma unroll
                for (int k = 0; k < D; ++k) acc = fma(s[D + sidx(i, k)], a[oJ + sidx(k, j)], acc);
                M[i * D + j] = acc;
            }
        T pe[D];
        mv_s<T, D>(s + D, a + oE, pe);
#pragma unroll
        for (int i = 0; i < D; ++i) {
            B[i * NR] = s[i] + pe[i];
#pragma unroll
            for (int j = 0; j < D; ++j) {
                T acc = T(0);
#pragma unroll
                for (int k = 0; k < D; ++k) acc = fma(s[D + sidx(i, k)], a[oA + j * D + k], acc);
                B[i * NR + 1 + j] = acc;
            }
        }
        lu_solve<T, D, NR>(M, B);
        T zb[D], ZC[D * D];
#pragma unroll
        for (int i = 0; i < D; ++i) {
            zb[i] = B[i * NR];
#pragma unroll
            for (int j = 0; j < D; ++j) ZC[i * D + j] = B[i * NR + 1 + j];
        }
        T t1[D];
        mv_f<T, D>(a + oA, zb, t1);
#pragma unroll
        for (int i = 0; i < D; ++i) s2[i] = t1[i] + a[ob + i];
        sym_half_ab_plus<T, D>(a + oA, ZC, a + oC, s2 + D);
    }

    PSSGP_DEV static void load_init(const Params& p, T* s) {
#pragma unroll
        for (int i = 0; i < D; ++i) s[i] = p.m0 ? p.m0[i] : T(0);
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) s[D + sidx(i, j)] = T(0.5) * (p.P0[i * D + j] + p.P0[j * D + i]);
#pragma unroll 1
        for (int i = 0; i < p.fold_count; ++i) {
            T b[NAGG], s2[NSTATE];
#pragma unroll
            for (int e = 0; e < NAGG; ++e) b[e] = p.fold[(long)i * p.fold_stride + e];
            apply(s, b, s2);
#pragma unroll
            for (int e = 0; e < NSTATE; ++e) s[e] = s2[e];
        }
    }

    // Seeded Kalman step k: s=(m,P) filtered at k-1 -> filtered at k; emits fms/fPs, accumulates ll.
    PSSGP_DEV static bool step_row(T* s, const Ctx& cx, const T (&in)[NIN][WMAX], T (&out)[NOUT][WMAX], long k,
                                   const Params& p, T* acc, Carry& cr) {
        const T* F = in[0];
        const T* h = cx.h;
        const T R = cx.R;
        const T yk = in[2][0];
        T Q[NS], FP[D * D], mp[D], Pp[NS];
        sym_pack(in[1], Q);
        mv_f<T, D>(F, s, mp);
        mm_fs<T, D>(F, s + D, FP);
        sym_xat_plus<T, D>(FP, F, Q, Pp);
        const bool obs = !t_isnan(yk);
        T u[D];
        mv_s<T, D>(Pp, h, u);
        T sv = dot<T, D>(h, u) + R;
        T e0 = yk - dot<T, D>(h, mp);
        T is = t_rcp(sv);
        if (obs) {
            // log N(y; H mp, sv): the log-determinant part goes through the running LogSum of the chunk
            cr.ls.add(sv);
            cr.quad = fma(e0 * e0, is, cr.quad);
            cr.nobs += 1;
        }
        if (k == 0 && p.first_special) {
            // the filter update at the global first step acts on (m0, P0) directly (parallel.py:24-30)
#pragma unroll
            for (int i = 0; i < D; ++i) mp[i] = s[i];
#pragma unroll
            for (int e = 0; e < NS; ++e) Pp[e] = s[D + e];
            mv_s<T, D>(Pp, h, u);
            sv = dot<T, D>(h, u) + R;
            e0 = yk - dot<T, D>(h, mp);
            is = t_rcp(sv);
        }
        if (obs) {
#pragma unroll
            for (int i = 0; i < D; ++i) s[i] = fma(u[i], e0 * is, mp[i]);
#pragma unroll
            for (int i = 0; i < D; ++i)
#pragma unroll
                for (int j = 0; j <= i; ++j) s[D + sidx(i, j)] = fma(-u[i] * is, u[j], Pp[sidx(i, j)]);
        } else {
#pragma unroll
            for (int i = 0; i < D; ++i) s[i] = mp[i];
#pragma unroll
            for (int e = 0; e < NS; ++e) s[D + e] = Pp[e];
        }
#pragma unroll
        for (int i = 0; i < D; ++i) out[0][i] = s[i];
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j < D; ++j) out[1][i * D + j] = s[D + sidx(i, j)];
        return true;
    }

    // fold output: m[D] | P full [D,D]  (so that the caller can pass m0 = out, P0 = out + D to pssgp_pkf)
    PSSGP_DEV static void expand_state(const T* s, T* out) {
#pragma unroll
        for (int i = 0; i < D; ++i) out[i] = s[i];
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j < D; ++j) out[D + i * D + j] = s[D + sidx(i, j)];
    }

    PSSGP_DEV static void finish(const Params&, int, T tot, T* acc_out) {
        if (acc_out) acc_out[0] = tot;
    }
};

}  // namespace pssgp
