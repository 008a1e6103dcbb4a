// RTS smoothing algebra for D <= 4 (one thread per chunk), reverse-time scan done by walking chunks
// and rows in descending time (scan_stream.cuh), never by physically reversing arrays.
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
    // streaming tables: inputs F, Q, fms, fPs at row k; outputs sms, sPs at row k.  F, Q of step k+1 are
    // carried in registers from the previous (later-in-time) row.
    static constexpr bool REVERSE = true;
    static constexpr int OUT_SHIFT = 0;
    static constexpr bool FLUSH = false;
    static constexpr bool HAS_DONE = false;
    static constexpr bool HAS_SIDE = false;  // fused_small.cuh: extra per-chunk aggregates built by K3
    static constexpr bool OUT8 = false;      // scan_stream.cuh: per-row output staging
    __host__ __device__ static constexpr int out_shift(int) { return OUT_SHIFT; }
    static constexpr int NIN = 4, NOUT = 2, WMAX = D * D;
    __host__ __device__ static constexpr int in_w(int a) { return a == 2 ? D : D * D; }
    __host__ __device__ static constexpr int out_w(int a) { return a == 0 ? D : D * D; }

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
        // time sharding: summaries (E, g, L) of the shards that follow, in rank order, fold_stride scalars apart;
        // folded (last to first) onto `init` (zeros if null) by load_init instead of by pssgp_smoother_fold
        const T* fold = nullptr;
        int fold_count = 0;
        long fold_stride = 0;
    };

    PSSGP_DEV static void identity(T* a) {
#pragma unroll
        for (int e = 0; e < NAGG; ++e) a[e] = T(0);
#pragma unroll
        for (int i = 0; i < D; ++i) a[oE + i * D + i] = T(1);
    }

    __host__ __device__ __forceinline__ static const T* in_ptr(const Params& p, int a) {
        return a == 0 ? p.Fs : (a == 1 ? p.Qs : (a == 2 ? p.fms : p.fPs));
    }
    __host__ __device__ __forceinline__ static T* out_ptr(const Params& p, int a) { return a == 0 ? p.sms : p.sPs; }

    struct Ctx {};
    PSSGP_DEV static void load_ctx(const Params&, Ctx&) {}

    PSSGP_DEV static void sym_pack(const T* qf, T* Q) {
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) Q[sidx(i, j)] = T(0.5) * (qf[i * D + j] + qf[j * D + i]);
    }

    // element (E, g, L) of time k (parallel.py:159-166) from F_{k+1}, Q_{k+1} and m_k, P_k (Q, P packed)
    PSSGP_DEV static void element(const T* F, const T* Q, const T* m, const T* P, T* E, T* g, T* L) {
        T Pp[NS];
#pragma unroll
        for (int e = 0; e < NS; ++e) Pp[e] = Q[e];
        T FP[D * D];
        mm_fs<T, D>(F, P, FP);
        sym_xat_plus<T, D>(FP, F, Pp, Pp);  // Pp = FP F^T + Q
        ldl_packed<T, D>(Pp);
        T X[D * D];
#pragma unroll
        for (int e = 0; e < D * D; ++e) X[e] = FP[e];
        ldl_solve<T, D, D>(Pp, X);  // X = Pp^{-1} F P ; E = X^T
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

    // F, Q of the step after the row being visited
    struct Carry {
        T F[D * D];
        T Q[NS];
    };
    PSSGP_DEV static void carry_load(Carry& c, const T* f, const T* q) {
        T qf[D * D];
#pragma unroll
        for (int e = 0; e < D * D; ++e) {
            c.F[e] = __ldg(f + e);
            qf[e] = __ldg(q + e);
        }
        sym_pack(qf, c.Q);
    }
    // k_hi = one past the chunk's last (first visited) row: its F, Q come from the next chunk or the halo
    PSSGP_DEV static void carry_init(Carry& c, const Ctx&, long, long k_hi, const Params& p) {
        if (k_hi < p.n)
            carry_load(c, p.Fs + k_hi * (D * D), p.Qs + k_hi * (D * D));
        else if (!p.last_special)
            carry_load(c, p.Fnext, p.Qnext);
    }
    PSSGP_DEV static void carry_set(Carry& c, const T (&in)[NIN][WMAX]) {
#pragma unroll
        for (int e = 0; e < D * D; ++e) c.F[e] = in[0][e];
        sym_pack(in[1], c.Q);
    }

    PSSGP_DEV static void append_row(T* a, const Ctx&, const T (&in)[NIN][WMAX], long k, const Params& p, Carry& c) {
        if (k == p.n - 1 && p.last_special) {
            // last_smoothing_element (parallel.py:155-156): nothing has been appended before it
#pragma unroll
            for (int e = 0; e < D * D; ++e) a[oE + e] = T(0);
#pragma unroll
            for (int i = 0; i < D; ++i) a[og + i] = in[2][i];
            sym_pack(in[3], a + oL);
            carry_set(c, in);
            return;
        }
        T E[D * D], g[D], L[NS], P[NS];
        sym_pack(in[3], P);
        element(c.F, c.Q, in[2], P, E, g, L);
        carry_set(c, in);
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
#pragma unroll 1
        for (int i = p.fold_count - 1; i >= 0; --i) {
            T b[NAGG], s2[NSTATE];
#pragma unroll
            for (int e = 0; e < NAGG; ++e) b[e] = p.fold[(long)i * p.fold_stride + e];
            apply(s, b, s2);
#pragma unroll
            for (int e = 0; e < NSTATE; ++e) s[e] = s2[e];
        }
    }

    // s = smoothed state at time k+1 -> at time k
    PSSGP_DEV static void advance(T* s, const T (&in)[NIN][WMAX], long k, const Params& p, Carry& c) {
        if (k == p.n - 1 && p.last_special) {
#pragma unroll
            for (int i = 0; i < D; ++i) s[i] = in[2][i];
            sym_pack(in[3], s + D);
        } else {
            T a[NAGG], s2[NSTATE], P[NS];
            sym_pack(in[3], P);
            element(c.F, c.Q, in[2], P, a + oE, a + og, a + oL);
            apply(s, a, s2);
#pragma unroll
            for (int e = 0; e < NSTATE; ++e) s[e] = s2[e];
        }
        carry_set(c, in);
    }

    PSSGP_DEV static bool step_row(T* s, const Ctx&, const T (&in)[NIN][WMAX], T (&out)[NOUT][WMAX], long k,
                                   const Params& p, T*, Carry& c) {
        advance(s, in, k, p, c);
#pragma unroll
        for (int i = 0; i < D; ++i) out[0][i] = s[i];
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int jj = 0; jj < D; ++jj) out[1][i * D + jj] = s[D + sidx(i, jj)];
        return true;
    }

    PSSGP_DEV static void expand_state(const T* s, T* out) {
#pragma unroll
        for (int e = 0; e < NSTATE; ++e) out[e] = s[e];
    }

    PSSGP_DEV static void finish(const Params&, int, T, T*) {}
};

// RTS smoother that emits only the projection of the smoothed state on the observation row H: proj[k] = (H m_k,
// H P_k H^T) — what predict_f keeps of (sms, sPs) (pssgp/model.py:107-111), 16 bytes per step instead of
// 8 d (d + 1).
template <typename T, int D>
struct SmootherProjAlg : SmootherAlg<T, D> {
    using Base = SmootherAlg<T, D>;
    static const char* name_apply() { return "pks_apply_proj"; }
    static constexpr int NIN = Base::NIN, NOUT = 1;
    static constexpr int WMAX = Base::WMAX < 2 ? 2 : Base::WMAX;  // the output row has two values (d = 1: D * D = 1)
    __host__ __device__ static constexpr int out_w(int) { return 2; }
    struct Params : Base::Params {
        const T* H;   // [D]
        T* proj;      // [n, 2]
    };
    __host__ __device__ __forceinline__ static T* out_ptr(const Params& p, int) { return p.proj; }
    struct Ctx {
        T h[D];
    };
    PSSGP_DEV static void load_ctx(const Params& p, Ctx& c) {
#pragma unroll
        for (int i = 0; i < D; ++i) c.h[i] = __ldg(p.H + i);
    }
    using Carry = typename Base::Carry;
    PSSGP_DEV static void carry_init(Carry& c, const Ctx&, long k_lo, long k_hi, const Params& p) {
        Base::carry_init(c, typename Base::Ctx{}, k_lo, k_hi, p);
    }
    PSSGP_DEV static bool step_row(T* s, const Ctx& cx, const T (&in)[NIN][WMAX], T (&out)[NOUT][WMAX], long k,
                                   const Params& p, T*, Carry& c) {
        T inb[NIN][Base::WMAX];
#pragma unroll
        for (int a = 0; a < NIN; ++a)
#pragma unroll
            for (int e = 0; e < Base::WMAX; ++e) inb[a][e] = in[a][e];
        Base::advance(s, inb, k, p, c);
        T Ph[D];
        mv_s<T, D>(s + D, cx.h, Ph);
        out[0][0] = dot<T, D>(cx.h, s);
        out[0][1] = dot<T, D>(cx.h, Ph);
        return true;
    }
};

}  // namespace pssgp
