// Adjoint (reverse-mode) algebra of the filter log-likelihood for D <= 4.
//
// The reference obtains d ll / d(P0, Fs, Qs, H, R) by TF autodiff through pkf
// (pssgp/kalman/parallel.py:121-152; contract: tests/test_gp_vs_kfs.py:53-67).  Here the adjoint of
// the mathematically identical sequential recursion is itself run as a reverse associative scan:
// the adjoint state (dm, dP) w.r.t. the filtered moments propagates backwards through the affine map
//     dm' = Abar^T dm + a ,   dP' = Abar^T dP Abar + sym(Abar^T dm a^T) + B
// whose compositions are closed under
//     (Abar1 Abar2,  Abar2^T a1 + a2,  Abar2^T B1 Abar2 + sym(Abar2^T a1 a2^T) + B2)      (1 = later in time)
// (matmul only, no solves).  Everything is computed for an upstream gradient of 1 and scaled by
// g_ll when written.
//
// Aggregate layout: Abar[D*D] | a[D] | B[NS] ; state: dm[D] | dP[NS] ; accumulators: dR, dH[D]
#pragma once
#include "smalld.cuh"
#include "workspace.h"

namespace pssgp {

template <typename T, int D>
struct AdjointAlg {
    using scalar = T;
    static constexpr int KIND = KIND_ADJOINT;
    static const char* name_reduce() { return "pkf_bwd_reduce"; }
    static const char* name_mid() { return "pkf_bwd_mid"; }
    static const char* name_apply() { return "pkf_bwd_apply"; }
    static constexpr int NS = nsym(D);
    static constexpr int oA = 0, oa = D * D, oB = oa + D;
    static constexpr int NAGG = oB + NS;
    static constexpr int NSTATE = D + NS;
    static constexpr int NACC = 1 + D;
    // streaming tables: inputs F, Q, y, fms, fPs at row k; outputs dFs, dQs.  Step k needs the filtered
    // moments of row k-1, so it is taken one row late: when row k-1 is visited, with (F, Q, y) of row k
    // carried in registers (OUT_SHIFT = 1); the chunk's first row is finished by step_flush from a halo load.
    static constexpr bool REVERSE = true;
    static constexpr int OUT_SHIFT = 1;
    static constexpr bool FLUSH = true;
    static constexpr bool HAS_DONE = false;
    static constexpr bool HAS_SIDE = false;  // fused_small.cuh: extra per-chunk aggregates built by K3
    static constexpr bool OUT8 = false;      // scan_stream.cuh: per-row output staging
    __host__ __device__ static constexpr int out_shift(int) { return OUT_SHIFT; }
    static constexpr int NIN = 5, NOUT = 2, WMAX = D * D;
    __host__ __device__ static constexpr int in_w(int a) { return a == 2 ? 1 : (a == 3 ? D : D * D); }
    __host__ __device__ static constexpr int out_w(int) { return D * D; }

    struct Params {
        const T* Fs;
        const T* Qs;
        const T* y;
        const T* H;
        const T* R;
        const T* P0;      // [D,D] prior covariance (first_special) or filtered covariance before the shard
        const T* m0;      // [D] or null
        const T* fms;     // [n,D]   filtered means from the forward pass
        const T* fPs;     // [n,D,D]
        const T* g;       // [1] upstream gradient of ll
        const T* init;    // [NSTATE] adjoint state entering from the right (null = zeros)
        T* dFs;           // [n,D,D]
        T* dQs;           // [n,D,D]
        T* dP0;           // [D,D]
        T* dH;            // [D]
        T* dR;            // [1]
        T* first_state;   // unused here (returned through final_state of the framework)
        long n;
        int first_special;
        // time sharding: summaries (Abar, a, B) of the shards that follow, in rank order, fold_stride scalars apart;
        // folded (last to first) onto `init` (zeros if null) by load_init instead of by pssgp_adjoint_fold
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

    struct Fwd {
        T F[D * D];
        T m[D], P[NS];     // filtered at k-1
        T mp[D], Pp[NS];   // predicted at k
        T h[D];
        T u[D];            // Pp h
        T s, r, R, yk;
        bool obs, first;
    };

    __host__ __device__ __forceinline__ static const T* in_ptr(const Params& p, int a) {
        return a == 0 ? p.Fs : (a == 1 ? p.Qs : (a == 2 ? p.y : (a == 3 ? p.fms : p.fPs)));
    }
    __host__ __device__ __forceinline__ static T* out_ptr(const Params& p, int a) { return a == 0 ? p.dFs : p.dQs; }

    struct Ctx {
        T h[D];
        T R, g;
    };
    PSSGP_DEV static void load_ctx(const Params& p, Ctx& c) {
#pragma unroll
        for (int i = 0; i < D; ++i) c.h[i] = __ldg(p.H + i);
        c.R = __ldg(p.R);
        c.g = p.g ? __ldg(p.g) : T(1);
    }

    // (F, Q, y) of the step that is taken when the next (earlier) row is visited
    struct Carry {
        T F[D * D];
        T Q[NS];
        T y;
        bool has;
    };
    PSSGP_DEV static void carry_init(Carry& c, const Ctx&, long, long, const Params&) { c.has = false; }
    PSSGP_DEV static void carry_set(Carry& c, const T (&in)[NIN][WMAX], long k, const Params& p) {
#pragma unroll
        for (int e = 0; e < D * D; ++e) c.F[e] = in[0][e];
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) c.Q[sidx(i, j)] = T(0.5) * (in[1][i * D + j] + in[1][j * D + i]);
        c.y = in[2][0];
        c.has = true;
    }
    // filtered moments of row k-1 straight from global memory (chunk boundary), or the prior at k = 0
    PSSGP_DEV static void halo_moments(long k, const Params& p, T* m, T* P) {
        if (k > 0) {
            const T* pm = p.fms + (k - 1) * D;
            const T* pP = p.fPs + (k - 1) * (D * D);
#pragma unroll
            for (int i = 0; i < D; ++i) m[i] = __ldg(pm + i);
#pragma unroll
            for (int i = 0; i < D; ++i)
#pragma unroll
                for (int j = 0; j <= i; ++j) P[sidx(i, j)] = T(0.5) * (__ldg(pP + i * D + j) + __ldg(pP + j * D + i));
        } else {
#pragma unroll
            for (int i = 0; i < D; ++i) m[i] = p.m0 ? p.m0[i] : T(0);
#pragma unroll
            for (int i = 0; i < D; ++i)
#pragma unroll
                for (int j = 0; j <= i; ++j) P[sidx(i, j)] = T(0.5) * (p.P0[i * D + j] + p.P0[j * D + i]);
        }
    }
    PSSGP_DEV static void row_moments(const T (&in)[NIN][WMAX], T* m, T* P) {
#pragma unroll
        for (int i = 0; i < D; ++i) m[i] = in[3][i];
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) P[sidx(i, j)] = T(0.5) * (in[4][i * D + j] + in[4][j * D + i]);
    }

    // Recomputes the forward quantities of time step k from (F, Q, y) of row k and the filtered moments
    // (m, P) of row k-1.
    PSSGP_DEV static void forward(const Ctx& cx, const Carry& c, const T* m, const T* P, long k, const Params& p,
                                  Fwd& f) {
#pragma unroll
        for (int i = 0; i < D; ++i) f.h[i] = cx.h[i];
        f.R = cx.R;
        f.yk = c.y;
        f.obs = !t_isnan(f.yk);
        f.first = (k == 0 && p.first_special);
#pragma unroll
        for (int e = 0; e < D * D; ++e) f.F[e] = c.F[e];
#pragma unroll
        for (int i = 0; i < D; ++i) f.m[i] = m[i];
#pragma unroll
        for (int e = 0; e < NS; ++e) f.P[e] = P[e];
        T FP[D * D];
        mv_f<T, D>(f.F, f.m, f.mp);
        mm_fs<T, D>(f.F, f.P, FP);
        sym_xat_plus<T, D>(FP, f.F, c.Q, f.Pp);
        mv_s<T, D>(f.Pp, f.h, f.u);
        f.s = dot<T, D>(f.h, f.u) + f.R;
        f.r = f.yk - dot<T, D>(f.h, f.mp);
    }

    // Element (Abar, a, B) of time step k for an upstream gradient of 1.
    PSSGP_DEV static void element(const Ctx& cx, const Carry& c, const T* m, const T* P, long k, const Params& p,
                                  T* x) {
        Fwd f;
        forward(cx, c, m, P, k, p, f);
        if (f.first) {
            // filter update acts on (m0, P0) directly, no likelihood term through this path
            identity(x);
            if (f.obs) {
                T u0[D];
                mv_s<T, D>(f.P, f.h, u0);
                const T s0 = dot<T, D>(f.h, u0) + f.R;
                const T is = t_rcp(s0);
#pragma unroll
                for (int i = 0; i < D; ++i)
#pragma unroll
                    for (int j = 0; j < D; ++j) x[oA + i * D + j] -= u0[i] * is * f.h[j];
            }
            return;
        }
        if (!f.obs) {
#pragma unroll
            for (int e = 0; e < D * D; ++e) x[oA + e] = f.F[e];
#pragma unroll
            for (int e = 0; e < D; ++e) x[oa + e] = T(0);
#pragma unroll
            for (int e = 0; e < NS; ++e) x[oB + e] = T(0);
            return;
        }
        const T is = t_rcp(f.s);
        T w[D];  // F^T h
        mv_t<T, D>(f.F, f.h, w);
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j < D; ++j) x[oA + i * D + j] = fma(-f.u[i] * is, w[j], f.F[i * D + j]);
        const T ris = f.r * is;
#pragma unroll
        for (int i = 0; i < D; ++i) x[oa + i] = w[i] * ris;
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) x[oB + sidx(i, j)] = T(0.5) * (ris * ris - is) * w[i] * w[j];
    }

    // x1 = later in time (earlier in the reversed sequence)
    PSSGP_DEV static void combine(const T* x1, const T* x2, T* r) {
        mm_ff<T, D>(x1 + oA, x2 + oA, r + oA);
        T t[D];
        mv_t<T, D>(x2 + oA, x1 + oa, t);
#pragma unroll
        for (int i = 0; i < D; ++i) r[oa + i] = t[i] + x2[oa + i];
        T X[D * D];  // B1 A2
        mm_sf<T, D>(x1 + oB, x2 + oA, X);
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) {
                T acc = x2[oB + sidx(i, j)] + T(0.5) * (t[i] * x2[oa + j] + t[j] * x2[oa + i]);
#pragma unroll
                for (int kk = 0; kk < D; ++kk) acc = fma(x2[oA + kk * D + i], X[kk * D + j], acc);
                r[oB + sidx(i, j)] = acc;
            }
    }

    PSSGP_DEV static void append_step(T* a, const Ctx& cx, const Carry& c, const T* m, const T* P, long k,
                                      const Params& p) {
        T x[NAGG], rr[NAGG];
        element(cx, c, m, P, k, p, x);
        combine(a, x, rr);
#pragma unroll
        for (int e = 0; e < NAGG; ++e) a[e] = rr[e];
    }
    // visiting row k: take step k+1 (if its inputs are carried), then carry (F, Q, y) of row k
    PSSGP_DEV static void append_row(T* a, const Ctx& cx, const T (&in)[NIN][WMAX], long k, const Params& p, Carry& c) {
        if (c.has) {
            T m[D], P[NS];
            row_moments(in, m, P);
            append_step(a, cx, c, m, P, k + 1, p);
        }
        carry_set(c, in, k, p);
    }
    // step k_lo of the chunk, with the filtered moments of row k_lo - 1 from the halo
    PSSGP_DEV static void append_flush(T* a, const Ctx& cx, long k_lo, const Params& p, Carry& c) {
        if (c.has) {
            T m[D], P[NS];
            halo_moments(k_lo, p, m, P);
            append_step(a, cx, c, m, P, k_lo, p);
        }
    }

    PSSGP_DEV static void apply(const T* s, const T* x, T* s2) {
        T t[D];
        mv_t<T, D>(x + oA, s, t);
#pragma unroll
        for (int i = 0; i < D; ++i) s2[i] = t[i] + x[oa + i];
        T X[D * D];
        mm_sf<T, D>(s + D, x + oA, X);
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) {
                T acc = x[oB + sidx(i, j)] + T(0.5) * (t[i] * x[oa + j] + t[j] * x[oa + i]);
#pragma unroll
                for (int kk = 0; kk < D; ++kk) acc = fma(x[oA + kk * D + i], X[kk * D + j], acc);
                s2[D + sidx(i, j)] = acc;
            }
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

    // Adjoint of the measurement update given (mp, Pp, u, s, r): in (dm+, dP+) -> out (dmp, dPp),
    // accumulating dR and dH.  with_ll: include the log-density term of this step.
    PSSGP_DEV static void update_adjoint(const T* h, const T* mp, const T* Pp, const T* u, T s, T r, bool with_ll,
                                         const T* dm, const T* dP, T* dmp, T* dPp, T* acc) {
        const T is = t_rcp(s);
        const T udm = dot<T, D>(u, dm);
        T Pu[D];
        mv_s<T, D>(dP, u, Pu);
        const T uPu = dot<T, D>(u, Pu);
        T rbar = udm * is;
        T sbar = (-udm * r + uPu) * is * is;
        if (with_ll) {
            rbar -= r * is;
            sbar += T(0.5) * (r * r * is * is - is);
        }
        T ut[D];
#pragma unroll
        for (int i = 0; i < D; ++i) ut[i] = dm[i] * r * is - T(2) * Pu[i] * is + sbar * h[i];
        T Pput[D];
        mv_s<T, D>(Pp, ut, Pput);
        acc[0] += sbar;
#pragma unroll
        for (int i = 0; i < D; ++i) acc[1 + i] += sbar * u[i] + Pput[i] - mp[i] * rbar;
#pragma unroll
        for (int i = 0; i < D; ++i) dmp[i] = dm[i] - h[i] * rbar;
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) dPp[sidx(i, j)] = dP[sidx(i, j)] + T(0.5) * (ut[i] * h[j] + ut[j] * h[i]);
    }

    // s = adjoint w.r.t. the filtered moments at time k; on exit w.r.t. those at time k-1.
    PSSGP_DEV static void step_core(T* s, const Ctx& cx, const Carry& c, const T* mprev, const T* Pprev,
                                    T (&out)[NOUT][WMAX], long k, const Params& p, T* acc) {
        const T g = cx.g;
        Fwd f;
        forward(cx, c, mprev, Pprev, k, p, f);
        T dmp[D], dPp[NS];
        T* oF = out[0];
        T* oQ = out[1];
        if (f.first) {
            // (i) likelihood term of step 0 through the prediction from (m0, P0)
            T dPp0[NS], dmp0[D];
#pragma unroll
            for (int e = 0; e < NS; ++e) dPp0[e] = T(0);
#pragma unroll
            for (int e = 0; e < D; ++e) dmp0[e] = T(0);
            if (f.obs) {
                const T is = t_rcp(f.s);
                const T sbar = T(0.5) * (f.r * f.r * is * is - is);
                const T rbar = -f.r * is;
                acc[0] += sbar;
#pragma unroll
                for (int i = 0; i < D; ++i) acc[1 + i] += T(2) * sbar * f.u[i] - f.mp[i] * rbar;
#pragma unroll
                for (int i = 0; i < D; ++i) dmp0[i] = -f.h[i] * rbar;
#pragma unroll
                for (int i = 0; i < D; ++i)
#pragma unroll
                    for (int jj = 0; jj <= i; ++jj) dPp0[sidx(i, jj)] = sbar * f.h[i] * f.h[jj];
            }
            T X[D * D], Y[D * D];
            mm_sf<T, D>(dPp0, f.F, X);   // dPp0 F
            mm_fs<T, D>(X, f.P, Y);      // dPp0 F P
#pragma unroll
            for (int i = 0; i < D; ++i)
#pragma unroll
                for (int jj = 0; jj < D; ++jj) {
                    oF[i * D + jj] = g * (dmp0[i] * f.m[jj] + T(2) * Y[i * D + jj]);
                    oQ[i * D + jj] = g * dPp0[sidx(i, jj)];
                }
            // (ii) filter update on (m0, P0) directly
            if (f.obs) {
                T u0[D];
                mv_s<T, D>(f.P, f.h, u0);
                const T s0 = dot<T, D>(f.h, u0) + f.R;
                const T r0 = f.yk - dot<T, D>(f.h, f.m);
                update_adjoint(f.h, f.m, f.P, u0, s0, r0, false, s, s + D, dmp, dPp, acc);
            } else {
#pragma unroll
                for (int e = 0; e < D; ++e) dmp[e] = s[e];
#pragma unroll
                for (int e = 0; e < NS; ++e) dPp[e] = s[D + e];
            }
            // dP0 = F0^T dPp0 F0 + dPp(update)
            if (p.dP0) {
#pragma unroll
                for (int i = 0; i < D; ++i)
#pragma unroll
                    for (int jj = 0; jj < D; ++jj) {
                        T a1 = dPp[sidx(i, jj)];
#pragma unroll
                        for (int kk = 0; kk < D; ++kk) a1 = fma(f.F[kk * D + i], X[kk * D + jj], a1);
                        p.dP0[i * D + jj] = g * a1;
                    }
            }
#pragma unroll
            for (int e = 0; e < D; ++e) s[e] = dmp[e];
#pragma unroll
            for (int e = 0; e < NS; ++e) s[D + e] = dPp[e];
            return;
        }
        if (f.obs) {
            update_adjoint(f.h, f.mp, f.Pp, f.u, f.s, f.r, true, s, s + D, dmp, dPp, acc);
        } else {
#pragma unroll
            for (int e = 0; e < D; ++e) dmp[e] = s[e];
#pragma unroll
            for (int e = 0; e < NS; ++e) dPp[e] = s[D + e];
        }
        T X[D * D], Y[D * D];
        mm_sf<T, D>(dPp, f.F, X);  // dPp F
        mm_fs<T, D>(X, f.P, Y);    // dPp F P
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int jj = 0; jj < D; ++jj) {
                oF[i * D + jj] = g * (dmp[i] * f.m[jj] + T(2) * Y[i * D + jj]);
                oQ[i * D + jj] = g * dPp[sidx(i, jj)];
            }
        mv_t<T, D>(f.F, dmp, s);
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int jj = 0; jj <= i; ++jj) {
                T a1 = T(0);
#pragma unroll
                for (int kk = 0; kk < D; ++kk) a1 = fma(f.F[kk * D + i], X[kk * D + jj], a1);
                s[D + sidx(i, jj)] = a1;
            }
        // when the shard does not start at the global origin, the state after its first step is the
        // adjoint w.r.t. the previous shard's last filtered moments: the framework returns it.
    }

    // visiting row k: take step k+1 (outputs belong to row k+1), then carry (F, Q, y) of row k
    PSSGP_DEV static bool step_row(T* s, const Ctx& cx, const T (&in)[NIN][WMAX], T (&out)[NOUT][WMAX], long k,
                                   const Params& p, T* acc, Carry& c) {
        const bool has = c.has;
        if (has) {
            T m[D], P[NS];
            row_moments(in, m, P);
            step_core(s, cx, c, m, P, out, k + 1, p, acc);
        }
        carry_set(c, in, k, p);
        return has;
    }
    PSSGP_DEV static bool step_flush(T* s, const Ctx& cx, T (&out)[NOUT][WMAX], long k_lo, const Params& p, T* acc,
                                     Carry& c) {
        if (c.has) {
            T m[D], P[NS];
            halo_moments(k_lo, p, m, P);
            step_core(s, cx, c, m, P, out, k_lo, p, acc);
        }
        return c.has;
    }

    PSSGP_DEV static void expand_state(const T* s, T* out) {
#pragma unroll
        for (int e = 0; e < NSTATE; ++e) out[e] = s[e];
    }

    PSSGP_DEV static void finish(const Params& p, int e, T tot, T*) {
        const T g = p.g[0];
        if (e == 0) {
            if (p.dR) p.dR[0] = g * tot;
        } else {
            if (p.dH) p.dH[e - 1] = g * tot;
        }
    }
};

}  // namespace pssgp
