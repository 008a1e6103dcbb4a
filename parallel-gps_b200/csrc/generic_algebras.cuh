// Filter / smoother / adjoint algebras for a generic (run-time) state dimension d: one CTA cooperates
// on one chunk or one aggregate, matrices live in shared memory (coop.cuh).  Same mathematics and the
// same reference correspondences as filter_small.cuh / smoother_small.cuh / adjoint_small.cuh; matrices
// are stored full (not packed).
//
// Every routine must be entered with the CTA synchronised and leaves it synchronised.
#pragma once
#include "coop.cuh"
#include "workspace.h"

namespace pssgp {

template <typename T> struct GScratch {
    int* piv;   // 1 int
    T* red;     // 32 values
};

// ------------------------------------------------------------------------------------------------
// Filter.  aggregate: A[dd] | C[dd] | J[dd] | b[d] | eta[d] ; state: m[d] | P[dd]
// ------------------------------------------------------------------------------------------------
template <typename T>
struct GFilter {
    using scalar = T;
    static constexpr int KIND = KIND_FILTER;
    static constexpr int NACC = 1;
    static const char* name(int i) {
        static const char* n[] = {"gpkf_reduce", "gpkf_up", "gpkf_top", "gpkf_down", "gpkf_apply"};
        return n[i];
    }
    struct Params {
        const T* Fs; const T* Qs; const T* y; const T* H; const T* R; const T* P0; const T* m0;
        T* fms; T* fPs;
        long n; int d; int first_special;
    };
    __host__ __device__ static int nagg(int d) { return 3 * d * d + 2 * d; }
    __host__ __device__ static int nstate(int d) { return d + d * d; }
    __host__ __device__ static int nwork(int d) { return 4 * d * d + (2 * d + 1) * d + 4 * d + 8; }

    __device__ static void identity(const Coop& c, int d, T* a) {
        co_fill(c, nagg(d), a, T(0));
        c.sync();
        for (int i = c.tid; i < d; i += c.nt) a[i * d + i] = T(1);
        c.sync();
    }

    __device__ static void append(const Coop& c, const Params& p, T* a, long k, T* w, const GScratch<T>&) {
        const int d = p.d, dd = d * d;
        T *A = a, *C = a + dd, *J = a + 2 * dd, *b = a + 3 * dd, *eta = b + d;
        T *F = w, *Q = w + dd, *T1 = w + 2 * dd, *T2 = w + 3 * dd, *h = w + 4 * dd, *u = h + d, *ww = u + d, *bp = ww + d;
        for (int i = c.tid; i < d; i += c.nt) h[i] = p.H[i];
        const bool first = (k == 0 && p.first_special);
        if (!first) {
            const T* gf = p.Fs + k * dd;
            const T* gq = p.Qs + k * dd;
            for (int i = c.tid; i < dd; i += c.nt) {
                F[i] = gf[i];
                const int r = i / d, cc = i - r * d;
                Q[i] = T(0.5) * (gq[i] + gq[cc * d + r]);
            }
            c.sync();
            co_mm(c, d, d, d, F, d, 1, A, d, 1, T1, d, (const T*)nullptr, T(1));   // F A
            co_mm(c, d, d, d, F, d, 1, C, d, 1, T2, d, (const T*)nullptr, T(1));   // F C
            co_mv(c, d, d, F, d, 1, b, bp, (const T*)nullptr, T(1));
            c.sync();
            co_copy(c, dd, T1, A);
            co_mm(c, d, d, d, T2, d, 1, F, 1, d, C, d, Q, T(1));                    // F C F^T + Q
            co_copy(c, d, bp, b);
            c.sync();
        } else {
            c.sync();
        }
        const T yk = p.y[k];
        if (t_isnan(yk)) return;
        co_mv(c, d, d, C, d, 1, h, u, (const T*)nullptr, T(1));    // u = C h
        co_mv(c, d, d, A, 1, d, h, ww, (const T*)nullptr, T(1));   // w = A^T h
        c.sync();
        T s = p.R[0], e = yk;
        for (int i = 0; i < d; ++i) {
            s = fma(h[i], u[i], s);
            e = fma(-h[i], b[i], e);
        }
        const T is = T(1) / s;
        c.sync();  // b is about to change; everyone has computed e
        for (int idx = c.tid; idx < dd; idx += c.nt) {
            const int i = idx / d, j = idx - i * d;
            J[idx] = fma(ww[i] * is, ww[j], J[idx]);
            A[idx] = fma(-u[i] * is, ww[j], A[idx]);
            C[idx] = fma(-u[i] * is, u[j], C[idx]);
        }
        for (int i = c.tid; i < d; i += c.nt) {
            eta[i] = fma(ww[i], e * is, eta[i]);
            b[i] = fma(u[i], e * is, b[i]);
        }
        c.sync();
    }

    // out = a1 (earlier) o a2 (later); a1, a2, out distinct shared-memory aggregates.
    __device__ static void combine(const Coop& c, int d, const T* a1, const T* a2, T* out, T* w, const GScratch<T>& g) {
        const int dd = d * d, nr = 2 * d + 1;
        const T *A1 = a1, *C1 = a1 + dd, *J1 = a1 + 2 * dd, *b1 = a1 + 3 * dd, *e1 = b1 + d;
        const T *A2 = a2, *C2 = a2 + dd, *J2 = a2 + 2 * dd, *b2 = a2 + 3 * dd, *e2 = b2 + d;
        T *Ao = out, *Co = out + dd, *Jo = out + 2 * dd, *bo = out + 3 * dd, *eo = bo + d;
        T *M = w, *B = w + dd, *T1 = B + nr * d, *v1 = T1 + dd, *v2 = v1 + d;
        // M = I + C1 J2 ; B = [A1 | b1 + C1 eta2 | C1 A2^T]
        co_eye(c, d, T1);
        c.sync();
        co_mm(c, d, d, d, C1, d, 1, J2, d, 1, M, d, T1, T(1));
        co_mv(c, d, d, C1, d, 1, e2, v1, b1, T(1));
        for (int idx = c.tid; idx < dd; idx += c.nt) {
            const int i = idx / d, j = idx - i * d;
            B[i * nr + j] = A1[idx];
        }
        co_mm(c, d, d, d, C1, d, 1, A2, 1, d, B + d + 1, nr, (const T*)nullptr, T(1));
        c.sync();
        for (int i = c.tid; i < d; i += c.nt) B[i * nr + d] = v1[i];
        c.sync();
        co_solve(c, d, nr, M, B, g.piv);
        // A = A2 ZA ; b = A2 zb + b2 ; C = sym(A2 ZC) + C2
        co_mm(c, d, d, d, A2, d, 1, B, nr, 1, Ao, d, (const T*)nullptr, T(1));
        co_mm(c, d, d, d, A2, d, 1, B + d + 1, nr, 1, Co, d, (const T*)nullptr, T(1));
        for (int i = c.tid; i < d; i += c.nt) {
            T acc = b2[i];
            for (int k = 0; k < d; ++k) acc = fma(A2[i * d + k], B[k * nr + d], acc);
            bo[i] = acc;
        }
        // v2 = eta2 - J2 zb ; T1 = J2 ZA
        for (int i = c.tid; i < d; i += c.nt) {
            T acc = e2[i];
            for (int k = 0; k < d; ++k) acc = fma(-J2[i * d + k], B[k * nr + d], acc);
            v2[i] = acc;
        }
        co_mm(c, d, d, d, J2, d, 1, B, nr, 1, T1, d, (const T*)nullptr, T(1));
        c.sync();
        co_symmetrise(c, d, Co, C2);
        // eta = A1^T v2 + eta1 ; J = sym(A1^T T1) + J1
        co_mv(c, d, d, A1, 1, d, v2, eo, e1, T(1));
        co_mm(c, d, d, d, A1, 1, d, T1, d, 1, Jo, d, (const T*)nullptr, T(1));
        c.sync();
        co_symmetrise(c, d, Jo, J1);
        c.sync();
    }

    // s2 = s o a   (s, s2 distinct)
    __device__ static void apply(const Coop& c, int d, const T* s, const T* a, T* s2, T* w, const GScratch<T>& g) {
        const int dd = d * d, nr = d + 1;
        const T *A = a, *C = a + dd, *J = a + 2 * dd, *b = a + 3 * dd, *eta = b + d;
        const T *m = s, *P = s + d;
        T *M = w, *B = w + dd, *T1 = B + nr * d, *v1 = T1 + dd;
        co_eye(c, d, T1);
        c.sync();
        co_mm(c, d, d, d, P, d, 1, J, d, 1, M, d, T1, T(1));
        co_mv(c, d, d, P, d, 1, eta, v1, m, T(1));
        co_mm(c, d, d, d, P, d, 1, A, 1, d, B + 1, nr, (const T*)nullptr, T(1));
        c.sync();
        for (int i = c.tid; i < d; i += c.nt) B[i * nr] = v1[i];
        c.sync();
        co_solve(c, d, nr, M, B, g.piv);
        for (int i = c.tid; i < d; i += c.nt) {
            T acc = b[i];
            for (int k = 0; k < d; ++k) acc = fma(A[i * d + k], B[k * nr], acc);
            s2[i] = acc;
        }
        co_mm(c, d, d, d, A, d, 1, B + 1, nr, 1, s2 + d, d, (const T*)nullptr, T(1));
        c.sync();
        co_symmetrise(c, d, s2 + d, C);
        c.sync();
    }

    __device__ static void load_init(const Coop& c, const Params& p, T* s) {
        const int d = p.d;
        for (int i = c.tid; i < d; i += c.nt) s[i] = p.m0 ? p.m0[i] : T(0);
        for (int idx = c.tid; idx < d * d; idx += c.nt) {
            const int i = idx / d, j = idx - i * d;
            s[d + idx] = T(0.5) * (p.P0[idx] + p.P0[j * d + i]);
        }
        c.sync();
    }

    // seeded Kalman step; acc[0] accumulates the log-likelihood on thread 0
    __device__ static void step(const Coop& c, const Params& p, T* s, long k, T* w, const GScratch<T>&, T* acc) {
        const int d = p.d, dd = d * d;
        T *m = s, *P = s + d;
        T *F = w, *Q = w + dd, *T1 = w + 2 * dd, *Pp = w + 3 * dd, *h = w + 4 * dd, *u = h + d, *mp = u + d;
        const T* gf = p.Fs + k * dd;
        const T* gq = p.Qs + k * dd;
        for (int i = c.tid; i < dd; i += c.nt) {
            F[i] = gf[i];
            const int r = i / d, cc = i - r * d;
            Q[i] = T(0.5) * (gq[i] + gq[cc * d + r]);
        }
        for (int i = c.tid; i < d; i += c.nt) h[i] = p.H[i];
        c.sync();
        co_mm(c, d, d, d, F, d, 1, P, d, 1, T1, d, (const T*)nullptr, T(1));
        co_mv(c, d, d, F, d, 1, m, mp, (const T*)nullptr, T(1));
        c.sync();
        co_mm(c, d, d, d, T1, d, 1, F, 1, d, Pp, d, Q, T(1));
        c.sync();
        co_mv(c, d, d, Pp, d, 1, h, u, (const T*)nullptr, T(1));
        c.sync();
        const T yk = p.y[k];
        const bool obs = !t_isnan(yk);
        T sv = p.R[0], e = yk;
        for (int i = 0; i < d; ++i) {
            sv = fma(h[i], u[i], sv);
            e = fma(-h[i], mp[i], e);
        }
        if (obs && c.tid == 0) acc[0] += T(-0.5) * (t_log(T(6.283185307179586476925286766559) * sv) + e * e / sv);
        if (k == 0 && p.first_special) {
            c.sync();
            co_copy(c, d, m, mp);
            co_copy(c, dd, P, Pp);
            c.sync();
            co_mv(c, d, d, Pp, d, 1, h, u, (const T*)nullptr, T(1));
            c.sync();
            sv = p.R[0];
            e = yk;
            for (int i = 0; i < d; ++i) {
                sv = fma(h[i], u[i], sv);
                e = fma(-h[i], mp[i], e);
            }
        }
        const T is = T(1) / sv;
        T* om = p.fms + k * d;
        T* oP = p.fPs + k * dd;
        if (obs) {
            for (int i = c.tid; i < d; i += c.nt) {
                const T v = fma(u[i], e * is, mp[i]);
                m[i] = v;
                om[i] = v;
            }
            for (int idx = c.tid; idx < dd; idx += c.nt) {
                const int i = idx / d, j = idx - i * d;
                const T v = fma(-u[i] * is, u[j], Pp[idx]);
                P[idx] = v;
                oP[idx] = v;
            }
        } else {
            for (int i = c.tid; i < d; i += c.nt) {
                m[i] = mp[i];
                om[i] = mp[i];
            }
            for (int idx = c.tid; idx < dd; idx += c.nt) {
                P[idx] = Pp[idx];
                oP[idx] = Pp[idx];
            }
        }
        c.sync();
    }

    // expanded state for the C ABI: m | P full (already the internal layout)
    __device__ static void expand_state(const Coop& c, int d, const T* s, T* out) {
        co_copy(c, nstate(d), s, out);
    }
    __device__ static void finish(const Params&, int, T tot, T* acc_out) {
        if (acc_out) acc_out[0] = tot;
    }
};

// ------------------------------------------------------------------------------------------------
// Smoother (reverse time).  aggregate: E[dd] | L[dd] | g[d] ; state: sm[d] | sP[dd]
// ------------------------------------------------------------------------------------------------
template <typename T>
struct GSmoother {
    using scalar = T;
    static constexpr int KIND = KIND_SMOOTHER;
    static constexpr int NACC = 0;
    static const char* name(int i) {
        static const char* n[] = {"gpks_reduce", "gpks_up", "gpks_top", "gpks_down", "gpks_apply"};
        return n[i];
    }
    struct Params {
        const T* Fs; const T* Qs; const T* fms; const T* fPs; T* sms; T* sPs;
        long n; int d; int last_special;
        const T* Fnext; const T* Qnext; const T* init;
    };
    __host__ __device__ static int nagg(int d) { return 2 * d * d + d; }
    __host__ __device__ static int nstate(int d) { return d + d * d; }
    __host__ __device__ static int nwork(int d) { return 6 * d * d + 4 * d + 8; }

    __device__ static void identity(const Coop& c, int d, T* a) {
        co_fill(c, nagg(d), a, T(0));
        c.sync();
        for (int i = c.tid; i < d; i += c.nt) a[i * d + i] = T(1);
        c.sync();
    }

    // element of time k into (E, L, g) (distinct from everything in w)
    __device__ static void element(const Coop& c, const Params& p, long k, T* E, T* L, T* gvec, T* w,
                                   const GScratch<T>& g) {
        const int d = p.d, dd = d * d;
        T *F = w, *Pp = w + dd, *FP = w + 2 * dd, *P = w + 3 * dd, *m = w + 4 * dd, *Fm = m + d;
        const T* gf = (k + 1 < p.n) ? p.Fs + (k + 1) * dd : p.Fnext;
        const T* gq = (k + 1 < p.n) ? p.Qs + (k + 1) * dd : p.Qnext;
        const T* gP = p.fPs + k * dd;
        for (int i = c.tid; i < dd; i += c.nt) {
            const int r = i / d, cc = i - r * d;
            F[i] = gf[i];
            Pp[i] = T(0.5) * (gq[i] + gq[cc * d + r]);
            P[i] = T(0.5) * (gP[i] + gP[cc * d + r]);
        }
        for (int i = c.tid; i < d; i += c.nt) m[i] = p.fms[k * d + i];
        c.sync();
        co_mm(c, d, d, d, F, d, 1, P, d, 1, FP, d, (const T*)nullptr, T(1));
        co_mv(c, d, d, F, d, 1, m, Fm, (const T*)nullptr, T(1));
        c.sync();
        co_mm(c, d, d, d, FP, d, 1, F, 1, d, Pp, d, Pp, T(1));  // Pp = FP F^T + Q (in place on Q: element-wise safe)
        co_copy(c, dd, FP, L);                                   // L used as the RHS / solution buffer X
        c.sync();
        co_solve(c, d, d, Pp, L, g.piv);                         // L = X = Pp^-1 F P
        for (int idx = c.tid; idx < dd; idx += c.nt) {
            const int i = idx / d, j = idx - i * d;
            E[idx] = L[j * d + i];                               // E = X^T
        }
        c.sync();
        // g = m - E Fm ; L = sym(P - E FP)
        for (int i = c.tid; i < d; i += c.nt) {
            T acc = m[i];
            for (int kk = 0; kk < d; ++kk) acc = fma(-E[i * d + kk], Fm[kk], acc);
            gvec[i] = acc;
        }
        co_mm(c, d, d, d, E, d, 1, FP, d, 1, L, d, (const T*)nullptr, T(-1));
        c.sync();
        co_symmetrise(c, d, L, P);
        c.sync();
    }

    __device__ static void append(const Coop& c, const Params& p, T* a, long j, T* w, const GScratch<T>& g) {
        const int d = p.d, dd = d * d;
        const long k = p.n - 1 - j;
        T *Ea = a, *La = a + dd, *ga = a + 2 * dd;
        if (j == 0 && p.last_special) {
            co_fill(c, dd, Ea, T(0));
            for (int i = c.tid; i < d; i += c.nt) ga[i] = p.fms[k * d + i];
            const T* gP = p.fPs + k * dd;
            for (int idx = c.tid; idx < dd; idx += c.nt) {
                const int r = idx / d, cc = idx - r * d;
                La[idx] = T(0.5) * (gP[idx] + gP[cc * d + r]);
            }
            c.sync();
            return;
        }
        T *E = w + 4 * dd + 2 * d, *L = E + dd, *gv = w + 4 * dd + 3 * d + 2 * dd;
        // note: element() uses w[0 .. 4dd+2d); E, L, gv live behind it
        element(c, p, k, E, L, gv, w, g);
        T *T1 = w, *T2 = w + dd, *t = w + 2 * dd;
        // new = elem o agg : E' = E Ea ; g' = E ga + g ; L' = E La E^T + L
        co_mm(c, d, d, d, E, d, 1, Ea, d, 1, T1, d, (const T*)nullptr, T(1));
        co_mm(c, d, d, d, E, d, 1, La, d, 1, T2, d, (const T*)nullptr, T(1));
        co_mv(c, d, d, E, d, 1, ga, t, gv, T(1));
        c.sync();
        co_copy(c, dd, T1, Ea);
        co_mm(c, d, d, d, T2, d, 1, E, 1, d, La, d, L, T(1));
        co_copy(c, d, t, ga);
        c.sync();
    }

    // a1 first in the reversed sequence (later in time)
    __device__ static void combine(const Coop& c, int d, const T* a1, const T* a2, T* out, T* w, const GScratch<T>&) {
        const int dd = d * d;
        const T *E1 = a1, *L1 = a1 + dd, *g1 = a1 + 2 * dd;
        const T *E2 = a2, *L2 = a2 + dd, *g2 = a2 + 2 * dd;
        T *Eo = out, *Lo = out + dd, *go = out + 2 * dd, *T1 = w;
        co_mm(c, d, d, d, E2, d, 1, E1, d, 1, Eo, d, (const T*)nullptr, T(1));
        co_mm(c, d, d, d, E2, d, 1, L1, d, 1, T1, d, (const T*)nullptr, T(1));
        co_mv(c, d, d, E2, d, 1, g1, go, g2, T(1));
        c.sync();
        co_mm(c, d, d, d, T1, d, 1, E2, 1, d, Lo, d, L2, T(1));
        c.sync();
    }

    __device__ static void apply(const Coop& c, int d, const T* s, const T* a, T* s2, T* w, const GScratch<T>&) {
        const int dd = d * d;
        const T *E = a, *L = a + dd, *gv = a + 2 * dd;
        T* T1 = w;
        co_mv(c, d, d, E, d, 1, s, s2, gv, T(1));
        co_mm(c, d, d, d, E, d, 1, s + d, d, 1, T1, d, (const T*)nullptr, T(1));
        c.sync();
        co_mm(c, d, d, d, T1, d, 1, E, 1, d, s2 + d, d, L, T(1));
        c.sync();
    }

    // packed (C-ABI) <-> full state conversion
    __device__ static void load_init(const Coop& c, const Params& p, T* s) {
        const int d = p.d;
        if (!p.init) {
            co_fill(c, nstate(d), s, T(0));
        } else {
            for (int i = c.tid; i < d; i += c.nt) s[i] = p.init[i];
            for (int idx = c.tid; idx < d * d; idx += c.nt) {
                const int i = idx / d, j = idx - i * d;
                s[d + idx] = p.init[d + sidx(i, j)];
            }
        }
        c.sync();
    }
    __device__ static void expand_state(const Coop& c, int d, const T* s, T* out) {
        for (int i = c.tid; i < d; i += c.nt) out[i] = s[i];
        for (int idx = c.tid; idx < d * d; idx += c.nt) {
            const int i = idx / d, j = idx - i * d;
            if (j <= i) out[d + sidx(i, j)] = s[d + idx];
        }
    }

    __device__ static void step(const Coop& c, const Params& p, T* s, long j, T* w, const GScratch<T>& g, T*) {
        const int d = p.d, dd = d * d;
        const long k = p.n - 1 - j;
        if (j == 0 && p.last_special) {
            for (int i = c.tid; i < d; i += c.nt) s[i] = p.fms[k * d + i];
            const T* gP = p.fPs + k * dd;
            for (int idx = c.tid; idx < dd; idx += c.nt) {
                const int r = idx / d, cc = idx - r * d;
                s[d + idx] = T(0.5) * (gP[idx] + gP[cc * d + r]);
            }
            c.sync();
        } else {
            T *E = w + 4 * dd + 2 * d, *L = E + dd, *gv = w + 4 * dd + 3 * d + 2 * dd;
            element(c, p, k, E, L, gv, w, g);
            T *T1 = w, *sn = w + dd;  // sn: d + dd
            co_mv(c, d, d, E, d, 1, s, sn, gv, T(1));
            co_mm(c, d, d, d, E, d, 1, s + d, d, 1, T1, d, (const T*)nullptr, T(1));
            c.sync();
            co_mm(c, d, d, d, T1, d, 1, E, 1, d, sn + d, d, L, T(1));
            c.sync();
            co_copy(c, d + dd, sn, s);
            c.sync();
        }
        co_copy(c, d, s, p.sms + k * d);
        co_copy(c, dd, s + d, p.sPs + k * dd);
        c.sync();
    }
    __device__ static void finish(const Params&, int, T, T*) {}
};

// ------------------------------------------------------------------------------------------------
// Adjoint of the log-likelihood (reverse time).  aggregate: Abar[dd] | B[dd] | a[d] ; state: dm[d] | dP[dd]
// accumulators: dR, dH[d]  (NACC = 1 + d, run-time)
// ------------------------------------------------------------------------------------------------
template <typename T>
struct GAdjoint {
    using scalar = T;
    static constexpr int KIND = KIND_ADJOINT;
    static constexpr int NACC = -1;  // run-time: 1 + d
    static const char* name(int i) {
        static const char* n[] = {"gbwd_reduce", "gbwd_up", "gbwd_top", "gbwd_down", "gbwd_apply"};
        return n[i];
    }
    struct Params {
        const T* Fs; const T* Qs; const T* y; const T* H; const T* R; const T* P0; const T* m0;
        const T* fms; const T* fPs; const T* g; const T* init;
        T* dFs; T* dQs; T* dP0; T* dH; T* dR;
        long n; int d; int first_special;
    };
    __host__ __device__ static int nagg(int d) { return 2 * d * d + d; }
    __host__ __device__ static int nstate(int d) { return d + d * d; }
    __host__ __device__ static int nwork(int d) { return 8 * d * d + 10 * d + 16; }

    __device__ static void identity(const Coop& c, int d, T* a) {
        co_fill(c, nagg(d), a, T(0));
        c.sync();
        for (int i = c.tid; i < d; i += c.nt) a[i * d + i] = T(1);
        c.sync();
    }

    // forward quantities of step k into w: F | P | Pp | T1 | h | m | mp | u ; returns s, r through sh scalars
    struct Fw { T *F, *P, *Pp, *T1, *h, *m, *mp, *u; T s, r, yk; bool obs, first; };

    __device__ static Fw forward(const Coop& c, const Params& p, long k, T* w) {
        const int d = p.d, dd = d * d;
        Fw f;
        f.F = w; f.P = w + dd; f.Pp = w + 2 * dd; f.T1 = w + 3 * dd;
        f.h = w + 4 * dd; f.m = f.h + d; f.mp = f.m + d; f.u = f.mp + d;
        f.yk = p.y[k];
        f.obs = !t_isnan(f.yk);
        f.first = (k == 0 && p.first_special);
        const T* gf = p.Fs + k * dd;
        const T* gq = p.Qs + k * dd;
        const T* gP = (k > 0) ? p.fPs + (k - 1) * dd : p.P0;
        for (int i = c.tid; i < dd; i += c.nt) {
            const int r = i / d, cc = i - r * d;
            f.F[i] = gf[i];
            f.Pp[i] = T(0.5) * (gq[i] + gq[cc * d + r]);
            f.P[i] = T(0.5) * (gP[i] + gP[cc * d + r]);
        }
        for (int i = c.tid; i < d; i += c.nt) {
            f.h[i] = p.H[i];
            f.m[i] = (k > 0) ? p.fms[(k - 1) * d + i] : (p.m0 ? p.m0[i] : T(0));
        }
        c.sync();
        co_mm(c, d, d, d, f.F, d, 1, f.P, d, 1, f.T1, d, (const T*)nullptr, T(1));
        co_mv(c, d, d, f.F, d, 1, f.m, f.mp, (const T*)nullptr, T(1));
        c.sync();
        co_mm(c, d, d, d, f.T1, d, 1, f.F, 1, d, f.Pp, d, f.Pp, T(1));
        c.sync();
        co_mv(c, d, d, f.Pp, d, 1, f.h, f.u, (const T*)nullptr, T(1));
        c.sync();
        T s = p.R[0], r = f.yk;
        for (int i = 0; i < d; ++i) {
            s = fma(f.h[i], f.u[i], s);
            r = fma(-f.h[i], f.mp[i], r);
        }
        f.s = s;
        f.r = r;
        return f;
    }

    // element of time k into x = (Abar | B | a); uses w[0 .. 4dd + 5d)
    __device__ static void element(const Coop& c, const Params& p, long k, T* x, T* w) {
        const int d = p.d, dd = d * d;
        Fw f = forward(c, p, k, w);
        T *Ab = x, *B = x + dd, *a = x + 2 * dd, *wv = f.u + d;
        if (f.first) {
            // update on (m0, P0) directly: Abar = I - K0 h^T, a = 0, B = 0
            co_mv(c, d, d, f.P, d, 1, f.h, f.u, (const T*)nullptr, T(1));
            c.sync();
            T s0 = p.R[0];
            for (int i = 0; i < d; ++i) s0 = fma(f.h[i], f.u[i], s0);
            const T is = f.obs ? T(1) / s0 : T(0);
            for (int idx = c.tid; idx < dd; idx += c.nt) {
                const int i = idx / d, j = idx - i * d;
                Ab[idx] = ((i == j) ? T(1) : T(0)) - f.u[i] * is * f.h[j];
                B[idx] = T(0);
            }
            co_fill(c, d, a, T(0));
            c.sync();
            return;
        }
        if (!f.obs) {
            co_copy(c, dd, f.F, Ab);
            co_fill(c, dd, B, T(0));
            co_fill(c, d, a, T(0));
            c.sync();
            return;
        }
        co_mv(c, d, d, f.F, 1, d, f.h, wv, (const T*)nullptr, T(1));  // w = F^T h
        c.sync();
        const T is = T(1) / f.s, ris = f.r * is;
        for (int idx = c.tid; idx < dd; idx += c.nt) {
            const int i = idx / d, j = idx - i * d;
            Ab[idx] = fma(-f.u[i] * is, wv[j], f.F[idx]);
            B[idx] = T(0.5) * (ris * ris - is) * wv[i] * wv[j];
        }
        for (int i = c.tid; i < d; i += c.nt) a[i] = wv[i] * ris;
        c.sync();
    }

    // x1 later in time
    __device__ static void combine(const Coop& c, int d, const T* x1, const T* x2, T* out, T* w, const GScratch<T>&) {
        const int dd = d * d;
        const T *A1 = x1, *B1 = x1 + dd, *a1 = x1 + 2 * dd;
        const T *A2 = x2, *B2 = x2 + dd, *a2 = x2 + 2 * dd;
        T *Ao = out, *Bo = out + dd, *ao = out + 2 * dd, *T1 = w, *t = w + dd;
        co_mm(c, d, d, d, A1, d, 1, A2, d, 1, Ao, d, (const T*)nullptr, T(1));
        co_mm(c, d, d, d, B1, d, 1, A2, d, 1, T1, d, (const T*)nullptr, T(1));
        co_mv(c, d, d, A2, 1, d, a1, t, (const T*)nullptr, T(1));
        c.sync();
        for (int idx = c.tid; idx < dd; idx += c.nt) {
            const int i = idx / d, j = idx - i * d;
            T acc = B2[idx] + T(0.5) * (t[i] * a2[j] + t[j] * a2[i]);
            for (int k = 0; k < d; ++k) acc = fma(A2[k * d + i], T1[k * d + j], acc);
            Bo[idx] = acc;
        }
        for (int i = c.tid; i < d; i += c.nt) ao[i] = t[i] + a2[i];
        c.sync();
    }

    __device__ static void append(const Coop& c, const Params& p, T* a, long j, T* w, const GScratch<T>& g) {
        const int d = p.d, dd = d * d;
        T* x = w + 4 * dd + 6 * d;        // element buffer (2dd + d)
        T* out = x + 2 * dd + d;           // combine output (2dd + d)
        element(c, p, p.n - 1 - j, x, w);
        combine(c, d, a, x, out, w, g);    // combine's work (dd + d) reuses the front of w
        co_copy(c, nagg(d), out, a);
        c.sync();
    }

    __device__ static void apply(const Coop& c, int d, const T* s, const T* x, T* s2, T* w, const GScratch<T>&) {
        const int dd = d * d;
        const T *Ab = x, *B = x + dd, *a = x + 2 * dd;
        T *T1 = w, *t = w + dd;
        co_mv(c, d, d, Ab, 1, d, s, t, (const T*)nullptr, T(1));
        co_mm(c, d, d, d, s + d, d, 1, Ab, d, 1, T1, d, (const T*)nullptr, T(1));
        c.sync();
        for (int idx = c.tid; idx < dd; idx += c.nt) {
            const int i = idx / d, j = idx - i * d;
            T acc = B[idx] + T(0.5) * (t[i] * a[j] + t[j] * a[i]);
            for (int k = 0; k < d; ++k) acc = fma(Ab[k * d + i], T1[k * d + j], acc);
            s2[d + idx] = acc;
        }
        for (int i = c.tid; i < d; i += c.nt) s2[i] = t[i] + a[i];
        c.sync();
    }

    __device__ static void load_init(const Coop& c, const Params& p, T* s) {
        const int d = p.d;
        if (!p.init) {
            co_fill(c, nstate(d), s, T(0));
        } else {
            for (int i = c.tid; i < d; i += c.nt) s[i] = p.init[i];
            for (int idx = c.tid; idx < d * d; idx += c.nt) {
                const int i = idx / d, j = idx - i * d;
                s[d + idx] = p.init[d + sidx(i, j)];
            }
        }
        c.sync();
    }
    __device__ static void expand_state(const Coop& c, int d, const T* s, T* out) {
        for (int i = c.tid; i < d; i += c.nt) out[i] = s[i];
        for (int idx = c.tid; idx < d * d; idx += c.nt) {
            const int i = idx / d, j = idx - i * d;
            if (j <= i) out[d + sidx(i, j)] = s[d + idx];
        }
    }

    // adjoint of the measurement update; vectors in shared memory: dm, dP (in), dmp, dPp (out); acc on thread 0
    __device__ static void update_adjoint(const Coop& c, int d, const T* h, const T* mp, const T* Pp, const T* u, T s,
                                          T r, bool with_ll, const T* dm, const T* dP, T* dmp, T* dPp, T* Pu, T* ut,
                                          T* Pput, T* acc) {
        const int dd = d * d;
        const T is = T(1) / s;
        co_mv(c, d, d, dP, d, 1, u, Pu, (const T*)nullptr, T(1));
        c.sync();
        T udm = T(0), uPu = T(0);
        for (int i = 0; i < d; ++i) {
            udm = fma(u[i], dm[i], udm);
            uPu = fma(u[i], Pu[i], uPu);
        }
        T rbar = udm * is;
        T sbar = (-udm * r + uPu) * is * is;
        if (with_ll) {
            rbar -= r * is;
            sbar += T(0.5) * (r * r * is * is - is);
        }
        for (int i = c.tid; i < d; i += c.nt) ut[i] = dm[i] * r * is - T(2) * Pu[i] * is + sbar * h[i];
        c.sync();
        co_mv(c, d, d, Pp, d, 1, ut, Pput, (const T*)nullptr, T(1));
        c.sync();
        if (c.tid == 0) {
            acc[0] += sbar;
            for (int i = 0; i < d; ++i) acc[1 + i] += sbar * u[i] + Pput[i] - mp[i] * rbar;
        }
        for (int i = c.tid; i < d; i += c.nt) dmp[i] = dm[i] - h[i] * rbar;
        for (int idx = c.tid; idx < dd; idx += c.nt) {
            const int i = idx / d, j = idx - i * d;
            dPp[idx] = dP[idx] + T(0.5) * (ut[i] * h[j] + ut[j] * h[i]);
        }
        c.sync();
    }

    __device__ static void step(const Coop& c, const Params& p, T* s, long j, T* w, const GScratch<T>&, T* acc) {
        const int d = p.d, dd = d * d;
        const long k = p.n - 1 - j;
        const T gl = p.g[0];
        Fw f = forward(c, p, k, w);
        T* base = w + 4 * dd + 4 * d;
        T *dmp = base, *Pu = dmp + d, *ut = Pu + d, *Pput = ut + d, *dmp0 = Pput + d;
        T *dPp = dmp0 + d, *X = dPp + dd, *Y = X + dd, *dPp0 = Y + dd;
        T* oF = p.dFs + k * dd;
        T* oQ = p.dQs + k * dd;
        if (f.first) {
            T sbar = T(0), rbar = T(0);
            if (f.obs) {
                const T is = T(1) / f.s;
                sbar = T(0.5) * (f.r * f.r * is * is - is);
                rbar = -f.r * is;
                if (c.tid == 0) {
                    acc[0] += sbar;
                    for (int i = 0; i < d; ++i) acc[1 + i] += T(2) * sbar * f.u[i] - f.mp[i] * rbar;
                }
            }
            for (int i = c.tid; i < d; i += c.nt) dmp0[i] = -f.h[i] * rbar;
            for (int idx = c.tid; idx < dd; idx += c.nt) {
                const int i = idx / d, jj = idx - i * d;
                dPp0[idx] = sbar * f.h[i] * f.h[jj];
            }
            c.sync();
            co_mm(c, d, d, d, dPp0, d, 1, f.F, d, 1, X, d, (const T*)nullptr, T(1));
            c.sync();
            co_mm(c, d, d, d, X, d, 1, f.P, d, 1, Y, d, (const T*)nullptr, T(1));
            c.sync();
            for (int idx = c.tid; idx < dd; idx += c.nt) {
                const int i = idx / d, jj = idx - i * d;
                oF[idx] = gl * (dmp0[i] * f.m[jj] + T(2) * Y[idx]);
                oQ[idx] = gl * dPp0[idx];
            }
            if (f.obs) {
                co_mv(c, d, d, f.P, d, 1, f.h, f.u, (const T*)nullptr, T(1));
                c.sync();
                T s0 = p.R[0], r0 = f.yk;
                for (int i = 0; i < d; ++i) {
                    s0 = fma(f.h[i], f.u[i], s0);
                    r0 = fma(-f.h[i], f.m[i], r0);
                }
                update_adjoint(c, d, f.h, f.m, f.P, f.u, s0, r0, false, s, s + d, dmp, dPp, Pu, ut, Pput, acc);
            } else {
                co_copy(c, d, s, dmp);
                co_copy(c, dd, s + d, dPp);
                c.sync();
            }
            if (p.dP0) {
                // dP0 = F0^T dPp0 F0 + dPp(update) = F^T X + dPp
                for (int idx = c.tid; idx < dd; idx += c.nt) {
                    const int i = idx / d, jj = idx - i * d;
                    T a1 = dPp[idx];
                    for (int kk = 0; kk < d; ++kk) a1 = fma(f.F[kk * d + i], X[kk * d + jj], a1);
                    p.dP0[idx] = gl * a1;
                }
            }
            co_copy(c, d, dmp, s);
            co_copy(c, dd, dPp, s + d);
            c.sync();
            return;
        }
        if (f.obs) {
            update_adjoint(c, d, f.h, f.mp, f.Pp, f.u, f.s, f.r, true, s, s + d, dmp, dPp, Pu, ut, Pput, acc);
        } else {
            co_copy(c, d, s, dmp);
            co_copy(c, dd, s + d, dPp);
            c.sync();
        }
        co_mm(c, d, d, d, dPp, d, 1, f.F, d, 1, X, d, (const T*)nullptr, T(1));
        c.sync();
        co_mm(c, d, d, d, X, d, 1, f.P, d, 1, Y, d, (const T*)nullptr, T(1));
        co_mv(c, d, d, f.F, 1, d, dmp, s, (const T*)nullptr, T(1));
        co_mm(c, d, d, d, f.F, 1, d, X, d, 1, s + d, d, (const T*)nullptr, T(1));
        c.sync();
        for (int idx = c.tid; idx < dd; idx += c.nt) {
            const int i = idx / d, jj = idx - i * d;
            oF[idx] = gl * (dmp[i] * f.m[jj] + T(2) * Y[idx]);
            oQ[idx] = gl * dPp[idx];
        }
        c.sync();
    }

    __device__ static void finish(const Params& p, int e, T tot, T*) {
        const T gl = p.g[0];
        if (e == 0) {
            if (p.dR) p.dR[0] = gl * tot;
        } else if (p.dH) {
            p.dH[e - 1] = gl * tot;
        }
    }
};

// ------------------------------------------------------------------------------------------------
// Combined reverse scan of the fused d > 4 step (mid.cuh): the log-likelihood adjoint (dm, dP) and the
// solve-free (modified Bryson-Frazier) form of the RTS smoother (lam, Lam) propagate backwards through the SAME
// transition matrices Abar_k = (I - K_k H) F_k:
//     dm'  = Abar^T dm  + a          dP'  = Abar^T dP  Abar + sym((Abar^T dm) a^T) + Ba
//     lam' = Abar^T lam - a          Lam' = Abar^T Lam Abar + Bm
// with a = F^T H^T r/S, Ba = (r^2/S^2 - 1/S)/2 * F^T H^T H F, Bm = F^T H^T H F / S per observed step (a missing
// observation: Abar = F, a = Ba = Bm = 0).  Smoothed moments: sm_k = m_k - P_k lam_k, sP_k = P_k - P_k Lam_k P_k
// (equal to pssgp/kalman/parallel.py:155-196 in exact arithmetic; no d x d solve per step).
// aggregate: Abar[dd] | Ba[dd] | Bm[dd] | a[d] ; state: dm[d] | lam[d] | dP[dd] | Lam[dd]
// Only the hierarchy over chunk aggregates runs through this CTA-cooperative algebra; the per-step work is in the
// warp-level kernels of mid.cuh.
// ------------------------------------------------------------------------------------------------
template <typename T>
struct GRev {
    using scalar = T;
    static constexpr int KIND = KIND_ADJOINT;
    static constexpr int NACC = -1;  // run-time: 1 + d
    static const char* name(int i) {
        static const char* n[] = {"mrev_reduce", "mrev_up", "mrev_top", "mrev_down", "mrev_apply"};
        return n[i];
    }
    struct Params {
        const T* Fs; const T* Qs; const T* y; const T* H; const T* R; const T* P0; const T* m0;
        const T* fms; const T* fPs; const T* g; const T* init;
        T* sms; T* sPs; T* dFs; T* dQs; T* dP0; T* dH; T* dR;
        long n; int d; int first_special;
    };
    __host__ __device__ static int nagg(int d) { return 3 * d * d + d; }
    __host__ __device__ static int nstate(int d) { return 2 * d * d + 2 * d; }
    __host__ __device__ static int nwork(int d) { return 2 * d * d + 2 * d + 16; }

    __device__ static void identity(const Coop& c, int d, T* a) {
        co_fill(c, nagg(d), a, T(0));
        c.sync();
        for (int i = c.tid; i < d; i += c.nt) a[i * d + i] = T(1);
        c.sync();
    }

    // x1 later in time (first in scan order), x2 earlier: out = "x1, then x2"
    __device__ static void combine(const Coop& c, int d, const T* x1, const T* x2, T* out, T* w, const GScratch<T>&) {
        const int dd = d * d;
        const T *A1 = x1, *Ba1 = x1 + dd, *Bm1 = x1 + 2 * dd, *a1 = x1 + 3 * dd;
        const T *A2 = x2, *Ba2 = x2 + dd, *Bm2 = x2 + 2 * dd, *a2 = x2 + 3 * dd;
        T *Ao = out, *Bao = out + dd, *Bmo = out + 2 * dd, *ao = out + 3 * dd, *T1 = w, *T2 = w + dd, *t = w + 2 * dd;
        co_mm(c, d, d, d, A1, d, 1, A2, d, 1, Ao, d, (const T*)nullptr, T(1));
        co_mm(c, d, d, d, Ba1, d, 1, A2, d, 1, T1, d, (const T*)nullptr, T(1));
        co_mm(c, d, d, d, Bm1, d, 1, A2, d, 1, T2, d, (const T*)nullptr, T(1));
        co_mv(c, d, d, A2, 1, d, a1, t, (const T*)nullptr, T(1));
        c.sync();
        for (int idx = c.tid; idx < dd; idx += c.nt) {
            const int i = idx / d, j = idx - i * d;
            T acc = Ba2[idx] + T(0.5) * (t[i] * a2[j] + t[j] * a2[i]);
            T acm = Bm2[idx];
            for (int k = 0; k < d; ++k) {
                acc = fma(A2[k * d + i], T1[k * d + j], acc);
                acm = fma(A2[k * d + i], T2[k * d + j], acm);
            }
            Bao[idx] = acc;
            Bmo[idx] = acm;
        }
        for (int i = c.tid; i < d; i += c.nt) ao[i] = t[i] + a2[i];
        c.sync();
    }

    __device__ static void apply(const Coop& c, int d, const T* s, const T* x, T* s2, T* w, const GScratch<T>&) {
        const int dd = d * d;
        const T *Ab = x, *Ba = x + dd, *Bm = x + 2 * dd, *a = x + 3 * dd;
        const T *dm = s, *lam = s + d, *dP = s + 2 * d, *Lam = s + 2 * d + dd;
        T *T1 = w, *T2 = w + dd, *t = w + 2 * dd, *tl = t + d;
        co_mv(c, d, d, Ab, 1, d, dm, t, (const T*)nullptr, T(1));
        co_mv(c, d, d, Ab, 1, d, lam, tl, (const T*)nullptr, T(1));
        co_mm(c, d, d, d, dP, d, 1, Ab, d, 1, T1, d, (const T*)nullptr, T(1));
        co_mm(c, d, d, d, Lam, d, 1, Ab, d, 1, T2, d, (const T*)nullptr, T(1));
        c.sync();
        for (int idx = c.tid; idx < dd; idx += c.nt) {
            const int i = idx / d, j = idx - i * d;
            T acc = Ba[idx] + T(0.5) * (t[i] * a[j] + t[j] * a[i]);
            T acm = Bm[idx];
            for (int k = 0; k < d; ++k) {
                acc = fma(Ab[k * d + i], T1[k * d + j], acc);
                acm = fma(Ab[k * d + i], T2[k * d + j], acm);
            }
            s2[2 * d + idx] = acc;
            s2[2 * d + dd + idx] = acm;
        }
        for (int i = c.tid; i < d; i += c.nt) {
            s2[i] = t[i] + a[i];
            s2[d + i] = tl[i] - a[i];
        }
        c.sync();
    }

    __device__ static void load_init(const Coop& c, const Params& p, T* s) {
        const int d = p.d;
        if (!p.init) co_fill(c, nstate(d), s, T(0));
        else co_copy(c, nstate(d), p.init, s);
        c.sync();
    }
    __device__ static void expand_state(const Coop& c, int d, const T* s, T* out) { co_copy(c, nstate(d), s, out); }
    __device__ static void finish(const Params& p, int e, T tot, T*) {
        const T gl = p.g ? p.g[0] : T(1);
        if (e == 0) {
            if (p.dR) p.dR[0] = gl * tot;
        } else if (p.dH) {
            p.dH[e - 1] = gl * tot;
        }
    }
};

}  // namespace pssgp
