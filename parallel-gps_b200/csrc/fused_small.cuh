// Fused filter + smoother + log-likelihood-gradient step for D <= 4: algebras that let the three scans
// of pkf / pks / pkf_backward share their passes over the LGSSM.
//
//   K1  stream_reduce<FilterAlg>          read F, Q, y                  (chunk aggregates of the filter)
//   K2' stream_apply<FusedFwdAlg>         read F, Q, y; write fms, fPs  + while the filtered moments of a
//                                         chunk are in registers, the chunk aggregates of BOTH reverse scans
//                                         (smoother (E, g, L), adjoint (Abar, a, B)), built in ascending time
//                                         by composing every new element on the "later" side; the CTA-level
//                                         reverse scans and the scan over CTA totals follow in the same kernel
//   K3' stream_apply<FusedRevAlg>         read F, Q, y, fms, fPs; write sms, sPs, dFs, dQs
//
// so the reduce passes of the smoother and of the adjoint (and their re-reads of F, Q, y, fms, fPs) disappear
// and the two reverse apply passes share one read of their inputs.  Same arithmetic per element as
// smoother_small.cuh / adjoint_small.cuh (reference: pssgp/kalman/parallel.py:155-184 for the smoother,
// TF autodiff of :121-152 for the gradient); only the association order of the chunk aggregates differs.
// pssgp_pkfs_grad runs the three kernels on a whole series; pssgp_pkf_with_summaries runs K2' on one shard of a
// time-sharded series and leaves the reverse aggregates for pssgp_pks / pssgp_pkf_backward.
#pragma once
#include "adjoint_small.cuh"
#include "filter_small.cuh"
#include "scan_stream.cuh"
#include "smoother_small.cuh"

namespace pssgp {

// Workspace of one reverse scan as the forward kernel fills it (same layout as run_scan's).
template <typename T>
struct SideWs {
    T* lane_excl;
    T* warp_excl;
    T* wagg;
    T* wstate;
};

// SM / AD: build the aggregates of the smoother / of the adjoint scan (both for the full step, SM only for pkfs,
// AD only for a training step without smoother).
template <typename T, int D, bool SM = true, bool AD = true>
struct FusedFwdAlg : FilterAlg<T, D> {
    using Base = FilterAlg<T, D>;
    using SA = SmootherAlg<T, D>;
    using AA = AdjointAlg<T, D>;
    static const char* name_apply() { return "pkf_apply_fused"; }
    static constexpr bool HAS_SIDE = true;
    static constexpr int NS = Base::NS, NIN = Base::NIN, NOUT = Base::NOUT, WMAX = Base::WMAX;
    using Ctx = typename Base::Ctx;

    struct Params : Base::Params {
        long n;
        SideWs<T> sm, ad;
        unsigned int* side_ticket;
        // time sharding: the shard's last step is not the global last one: F, Q of the step after it (halo)
        int last_special;
        const T* Fnext;
        const T* Qnext;
        // non-null: the last CTA writes the shard summaries of both reverse scans and the exclusive prefix aggregate
        // of every CTA (into sm.wstate / ad.wstate) instead of CTA states (the states entering this shard from the
        // right are not known yet: pssgp_pks / pssgp_pkf_backward finish with state o prefix)
        T* sm_summary;
        T* ad_summary;
        // non-null (with Base::fold): receives the folded state entering the shard, m [D] | P [D, D], for the calls
        // that need it afterwards (pssgp_pkf_backward's P0 / m0)
        T* state_in_out = nullptr;
    };
    // called by every thread of the apply kernel with the state entering the shard (time sharding, prefix mode)
    PSSGP_DEV static void publish_init(const Params& p, const T* s) {
        if (p.state_in_out != nullptr && blockIdx.x == 0 && threadIdx.x == 0) Base::expand_state(s, p.state_in_out);
    }

    struct Carry : Base::Carry {
        T sa[SA::NAGG];   // e_{k_lo} o ... (smoother elements of this chunk appended so far)
        T aa[AA::NAGG];   // adjoint elements of this chunk appended so far
        long k_lo;
    };
    PSSGP_DEV static void carry_init(Carry& c, const Ctx& cx, long k_lo, long k_hi, const Params& p) {
        Base::carry_init(c, cx, k_lo, k_hi, p);
        SA::identity(c.sa);
        AA::identity(c.aa);
        c.k_lo = k_lo;
    }
    PSSGP_DEV static void step_done(T* acc, const Carry& c) { Base::step_done(acc, c); }

    // sa <- sa o e  with e = (E, g, L) later in time than everything in sa
    PSSGP_DEV static void push_smoother(Carry& c, const T* el) {
        T r[SA::NAGG];
        SA::combine(el, c.sa, r);
#pragma unroll
        for (int e = 0; e < SA::NAGG; ++e) c.sa[e] = r[e];
    }
    // smoothing element of time k-1 from FP = F_k P_{k-1}, Pp = F_k P_{k-1} F_k^T + Q_k, mp = F_k m_{k-1}
    PSSGP_DEV static void smoother_element(const T* FP, const T* Pp, const T* m, const T* P, const T* mp, T* el) {
        T S[NS], X[D * D];
#pragma unroll
        for (int e = 0; e < NS; ++e) S[e] = Pp[e];
        ldl_packed<T, D>(S);
#pragma unroll
        for (int e = 0; e < D * D; ++e) X[e] = FP[e];
        ldl_solve<T, D, D>(S, X);  // X = Pp^{-1} F P ; E = X^T
        T* E = el + SA::oE;
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j < D; ++j) E[i * D + j] = X[j * D + i];
        T Emp[D];
        mv_f<T, D>(E, mp, Emp);
#pragma unroll
        for (int i = 0; i < D; ++i) el[SA::og + i] = m[i] - Emp[i];
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
                el[SA::oL + sidx(i, j)] = P[sidx(i, j)] - T(0.5) * (a1 + a2);
            }
    }

    // Seeded Kalman step k (FilterAlg::step_row) + smoothing element of time k-1 + adjoint element of step k.
    PSSGP_DEV static bool step_row(T* s, const Ctx& cx, const T (&in)[NIN][WMAX], T (&out)[NOUT][WMAX], long k,
                                   const Params& p, T* acc, Carry& cr) {
        const T* F = in[0];
        const T* h = cx.h;
        const T R = cx.R;
        const T yk = in[2][0];
        T Q[NS], FP[D * D], mp[D], Pp[NS];
        Base::sym_pack(in[1], Q);
        mv_f<T, D>(F, s, mp);
        mm_fs<T, D>(F, s + D, FP);
        sym_xat_plus<T, D>(FP, F, Q, Pp);
        const bool obs = !t_isnan(yk);
        const bool first = (k == 0 && p.first_special);
        T u[D];
        mv_s<T, D>(Pp, h, u);
        T sv = dot<T, D>(h, u) + R;
        T e0 = yk - dot<T, D>(h, mp);
        T is = t_rcp(sv);
        if (obs) {
            cr.ls.add(sv);
            cr.quad = fma(e0 * e0, is, cr.quad);
            cr.nobs += 1;
        }
        // smoothing element of time k-1 (parallel.py:159-166); the one before the chunk's first row belongs
        // to the previous chunk (its owner builds it from a halo row in side_flush)
        if (SM && k > cr.k_lo) {
            T el[SA::NAGG];
            smoother_element(FP, Pp, s, s + D, mp, el);
            push_smoother(cr, el);
        }
        if (first) {
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
        // adjoint element of step k (adjoint_small.cuh element()): x = (Abar, a, B), appended on the later side
        if constexpr (AD) {
            T x[AA::NAGG];
            if (first) {
                AA::identity(x);
                if (obs) {
#pragma unroll
                    for (int i = 0; i < D; ++i)
#pragma unroll
                        for (int j = 0; j < D; ++j) x[AA::oA + i * D + j] -= u[i] * is * h[j];
                }
            } else if (!obs) {
#pragma unroll
                for (int e = 0; e < D * D; ++e) x[AA::oA + e] = F[e];
#pragma unroll
                for (int e = 0; e < D; ++e) x[AA::oa + e] = T(0);
#pragma unroll
                for (int e = 0; e < NS; ++e) x[AA::oB + e] = T(0);
            } else {
                T w[D];  // F^T h
                mv_t<T, D>(F, h, w);
#pragma unroll
                for (int i = 0; i < D; ++i)
#pragma unroll
                    for (int j = 0; j < D; ++j) x[AA::oA + i * D + j] = fma(-u[i] * is, w[j], F[i * D + j]);
                const T ris = e0 * is;
#pragma unroll
                for (int i = 0; i < D; ++i) x[AA::oa + i] = w[i] * ris;
#pragma unroll
                for (int i = 0; i < D; ++i)
#pragma unroll
                    for (int j = 0; j <= i; ++j) x[AA::oB + sidx(i, j)] = T(0.5) * (ris * ris - is) * w[i] * w[j];
            }
            T r[AA::NAGG];
            AA::combine(x, cr.aa, r);
#pragma unroll
            for (int e = 0; e < AA::NAGG; ++e) cr.aa[e] = r[e];
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

    // After the chunk's last row (s = filtered moments at k_hi - 1): the smoothing element of time k_hi - 1,
    // from the halo row k_hi, or last_smoothing_element (parallel.py:155-156) at the end of the series.
    PSSGP_DEV static void side_flush(const T* s, long k_hi, const Params& p, Carry& cr) {
        if constexpr (!SM) return;
        T el[SA::NAGG];
        if (k_hi >= p.n && p.last_special) {
#pragma unroll
            for (int e = 0; e < D * D; ++e) el[SA::oE + e] = T(0);
#pragma unroll
            for (int i = 0; i < D; ++i) el[SA::og + i] = s[i];
#pragma unroll
            for (int e = 0; e < NS; ++e) el[SA::oL + e] = s[D + e];
        } else {
            T F[D * D], qf[D * D], Q[NS], FP[D * D], mp[D], Pp[NS];
            const T* pf = k_hi >= p.n ? p.Fnext : p.Fs + k_hi * (D * D);
            const T* pq = k_hi >= p.n ? p.Qnext : p.Qs + k_hi * (D * D);
#pragma unroll
            for (int e = 0; e < D * D; ++e) {
                F[e] = __ldg(pf + e);
                qf[e] = __ldg(pq + e);
            }
            Base::sym_pack(qf, Q);
            mv_f<T, D>(F, s, mp);
            mm_fs<T, D>(F, s + D, FP);
            sym_xat_plus<T, D>(FP, F, Q, Pp);
            smoother_element(FP, Pp, s, s + D, mp, el);
        }
        push_smoother(cr, el);
    }

    // CTA-level reverse scans of both side aggregates, then (last CTA to arrive) the scans over the CTA totals:
    // the two halves of the CTA run them side by side.  All threads of the CTA call this.
    template <int NW>
    PSSGP_DEV static void side_finish(const Params& p, Carry& cr, int lane, int wid, long nCta, long nChunksPad) {
        __shared__ T shw_s[SM ? NW * SA::NAGG : 1];
        __shared__ T shw_a[AD ? NW * AA::NAGG : 1];
        const long blk_s = nCta - 1 - (long)blockIdx.x;
        if constexpr (SM)
            cta_scan_publish<SA, NW, true>(cr.sa, lane, wid, blk_s, nCta, nChunksPad, p.sm.lane_excl, p.sm.warp_excl,
                                           p.sm.wagg, shw_s);
        if constexpr (AD)
            cta_scan_publish<AA, NW, true>(cr.aa, lane, wid, blk_s, nCta, nChunksPad, p.ad.lane_excl, p.ad.warp_excl,
                                           p.ad.wagg, shw_a);
        __shared__ bool is_last;
        __shared__ T sh_mid_s[SM ? 32 * SA::NAGG : 1];
        __shared__ T sh_mid_a[AD ? 32 * AA::NAGG : 1];
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned int t = atomicAdd(p.side_ticket, 1u);
            is_last = (t == gridDim.x - 1);
        }
        __syncthreads();
        if (is_last) {
            __threadfence();
            // both scans: side by side in the two halves of the CTA (named barriers 1 and 2); one scan: the whole CTA
            constexpr bool BOTH = SM && AD && NW >= 2;
            constexpr int GROUP = BOTH ? (NW / 2) * 32 : NW * 32;
            const bool summaries = p.sm_summary != nullptr || p.ad_summary != nullptr;
            const int tid = (int)threadIdx.x;
            if constexpr (SM) {
                if (tid < GROUP) {
                    if (summaries) {  // sm.wstate holds the per-CTA prefix aggregates in this mode
                        scan_prefix_body<SA>(p.sm.wagg, nCta, p.sm.wstate, p.sm_summary, sh_mid_s, tid, GROUP, 1);
                    } else {
                        typename SA::Params sp{};
                        scan_mid_body<SA>(sp, p.sm.wagg, nCta, p.sm.wstate, (T*)nullptr, sh_mid_s, tid, GROUP, 1);
                    }
                }
            }
            if constexpr (AD) {
                const int t2 = BOTH ? tid - GROUP : tid;
                if (t2 >= 0 && t2 < GROUP) {
                    if (summaries) {
                        scan_prefix_body<AA>(p.ad.wagg, nCta, p.ad.wstate, p.ad_summary, sh_mid_a, t2, GROUP, 2);
                    } else {
                        typename AA::Params ap{};
                        scan_mid_body<AA>(ap, p.ad.wagg, nCta, p.ad.wstate, (T*)nullptr, sh_mid_a, t2, GROUP, 2);
                    }
                }
            }
            __syncthreads();
            if (threadIdx.x == 0) *p.side_ticket = 0u;
        }
    }
};

// K3': RTS smoother recursion + adjoint recursion of the log-likelihood in one descending pass: both need
// (F, Q) of row k+1 carried in registers and the filtered moments of row k from the same staged row.
// State / aggregates are the smoother's followed by the adjoint's.  Outputs sms, sPs belong to the visited
// row (out_shift 0), dFs, dQs to the row above it (out_shift 1, see adjoint_small.cuh).
template <typename T, int D>
struct FusedRevAlg {
    using scalar = T;
    using SA = SmootherAlg<T, D>;
    using AA = AdjointAlg<T, D>;
    static const char* name_apply() { return "pks_bwd_apply_fused"; }
    static constexpr int KIND = KIND_ADJOINT;
    static constexpr int NS = nsym(D);
    static constexpr int NAGG = SA::NAGG + AA::NAGG;
    static constexpr int NSTATE = SA::NSTATE + AA::NSTATE;
    static constexpr int NACC = AA::NACC;
    static constexpr bool REVERSE = true;
    static constexpr int OUT_SHIFT = 1;
    static constexpr bool FLUSH = true;
    static constexpr bool HAS_DONE = false;
    static constexpr bool HAS_SIDE = false;
    static constexpr bool OUT8 = true;
    static constexpr int NIN = 5, NOUT = 4, WMAX = D * D;
    __host__ __device__ static constexpr int in_w(int a) { return AA::in_w(a); }
    __host__ __device__ static constexpr int out_w(int a) { return a == 0 ? D : D * D; }
    __host__ __device__ static constexpr int out_shift(int a) { return a < 2 ? 0 : 1; }

    struct Params {
        typename SA::Params s;
        typename AA::Params a;
    };
    __host__ __device__ __forceinline__ static const T* in_ptr(const Params& p, int a) { return AA::in_ptr(p.a, a); }
    __host__ __device__ __forceinline__ static T* out_ptr(const Params& p, int a) {
        return a == 0 ? p.s.sms : (a == 1 ? p.s.sPs : (a == 2 ? p.a.dFs : p.a.dQs));
    }
    using Ctx = typename AA::Ctx;
    PSSGP_DEV static void load_ctx(const Params& p, Ctx& c) { AA::load_ctx(p.a, c); }

    // (F, Q, y) of the row above the one being visited; `has`: the adjoint step of that row is this chunk's
    using Carry = typename AA::Carry;
    PSSGP_DEV static void carry_init(Carry& c, const Ctx&, long, long k_hi, const Params& p) {
        c.has = false;
        if (k_hi < p.s.n) {
            const T* pf = p.s.Fs + k_hi * (D * D);
            const T* pq = p.s.Qs + k_hi * (D * D);
            T qf[D * D];
#pragma unroll
            for (int e = 0; e < D * D; ++e) {
                c.F[e] = __ldg(pf + e);
                qf[e] = __ldg(pq + e);
            }
            SA::sym_pack(qf, c.Q);
        }
    }

    PSSGP_DEV static void apply(const T* s, const T* a, T* s2) {
        SA::apply(s, a, s2);
        AA::apply(s + SA::NSTATE, a + SA::NAGG, s2 + SA::NSTATE);
    }
    PSSGP_DEV static void load_init(const Params& p, T* s) {
        SA::load_init(p.s, s);
        AA::load_init(p.a, s + SA::NSTATE);
    }

    PSSGP_DEV static int step_row(T* s, const Ctx& cx, const T (&in)[NIN][WMAX], T (&out)[NOUT][WMAX], long k,
                                  const Params& p, T* acc, Carry& c) {
        T m[D], P[NS];
        AA::row_moments(in, m, P);
        // smoother (smoother_small.cuh step_row)
        if (k == p.s.n - 1 && p.s.last_special) {
#pragma unroll
            for (int i = 0; i < D; ++i) s[i] = m[i];
#pragma unroll
            for (int e = 0; e < NS; ++e) s[D + e] = P[e];
        } else {
            T a[SA::NAGG], s2[SA::NSTATE];
            SA::element(c.F, c.Q, m, P, a + SA::oE, a + SA::og, a + SA::oL);
            SA::apply(s, a, s2);
#pragma unroll
            for (int e = 0; e < SA::NSTATE; ++e) s[e] = s2[e];
        }
#pragma unroll
        for (int i = 0; i < D; ++i) out[0][i] = s[i];
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j < D; ++j) out[1][i * D + j] = s[D + sidx(i, j)];
        // adjoint step of row k+1 (adjoint_small.cuh step_row)
        int mask = 1;
        if (c.has) {
            AA::step_core(s + SA::NSTATE, cx, c, m, P, *reinterpret_cast<T(*)[2][WMAX]>(&out[2]), k + 1, p.a, acc);
            mask |= 2;
        }
        AA::carry_set(c, in, k, p.a);
        return mask;
    }
    PSSGP_DEV static int step_flush(T* s, const Ctx& cx, T (&out)[NOUT][WMAX], long k_lo, const Params& p, T* acc,
                                    Carry& c) {
        if (!c.has) return 0;
        T m[D], P[NS];
        AA::halo_moments(k_lo, p.a, m, P);
        AA::step_core(s + SA::NSTATE, cx, c, m, P, *reinterpret_cast<T(*)[2][WMAX]>(&out[2]), k_lo, p.a, acc);
        return 2;
    }
    PSSGP_DEV static void finish(const Params& p, int e, T tot, T* acc_out) { AA::finish(p.a, e, tot, acc_out); }
};

}  // namespace pssgp
