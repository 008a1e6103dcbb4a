// Warp-level kernels for state dimensions 5 <= d <= 32 (compile-time D): RBF order 6 (d = 6), Matern52 + RBF6
// (d = 9), quasi-periodic Periodic x Matern32 (d = 16 / 24 / 28), ... — BASELINE configs[2..4].
//
// One warp (or a group of WG warps for the larger D) owns one chunk of L consecutive time steps and walks it
// sequentially; the d x d state lives in that group's shared memory as zero-padded DP x LD tiles (DP = D rounded up
// to 8, LD = DP + 4 so that every fragment access pattern of mma.m8n8k4 is bank-conflict free), and every d x d
// product is a handful of FP64 tensor-core instructions (DMMA, mma.sync.m8n8k4.f64: the full FP64 rate of the SM
// for 1/8 of the issue slots of DFMA, measured with scripts/sm_probe.cu).  The rows of the LGSSM stream through a
// per-group ring of shared-memory slots filled by cp.async (LDGSTS straight into the padded layout), NSLOT - 1 rows
// ahead of the arithmetic.  No d x d solve is executed per time step anywhere:
//
//   K1 filter_reduce : chunk aggregate (A, b, C, J, eta) of pssgp/kalman/parallel.py:56-118 by the conditional
//                      recursion  A <- F A, C <- F C F^T + Q, rank-one measurement update     (3 products / step)
//   K2 forward       : seeded Kalman recursion (parallel.py:121-152 incl. the log-likelihood block :135-151) +, in
//                      the same pass, the chunk aggregate of the combined reverse scan (GRev, generic_algebras.cuh)
//                      appended on the LATER side, which costs one product because Abar_k = F_k - u t^T / s
//                                                                                            (3 products / step)
//   K3 reverse       : RTS smoother in modified Bryson-Frazier form (equal to parallel.py:155-196 in exact
//                      arithmetic) + adjoint of the log-likelihood (what TF autodiff provides in the reference),
//                      both seeded by the hierarchy over the chunk aggregates                 (9 products / step)
//
// The scans over the chunk aggregates (N / L of them) run through the CTA-cooperative hierarchy of generic.cu.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/pssgp_b200.h"
#include "generic_algebras.cuh"
#include "workspace.h"

namespace pssgp {
namespace mid {

#define MDEV __device__ __forceinline__

template <int D_, int WG_> struct Geo {
    static constexpr int D = D_, WG = WG_;
    static constexpr int DP = (D + 7) / 8 * 8;  // padded rows / columns (8 x 8 output tiles)
    static constexpr int MT = DP / 8;           // tiles per side
    static constexpr int KS = (D + 3) / 4;      // k-steps of 4
    static constexpr int LD = DP + 4;           // pitch in doubles: LD = 4 (mod 8) -> conflict-free fragment loads
    static constexpr int MSZ = DP * LD;         // doubles per matrix
    static constexpr int DD = D * D;
    static constexpr int NT = WG * 32;          // threads per group
};

struct Grp {
    int tid;     // thread in group
    int lane, r, c;
    int wig;     // warp in group
    int bar;     // named barrier of this group (WG > 1)
};

template <int WG> MDEV void gsync(const Grp& g) {
    if constexpr (WG == 1) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(g.bar), "n"(WG * 32) : "memory");
}

MDEV void dmma(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c[0]), "+d"(c[1])
                 : "d"(a), "d"(b));
}

MDEV void cp8(double* dst_smem, const double* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src)
                 : "memory");
}
MDEV void cp_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> MDEV void cp_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// C = (Cinit ? Cinit : 0) + op(A) op(B); every operand a zero-padded DP x LD tile set in shared memory.  The warps
// of the group take tile rows round-robin.  C must not alias A or B.
template <class G, bool TA, bool TB>
MDEV void mm(const Grp& g, const double* __restrict__ A, const double* __restrict__ B, double* __restrict__ C,
             const double* __restrict__ Cinit = nullptr) {
    constexpr int MT = G::MT, KS = G::KS, LD = G::LD;
    const int r = g.r, c = g.c;
#pragma unroll 1
    for (int mt = g.wig; mt < MT; mt += G::WG) {
        double acc[MT][2];
#pragma unroll
        for (int nt = 0; nt < MT; ++nt) {
            if (Cinit != nullptr) {
                const double2 v = *reinterpret_cast<const double2*>(Cinit + (8 * mt + r) * LD + 8 * nt + 2 * c);
                acc[nt][0] = v.x;
                acc[nt][1] = v.y;
            } else {
                acc[nt][0] = 0.0;
                acc[nt][1] = 0.0;
            }
        }
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            const double a = TA ? A[(4 * ks + c) * LD + 8 * mt + r] : A[(8 * mt + r) * LD + 4 * ks + c];
            double b[MT];
#pragma unroll
            for (int nt = 0; nt < MT; ++nt)
                b[nt] = TB ? B[(8 * nt + r) * LD + 4 * ks + c] : B[(4 * ks + c) * LD + 8 * nt + r];
#pragma unroll
            for (int nt = 0; nt < MT; ++nt) dmma(acc[nt], a, b[nt]);
        }
#pragma unroll
        for (int nt = 0; nt < MT; ++nt)
            *reinterpret_cast<double2*>(C + (8 * mt + r) * LD + 8 * nt + 2 * c) = make_double2(acc[nt][0], acc[nt][1]);
    }
}

// out = op(M) v (+ add); v, out, add: zero-padded DP vectors in shared memory (out must not alias v).
template <class G, bool TRANS>
MDEV void mv(const Grp& g, const double* __restrict__ M, const double* __restrict__ v, double* __restrict__ out,
             const double* __restrict__ add = nullptr) {
    constexpr int MT = G::MT, KS = G::KS, LD = G::LD;
    const int r = g.r, c = g.c;
#pragma unroll
    for (int t = g.wig; t < MT; t += G::WG) {
        double s = 0.0;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            const double m = TRANS ? M[(4 * ks + c) * LD + 8 * t + r] : M[(8 * t + r) * LD + 4 * ks + c];
            s = fma(m, v[4 * ks + c], s);
        }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (c == 0) out[8 * t + r] = add != nullptr ? s + add[8 * t + r] : s;
    }
}

// N dot products of zero-padded DP vectors at once; every lane of every warp gets the results.
template <class G, int N> MDEV void dots(const Grp& g, const double* const (&a)[N], const double* const (&b)[N], double (&out)[N]) {
#pragma unroll
    for (int i = 0; i < N; ++i) out[i] = g.lane < G::DP ? a[i][g.lane] * b[i][g.lane] : 0.0;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
        for (int i = 0; i < N; ++i) out[i] += __shfl_xor_sync(0xffffffffu, out[i], off);
    }
}

template <class G, class Fn> MDEV void sweep(const Grp& g, Fn fn) {
#pragma unroll 1
    for (int idx = g.tid; idx < G::DD; idx += G::NT) {
        const int i = idx / G::D, j = idx - i * G::D;
        fn(i, j, idx);
    }
}
template <class G, class Fn> MDEV void vsweep(const Grp& g, Fn fn) {
    for (int i = g.tid; i < G::D; i += G::NT) fn(i);
}

template <class G> MDEV void load_mat(const Grp& g, double* dst, const double* src) {
    sweep<G>(g, [&](int i, int j, int idx) { cp8(dst + i * G::LD + j, src + idx); });
}
// M <- (M + M^T) / 2 in place (every unordered pair is handled by one thread)
template <class G> MDEV void sym_inplace(const Grp& g, double* M) {
    sweep<G>(g, [&](int i, int j, int) {
        if (j > i) {
            const double v = 0.5 * (M[i * G::LD + j] + M[j * G::LD + i]);
            M[i * G::LD + j] = v;
            M[j * G::LD + i] = v;
        }
    });
}
template <class G> MDEV void load_vec(const Grp& g, double* dst, const double* src) {
    vsweep<G>(g, [&](int i) { cp8(dst + i, src + i); });
}

template <int NT_, int NSLOT_> MDEV int slot_of(long row) { return (int)((row + 4L * NSLOT_) % NSLOT_); }

template <int D> constexpr int default_wg() { return D <= 16 ? 1 : (D <= 24 ? 3 : 4); }
template <int D> constexpr int default_nslot() { return D <= 8 ? 4 : (D <= 16 ? 3 : 2); }

struct Params {
    const double *Fs, *Qs, *y, *H, *R, *P0, *m0, *g;
    const double *fms_in, *fPs_in;  // filtered moments as inputs (K3; K2 in STORED mode)
    double *fms, *fPs, *sms, *sPs, *dFs, *dQs, *dP0;
    double* proj;  // fragment K3 only: [n,2] = (H sm_k, H sP_k H^T) instead of sms / sPs (predict_f, model.py:107-111)
    long n;
    int first_special;
};

template <class G> MDEV Grp make_grp(int gi) {
    Grp g;
    g.tid = threadIdx.x - gi * G::NT;
    g.lane = threadIdx.x & 31;
    g.r = g.lane >> 2;
    g.c = g.lane & 3;
    g.wig = g.tid >> 5;
    g.bar = 1 + gi;
    return g;
}

// ------------------------------------------------------------------------------------------------
// K1: chunk aggregates of the filter (layout of GFilter: A | C | J | b | eta, dense d x d)
// ------------------------------------------------------------------------------------------------
template <int D, int WG> struct K1 {
    using G = Geo<D, WG>;
    static constexpr int NSLOT = default_nslot<D>();
    static constexpr int NMAT = 6 + 2 * NSLOT;  // Aa Ab C T2 J (+1 spare for alignment of thought) + ring (F, Q)
    static constexpr int NVEC = 6;              // ba bb eta h u w
    static constexpr int GROUP_DOUBLES = NMAT * G::MSZ + NVEC * G::DP;
    static constexpr int GPC_MAX = (216 * 1024) / (GROUP_DOUBLES * 8);
    static constexpr int GPC = GPC_MAX * WG > 16 ? 16 / WG : (GPC_MAX < 1 ? 1 : GPC_MAX);
};

template <int D, int WG>
__global__ void __launch_bounds__(K1<D, WG>::GPC* WG * 32)
k1_filter_reduce(Params p, int L, long nchunks, double* __restrict__ aggs) {
    using K = K1<D, WG>;
    using G = typename K::G;
    constexpr int MSZ = G::MSZ, DP = G::DP, LD = G::LD, NSLOT = K::NSLOT, DD = G::DD;
    extern __shared__ __align__(16) double smem[];
    const int gi = threadIdx.x / G::NT;
    const long chunk = (long)blockIdx.x * (blockDim.x / G::NT) + gi;  // groups per CTA: run-time (<= K::GPC)
    if (chunk >= nchunks) return;
    const Grp g = make_grp<G>(gi);
    double* S = smem + (size_t)gi * K::GROUP_DOUBLES;
    for (int i = g.tid; i < K::GROUP_DOUBLES; i += G::NT) S[i] = 0.0;
    double *A = S, *A2 = S + MSZ, *C = S + 2 * MSZ, *T2 = S + 3 * MSZ, *J = S + 4 * MSZ;
    double* ring = S + 6 * MSZ;
    double* V = S + K::NMAT * MSZ;
    double *b = V, *b2 = V + DP, *eta = V + 2 * DP, *h = V + 3 * DP, *u = V + 4 * DP, *w = V + 5 * DP;
    gsync<WG>(g);
    vsweep<G>(g, [&](int i) {
        h[i] = p.H[i];
        A[i * LD + i] = 1.0;
    });
    const double Rv = p.R[0];
    const long k_lo = chunk * (long)L;
    const long k_hi = (k_lo + L < p.n) ? k_lo + L : p.n;
    const int nrows = (int)(k_hi - k_lo);
    constexpr int PD = NSLOT - 1;
    auto issue = [&](int i) {
        if (i < nrows) {
            double* sl = ring + (i % NSLOT) * 2 * MSZ;
            load_mat<G>(g, sl, p.Fs + (k_lo + i) * DD);
            load_mat<G>(g, sl + MSZ, p.Qs + (k_lo + i) * DD);
        }
        cp_commit();
    };
#pragma unroll 1
    for (int i = 0; i < PD; ++i) issue(i);
    double ynext = p.y[k_lo];
#pragma unroll 1
    for (int i = 0; i < nrows; ++i) {
        const long k = k_lo + i;
        issue(i + PD);
        cp_wait<PD>();
        gsync<WG>(g);
        const double yk = ynext;
        if (i + 1 < nrows) ynext = p.y[k + 1];
        const double* F = ring + (i % NSLOT) * 2 * MSZ;
        double* Q = ring + (i % NSLOT) * 2 * MSZ + MSZ;
        sym_inplace<G>(g, Q);
        const bool first = (k == 0 && p.first_special);
        if (!first) {
            mm<G, false, false>(g, F, A, A2);
            mm<G, false, false>(g, F, C, T2);
            mv<G, false>(g, F, b, b2);
            gsync<WG>(g);
            mm<G, false, true>(g, T2, F, C, Q);
            double* t = A; A = A2; A2 = t;
            t = b; b = b2; b2 = t;
            gsync<WG>(g);
        }
        if (!isnan(yk)) {
            mv<G, false>(g, C, h, u);
            mv<G, true>(g, A, h, w);
            gsync<WG>(g);
            double dv[2];
            {
                const double* const da[2] = {h, h};
                const double* const db[2] = {u, b};
                dots<G, 2>(g, da, db, dv);
            }
            const double is = 1.0 / (Rv + dv[0]);
            const double eis = (yk - dv[1]) * is;
            if constexpr (WG > 1) gsync<WG>(g);
            sweep<G>(g, [&](int i2, int j2, int) {
                const int o = i2 * LD + j2;
                const double ui = u[i2] * is, wj = w[j2];
                J[o] = fma(w[i2] * is, wj, J[o]);
                A[o] = fma(-ui, wj, A[o]);
                C[o] = fma(-ui, u[j2], C[o]);
            });
            vsweep<G>(g, [&](int i2) {
                eta[i2] = fma(w[i2], eis, eta[i2]);
                b[i2] = fma(u[i2], eis, b[i2]);
            });
        }
        gsync<WG>(g);
    }
    cp_wait<0>();
    double* out = aggs + chunk * (3 * DD + 2 * D);
    sweep<G>(g, [&](int i, int j, int idx) {
        out[idx] = A[i * LD + j];
        out[DD + idx] = 0.5 * (C[i * LD + j] + C[j * LD + i]);
        out[2 * DD + idx] = 0.5 * (J[i * LD + j] + J[j * LD + i]);
    });
    vsweep<G>(g, [&](int i) {
        out[3 * DD + i] = b[i];
        out[3 * DD + D + i] = eta[i];
    });
}

// ------------------------------------------------------------------------------------------------
// K2: seeded filter recursion (+ log-likelihood) and, with REV, the chunk aggregate of the combined reverse scan.
// STORED: the filtered moments are read (pssgp_pkf_backward on its own) instead of recomputed; nothing is written
// but the reverse aggregate.
// ------------------------------------------------------------------------------------------------
template <int D, int WG, bool REV, bool STORED> struct K2 {
    using G = Geo<D, WG>;
    static constexpr int NSLOT = default_nslot<D>();
    static constexpr int SLOT_M = STORED ? 3 : 2;
    static constexpr int SLOT = SLOT_M * G::MSZ + (STORED ? G::DP : 0);
    static constexpr int NMAT = 3 + (REV ? 4 : 0);  // Pa Pb T1 | Ab Gb Ba Bm
    static constexpr int NVEC = 4 + (REV ? 4 : 0);  // ma mb h u | w t aa ab
    static constexpr int GROUP_DOUBLES = NMAT * G::MSZ + NVEC * G::DP + NSLOT * SLOT;
    static constexpr int GPC_MAX = (216 * 1024) / (GROUP_DOUBLES * 8);
    static constexpr int GPC = GPC_MAX * WG > 16 ? 16 / WG : (GPC_MAX < 1 ? 1 : GPC_MAX);
};

template <int D, int WG, bool REV, bool STORED>
__global__ void __launch_bounds__(K2<D, WG, REV, STORED>::GPC* WG * 32)
k2_forward(Params p, int L, long nchunks, const double* __restrict__ fstates, double* __restrict__ part,
           double* __restrict__ raggs) {
    using K = K2<D, WG, REV, STORED>;
    using G = typename K::G;
    constexpr int MSZ = G::MSZ, DP = G::DP, LD = G::LD, NSLOT = K::NSLOT, DD = G::DD;
    extern __shared__ __align__(16) double smem[];
    const int gi = threadIdx.x / G::NT;
    const long chunk = (long)blockIdx.x * (blockDim.x / G::NT) + gi;  // groups per CTA: run-time (<= K::GPC)
    if (chunk >= nchunks) return;
    const Grp g = make_grp<G>(gi);
    double* S = smem + (size_t)gi * K::GROUP_DOUBLES;
    for (int i = g.tid; i < K::GROUP_DOUBLES; i += G::NT) S[i] = 0.0;
    double *P = S, *PP = S + MSZ, *T1 = S + 2 * MSZ;
    double *Ab = S + 3 * MSZ, *Gb = S + 4 * MSZ, *Ba = S + 5 * MSZ, *Bm = S + 6 * MSZ;  // REV only
    double* V = S + K::NMAT * MSZ;
    double *m = V, *mp = V + DP, *h = V + 2 * DP, *u = V + 3 * DP;
    double *w = V + 4 * DP, *t = V + 5 * DP, *a = V + 6 * DP, *a2 = V + 7 * DP;  // REV only
    double* ring = V + K::NVEC * DP;
    gsync<WG>(g);
    const long k_lo = chunk * (long)L;
    const long k_hi = (k_lo + L < p.n) ? k_lo + L : p.n;
    const int nrows = (int)(k_hi - k_lo);
    vsweep<G>(g, [&](int i) {
        h[i] = p.H[i];
        if constexpr (REV) Ab[i * LD + i] = 1.0;
    });
    if constexpr (!STORED) {
        const double* st = fstates + chunk * (D + DD);
        vsweep<G>(g, [&](int i) { m[i] = st[i]; });
        sweep<G>(g, [&](int i, int j, int idx) { P[i * LD + j] = st[D + idx]; });
    }
    const double Rv = p.R[0];
    constexpr int PD = NSLOT - 1;
    auto issue = [&](int i) {
        if (i < nrows) {
            double* sl = ring + (i % NSLOT) * K::SLOT;
            const long k = k_lo + i;
            load_mat<G>(g, sl, p.Fs + k * DD);
            load_mat<G>(g, sl + MSZ, p.Qs + k * DD);
            if constexpr (STORED) {
                load_mat<G>(g, sl + 2 * MSZ, k > 0 ? p.fPs_in + (k - 1) * DD : p.P0);
                if (k > 0) load_vec<G>(g, sl + 3 * MSZ, p.fms_in + (k - 1) * D);
                else if (p.m0 != nullptr) load_vec<G>(g, sl + 3 * MSZ, p.m0);
                else vsweep<G>(g, [&](int i2) { sl[3 * MSZ + i2] = 0.0; });
            }
        }
        cp_commit();
    };
#pragma unroll 1
    for (int i = 0; i < PD; ++i) issue(i);
    double ynext = p.y[k_lo];
    double ll = 0.0;
#pragma unroll 1
    for (int i = 0; i < nrows; ++i) {
        const long k = k_lo + i;
        issue(i + PD);
        cp_wait<PD>();
        gsync<WG>(g);
        const double yk = ynext;
        if (i + 1 < nrows) ynext = p.y[k + 1];
        const bool obs = !isnan(yk);
        double* sl = ring + (i % NSLOT) * K::SLOT;
        const double* F = sl;
        double* Q = sl + MSZ;
        sym_inplace<G>(g, Q);
        const double* Pc = STORED ? sl + 2 * MSZ : P;
        const double* mc = STORED ? sl + 3 * MSZ : m;
        const bool first = (k == 0 && p.first_special);
        mm<G, false, false>(g, F, Pc, T1);
        mv<G, false>(g, F, mc, mp);
        if (REV && !first) {
            mm<G, false, false>(g, F, Ab, Gb);
            mv<G, true>(g, F, h, w);
        }
        gsync<WG>(g);
        mm<G, false, true>(g, T1, F, PP, Q);
        if (REV && !first) mv<G, true>(g, Ab, w, t);
        gsync<WG>(g);
        mv<G, false>(g, PP, h, u);
        gsync<WG>(g);
        double dv[2];
        {
            const double* const da[2] = {h, h};
            const double* const db[2] = {u, mp};
            dots<G, 2>(g, da, db, dv);
        }
        double s = Rv + dv[0];
        double e = yk - dv[1];
        if (!STORED && obs) ll += -0.5 * (log(6.283185307179586476925286766559 * s) + e * e / s);
        const double* Psrc = PP;
        const double* msrc = mp;
        if (first) {
            // parallel.py:24-30: the first update is made on (m0, P0) directly, without prediction
            gsync<WG>(g);
            mv<G, false>(g, Pc, h, u);
            gsync<WG>(g);
            const double* const da[2] = {h, h};
            const double* const db[2] = {u, mc};
            dots<G, 2>(g, da, db, dv);
            s = Rv + dv[0];
            e = yk - dv[1];
            Psrc = Pc;
            msrc = mc;
        }
        const double is = obs ? 1.0 / s : 0.0;
        const double eis = obs ? e * is : 0.0;
        if constexpr (WG > 1) gsync<WG>(g);
        if constexpr (!STORED) {
            double* Pdst = (Psrc == P) ? PP : P;
            double* mdst = (msrc == m) ? mp : m;
            double* oP = p.fPs + k * DD;
            double* om = p.fms + k * D;
            sweep<G>(g, [&](int i2, int j2, int idx) {
                const double v = fma(-u[i2] * is, u[j2], 0.5 * (Psrc[i2 * LD + j2] + Psrc[j2 * LD + i2]));
                Pdst[i2 * LD + j2] = v;
                oP[idx] = v;
            });
            vsweep<G>(g, [&](int i2) {
                const double v = fma(u[i2], eis, msrc[i2]);
                mdst[i2] = v;
                om[i2] = v;
            });
            // P, m <- the buffers just written; PP, mp <- the others
            if (Pdst != P) { PP = P; P = Pdst; }
            if (mdst != m) { mp = m; m = mdst; }
        }
        if (REV && !first) {
            // append step k on the later side of the chunk's reverse aggregate
            if (obs) {
                const double beta = 0.5 * (eis * eis - is);
                sweep<G>(g, [&](int i2, int j2, int) {
                    const int o = i2 * LD + j2;
                    const double ti = t[i2], tj = t[j2];
                    Gb[o] = fma(-u[i2] * is, tj, Gb[o]);
                    Ba[o] += beta * ti * tj + 0.5 * eis * (ti * a[j2] + a[i2] * tj);
                    Bm[o] = fma(ti * is, tj, Bm[o]);
                });
                vsweep<G>(g, [&](int i2) { a2[i2] = fma(t[i2], eis, a[i2]); });
                double* tt = a; a = a2; a2 = tt;
            }
            double* tt = Ab; Ab = Gb; Gb = tt;
        }
        gsync<WG>(g);
    }
    cp_wait<0>();
    if (!STORED && part != nullptr && g.tid == 0) part[chunk] = ll;
    if constexpr (REV) {
        double* out = raggs + (nchunks - 1 - chunk) * (3 * DD + D);
        sweep<G>(g, [&](int i, int j, int idx) {
            out[idx] = Ab[i * LD + j];
            out[DD + idx] = 0.5 * (Ba[i * LD + j] + Ba[j * LD + i]);
            out[2 * DD + idx] = 0.5 * (Bm[i * LD + j] + Bm[j * LD + i]);
        });
        vsweep<G>(g, [&](int i) { out[3 * DD + i] = a[i]; });
    }
}

// ------------------------------------------------------------------------------------------------
// K3: reverse pass — smoothed moments (SMOOTH) and / or gradient of the log-likelihood (ADJ).
// State entering a chunk from above (hierarchy, layout of GRev): dm | lam | dP | Lam.
// ------------------------------------------------------------------------------------------------
template <int D, int WG, bool SMOOTH, bool ADJ> struct K3 {
    using G = Geo<D, WG>;
    static constexpr int NSLOT = default_nslot<D>() < 3 ? 3 : default_nslot<D>();
    static constexpr int SLOT = 3 * G::MSZ + G::DP;  // F | Q | fP | fm
    static constexpr int NMAT = 6;                   // Lam dP W0 W1 W2 W3
    static constexpr int NVEC = 15;
    static constexpr int GROUP_DOUBLES = NMAT * G::MSZ + NVEC * G::DP + NSLOT * SLOT;
    static constexpr int GPC_MAX = (216 * 1024) / (GROUP_DOUBLES * 8);
    static constexpr int GPC = GPC_MAX * WG > 16 ? 16 / WG : (GPC_MAX < 1 ? 1 : GPC_MAX);
};

template <int D, int WG, bool SMOOTH, bool ADJ>
__global__ void __launch_bounds__(K3<D, WG, SMOOTH, ADJ>::GPC* WG * 32)
k3_reverse(Params p, int L, long nchunks, const double* __restrict__ rstates, double* __restrict__ part) {
    using K = K3<D, WG, SMOOTH, ADJ>;
    using G = typename K::G;
    constexpr int MSZ = G::MSZ, DP = G::DP, LD = G::LD, NSLOT = K::NSLOT, DD = G::DD;
    extern __shared__ __align__(16) double smem[];
    const int gi = threadIdx.x / G::NT;
    const long chunk = (long)blockIdx.x * (blockDim.x / G::NT) + gi;  // groups per CTA: run-time (<= K::GPC)
    if (chunk >= nchunks) return;
    const Grp g = make_grp<G>(gi);
    double* S = smem + (size_t)gi * K::GROUP_DOUBLES;
    for (int i = g.tid; i < K::GROUP_DOUBLES; i += G::NT) S[i] = 0.0;
    double *Lam = S, *dP = S + MSZ, *W0 = S + 2 * MSZ, *W1 = S + 3 * MSZ, *W2 = S + 4 * MSZ, *W3 = S + 5 * MSZ;
    double* V = S + K::NMAT * MSZ;
    double *lam = V, *lam2 = V + DP, *dm = V + 2 * DP, *dm2 = V + 3 * DP, *h = V + 4 * DP, *u = V + 5 * DP;
    double *gv = V + 6 * DP, *Pu = V + 7 * DP, *ut = V + 8 * DP, *Pput = V + 9 * DP, *mp = V + 10 * DP;
    double *dmp = V + 11 * DP, *lt = V + 12 * DP, *v1 = V + 13 * DP;
    double* ring = V + K::NVEC * DP;
    gsync<WG>(g);
    const long k_lo = chunk * (long)L;
    const long k_hi = (k_lo + L < p.n) ? k_lo + L : p.n;
    const int nrows = (int)(k_hi - k_lo);
    {
        const double* st = rstates + (nchunks - 1 - chunk) * (2 * DD + 2 * D);
        vsweep<G>(g, [&](int i) {
            h[i] = p.H[i];
            dm[i] = st[i];
            lam[i] = st[D + i];
        });
        sweep<G>(g, [&](int i, int j, int idx) {
            dP[i * LD + j] = st[2 * D + idx];
            Lam[i * LD + j] = st[2 * D + DD + idx];
        });
    }
    const double Rv = p.R[0];
    const double gl = ADJ ? p.g[0] : 1.0;
    // rows are visited in descending order: visit j <-> row k_hi - 1 - j; one more "visit" loads the filtered
    // moments of row k_lo - 1 (or the prior) that the step of row k_lo needs
    constexpr int PD = NSLOT - 1;
    auto slot = [&](long row) { return ring + (int)((row + 8L * NSLOT) % NSLOT) * K::SLOT; };
    auto issue = [&](int j) {
        if (j <= nrows) {
            const long row = k_hi - 1 - j;
            double* sl = slot(row);
            if (j < nrows) {
                load_mat<G>(g, sl, p.Fs + row * DD);
                load_mat<G>(g, sl + MSZ, p.Qs + row * DD);
            }
            if (row >= 0) {
                load_mat<G>(g, sl + 2 * MSZ, p.fPs_in + row * DD);
                load_vec<G>(g, sl + 3 * MSZ, p.fms_in + row * D);
            } else {
                load_mat<G>(g, sl + 2 * MSZ, p.P0);
                if (p.m0 != nullptr) load_vec<G>(g, sl + 3 * MSZ, p.m0);
                else vsweep<G>(g, [&](int i2) { sl[3 * MSZ + i2] = 0.0; });
            }
        }
        cp_commit();
    };
#pragma unroll 1
    for (int j = 0; j < PD; ++j) issue(j);
    double ynext = p.y[k_hi - 1];
    double dRacc = 0.0, dHacc = 0.0;
#pragma unroll 1
    for (int j = 0; j < nrows; ++j) {
        const long k = k_hi - 1 - j;
        issue(j + PD);
        cp_wait<PD - 1>();  // rows k and k - 1 have landed
        gsync<WG>(g);
        const double yk = ynext;
        if (j + 1 < nrows) ynext = p.y[k - 1];
        const bool obs = !isnan(yk);
        double* sl = slot(k);
        const double* F = sl;
        double* Q = sl + MSZ;
        sym_inplace<G>(g, Q);
        const double* Pk = sl + 2 * MSZ;
        const double* mk = sl + 3 * MSZ;
        const double* slp = slot(k - 1);
        const double* Pprev = slp + 2 * MSZ;
        const double* mprev = slp + 3 * MSZ;
        const bool first = (k == 0 && p.first_special);
        if constexpr (SMOOTH) {
            // sm_k = m_k - P_k lam_k ; sP_k = P_k - P_k Lam_k P_k   (state entering from above)
            mm<G, false, false>(g, Lam, Pk, W0);
            mv<G, false>(g, Pk, lam, v1);
            gsync<WG>(g);
            mm<G, false, false>(g, Pk, W0, W1);
            gsync<WG>(g);
            double* oP = p.sPs + k * DD;
            double* om = p.sms + k * D;
            sweep<G>(g, [&](int i2, int j2, int idx) {
                oP[idx] = 0.5 * ((Pk[i2 * LD + j2] + Pk[j2 * LD + i2]) - (W1[i2 * LD + j2] + W1[j2 * LD + i2]));
            });
            vsweep<G>(g, [&](int i2) { om[i2] = mk[i2] - v1[i2]; });
        }
        // forward quantities of step k
        mm<G, false, false>(g, F, Pprev, W2);
        mv<G, false>(g, F, mprev, mp);
        gsync<WG>(g);
        mm<G, false, true>(g, W2, F, W3, Q);  // Pp
        gsync<WG>(g);
        mv<G, false>(g, W3, h, u);
        gsync<WG>(g);
        double s, r;
        {
            double dv[2];
            const double* const da[2] = {h, h};
            const double* const db[2] = {u, mp};
            dots<G, 2>(g, da, db, dv);
            s = Rv + dv[0];
            r = yk - dv[1];
        }
        const double* dPp = dP;  // adjoint w.r.t. the predicted covariance (aliases dP when nothing is observed)
        const double* Lt = Lam;
        const double* dmpv = dm;
        const double* ltv = lam;
        if (first) {
            // step 0 of the global series: the log-likelihood term sees (F0 m0, F0 P0 F0^T + Q0), the update is made
            // on (m0, P0) directly (parallel.py:24-30, :136-141)
            double sbar0 = 0.0, rbar0 = 0.0;
            if (obs) {
                const double is = 1.0 / s;
                sbar0 = 0.5 * (r * r * is * is - is);
                rbar0 = -r * is;
                if constexpr (ADJ) {
                    dRacc += sbar0;
                    if (g.tid < D) dHacc += 2.0 * sbar0 * u[g.tid] - mp[g.tid] * rbar0;
                }
            }
            if constexpr (ADJ) {
                // dPp0 = sbar0 h h^T -> W0 ; dmp0 = -h rbar0 -> dmp
                if constexpr (WG > 1) gsync<WG>(g);
                sweep<G>(g, [&](int i2, int j2, int idx) {
                    const double v = sbar0 * h[i2] * h[j2];
                    W0[i2 * LD + j2] = v;
                    p.dQs[k * DD + idx] = gl * v;
                });
                vsweep<G>(g, [&](int i2) { dmp[i2] = -h[i2] * rbar0; });
                gsync<WG>(g);
                mm<G, false, false>(g, W0, F, W2);  // X
                gsync<WG>(g);
                mm<G, false, false>(g, W2, Pprev, W1);  // Y
                mm<G, true, false>(g, F, W2, W3);       // F^T X
                gsync<WG>(g);
                sweep<G>(g, [&](int i2, int j2, int idx) {
                    p.dFs[k * DD + idx] = gl * fma(2.0, W1[i2 * LD + j2], dmp[i2] * mprev[j2]);
                });
                // adjoint of the update on (m0, P0)
                mv<G, false>(g, Pprev, h, u);
                gsync<WG>(g);
                double s0, r0;
                {
                    double dv[2];
                    const double* const da[2] = {h, h};
                    const double* const db[2] = {u, mprev};
                    dots<G, 2>(g, da, db, dv);
                    s0 = Rv + dv[0];
                    r0 = yk - dv[1];
                }
                if (obs) {
                    mv<G, false>(g, dP, u, Pu);
                    gsync<WG>(g);
                    double dv[2];
                    const double* const da[2] = {u, u};
                    const double* const db[2] = {dm, Pu};
                    dots<G, 2>(g, da, db, dv);
                    const double is0 = 1.0 / s0;
                    const double rbar = dv[0] * is0;
                    const double sbar = (-dv[0] * r0 + dv[1]) * is0 * is0;
                    if constexpr (WG > 1) gsync<WG>(g);
                    vsweep<G>(g, [&](int i2) { ut[i2] = dm[i2] * r0 * is0 - 2.0 * Pu[i2] * is0 + sbar * h[i2]; });
                    gsync<WG>(g);
                    mv<G, false>(g, Pprev, ut, Pput);
                    gsync<WG>(g);
                    dRacc += sbar;
                    if (g.tid < D) dHacc += sbar * u[g.tid] + Pput[g.tid] - mprev[g.tid] * rbar;
                    if (p.dP0 != nullptr)
                        sweep<G>(g, [&](int i2, int j2, int idx) {
                            p.dP0[idx] = gl * (W3[i2 * LD + j2] + dP[i2 * LD + j2] + 0.5 * (ut[i2] * h[j2] + h[i2] * ut[j2]));
                        });
                } else if (p.dP0 != nullptr) {
                    sweep<G>(g, [&](int i2, int j2, int idx) { p.dP0[idx] = gl * (W3[i2 * LD + j2] + dP[i2 * LD + j2]); });
                }
            }
            gsync<WG>(g);
            continue;  // k == 0: nothing below this row
        }
        if (obs) {
            const double is = 1.0 / s;
            if constexpr (ADJ) mv<G, false>(g, dP, u, Pu);
            if constexpr (SMOOTH) mv<G, false>(g, Lam, u, gv);
            gsync<WG>(g);
            double dv[4];
            {
                const double* const da[4] = {u, u, u, u};
                const double* const db[4] = {dm, Pu, lam, gv};
                dots<G, 4>(g, da, db, dv);
            }
            const double udm = ADJ ? dv[0] : 0.0, uPu = ADJ ? dv[1] : 0.0, ulam = SMOOTH ? dv[2] : 0.0,
                         alpha = SMOOTH ? dv[3] : 0.0;
            const double rbar = (udm - r) * is;
            const double sbar = (-udm * r + uPu) * is * is + 0.5 * (r * r * is * is - is);
            if constexpr (WG > 1) gsync<WG>(g);
            vsweep<G>(g, [&](int i2) {
                if constexpr (ADJ) {
                    ut[i2] = dm[i2] * r * is - 2.0 * Pu[i2] * is + sbar * h[i2];
                    dmp[i2] = dm[i2] - h[i2] * rbar;
                }
                if constexpr (SMOOTH) lt[i2] = lam[i2] - h[i2] * (ulam + r) * is;
            });
            gsync<WG>(g);
            if constexpr (ADJ) {
                mv<G, false>(g, W3, ut, Pput);
                gsync<WG>(g);
                dRacc += sbar;
                if (g.tid < D) dHacc += sbar * u[g.tid] + Pput[g.tid] - mp[g.tid] * rbar;
            }
            const double cm = is + alpha * is * is;
            sweep<G>(g, [&](int i2, int j2, int) {
                const int o = i2 * LD + j2;
                if constexpr (ADJ) W0[o] = dP[o] + 0.5 * (ut[i2] * h[j2] + h[i2] * ut[j2]);
                if constexpr (SMOOTH) W1[o] = Lam[o] - (h[i2] * gv[j2] + gv[i2] * h[j2]) * is + h[i2] * h[j2] * cm;
            });
            dPp = W0;
            Lt = W1;
            dmpv = dmp;
            ltv = lt;
            gsync<WG>(g);
        }
        if constexpr (ADJ) {
            sweep<G>(g, [&](int i2, int j2, int idx) { p.dQs[k * DD + idx] = gl * dPp[i2 * LD + j2]; });
            mm<G, false, false>(g, dPp, F, W2);  // X = dPp F
            mv<G, true>(g, F, dmpv, dm2);
        }
        if constexpr (SMOOTH) {
            mm<G, false, false>(g, Lt, F, W3);  // X2 = Lt F   (Pp is dead by now)
            mv<G, true>(g, F, ltv, lam2);
        }
        gsync<WG>(g);
        if constexpr (ADJ) {
            mm<G, false, false>(g, W2, Pprev, W0);  // Y = X P_{k-1}
            mm<G, true, false>(g, F, W2, dP);       // dP' = F^T X
        }
        if constexpr (SMOOTH) mm<G, true, false>(g, F, W3, Lam);  // Lam' = F^T X2
        gsync<WG>(g);
        if constexpr (ADJ) {
            sweep<G>(g, [&](int i2, int j2, int idx) {
                p.dFs[k * DD + idx] = gl * fma(2.0, W0[i2 * LD + j2], dmpv[i2] * mprev[j2]);
            });
            double* tt = dm; dm = dm2; dm2 = tt;
        }
        if constexpr (SMOOTH) {
            double* tt = lam; lam = lam2; lam2 = tt;
        }
        gsync<WG>(g);
    }
    cp_wait<0>();
    if constexpr (ADJ) {
        if (g.tid == 0) part[chunk * (1 + D)] = dRacc;
        if (g.tid < D) part[chunk * (1 + D) + 1 + g.tid] = dHacc;
    }
}

}  // namespace mid
}  // namespace pssgp
