// Warp-level hierarchy over the chunk aggregates of the d > 4 path (replaces the CTA-cooperative kernels of
// generic.cu for 5 <= d <= 32): one warp combines a group of GF consecutive aggregates (up-sweep), one warp walks the
// few aggregates of the top level from the initial state, one warp per group turns the group-entry state into the
// entry state of each member (down-sweep).  Matrices are zero-padded DP x LD tiles in the warp's shared memory and
// every d x d product is DMMA (mid.cuh); the one solve of the filtering operator, (I + C1 J2)^-1 [A1 | b1 + C1 eta2 |
// C1 A2^T]  (pssgp/kalman/parallel.py:100-118: both tf.linalg.solve calls share this matrix), is a warp-level
// Gauss-Jordan elimination with partial pivoting.  The combined reverse operator (GRev) is products only.
// Global layouts are those of GFilter / GRev (generic_algebras.cuh).
#pragma once
#include "mid.cuh"

namespace pssgp {
namespace mid {
namespace hier {

constexpr int GF = 4;        // fan-in of a hierarchy level (short sequential chains: the nodes are latency-bound)
constexpr int TOPMAX = 4;    // aggregates the top group walks sequentially
// warps that cooperate on one node of the hierarchy (one warp up to d = 16; the d^3 work of larger d is spread over 4)
template <int D> constexpr int hier_wg() { return D <= 16 ? 1 : 4; }

// [M | B] (D x (D + NR), pitch WP) -> [. | M^-1 B]   (columns of M are not cleaned up)
template <class G, int NR, int WP> MDEV void wsolve(const Grp& g, double* W) {
    constexpr int D = G::D, NC = D + NR, NT = G::NT;
    const int lane = g.lane;
#pragma unroll 1
    for (int col = 0; col < D; ++col) {
        // every warp of the group finds the pivot row on its own (same data, same answer)
        double v = (lane >= col && lane < D) ? fabs(W[lane * WP + col]) : -1.0;
        int idx = lane;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, v, off);
            const int oi = __shfl_xor_sync(0xffffffffu, idx, off);
            if (ov > v || (ov == v && oi < idx)) {
                v = ov;
                idx = oi;
            }
        }
        const int p = idx;
        const double inv = 1.0 / W[p * WP + col];
        gsync<G::WG>(g);
#pragma unroll 1
        for (int j = col + g.tid; j < NC; j += NT) {
            const double a = W[col * WP + j], b = W[p * WP + j];
            W[p * WP + j] = a;
            W[col * WP + j] = b * inv;
        }
        gsync<G::WG>(g);
#pragma unroll 1
        for (int j = col + 1 + g.tid; j < NC; j += NT) {
            const double pj = W[col * WP + j];
#pragma unroll(D <= 16 ? D : 8)
            for (int i = 0; i < D; ++i)
                if (i != col) W[i * WP + j] = fma(-W[i * WP + col], pj, W[i * WP + j]);
        }
        gsync<G::WG>(g);
    }
}

// out = 0.5 (X + X^T) + S   (out distinct from X)
template <class G> MDEV void sym_add(const Grp& g, const double* X, const double* S, double* out) {
    sweep<G>(g, [&](int i, int j, int) { out[i * G::LD + j] = 0.5 * (X[i * G::LD + j] + X[j * G::LD + i]) + S[i * G::LD + j]; });
}
template <class G> MDEV void gload_mat(const Grp& g, double* dst, const double* src) {
    sweep<G>(g, [&](int i, int j, int idx) { dst[i * G::LD + j] = src[idx]; });
}
template <class G> MDEV void gstore_mat(const Grp& g, double* dst, const double* src) {
    sweep<G>(g, [&](int i, int j, int idx) { dst[idx] = src[i * G::LD + j]; });
}
// A dense d x d matrix / d vector staged in registers: every global load is issued before the first use.
template <class G> struct MatRegs {
    static constexpr int NQ = (G::DD + G::NT - 1) / G::NT;
    double v[NQ];
    MDEV void fetch(const Grp& g, const double* src) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) v[q] = (g.tid + G::NT * q < G::DD) ? __ldg(src + g.tid + G::NT * q) : 0.0;
    }
    MDEV void put(const Grp& g, double* dst) const {
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const int idx = g.tid + G::NT * q;
            if (idx < G::DD) {
                const int i = idx / G::D, j = idx - i * G::D;
                dst[i * G::LD + j] = v[q];
            }
        }
    }
};
template <class G> struct VecRegs {
    double v;
    MDEV void fetch(const Grp& g, const double* src) { v = g.tid < G::D ? __ldg(src + g.tid) : 0.0; }
    MDEV void put(const Grp& g, double* dst) const {
        if (g.tid < G::D) dst[g.tid] = v;
    }
};
template <class G> MDEV void gload_vec(const Grp& g, double* dst, const double* src) {
    vsweep<G>(g, [&](int i) { dst[i] = src[i]; });
}
template <class G> MDEV void gstore_vec(const Grp& g, double* dst, const double* src) {
    vsweep<G>(g, [&](int i) { dst[i] = src[i]; });
}

// ---------------------------------------------------------------------------------------------------------------
// Filter.  aggregate (smem): A | C | J | b | eta ; state: P | m
// ---------------------------------------------------------------------------------------------------------------
template <int D> struct FilterH {
    using G = Geo<D, hier_wg<D>()>;
    static constexpr int MSZ = G::MSZ, DP = G::DP, LD = G::LD, DD = G::DD;
    static constexpr int AGG = 3 * MSZ + 2 * DP;      // smem doubles
    static constexpr int STATE = MSZ + DP;
    static constexpr int NAGG_G = 3 * DD + 2 * D;     // global doubles (GFilter layout)
    static constexpr int NSTATE_G = D + DD;
    static constexpr int WP = (3 * D + 1) | 1;
    static constexpr int WORK = 5 * MSZ + 3 * DP + ((D * WP + 1) & ~1);  // even: 16-byte alignment of the next warp
    struct Init { const double* P0; const double* m0; };

    struct AggRegs {
        MatRegs<G> A, C, J;
        VecRegs<G> b, eta;
        MDEV void fetch(const Grp& g, const double* src) {
            A.fetch(g, src);
            C.fetch(g, src + DD);
            J.fetch(g, src + 2 * DD);
            b.fetch(g, src + 3 * DD);
            eta.fetch(g, src + 3 * DD + D);
        }
        MDEV void put(const Grp& g, double* a) const {
            A.put(g, a);
            C.put(g, a + MSZ);
            J.put(g, a + 2 * MSZ);
            b.put(g, a + 3 * MSZ);
            eta.put(g, a + 3 * MSZ + DP);
        }
    };
    MDEV static void store_agg(const Grp& g, double* dst, const double* a) {
        gstore_mat<G>(g, dst, a);
        gstore_mat<G>(g, dst + DD, a + MSZ);
        gstore_mat<G>(g, dst + 2 * DD, a + 2 * MSZ);
        gstore_vec<G>(g, dst + 3 * DD, a + 3 * MSZ);
        gstore_vec<G>(g, dst + 3 * DD + D, a + 3 * MSZ + DP);
    }
    MDEV static void load_state(const Grp& g, double* s, const double* src) {  // global: m | P
        gload_vec<G>(g, s + MSZ, src);
        gload_mat<G>(g, s, src + D);
    }
    MDEV static void store_state(const Grp& g, double* dst, const double* s) {
        gstore_vec<G>(g, dst, s + MSZ);
        gstore_mat<G>(g, dst + D, s);
    }
    MDEV static void init_state(const Grp& g, double* s, const Init& in) {
        sweep<G>(g, [&](int i, int j, int idx) { s[i * LD + j] = 0.5 * (in.P0[idx] + in.P0[j * D + i]); });
        vsweep<G>(g, [&](int i) { s[MSZ + i] = in.m0 != nullptr ? in.m0[i] : 0.0; });
    }

    // out = a1 (earlier) o a2 (later); a1, a2, out distinct
    MDEV static void combine(const Grp& g, const double* a1, const double* a2, double* o, double* w) {
        const double *A1 = a1, *C1 = a1 + MSZ, *J1 = a1 + 2 * MSZ, *b1 = a1 + 3 * MSZ, *e1 = b1 + DP;
        const double *A2 = a2, *C2 = a2 + MSZ, *J2 = a2 + 2 * MSZ, *b2 = a2 + 3 * MSZ, *e2 = b2 + DP;
        double *Ao = o, *Co = o + MSZ, *Jo = o + 2 * MSZ, *bo = o + 3 * MSZ, *eo = bo + DP;
        double *T1 = w, *T2 = w + MSZ, *T3 = w + 2 * MSZ, *X1 = w + 3 * MSZ, *X2 = w + 4 * MSZ;
        double *v1 = w + 5 * MSZ, *v2 = v1 + DP, *zb = v2 + DP, *W = zb + DP;
        mm<G, false, false>(g, C1, J2, T1);      // C1 J2
        mm<G, false, true>(g, C1, A2, T2);       // C1 A2^T
        mv<G, false>(g, C1, e2, v1, b1);         // b1 + C1 eta2
        gsync<G::WG>(g);
        sweep<G>(g, [&](int i, int j, int) {
            W[i * WP + j] = T1[i * LD + j] + (i == j ? 1.0 : 0.0);
            W[i * WP + D + j] = A1[i * LD + j];
            W[i * WP + 2 * D + 1 + j] = T2[i * LD + j];
        });
        vsweep<G>(g, [&](int i) { W[i * WP + 2 * D] = v1[i]; });
        gsync<G::WG>(g);
        wsolve<G, 2 * D + 1, WP>(g, W);
        sweep<G>(g, [&](int i, int j, int) {
            T1[i * LD + j] = W[i * WP + D + j];            // ZA = M^-1 A1
            T2[i * LD + j] = W[i * WP + 2 * D + 1 + j];    // ZC = M^-1 C1 A2^T
        });
        vsweep<G>(g, [&](int i) { zb[i] = W[i * WP + 2 * D]; });
        gsync<G::WG>(g);
        mm<G, false, false>(g, A2, T1, Ao);      // A = A2 ZA
        mm<G, false, false>(g, A2, T2, X1);      // A2 ZC
        mm<G, false, false>(g, J2, T1, T3);      // J2 ZA
        mv<G, false>(g, A2, zb, bo, b2);         // b = A2 zb + b2
        mv<G, false>(g, J2, zb, v2);             // J2 zb
        gsync<G::WG>(g);
        vsweep<G>(g, [&](int i) { v1[i] = e2[i] - v2[i]; });
        mm<G, true, false>(g, A1, T3, X2);       // A1^T J2 ZA
        sym_add<G>(g, X1, C2, Co);               // C = sym(A2 ZC) + C2
        gsync<G::WG>(g);
        mv<G, true>(g, A1, v1, eo, e1);          // eta = A1^T (eta2 - J2 zb) + eta1
        sym_add<G>(g, X2, J1, Jo);               // J = sym(A1^T J2 ZA) + J1
        gsync<G::WG>(g);
    }

    // s2 = s o a   (s, s2 distinct)
    MDEV static void apply(const Grp& g, const double* s, const double* a, double* s2, double* w) {
        const double *A = a, *C = a + MSZ, *J = a + 2 * MSZ, *b = a + 3 * MSZ, *eta = b + DP;
        const double *P = s, *m = s + MSZ;
        double *T1 = w, *T2 = w + MSZ, *X1 = w + 3 * MSZ;
        double *v1 = w + 5 * MSZ, *zb = v1 + 2 * DP, *W = zb + DP;
        mm<G, false, false>(g, P, J, T1);        // P J
        mm<G, false, true>(g, P, A, T2);         // P A^T
        mv<G, false>(g, P, eta, v1, m);          // m + P eta
        gsync<G::WG>(g);
        sweep<G>(g, [&](int i, int j, int) {
            W[i * WP + j] = T1[i * LD + j] + (i == j ? 1.0 : 0.0);
            W[i * WP + D + 1 + j] = T2[i * LD + j];
        });
        vsweep<G>(g, [&](int i) { W[i * WP + D] = v1[i]; });
        gsync<G::WG>(g);
        wsolve<G, D + 1, WP>(g, W);
        sweep<G>(g, [&](int i, int j, int) { T2[i * LD + j] = W[i * WP + D + 1 + j]; });
        vsweep<G>(g, [&](int i) { zb[i] = W[i * WP + D]; });
        gsync<G::WG>(g);
        mm<G, false, false>(g, A, T2, X1);       // A Z
        mv<G, false>(g, A, zb, s2 + MSZ, b);     // m' = A z + b
        gsync<G::WG>(g);
        sym_add<G>(g, X1, C, s2);                // P' = sym(A Z) + C
        gsync<G::WG>(g);
    }
};

// ---------------------------------------------------------------------------------------------------------------
// Combined reverse scan (GRev).  aggregate (smem): Ab | Ba | Bm | a ; state: dP | Lam | dm | lam
// ---------------------------------------------------------------------------------------------------------------
template <int D> struct RevH {
    using G = Geo<D, hier_wg<D>()>;
    static constexpr int MSZ = G::MSZ, DP = G::DP, LD = G::LD, DD = G::DD;
    static constexpr int AGG = 3 * MSZ + DP;
    static constexpr int STATE = 2 * MSZ + 2 * DP;
    static constexpr int NAGG_G = 3 * DD + D;
    static constexpr int NSTATE_G = 2 * DD + 2 * D;
    static constexpr int WORK = 4 * MSZ + 2 * DP;
    struct Init { const double* init; };

    struct AggRegs {
        MatRegs<G> A, Ba, Bm;
        VecRegs<G> a;
        MDEV void fetch(const Grp& g, const double* src) {
            A.fetch(g, src);
            Ba.fetch(g, src + DD);
            Bm.fetch(g, src + 2 * DD);
            a.fetch(g, src + 3 * DD);
        }
        MDEV void put(const Grp& g, double* x) const {
            A.put(g, x);
            Ba.put(g, x + MSZ);
            Bm.put(g, x + 2 * MSZ);
            a.put(g, x + 3 * MSZ);
        }
    };
    MDEV static void store_agg(const Grp& g, double* dst, const double* a) {
        gstore_mat<G>(g, dst, a);
        gstore_mat<G>(g, dst + DD, a + MSZ);
        gstore_mat<G>(g, dst + 2 * DD, a + 2 * MSZ);
        gstore_vec<G>(g, dst + 3 * DD, a + 3 * MSZ);
    }
    MDEV static void load_state(const Grp& g, double* s, const double* src) {  // global: dm | lam | dP | Lam
        gload_vec<G>(g, s + 2 * MSZ, src);
        gload_vec<G>(g, s + 2 * MSZ + DP, src + D);
        gload_mat<G>(g, s, src + 2 * D);
        gload_mat<G>(g, s + MSZ, src + 2 * D + DD);
    }
    MDEV static void store_state(const Grp& g, double* dst, const double* s) {
        gstore_vec<G>(g, dst, s + 2 * MSZ);
        gstore_vec<G>(g, dst + D, s + 2 * MSZ + DP);
        gstore_mat<G>(g, dst + 2 * D, s);
        gstore_mat<G>(g, dst + 2 * D + DD, s + MSZ);
    }
    MDEV static void init_state(const Grp& g, double* s, const Init& in) {
        if (in.init != nullptr) load_state(g, s, in.init);
        // else: the shared memory was zero-filled
    }

    // x1 later in time (first in scan order), x2 earlier: out = "x1, then x2"
    MDEV static void combine(const Grp& g, const double* x1, const double* x2, double* o, double* w) {
        const double *A1 = x1, *Ba1 = x1 + MSZ, *Bm1 = x1 + 2 * MSZ, *a1 = x1 + 3 * MSZ;
        const double *A2 = x2, *Ba2 = x2 + MSZ, *Bm2 = x2 + 2 * MSZ, *a2 = x2 + 3 * MSZ;
        double *Ao = o, *Bao = o + MSZ, *Bmo = o + 2 * MSZ, *ao = o + 3 * MSZ;
        double *T1 = w, *T2 = w + MSZ, *X1 = w + 2 * MSZ, *X2 = w + 3 * MSZ, *t = w + 4 * MSZ;
        mm<G, false, false>(g, A1, A2, Ao);
        mm<G, false, false>(g, Ba1, A2, T1);
        mm<G, false, false>(g, Bm1, A2, T2);
        mv<G, true>(g, A2, a1, t);               // A2^T a1
        gsync<G::WG>(g);
        mm<G, true, false>(g, A2, T1, X1);       // A2^T Ba1 A2
        mm<G, true, false>(g, A2, T2, X2);       // A2^T Bm1 A2
        gsync<G::WG>(g);
        sweep<G>(g, [&](int i, int j, int) {
            const int q = i * LD + j;
            Bao[q] = Ba2[q] + 0.5 * (t[i] * a2[j] + t[j] * a2[i]) + 0.5 * (X1[q] + X1[j * LD + i]);
            Bmo[q] = Bm2[q] + 0.5 * (X2[q] + X2[j * LD + i]);
        });
        vsweep<G>(g, [&](int i) { ao[i] = t[i] + a2[i]; });
        gsync<G::WG>(g);
    }

    MDEV static void apply(const Grp& g, const double* s, const double* x, double* s2, double* w) {
        const double *Ab = x, *Ba = x + MSZ, *Bm = x + 2 * MSZ, *a = x + 3 * MSZ;
        const double *dP = s, *Lam = s + MSZ, *dm = s + 2 * MSZ, *lam = dm + DP;
        double *T1 = w, *T2 = w + MSZ, *X1 = w + 2 * MSZ, *X2 = w + 3 * MSZ, *t = w + 4 * MSZ, *tl = t + DP;
        mm<G, false, false>(g, dP, Ab, T1);
        mm<G, false, false>(g, Lam, Ab, T2);
        mv<G, true>(g, Ab, dm, t);
        mv<G, true>(g, Ab, lam, tl);
        gsync<G::WG>(g);
        mm<G, true, false>(g, Ab, T1, X1);
        mm<G, true, false>(g, Ab, T2, X2);
        gsync<G::WG>(g);
        sweep<G>(g, [&](int i, int j, int) {
            const int q = i * LD + j;
            s2[q] = Ba[q] + 0.5 * (t[i] * a[j] + t[j] * a[i]) + 0.5 * (X1[q] + X1[j * LD + i]);
            s2[MSZ + q] = Bm[q] + 0.5 * (X2[q] + X2[j * LD + i]);
        });
        vsweep<G>(g, [&](int i) {
            s2[2 * MSZ + i] = t[i] + a[i];
            s2[2 * MSZ + DP + i] = tl[i] - a[i];
        });
        gsync<G::WG>(g);
    }
};

template <class H> __host__ __device__ constexpr int up_warp_doubles() { return 3 * H::AGG + H::WORK; }
template <class H> __host__ __device__ constexpr int walk_warp_doubles() { return H::AGG + 2 * H::STATE + H::WORK; }
template <class H> __host__ __device__ constexpr int top_warp_doubles() {
    return up_warp_doubles<H>() > walk_warp_doubles<H>() ? up_warp_doubles<H>() : walk_warp_doubles<H>();
}
// groups per CTA
template <class H> __host__ __device__ constexpr int warps_per_cta(int per_warp_doubles) {
    const int fit = (200 * 1024) / (per_warp_doubles * 8);
    const int cap = 32 / H::G::WG < 8 ? 32 / H::G::WG : 8;
    return fit > cap ? cap : (fit < 1 ? 1 : fit);
}

// level l -> l + 1: one warp per group of GF aggregates
template <class H>
__global__ void __launch_bounds__(warps_per_cta<H>(up_warp_doubles<H>()) * H::G::NT) hier_up_kernel(const double* __restrict__ in, long nin, double* __restrict__ out, long nout) {
    using G = typename H::G;
    extern __shared__ __align__(16) double smem[];
    constexpr int PW = up_warp_doubles<H>();
    const int gi = threadIdx.x / G::NT;
    const long grp = (long)blockIdx.x * (blockDim.x / G::NT) + gi;
    if (grp >= nout) return;
    const Grp g = make_grp<G>(gi);
    double* S = smem + (size_t)gi * PW;
    for (int i = g.tid; i < PW; i += G::NT) S[i] = 0.0;
    gsync<G::WG>(g);
    double *a = S, *b = S + H::AGG, *o = S + 2 * H::AGG, *w = S + 3 * H::AGG;
    const long i0 = grp * GF;
    const long i1 = (i0 + GF < nin) ? i0 + GF : nin;
    typename H::AggRegs rg;
    rg.fetch(g, in + i0 * H::NAGG_G);
    rg.put(g, a);
    if (i0 + 1 < i1) rg.fetch(g, in + (i0 + 1) * H::NAGG_G);
#pragma unroll 1
    for (long i = i0 + 1; i < i1; ++i) {
        rg.put(g, b);
        gsync<G::WG>(g);
        if (i + 1 < i1) rg.fetch(g, in + (i + 1) * H::NAGG_G);  // in flight while this pair is combined
        H::combine(g, a, b, o, w);
        double* t = a;
        a = o;
        o = t;
    }
    gsync<G::WG>(g);
    H::store_agg(g, out + grp * H::NAGG_G, a);
}

// top level (single warp): walk the n aggregates from the initial state -> states[i] = state entering aggregate i,
// final_state = state after all; or (summary != nullptr) reduce them to one aggregate.
// `stride` = distance in doubles between consecutive aggregates in scan order (negative: a gathered buffer of shard
// summaries in rank order walked from the last rank down).  states may be nullptr (fold: only final_state wanted).
template <class H>
__global__ void __launch_bounds__(H::G::NT) hier_top_kernel(typename H::Init init, const double* __restrict__ aggs, long n,
                                                       long stride, double* __restrict__ states,
                                                       double* __restrict__ final_state, double* __restrict__ summary) {
    using G = typename H::G;
    extern __shared__ __align__(16) double smem[];
    constexpr int PW = top_warp_doubles<H>();
    const Grp g = make_grp<G>(0);
    for (int i = g.tid; i < PW; i += G::NT) smem[i] = 0.0;
    gsync<G::WG>(g);
    if (summary != nullptr) {
        double *a = smem, *b = smem + H::AGG, *o = smem + 2 * H::AGG, *w = smem + 3 * H::AGG;
        typename H::AggRegs rg;
        rg.fetch(g, aggs);
        rg.put(g, a);
        if (n > 1) rg.fetch(g, aggs + stride);
#pragma unroll 1
        for (long i = 1; i < n; ++i) {
            rg.put(g, b);
            gsync<G::WG>(g);
            if (i + 1 < n) rg.fetch(g, aggs + (i + 1) * stride);
            H::combine(g, a, b, o, w);
            double* t = a;
            a = o;
            o = t;
        }
        gsync<G::WG>(g);
        H::store_agg(g, summary, a);
        return;
    }
    double *a = smem, *s = smem + H::AGG, *s2 = s + H::STATE, *w = s2 + H::STATE;
    H::init_state(g, s, init);
    typename H::AggRegs rg;
    if (n > 0) rg.fetch(g, aggs);
    gsync<G::WG>(g);
#pragma unroll 1
    for (long i = 0; i < n; ++i) {
        if (states != nullptr) H::store_state(g, states + i * H::NSTATE_G, s);
        rg.put(g, a);
        gsync<G::WG>(g);
        if (i + 1 < n) rg.fetch(g, aggs + (i + 1) * stride);
        H::apply(g, s, a, s2, w);
        double* t = s;
        s = s2;
        s2 = t;
    }
    if (final_state != nullptr) H::store_state(g, final_state, s);
}

// level l + 1 -> l: one warp per group turns the group-entry state into the entry state of every member
template <class H>
__global__ void __launch_bounds__(warps_per_cta<H>(walk_warp_doubles<H>()) * H::G::NT) hier_down_kernel(const double* __restrict__ aggs, long n,
                                                         const double* __restrict__ gstates, long ngroups,
                                                         double* __restrict__ states) {
    using G = typename H::G;
    extern __shared__ __align__(16) double smem[];
    constexpr int PW = walk_warp_doubles<H>();
    const int gi = threadIdx.x / G::NT;
    const long grp = (long)blockIdx.x * (blockDim.x / G::NT) + gi;
    if (grp >= ngroups) return;
    const Grp g = make_grp<G>(gi);
    double* S = smem + (size_t)gi * PW;
    for (int i = g.tid; i < PW; i += G::NT) S[i] = 0.0;
    gsync<G::WG>(g);
    double *a = S, *s = S + H::AGG, *s2 = s + H::STATE, *w = s2 + H::STATE;
    const long i0 = grp * GF;
    const long i1 = (i0 + GF < n) ? i0 + GF : n;
    typename H::AggRegs rg;
    if (i0 + 1 < i1) rg.fetch(g, aggs + i0 * H::NAGG_G);
    H::load_state(g, s, gstates + grp * H::NSTATE_G);
    gsync<G::WG>(g);
#pragma unroll 1
    for (long i = i0; i < i1; ++i) {
        H::store_state(g, states + i * H::NSTATE_G, s);
        if (i + 1 < i1) {
            rg.put(g, a);
            gsync<G::WG>(g);
            if (i + 2 < i1) rg.fetch(g, aggs + (i + 1) * H::NAGG_G);
            H::apply(g, s, a, s2, w);
            double* t = s;
            s = s2;
            s2 = t;
        }
    }
}

struct Levels {
    static constexpr int MAXL = 16;
    int64_t cnt[MAXL];
    size_t off[MAXL];
    int nl;
    size_t tot;
};
inline Levels levels(int64_t cnt0) {
    Levels hl;
    hl.nl = 0;
    hl.cnt[hl.nl++] = cnt0;
    while (hl.cnt[hl.nl - 1] > TOPMAX && hl.nl < Levels::MAXL) {
        hl.cnt[hl.nl] = (hl.cnt[hl.nl - 1] + GF - 1) / GF;
        ++hl.nl;
    }
    hl.tot = 0;
    for (int l = 0; l < hl.nl; ++l) {
        hl.off[l] = hl.tot;
        hl.tot += (size_t)hl.cnt[l];
    }
    return hl;
}
inline size_t total(int64_t cnt0) { return levels(cnt0).tot; }

template <class H>
int run(pssgp_handle* h, const typename H::Init& init, int64_t cnt0, double* aggs, double* states, double* final_state,
        double* summary, bool have_up, const char* const (&names)[3], cudaStream_t st, int* launches) {
    constexpr int PWU = up_warp_doubles<H>(), PWD = walk_warp_doubles<H>(), PWT = top_warp_doubles<H>();
    constexpr int WU = warps_per_cta<H>(PWU), WD = warps_per_cta<H>(PWD);
    static_assert((size_t)PWT * 8 <= 227 * 1024, "state dimension too large for the warp-level hierarchy");
    cudaError_t e = cudaFuncSetAttribute(hier_up_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, WU * PWU * 8);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(hier_top_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, PWT * 8);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(hier_down_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, WD * PWD * 8);
    if (e != cudaSuccess) return set_err(PSSGP_ERR_CUDA, "cudaFuncSetAttribute(hierarchy): %s", cudaGetErrorString(e));
    const Levels hl = levels(cnt0);
    const int nl = hl.nl;
    constexpr int NA = H::NAGG_G, NS = H::NSTATE_G;
    if (!have_up) {
        for (int l = 0; l + 1 < nl; ++l) {
            const unsigned grid = (unsigned)((hl.cnt[l + 1] + WU - 1) / WU);
            PSSGP_LAUNCH(h, names[0], st,
                         (hier_up_kernel<H><<<grid, WU * H::G::NT, (size_t)WU * PWU * 8, st>>>(aggs + hl.off[l] * NA, hl.cnt[l],
                                                                                        aggs + hl.off[l + 1] * NA, hl.cnt[l + 1])));
            ++*launches;
        }
    }
    if (summary != nullptr) {
        PSSGP_LAUNCH(h, names[1], st,
                     (hier_top_kernel<H><<<1, H::G::NT, (size_t)PWT * 8, st>>>(init, aggs + hl.off[nl - 1] * NA, hl.cnt[nl - 1], (long)NA,
                                                                        nullptr, nullptr, summary)));
        ++*launches;
        return PSSGP_OK;
    }
    PSSGP_LAUNCH(h, names[1], st,
                 (hier_top_kernel<H><<<1, H::G::NT, (size_t)PWT * 8, st>>>(init, aggs + hl.off[nl - 1] * NA, hl.cnt[nl - 1], (long)NA,
                                                                    states + hl.off[nl - 1] * NS, final_state, nullptr)));
    ++*launches;
    for (int l = nl - 2; l >= 0; --l) {
        const unsigned grid = (unsigned)((hl.cnt[l + 1] + WD - 1) / WD);
        PSSGP_LAUNCH(h, names[2], st,
                     (hier_down_kernel<H><<<grid, WD * H::G::NT, (size_t)WD * PWD * 8, st>>>(aggs + hl.off[l] * NA, hl.cnt[l],
                                                                                      states + hl.off[l + 1] * NS, hl.cnt[l + 1],
                                                                                      states + hl.off[l] * NS)));
        ++*launches;
    }
    return PSSGP_OK;
}

// state_out = init o summaries[0] o summaries[stride] o ... (count of them): fold of gathered shard summaries
template <class H>
int fold(pssgp_handle* h, const typename H::Init& init, const double* summaries, int count, long stride, double* state_out,
         const char* name, cudaStream_t st) {
    constexpr int PWT = top_warp_doubles<H>();
    cudaError_t e = cudaFuncSetAttribute(hier_top_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, PWT * 8);
    if (e != cudaSuccess) return set_err(PSSGP_ERR_CUDA, "cudaFuncSetAttribute(hierarchy): %s", cudaGetErrorString(e));
    PSSGP_LAUNCH(h, name, st,
                 (hier_top_kernel<H><<<1, H::G::NT, (size_t)PWT * 8, st>>>(init, summaries, (long)count, stride, nullptr, state_out,
                                                                    nullptr)));
    return PSSGP_OK;
}

}  // namespace hier
}  // namespace mid
}  // namespace pssgp
