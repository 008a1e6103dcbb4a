// Fragment-resident variant of the warp-level d > 4 kernels (mid.cuh) for D <= 16: one warp per chunk, and the
// d x d state of the recursion never leaves the register file.
//
// Every matrix is held in the accumulator layout of mma.m8n8k4.f64 ("CF": lane (r, c) = (lane >> 2, lane & 3) holds
// M[8 mt + r][8 nt + 2c + {0, 1}] of tile (mt, nt)).  With the k-slot of lane c standing for the physical columns
// k = 2c + s of k-step s, the accumulator layout IS the A-operand layout, and it is also the B-operand layout of the
// TRANSPOSE: the product  X * S^T  of two CF matrices needs no data movement at all.  All recursions are arranged in
// that form (symmetric matrices are their own transpose; A^T is tracked instead of A; F is loaded in both
// orientations from the shared-memory ring the cp.async copies fill), rank-one / rank-two / rank-three updates are
// single DMMAs whose operands are vectors ("VR": lane (r, .) holds v[8t + r]), matrix-vector products reduce over the
// four lanes of a row with two shuffles.  Shared memory only holds the ring of input rows (tile-major, 64 doubles
// per 8 x 8 tile: every fragment load is a conflict-free 128-bit access).
//
// Same mathematics, arguments and workspace layouts as the kernels of mid.cuh (see there for the reference lines).
#pragma once
#include <type_traits>

#include "mid.cuh"

namespace pssgp {
namespace mid {
namespace frag {

template <int D_> struct FGeo {
    static constexpr int D = D_;
    static constexpr int DP = (D + 7) / 8 * 8;
    static constexpr int MT = DP / 8;
    static constexpr int MSZ = DP * DP;  // doubles per matrix copy in a ring slot (tile-major)
    static constexpr int DD = D * D;
};

template <int MT> struct Mat { double v[MT][MT][2]; };
template <int MT> struct VecR { double v[MT]; };
template <int MT> struct VecP { double v[MT][2]; };

constexpr unsigned FULL = 0xffffffffu;

template <int MT> MDEV void mzero(Mat<MT>& m) {
#pragma unroll
    for (int a = 0; a < MT; ++a)
#pragma unroll
        for (int b = 0; b < MT; ++b) m.v[a][b][0] = m.v[a][b][1] = 0.0;
}

// Operand of a product that still sits in a ring slot (tile-major): fragments are loaded inside the product loop, so
// an input matrix never occupies registers for longer than one DMMA (MT = 3: a matrix is 36 registers per thread).
template <int MT> struct SMat {
    const double* p;
    int r, c;
    bool tr;  // fragments of the TRANSPOSE of the stored matrix
    MDEV double get(int a, int b, int s) const {
        return tr ? p[(b * MT + a) * 64 + (2 * c + s) * 8 + r] : p[(a * MT + b) * 64 + r * 8 + 2 * c + s];
    }
};
template <int MT> MDEV double oget(const Mat<MT>& m, int a, int b, int s) { return m.v[a][b][s]; }
template <int MT> MDEV double oget(const SMat<MT>& m, int a, int b, int s) { return m.get(a, b, s); }

// acc += X S^T.  KD = number of valid columns of X and S (the state dimension): k-steps whose four columns
// 8 kt + 2c + s all lie in the zero padding are skipped (d = 9: 3 DMMAs per output tile instead of 4).
template <int MT, int KD, class XO, class SO> MDEV void mmT(Mat<MT>& acc, const XO& X, const SO& S) {
#pragma unroll
    for (int kt = 0; kt < MT; ++kt)
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            if (8 * kt + s >= KD) continue;
            double xa[MT], sb[MT];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) xa[mt] = oget<MT>(X, mt, kt, s);
#pragma unroll
            for (int nt = 0; nt < MT; ++nt) sb[nt] = oget<MT>(S, nt, kt, s);
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int nt = 0; nt < MT; ++nt) dmma(acc.v[mt][nt], xa[mt], sb[nt]);
        }
}
template <int MT, int KD, class XO, class SO> MDEV Mat<MT> mulT(const XO& X, const SO& S) {
    Mat<MT> acc;
    mzero(acc);
    mmT<MT, KD>(acc, X, S);
    return acc;
}

// acc += sum_{q < NQ} x[q] y[q]^T   (NQ <= 4 rank-one terms in ONE DMMA per tile: term q rides in k-slot q)
template <int MT, int NQ> MDEV void rank_update(Mat<MT>& acc, const VecR<MT> (&x)[NQ], const VecR<MT> (&y)[NQ], int c) {
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
        double a = 0.0;
#pragma unroll
        for (int q = 0; q < NQ; ++q) a = (c == q) ? x[q].v[mt] : a;
#pragma unroll
        for (int nt = 0; nt < MT; ++nt) {
            double b = 0.0;
#pragma unroll
            for (int q = 0; q < NQ; ++q) b = (c == q) ? y[q].v[nt] : b;
            dmma(acc.v[mt][nt], a, b);
        }
    }
}

// y = M v : M in CF (registers or ring slot), v in VP -> VR
template <int MT, class MO> MDEV VecR<MT> mv(const MO& M, const VecP<MT>& v) {
    VecR<MT> y;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
        double s = 0.0;
#pragma unroll
        for (int nt = 0; nt < MT; ++nt) {
            s = fma(oget<MT>(M, mt, nt, 0), v.v[nt][0], s);
            s = fma(oget<MT>(M, mt, nt, 1), v.v[nt][1], s);
        }
        s += __shfl_xor_sync(FULL, s, 1);
        s += __shfl_xor_sync(FULL, s, 2);
        y.v[mt] = s;
    }
    return y;
}

template <int MT> MDEV VecP<MT> vr2vp(const VecR<MT>& x, int c) {
    VecP<MT> y;
#pragma unroll
    for (int t = 0; t < MT; ++t) {
        y.v[t][0] = __shfl_sync(FULL, x.v[t], (2 * c) * 4);
        y.v[t][1] = __shfl_sync(FULL, x.v[t], (2 * c + 1) * 4);
    }
    return y;
}

// NQ dot products of VR vectors at once (every lane gets all results)
template <int MT, int NQ> MDEV void dots(const VecR<MT>* const (&x)[NQ], const VecR<MT>* const (&y)[NQ], double (&out)[NQ]) {
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        double p = 0.0;
#pragma unroll
        for (int t = 0; t < MT; ++t) p = fma(x[q]->v[t], y[q]->v[t], p);
        out[q] = p;
    }
#pragma unroll
    for (int off = 4; off < 32; off <<= 1) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) out[q] += __shfl_xor_sync(FULL, out[q], off);
    }
}

// ---- shared-memory ring slots (tile-major) -----------------------------------------------------------------
// Per-lane plan of the cp.async copy of one dense row-major D x D matrix into a tile-major slot: computed once,
// so that issuing a matrix costs one LDGSTS (+ one address add) per piece.  D even: 16-byte pieces (a pair of
// consecutive columns never straddles a row), else 8-byte pieces.
template <int D, int MT> struct CpPlan {
    static constexpr int PB = (D % 2 == 0) ? 2 : 1;           // doubles per piece
    static constexpr int NPIECE = D * D / PB;
    static constexpr int NQ = (NPIECE + 31) / 32;
    int soff[NQ], doff[NQ];
    MDEV void init(int lane) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const int idx = (lane + 32 * q) * PB;
            const int i = idx / D, j = idx - i * D;
            soff[q] = idx;
            doff[q] = ((i >> 3) * MT + (j >> 3)) * 64 + (i & 7) * 8 + (j & 7);
        }
    }
    MDEV void issue(int lane, double* dst, const double* src) const {
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            if (32 * (q + 1) <= NPIECE || lane + 32 * q < NPIECE) {
                if constexpr (PB == 2)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(
                                     (unsigned)__cvta_generic_to_shared(dst + doff[q])),
                                 "l"(src + soff[q])
                                 : "memory");
                else
                    cp8(dst + doff[q], src + soff[q]);
            }
        }
    }
};
template <int D> MDEV void issue_vec(int lane, double* dst, const double* src) {
    if (lane < D) cp8(dst + lane, src + lane);
}
template <int MT> MDEV Mat<MT> ld_mat(const double* slot, int r, int c) {
    Mat<MT> m;
#pragma unroll
    for (int a = 0; a < MT; ++a)
#pragma unroll
        for (int b = 0; b < MT; ++b) {
            const double2 v = *reinterpret_cast<const double2*>(slot + (a * MT + b) * 64 + r * 8 + 2 * c);
            m.v[a][b][0] = v.x;
            m.v[a][b][1] = v.y;
        }
    return m;
}
// CF of the TRANSPOSE of the matrix held in the slot (two strided 64-bit loads per tile)
template <int MT> MDEV Mat<MT> ld_matT(const double* slot, int r, int c) {
    Mat<MT> m;
#pragma unroll
    for (int a = 0; a < MT; ++a)
#pragma unroll
        for (int b = 0; b < MT; ++b) {
            m.v[a][b][0] = slot[(b * MT + a) * 64 + (2 * c) * 8 + r];
            m.v[a][b][1] = slot[(b * MT + a) * 64 + (2 * c + 1) * 8 + r];
        }
    return m;
}
// symmetrised load: 0.5 (M + M^T)
template <int MT> MDEV Mat<MT> ld_sym(const double* slot, int r, int c) {
    Mat<MT> m = ld_mat<MT>(slot, r, c);
    const Mat<MT> t = ld_matT<MT>(slot, r, c);
#pragma unroll
    for (int a = 0; a < MT; ++a)
#pragma unroll
        for (int b = 0; b < MT; ++b) {
            m.v[a][b][0] = 0.5 * (m.v[a][b][0] + t.v[a][b][0]);
            m.v[a][b][1] = 0.5 * (m.v[a][b][1] + t.v[a][b][1]);
        }
    return m;
}
// input operand: held in registers when a matrix is small (MT <= 2), read from the slot on use otherwise
template <int MT, bool LAZY = (MT >= 3)> struct InOp;
template <int MT> struct InOp<MT, false> {
    using type = Mat<MT>;
    MDEV static type make(const double* slot, bool tr, int r, int c) { return tr ? ld_matT<MT>(slot, r, c) : ld_mat<MT>(slot, r, c); }
};
template <int MT> struct InOp<MT, true> {
    using type = SMat<MT>;
    MDEV static type make(const double* slot, bool tr, int r, int c) { return SMat<MT>{slot, r, c, tr}; }
};
template <int MT> MDEV typename InOp<MT>::type in_op(const double* slot, bool tr, int r, int c) { return InOp<MT>::make(slot, tr, r, c); }

template <int MT> MDEV VecR<MT> ld_vr(const double* v, int r) {
    VecR<MT> y;
#pragma unroll
    for (int t = 0; t < MT; ++t) y.v[t] = v[8 * t + r];
    return y;
}
template <int MT> MDEV VecP<MT> ld_vp(const double* v, int c) {
    VecP<MT> y;
#pragma unroll
    for (int t = 0; t < MT; ++t) {
        const double2 d2 = *reinterpret_cast<const double2*>(v + 8 * t + 2 * c);
        y.v[t][0] = d2.x;
        y.v[t][1] = d2.y;
    }
    return y;
}

// ---- global memory (dense row-major d x d / d) ---------------------------------------------------------------
template <int D, int MT> MDEV Mat<MT> gld_mat(const double* src, int r, int c) {
    Mat<MT> m;
#pragma unroll
    for (int a = 0; a < MT; ++a)
#pragma unroll
        for (int b = 0; b < MT; ++b)
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                const int i = 8 * a + r, j = 8 * b + 2 * c + s;
                m.v[a][b][s] = (i < D && j < D) ? src[i * D + j] : 0.0;
            }
    return m;
}
template <int D, int MT> MDEV VecR<MT> gld_vr(const double* src, int r) {
    VecR<MT> y;
#pragma unroll
    for (int t = 0; t < MT; ++t) y.v[t] = (src != nullptr && 8 * t + r < D) ? src[8 * t + r] : 0.0;
    return y;
}
template <int D, int MT> MDEV VecP<MT> gld_vp(const double* src, int c) {
    VecP<MT> y;
#pragma unroll
    for (int t = 0; t < MT; ++t)
#pragma unroll
        for (int s = 0; s < 2; ++s) y.v[t][s] = (src != nullptr && 8 * t + 2 * c + s < D) ? src[8 * t + 2 * c + s] : 0.0;
    return y;
}
// Per-lane offsets / validity of the CF elements in a dense row-major D x D matrix (computed once per kernel).
template <int D, int MT> struct StPlan {
    int off;            // r * D + 2c
    int offT;           // (2c) * D + r
    bool rok[MT], cok[MT][2];
    MDEV void init(int r, int c) {
        off = r * D + 2 * c;
        offT = 2 * c * D + r;
#pragma unroll
        for (int a = 0; a < MT; ++a) {
            rok[a] = 8 * a + r < D;
            cok[a][0] = 8 * a + 2 * c < D;
            cok[a][1] = 8 * a + 2 * c + 1 < D;
        }
    }
};
// dst = scale * M  (TRANS: dst = scale * M^T)
template <int D, int MT, bool TRANS = false>
MDEV void gst_mat(double* dst, const Mat<MT>& m, double scale, const StPlan<D, MT>& sp) {
#pragma unroll
    for (int a = 0; a < MT; ++a)
#pragma unroll
        for (int b = 0; b < MT; ++b) {
            if (TRANS) {
                double* q = dst + sp.offT + 8 * b * D + 8 * a;
                if (sp.rok[a] && sp.cok[b][0]) q[0] = scale * m.v[a][b][0];
                if (sp.rok[a] && sp.cok[b][1]) q[D] = scale * m.v[a][b][1];
            } else {
                double* q = dst + sp.off + 8 * a * D + 8 * b;
                if (D % 2 == 0) {
                    if (sp.rok[a] && sp.cok[b][0])
                        __stcs(reinterpret_cast<double2*>(q), make_double2(scale * m.v[a][b][0], scale * m.v[a][b][1]));
                } else {
                    if (sp.rok[a] && sp.cok[b][0]) __stcs(q, scale * m.v[a][b][0]);
                    if (sp.rok[a] && sp.cok[b][1]) __stcs(q + 1, scale * m.v[a][b][1]);
                }
            }
        }
}
template <int D, int MT> MDEV void gst_vr(double* dst, const VecR<MT>& v, int r, int c) {
    if (c == 0) {
#pragma unroll
        for (int t = 0; t < MT; ++t)
            if (8 * t + r < D) dst[8 * t + r] = v.v[t];
    }
}

// 4 dot products x[q] . y[q] of VR vectors with 3 + 4 shuffles: lane (r, c) carries the partial sum of product c
template <int MT> MDEV void dots4(const VecR<MT>* const (&x)[4], const VecR<MT>* const (&y)[4], int lane, int c, double (&out)[4]) {
    double p = 0.0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < MT; ++k) t = fma(x[q]->v[k], y[q]->v[k], t);
        p = (c == q) ? t : p;
    }
    p += __shfl_xor_sync(FULL, p, 4);
    p += __shfl_xor_sync(FULL, p, 8);
    p += __shfl_xor_sync(FULL, p, 16);
#pragma unroll
    for (int q = 0; q < 4; ++q) out[q] = __shfl_sync(FULL, p, (lane & ~3) | q);
}
// 2 dot products with 3 + 2 shuffles
template <int MT> MDEV void dots2(const VecR<MT>& x0, const VecR<MT>& y0, const VecR<MT>& x1, const VecR<MT>& y1, int lane, int c,
                                 double& o0, double& o1) {
    double t0 = 0.0, t1 = 0.0;
#pragma unroll
    for (int k = 0; k < MT; ++k) {
        t0 = fma(x0.v[k], y0.v[k], t0);
        t1 = fma(x1.v[k], y1.v[k], t1);
    }
    double p = (c & 1) ? t1 : t0;
    p += __shfl_xor_sync(FULL, p, 4);
    p += __shfl_xor_sync(FULL, p, 8);
    p += __shfl_xor_sync(FULL, p, 16);
    o0 = __shfl_sync(FULL, p, lane & ~3);
    o1 = __shfl_sync(FULL, p, (lane & ~3) | 1);
}

template <int MT> MDEV Mat<MT> identity_cf(int r, int c) {
    Mat<MT> m;
    mzero(m);
#pragma unroll
    for (int t = 0; t < MT; ++t)
#pragma unroll
        for (int s = 0; s < 2; ++s)
            if (r == 2 * c + s) m.v[t][t][s] = 1.0;
    return m;
}
// identity restricted to the leading D x D block
template <int D, int MT> MDEV Mat<MT> identity_d(int r, int c) {
    Mat<MT> m;
    mzero(m);
#pragma unroll
    for (int t = 0; t < MT; ++t)
#pragma unroll
        for (int s = 0; s < 2; ++s)
            if (r == 2 * c + s && 8 * t + r < D) m.v[t][t][s] = 1.0;
    return m;
}

// running sum of log(s) with one log per several factors
struct LogAcc {
    double prod, sum;
    MDEV void init() { prod = 1.0; sum = 0.0; }
    MDEV void mul(double s) {
        prod *= s;
        if (prod > 1e100 || prod < 1e-100) flush();
    }
    MDEV void flush() {
        sum += log(prod);
        prod = 1.0;
    }
};

template <int MT> MDEV VecR<MT> vaxpy(double a, const VecR<MT>& x, const VecR<MT>& y) {  // a x + y
    VecR<MT> o;
#pragma unroll
    for (int t = 0; t < MT; ++t) o.v[t] = fma(a, x.v[t], y.v[t]);
    return o;
}
template <int MT> MDEV VecR<MT> vscale(double a, const VecR<MT>& x) {
    VecR<MT> o;
#pragma unroll
    for (int t = 0; t < MT; ++t) o.v[t] = a * x.v[t];
    return o;
}
template <int MT> MDEV Mat<MT> msub(const Mat<MT>& a, const Mat<MT>& b) {
    Mat<MT> o;
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < MT; ++j) {
            o.v[i][j][0] = a.v[i][j][0] - b.v[i][j][0];
            o.v[i][j][1] = a.v[i][j][1] - b.v[i][j][1];
        }
    return o;
}

template <int D> constexpr int frag_nslot() { return D <= 8 ? 4 : 3; }
// warps per CTA: as many as the ring slots of a CTA fit in shared memory, at most `cap` (the launch bound; the option
// "mid_warps" lowers it at run time)
template <int D> constexpr int frag_warps(int slot_doubles, int cap) {
    const int fit = (200 * 1024) / (frag_nslot<D>() * slot_doubles * 8);
    return fit > cap ? cap : (fit < 1 ? 1 : fit);
}
#ifndef PSSGP_FRAG_CAP_SMALL
#define PSSGP_FRAG_CAP_SMALL 24
#endif

// ------------------------------------------------------------------------------------------------
// K1: chunk aggregates of the filter.  Tracks At = A^T, C, J (CF), b, eta (VR).
// ------------------------------------------------------------------------------------------------
template <int D> struct FK1 {
    using G = FGeo<D>;
    static constexpr int NSLOT = frag_nslot<D>();
    static constexpr int SLOT = 2 * G::MSZ;  // F | Q
    static constexpr int WPC = frag_warps<D>(SLOT, D <= 8 ? PSSGP_FRAG_CAP_SMALL : 12);
    static constexpr size_t WARP_SMEM = (size_t)NSLOT * SLOT * 8;
    static constexpr size_t SMEM = (size_t)WPC * WARP_SMEM;
};

template <int D>
__global__ void __launch_bounds__(FK1<D>::WPC * 32) fk1_filter_reduce(Params p, int L, long nchunks, double* __restrict__ aggs) {
    using K = FK1<D>;
    using G = typename K::G;
    constexpr int MT = G::MT, MSZ = G::MSZ, NSLOT = K::NSLOT, DD = G::DD;
    extern __shared__ __align__(16) double smem[];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31, r = lane >> 2, c = lane & 3;
    const long chunk = (long)blockIdx.x * (blockDim.x >> 5) + wid;
    if (chunk >= nchunks) return;
    double* ring = smem + (size_t)wid * NSLOT * K::SLOT;
    for (int i = lane; i < NSLOT * K::SLOT; i += 32) ring[i] = 0.0;
    __syncwarp();
    const long k_lo = chunk * (long)L;
    const long k_hi = (k_lo + L < p.n) ? k_lo + L : p.n;
    const int nrows = (int)(k_hi - k_lo);
    constexpr int PD = NSLOT - 1;
    CpPlan<D, MT> cp;
    cp.init(lane);
    StPlan<D, MT> sp;
    sp.init(r, c);
    auto issue = [&](int i) {
        if (i < nrows) {
            double* sl = ring + (i % NSLOT) * K::SLOT;
            cp.issue(lane, sl, p.Fs + (k_lo + i) * DD);
            cp.issue(lane, sl + MSZ, p.Qs + (k_lo + i) * DD);
        }
        cp_commit();
    };
#pragma unroll 1
    for (int i = 0; i < PD; ++i) issue(i);
    const VecP<MT> hP = gld_vp<D, MT>(p.H, c);
    const VecR<MT> hR = gld_vr<D, MT>(p.H, r);
    const double Rv = p.R[0];
    Mat<MT> At = identity_d<D, MT>(r, c), C, J;
    mzero(C);
    mzero(J);
    VecR<MT> b, eta;
#pragma unroll
    for (int t = 0; t < MT; ++t) b.v[t] = eta.v[t] = 0.0;
    double ynext = p.y[k_lo];
#pragma unroll 1
    for (int i = 0; i < nrows; ++i) {
        const long k = k_lo + i;
        issue(i + PD);
        cp_wait<PD>();
        __syncwarp();
        const double yk = ynext;
        if (i + 1 < nrows) ynext = p.y[k + 1];
        const double* sl = ring + (i % NSLOT) * K::SLOT;
        // the first step of the global series is an update without propagation (parallel.py:24-30)
        if (!(k == 0 && p.first_special)) {
            const auto F = in_op<MT>(sl, false, r, c);
            Mat<MT> Cn = ld_sym<MT>(sl + MSZ, r, c);
            At = mulT<MT, D>(At, F);                   // (F A)^T = A^T F^T
            {
                const Mat<MT> T2 = mulT<MT, D>(F, C);  // F C   (C symmetric)
                mmT<MT, D>(Cn, T2, F);                 // F C F^T + Q
            }
            C = Cn;
            b = mv(F, vr2vp(b, c));
        }
        const bool obs = !isnan(yk);
        const VecR<MT> u = mv(C, hP);
        const VecR<MT> w = mv(At, hP);      // A^T h
        double hu, hb;
        dots2<MT>(hR, u, hR, b, lane, c, hu, hb);
        const double is = obs ? 1.0 / (Rv + hu) : 0.0;
        const double eis = obs ? (yk - hb) * is : 0.0;
        {
            const VecR<MT> x1[1] = {vscale(is, w)}, y1[1] = {w};
            rank_update<MT, 1>(J, x1, y1, c);   // J += w w^T / s
        }
        {
            const VecR<MT> x1[1] = {vscale(-is, w)}, y1[1] = {u};
            rank_update<MT, 1>(At, x1, y1, c);  // A -= u w^T / s
        }
        {
            const VecR<MT> x1[1] = {vscale(-is, u)}, y1[1] = {u};
            rank_update<MT, 1>(C, x1, y1, c);   // C -= u u^T / s
        }
        eta = vaxpy(eis, w, eta);
        b = vaxpy(eis, u, b);
        __syncwarp();  // every lane has read the slot: it may be refilled
    }
    cp_wait<0>();
    double* out = aggs + chunk * (3 * DD + 2 * D);
    gst_mat<D, MT, true>(out, At, 1.0, sp);
    gst_mat<D, MT>(out + DD, C, 1.0, sp);
    gst_mat<D, MT>(out + 2 * DD, J, 1.0, sp);
    gst_vr<D, MT>(out + 3 * DD, b, r, c);
    gst_vr<D, MT>(out + 3 * DD + D, eta, r, c);
}

// ------------------------------------------------------------------------------------------------
// K2: seeded filter recursion (+ log-likelihood) and, with REV, the chunk aggregate of the combined reverse scan
// (tracked as Abt = Abar^T, Ba, Bm, a).  STORED: filtered moments are read instead of recomputed.
// ------------------------------------------------------------------------------------------------
template <int D, bool REV, bool STORED> struct FK2 {
    using G = FGeo<D>;
    static constexpr int NSLOT = frag_nslot<D>();
    static constexpr int NM = 2 + (STORED ? 1 : 0);  // F | Q | P_{k-1}
    static constexpr int SLOT = NM * G::MSZ + (STORED ? G::DP : 0);
    static constexpr int WPC = frag_warps<D>(SLOT, D <= 8 ? PSSGP_FRAG_CAP_SMALL : 12);
    static constexpr size_t WARP_SMEM = (size_t)NSLOT * SLOT * 8;
    static constexpr size_t SMEM = (size_t)WPC * WARP_SMEM;
};

template <int D, bool REV, bool STORED>
__global__ void __launch_bounds__(FK2<D, REV, STORED>::WPC * 32)
fk2_forward(Params p, int L, long nchunks, const double* __restrict__ fstates, double* __restrict__ part,
            double* __restrict__ raggs) {
    using K = FK2<D, REV, STORED>;
    using G = typename K::G;
    constexpr int MT = G::MT, MSZ = G::MSZ, NSLOT = K::NSLOT, DD = G::DD;
    constexpr int O_P = 2 * MSZ, O_M = 3 * MSZ;
    extern __shared__ __align__(16) double smem[];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31, r = lane >> 2, c = lane & 3;
    const long chunk = (long)blockIdx.x * (blockDim.x >> 5) + wid;
    if (chunk >= nchunks) return;
    double* ring = smem + (size_t)wid * NSLOT * K::SLOT;
    for (int i = lane; i < NSLOT * K::SLOT; i += 32) ring[i] = 0.0;
    __syncwarp();
    const long k_lo = chunk * (long)L;
    const long k_hi = (k_lo + L < p.n) ? k_lo + L : p.n;
    const int nrows = (int)(k_hi - k_lo);
    constexpr int PD = NSLOT - 1;
    CpPlan<D, MT> cp;
    cp.init(lane);
    StPlan<D, MT> sp;
    sp.init(r, c);
    auto issue = [&](int i) {
        if (i < nrows) {
            double* sl = ring + (i % NSLOT) * K::SLOT;
            const long k = k_lo + i;
            cp.issue(lane, sl, p.Fs + k * DD);
            cp.issue(lane, sl + MSZ, p.Qs + k * DD);
            if constexpr (STORED) {
                cp.issue(lane, sl + O_P, k > 0 ? p.fPs_in + (k - 1) * DD : p.P0);
                if (k > 0) issue_vec<D>(lane, sl + O_M, p.fms_in + (k - 1) * D);
                else if (p.m0 != nullptr) issue_vec<D>(lane, sl + O_M, p.m0);
                else if (lane < D) sl[O_M + lane] = 0.0;
            }
        }
        cp_commit();
    };
#pragma unroll 1
    for (int i = 0; i < PD; ++i) issue(i);
    const VecP<MT> hP = gld_vp<D, MT>(p.H, c);
    const VecR<MT> hR = gld_vr<D, MT>(p.H, r);
    const double Rv = p.R[0];
    Mat<MT> P;
    VecR<MT> m;
    if constexpr (!STORED) {
        const double* st = fstates + chunk * (D + DD);
        m = gld_vr<D, MT>(st, r);
        P = gld_mat<D, MT>(st + D, r, c);
    } else {
        mzero(P);
#pragma unroll
        for (int t = 0; t < MT; ++t) m.v[t] = 0.0;
    }
    Mat<MT> Abt, Ba, Bm;
    VecR<MT> av;
    if constexpr (REV) {
        Abt = identity_d<D, MT>(r, c);
        mzero(Ba);
        mzero(Bm);
#pragma unroll
        for (int t = 0; t < MT; ++t) av.v[t] = 0.0;
    }
    double ynext = p.y[k_lo];
    double quad = 0.0;
    int nobs = 0;
    LogAcc lacc;
    lacc.init();
    // one time step; FIRST (compile-time) = step 0 of the global series, peeled so that the steady-state loop has no
    // data-dependent branch around warp shuffles
    auto body = [&](int i, auto first_tag) {
        constexpr bool first = decltype(first_tag)::value;
        const long k = k_lo + i;
        issue(i + PD);
        cp_wait<PD>();
        __syncwarp();
        const double yk = ynext;
        if (i + 1 < nrows) ynext = p.y[k + 1];
        const bool obs = !isnan(yk);
        const double* sl = ring + (i % NSLOT) * K::SLOT;
        const auto F = in_op<MT>(sl, false, r, c);
        if constexpr (STORED) {
            P = ld_mat<MT>(sl + O_P, r, c);
            m = ld_vr<MT>(sl + O_M, r);
        }
        Mat<MT> Pp = ld_sym<MT>(sl + MSZ, r, c);
        {
            const Mat<MT> T1 = mulT<MT, D>(F, P);  // F P (P symmetric)
            mmT<MT, D>(Pp, T1, F);                 // F P F^T + Q
        }
        VecR<MT> mp = mv(F, vr2vp(m, c));
        VecR<MT> u = mv(Pp, hP);
        double hu, hm;
        dots2<MT>(hR, u, hR, mp, lane, c, hu, hm);
        double s = Rv + hu;
        double e = obs ? yk - hm : 0.0;
        if constexpr (!STORED) {
            if (obs) {
                lacc.mul(s);
                quad = fma(e, e / s, quad);
                ++nobs;
            }
        }
        if constexpr (first) {
            // parallel.py:24-30: the first update is made on (m0, P0) directly, without prediction
            Pp = P;
            mp = m;
            u = mv(Pp, hP);
            dots2<MT>(hR, u, hR, mp, lane, c, hu, hm);
            s = Rv + hu;
            e = obs ? yk - hm : 0.0;
        }
        const double is = obs ? 1.0 / s : 0.0;
        const double eis = e * is;
        if constexpr (!STORED) {
            P = Pp;
            {
                const VecR<MT> x1[1] = {vscale(-is, u)}, y1[1] = {u};
                rank_update<MT, 1>(P, x1, y1, c);
            }
            m = vaxpy(eis, u, mp);
            gst_mat<D, MT>(p.fPs + k * DD, P, 1.0, sp);
            gst_vr<D, MT>(p.fms + k * D, m, r, c);
        }
        if constexpr (REV) {
            // append step k on the later side of the chunk's reverse aggregate (nothing for the global first step)
            const auto Ft = in_op<MT>(sl, true, r, c);
            const VecR<MT> w = mv(Ft, hP);                 // F^T h
            const VecR<MT> t = mv(Abt, vr2vp(w, c));       // Abar_old^T w
            const double isr = first ? 0.0 : is, eisr = first ? 0.0 : eis;
            if (!first) Abt = mulT<MT, D>(Abt, F);                // (F Abar_old)^T
            {
                const VecR<MT> x1[1] = {vscale(-isr, t)}, y1[1] = {u};
                rank_update<MT, 1>(Abt, x1, y1, c);        // Abar = F Abar_old - u t^T / s
            }
            const double beta = 0.5 * (eisr * eisr - isr);
            {
                const VecR<MT> x3[3] = {vscale(beta, t), vscale(0.5 * eisr, t), vscale(0.5 * eisr, av)};
                const VecR<MT> y3[3] = {t, av, t};
                rank_update<MT, 3>(Ba, x3, y3, c);
            }
            {
                const VecR<MT> x1[1] = {vscale(isr, t)}, y1[1] = {t};
                rank_update<MT, 1>(Bm, x1, y1, c);
            }
            av = vaxpy(eisr, t, av);
        }
        __syncwarp();
    };
    int i0 = 0;
    if (k_lo == 0 && p.first_special) {
        body(0, std::true_type{});
        i0 = 1;
    }
#pragma unroll 1
    for (int i = i0; i < nrows; ++i) body(i, std::false_type{});
    cp_wait<0>();
    if constexpr (!STORED) {
        lacc.flush();
        if (part != nullptr && lane == 0)
            part[chunk] = -0.5 * (lacc.sum + quad + nobs * 1.8378770664093454835606594728112);  // log(2 pi)
    }
    if constexpr (REV) {
        double* out = raggs + (nchunks - 1 - chunk) * (3 * DD + D);
        gst_mat<D, MT, true>(out, Abt, 1.0, sp);
        gst_mat<D, MT>(out + DD, Ba, 1.0, sp);
        gst_mat<D, MT>(out + 2 * DD, Bm, 1.0, sp);
        gst_vr<D, MT>(out + 3 * DD, av, r, c);
    }
}

// ------------------------------------------------------------------------------------------------
// K3: reverse pass — smoothed moments (SMOOTH) and / or gradient of the log-likelihood (ADJ).
// ------------------------------------------------------------------------------------------------
template <int D, bool SMOOTH, bool ADJ> struct FK3 {
    using G = FGeo<D>;
    static constexpr int NSLOT = frag_nslot<D>() < 3 ? 3 : frag_nslot<D>();
    static constexpr int SLOT = 3 * G::MSZ + G::DP;  // F | Q | fP | fm
    static constexpr int WPC = frag_warps<D>(SLOT, D <= 8 ? PSSGP_FRAG_CAP_SMALL : 10);
    static constexpr size_t WARP_SMEM = (size_t)NSLOT * SLOT * 8;
    static constexpr size_t SMEM = (size_t)WPC * WARP_SMEM;
};

template <int D, bool SMOOTH, bool ADJ>
__global__ void __launch_bounds__(FK3<D, SMOOTH, ADJ>::WPC * 32)
fk3_reverse(Params p, int L, long nchunks, const double* __restrict__ rstates, double* __restrict__ part) {
    using K = FK3<D, SMOOTH, ADJ>;
    using G = typename K::G;
    constexpr int MT = G::MT, MSZ = G::MSZ, NSLOT = K::NSLOT, DD = G::DD;
    constexpr int O_P = 2 * MSZ, O_M = 3 * MSZ;
    extern __shared__ __align__(16) double smem[];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31, r = lane >> 2, c = lane & 3;
    const long chunk = (long)blockIdx.x * (blockDim.x >> 5) + wid;
    if (chunk >= nchunks) return;
    double* ring = smem + (size_t)wid * NSLOT * K::SLOT;
    for (int i = lane; i < NSLOT * K::SLOT; i += 32) ring[i] = 0.0;
    __syncwarp();
    const long k_lo = chunk * (long)L;
    const long k_hi = (k_lo + L < p.n) ? k_lo + L : p.n;
    const int nrows = (int)(k_hi - k_lo);
    constexpr int PD = NSLOT - 1;
    CpPlan<D, MT> cp;
    cp.init(lane);
    StPlan<D, MT> sp;
    sp.init(r, c);
    auto slot = [&](long row) { return ring + (int)((row + 8L * NSLOT) % NSLOT) * K::SLOT; };
    auto issue = [&](int j) {
        if (j <= nrows) {
            const long row = k_hi - 1 - j;
            double* sl = slot(row);
            if (j < nrows) {
                cp.issue(lane, sl, p.Fs + row * DD);
                cp.issue(lane, sl + MSZ, p.Qs + row * DD);
            }
            if (row >= 0) {
                cp.issue(lane, sl + O_P, p.fPs_in + row * DD);
                issue_vec<D>(lane, sl + O_M, p.fms_in + row * D);
            } else {
                cp.issue(lane, sl + O_P, p.P0);
                if (p.m0 != nullptr) issue_vec<D>(lane, sl + O_M, p.m0);
                else if (lane < D) sl[O_M + lane] = 0.0;
            }
        }
        cp_commit();
    };
#pragma unroll 1
    for (int j = 0; j < PD; ++j) issue(j);
    const VecP<MT> hP = gld_vp<D, MT>(p.H, c);
    const VecR<MT> hR = gld_vr<D, MT>(p.H, r);
    const double Rv = p.R[0];
    const double gl = ADJ ? p.g[0] : 1.0;
    const double* st = rstates + (nchunks - 1 - chunk) * (2 * DD + 2 * D);
    VecR<MT> dm = gld_vr<D, MT>(st, r), lam = gld_vr<D, MT>(st + D, r);
    Mat<MT> dP = gld_mat<D, MT>(st + 2 * D, r, c), Lam = gld_mat<D, MT>(st + 2 * D + DD, r, c);
    VecR<MT> dHacc;
#pragma unroll
    for (int t = 0; t < MT; ++t) dHacc.v[t] = 0.0;
    double dRacc = 0.0;
    double ynext = p.y[k_hi - 1];
    // one visited row; FIRST (compile-time) = step 0 of the global series, peeled out of the steady-state loop
    auto body = [&](int j, auto first_tag) {
        constexpr bool first = decltype(first_tag)::value;
        const long k = k_hi - 1 - j;
        issue(j + PD);
        cp_wait<PD - 1>();  // rows k and k - 1 have landed
        __syncwarp();
        const double yk = ynext;
        if (j + 1 < nrows) ynext = p.y[k - 1];
        const bool obs = !isnan(yk);
        const double* sl = slot(k);
        const double* slp = slot(k - 1);
        if constexpr (SMOOTH) {
            // sm_k = m_k - P_k lam_k ; sP_k = P_k - P_k Lam_k P_k   (state entering from above)
            const auto Pk = in_op<MT>(sl + O_P, false, r, c);
            const Mat<MT> Zt = mulT<MT, D>(Pk, Lam);  // P_k Lam = (Lam P_k)^T
            Mat<MT> PZ = mulT<MT, D>(Pk, Zt);         // P_k (Lam P_k)
#pragma unroll
            for (int a = 0; a < MT; ++a)
#pragma unroll
                for (int b2 = 0; b2 < MT; ++b2) {
                    PZ.v[a][b2][0] = oget<MT>(Pk, a, b2, 0) - PZ.v[a][b2][0];
                    PZ.v[a][b2][1] = oget<MT>(Pk, a, b2, 1) - PZ.v[a][b2][1];
                }
            gst_mat<D, MT>(p.sPs + k * DD, PZ, 1.0, sp);
            const VecR<MT> v1 = mv(Pk, vr2vp(lam, c));
            const VecR<MT> mk = ld_vr<MT>(sl + O_M, r);
            gst_vr<D, MT>(p.sms + k * D, vaxpy(-1.0, v1, mk), r, c);
        }
        // forward quantities of step k
        const auto F = in_op<MT>(sl, false, r, c);
        const auto Ft = in_op<MT>(sl, true, r, c);
        const auto Pprev = in_op<MT>(slp + O_P, false, r, c);
        const VecR<MT> mprev = ld_vr<MT>(slp + O_M, r);
        Mat<MT> Pp = ld_sym<MT>(sl + MSZ, r, c);
        {
            const Mat<MT> T1 = mulT<MT, D>(F, Pprev);
            mmT<MT, D>(Pp, T1, F);
        }
        const VecR<MT> mp = mv(F, ld_vp<MT>(slp + O_M, c));
        VecR<MT> u = mv(Pp, hP);
        double hu, hm;
        dots2<MT>(hR, u, hR, mp, lane, c, hu, hm);
        const double s = Rv + hu;
        const double rr = obs ? yk - hm : 0.0;
        if constexpr (first) {
            // step 0 of the global series: the log-likelihood term sees (F0 m0, F0 P0 F0^T + Q0), the update is made
            // on (m0, P0) directly (parallel.py:24-30, :136-141)
            if constexpr (ADJ) {
                double sbar0 = 0.0, rbar0 = 0.0;
                if (obs) {
                    const double is = 1.0 / s;
                    sbar0 = 0.5 * (rr * rr * is * is - is);
                    rbar0 = -rr * is;
                    dRacc += sbar0;
#pragma unroll
                    for (int t = 0; t < MT; ++t) dHacc.v[t] += 2.0 * sbar0 * u.v[t] - mp.v[t] * rbar0;
                }
                Mat<MT> dPp0;
                mzero(dPp0);
                {
                    const VecR<MT> x1[1] = {vscale(sbar0, hR)}, y1[1] = {hR};
                    rank_update<MT, 1>(dPp0, x1, y1, c);
                }
                const VecR<MT> dmp0 = vscale(-rbar0, hR);
                gst_mat<D, MT>(p.dQs + k * DD, dPp0, gl, sp);
                const Mat<MT> X = mulT<MT, D>(dPp0, Ft);   // dPp0 F
                const Mat<MT> Xt = mulT<MT, D>(Ft, dPp0);  // F^T dPp0
                Mat<MT> Y = mulT<MT, D>(X, Pprev);
                {
                    const VecR<MT> x1[1] = {vscale(0.5, dmp0)}, y1[1] = {mprev};
                    rank_update<MT, 1>(Y, x1, y1, c);
                }
                gst_mat<D, MT>(p.dFs + k * DD, Y, 2.0 * gl, sp);
                Mat<MT> dP0 = mulT<MT, D>(Xt, Ft);  // F^T dPp0 F
                // adjoint of the update on (m0, P0)
                u = mv(Pprev, hP);
                double hu0, hm0;
                dots2<MT>(hR, u, hR, mprev, lane, c, hu0, hm0);
                const double s0 = Rv + hu0, r0 = yk - hm0;
                if (obs) {
                    const VecR<MT> Pu = mv(dP, vr2vp(u, c));
                    double udm, uPu;
                    dots2<MT>(u, dm, u, Pu, lane, c, udm, uPu);
                    const double is0 = 1.0 / s0;
                    const double rbar = udm * is0;
                    const double sbar = (-udm * r0 + uPu) * is0 * is0;
                    VecR<MT> ut;
#pragma unroll
                    for (int t = 0; t < MT; ++t) ut.v[t] = dm.v[t] * r0 * is0 - 2.0 * Pu.v[t] * is0 + sbar * hR.v[t];
                    const VecR<MT> Pput = mv(Pprev, vr2vp(ut, c));
                    dRacc += sbar;
#pragma unroll
                    for (int t = 0; t < MT; ++t) dHacc.v[t] += sbar * u.v[t] + Pput.v[t] - mprev.v[t] * rbar;
                    const VecR<MT> x2[2] = {vscale(0.5, ut), vscale(0.5, hR)}, y2[2] = {hR, ut};
                    rank_update<MT, 2>(dP, x2, y2, c);
                }
                if (p.dP0 != nullptr) {
#pragma unroll
                    for (int a = 0; a < MT; ++a)
#pragma unroll
                        for (int b2 = 0; b2 < MT; ++b2) {
                            dP0.v[a][b2][0] += dP.v[a][b2][0];
                            dP0.v[a][b2][1] += dP.v[a][b2][1];
                        }
                    gst_mat<D, MT>(p.dP0, dP0, gl, sp);
                }
            }
            __syncwarp();
            return;  // k == 0: nothing below this row
        }
        // measurement part (all terms vanish with is = 0 when nothing is observed)
        const double is = obs ? 1.0 / s : 0.0;
        const VecP<MT> uP = vr2vp(u, c);
        VecR<MT> Pu = u, gv = u;
        if constexpr (ADJ) Pu = mv(dP, uP);
        if constexpr (SMOOTH) gv = mv(Lam, uP);
        double dv[4];
        {
            const VecR<MT>* const xa[4] = {&u, &u, &u, &u};
            const VecR<MT>* const xb[4] = {&dm, &Pu, &lam, &gv};
            dots4<MT>(xa, xb, lane, c, dv);
        }
        const double udm = ADJ ? dv[0] : 0.0, uPu = ADJ ? dv[1] : 0.0, ulam = SMOOTH ? dv[2] : 0.0,
                     alpha = SMOOTH ? dv[3] : 0.0;
        const double rbar = (udm - rr) * is;
        const double sbar = (-udm * rr + uPu) * is * is + 0.5 * (rr * rr * is * is - is);
        VecR<MT> dmp = dm, lt = lam;
        if constexpr (ADJ) {
            VecR<MT> ut;
#pragma unroll
            for (int t = 0; t < MT; ++t) {
                ut.v[t] = dm.v[t] * rr * is - 2.0 * Pu.v[t] * is + sbar * hR.v[t];
                dmp.v[t] = dm.v[t] - hR.v[t] * rbar;
            }
            const VecR<MT> Pput = mv(Pp, vr2vp(ut, c));
            dRacc += sbar;
#pragma unroll
            for (int t = 0; t < MT; ++t) dHacc.v[t] += sbar * u.v[t] + Pput.v[t] - mp.v[t] * rbar;
            const VecR<MT> x2[2] = {vscale(0.5, ut), vscale(0.5, hR)}, y2[2] = {hR, ut};
            rank_update<MT, 2>(dP, x2, y2, c);  // dPp = dP + sym(ut h^T)
            gst_mat<D, MT>(p.dQs + k * DD, dP, gl, sp);
            const Mat<MT> X = mulT<MT, D>(dP, Ft);     // dPp F
            const Mat<MT> Xt = mulT<MT, D>(Ft, dP);    // F^T dPp
            Mat<MT> Y = mulT<MT, D>(X, Pprev);         // dPp F P_{k-1}
            {
                const VecR<MT> x1[1] = {vscale(0.5, dmp)}, y1[1] = {mprev};
                rank_update<MT, 1>(Y, x1, y1, c);
            }
            gst_mat<D, MT>(p.dFs + k * DD, Y, 2.0 * gl, sp);
            dP = mulT<MT, D>(Xt, Ft);                  // F^T dPp F
            dm = mv(Ft, vr2vp(dmp, c));         // F^T dmp
        }
        if constexpr (SMOOTH) {
            const double cm = is + alpha * is * is;
            VecR<MT> xv;
#pragma unroll
            for (int t = 0; t < MT; ++t) {
                lt.v[t] = lam.v[t] - hR.v[t] * (ulam + rr) * is;
                xv.v[t] = -gv.v[t] * is + 0.5 * cm * hR.v[t];
            }
            const VecR<MT> x2[2] = {hR, xv}, y2[2] = {xv, hR};
            rank_update<MT, 2>(Lam, x2, y2, c);  // Lt = Lam - (h g^T + g h^T)/s + h h^T (1/s + alpha/s^2)
            const Mat<MT> X2t = mulT<MT, D>(Ft, Lam);   // F^T Lt
            Lam = mulT<MT, D>(X2t, Ft);                 // F^T Lt F
            lam = mv(Ft, vr2vp(lt, c));
        }
        __syncwarp();
    };
    const int nmain = (k_lo == 0 && p.first_special) ? nrows - 1 : nrows;
#pragma unroll 1
    for (int j = 0; j < nmain; ++j) body(j, std::false_type{});
    if (nmain < nrows) body(nmain, std::true_type{});
    cp_wait<0>();
    if constexpr (ADJ) {
        if (lane == 0) part[chunk * (1 + D)] = dRacc;
        if (c == 0) {
#pragma unroll
            for (int t = 0; t < MT; ++t)
                if (8 * t + r < D) part[chunk * (1 + D) + 1 + 8 * t + r] = dHacc.v[t];
        }
    }
}

}  // namespace frag
}  // namespace mid
}  // namespace pssgp
