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

// ================================================================================================
// Traits: how a d x d matrix / a d vector is held by one warp.  The kernels below are written once against this
// interface.  Std<D>: MT x MT tiles of 8 x 8 (any D).  Bord9: D = 9 as an 8 x 8 core tile + border column + border row +
// corner — the 16 x 16 padding of the tile form wastes 3/4 of the DMMA work at D = 9 (BASELINE configs[4]).
// ================================================================================================
template <int D_> struct Std {
    static constexpr int D = D_, MT = FGeo<D>::MT, DD = D * D;
    static constexpr int MSZ = FGeo<D>::MSZ;   // doubles per matrix in a ring slot
    static constexpr int VSZ = FGeo<D>::DP;    // doubles per vector in a ring slot
    using M = Mat<MT>;
    using VR = VecR<MT>;
    using VP = VecP<MT>;
    using Op = typename InOp<MT>::type;
    using Cp = CpPlan<D, MT>;
    using St = StPlan<D, MT>;
    MDEV static M zero() { M m; mzero(m); return m; }
    MDEV static M identity(int r, int c) { return identity_d<D, MT>(r, c); }
    MDEV static VR vzero() { VR v;
#pragma unroll
        for (int t = 0; t < MT; ++t) v.v[t] = 0.0;
        return v; }
    MDEV static Op in_op(const double* slot, bool tr, int r, int c) { return InOp<MT>::make(slot, tr, r, c); }
    MDEV static M ld_mat(const double* slot, int r, int c) { return frag::ld_mat<MT>(slot, r, c); }
    MDEV static M ld_sym(const double* slot, int r, int c) { return frag::ld_sym<MT>(slot, r, c); }
    MDEV static VR ld_vr(const double* v, int r) { return frag::ld_vr<MT>(v, r); }
    MDEV static VP ld_vp(const double* v, int c) { return frag::ld_vp<MT>(v, c); }
    MDEV static void issue_vec(int lane, double* dst, const double* src) { frag::issue_vec<D>(lane, dst, src); }
    MDEV static void zero_vec(int lane, double* dst) { if (lane < D) dst[lane] = 0.0; }
    MDEV static M gld_mat(const double* src, int r, int c) { return frag::gld_mat<D, MT>(src, r, c); }
    MDEV static VR gld_vr(const double* src, int r) { return frag::gld_vr<D, MT>(src, r); }
    MDEV static VP gld_vp(const double* src, int c) { return frag::gld_vp<D, MT>(src, c); }
    template <bool TRANS = false> MDEV static void gst_mat(double* dst, const M& m, double scale, const St& sp, int, int) {
        frag::gst_mat<D, MT, TRANS>(dst, m, scale, sp);
    }
    MDEV static void gst_vr(double* dst, const VR& v, int r, int c) { frag::gst_vr<D, MT>(dst, v, r, c); }
    template <class XO, class SO> MDEV static void mmT(M& acc, const XO& X, const SO& S, int) { frag::mmT<MT, D>(acc, X, S); }
    template <class XO, class SO> MDEV static M mulT(const XO& X, const SO& S, int) { return frag::mulT<MT, D>(X, S); }
    // yP (the y vectors in VP layout) is only needed by the bordered form
    template <int NQ> MDEV static void rank_update(M& acc, const VR (&x)[NQ], const VR (&y)[NQ], const VP (&)[NQ], int c) {
        frag::rank_update<MT, NQ>(acc, x, y, c);
    }
    MDEV static VP vp_for_rank(const VR&, int) { return VP{}; }   // not needed: no shuffles spent
    template <class MO> MDEV static VR mv(const MO& Mx, const VP& v) { return frag::mv<MT>(Mx, v); }
    MDEV static VP vr2vp(const VR& x, int c) { return frag::vr2vp<MT>(x, c); }
    MDEV static void dots2(const VR& x0, const VR& y0, const VR& x1, const VR& y1, int lane, int c, double& o0, double& o1) {
        frag::dots2<MT>(x0, y0, x1, y1, lane, c, o0, o1);
    }
    MDEV static void dots4(const VR* const (&x)[4], const VR* const (&y)[4], int lane, int c, double (&out)[4]) {
        frag::dots4<MT>(x, y, lane, c, out);
    }
    MDEV static VR vscale(double a, const VR& x) { return frag::vscale(a, x); }
    MDEV static VR vaxpy(double a, const VR& x, const VR& y) { return frag::vaxpy(a, x, y); }
    MDEV static VR vlin3(double a, const VR& x, double b, const VR& y, double cc, const VR& z) {
        VR o;
#pragma unroll
        for (int t = 0; t < MT; ++t) o.v[t] = a * x.v[t] + b * y.v[t] + cc * z.v[t];
        return o;
    }
    MDEV static void madd(M& a, const M& b) {
#pragma unroll
        for (int i = 0; i < MT; ++i)
#pragma unroll
            for (int j = 0; j < MT; ++j) { a.v[i][j][0] += b.v[i][j][0]; a.v[i][j][1] += b.v[i][j][1]; }
    }
    // X - m with X an input operand
    MDEV static M op_minus(const Op& X, const M& m) {
        M o;
#pragma unroll
        for (int i = 0; i < MT; ++i)
#pragma unroll
            for (int j = 0; j < MT; ++j) {
                o.v[i][j][0] = oget<MT>(X, i, j, 0) - m.v[i][j][0];
                o.v[i][j][1] = oget<MT>(X, i, j, 1) - m.v[i][j][1];
            }
        return o;
    }
};

// ---- D = 9: 8 x 8 core + border -------------------------------------------------------------------------------
struct BM { double c[2]; double col; double row[2]; double cor; };  // core CF | M[r][8] | M[8][2c+s] | M[8][8]
struct BVR { double v, e; };      // x[r] | x[8]
struct BVP { double v[2], e; };   // x[2c+s] | x[8]

MDEV double red_c(double x) {  // sum over the four lanes of a row (result on all of them)
    x += __shfl_xor_sync(FULL, x, 1);
    x += __shfl_xor_sync(FULL, x, 2);
    return x;
}

struct Bord9 {
    static constexpr int D = 9, DD = 81;
    static constexpr int MSZ = 88;   // slot: core 64 (row-major 8 x 8) | col 8 | row 8 | corner (+ pad)
    static constexpr int VSZ = 10;   // slot: x[0..8] (+ pad)
    using M = BM;
    using VR = BVR;
    using VP = BVP;
    using Op = BM;
    struct Cp {
        int soff[3], doff[3];
        MDEV void init(int lane) {
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const int idx = lane + 32 * q;
                const int i = idx / 9, j = idx - i * 9;
                soff[q] = idx;
                doff[q] = (i < 8 && j < 8) ? i * 8 + j : (i < 8 ? 64 + i : (j < 8 ? 72 + j : 80));
            }
        }
        MDEV void issue(int lane, double* dst, const double* src) const {
#pragma unroll
            for (int q = 0; q < 3; ++q)
                if (q < 2 || lane + 64 < 81) cp8(dst + doff[q], src + soff[q]);
        }
    };
    struct St { MDEV void init(int, int) {} };
    MDEV static M zero() { return BM{{0.0, 0.0}, 0.0, {0.0, 0.0}, 0.0}; }
    MDEV static M identity(int r, int c) {
        BM m = zero();
        if (r == 2 * c) m.c[0] = 1.0;
        if (r == 2 * c + 1) m.c[1] = 1.0;
        m.cor = 1.0;
        return m;
    }
    MDEV static VR vzero() { return BVR{0.0, 0.0}; }
    MDEV static M ld_mat(const double* sl, int r, int c) {
        BM m;
        const double2 a = *reinterpret_cast<const double2*>(sl + r * 8 + 2 * c);
        const double2 b = *reinterpret_cast<const double2*>(sl + 72 + 2 * c);
        m.c[0] = a.x; m.c[1] = a.y;
        m.col = sl[64 + r];
        m.row[0] = b.x; m.row[1] = b.y;
        m.cor = sl[80];
        return m;
    }
    MDEV static M ld_matT(const double* sl, int r, int c) {
        BM m;
        m.c[0] = sl[(2 * c) * 8 + r];
        m.c[1] = sl[(2 * c + 1) * 8 + r];
        m.col = sl[72 + r];
        const double2 b = *reinterpret_cast<const double2*>(sl + 64 + 2 * c);
        m.row[0] = b.x; m.row[1] = b.y;
        m.cor = sl[80];
        return m;
    }
    MDEV static Op in_op(const double* slot, bool tr, int r, int c) { return tr ? ld_matT(slot, r, c) : ld_mat(slot, r, c); }
    MDEV static M ld_sym(const double* sl, int r, int c) {
        BM a = ld_mat(sl, r, c);
        const BM t = ld_matT(sl, r, c);
        a.c[0] = 0.5 * (a.c[0] + t.c[0]); a.c[1] = 0.5 * (a.c[1] + t.c[1]);
        a.col = 0.5 * (a.col + t.col);
        a.row[0] = 0.5 * (a.row[0] + t.row[0]); a.row[1] = 0.5 * (a.row[1] + t.row[1]);
        return a;
    }
    MDEV static VR ld_vr(const double* v, int r) { return BVR{v[r], v[8]}; }
    MDEV static VP ld_vp(const double* v, int c) {
        const double2 a = *reinterpret_cast<const double2*>(v + 2 * c);
        return BVP{{a.x, a.y}, v[8]};
    }
    MDEV static void issue_vec(int lane, double* dst, const double* src) { if (lane < 9) cp8(dst + lane, src + lane); }
    MDEV static void zero_vec(int lane, double* dst) { if (lane < 9) dst[lane] = 0.0; }
    MDEV static M gld_mat(const double* src, int r, int c) {
        BM m;
        m.c[0] = src[r * 9 + 2 * c]; m.c[1] = src[r * 9 + 2 * c + 1];
        m.col = src[r * 9 + 8];
        m.row[0] = src[72 + 2 * c]; m.row[1] = src[72 + 2 * c + 1];
        m.cor = src[80];
        return m;
    }
    MDEV static VR gld_vr(const double* src, int r) { return src != nullptr ? BVR{src[r], src[8]} : BVR{0.0, 0.0}; }
    MDEV static VP gld_vp(const double* src, int c) {
        return src != nullptr ? BVP{{src[2 * c], src[2 * c + 1]}, src[8]} : BVP{{0.0, 0.0}, 0.0};
    }
    template <bool TRANS = false> MDEV static void gst_mat(double* dst, const M& m, double scale, const St&, int r, int c) {
        if (TRANS) {
            dst[(2 * c) * 9 + r] = scale * m.c[0];
            dst[(2 * c + 1) * 9 + r] = scale * m.c[1];
            if (c == 0) dst[72 + r] = scale * m.col;
            if (r == 0) { dst[(2 * c) * 9 + 8] = scale * m.row[0]; dst[(2 * c + 1) * 9 + 8] = scale * m.row[1]; }
        } else {
            __stcs(dst + r * 9 + 2 * c, scale * m.c[0]);
            __stcs(dst + r * 9 + 2 * c + 1, scale * m.c[1]);
            if (c == 0) __stcs(dst + r * 9 + 8, scale * m.col);
            if (r == 0) { __stcs(dst + 72 + 2 * c, scale * m.row[0]); __stcs(dst + 72 + 2 * c + 1, scale * m.row[1]); }
        }
        if (r == 0 && c == 0) dst[80] = scale * m.cor;
    }
    MDEV static void gst_vr(double* dst, const VR& v, int r, int c) {
        if (c == 0) dst[r] = v.v;
        if (r == 0 && c == 0) dst[8] = v.e;
    }
    // acc += X S^T
    MDEV static void mmT(M& acc, const M& X, const M& S, int c) {
        dmma(acc.c, X.c[0], S.c[0]);
        dmma(acc.c, X.c[1], S.c[1]);
        dmma(acc.c, c == 0 ? X.col : 0.0, c == 0 ? S.col : 0.0);
        const double pc = red_c(fma(X.c[0], S.row[0], X.c[1] * S.row[1]));      // sum_k X[r][k] S[8][k]
        const double pr = red_c(fma(S.c[0], X.row[0], S.c[1] * X.row[1]));      // sum_k X[8][k] S[n][k], n = r
        const double pz = red_c(fma(X.row[0], S.row[0], X.row[1] * S.row[1]));  // sum_k X[8][k] S[8][k]
        acc.col += fma(X.col, S.cor, pc);
        const double zr = fma(X.cor, S.col, pr);                                // Z[8][n] at lane (n, .)
        acc.row[0] += __shfl_sync(FULL, zr, (2 * c) * 4);
        acc.row[1] += __shfl_sync(FULL, zr, (2 * c + 1) * 4);
        acc.cor += fma(X.cor, S.cor, pz);
    }
    MDEV static M mulT(const M& X, const M& S, int c) {
        BM acc = zero();
        mmT(acc, X, S, c);
        return acc;
    }
    template <int NQ> MDEV static void rank_update(M& acc, const VR (&x)[NQ], const VR (&y)[NQ], const VP (&yP)[NQ], int c) {
        double a = 0.0, b = 0.0;
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            a = (c == q) ? x[q].v : a;
            b = (c == q) ? y[q].v : b;
            acc.col = fma(x[q].v, y[q].e, acc.col);
            acc.row[0] = fma(x[q].e, yP[q].v[0], acc.row[0]);
            acc.row[1] = fma(x[q].e, yP[q].v[1], acc.row[1]);
            acc.cor = fma(x[q].e, y[q].e, acc.cor);
        }
        dmma(acc.c, a, b);
    }
    MDEV static VP vp_for_rank(const VR& x, int c) { return vr2vp(x, c); }
    MDEV static VR mv(const M& Mx, const VP& v) {
        BVR y;
        y.v = fma(Mx.col, v.e, red_c(fma(Mx.c[0], v.v[0], Mx.c[1] * v.v[1])));
        y.e = fma(Mx.cor, v.e, red_c(fma(Mx.row[0], v.v[0], Mx.row[1] * v.v[1])));
        return y;
    }
    MDEV static VP vr2vp(const VR& x, int c) {
        BVP y;
        y.v[0] = __shfl_sync(FULL, x.v, (2 * c) * 4);
        y.v[1] = __shfl_sync(FULL, x.v, (2 * c + 1) * 4);
        y.e = x.e;
        return y;
    }
    MDEV static void dots2(const VR& x0, const VR& y0, const VR& x1, const VR& y1, int lane, int c, double& o0, double& o1) {
        double p = (c & 1) ? x1.v * y1.v : x0.v * y0.v;
        p += __shfl_xor_sync(FULL, p, 4);
        p += __shfl_xor_sync(FULL, p, 8);
        p += __shfl_xor_sync(FULL, p, 16);
        o0 = fma(x0.e, y0.e, __shfl_sync(FULL, p, lane & ~3));
        o1 = fma(x1.e, y1.e, __shfl_sync(FULL, p, (lane & ~3) | 1));
    }
    MDEV static void dots4(const VR* const (&x)[4], const VR* const (&y)[4], int lane, int c, double (&out)[4]) {
        double p = 0.0;
#pragma unroll
        for (int q = 0; q < 4; ++q) p = (c == q) ? x[q]->v * y[q]->v : p;
        p += __shfl_xor_sync(FULL, p, 4);
        p += __shfl_xor_sync(FULL, p, 8);
        p += __shfl_xor_sync(FULL, p, 16);
#pragma unroll
        for (int q = 0; q < 4; ++q) out[q] = fma(x[q]->e, y[q]->e, __shfl_sync(FULL, p, (lane & ~3) | q));
    }
    MDEV static VR vscale(double a, const VR& x) { return BVR{a * x.v, a * x.e}; }
    MDEV static VR vaxpy(double a, const VR& x, const VR& y) { return BVR{fma(a, x.v, y.v), fma(a, x.e, y.e)}; }
    MDEV static VR vlin3(double a, const VR& x, double b, const VR& y, double cc, const VR& z) {
        return BVR{a * x.v + b * y.v + cc * z.v, a * x.e + b * y.e + cc * z.e};
    }
    MDEV static void madd(M& a, const M& b) {
        a.c[0] += b.c[0]; a.c[1] += b.c[1]; a.col += b.col; a.row[0] += b.row[0]; a.row[1] += b.row[1]; a.cor += b.cor;
    }
    MDEV static M op_minus(const Op& X, const M& m) {
        return BM{{X.c[0] - m.c[0], X.c[1] - m.c[1]}, X.col - m.col, {X.row[0] - m.row[0], X.row[1] - m.row[1]}, X.cor - m.cor};
    }
};

#ifndef PSSGP_NO_BORDERED
template <int D> struct TraitsFor { using type = Std<D>; };
template <> struct TraitsFor<9> { using type = Bord9; };
#else
template <int D> struct TraitsFor { using type = Std<D>; };
#endif

template <int D> constexpr int frag_nslot() { return D <= 9 ? 4 : 3; }
// warps per CTA: as many as the ring slots of a CTA fit in shared memory, at most `cap` (the launch bound; the option
// "mid_warps" lowers it at run time)
template <int D> constexpr int frag_warps(int slot_doubles, int cap) {
    const int fit = (200 * 1024) / (frag_nslot<D>() * slot_doubles * 8);
    return fit > cap ? cap : (fit < 1 ? 1 : fit);
}
#ifndef PSSGP_FRAG_CAP_D9
#define PSSGP_FRAG_CAP_D9 16
#endif
#ifndef PSSGP_FRAG_CAP_SMALL
#define PSSGP_FRAG_CAP_SMALL 24
#endif
template <int D> constexpr int frag_cap(int big) { return (D <= 8) ? PSSGP_FRAG_CAP_SMALL : (D == 9 ? PSSGP_FRAG_CAP_D9 : big); }

// ------------------------------------------------------------------------------------------------
// K1: chunk aggregates of the filter.  Tracks At = A^T, C, J, b, eta.
// ------------------------------------------------------------------------------------------------
template <int D> struct FK1 {
    using TR = typename TraitsFor<D>::type;
    static constexpr int NSLOT = frag_nslot<D>();
    static constexpr int SLOT = 2 * TR::MSZ;  // F | Q
    static constexpr int WPC = frag_warps<D>(SLOT, frag_cap<D>(12));
    static constexpr size_t WARP_SMEM = (size_t)NSLOT * SLOT * 8;
    static constexpr size_t SMEM = (size_t)WPC * WARP_SMEM;
};

template <int D>
__global__ void __launch_bounds__(FK1<D>::WPC * 32) fk1_filter_reduce(Params p, int L, long nchunks, double* __restrict__ aggs) {
    using K = FK1<D>;
    using TR = typename K::TR;
    using M = typename TR::M;
    using VR = typename TR::VR;
    using VP = typename TR::VP;
    constexpr int MSZ = TR::MSZ, NSLOT = K::NSLOT, DD = D * D;
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
    typename TR::Cp cp;
    cp.init(lane);
    typename TR::St sp;
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
    const VP hP = TR::gld_vp(p.H, c);
    const VR hR = TR::gld_vr(p.H, r);
    const double Rv = p.R[0];
    M At = TR::identity(r, c), C = TR::zero(), J = TR::zero();
    VR b = TR::vzero(), eta = TR::vzero();
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
            const auto F = TR::in_op(sl, false, r, c);
            M Cn = TR::ld_sym(sl + MSZ, r, c);
            At = TR::mulT(At, F, c);                   // (F A)^T = A^T F^T
            {
                const M T2 = TR::mulT(F, C, c);        // F C   (C symmetric)
                TR::mmT(Cn, T2, F, c);                 // F C F^T + Q
            }
            C = Cn;
            b = TR::mv(F, TR::vr2vp(b, c));
        }
        const bool obs = !isnan(yk);
        const VR u = TR::mv(C, hP);
        const VR w = TR::mv(At, hP);      // A^T h
        double hu, hb;
        TR::dots2(hR, u, hR, b, lane, c, hu, hb);
        const double is = obs ? 1.0 / (Rv + hu) : 0.0;
        const double eis = obs ? (yk - hb) * is : 0.0;
        const VP uP = TR::vp_for_rank(u, c), wP = TR::vp_for_rank(w, c);
        {
            const VR x1[1] = {TR::vscale(is, w)}, y1[1] = {w};
            const VP p1[1] = {wP};
            TR::template rank_update<1>(J, x1, y1, p1, c);   // J += w w^T / s
        }
        {
            const VR x1[1] = {TR::vscale(-is, w)}, y1[1] = {u};
            const VP p1[1] = {uP};
            TR::template rank_update<1>(At, x1, y1, p1, c);  // A -= u w^T / s
        }
        {
            const VR x1[1] = {TR::vscale(-is, u)}, y1[1] = {u};
            const VP p1[1] = {uP};
            TR::template rank_update<1>(C, x1, y1, p1, c);   // C -= u u^T / s
        }
        eta = TR::vaxpy(eis, w, eta);
        b = TR::vaxpy(eis, u, b);
        __syncwarp();  // every lane has read the slot: it may be refilled
    }
    cp_wait<0>();
    double* out = aggs + chunk * (3 * DD + 2 * D);
    TR::template gst_mat<true>(out, At, 1.0, sp, r, c);
    TR::template gst_mat<false>(out + DD, C, 1.0, sp, r, c);
    TR::template gst_mat<false>(out + 2 * DD, J, 1.0, sp, r, c);
    TR::gst_vr(out + 3 * DD, b, r, c);
    TR::gst_vr(out + 3 * DD + D, eta, r, c);
}

// ------------------------------------------------------------------------------------------------
// K2: seeded filter recursion (+ log-likelihood) and, with REV, the chunk aggregate of the combined reverse scan
// (tracked as Abt = Abar^T, Ba, Bm, a).  STORED: filtered moments are read instead of recomputed.
// ------------------------------------------------------------------------------------------------
template <int D, bool REV, bool STORED> struct FK2 {
    using TR = typename TraitsFor<D>::type;
    static constexpr int NSLOT = frag_nslot<D>();
    static constexpr int NM = 2 + (STORED ? 1 : 0);  // F | Q | P_{k-1}
    static constexpr int SLOT = NM * TR::MSZ + (STORED ? TR::VSZ : 0);
    static constexpr int WPC = frag_warps<D>(SLOT, frag_cap<D>(12));
    static constexpr size_t WARP_SMEM = (size_t)NSLOT * SLOT * 8;
    static constexpr size_t SMEM = (size_t)WPC * WARP_SMEM;
};

template <int D, bool REV, bool STORED>
__global__ void __launch_bounds__(FK2<D, REV, STORED>::WPC * 32)
fk2_forward(Params p, int L, long nchunks, const double* __restrict__ fstates, double* __restrict__ part,
            double* __restrict__ raggs) {
    using K = FK2<D, REV, STORED>;
    using TR = typename K::TR;
    using M = typename TR::M;
    using VR = typename TR::VR;
    using VP = typename TR::VP;
    constexpr int MSZ = TR::MSZ, NSLOT = K::NSLOT, DD = D * D;
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
    typename TR::Cp cp;
    cp.init(lane);
    typename TR::St sp;
    sp.init(r, c);
    auto issue = [&](int i) {
        if (i < nrows) {
            double* sl = ring + (i % NSLOT) * K::SLOT;
            const long k = k_lo + i;
            cp.issue(lane, sl, p.Fs + k * DD);
            cp.issue(lane, sl + MSZ, p.Qs + k * DD);
            if constexpr (STORED) {
                cp.issue(lane, sl + O_P, k > 0 ? p.fPs_in + (k - 1) * DD : p.P0);
                if (k > 0) TR::issue_vec(lane, sl + O_M, p.fms_in + (k - 1) * D);
                else if (p.m0 != nullptr) TR::issue_vec(lane, sl + O_M, p.m0);
                else TR::zero_vec(lane, sl + O_M);
            }
        }
        cp_commit();
    };
#pragma unroll 1
    for (int i = 0; i < PD; ++i) issue(i);
    const VP hP = TR::gld_vp(p.H, c);
    const VR hR = TR::gld_vr(p.H, r);
    const double Rv = p.R[0];
    M P = TR::zero();
    VR m = TR::vzero();
    if constexpr (!STORED) {
        const double* st = fstates + chunk * (D + DD);
        m = TR::gld_vr(st, r);
        P = TR::gld_mat(st + D, r, c);
    }
    M Abt = TR::identity(r, c), Ba = TR::zero(), Bm = TR::zero();
    VR av = TR::vzero();
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
        const auto F = TR::in_op(sl, false, r, c);
        if constexpr (STORED) {
            P = TR::ld_mat(sl + O_P, r, c);
            m = TR::ld_vr(sl + O_M, r);
        }
        M Pp = TR::ld_sym(sl + MSZ, r, c);
        {
            const M T1 = TR::mulT(F, P, c);  // F P (P symmetric)
            TR::mmT(Pp, T1, F, c);           // F P F^T + Q
        }
        VR mp = TR::mv(F, TR::vr2vp(m, c));
        VR u = TR::mv(Pp, hP);
        double hu, hm;
        TR::dots2(hR, u, hR, mp, lane, c, hu, hm);
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
            u = TR::mv(Pp, hP);
            TR::dots2(hR, u, hR, mp, lane, c, hu, hm);
            s = Rv + hu;
            e = obs ? yk - hm : 0.0;
        }
        const double is = obs ? 1.0 / s : 0.0;
        const double eis = e * is;
        const VP uP = TR::vp_for_rank(u, c);
        if constexpr (!STORED) {
            P = Pp;
            {
                const VR x1[1] = {TR::vscale(-is, u)}, y1[1] = {u};
                const VP p1[1] = {uP};
                TR::template rank_update<1>(P, x1, y1, p1, c);
            }
            m = TR::vaxpy(eis, u, mp);
            TR::template gst_mat<false>(p.fPs + k * DD, P, 1.0, sp, r, c);
            TR::gst_vr(p.fms + k * D, m, r, c);
        }
        if constexpr (REV) {
            // append step k on the later side of the chunk's reverse aggregate (nothing for the global first step)
            const auto Ft = TR::in_op(sl, true, r, c);
            const VR w = TR::mv(Ft, hP);                    // F^T h
            const VR t = TR::mv(Abt, TR::vr2vp(w, c));      // Abar_old^T w
            const double isr = first ? 0.0 : is, eisr = first ? 0.0 : eis;
            if (!first) Abt = TR::mulT(Abt, F, c);          // (F Abar_old)^T
            const VP tP = TR::vp_for_rank(t, c), aP = TR::vp_for_rank(av, c);
            {
                const VR x1[1] = {TR::vscale(-isr, t)}, y1[1] = {u};
                const VP p1[1] = {uP};
                TR::template rank_update<1>(Abt, x1, y1, p1, c);  // Abar = F Abar_old - u t^T / s
            }
            const double beta = 0.5 * (eisr * eisr - isr);
            {
                const VR x3[3] = {TR::vscale(beta, t), TR::vscale(0.5 * eisr, t), TR::vscale(0.5 * eisr, av)};
                const VR y3[3] = {t, av, t};
                const VP p3[3] = {tP, aP, tP};
                TR::template rank_update<3>(Ba, x3, y3, p3, c);
            }
            {
                const VR x1[1] = {TR::vscale(isr, t)}, y1[1] = {t};
                const VP p1[1] = {tP};
                TR::template rank_update<1>(Bm, x1, y1, p1, c);
            }
            av = TR::vaxpy(eisr, t, av);
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
        TR::template gst_mat<true>(out, Abt, 1.0, sp, r, c);
        TR::template gst_mat<false>(out + DD, Ba, 1.0, sp, r, c);
        TR::template gst_mat<false>(out + 2 * DD, Bm, 1.0, sp, r, c);
        TR::gst_vr(out + 3 * DD, av, r, c);
    }
}

// ------------------------------------------------------------------------------------------------
// K3: reverse pass — smoothed moments (SMOOTH) and / or gradient of the log-likelihood (ADJ).
// ------------------------------------------------------------------------------------------------
template <int D, bool SMOOTH, bool ADJ> struct FK3 {
    using TR = typename TraitsFor<D>::type;
    static constexpr int NSLOT = frag_nslot<D>() < 3 ? 3 : frag_nslot<D>();
    static constexpr int SLOT = 3 * TR::MSZ + TR::VSZ;  // F | Q | fP | fm
    static constexpr int WPC = frag_warps<D>(SLOT, frag_cap<D>(10));
    static constexpr size_t WARP_SMEM = (size_t)NSLOT * SLOT * 8;
    static constexpr size_t SMEM = (size_t)WPC * WARP_SMEM;
};

template <int D, bool SMOOTH, bool ADJ>
__global__ void __launch_bounds__(FK3<D, SMOOTH, ADJ>::WPC * 32)
fk3_reverse(Params p, int L, long nchunks, const double* __restrict__ rstates, double* __restrict__ part) {
    using K = FK3<D, SMOOTH, ADJ>;
    using TR = typename K::TR;
    using M = typename TR::M;
    using VR = typename TR::VR;
    using VP = typename TR::VP;
    constexpr int MSZ = TR::MSZ, NSLOT = K::NSLOT, DD = D * D;
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
    typename TR::Cp cp;
    cp.init(lane);
    typename TR::St sp;
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
                TR::issue_vec(lane, sl + O_M, p.fms_in + row * D);
            } else {
                cp.issue(lane, sl + O_P, p.P0);
                if (p.m0 != nullptr) TR::issue_vec(lane, sl + O_M, p.m0);
                else TR::zero_vec(lane, sl + O_M);
            }
        }
        cp_commit();
    };
#pragma unroll 1
    for (int j = 0; j < PD; ++j) issue(j);
    const VP hP = TR::gld_vp(p.H, c);
    const VR hR = TR::gld_vr(p.H, r);
    const double Rv = p.R[0];
    const double gl = ADJ ? p.g[0] : 1.0;
    const double* st = rstates + (nchunks - 1 - chunk) * (2 * DD + 2 * D);
    VR dm = TR::gld_vr(st, r), lam = TR::gld_vr(st + D, r);
    M dP = TR::gld_mat(st + 2 * D, r, c), Lam = TR::gld_mat(st + 2 * D + DD, r, c);
    VR dHacc = TR::vzero();
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
          if (p.proj != nullptr) {
            // projected output only: H sm_k = h.m_k - h.(P_k lam_k),  H sP_k H^T = h.w - w.(Lam_k w) with w = P_k h
            // (four matrix-vector products instead of the two matrix products below, and 2 values stored instead of d^2 + d)
            const auto Pk = TR::in_op(sl + O_P, false, r, c);
            const VR w = TR::mv(Pk, hP);
            const VR v1 = TR::mv(Pk, TR::vr2vp(lam, c));
            const VR Lw = TR::mv(Lam, TR::vr2vp(w, c));
            const VR mk = TR::ld_vr(sl + O_M, r);
            double hw, wLw, hmk, hv1;
            TR::dots2(hR, w, w, Lw, lane, c, hw, wLw);
            TR::dots2(hR, mk, hR, v1, lane, c, hmk, hv1);
            if (lane == 0) *reinterpret_cast<double2*>(p.proj + 2 * k) = make_double2(hmk - hv1, hw - wLw);
          } else {
            // sm_k = m_k - P_k lam_k ; sP_k = P_k - P_k Lam_k P_k   (state entering from above)
            const auto Pk = TR::in_op(sl + O_P, false, r, c);
            const M Zt = TR::mulT(Pk, Lam, c);  // P_k Lam = (Lam P_k)^T
            const M PZ = TR::mulT(Pk, Zt, c);   // P_k (Lam P_k)
            TR::template gst_mat<false>(p.sPs + k * DD, TR::op_minus(Pk, PZ), 1.0, sp, r, c);
            const VR v1 = TR::mv(Pk, TR::vr2vp(lam, c));
            const VR mk = TR::ld_vr(sl + O_M, r);
            TR::gst_vr(p.sms + k * D, TR::vaxpy(-1.0, v1, mk), r, c);
          }
        }
        // forward quantities of step k
        const auto F = TR::in_op(sl, false, r, c);
        const auto Ft = TR::in_op(sl, true, r, c);
        const auto Pprev = TR::in_op(slp + O_P, false, r, c);
        const VR mprev = TR::ld_vr(slp + O_M, r);
        M Pp = TR::ld_sym(sl + MSZ, r, c);
        {
            const M T1 = TR::mulT(F, Pprev, c);
            TR::mmT(Pp, T1, F, c);
        }
        const VP mprevP = TR::ld_vp(slp + O_M, c);
        const VR mp = TR::mv(F, mprevP);
        VR u = TR::mv(Pp, hP);
        double hu, hm;
        TR::dots2(hR, u, hR, mp, lane, c, hu, hm);
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
                    dHacc = TR::vlin3(1.0, dHacc, 2.0 * sbar0, u, -rbar0, mp);
                }
                M dPp0 = TR::zero();
                {
                    const VR x1[1] = {TR::vscale(sbar0, hR)}, y1[1] = {hR};
                    const VP p1[1] = {hP};
                    TR::template rank_update<1>(dPp0, x1, y1, p1, c);
                }
                const VR dmp0 = TR::vscale(-rbar0, hR);
                TR::template gst_mat<false>(p.dQs + k * DD, dPp0, gl, sp, r, c);
                const M X = TR::mulT(dPp0, Ft, c);   // dPp0 F
                const M Xt = TR::mulT(Ft, dPp0, c);  // F^T dPp0
                M Y = TR::mulT(X, Pprev, c);
                {
                    const VR x1[1] = {TR::vscale(0.5, dmp0)}, y1[1] = {mprev};
                    const VP p1[1] = {mprevP};
                    TR::template rank_update<1>(Y, x1, y1, p1, c);
                }
                TR::template gst_mat<false>(p.dFs + k * DD, Y, 2.0 * gl, sp, r, c);
                M dP0 = TR::mulT(Xt, Ft, c);  // F^T dPp0 F
                // adjoint of the update on (m0, P0)
                u = TR::mv(Pprev, hP);
                double hu0, hm0;
                TR::dots2(hR, u, hR, mprev, lane, c, hu0, hm0);
                const double s0 = Rv + hu0, r0 = yk - hm0;
                if (obs) {
                    const VR Pu = TR::mv(dP, TR::vr2vp(u, c));
                    double udm, uPu;
                    TR::dots2(u, dm, u, Pu, lane, c, udm, uPu);
                    const double is0 = 1.0 / s0;
                    const double rbar = udm * is0;
                    const double sbar = (-udm * r0 + uPu) * is0 * is0;
                    const VR ut = TR::vlin3(r0 * is0, dm, -2.0 * is0, Pu, sbar, hR);
                    const VP utP = TR::vr2vp(ut, c);
                    const VR Pput = TR::mv(Pprev, utP);
                    dRacc += sbar;
                    dHacc = TR::vaxpy(sbar, u, dHacc);
                    dHacc = TR::vlin3(1.0, dHacc, 1.0, Pput, -rbar, mprev);
                    const VR x2[2] = {TR::vscale(0.5, ut), TR::vscale(0.5, hR)}, y2[2] = {hR, ut};
                    const VP p2[2] = {hP, utP};
                    TR::template rank_update<2>(dP, x2, y2, p2, c);
                }
                if (p.dP0 != nullptr) {
                    TR::madd(dP0, dP);
                    TR::template gst_mat<false>(p.dP0, dP0, gl, sp, r, c);
                }
            }
            __syncwarp();
            return;  // k == 0: nothing below this row
        }
        // measurement part (all terms vanish with is = 0 when nothing is observed)
        const double is = obs ? 1.0 / s : 0.0;
        const VP uP = TR::vr2vp(u, c);
        VR Pu = u, gv = u;
        if constexpr (ADJ) Pu = TR::mv(dP, uP);
        if constexpr (SMOOTH) gv = TR::mv(Lam, uP);
        double dv[4];
        {
            const VR* const xa[4] = {&u, &u, &u, &u};
            const VR* const xb[4] = {&dm, &Pu, &lam, &gv};
            TR::dots4(xa, xb, lane, c, dv);
        }
        const double udm = ADJ ? dv[0] : 0.0, uPu = ADJ ? dv[1] : 0.0, ulam = SMOOTH ? dv[2] : 0.0,
                     alpha = SMOOTH ? dv[3] : 0.0;
        const double rbar = (udm - rr) * is;
        const double sbar = (-udm * rr + uPu) * is * is + 0.5 * (rr * rr * is * is - is);
        VR dmp = dm, lt = lam;
        if constexpr (ADJ) {
            const VR ut = TR::vlin3(rr * is, dm, -2.0 * is, Pu, sbar, hR);
            dmp = TR::vaxpy(-rbar, hR, dm);
            const VP utP = TR::vr2vp(ut, c);
            const VR Pput = TR::mv(Pp, utP);
            dRacc += sbar;
            dHacc = TR::vaxpy(sbar, u, dHacc);
            dHacc = TR::vlin3(1.0, dHacc, 1.0, Pput, -rbar, mp);
            {
                const VR x2[2] = {TR::vscale(0.5, ut), TR::vscale(0.5, hR)}, y2[2] = {hR, ut};
                const VP p2[2] = {hP, utP};
                TR::template rank_update<2>(dP, x2, y2, p2, c);  // dPp = dP + sym(ut h^T)
            }
            TR::template gst_mat<false>(p.dQs + k * DD, dP, gl, sp, r, c);
            const M X = TR::mulT(dP, Ft, c);     // dPp F
            const M Xt = TR::mulT(Ft, dP, c);    // F^T dPp
            M Y = TR::mulT(X, Pprev, c);         // dPp F P_{k-1}
            {
                const VR x1[1] = {TR::vscale(0.5, dmp)}, y1[1] = {mprev};
                const VP p1[1] = {mprevP};
                TR::template rank_update<1>(Y, x1, y1, p1, c);
            }
            TR::template gst_mat<false>(p.dFs + k * DD, Y, 2.0 * gl, sp, r, c);
            dP = TR::mulT(Xt, Ft, c);            // F^T dPp F
            dm = TR::mv(Ft, TR::vr2vp(dmp, c));  // F^T dmp
        }
        if constexpr (SMOOTH) {
            const double cm = is + alpha * is * is;
            lt = TR::vaxpy(-(ulam + rr) * is, hR, lam);
            const VR xv = TR::vlin3(-is, gv, 0.5 * cm, hR, 0.0, hR);
            {
                const VR x2[2] = {hR, xv}, y2[2] = {xv, hR};
                const VP p2[2] = {TR::vp_for_rank(xv, c), hP};
                TR::template rank_update<2>(Lam, x2, y2, p2, c);  // Lt = Lam - (h g^T + g h^T)/s + h h^T (1/s + alpha/s^2)
            }
            const M X2t = TR::mulT(Ft, Lam, c);   // F^T Lt
            Lam = TR::mulT(X2t, Ft, c);           // F^T Lt F
            lam = TR::mv(Ft, TR::vr2vp(lt, c));
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
        TR::gst_vr(part + chunk * (1 + D) + 1, dHacc, r, c);
    }
}

}  // namespace frag
}  // namespace mid
}  // namespace pssgp
