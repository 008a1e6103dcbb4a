// Warp-level FP64 tensor-core discretisation for 5 <= d <= 32 (reference pssgp/kernels/base.py:29-47, _get_ssm):
//   A_k = expm(F dt_k)  (scaled Taylor polynomial with the precomputed matrix coefficients C_j of discretise.cu,
//                        then s squarings),      Q_k = Pinf - sym(A_k Pinf A_k^T).
// ONE WARP PER TIME STEP.  Every matrix lives in registers in the accumulator layout of mma.m8n8k4.f64 ("CF": lane
// (r, c) = (lane >> 2, lane & 3) holds M[8 a + r][8 b + 2c + {0,1}] of tile (a, b)); with the k-slot of lane c standing
// for the physical columns 2c + s of k-step s, a CF matrix is at once the A operand of X S^T and the B operand of its
// own transpose, so A P (P symmetric), (A P) A^T and A (A P)^T need no data movement.  The Horner evaluation reads the
// coefficient table from shared memory in the same tile-major layout (one conflict-free 128-bit load per tile and
// coefficient).  The CTA-cooperative kernel this replaces (discretise_generic_kernel: one CTA per step, a barrier
// after every product, 16-way bank conflicts for d = 16 / 24) ran at 2-9 % of the HBM roofline of its output.
#include "../../include/pssgp_b200.h"

#include <cuda_runtime.h>
#include <stdint.h>

#include "workspace.h"

namespace pssgp {

namespace {

constexpr int kDeg = 18;          // Taylor<double>::DEG of discretise.cu
constexpr int kMaxSquarings = 30;  // as in discretise.cu

#define DDEV __device__ __forceinline__

DDEV void dmma(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c[0]), "+d"(c[1])
                 : "d"(a), "d"(b));
}

template <int MT> struct Mat { double v[MT][MT][2]; };

// acc = X S^T; k-steps whose four columns 8 kt + s + {0,2,4,6} all lie in the zero padding (>= d) are skipped
template <int MT> DDEV void mmT(Mat<MT>& acc, const Mat<MT>& X, const Mat<MT>& S, int d) {
#pragma unroll
    for (int kt = 0; kt < MT; ++kt)
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            if (8 * kt + s < d) {
#pragma unroll
                for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                    for (int nt = 0; nt < MT; ++nt) dmma(acc.v[mt][nt], X.v[mt][kt][s], S.v[nt][kt][s]);
            }
        }
}
template <int MT> DDEV Mat<MT> mulT(const Mat<MT>& X, const Mat<MT>& S, int d) {
    Mat<MT> acc;
#pragma unroll
    for (int a = 0; a < MT; ++a)
#pragma unroll
        for (int b = 0; b < MT; ++b) acc.v[a][b][0] = acc.v[a][b][1] = 0.0;
    mmT<MT>(acc, X, S, d);
    return acc;
}

template <int MT> struct Cfg {
    static constexpr int WARPS = MT <= 2 ? 16 : (MT == 3 ? 12 : 8);
    static constexpr int TAB = (kDeg + 1) * MT * MT * 64;  // doubles: coefficient table, tile-major
    static constexpr int SCR = MT * MT * 64;               // doubles per warp: transpose scratch (squarings)
    static constexpr size_t SMEM = sizeof(double) * (size_t)(TAB + WARPS * SCR);
};

template <int MT>
__global__ void __launch_bounds__(Cfg<MT>::WARPS * 32)
disc_frag_kernel(const double* __restrict__ coef, const double* __restrict__ Pinf, int d, const double* __restrict__ dts,
                 long n, double* __restrict__ Fs, double* __restrict__ Qs) {
    using C = Cfg<MT>;
    extern __shared__ __align__(16) double dsm[];
    double* tab = dsm;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int r = lane >> 2, c = lane & 3;
    double* scr = dsm + C::TAB + w * C::SCR;
    const int dd = d * d;
    // coefficient table: C_j (row-major d x d in global memory) -> tile-major with zero padding
    for (int idx = threadIdx.x; idx < C::TAB; idx += blockDim.x) {
        const int j = idx / (MT * MT * 64), rem = idx - j * (MT * MT * 64);
        const int t = rem >> 6, e = rem & 63;
        const int row = 8 * (t / MT) + (e >> 3), col = 8 * (t % MT) + (e & 7);
        tab[idx] = (row < d && col < d) ? coef[8 + (size_t)j * dd + row * d + col] : 0.0;
    }
    const double normF = coef[0];
    // Pinf, symmetrised, in registers
    Mat<MT> P;
#pragma unroll
    for (int a = 0; a < MT; ++a)
#pragma unroll
        for (int b = 0; b < MT; ++b)
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                const int i = 8 * a + r, j = 8 * b + 2 * c + s;
                P.v[a][b][s] = (i < d && j < d) ? 0.5 * (Pinf[i * d + j] + Pinf[j * d + i]) : 0.0;
            }
    __syncthreads();
    // pairs (j, j + 1) are stored as one 16-byte piece when rows start on 16-byte boundaries
    const bool even = (d & 1) == 0 && ((((size_t)Fs) | ((size_t)Qs)) & 15) == 0;
    const long stride = (long)gridDim.x * C::WARPS;
    for (long k = (long)blockIdx.x * C::WARPS + w; k < n; k += stride) {
        const double dt = dts[k];
        double x = normF * fabs(dt);
        int s = 0;
        while (x > 1.0 && s < kMaxSquarings) {
            x *= 0.5;
            ++s;
        }
        if (dt < 0.0) x = -x;
        // Horner: A = sum_j C_j x^j
        Mat<MT> A;
#pragma unroll
        for (int a = 0; a < MT; ++a)
#pragma unroll
            for (int b = 0; b < MT; ++b) {
                const double2 v = *reinterpret_cast<const double2*>(tab + (kDeg * MT * MT + a * MT + b) * 64 + lane * 2);
                A.v[a][b][0] = v.x;
                A.v[a][b][1] = v.y;
            }
#pragma unroll 2
        for (int p = kDeg - 1; p >= 0; --p) {
#pragma unroll
            for (int a = 0; a < MT; ++a)
#pragma unroll
                for (int b = 0; b < MT; ++b) {
                    const double2 v = *reinterpret_cast<const double2*>(tab + (p * MT * MT + a * MT + b) * 64 + lane * 2);
                    A.v[a][b][0] = fma(A.v[a][b][0], x, v.x);
                    A.v[a][b][1] = fma(A.v[a][b][1], x, v.y);
                }
        }
        // squarings A <- A A = A (A^T)^T: the transpose comes back from the warp's scratch tile
        for (int q = 0; q < s; ++q) {
#pragma unroll
            for (int a = 0; a < MT; ++a)
#pragma unroll
                for (int b = 0; b < MT; ++b)
                    *reinterpret_cast<double2*>(scr + (a * MT + b) * 64 + lane * 2) = make_double2(A.v[a][b][0], A.v[a][b][1]);
            __syncwarp();
            Mat<MT> At;
#pragma unroll
            for (int a = 0; a < MT; ++a)
#pragma unroll
                for (int b = 0; b < MT; ++b) {
                    At.v[a][b][0] = scr[(b * MT + a) * 64 + (2 * c) * 8 + r];
                    At.v[a][b][1] = scr[(b * MT + a) * 64 + (2 * c + 1) * 8 + r];
                }
            __syncwarp();
            A = mulT<MT>(A, At, d);
        }
        // X = A P,  M1 = X A^T,  M2 = A X^T = M1^T,  Q = P - (M1 + M2) / 2.  MT <= 3: two accumulators, so that Q is
        // symmetric to the last bit; MT = 4: one accumulator takes both products (register budget), symmetric to rounding
        const Mat<MT> X = mulT<MT>(A, P, d);
        Mat<MT> M1 = mulT<MT>(X, A, d);
        Mat<MT> M2;
        if constexpr (MT <= 3) {
            M2 = mulT<MT>(A, X, d);
        } else {
            mmT<MT>(M1, A, X, d);
        }
        double* fo = Fs + k * dd;
        double* qo = Qs + k * dd;
#pragma unroll
        for (int a = 0; a < MT; ++a)
#pragma unroll
            for (int b = 0; b < MT; ++b) {
                const int i = 8 * a + r, j = 8 * b + 2 * c;
                double q0, q1;
                if constexpr (MT <= 3) {
                    q0 = P.v[a][b][0] - 0.5 * (M1.v[a][b][0] + M2.v[a][b][0]);
                    q1 = P.v[a][b][1] - 0.5 * (M1.v[a][b][1] + M2.v[a][b][1]);
                } else {
                    q0 = fma(-0.5, M1.v[a][b][0], P.v[a][b][0]);
                    q1 = fma(-0.5, M1.v[a][b][1], P.v[a][b][1]);
                }
                if (i < d && j < d) {
                    if (even) {  // j and d even: 16-byte aligned pair inside the row
                        __stcs(reinterpret_cast<double2*>(fo + i * d + j), make_double2(A.v[a][b][0], A.v[a][b][1]));
                        __stcs(reinterpret_cast<double2*>(qo + i * d + j), make_double2(q0, q1));
                    } else {
                        __stcs(fo + i * d + j, A.v[a][b][0]);
                        __stcs(qo + i * d + j, q0);
                        if (j + 1 < d) {
                            __stcs(fo + i * d + j + 1, A.v[a][b][1]);
                            __stcs(qo + i * d + j + 1, q1);
                        }
                    }
                }
            }
    }
}

template <int MT>
int launch(pssgp_handle* h, int64_t n, int d, const double* coef, const double* Pinf, const double* dts, double* Fs,
           double* Qs, cudaStream_t st) {
    using C = Cfg<MT>;
    cudaError_t e = cudaFuncSetAttribute(disc_frag_kernel<MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
    if (e != cudaSuccess) return set_err(PSSGP_ERR_CUDA, "discretise: %s", cudaGetErrorString(e));
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, disc_frag_kernel<MT>, C::WARPS * 32, C::SMEM);
    if (per_sm < 1) per_sm = 1;
    long grid = (long)h->num_sms * per_sm;
    const long need = (n + C::WARPS - 1) / C::WARPS;
    if (grid > need) grid = need;
    PSSGP_LAUNCH(h, "discretise", st,
                 (disc_frag_kernel<MT><<<(unsigned)grid, C::WARPS * 32, C::SMEM, st>>>(coef, Pinf, d, dts, (long)n, Fs, Qs)));
    return check_launch(h, "discretise", 2);  // + the coefficient set-up launch
}

// ------------------------------------------------------------------------------------------------------------------
// Backward (d <= 24): per step  T1 = dQ A,  dA_h = dA - 2 T1 P (then back through the squarings),
// dPinf += dQ - A^T dQ A,  W_p += x^p dA_h  (the moment matrices of discretise.cu; dF is assembled from them by
// discretise_bwd_final_kernel).  The products run warp-per-step on the tensor cores as in the forward kernel; each
// warp leaves its dA_h in a shared-memory slot, and after a CTA barrier the moments are accumulated one THREAD PER
// MATRIX ELEMENT in registers (G = NT / d^2 thread groups take the slots round-robin; every group is one partial;
// d > 16: up to three elements per thread, moments in shared memory).
// ------------------------------------------------------------------------------------------------------------------
template <int MT> struct BCfg {
    static constexpr int WARPS = MT == 1 ? 16 : (MT == 2 ? 12 : 8);
    static constexpr int NT = WARPS * 32;
    static constexpr int TAB = (kDeg + 1) * MT * MT * 64;
    static constexpr int SCR = MT * MT * 64;
    // MT = 3 (d = 17..24): a thread owns up to three elements, whose 18 moments each live in shared memory
    static constexpr int MOM = MT <= 2 ? 0 : kDeg * MT * MT * 64;
    static constexpr size_t SMEM = sizeof(double) * (size_t)(TAB + WARPS * SCR + 2 * WARPS + MOM);
};

template <int MT> DDEV Mat<MT> transp(const Mat<MT>& A, double* scr, int lane, int r, int c) {
#pragma unroll
    for (int a = 0; a < MT; ++a)
#pragma unroll
        for (int b = 0; b < MT; ++b)
            *reinterpret_cast<double2*>(scr + (a * MT + b) * 64 + lane * 2) = make_double2(A.v[a][b][0], A.v[a][b][1]);
    __syncwarp();
    Mat<MT> At;
#pragma unroll
    for (int a = 0; a < MT; ++a)
#pragma unroll
        for (int b = 0; b < MT; ++b) {
            At.v[a][b][0] = scr[(b * MT + a) * 64 + (2 * c) * 8 + r];
            At.v[a][b][1] = scr[(b * MT + a) * 64 + (2 * c + 1) * 8 + r];
        }
    __syncwarp();
    return At;
}

template <int MT>
__global__ void __launch_bounds__(BCfg<MT>::NT)
disc_bwd_frag_kernel(const double* __restrict__ coef, const double* __restrict__ Pinf, int d, const double* __restrict__ dts,
                     long n, const double* __restrict__ Fs, const double* __restrict__ dFs, const double* __restrict__ dQs,
                     double* __restrict__ part) {
    using C = BCfg<MT>;
    extern __shared__ __align__(16) double dsm[];
    double* tab = dsm;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int r = lane >> 2, c = lane & 3;
    double* slot = dsm + C::TAB + w * C::SCR;
    double* xs = dsm + C::TAB + C::WARPS * C::SCR;   // x of the step each warp holds (NaN: no step)
    double* pms = xs + C::WARPS;                      // highest moment degree that matters for that x
    double* mom = pms + C::WARPS;                     // MT = 3: [kDeg][d^2] moment accumulators
    const int dd = d * d;
    if constexpr (MT > 2)
        for (int idx = threadIdx.x; idx < kDeg * dd; idx += blockDim.x) mom[idx] = 0.0;
    for (int idx = threadIdx.x; idx < C::TAB; idx += blockDim.x) {
        const int j = idx / (MT * MT * 64), rem = idx - j * (MT * MT * 64);
        const int t = rem >> 6, e = rem & 63;
        const int row = 8 * (t / MT) + (e >> 3), col = 8 * (t % MT) + (e & 7);
        tab[idx] = (row < d && col < d) ? coef[8 + (size_t)j * dd + row * d + col] : 0.0;
    }
    const double normF = coef[0];
    Mat<MT> P, acc0;
#pragma unroll
    for (int a = 0; a < MT; ++a)
#pragma unroll
        for (int b = 0; b < MT; ++b)
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                const int i = 8 * a + r, j = 8 * b + 2 * c + s;
                P.v[a][b][s] = (i < d && j < d) ? 0.5 * (Pinf[i * d + j] + Pinf[j * d + i]) : 0.0;
                acc0.v[a][b][s] = 0.0;
            }
    // moment phase: thread -> (group g, element e)
    const int G = MT <= 2 ? C::NT / dd : 1;
    const int g = MT <= 2 ? threadIdx.x / dd : 0, e = threadIdx.x - g * dd;
    const bool momt = g < G && e < dd;
    const int ei = e / d, ej = e - ei * d;
    const int eoff = ((ei >> 3) * MT + (ej >> 3)) * 64 + (ei & 7) * 8 + (ej & 7);
    double acc[kDeg];
#pragma unroll
    for (int p = 0; p < kDeg; ++p) acc[p] = 0.0;
    __syncthreads();
    const long stride = (long)gridDim.x * C::WARPS;
    const long rounds = (n + stride - 1) / stride;
    for (long rd = 0; rd < rounds; ++rd) {
        const long k = rd * stride + (long)blockIdx.x * C::WARPS + w;
        if (k < n) {  // warp-uniform
            const double dt = dts[k];
            double x = normF * fabs(dt);
            int s = 0;
            while (x > 1.0 && s < kMaxSquarings) {
                x *= 0.5;
                ++s;
            }
            if (dt < 0.0) x = -x;
            const double* Ak = Fs + k * dd;
            const double* dAk = dFs + k * dd;
            const double* dQk = dQs + k * dd;
            Mat<MT> At, da, dQ;
#pragma unroll
            for (int a = 0; a < MT; ++a)
#pragma unroll
                for (int b = 0; b < MT; ++b)
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const int i = 8 * a + r, j = 8 * b + 2 * c + q;
                        const bool ok = i < d && j < d;
                        At.v[a][b][q] = ok ? __ldg(Ak + j * d + i) : 0.0;
                        da.v[a][b][q] = ok ? __ldg(dAk + i * d + j) : 0.0;
                        dQ.v[a][b][q] = ok ? 0.5 * (__ldg(dQk + i * d + j) + __ldg(dQk + j * d + i)) : 0.0;
                    }
            const Mat<MT> T1 = mulT<MT>(dQ, At, d);     // dQ A
            const Mat<MT> T1t = mulT<MT>(At, dQ, d);    // A^T dQ
            const Mat<MT> TP = mulT<MT>(T1, P, d);      // dQ A P
            const Mat<MT> AQA = mulT<MT>(T1t, At, d);   // A^T dQ A
#pragma unroll
            for (int a = 0; a < MT; ++a)
#pragma unroll
                for (int b = 0; b < MT; ++b)
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        da.v[a][b][q] = fma(-2.0, TP.v[a][b][q], da.v[a][b][q]);
                        acc0.v[a][b][q] += dQ.v[a][b][q] - AQA.v[a][b][q];
                    }
            if (s > 0) {  // back through A = Ah^(2^s): dA_lvl = A_lvl^T dA + dA A_lvl^T, A_lvl = Ah^(2^lvl) recomputed
                Mat<MT> Ah;
#pragma unroll
                for (int a = 0; a < MT; ++a)
#pragma unroll
                    for (int b = 0; b < MT; ++b) {
                        const double2 v = *reinterpret_cast<const double2*>(tab + (kDeg * MT * MT + a * MT + b) * 64 + lane * 2);
                        Ah.v[a][b][0] = v.x;
                        Ah.v[a][b][1] = v.y;
                    }
#pragma unroll 2
                for (int p = kDeg - 1; p >= 0; --p) {
#pragma unroll
                    for (int a = 0; a < MT; ++a)
#pragma unroll
                        for (int b = 0; b < MT; ++b) {
                            const double2 v = *reinterpret_cast<const double2*>(tab + (p * MT * MT + a * MT + b) * 64 + lane * 2);
                            Ah.v[a][b][0] = fma(Ah.v[a][b][0], x, v.x);
                            Ah.v[a][b][1] = fma(Ah.v[a][b][1], x, v.y);
                        }
                }
                for (int lvl = s - 1; lvl >= 0; --lvl) {
                    Mat<MT> cur = Ah;
                    for (int q = 0; q < lvl; ++q) {
                        const Mat<MT> ct = transp<MT>(cur, slot, lane, r, c);
                        cur = mulT<MT>(cur, ct, d);
                    }
                    const Mat<MT> curT = transp<MT>(cur, slot, lane, r, c);
                    const Mat<MT> daT = transp<MT>(da, slot, lane, r, c);
                    Mat<MT> nd = mulT<MT>(curT, daT, d);   // A_lvl^T dA
                    mmT<MT>(nd, da, cur, d);               // + dA A_lvl^T
                    da = nd;
                }
            }
#pragma unroll
            for (int a = 0; a < MT; ++a)
#pragma unroll
                for (int b = 0; b < MT; ++b)
                    *reinterpret_cast<double2*>(slot + (a * MT + b) * 64 + lane * 2) = make_double2(da.v[a][b][0], da.v[a][b][1]);
            if (lane == 0) {
                // degrees whose weight x^(p-1)/p! is negligible are skipped (as in discretise_bwd_generic_kernel)
                const double ax = fabs(x);
                int pmax = kDeg;
                double wgt = 1.0;
                for (int p = 1; p <= kDeg; ++p) {
                    if (p > 1) wgt *= ax / (double)p;
                    if (wgt < 1e-19) {
                        pmax = p - 1;
                        break;
                    }
                }
                xs[w] = x;
                pms[w] = (double)(pmax < 1 ? 1 : pmax);
            }
        } else if (lane == 0) {
            pms[w] = 0.0;
        }
        __syncthreads();
        if constexpr (MT <= 2) {
            if (momt) {
                for (int ws = g; ws < C::WARPS; ws += G) {
                    const int pmax = (int)pms[ws];
                    if (pmax == 0) continue;
                    const double x = xs[ws];
                    const double v = dsm[C::TAB + ws * C::SCR + eoff];
                    double xp = 1.0;
#pragma unroll
                    for (int p = 1; p <= kDeg; ++p) {
                        xp *= x;
                        if (p <= pmax) acc[p - 1] = fma(xp, v, acc[p - 1]);
                    }
                }
            }
        } else {
            // the round's contributions are summed in registers, then added to the shared-memory moments once
            for (int e2 = threadIdx.x; e2 < dd; e2 += C::NT) {
                const int i2 = e2 / d, j2 = e2 - i2 * d;
                const int off2 = ((i2 >> 3) * MT + (j2 >> 3)) * 64 + (i2 & 7) * 8 + (j2 & 7);
#pragma unroll
                for (int p = 0; p < kDeg; ++p) acc[p] = 0.0;
                int pall = 0;
                for (int ws = 0; ws < C::WARPS; ++ws) {
                    const int pmax = (int)pms[ws];
                    if (pmax == 0) continue;
                    pall = pmax > pall ? pmax : pall;
                    const double x = xs[ws];
                    const double v = dsm[C::TAB + ws * C::SCR + off2];
                    double xp = 1.0;
#pragma unroll
                    for (int p = 1; p <= kDeg; ++p) {
                        xp *= x;
                        if (p <= pmax) acc[p - 1] = fma(xp, v, acc[p - 1]);
                    }
                }
#pragma unroll
                for (int p = 1; p <= kDeg; ++p)
                    if (p <= pall) mom[(p - 1) * dd + e2] += acc[p - 1];
            }
        }
        __syncthreads();
    }
    // partial g of this CTA: p = 0 (dPinf) is the sum of the warps' accumulators (fixed order), carried by group 0
#pragma unroll
    for (int a = 0; a < MT; ++a)
#pragma unroll
        for (int b = 0; b < MT; ++b)
            *reinterpret_cast<double2*>(slot + (a * MT + b) * 64 + lane * 2) = make_double2(acc0.v[a][b][0], acc0.v[a][b][1]);
    __syncthreads();
    if constexpr (MT <= 2) {
        if (momt) {
            double* mine = part + ((size_t)blockIdx.x * G + g) * (size_t)(kDeg + 1) * dd;
            double p0 = 0.0;
            if (g == 0)
                for (int ws = 0; ws < C::WARPS; ++ws) p0 += dsm[C::TAB + ws * C::SCR + eoff];
            mine[e] = p0;
#pragma unroll
            for (int p = 1; p <= kDeg; ++p) mine[(size_t)p * dd + e] = acc[p - 1];
        }
    } else {
        double* mine = part + (size_t)blockIdx.x * (size_t)(kDeg + 1) * dd;
        for (int e2 = threadIdx.x; e2 < dd; e2 += C::NT) {
            const int i2 = e2 / d, j2 = e2 - i2 * d;
            const int off2 = ((i2 >> 3) * MT + (j2 >> 3)) * 64 + (i2 & 7) * 8 + (j2 & 7);
            double p0 = 0.0;
            for (int ws = 0; ws < C::WARPS; ++ws) p0 += dsm[C::TAB + ws * C::SCR + off2];
            mine[e2] = p0;
            for (int p = 1; p <= kDeg; ++p) mine[(size_t)p * dd + e2] = mom[(p - 1) * dd + e2];
        }
    }
}

template <int MT>
int launch_bwd(pssgp_handle* h, int64_t n, int d, const double* coef, const double* Pinf, const double* dts,
               const double* Fs, const double* dFs, const double* dQs, double* part, long grid, cudaStream_t st) {
    using C = BCfg<MT>;
    cudaError_t e = cudaFuncSetAttribute(disc_bwd_frag_kernel<MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
    if (e != cudaSuccess) return set_err(PSSGP_ERR_CUDA, "discretise_backward: %s", cudaGetErrorString(e));
    PSSGP_LAUNCH(h, "discretise_bwd", st,
                 (disc_bwd_frag_kernel<MT><<<(unsigned)grid, C::NT, C::SMEM, st>>>(coef, Pinf, d, dts, (long)n, Fs, dFs, dQs, part)));
    return check_launch(h, "discretise_bwd", 1);
}

}  // namespace

// coef: the table of taylor_setup_kernel (discretise.cu): [0] = ||F||_1, C_j at 8 + j d^2.
int discretise_frag_f64(pssgp_handle* h, int64_t n, int d, const double* coef, const double* Pinf, const double* dts,
                        double* Fs, double* Qs, cudaStream_t st) {
    if (d <= 8) return launch<1>(h, n, d, coef, Pinf, dts, Fs, Qs, st);
    if (d <= 16) return launch<2>(h, n, d, coef, Pinf, dts, Fs, Qs, st);
    if (d <= 24) return launch<3>(h, n, d, coef, Pinf, dts, Fs, Qs, st);
    return launch<4>(h, n, d, coef, Pinf, dts, Fs, Qs, st);
}


// Backward partials for d <= 24 (see disc_bwd_frag_kernel).  *grid_out CTAs, each writing *per_cta partials of
// (DEG + 1) d^2 doubles into `part` (query with part == nullptr to size the buffer).
int discretise_bwd_frag_f64(pssgp_handle* h, int64_t n, int d, const double* coef, const double* Pinf, const double* dts,
                            const double* Fs, const double* dFs, const double* dQs, double* part, long* grid_out,
                            int* per_cta, cudaStream_t st) {
    const int nt = d <= 8 ? BCfg<1>::NT : (d <= 16 ? BCfg<2>::NT : BCfg<3>::NT);
    const int warps = nt / 32;
    int per_sm = 1;
    if (d > 16) {
        cudaFuncSetAttribute(disc_bwd_frag_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BCfg<3>::SMEM);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, disc_bwd_frag_kernel<3>, nt, BCfg<3>::SMEM);
    } else if (d <= 8) {
        cudaFuncSetAttribute(disc_bwd_frag_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BCfg<1>::SMEM);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, disc_bwd_frag_kernel<1>, nt, BCfg<1>::SMEM);
    } else {
        cudaFuncSetAttribute(disc_bwd_frag_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BCfg<2>::SMEM);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, disc_bwd_frag_kernel<2>, nt, BCfg<2>::SMEM);
    }
    if (per_sm < 1) per_sm = 1;
    long grid = (long)h->num_sms * per_sm;
    const long need = (n + warps - 1) / warps;
    if (grid > need) grid = need;
    *grid_out = grid;
    *per_cta = d <= 16 ? nt / (d * d) : 1;
    if (!part) return PSSGP_OK;
    if (d <= 8) return launch_bwd<1>(h, n, d, coef, Pinf, dts, Fs, dFs, dQs, part, grid, st);
    if (d <= 16) return launch_bwd<2>(h, n, d, coef, Pinf, dts, Fs, dFs, dQs, part, grid, st);
    return launch_bwd<3>(h, n, d, coef, Pinf, dts, Fs, dFs, dQs, part, grid, st);
}

}  // namespace pssgp
