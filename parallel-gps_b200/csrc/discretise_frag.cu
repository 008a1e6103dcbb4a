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
    int per_sm = (int)((size_t)220 * 1024 / (C::SMEM + 1024));
    const int by_threads = 2048 / (C::WARPS * 32);
    if (per_sm > by_threads) per_sm = by_threads;
    if (per_sm < 1) per_sm = 1;
    long grid = (long)h->num_sms * per_sm;
    const long need = (n + C::WARPS - 1) / C::WARPS;
    if (grid > need) grid = need;
    PSSGP_LAUNCH(h, "discretise", st,
                 (disc_frag_kernel<MT><<<(unsigned)grid, C::WARPS * 32, C::SMEM, st>>>(coef, Pinf, d, dts, (long)n, Fs, Qs)));
    return check_launch(h, "discretise", 2);  // + the coefficient set-up launch
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

}  // namespace pssgp
