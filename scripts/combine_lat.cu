// Microbenchmark (tuning aid): latency of one associative combine / apply of each algebra, as a dependent chain.
#include <cstdio>
#include <cuda_runtime.h>
#include "../parallel-gps_b200/csrc/filter_small.cuh"
#include "../parallel-gps_b200/csrc/smoother_small.cuh"
#include "../parallel-gps_b200/csrc/adjoint_small.cuh"
using namespace pssgp;

template <typename Alg, bool SHFL>
__global__ void chain(const double* in, double* out, long long* cyc, int iters) {
    using T = double;
    T a[Alg::NAGG], b[Alg::NAGG];
    for (int e = 0; e < Alg::NAGG; ++e) { a[e] = in[e]; b[e] = in[Alg::NAGG + e]; }
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
        T o[Alg::NAGG], r[Alg::NAGG];
        if (SHFL) {
#pragma unroll
            for (int e = 0; e < Alg::NAGG; ++e) o[e] = shfl_up_t(a[e], 1);
        } else {
#pragma unroll
            for (int e = 0; e < Alg::NAGG; ++e) o[e] = b[e];
        }
        Alg::combine(o, a, r);
#pragma unroll
        for (int e = 0; e < Alg::NAGG; ++e) a[e] = r[e];
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    for (int e = 0; e < Alg::NAGG; ++e) out[threadIdx.x * Alg::NAGG + e] = a[e];
}

template <typename Alg>
void run(const char* name, const double* d_in, double* d_out, long long* d_c) {
    long long hc; const int iters = 200;
    for (int warps : {1, 4, 8}) {
        chain<Alg, false><<<1, 32 * warps>>>(d_in, d_out, d_c, iters); cudaDeviceSynchronize();
        cudaMemcpy(&hc, d_c, 8, cudaMemcpyDeviceToHost);
        printf("%s combine, %d warps: %.0f cycles per combine", name, warps, (double)hc / iters);
        chain<Alg, true><<<1, 32 * warps>>>(d_in, d_out, d_c, iters); cudaDeviceSynchronize();
        cudaMemcpy(&hc, d_c, 8, cudaMemcpyDeviceToHost);
        printf("   with shuffles: %.0f\n", (double)hc / iters);
    }
}

int main() {
    double h[128];
    // a benign pair of elements: A ~ 0.9 I, C, J small SPD
    for (int i = 0; i < 128; ++i) h[i] = 0.0;
    using FA = FilterAlg<double, 3>;
    for (int rep = 0; rep < 2; ++rep) {
        double* a = h + rep * 64;
        for (int i = 0; i < 3; ++i) { a[FA::oA + i * 3 + i] = 0.9; a[FA::ob + i] = 0.1 * i; a[FA::oE + i] = 0.05; }
        a[FA::oC + 0] = 0.1; a[FA::oC + 2] = 0.1; a[FA::oC + 5] = 0.1; a[FA::oC + 1] = 0.01;
        a[FA::oJ + 0] = 0.2; a[FA::oJ + 2] = 0.2; a[FA::oJ + 5] = 0.2; a[FA::oJ + 3] = 0.02;
    }
    double *d_in, *d_out; long long* d_c;
    cudaMalloc(&d_in, sizeof(h)); cudaMalloc(&d_out, 8 * 64 * 1024); cudaMalloc(&d_c, 8);
    // pack: second element must start at NAGG
    double hin[128];
    for (int i = 0; i < 64; ++i) { hin[i] = h[i]; }
    for (int i = 0; i < 64 - 33; ++i) hin[33 + i] = h[64 + i];
    cudaMemcpy(d_in, hin, sizeof(hin), cudaMemcpyHostToDevice);
    run<FilterAlg<double, 3>>("filter  ", d_in, d_out, d_c);
    run<SmootherAlg<double, 3>>("smoother", d_in, d_out, d_c);
    run<AdjointAlg<double, 3>>("adjoint ", d_in, d_out, d_c);
    return 0;
}
