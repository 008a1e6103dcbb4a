// Microbenchmark (tuning aid for the d > 4 kernels): per-SM issue rates on B200 of the instructions a
// lane-distributed d x d FP64 product can be built from:
//   DFMA, DMMA (mma.sync.m8n8k4.f64), SHFL.IDX (64-bit = 2 x 32-bit), LDS.128 broadcast (G distinct
//   addresses per warp), LDS.64 all-distinct, and DFMA fed by LDS.128 broadcast at 1 load : 2 DFMA.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int ILP> __global__ void k_dfma(double* out, long long* cyc, int iters) {
    double a = out[0], b = out[1], x[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) x[k] = out[2] + k;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i)
#pragma unroll
        for (int k = 0; k < ILP; ++k) x[k] = fma(a, x[k], b);
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    double s = 0;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s += x[k];
    out[8 + threadIdx.x + blockIdx.x * blockDim.x] = s;
}

template <int ILP> __global__ void k_dmma(double* out, long long* cyc, int iters) {
    double a = out[0] + threadIdx.x * 1e-9, b = out[1], c[ILP][2];
#pragma unroll
    for (int k = 0; k < ILP; ++k) c[k][0] = c[k][1] = out[2] + k;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i)
#pragma unroll
        for (int k = 0; k < ILP; ++k) dmma(c[k][0], c[k][1], a, b);
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    double s = 0;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s += c[k][0] + c[k][1];
    out[8 + threadIdx.x + blockIdx.x * blockDim.x] = s;
}

// DMMA and DFMA interleaved 1 : NF (do they share the pipe?)
template <int NF> __global__ void k_mix(double* out, long long* cyc, int iters) {
    double a = out[0] + threadIdx.x * 1e-9, b = out[1], c[4][2], x[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) c[k][0] = c[k][1] = out[2] + k;
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = out[2] + k;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            dmma(c[k][0], c[k][1], a, b);
#pragma unroll
            for (int f = 0; f < NF; ++f) x[(k * NF + f) % 8] = fma(a, x[(k * NF + f) % 8], b);
        }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    double s = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) s += c[k][0] + c[k][1];
#pragma unroll
    for (int k = 0; k < 8; ++k) s += x[k];
    out[8 + threadIdx.x + blockIdx.x * blockDim.x] = s;
}

__global__ void k_shfl(double* out, long long* cyc, int iters) {
    double x[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = out[2] + k + threadIdx.x;
    const int src = (threadIdx.x & 24) | ((threadIdx.x + 1) & 7);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i)
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] = __shfl_sync(0xffffffffu, x[k], src);
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    double s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += x[k];
    out[8 + threadIdx.x + blockIdx.x * blockDim.x] = s;
}

// LDS.128 with G distinct 16-byte addresses per warp (lanes of a group read the same address)
template <int GL> __global__ void k_lds128(double* out, long long* cyc, int iters) {
    extern __shared__ __align__(16) double sm[];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = out[2] + i;
    __syncthreads();
    const int grp = (threadIdx.x & 31) / GL, w = threadIdx.x >> 5;
    const double2* base = reinterpret_cast<const double2*>(sm) + w * 64 + grp * 9;  // odd pitch between groups
    double s0 = 0, s1 = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            double vx, vy;
            asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(vx), "=d"(vy) : "r"((unsigned)__cvta_generic_to_shared(base + ((k + i) & 7))));
            s0 += vx;
            s1 += vy;
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[8 + threadIdx.x + blockIdx.x * blockDim.x] = s0 + s1;
}

// the inner loop of a row-distributed product: 1 LDS.128 broadcast feeds 2*R DFMA (R rows of C per lane)
template <int GL, int R> __global__ void k_ldsfma(double* out, long long* cyc, int iters) {
    extern __shared__ __align__(16) double sm[];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = out[1] * (1 + (i & 7));
    __syncthreads();
    const int grp = (threadIdx.x & 31) / GL, w = threadIdx.x >> 5;
    const double2* base = reinterpret_cast<const double2*>(sm) + w * 64 + grp * 9;
    double c[R][8], a[R][4];
#pragma unroll
    for (int r = 0; r < R; ++r) {
#pragma unroll
        for (int k = 0; k < 8; ++k) c[r][k] = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) a[r][k] = out[0] + k + r + threadIdx.x * 1e-9;
    }
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {      // 4 rows of B, each row = 8 doubles = 4 LDS.128
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const double2 v = base[(k * 4 + j + i) & 7];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    c[r][2 * j] = fma(a[r][k], v.x, c[r][2 * j]);
                    c[r][2 * j + 1] = fma(a[r][k], v.y, c[r][2 * j + 1]);
                }
            }
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    double s = 0;
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int k = 0; k < 8; ++k) s += c[r][k];
    out[8 + threadIdx.x + blockIdx.x * blockDim.x] = s;
}

// LDS.64, every lane its own address (odd pitch: conflict-free)
__global__ void k_lds64(double* out, long long* cyc, int iters) {
    extern __shared__ __align__(16) double sm[];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = out[2] + i;
    __syncthreads();
    const double* base = sm + (threadIdx.x & 255) * 9;
    double s = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) s += *(volatile const double*)(base + ((k + i) & 7));
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[8 + threadIdx.x + blockIdx.x * blockDim.x] = s;
}

static double* d;
static long long* c;
template <class F> static double run(F f, int warps, size_t smem = 0) {
    f(1, 32 * warps, smem);
    cudaDeviceSynchronize();
    f(1, 32 * warps, smem);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return -1; }
    long long hc;
    cudaMemcpy(&hc, c, 8, cudaMemcpyDeviceToHost);
    return (double)hc;
}

int main() {
    cudaMalloc(&d, 8 * (8 + 148 * 8 * 1024));
    cudaMalloc(&c, 8 * 4096);
    double h[3] = {0.999, 0.001, 1.0};
    cudaMemcpy(d, h, 24, cudaMemcpyHostToDevice);
    const int iters = 4000;
    const size_t SM = 72 * 1024;
    cudaFuncSetAttribute(k_lds128<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM);
    cudaFuncSetAttribute(k_lds128<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM);
    cudaFuncSetAttribute(k_lds128<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM);
    cudaFuncSetAttribute(k_lds128<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM);
    cudaFuncSetAttribute(k_lds64, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM);
    cudaFuncSetAttribute(k_ldsfma<8, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM);
    cudaFuncSetAttribute(k_ldsfma<6, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM);
    cudaFuncSetAttribute(k_ldsfma<8, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM);
    cudaFuncSetAttribute(k_ldsfma<32, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM);
    for (int w : {1, 4, 8, 16, 32}) {
        double cy = run([&](int g, int b, size_t s) { k_dfma<8><<<g, b, s>>>(d, c, iters); }, w);
        printf("DFMA ilp8      %2d warps: %.3f warp-instr/clk/SM\n", w, iters * 8.0 * w / cy);
    }
    for (int w : {1, 4, 8, 16, 32}) {
        double cy = run([&](int g, int b, size_t s) { k_dmma<1><<<g, b, s>>>(d, c, iters); }, w);
        printf("DMMA ilp1      %2d warps: %.4f warp-instr/clk/SM  (%.1f clk latency if 1 warp)  = %.1f FMA/clk/SM\n", w,
               iters * 1.0 * w / cy, cy / iters, iters * 1.0 * w / cy * 256);
    }
    for (int w : {1, 4, 8, 16, 32}) {
        double cy = run([&](int g, int b, size_t s) { k_dmma<6><<<g, b, s>>>(d, c, iters); }, w);
        printf("DMMA ilp6      %2d warps: %.4f warp-instr/clk/SM = %.1f FMA/clk/SM\n", w, iters * 6.0 * w / cy,
               iters * 6.0 * w / cy * 256);
    }
    for (int w : {4, 8, 16}) {
        double cy = run([&](int g, int b, size_t s) { k_mix<2><<<g, b, s>>>(d, c, iters); }, w);
        printf("DMMA+2 DFMA    %2d warps: %.1f FMA/clk/SM total\n", w, iters * 4.0 * w / cy * (256 + 2 * 32));
        cy = run([&](int g, int b, size_t s) { k_mix<8><<<g, b, s>>>(d, c, iters); }, w);
        printf("DMMA+8 DFMA    %2d warps: %.1f FMA/clk/SM total\n", w, iters * 4.0 * w / cy * (256 + 8 * 32));
    }
    for (int w : {4, 8, 16, 32}) {
        double cy = run([&](int g, int b, size_t s) { k_shfl<<<g, b, s>>>(d, c, iters); }, w);
        printf("SHFL (64-bit)  %2d warps: %.3f 64-bit shuffles/clk/SM (= %.3f SHFL.32)\n", w, iters * 8.0 * w / cy,
               2 * iters * 8.0 * w / cy);
    }
    for (int w : {4, 8, 16, 32}) {
        double c32 = run([&](int g, int b, size_t s) { k_lds128<32><<<g, b, s>>>(d, c, iters); }, w, SM);
        double c8 = run([&](int g, int b, size_t s) { k_lds128<8><<<g, b, s>>>(d, c, iters); }, w, SM);
        double c6 = run([&](int g, int b, size_t s) { k_lds128<6><<<g, b, s>>>(d, c, iters); }, w, SM);
        double c1 = run([&](int g, int b, size_t s) { k_lds128<1><<<g, b, s>>>(d, c, iters); }, w, SM);
        double c64 = run([&](int g, int b, size_t s) { k_lds64<<<g, b, s>>>(d, c, iters); }, w, SM);
        printf("LDS.128 %2d warps: 1 addr %.3f | 4 addr %.3f | 6 addr %.3f | 32 addr %.3f ; LDS.64 distinct %.3f  (warp-instr/clk/SM)\n",
               w, iters * 8.0 * w / c32, iters * 8.0 * w / c8, iters * 8.0 * w / c6, iters * 8.0 * w / c1,
               iters * 8.0 * w / c64);
    }
    for (int w : {4, 8, 16, 32}) {
        double a = run([&](int g, int b, size_t s) { k_ldsfma<8, 1><<<g, b, s>>>(d, c, iters); }, w, SM);
        double b6 = run([&](int g, int b, size_t s) { k_ldsfma<6, 1><<<g, b, s>>>(d, c, iters); }, w, SM);
        double b2 = run([&](int g, int b, size_t s) { k_ldsfma<8, 2><<<g, b, s>>>(d, c, iters); }, w, SM);
        double b1 = run([&](int g, int b, size_t s) { k_ldsfma<32, 1><<<g, b, s>>>(d, c, iters); }, w, SM);
        printf("LDS.128->DFMA  %2d warps: G8 R1 %.3f | G6 R1 %.3f | G8 R2 %.3f | G32 R1 %.3f  warp-DFMA/clk/SM\n", w,
               iters * 32.0 * w / a, iters * 32.0 * w / b6, iters * 64.0 * w / b2, iters * 32.0 * w / b1);
    }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float ms;
    k_dfma<8><<<148 * 4, 512>>>(d, c, iters);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k_dfma<8><<<148 * 4, 512>>>(d, c, iters);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    cudaEventElapsedTime(&ms, e0, e1);
    printf("full chip DFMA: %.2f TFLOP/s\n", 2.0 * 148.0 * 4 * 512 * (double)iters * 8 / (ms * 1e-3) / 1e12);
    k_dmma<6><<<148 * 4, 512>>>(d, c, iters);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k_dmma<6><<<148 * 4, 512>>>(d, c, iters);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    cudaEventElapsedTime(&ms, e0, e1);
    printf("full chip DMMA: %.2f TFLOP/s\n", 2.0 * 148.0 * 4 * 16 * (double)iters * 6 * 256 / (ms * 1e-3) / 1e12);
    return 0;
}
