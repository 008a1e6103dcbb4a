// Microbenchmark: dependent-issue latency and throughput of FP64 DFMA on the SM (tuning aid).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void lat(double* out, long long* cyc, int iters) {
    double a = out[0], b = out[1], x = out[2];
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) x = fma(a, x, b);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
    out[3 + threadIdx.x + blockIdx.x * blockDim.x] = x;
}
template <int ILP>
__global__ void thr(double* out, long long* cyc, int iters) {
    double a = out[0], b = out[1];
    double x[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) x[k] = out[2] + k;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int k = 0; k < ILP; ++k) x[k] = fma(a, x[k], b);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
    double s = 0;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s += x[k];
    out[3 + threadIdx.x + blockIdx.x * blockDim.x] = s;
}
// like thr<>, but every DFMA reads three distinct register pairs (no operand reuse): x[k] = fma(a[k], x[k], b[(k+j)%ILP])
template <int ILP>
__global__ void thr3(double* out, long long* cyc, int iters) {
    double a[ILP], b[ILP], x[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) { a[k] = out[0] + 1e-9 * k; b[k] = out[1] + 1e-9 * k; x[k] = out[2] + k; }
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int k = 0; k < ILP; ++k) x[k] = fma(a[(k + j) % ILP], x[k], b[(k + 2 * j + 1) % ILP]);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
    double s = 0;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s += x[k];
    out[3 + threadIdx.x + blockIdx.x * blockDim.x] = s;
}
int main() {
    double* d; long long* c; cudaMalloc(&d, 8 * (3 + 148 * 1024)); cudaMalloc(&c, 8);
    double h[3] = {0.999, 0.001, 1.0}; cudaMemcpy(d, h, 24, cudaMemcpyHostToDevice);
    long long hc; int iters = 2000;
    for (int warps : {1, 2, 4, 8, 16, 32}) {
        lat<<<1, 32 * warps>>>(d, c, iters); cudaDeviceSynchronize(); cudaMemcpy(&hc, c, 8, cudaMemcpyDeviceToHost);
        printf("dependent DFMA chain, %2d warps/SM: %.2f cycles per DFMA per warp\n", warps, (double)hc / (iters * 16));
    }
    for (int warps : {4, 8, 16, 32}) {
        thr<8><<<1, 32 * warps>>>(d, c, iters); cudaDeviceSynchronize(); cudaMemcpy(&hc, c, 8, cudaMemcpyDeviceToHost);
        printf("ILP 8, %2d warps/SM: %.3f warp-DFMA per cycle per SM\n", warps, (double)(iters * 4 * 8) * warps / hc);
    }
    for (int warps : {4, 8, 16}) {
        thr3<8><<<1, 32 * warps>>>(d, c, iters); cudaDeviceSynchronize(); cudaMemcpy(&hc, c, 8, cudaMemcpyDeviceToHost);
        printf("ILP 8, distinct operands, %2d warps/SM: %.3f warp-DFMA per cycle per SM\n", warps, (double)(iters * 4 * 8) * warps / hc);
    }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    thr<8><<<148 * 4, 512>>>(d, c, iters); cudaDeviceSynchronize();
    cudaEventRecord(e0); thr<8><<<148 * 4, 512>>>(d, c, iters); cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("full chip: %.2f TFLOP/s FP64 (DFMA)\n", 2.0 * 148.0 * 4 * 512 * (double)iters * 32 / (ms * 1e-3) / 1e12);
    return 0;
}
