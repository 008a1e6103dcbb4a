// Memory-pipeline probe for the chunked scan: how fast can NT threads each walk their own contiguous
// chunk of 72-byte rows (two arrays) when the rows are staged through shared memory in different ways?
// No arithmetic beyond a checksum.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a membench.cu -o membench
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp16(unsigned dst, const void* src) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src) : "memory"); }
__device__ __forceinline__ void cp16ca(unsigned dst, const void* src) { asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src) : "memory"); }
__device__ __forceinline__ void commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void waitg() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// V0: plain coalesced read
__global__ void k_coalesced(const double2* __restrict__ a, long n16, double* out) {
    double s = 0;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n16; i += (long)gridDim.x * blockDim.x) {
        double2 v = __ldg(a + i);
        s += v.x + v.y;
    }
    if (s == 1.2345) out[0] = s;
}

// V1: cooperative LDGSTS, LSR rows per sub-step, NSTG stages, NARR arrays of W=9 doubles per row.
// lane -> (g, off): GR = 32 / NP segments per instruction.
template <int LSR, int NSTG, int NARR, bool CA>
__global__ void k_coop(const double* __restrict__ base, long rows_per_array, int L, double* out) {
    constexpr int W = 9;
    constexpr int SEG = LSR * W * 8;          // bytes per segment
    constexpr int NP = SEG / 16;
    constexpr int GR = 32 / NP > 0 ? 32 / NP : 1;
    constexpr int IT = (32 + GR - 1) / GR;
    constexpr int PITCH = (NP | 1) * 16;
    constexpr int STAGE = NARR * 32 * PITCH;
    extern __shared__ __align__(16) unsigned char sm[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    unsigned char* wsm = sm + wid * (NSTG * STAGE);
    const unsigned wsa = smem_u32(wsm);
    const long gw = (long)blockIdx.x * nw + wid;
    const long k_lo0 = gw * 32 * (long)L;
    const int nsub = L / LSR;
    const int g = lane / NP, off = lane - g * NP;
    const bool act = g < GR;
    const long step = (long)L * W * 8 * GR;
    const unsigned char* ptr[NARR];
#pragma unroll
    for (int a = 0; a < NARR; ++a)
        ptr[a] = (const unsigned char*)(base + (long)a * rows_per_array * W) + (k_lo0 + (long)g * L) * (W * 8) + off * 16;
    const unsigned soff = g * PITCH + off * 16;
    auto issue = [&](int stage) {
#pragma unroll
        for (int a = 0; a < NARR; ++a) {
            if (act) {
#pragma unroll
                for (int i = 0; i < IT; ++i)
                    if (g + GR * i < 32) {
                        if (CA) cp16ca(wsa + stage * STAGE + a * 32 * PITCH + soff + i * GR * PITCH, ptr[a] + i * step);
                        else cp16(wsa + stage * STAGE + a * 32 * PITCH + soff + i * GR * PITCH, ptr[a] + i * step);
                    }
            }
            ptr[a] += SEG;
        }
    };
#pragma unroll 1
    for (int s = 0; s < NSTG; ++s) { if (s < nsub) issue(s); commit(); }
    double acc = 0;
#pragma unroll 1
    for (int s = 0; s < nsub; ++s) {
        const int st = s % NSTG;
        waitg<NSTG - 1>();
        __syncwarp();
#pragma unroll
        for (int a = 0; a < NARR; ++a) {
            const unsigned char* src = wsm + st * STAGE + a * 32 * PITCH + lane * PITCH;
#pragma unroll
            for (int u = 0; u < NP; ++u) { double2 v = *(const double2*)(src + u * 16); acc += v.x + v.y; }
        }
        __syncwarp();
        if (s + NSTG < nsub) issue(st);
        commit();
    }
    waitg<0>();
    if (acc == 1.2345) out[0] = acc;
}

// V5: per-lane bulk async copy (cp.async.bulk) of its own segment, mbarrier per warp-stage
template <int LSR, int NSTG, int NARR>
__global__ void k_bulk(const double* __restrict__ base, long rows_per_array, int L, double* out) {
    constexpr int W = 9;
    constexpr int SEG = LSR * W * 8;
    constexpr int NP = SEG / 16;
    constexpr int PITCH = (NP | 1) * 16;
    constexpr int STAGE = NARR * 32 * PITCH;
    extern __shared__ __align__(16) unsigned char sm[];
    __shared__ __align__(8) unsigned long long bars[32 * 8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    unsigned char* wsm = sm + wid * (NSTG * STAGE);
    const unsigned wsa = smem_u32(wsm);
    const long gw = (long)blockIdx.x * nw + wid;
    const long k_lo = (gw * 32 + lane) * (long)L;
    const int nsub = L / LSR;
    unsigned bar[NSTG];
#pragma unroll
    for (int s = 0; s < NSTG; ++s) bar[s] = smem_u32(&bars[wid * 8 + s]);
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < NSTG; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 32;\n" ::"r"(bar[s]));
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    __syncwarp();
    const unsigned char* ptr[NARR];
#pragma unroll
    for (int a = 0; a < NARR; ++a) ptr[a] = (const unsigned char*)(base + (long)a * rows_per_array * W) + k_lo * (W * 8);
    auto issue = [&](int stage) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar[stage]), "r"(NARR * SEG) : "memory");
#pragma unroll
        for (int a = 0; a < NARR; ++a) {
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                             wsa + stage * STAGE + a * 32 * PITCH + lane * PITCH),
                         "l"(ptr[a]), "r"(SEG), "r"(bar[stage])
                         : "memory");
            ptr[a] += SEG;
        }
    };
#pragma unroll 1
    for (int s = 0; s < NSTG; ++s) if (s < nsub) issue(s);
    double acc = 0;
    unsigned phase = 0;
#pragma unroll 1
    for (int s = 0; s < nsub; ++s) {
        const int st = s % NSTG;
        unsigned done = 0;
        while (!done) {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(bar[st]), "r"((phase >> st) & 1u) : "memory");
        }
        phase ^= (1u << st);
#pragma unroll
        for (int a = 0; a < NARR; ++a) {
            const unsigned char* src = wsm + st * STAGE + a * 32 * PITCH + lane * PITCH;
#pragma unroll
            for (int u = 0; u < NP; ++u) { double2 v = *(const double2*)(src + u * 16); acc += v.x + v.y; }
        }
        __syncwarp();
        if (s + NSTG < nsub) issue(st);
    }
    if (acc == 1.2345) out[0] = acc;
}

// V7: contiguous warp tiles: every warp streams a contiguous region tile by tile (coalesced cp.async), each lane
// reads LSR rows of the tile from shared memory (the "tile" structure: rows of a lane are only LSR long).
template <int LSR, int NSTG, int NARR>
__global__ void k_tile(const double* __restrict__ base, long rows_per_array, long rows_per_warp, double* out) {
    constexpr int W = 9;
    constexpr int TILE = 32 * LSR * W * 8;  // bytes per array per tile
    constexpr int STAGE = NARR * TILE;
    extern __shared__ __align__(16) unsigned char sm[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    unsigned char* wsm = sm + wid * (NSTG * STAGE);
    const unsigned wsa = smem_u32(wsm);
    const long gw = (long)blockIdx.x * nw + wid;
    const int ntile = (int)(rows_per_warp / (32 * LSR));
    const unsigned char* ptr[NARR];
#pragma unroll
    for (int a = 0; a < NARR; ++a) ptr[a] = (const unsigned char*)(base + (long)a * rows_per_array * W) + gw * rows_per_warp * (W * 8) + lane * 16;
    auto issue = [&](int stage) {
#pragma unroll
        for (int a = 0; a < NARR; ++a) {
#pragma unroll
            for (int i = 0; i < TILE / 512; ++i) cp16(wsa + stage * STAGE + a * TILE + lane * 16 + i * 512, ptr[a] + i * 512);
            ptr[a] += TILE;
        }
    };
#pragma unroll 1
    for (int s = 0; s < NSTG; ++s) { if (s < ntile) issue(s); commit(); }
    double acc = 0;
#pragma unroll 1
    for (int s = 0; s < ntile; ++s) {
        const int st = s % NSTG;
        waitg<NSTG - 1>();
        __syncwarp();
#pragma unroll
        for (int a = 0; a < NARR; ++a) {
            const unsigned char* src = wsm + st * STAGE + a * TILE + lane * (LSR * W * 8);
#pragma unroll
            for (int u = 0; u < LSR * W / 2; ++u) { double2 v = *(const double2*)(src + u * 16); acc += v.x + v.y; }
        }
        __syncwarp();
        if (s + NSTG < ntile) issue(st);
        commit();
    }
    waitg<0>();
    if (acc == 1.2345) out[0] = acc;
}

template <typename F> float timeit(F f, int reps = 10) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) f();
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) f();
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms / reps * 1000.f;
}

int main() {
    const int SMS = 148;
    const int NARR = 2;
    constexpr int W = 9;
    // rows: multiple of everything we need
    const long rows = 148L * 8 * 32 * 28;  // 1,060,864 rows
    const size_t bytes = (size_t)rows * W * 8 * NARR;
    double *buf, *out;
    CK(cudaMalloc(&buf, bytes + (1 << 20)));
    CK(cudaMalloc(&out, 64));
    CK(cudaMemset(buf, 0, bytes + (1 << 20)));
    auto report = [&](const char* name, float us) { printf("%-44s %8.1f us  %7.1f GB/s\n", name, us, bytes / us * 1e-3); };

    report("V0 coalesced LDG.128", timeit([&] { k_coalesced<<<SMS * 8, 256>>>((const double2*)buf, bytes / 16, out); }));

#define RUN_COOP(LSR, NSTG, CA, NWARPS)                                                                      \
    {                                                                                                        \
        constexpr int SEG = LSR * W * 8, NP = SEG / 16, PITCH = (NP | 1) * 16, STAGE = NARR * 32 * PITCH;    \
        int smem = NWARPS * NSTG * STAGE;                                                                    \
        int L = (int)((rows + (long)SMS * NWARPS * 32 - 1) / ((long)SMS * NWARPS * 32));                     \
        L = (L + LSR - 1) / LSR * LSR;                                                                       \
        int ctas = (int)(rows / ((long)NWARPS * 32 * L));                                                    \
        if (smem <= 227 * 1024) {                                                                            \
            CK(cudaFuncSetAttribute(k_coop<LSR, NSTG, NARR, CA>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
            char nm[128];                                                                                    \
            snprintf(nm, 128, "coop LS=%d NST=%d %s warps=%d L=%d ctas=%d", LSR, NSTG, CA ? "ca" : "cg", NWARPS, L, ctas); \
            float us = timeit([&] { k_coop<LSR, NSTG, NARR, CA><<<ctas, NWARPS * 32, smem>>>(buf, rows, L, out); }); \
            printf("%-44s %8.1f us  %7.1f GB/s\n", nm, us, (double)ctas * NWARPS * 32 * L * W * 8 * NARR / us * 1e-3);   \
        }                                                                                                    \
    }
    RUN_COOP(2, 2, false, 8)
    RUN_COOP(2, 2, true, 8)
    RUN_COOP(2, 3, false, 8)
    RUN_COOP(2, 2, false, 11)
    RUN_COOP(2, 2, false, 6)
    RUN_COOP(2, 2, false, 4)
    RUN_COOP(2, 4, false, 4)
    RUN_COOP(4, 2, false, 5)
    RUN_COOP(4, 2, false, 4)
    RUN_COOP(4, 3, false, 3)
    RUN_COOP(8, 2, false, 3)
    RUN_COOP(8, 2, false, 2)

#define RUN_BULK(LSR, NSTG, NWARPS)                                                                          \
    {                                                                                                        \
        constexpr int SEG = LSR * W * 8, NP = SEG / 16, PITCH = (NP | 1) * 16, STAGE = NARR * 32 * PITCH;    \
        int smem = NWARPS * NSTG * STAGE;                                                                    \
        int L = (int)((rows + (long)SMS * NWARPS * 32 - 1) / ((long)SMS * NWARPS * 32));                     \
        L = (L + LSR - 1) / LSR * LSR;                                                                       \
        int ctas = (int)(rows / ((long)NWARPS * 32 * L));                                                    \
        if (smem <= 220 * 1024) {                                                                            \
            CK(cudaFuncSetAttribute(k_bulk<LSR, NSTG, NARR>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
            char nm[128];                                                                                    \
            snprintf(nm, 128, "bulk LS=%d NST=%d warps=%d L=%d ctas=%d", LSR, NSTG, NWARPS, L, ctas);          \
            report(nm, timeit([&] { k_bulk<LSR, NSTG, NARR><<<ctas, NWARPS * 32, smem>>>(buf, rows, L, out); })); \
        }                                                                                                    \
    }
    RUN_BULK(2, 2, 8)
    RUN_BULK(2, 4, 8)
    RUN_BULK(4, 2, 5)
    RUN_BULK(8, 2, 3)
    RUN_BULK(8, 3, 2)

#define RUN_TILE(LSR, NSTG, NWARPS)                                                                          \
    {                                                                                                        \
        constexpr int STAGE = NARR * 32 * LSR * W * 8;                                                       \
        int smem = NWARPS * NSTG * STAGE;                                                                    \
        long rpw = rows / ((long)SMS * NWARPS);                                                              \
        rpw = rpw / (32 * LSR) * (32 * LSR);                                                                 \
        if (smem <= 227 * 1024) {                                                                            \
            CK(cudaFuncSetAttribute(k_tile<LSR, NSTG, NARR>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
            char nm[128];                                                                                    \
            snprintf(nm, 128, "tile LS=%d NST=%d warps=%d rows/warp=%ld", LSR, NSTG, NWARPS, rpw);              \
            report(nm, timeit([&] { k_tile<LSR, NSTG, NARR><<<SMS, NWARPS * 32, smem>>>(buf, rows, rpw, out); })); \
        }                                                                                                    \
    }
    RUN_TILE(2, 2, 8)
    RUN_TILE(2, 4, 8)
    RUN_TILE(4, 3, 8)
    RUN_TILE(8, 2, 8)
    CK(cudaDeviceSynchronize());
    return 0;
}
