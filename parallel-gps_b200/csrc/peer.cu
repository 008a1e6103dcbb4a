// Shard-summary exchange over NVLink peer memory (time sharding, one process per GPU).
//
// The messages of the time-sharded scans are a few hundred bytes per rank (one aggregate element):
// an NCCL all-gather costs a launch + protocol latency several times their transfer time.  Here every rank owns one
// slot array in a SYMMETRIC buffer (same layout on every GPU, peer-mapped: torch symmetric memory / cudaIpc) and ONE
// small kernel per exchange does the whole all-gather:
//   warp j : copies this rank's message into row `rank` of the slot in peer j's buffer (plain stores to the mapped
//            peer pointer — NVLink writes), fences system-wide, release-stores the sequence number into flag `rank`
//            of peer j, then spins (acquire loads) on flag j of its OWN buffer until rank j's message has landed.
// After the kernel the local slot holds all rows, exactly like the output of all_gather_into_tensor, and the scans'
// fold prologues read it in place.  The sequence number lives in device memory (incremented by the kernel itself), so
// the launch has no host-dependent argument and the step can be replayed from a CUDA graph.
#include "../../include/pssgp_b200.h"

#include <cuda_runtime.h>
#include <stdint.h>

#include "workspace.h"

namespace pssgp {

constexpr int kMaxPeers = 16;

struct PeerPtrs {
    double* buf[kMaxPeers];
    unsigned long long* flag[kMaxPeers];
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// grid 1, block world * 32.  slot_off / row_stride in doubles; flags: [nslots][world] at flag_off (in u64) of the
// flag region, seq_counter: this rank's per-slot sequence numbers (local memory).
__global__ void peer_exchange_kernel(const double* __restrict__ msg, int nvals, PeerPtrs pp, int world, int rank,
                                     long slot_off, long row_stride, long flag_off, unsigned long long* seq_counter,
                                     unsigned long long* err_counter) {
    __shared__ unsigned long long seq_s;
    const int j = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) seq_s = ++(*seq_counter);
    __syncthreads();
    const unsigned long long seq = seq_s;
    if (j < world) {
        double* dst = pp.buf[j] + slot_off + (long)rank * row_stride;
        for (int i = lane; i < nvals; i += 32) dst[i] = msg[i];
        __threadfence_system();
        __syncwarp();
        if (lane == 0) {
            st_release_sys(pp.flag[j] + flag_off + rank, seq);
            const unsigned long long* mine = pp.flag[rank] + flag_off + j;
            // a peer that never shows up (crashed rank, mismatched exchange sequence) must not hang the GPU: give up
            // after ~10 s of SM clocks and count the failure where the host can see it
            const long long t0 = clock64();
            while (ld_acquire_sys(mine) < seq) {
                __nanosleep(20);
                if (clock64() - t0 > 20000000000LL) {
                    atomicAdd(err_counter, 1ull);
                    break;
                }
            }
        }
        __syncwarp();
    }
    __threadfence_system();
}

}  // namespace pssgp

using namespace pssgp;

extern "C" int pssgp_peer_exchange(pssgp_handle* h, const void* msg, int64_t nvals, void* const* peer_bufs,
                                   void* const* peer_flags, int world, int rank, int64_t slot_offset,
                                   int64_t row_stride, int64_t flag_offset, void* seq_counter, void* err_counter,
                                   void* stream) {
    if (!h) return set_err(PSSGP_ERR_INVALID, "null handle");
    if (!msg || !peer_bufs || !peer_flags || !seq_counter || !err_counter) return set_err(PSSGP_ERR_INVALID, "peer_exchange: null argument");
    if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world)
        return set_err(PSSGP_ERR_INVALID, "peer_exchange: world %d / rank %d out of range (max %d peers)", world, rank, kMaxPeers);
    if (nvals < 1 || nvals > row_stride) return set_err(PSSGP_ERR_INVALID, "peer_exchange: message longer than a row");
    cudaSetDevice(h->device);
    PeerPtrs pp;
    for (int j = 0; j < world; ++j) {
        if (!peer_bufs[j] || !peer_flags[j]) return set_err(PSSGP_ERR_INVALID, "peer_exchange: null peer pointer");
        pp.buf[j] = (double*)peer_bufs[j];
        pp.flag[j] = (unsigned long long*)peer_flags[j];
    }
    cudaStream_t st = (cudaStream_t)stream;
    PSSGP_LAUNCH(h, "peer_exchange", st,
                 (peer_exchange_kernel<<<1, world * 32, 0, st>>>((const double*)msg, (int)nvals, pp, world, rank, (long)slot_offset,
                                                                 (long)row_stride, (long)flag_offset,
                                                                 (unsigned long long*)seq_counter,
                                                                 (unsigned long long*)err_counter)));
    return check_launch(h, "peer_exchange", 1);
}
