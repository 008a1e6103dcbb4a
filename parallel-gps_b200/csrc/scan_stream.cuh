// Streaming chunked scan for register-resident state dimensions (D <= 4).
//
// The time axis is cut into chunks of L consecutive steps, one thread per chunk, one warp per 32
// consecutive chunks, NW warps per CTA, and (by choice of L) one CTA per SM: the whole grid is a
// single resident wave.  Three kernels per scan (filter / smoother / adjoint):
//
//   K1 stream_reduce : every thread folds its L steps into the chunk aggregate with the algebra's cheap
//                      sequential `append_row`; warp Kogge-Stone scan + CTA fold with the generic
//                      associative operator.  Stores the CTA-exclusive aggregate of every chunk and the
//                      total of every CTA.
//   K2 scan_mid      : one CTA scans the CTA totals and turns them into the *state* entering each CTA.
//   K3 stream_apply  : every thread applies state o exclusive-aggregate and re-runs the seeded recursion
//                      (`step_row`) over its L steps, producing the outputs.
//
// Memory path (K1 and K3).  A thread walks its chunk sequentially (72-byte records at D = 3), which
// read with plain loads is a dependent, latency-bound pattern at the 6-8 warps per SM that the
// register-resident FP64 state allows.  Instead every thread owns a private shared-memory FIFO of NST
// stages and prefetches its chunk through it with cp.async (LDGSTS, no register staging): per sub-step
// the LS-row segment (16 * W bytes, 16-byte aligned because LS * sizeof(T) = 16) of every input array,
// as 16-byte pieces with immediate offsets.  A slot's pitch is an odd number of 16-byte units, so the
// 128-bit shared loads of a warp are bank-conflict free.  A stage is handed back to the copy engine as
// soon as its last row has been fetched into registers, so the next sub-steps are in flight while the
// FP64 pipe works.  Threads only ever read what they copied themselves: the streaming loop has no
// barrier of any kind (cp.async.wait_group is the only synchronisation).  Outputs go straight from
// registers to global memory with 16-byte streaming stores.
//
// Reverse scans (smoother, adjoint) use the same time partition and walk chunks and rows in
// descending time; nothing is physically reversed.  Arrays the algebra needs at row k+1 / k-1
// (smoother: F, Q of the next step; adjoint: filtered moments of the previous step) are streamed with
// a one-row shift (element-sized pieces, since a one-row shift breaks 16-byte alignment).
//
// An "Algebra" supplies: NAGG, NSTATE, NACC, REVERSE, Params, Ctx, the array tables (NIN, in_w,
// in_shift, in_ptr, NOUT, out_w, out_ptr), identity, combine, apply, load_init, expand_state, finish,
// append_row and step_row.
#pragma once
#include "smalld.cuh"

namespace pssgp {

// ---------------------------------------------------------------------------------------------
// cp.async + 128-bit shared/global helpers
// ---------------------------------------------------------------------------------------------
PSSGP_DEV unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int BYTES> PSSGP_DEV void cp_async(unsigned dst, const void* src) {
    static_assert(BYTES == 4 || BYTES == 8 || BYTES == 16, "cp.async size");
    if constexpr (BYTES == 16)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src) : "memory");
    else if constexpr (BYTES == 8)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(dst), "l"(src) : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(dst), "l"(src) : "memory");
}
PSSGP_DEV void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> PSSGP_DEV void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// 16 bytes shared -> registers / registers -> shared / registers -> global (streaming store)
PSSGP_DEV void ld_shared16(double* dst, const unsigned char* src) {
    const double2 v = *reinterpret_cast<const double2*>(src);
    dst[0] = v.x;
    dst[1] = v.y;
}
PSSGP_DEV void ld_shared16(float* dst, const unsigned char* src) {
    const float4 v = *reinterpret_cast<const float4*>(src);
    dst[0] = v.x;
    dst[1] = v.y;
    dst[2] = v.z;
    dst[3] = v.w;
}
PSSGP_DEV void st_shared16(unsigned char* dst, const double* src) {
    *reinterpret_cast<double2*>(dst) = make_double2(src[0], src[1]);
}
PSSGP_DEV void st_shared16(unsigned char* dst, const float* src) {
    *reinterpret_cast<float4*>(dst) = make_float4(src[0], src[1], src[2], src[3]);
}

// ---------------------------------------------------------------------------------------------
// Compile-time geometry of the per-warp staging areas
// ---------------------------------------------------------------------------------------------
template <typename T> struct StreamGeom {
    static constexpr int LS = 16 / (int)sizeof(T);  // rows per sub-step: LS * W * sizeof(T) is a multiple of 16
    __host__ __device__ static constexpr int seg_bytes(int w) { return LS * w * (int)sizeof(T); }
    __host__ __device__ static constexpr int seg_units(int w) { return seg_bytes(w) / 16; }
    __host__ __device__ static constexpr int pitch(int w) { return (seg_units(w) | 1) * 16; }  // odd number of 16-byte units
};

template <typename Alg> struct StreamLayout {
    using T = typename Alg::scalar;
    using G = StreamGeom<T>;
    static constexpr int LS = G::LS;
    __host__ __device__ static constexpr int in_off(int a) { return a == 0 ? 0 : in_off(a - 1) + 32 * G::pitch(Alg::in_w(a - 1)); }
    static constexpr int STAGE_BYTES = in_off(Alg::NIN);
    static constexpr int NST = 2;
    static constexpr int WARP_BYTES = NST * STAGE_BYTES;
    static constexpr int SMEM_BUDGET = 216 * 1024;
    __host__ __device__ static constexpr int nw_fit() { return SMEM_BUDGET / WARP_BYTES; }
    static constexpr int NW = nw_fit() > 8 ? 8 : (nw_fit() < 1 ? 1 : nw_fit());
};

// time row at which lane-chunk `c` starts sub-step `s`
template <bool REVERSE> PSSGP_DEV long seg_row(long c, int s, int nsub, int L, int LS) {
    return c * (long)L + (long)(REVERSE ? (nsub - 1 - s) : s) * LS;
}

// Issues this lane's cp.async copies of input array A for one sub-step: the LS-row segment that starts
// at time row `row0` of the lane's own chunk, into the lane's slot of the stage at `stage_addr`.
template <typename Alg, int A>
PSSGP_DEV void stream_issue_array(const typename Alg::Params& p, long n, long row0, int lane, unsigned stage_addr) {
    using T = typename Alg::scalar;
    using G = StreamGeom<T>;
    using Lay = StreamLayout<Alg>;
    constexpr int W = Alg::in_w(A);
    constexpr int SH = Alg::in_shift(A);
    constexpr int SZ = (int)sizeof(T);
    constexpr int PB = (SH == 0) ? 16 : SZ;  // a one-row shift breaks 16-byte alignment: element-sized pieces
    constexpr int SEG = G::seg_bytes(W);
    constexpr int NP = SEG / PB;  // pieces per segment
    const unsigned char* base = reinterpret_cast<const unsigned char*>(Alg::in_ptr(p, A));
    const long total = n * (long)(W * SZ);
    const long gb = (row0 + SH) * (long)(W * SZ);
    const unsigned dst = stage_addr + Lay::in_off(A) + lane * G::pitch(W);
    const unsigned char* src = base + gb;
    if (gb >= 0 && gb + SEG <= total) {
#pragma unroll
        for (int i = 0; i < NP; ++i) cp_async<PB>(dst + i * PB, src + i * PB);
    } else {
        // segment crosses an end of the array: copy the elements that exist one by one
#pragma unroll
        for (int e = 0; e < SEG / SZ; ++e) {
            const long ge = gb + (long)e * SZ;
            if (ge >= 0 && ge + SZ <= total) cp_async<SZ>(dst + e * SZ, src + e * SZ);
        }
    }
}

template <typename Alg, int A = 0>
PSSGP_DEV void stream_issue(const typename Alg::Params& p, long n, long row0, int lane, unsigned stage_addr) {
    if constexpr (A < Alg::NIN) {
        stream_issue_array<Alg, A>(p, n, row0, lane, stage_addr);
        stream_issue<Alg, A + 1>(p, n, row0, lane, stage_addr);
    }
}

// Copies row r of this lane's slot of every input array from a stage into registers (128-bit shared
// loads; a 16-byte unit shared by two rows is simply read by both).
template <typename Alg, int A = 0>
PSSGP_DEV void stream_fetch_row(const unsigned char* stage, int lane, int r,
                                typename Alg::scalar (&row)[Alg::NIN][Alg::WMAX]) {
    using T = typename Alg::scalar;
    using G = StreamGeom<T>;
    using Lay = StreamLayout<Alg>;
    if constexpr (A < Alg::NIN) {
        constexpr int W = Alg::in_w(A);
        constexpr int EPU = 16 / (int)sizeof(T);  // elements per 16-byte unit
        const unsigned char* src = stage + Lay::in_off(A) + lane * G::pitch(W);
        const int e_lo = r * W, e_hi = (r + 1) * W;
#pragma unroll
        for (int u = 0; u < G::seg_units(W); ++u) {
            if ((u + 1) * EPU > e_lo && u * EPU < e_hi) {
                T tmp[EPU];
                ld_shared16(tmp, src + u * 16);
#pragma unroll
                for (int j = 0; j < EPU; ++j) {
                    const int e = u * EPU + j;
                    if (e >= e_lo && e < e_hi) row[A][e - e_lo] = tmp[j];
                }
            }
        }
        stream_fetch_row<Alg, A + 1>(stage, lane, r, row);
    }
}

// Registers -> global for row r of the LS-row segment that starts at time row `row0`: 16-byte streaming
// stores for the units that lie inside the row, element stores for a unit shared with the next row.
PSSGP_DEV void st_global16(unsigned char* dst, const double* v) {
    __stcs(reinterpret_cast<double2*>(dst), make_double2(v[0], v[1]));
}
PSSGP_DEV void st_global16(unsigned char* dst, const float* v) {
    __stcs(reinterpret_cast<float4*>(dst), make_float4(v[0], v[1], v[2], v[3]));
}
template <typename Alg, int A = 0>
PSSGP_DEV void stream_store_row(const typename Alg::Params& p, long row0, int r,
                                const typename Alg::scalar (&orow)[Alg::NOUT][Alg::WMAX]) {
    using T = typename Alg::scalar;
    using G = StreamGeom<T>;
    if constexpr (A < Alg::NOUT) {
        constexpr int W = Alg::out_w(A);
        constexpr int SZ = (int)sizeof(T);
        constexpr int EPU = 16 / SZ;
        unsigned char* dst = reinterpret_cast<unsigned char*>(Alg::out_ptr(p, A)) + row0 * (long)(W * SZ);
        const int e_lo = r * W, e_hi = (r + 1) * W;
#pragma unroll
        for (int u = 0; u < G::seg_units(W); ++u) {
            if (u * EPU >= e_lo && (u + 1) * EPU <= e_hi) {
                T tmp[EPU];
#pragma unroll
                for (int j = 0; j < EPU; ++j) tmp[j] = orow[A][u * EPU + j - e_lo];
                st_global16(dst + u * 16, tmp);
            } else if ((u + 1) * EPU > e_lo && u * EPU < e_hi) {
#pragma unroll
                for (int j = 0; j < EPU; ++j) {
                    const int e = u * EPU + j;
                    if (e >= e_lo && e < e_hi) __stcs(reinterpret_cast<T*>(dst + e * SZ), orow[A][e - e_lo]);
                }
            }
        }
        stream_store_row<Alg, A + 1>(p, row0, r, orow);
    }
}

// ---------------------------------------------------------------------------------------------
// K1: chunk aggregates, CTA-exclusive prefixes, CTA totals
// ---------------------------------------------------------------------------------------------
template <typename Alg>
__global__ void __launch_bounds__(StreamLayout<Alg>::NW * 32)
stream_reduce_kernel(typename Alg::Params p, long n, int L, long nChunks, long nChunksPad,
                     typename Alg::scalar* __restrict__ lane_excl, typename Alg::scalar* __restrict__ wagg,
                     long nCta) {
    using T = typename Alg::scalar;
    using Lay = StreamLayout<Alg>;
    constexpr int NW = Lay::NW, NST = Lay::NST, LS = Lay::LS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned char* wsm = smem_raw + wid * Lay::WARP_BYTES;
    const unsigned wsm_addr = smem_addr(wsm);
    const long lc = ((long)blockIdx.x * NW + wid) * 32 + lane;  // logical chunk (scan order)
    const long c = Alg::REVERSE ? (nChunks - 1 - lc) : lc;      // time chunk
    const int nsub = L / LS;

    T a[Alg::NAGG];
    Alg::identity(a);
    if (lc < nChunks) {
        typename Alg::Ctx ctx;
        Alg::load_ctx(p, ctx);
#pragma unroll
        for (int s = 0; s < NST; ++s) {
            if (s < nsub)
                stream_issue<Alg>(p, n, seg_row<Alg::REVERSE>(c, s, nsub, L, LS), lane, wsm_addr + s * Lay::STAGE_BYTES);
            cp_async_commit();
        }
#pragma unroll 1
        for (int s = 0; s < nsub; ++s) {
            const int st = s % NST;
            cp_async_wait<NST - 1>();
            const long k0 = seg_row<Alg::REVERSE>(c, s, nsub, L, LS);
#pragma unroll
            for (int rr = 0; rr < LS; ++rr) {
                const int r = Alg::REVERSE ? (LS - 1 - rr) : rr;
                const long k = k0 + r;
                T row[Alg::NIN][Alg::WMAX];
                stream_fetch_row<Alg>(wsm + st * Lay::STAGE_BYTES, lane, r, row);
                if (rr == LS - 1) {
                    // the stage is drained: hand it back to the copy engine before the last row's arithmetic
                    if (s + NST < nsub)
                        stream_issue<Alg>(p, n, seg_row<Alg::REVERSE>(c, s + NST, nsub, L, LS), lane,
                                          wsm_addr + st * Lay::STAGE_BYTES);
                    cp_async_commit();
                }
                if (k < n) Alg::append_row(a, ctx, row, 0, k, p);
            }
        }
        cp_async_wait<0>();
    }
    // warp inclusive scan (earlier lane is the left operand)
#pragma unroll 1
    for (int off = 1; off < 32; off <<= 1) {
        T o[Alg::NAGG];
#pragma unroll
        for (int e = 0; e < Alg::NAGG; ++e) o[e] = shfl_up_t(a[e], off);
        if (lane >= off) {
            T r[Alg::NAGG];
            Alg::combine(o, a, r);
#pragma unroll
            for (int e = 0; e < Alg::NAGG; ++e) a[e] = r[e];
        }
    }
    // CTA level: every warp folds the totals of the warps before it
    __shared__ T shw[NW * Alg::NAGG];
    if (lane == 31) {
#pragma unroll
        for (int e = 0; e < Alg::NAGG; ++e) shw[wid * Alg::NAGG + e] = a[e];
    }
    T ex[Alg::NAGG];
#pragma unroll
    for (int e = 0; e < Alg::NAGG; ++e) ex[e] = shfl_up_t(a[e], 1);
    if (lane == 0) Alg::identity(ex);
    __syncthreads();
    T wp[Alg::NAGG];  // aggregate of the warps before this one (wid > 0)
    if (wid > 0) {
#pragma unroll
        for (int e = 0; e < Alg::NAGG; ++e) wp[e] = shw[e];
#pragma unroll 1
        for (int w = 1; w < wid; ++w) {
            T b[Alg::NAGG], r[Alg::NAGG];
#pragma unroll
            for (int e = 0; e < Alg::NAGG; ++e) b[e] = shw[w * Alg::NAGG + e];
            Alg::combine(wp, b, r);
#pragma unroll
            for (int e = 0; e < Alg::NAGG; ++e) wp[e] = r[e];
        }
        T r[Alg::NAGG];
        Alg::combine(wp, ex, r);
#pragma unroll
        for (int e = 0; e < Alg::NAGG; ++e) ex[e] = r[e];
    }
    if (threadIdx.x != 0 && lc < nChunks) {
#pragma unroll
        for (int e = 0; e < Alg::NAGG; ++e) lane_excl[(long)e * nChunksPad + lc] = ex[e];
    }
    if (threadIdx.x == NW * 32 - 1) {
        T tot[Alg::NAGG];
        if (wid > 0) {
            Alg::combine(wp, a, tot);
        } else {
#pragma unroll
            for (int e = 0; e < Alg::NAGG; ++e) tot[e] = a[e];
        }
#pragma unroll
        for (int e = 0; e < Alg::NAGG; ++e) wagg[(long)e * nCta + blockIdx.x] = tot[e];
    }
}

// ---------------------------------------------------------------------------------------------
// K3: seeded recursion + outputs + deterministic reduction of the accumulators
// ---------------------------------------------------------------------------------------------
template <typename Alg>
__global__ void __launch_bounds__(StreamLayout<Alg>::NW * 32)
stream_apply_kernel(typename Alg::Params p, long n, int L, long nChunks, long nChunksPad,
                    const typename Alg::scalar* __restrict__ lane_excl,
                    const typename Alg::scalar* __restrict__ wstate, long nCta,
                    typename Alg::scalar* __restrict__ acc_part, unsigned int* __restrict__ ticket,
                    typename Alg::scalar* __restrict__ acc_out) {
    using T = typename Alg::scalar;
    using Lay = StreamLayout<Alg>;
    constexpr int NW = Lay::NW, NST = Lay::NST, LS = Lay::LS;
    constexpr int NACC1 = Alg::NACC > 0 ? Alg::NACC : 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned char* wsm = smem_raw + wid * Lay::WARP_BYTES;
    const unsigned wsm_addr = smem_addr(wsm);
    const long lc = ((long)blockIdx.x * NW + wid) * 32 + lane;
    const long c = Alg::REVERSE ? (nChunks - 1 - lc) : lc;
    const int nsub = L / LS;

    T acc[NACC1];
#pragma unroll
    for (int e = 0; e < NACC1; ++e) acc[e] = T(0);
    if (lc < nChunks) {
        typename Alg::Ctx ctx;
        Alg::load_ctx(p, ctx);
#pragma unroll
        for (int s = 0; s < NST; ++s) {
            if (s < nsub)
                stream_issue<Alg>(p, n, seg_row<Alg::REVERSE>(c, s, nsub, L, LS), lane, wsm_addr + s * Lay::STAGE_BYTES);
            cp_async_commit();
        }
        T st8[Alg::NSTATE];
#pragma unroll
        for (int e = 0; e < Alg::NSTATE; ++e) st8[e] = wstate[(long)e * nCta + blockIdx.x];
        if (threadIdx.x != 0) {
            T ex[Alg::NAGG], s2[Alg::NSTATE];
#pragma unroll
            for (int e = 0; e < Alg::NAGG; ++e) ex[e] = lane_excl[(long)e * nChunksPad + lc];
            Alg::apply(st8, ex, s2);
#pragma unroll
            for (int e = 0; e < Alg::NSTATE; ++e) st8[e] = s2[e];
        }
#pragma unroll 1
        for (int s = 0; s < nsub; ++s) {
            const int st = s % NST;
            cp_async_wait<NST - 1>();
            const long k0 = seg_row<Alg::REVERSE>(c, s, nsub, L, LS);
#pragma unroll
            for (int rr = 0; rr < LS; ++rr) {
                const int r = Alg::REVERSE ? (LS - 1 - rr) : rr;
                const long k = k0 + r;
                T row[Alg::NIN][Alg::WMAX];
                stream_fetch_row<Alg>(wsm + st * Lay::STAGE_BYTES, lane, r, row);
                if (rr == LS - 1) {
                    // the stage is drained: hand it back to the copy engine before the last row's arithmetic
                    if (s + NST < nsub)
                        stream_issue<Alg>(p, n, seg_row<Alg::REVERSE>(c, s + NST, nsub, L, LS), lane,
                                          wsm_addr + st * Lay::STAGE_BYTES);
                    cp_async_commit();
                }
                if (k < n) {
                    T orow[Alg::NOUT][Alg::WMAX];
                    Alg::step_row(st8, ctx, row, orow, 0, k, p, acc);
                    stream_store_row<Alg>(p, k0, r, orow);
                }
            }
        }
        cp_async_wait<0>();
    }
    if (Alg::NACC > 0) {
        // deterministic grid reduction: warp shuffle -> smem -> per-CTA partial -> last CTA sums in a fixed order
        __shared__ T red[NW * NACC1];
        __shared__ bool is_last;
#pragma unroll
        for (int e = 0; e < Alg::NACC; ++e) {
            T v = acc[e];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) v += shfl_down_t(v, off);
            if (lane == 0) red[wid * Alg::NACC + e] = v;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int e = 0; e < Alg::NACC; ++e) {
                T v = T(0);
                for (int w = 0; w < NW; ++w) v += red[w * Alg::NACC + e];
                acc_part[(long)blockIdx.x * Alg::NACC + e] = v;
            }
            __threadfence();
            unsigned int t = atomicAdd(ticket, 1u);
            is_last = (t == gridDim.x - 1);
        }
        __syncthreads();
        if (is_last) {
            __threadfence();
            const int nb = gridDim.x;
            for (int e = 0; e < Alg::NACC; ++e) {
                T v = T(0);
                for (int b = threadIdx.x; b < nb; b += blockDim.x)
                    v += ((volatile T*)acc_part)[(long)b * Alg::NACC + e];
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) v += shfl_down_t(v, off);
                __syncthreads();
                if (lane == 0) red[wid] = v;
                __syncthreads();
                if (threadIdx.x == 0) {
                    T tot = T(0);
                    for (int w = 0; w < NW; ++w) tot += red[w];
                    Alg::finish(p, e, tot, acc_out);
                }
            }
            if (threadIdx.x == 0) *ticket = 0u;
        }
    }
}

}  // namespace pssgp
