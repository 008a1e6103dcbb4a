// Streaming chunked scan for register-resident state dimensions (D <= 4).
//
// The time axis is cut into chunks of L consecutive steps, one thread per chunk, one warp per 32
// consecutive chunks, NW warps per CTA, and (by choice of L) one CTA per SM: the whole grid is a
// single resident wave.  Three kernels per scan (filter / smoother / adjoint):
//
//   K1 stream_reduce : every thread folds its L steps into the chunk aggregate with the algebra's cheap
//                      sequential `append_row`; warp Kogge-Stone scan with the generic associative
//                      operator, then one warp scans the NW warp totals.  Stores the warp-exclusive
//                      aggregate of every chunk, the CTA-exclusive aggregate of every warp and the total
//                      of every CTA.
//   K2 scan_mid      : one CTA scans the CTA totals and turns them into the *state* entering each CTA.
//   K3 stream_apply  : every thread applies state o warp-prefix o lane-prefix and re-runs the seeded
//                      recursion (`step_row`) over its L steps, producing the outputs.
//
// Memory path (K1 and K3).  A thread walks its chunk sequentially, so its rows are 72-byte (D = 3)
// records strided by L rows between neighbouring lanes.  Every warp streams its 32 chunks through its
// own shared-memory stages with cp.async (LDGSTS, no register staging): per sub-step, for each input
// array, the LS-row segment (16 * W bytes, 16-byte aligned because LS * sizeof(T) = 16) of each of its
// 32 chunks.  The copy is cooperative: a segment is NP 16-byte pieces, so 32 / NP whole segments fit one
// warp instruction (lane -> (group g, piece off), fixed for the whole kernel; per instruction only a
// pointer bump), and every global request is made of full contiguous segments.  The destination is a
// per-lane slot whose pitch is an odd number of 16-byte units (bank-conflict-free 128-bit shared
// loads).  A stage is handed back to the copy engine as soon as its last row has been fetched into
// registers, so the copies of the next sub-steps are in flight while the FP64 pipe works; there is no
// CTA-wide barrier in the streaming loop (warps are autonomous: cp.async.wait_group + __syncwarp).
// Outputs take the mirror path: registers -> per-lane staging slot -> cooperative 16-byte streaming
// stores.
//
// Reverse scans (smoother, adjoint) use the same time partition and walk chunks and rows in descending
// time; nothing is physically reversed, and nothing is loaded with a row shift: what an algebra needs
// from the neighbouring row (smoother: F, Q of step k+1; adjoint: the step itself is delayed by one row,
// OUT_SHIFT = 1) is carried in registers from the previous iteration, the chunk-boundary row comes from a
// direct (halo) load.
//
// An "Algebra" supplies: NAGG, NSTATE, NACC, REVERSE, OUT_SHIFT, FLUSH, Params, Ctx, Carry, the array
// tables (NIN, in_w, in_ptr, NOUT, out_w, out_ptr), identity, combine, apply, load_init, expand_state,
// finish, carry_init, append_row / append_flush and step_row / step_flush.
#pragma once
#include "scan_small.cuh"
#include "smalld.cuh"

#ifndef PSSGP_NWCAP
#define PSSGP_NWCAP 8
#endif
#ifndef PSSGP_OUT_POLICY
#define PSSGP_OUT_POLICY 0  // tuning: 0 = streaming (evict-first) output stores, 1 = default policy, 2 = write-back hint, 3 = write-through
#endif
#ifndef PSSGP_FETCH_AHEAD_KINDS
#define PSSGP_FETCH_AHEAD_KINDS 2  // bit mask over KIND_FILTER (1) / KIND_SMOOTHER (2) / KIND_ADJOINT (4), apply kernels
#endif
#ifndef PSSGP_NST
#define PSSGP_NST 1  // cp.async stages per warp (1: a stage is refilled while its last row is being processed)
#endif

namespace pssgp {

// ---------------------------------------------------------------------------------------------
// cp.async + 128-bit shared/global helpers
// ---------------------------------------------------------------------------------------------
PSSGP_DEV unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int BYTES> PSSGP_DEV void cp_async(unsigned dst, const void* src) {
    static_assert(BYTES == 4 || BYTES == 8 || BYTES == 16, "cp.async size");
    if constexpr (BYTES == 16)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src) : "memory");
    else if constexpr (BYTES == 8)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(dst), "l"(src) : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(dst), "l"(src) : "memory");
}
PSSGP_DEV void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> PSSGP_DEV void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

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

// Programmatic dependent launch (PDL): a kernel launched with the programmatic-stream-serialization attribute may
// become resident while its predecessor is still running its (single-CTA) scan tail; it must not touch anything
// the predecessor produces before pdl_wait().  Both are no-ops for ordinary launches.
PSSGP_DEV void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
PSSGP_DEV void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// Compile-time geometry of the per-warp staging areas
// ---------------------------------------------------------------------------------------------
template <typename T> struct StreamGeom {
    static constexpr int LS = 16 / (int)sizeof(T);  // rows per sub-step: LS * W * sizeof(T) is a multiple of 16
    __host__ __device__ static constexpr int seg_bytes(int w) { return LS * w * (int)sizeof(T); }
    __host__ __device__ static constexpr int seg_units(int w) { return seg_bytes(w) / 16; }
    __host__ __device__ static constexpr int pitch(int w) { return (seg_units(w) | 1) * 16; }  // odd # of 16-byte units
    // cooperative copy: GROUPS whole segments per warp instruction, ITERS instructions for 32 segments
    __host__ __device__ static constexpr int groups(int w) { return 32 / seg_units(w); }
    __host__ __device__ static constexpr int iters(int w) { return (32 + groups(w) - 1) / groups(w); }
};

// Per-row output staging (OUT8): a segment is ONE row of W elements, a piece is one element.
template <typename T> struct RowGeom {
    __host__ __device__ static constexpr int pitch(int w) { return (w | 1) * (int)sizeof(T); }  // odd # of elements
    __host__ __device__ static constexpr int groups(int w) { return 32 / w; }                   // whole rows per instruction
    __host__ __device__ static constexpr int iters(int w) { return (32 + groups(w) - 1) / groups(w); }
};

template <typename Alg> struct StreamLayout {
    using T = typename Alg::scalar;
    using G = StreamGeom<T>;
    static constexpr int LS = G::LS;
    __host__ __device__ static constexpr int in_off(int a) {
        return a == 0 ? 0 : in_off(a - 1) + 32 * G::pitch(Alg::in_w(a - 1));
    }
    __host__ __device__ static constexpr int out_off(int a) {
        return a == 0 ? 0 : out_off(a - 1) + 32 * G::pitch(Alg::out_w(a - 1));
    }
    static constexpr int STAGE_BYTES = in_off(Alg::NIN);
    // OUT8 algebras stage and store their outputs one row at a time in element-sized pieces (half the staging
    // memory of the 16-byte path, which needs LS rows per segment; twice the store instructions)
    static constexpr bool OUT8 = Alg::OUT8;
    __host__ __device__ static constexpr int out_off8(int a) {
        return a == 0 ? 0 : out_off8(a - 1) + 32 * RowGeom<T>::pitch(Alg::out_w(a - 1));
    }
    static constexpr int OUT_BYTES = OUT8 ? out_off8(Alg::NOUT) : out_off(Alg::NOUT);
    static constexpr int NST = PSSGP_NST;
    static constexpr int WARP_BYTES_REDUCE = NST * STAGE_BYTES;
    static constexpr int WARP_BYTES_APPLY = NST * STAGE_BYTES + OUT_BYTES;
    static constexpr int SMEM_BUDGET = 216 * 1024;
    __host__ __device__ static constexpr int nw_fit() { return SMEM_BUDGET / WARP_BYTES_APPLY; }
    static constexpr int NWCAP = PSSGP_NWCAP;
    static constexpr int NW = nw_fit() > NWCAP ? NWCAP : (nw_fit() < 1 ? 1 : nw_fit());
};

// Per-lane cursor of the cooperative copy of one array: this lane moves piece `off` of the segments of
// lane-chunks g, g + GROUPS, g + 2 GROUPS, ...  `ptr` points at its piece of lane-chunk g for the next
// sub-step to be issued and is bumped by one segment per sub-step.
struct PieceCursor {
    unsigned char* ptr;  // global
    unsigned soff;       // byte offset of (slot g, piece off) inside the array's staging region
    int g;               // first lane-chunk handled by this lane; >= 32 when the lane idles
};

template <typename T, int W, bool REVERSE>
PSSGP_DEV PieceCursor make_cursor(const void* base, long k_lo0, int L, int nsub, int lane) {
    // k_lo0 = first row of lane 0's chunk; lane-chunk j starts at k_lo0 + j L (forward) / k_lo0 - j L (reverse)
    using G = StreamGeom<T>;
    constexpr int NP = G::seg_units(W);
    constexpr int SZ = (int)sizeof(T);
    PieceCursor pc;
    const int g = lane / NP;
    const int off = lane - g * NP;
    pc.g = (g < G::groups(W)) ? g : 32;
    pc.soff = (unsigned)(g * G::pitch(W) + off * 16);
    const long row0 = k_lo0 + (REVERSE ? -(long)g * L + (long)(nsub - 1) * G::LS : (long)g * L);  // first sub-step
    pc.ptr = (unsigned char*)base + row0 * (long)(W * SZ) + off * 16;
    return pc;
}

// Issues the copies of one array for one sub-step and advances the cursor to the next sub-step.
// step = byte distance between the segments handled by consecutive instructions of this lane.
// The fast variant (every piece of the warp lies inside the array: all but the warps at the two ends of
// the series) is one address IMAD.WIDE + one LDGSTS per instruction; the checked variant is kept out of
// line so that it costs neither registers nor instruction-cache footprint in the streaming loop.
template <typename T, int W, bool REVERSE>
PSSGP_DEV void issue_array_fast(PieceCursor& pc, int step, unsigned region_addr) {
    using G = StreamGeom<T>;
    constexpr int GR = G::groups(W), IT = G::iters(W), PITCH = G::pitch(W);
    if (pc.g < 32) {
        const unsigned dst = region_addr + pc.soff;
#pragma unroll
        for (int i = 0; i < IT - 1; ++i) cp_async<16>(dst + i * GR * PITCH, pc.ptr + (long)i * step);
        if (GR * IT <= 32 || pc.g + GR * (IT - 1) < 32)
            cp_async<16>(dst + (IT - 1) * GR * PITCH, pc.ptr + (long)(IT - 1) * step);
    }
    pc.ptr += REVERSE ? -(long)G::seg_bytes(W) : (long)G::seg_bytes(W);
}

template <typename T, int W, bool REVERSE>
__device__ __noinline__ void issue_array_checked(unsigned char* ptr, int g, const void* base, long total_bytes,
                                                 int step, unsigned dst) {
    using G = StreamGeom<T>;
    constexpr int GR = G::groups(W), IT = G::iters(W), PITCH = G::pitch(W);
    constexpr int SZ = (int)sizeof(T);
#pragma unroll 1
    for (int i = 0; i < IT; ++i) {
        if (g + GR * i < 32) {
            const unsigned char* sp = ptr + (long)i * step;
            const long gb = sp - (const unsigned char*)base;
            if (gb >= 0 && gb + 16 <= total_bytes) {
                cp_async<16>(dst + i * GR * PITCH, sp);
            } else {
#pragma unroll
                for (int e = 0; e < 16 / SZ; ++e)
                    if (gb + e * SZ >= 0 && gb + (e + 1) * SZ <= total_bytes)
                        cp_async<SZ>(dst + i * GR * PITCH + e * SZ, sp + e * SZ);
            }
        }
    }
}

PSSGP_DEV void st_out16(void* dst, const int4& v) {
#if PSSGP_OUT_POLICY == 0
    __stcs(reinterpret_cast<int4*>(dst), v);
#elif PSSGP_OUT_POLICY == 1
    *reinterpret_cast<int4*>(dst) = v;
#elif PSSGP_OUT_POLICY == 2
    __stwb(reinterpret_cast<int4*>(dst), v);
#else
    __stwt(reinterpret_cast<int4*>(dst), v);
#endif
}

// Mirror of issue_array for an output array: staging slots -> global.
template <typename T, int W, bool REVERSE>
PSSGP_DEV void store_array_fast(PieceCursor& pc, int step, const unsigned char* region) {
    using G = StreamGeom<T>;
    constexpr int GR = G::groups(W), IT = G::iters(W), PITCH = G::pitch(W);
    if (pc.g < 32) {
        const unsigned char* src = region + pc.soff;
#pragma unroll
        for (int i = 0; i < IT - 1; ++i)
            st_out16(pc.ptr + (long)i * step, *reinterpret_cast<const int4*>(src + i * GR * PITCH));
        if (GR * IT <= 32 || pc.g + GR * (IT - 1) < 32)
            st_out16(pc.ptr + (long)(IT - 1) * step, *reinterpret_cast<const int4*>(src + (IT - 1) * GR * PITCH));
    }
    pc.ptr += REVERSE ? -(long)G::seg_bytes(W) : (long)G::seg_bytes(W);
}

template <typename T, int W, bool REVERSE>
__device__ __noinline__ void store_array_checked(unsigned char* ptr, int g, const void* base, long total_bytes,
                                                 int step, const unsigned char* src) {
    using G = StreamGeom<T>;
    constexpr int GR = G::groups(W), IT = G::iters(W), PITCH = G::pitch(W);
    constexpr int SZ = (int)sizeof(T);
#pragma unroll 1
    for (int i = 0; i < IT; ++i) {
        if (g + GR * i < 32) {
            unsigned char* sp = ptr + (long)i * step;
            const long gb = sp - (const unsigned char*)base;
            if (gb >= 0 && gb + 16 <= total_bytes) {
                __stcs(reinterpret_cast<int4*>(sp), *reinterpret_cast<const int4*>(src + i * GR * PITCH));
            } else {
#pragma unroll
                for (int e = 0; e < 16 / SZ; ++e)
                    if (gb + e * SZ >= 0 && gb + (e + 1) * SZ <= total_bytes)
                        *reinterpret_cast<T*>(sp + e * SZ) = *reinterpret_cast<const T*>(src + i * GR * PITCH + e * SZ);
            }
        }
    }
}

// All input arrays of an algebra.
template <typename Alg> struct StreamIn {
    using T = typename Alg::scalar;
    using G = StreamGeom<T>;
    using Lay = StreamLayout<Alg>;
    PieceCursor pc[Alg::NIN];

    template <int A = 0> PSSGP_DEV void init(const typename Alg::Params& p, long k_lo0, int L, int nsub, int lane) {
        if constexpr (A < Alg::NIN) {
            pc[A] = make_cursor<T, Alg::in_w(A), Alg::REVERSE>(Alg::in_ptr(p, A), k_lo0, L, nsub, lane);
            init<A + 1>(p, k_lo0, L, nsub, lane);
        }
    }
    template <int A = 0>
    PSSGP_DEV void issue(const typename Alg::Params& p, long n, int L, bool fast, unsigned stage_addr) {
        if constexpr (A < Alg::NIN) {
            constexpr int W = Alg::in_w(A);
            const int step = (Alg::REVERSE ? -L : L) * (int)(W * sizeof(T)) * G::groups(W);
            if (fast) {
                issue_array_fast<T, W, Alg::REVERSE>(pc[A], step, stage_addr + Lay::in_off(A));
            } else {
                issue_array_checked<T, W, Alg::REVERSE>(pc[A].ptr, pc[A].g, Alg::in_ptr(p, A), n * (long)(W * sizeof(T)),
                                                        step, stage_addr + Lay::in_off(A) + pc[A].soff);
                pc[A].ptr += Alg::REVERSE ? -(long)G::seg_bytes(W) : (long)G::seg_bytes(W);
            }
            issue<A + 1>(p, n, L, fast, stage_addr);
        }
    }
};

template <typename Alg> struct StreamOut {
    using T = typename Alg::scalar;
    using G = StreamGeom<T>;
    using Lay = StreamLayout<Alg>;
    PieceCursor pc[Alg::NOUT];

    template <int A = 0> PSSGP_DEV void init(const typename Alg::Params& p, long k_lo0, int L, int nsub, int lane) {
        if constexpr (A < Alg::NOUT) {
            pc[A] = make_cursor<T, Alg::out_w(A), Alg::REVERSE>(Alg::out_ptr(p, A), k_lo0, L, nsub, lane);
            init<A + 1>(p, k_lo0, L, nsub, lane);
        }
    }
    template <int A = 0>
    PSSGP_DEV void store(const typename Alg::Params& p, long n, int L, bool fast, const unsigned char* ostage) {
        if constexpr (A < Alg::NOUT) {
            constexpr int W = Alg::out_w(A);
            const int step = (Alg::REVERSE ? -L : L) * (int)(W * sizeof(T)) * G::groups(W);
            if (fast) {
                store_array_fast<T, W, Alg::REVERSE>(pc[A], step, ostage + Lay::out_off(A));
            } else {
                store_array_checked<T, W, Alg::REVERSE>(pc[A].ptr, pc[A].g, Alg::out_ptr(p, A), n * (long)(W * sizeof(T)),
                                                        step, ostage + Lay::out_off(A) + pc[A].soff);
                pc[A].ptr += Alg::REVERSE ? -(long)G::seg_bytes(W) : (long)G::seg_bytes(W);
            }
            store<A + 1>(p, n, L, fast, ostage);
        }
    }
};

// Bounds-checked variant of the per-row store (tail CTA only; out of line like issue_array_checked).
template <typename T, int W>
__device__ __noinline__ void store_row_checked(unsigned char* ptr, int g, const void* base, long total_bytes, long step,
                                               const unsigned char* src) {
    using G = RowGeom<T>;
    constexpr int GR = G::groups(W), IT = G::iters(W), PITCH = G::pitch(W);
#pragma unroll 1
    for (int i = 0; i < IT; ++i) {
        if (g + GR * i < 32) {
            unsigned char* dp = ptr + (long)i * step;
            const long gb = dp - (const unsigned char*)base;
            if (gb >= 0 && gb + (long)sizeof(T) <= total_bytes)
                *reinterpret_cast<T*>(dp) = *reinterpret_cast<const T*>(src + i * GR * PITCH);
        }
    }
}

// Output arrays of an OUT8 algebra: one cursor per array pointing at this lane's piece of the NEXT row to be
// written (rows are written in visiting order, one per call, for every array of a shift class at once).
template <typename Alg> struct StreamOut8 {
    using T = typename Alg::scalar;
    using G = RowGeom<T>;
    using Lay = StreamLayout<Alg>;
    PieceCursor pc[Alg::NOUT];

    template <int A = 0> PSSGP_DEV void init(const typename Alg::Params& p, long k_lo0, int L, int lane) {
        if constexpr (A < Alg::NOUT) {
            constexpr int W = Alg::out_w(A);
            const int g = lane / W, off = lane - g * W;
            pc[A].g = (g < G::groups(W)) ? g : 32;
            pc[A].soff = (unsigned)(g * G::pitch(W) + off * (int)sizeof(T));
            const long row0 = k_lo0 + (Alg::REVERSE ? -(long)g * L + (L - 1) : (long)g * L);
            pc[A].ptr = (unsigned char*)Alg::out_ptr(p, A) + (row0 * W + off) * (long)sizeof(T);
            init<A + 1>(p, k_lo0, L, lane);
        }
    }
    // stores the staged row of every array whose out_shift is SHIFT and moves their cursors to the next row
    template <int SHIFT, int A = 0>
    PSSGP_DEV void store(const typename Alg::Params& p, long n, int L, bool fast, const unsigned char* ostage) {
        if constexpr (A < Alg::NOUT) {
            if constexpr (Alg::out_shift(A) == SHIFT) {
                constexpr int W = Alg::out_w(A);
                constexpr int GR = G::groups(W), IT = G::iters(W), PITCH = G::pitch(W);
                const long step = (long)(Alg::REVERSE ? -L : L) * (long)(W * sizeof(T)) * GR;
                const unsigned char* src = ostage + Lay::out_off8(A) + pc[A].soff;
                if (fast) {
                    if (pc[A].g < 32) {
#pragma unroll
                        for (int i = 0; i < IT; ++i)
                            if (GR * IT <= 32 || i < IT - 1 || pc[A].g + GR * i < 32)
                                __stcs(reinterpret_cast<T*>(pc[A].ptr + (long)i * step),
                                       *reinterpret_cast<const T*>(src + i * GR * PITCH));
                    }
                } else {
                    store_row_checked<T, W>(pc[A].ptr, pc[A].g, Alg::out_ptr(p, A), n * (long)(W * sizeof(T)), step, src);
                }
                pc[A].ptr += Alg::REVERSE ? -(long)(W * sizeof(T)) : (long)(W * sizeof(T));
            }
            store<SHIFT, A + 1>(p, n, L, fast, ostage);
        }
    }
};

// Registers -> this lane's one-row staging slot of every output array whose bit (1 << out_shift) is set in mask.
template <typename Alg, int A = 0>
PSSGP_DEV void stream_stage_out_row8(unsigned char* ostage, int lane, int mask,
                                     const typename Alg::scalar (&orow)[Alg::NOUT][Alg::WMAX]) {
    using T = typename Alg::scalar;
    using Lay = StreamLayout<Alg>;
    if constexpr (A < Alg::NOUT) {
        constexpr int W = Alg::out_w(A);
        if ((mask >> Alg::out_shift(A)) & 1) {
            T* dst = reinterpret_cast<T*>(ostage + Lay::out_off8(A) + lane * RowGeom<T>::pitch(W));
#pragma unroll
            for (int e = 0; e < W; ++e) dst[e] = orow[A][e];
        }
        stream_stage_out_row8<Alg, A + 1>(ostage, lane, mask, orow);
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

// Same with a run-time row index (element-sized shared loads): lets the row loop stay rolled, which halves the
// code of the loop body (OUT8 algebras: the fused reverse step is otherwise too large for the instruction cache).
template <typename Alg, int A = 0>
PSSGP_DEV void stream_fetch_row_rt(const unsigned char* stage, int lane, int r,
                                   typename Alg::scalar (&row)[Alg::NIN][Alg::WMAX]) {
    using T = typename Alg::scalar;
    using G = StreamGeom<T>;
    using Lay = StreamLayout<Alg>;
    if constexpr (A < Alg::NIN) {
        constexpr int W = Alg::in_w(A);
        const T* src = reinterpret_cast<const T*>(stage + Lay::in_off(A) + lane * G::pitch(W)) + r * W;
#pragma unroll
        for (int e = 0; e < W; ++e) row[A][e] = src[e];
        stream_fetch_row_rt<Alg, A + 1>(stage, lane, r, row);
    }
}

// Registers (row r) -> this lane's staging slot of every output array; units that straddle two rows are
// written element by element.
template <typename Alg, int A = 0>
PSSGP_DEV void stream_stage_out_row(unsigned char* ostage, int lane, int r,
                                    const typename Alg::scalar (&orow)[Alg::NOUT][Alg::WMAX]) {
    using T = typename Alg::scalar;
    using G = StreamGeom<T>;
    using Lay = StreamLayout<Alg>;
    if constexpr (A < Alg::NOUT) {
        constexpr int W = Alg::out_w(A);
        constexpr int EPU = 16 / (int)sizeof(T);
        unsigned char* dst = ostage + Lay::out_off(A) + lane * G::pitch(W);
        const int e_lo = r * W, e_hi = (r + 1) * W;
#pragma unroll
        for (int u = 0; u < G::seg_units(W); ++u) {
            if (u * EPU >= e_lo && (u + 1) * EPU <= e_hi) {
                T tmp[EPU];
#pragma unroll
                for (int j = 0; j < EPU; ++j) tmp[j] = orow[A][u * EPU + j - e_lo];
                st_shared16(dst + u * 16, tmp);
            } else if ((u + 1) * EPU > e_lo && u * EPU < e_hi) {
#pragma unroll
                for (int j = 0; j < EPU; ++j) {
                    const int e = u * EPU + j;
                    if (e >= e_lo && e < e_hi) *reinterpret_cast<T*>(dst + e * (int)sizeof(T)) = orow[A][e - e_lo];
                }
            }
        }
        stream_stage_out_row<Alg, A + 1>(ostage, lane, r, orow);
    }
}

// Partition of the time axis (host-computed): nMain "main" CTAs of NW warps x 32 complete chunks, followed in
// time by at most one "tail" CTA that covers the remaining rows with its own (short) chunk length Ltail.  Inside a
// main CTA the first warps (in time order) have chunks of L rows and the others of L - LS rows: wl + 1 long warps
// in the first cl CTAs, wl in the others, so that nMain can be (number of SMs - 1) whatever n is, every main
// CTA carries the same number of rows to within 32 * LS, and fewer than 32 * LS rows are left for the tail CTA.
// Only the tail CTA ever sees a missing row, so only it pays for bounds checks.
struct StreamPart {
    long n;
    int L, Ltail;
    int wl, cl;
    int nMain, nCta;
};

// Geometry shared by K1 and K3.
template <typename Alg> struct WarpGeom {
    long lc;      // global logical chunk (scan order): index into the workspace arrays
    long k_lo;    // first row of this lane's chunk
    long k_lo0;   // first row of lane 0's chunk
    int L, nsub;  // rows per chunk and sub-steps of this warp
    bool fast;    // main CTA: every row of every chunk exists
    PSSGP_DEV WarpGeom(const StreamPart& sp, int NW, int lane, int wid) {
        constexpr int LS = StreamGeom<typename Alg::scalar>::LS;
        const int tb = Alg::REVERSE ? (sp.nCta - 1 - (int)blockIdx.x) : (int)blockIdx.x;  // CTA in time order
        fast = tb < sp.nMain;
        const int tbm = fast ? tb : sp.nMain;  // main CTAs before this one
        const long long_before = (long)tbm * sp.wl + (tbm < sp.cl ? tbm : sp.cl);  // long warps before this CTA
        const long row_begin = ((long)tbm * NW * (sp.L - LS) + long_before * LS) * 32;
        const int tl = wid * 32 + lane;                             // thread in scan order
        const int ct = Alg::REVERSE ? (NW * 32 - 1 - tl) : tl;      // chunk of the CTA in time order
        lc = (long)blockIdx.x * (NW * 32) + tl;
        if (fast) {
            const int nl = sp.wl + (tb < sp.cl ? 1 : 0);  // long warps of this CTA
            const int wt = ct >> 5, lw = ct & 31;         // warp (time order) and chunk inside it
            L = wt < nl ? sp.L : sp.L - LS;
            k_lo = row_begin + ((long)wt * (sp.L - LS) + (long)(wt < nl ? wt : nl) * LS) * 32 + (long)lw * L;
        } else {
            L = sp.Ltail;
            k_lo = row_begin + (long)ct * L;
        }
        nsub = L / LS;
        k_lo0 = Alg::REVERSE ? (k_lo + (long)lane * L) : (k_lo - (long)lane * L);
    }
};

// ---------------------------------------------------------------------------------------------
// CTA-level scan of the per-thread chunk aggregates (tail of K1; also used by the fused forward kernel for
// the aggregates of the reverse scans, FLIP = true: the thread that holds time-chunk c of a CTA plays the
// role of scan-order thread NW*32-1-c, exactly as WarpGeom maps the threads of a REVERSE algebra).
// Stores the lane-exclusive prefix of every chunk, the CTA-exclusive prefix of every warp and the CTA total.
// blk_s = this CTA's index in scan order.  shw: NW * NAGG scalars of shared memory.  All threads must call.
// ---------------------------------------------------------------------------------------------
template <typename Alg, int NW, bool FLIP>
PSSGP_DEV void cta_scan_publish(const typename Alg::scalar (&a_in)[Alg::NAGG], int lane, int wid, long blk_s, long nCta,
                                long nChunksPad, typename Alg::scalar* __restrict__ lane_excl,
                                typename Alg::scalar* __restrict__ warp_excl, typename Alg::scalar* __restrict__ wagg,
                                typename Alg::scalar* shw) {
    using T = typename Alg::scalar;
    // private copy: it is handed to the out-of-line operators by address, which would otherwise pull the caller's
    // register-resident aggregate into local memory for the whole kernel
    T a[Alg::NAGG];
#pragma unroll
    for (int e = 0; e < Alg::NAGG; ++e) a[e] = FLIP ? shfl_idx_t(a_in[e], 31 - lane) : a_in[e];
    if (FLIP) wid = NW - 1 - wid;
    const long lc = blk_s * (NW * 32) + wid * 32 + lane;
    // Warp-inclusive scan (5 shuffle steps; the earlier lane is the left operand), then the scan over the NW warp
    // totals by warp 0, through ONE loop body so that the CTA executes a single inlined copy of combine
    // (see scan_mid_body about cold instruction fetches).
    constexpr int LOGNW = NW > 16 ? 5 : (NW > 8 ? 4 : (NW > 4 ? 3 : (NW > 2 ? 2 : (NW > 1 ? 1 : 0))));
#pragma unroll 1
    for (int lvl = 0;; ++lvl) {
        if (lvl == 5) {
            // lane-exclusive prefix inside the warp
            T ex[Alg::NAGG];
#pragma unroll
            for (int e = 0; e < Alg::NAGG; ++e) ex[e] = shfl_up_t(a[e], 1);
            if (lane != 0) {
#pragma unroll
                for (int e = 0; e < Alg::NAGG; ++e) lane_excl[(long)e * nChunksPad + lc] = ex[e];
            }
            if (lane == 31) {
#pragma unroll
                for (int e = 0; e < Alg::NAGG; ++e) shw[wid * Alg::NAGG + e] = a[e];
            }
            __syncthreads();
            if (wid != 0) break;
            if (lane < NW) {
#pragma unroll
                for (int e = 0; e < Alg::NAGG; ++e) a[e] = shw[lane * Alg::NAGG + e];
            } else {
                Alg::identity(a);
            }
        }
        if (lvl == 5 + LOGNW) break;
        const int off = 1 << (lvl < 5 ? lvl : lvl - 5);
        T o[Alg::NAGG];
#pragma unroll
        for (int e = 0; e < Alg::NAGG; ++e) o[e] = shfl_up_t(a[e], off);
        if (lane >= off && (lvl < 5 || lane < NW)) {
            T r[Alg::NAGG];
            Alg::combine(o, a, r);
#pragma unroll
            for (int e = 0; e < Alg::NAGG; ++e) a[e] = r[e];
        }
    }
    if (wid == 0) {
        // a = inclusive prefix over warps: warp l+1's exclusive prefix, and the CTA total at lane NW-1
        const long gw = blk_s * NW + lane + 1;
        if (lane < NW - 1) {
#pragma unroll
            for (int e = 0; e < Alg::NAGG; ++e) warp_excl[(long)e * (nCta * NW) + gw] = a[e];
        }
        if (lane == NW - 1) {
#pragma unroll
            for (int e = 0; e < Alg::NAGG; ++e) wagg[(long)e * nCta + blk_s] = a[e];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K1: chunk aggregates, warp-exclusive prefixes, CTA-exclusive warp prefixes, CTA totals
// ---------------------------------------------------------------------------------------------
// PREFIX = false: the last CTA turns the CTA totals into the states entering the CTAs (wstate; nullptr: none);
// PREFIX = true (time sharding): into per-CTA prefix aggregates + the shard summary (scan_prefix_body).  Two
// instantiations, so that the single-shard kernel does not carry the code of the other tail.
template <typename Alg, bool PREFIX = false>
__global__ void __launch_bounds__(StreamLayout<Alg>::NW * 32)
stream_reduce_kernel(typename Alg::Params p, StreamPart sp, long nChunksPad,
                     typename Alg::scalar* __restrict__ lane_excl, typename Alg::scalar* __restrict__ warp_excl,
                     typename Alg::scalar* __restrict__ wagg, typename Alg::scalar* wstate,
                     typename Alg::scalar* final_state, unsigned int* ticket, typename Alg::scalar* wprefix,
                     typename Alg::scalar* total_out) {
    using T = typename Alg::scalar;
    using Lay = StreamLayout<Alg>;
    constexpr int NW = Lay::NW, NST = Lay::NST, LS = Lay::LS;
    const long nCta = sp.nCta;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned char* wsm = smem_raw + wid * Lay::WARP_BYTES_REDUCE;
    const unsigned wsm_addr = smem_addr(wsm);
    const WarpGeom<Alg> wg(sp, NW, lane, wid);
    const int nsub = wg.nsub, L = wg.L;
    const long n = sp.n;
    PSSGP_PHASE(0);

    T a[Alg::NAGG];
    Alg::identity(a);
    {
        typename Alg::Ctx ctx;
        Alg::load_ctx(p, ctx);
        StreamIn<Alg> in;
        in.init(p, wg.k_lo0, L, nsub, lane);
#pragma unroll 1
        for (int s = 0; s < NST; ++s) {
            if (s < nsub) in.issue(p, n, L, wg.fast, wsm_addr + s * Lay::STAGE_BYTES);
            cp_async_commit();
        }
        const bool mine = wg.k_lo < n;
        typename Alg::Carry cr;
        const long k_hi = wg.k_lo + L;
        if (mine) Alg::carry_init(cr, ctx, wg.k_lo, k_hi < n ? k_hi : n, p);
#pragma unroll 1
        for (int s = 0; s < nsub; ++s) {
            const int st = s % NST;
            cp_async_wait<NST - 1>();
            __syncwarp();
            const long k0 = wg.k_lo + (long)(Alg::REVERSE ? (nsub - 1 - s) : s) * LS;
#pragma unroll
            for (int rr = 0; rr < LS; ++rr) {
                const int r = Alg::REVERSE ? (LS - 1 - rr) : rr;
                const long k = k0 + r;
                T row[Alg::NIN][Alg::WMAX];
                stream_fetch_row<Alg>(wsm + st * Lay::STAGE_BYTES, lane, r, row);
                if (rr == LS - 1) {
                    // the stage is drained: hand it back to the copy engine before the last row's arithmetic
                    __syncwarp();
                    if (s + NST < nsub) in.issue(p, n, L, wg.fast, wsm_addr + st * Lay::STAGE_BYTES);
                    cp_async_commit();
                }
#ifdef PSSGP_DRYRUN  // memory-pipeline probe: no arithmetic (results are wrong)
                if (mine && k < n) a[0] += row[0][0] + row[1][Alg::in_w(1) - 1] + row[Alg::NIN - 1][0];
#else
                if (mine && k < n) Alg::append_row(a, ctx, row, k, p, cr);
#endif
            }
        }
        cp_async_wait<0>();
        if constexpr (Alg::FLUSH) {
            if (mine) Alg::append_flush(a, ctx, wg.k_lo, p, cr);
        }
    }
    __shared__ T shw[NW * Alg::NAGG];
    pdl_launch_dependents();  // the next kernel may start occupying the SMs this kernel's scan tail leaves idle
    PSSGP_PHASE(1);
    cta_scan_publish<Alg, NW, false>(a, lane, wid, (long)blockIdx.x, nCta, nChunksPad, lane_excl, warp_excl, wagg, shw);
    PSSGP_PHASE(2);
    // K2 folded into K1: the CTA that finishes last scans the CTA totals (wstate == nullptr: the caller runs
    // scan_mid_kernel / scan_total_kernel itself)
    if (PREFIX || wstate != nullptr) {
        __shared__ bool is_last;
        __shared__ T sh_mid[32 * Alg::NAGG];
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned int t = atomicAdd(ticket, 1u);
            is_last = (t == gridDim.x - 1);
        }
        __syncthreads();
        PSSGP_PHASE(3);
        if (is_last) {
            __threadfence();
            if constexpr (PREFIX)
                scan_prefix_body<Alg>(wagg, nCta, wprefix, total_out, sh_mid, (int)threadIdx.x, (int)blockDim.x, 0);
            else
                scan_mid_body<Alg>(p, wagg, nCta, wstate, final_state, sh_mid, (int)threadIdx.x, (int)blockDim.x, 0);
            if (threadIdx.x == 0) *ticket = 0u;
            __syncthreads();
            PSSGP_PHASE(7);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K3: seeded recursion + outputs + deterministic reduction of the accumulators
// ---------------------------------------------------------------------------------------------
template <typename Alg>
__global__ void __launch_bounds__(StreamLayout<Alg>::NW * 32)
stream_apply_kernel(typename Alg::Params p, StreamPart sp, long nChunksPad,
                    const typename Alg::scalar* __restrict__ lane_excl,
                    const typename Alg::scalar* __restrict__ warp_excl,
                    const typename Alg::scalar* __restrict__ wstate,
                    typename Alg::scalar* __restrict__ acc_part, unsigned int* __restrict__ ticket,
                    typename Alg::scalar* __restrict__ acc_out, const typename Alg::scalar* __restrict__ wprefix) {
    using T = typename Alg::scalar;
    using Lay = StreamLayout<Alg>;
    constexpr int NW = Lay::NW, NST = Lay::NST, LS = Lay::LS;
    constexpr int NACC1 = Alg::NACC > 0 ? Alg::NACC : 1;
    constexpr int ROW_UNROLL = Lay::OUT8 ? 1 : LS;  // OUT8: one rolled copy of the row body (run-time row index)
    // FETCH_AHEAD (macro PSSGP_FETCH_AHEAD_KINDS, bit per Alg::KIND): all LS rows of a stage go to registers at once
    // and the stage is refilled before the first of them is processed: the copy has LS rows of arithmetic to land
    constexpr bool FETCH_AHEAD = !Lay::OUT8 && !Alg::HAS_SIDE && ((PSSGP_FETCH_AHEAD_KINDS >> Alg::KIND) & 1) != 0;
    const long nCta = sp.nCta;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned char* wsm = smem_raw + wid * Lay::WARP_BYTES_APPLY;
    unsigned char* osm = wsm + NST * Lay::STAGE_BYTES;
    const unsigned wsm_addr = smem_addr(wsm);
    const WarpGeom<Alg> wg(sp, NW, lane, wid);
    const int nsub = wg.nsub, L = wg.L;
    const long n = sp.n;

    T acc[NACC1];
#pragma unroll
    for (int e = 0; e < NACC1; ++e) acc[e] = T(0);
    pdl_wait();  // everything below reads what the previous kernel of the scan wrote (states, aggregates, moments)
    {
        typename Alg::Ctx ctx;
        Alg::load_ctx(p, ctx);
        StreamIn<Alg> in;
        in.init(p, wg.k_lo0, L, nsub, lane);
        StreamOut<Alg> out;
        StreamOut8<Alg> out8;
        if constexpr (Lay::OUT8) out8.init(p, wg.k_lo0, L, lane);
        else out.init(p, wg.k_lo0, L, nsub, lane);
#pragma unroll 1
        for (int s = 0; s < NST; ++s) {
            if (s < nsub) in.issue(p, n, L, wg.fast, wsm_addr + s * Lay::STAGE_BYTES);
            cp_async_commit();
        }
        const bool mine = wg.k_lo < n;
        T st8[Alg::NSTATE];
        {
            // state entering this chunk = CTA state o warp prefix o lane prefix, on temporaries; one rolled loop so
            // that the kernel carries ONE inlined copy of apply (cold instruction fetch, see scan_mid_body).
            // Time sharding (wprefix != nullptr): the CTA state is (state entering the shard) o (prefix of this CTA).
            T sl[Alg::NSTATE];
            if (wprefix != nullptr) {
                Alg::load_init(p, sl);
                if constexpr (Alg::HAS_SIDE) Alg::publish_init(p, sl);
            } else {
#pragma unroll
                for (int e = 0; e < Alg::NSTATE; ++e) sl[e] = wstate[(long)e * nCta + blockIdx.x];
            }
#pragma unroll 1
            for (int j = 0; j < 3; ++j) {
                const typename Alg::scalar* src;
                long stride, idx;
                bool act;
                if (j == 0) {
                    act = wprefix != nullptr && blockIdx.x != 0;
                    src = wprefix, stride = nCta, idx = blockIdx.x;
                } else if (j == 1) {
                    act = wid != 0;
                    src = warp_excl, stride = nCta * NW, idx = (long)blockIdx.x * NW + wid;
                } else {
                    act = lane != 0 && mine;
                    src = lane_excl, stride = nChunksPad, idx = wg.lc;
                }
                if (act) {
                    T ex[Alg::NAGG], s2[Alg::NSTATE];
#pragma unroll
                    for (int e = 0; e < Alg::NAGG; ++e) ex[e] = src[(long)e * stride + idx];
                    Alg::apply(sl, ex, s2);
#pragma unroll
                    for (int e = 0; e < Alg::NSTATE; ++e) sl[e] = s2[e];
                }
            }
#pragma unroll
            for (int e = 0; e < Alg::NSTATE; ++e) st8[e] = sl[e];
        }
        typename Alg::Carry cr;
        const long k_hi = wg.k_lo + L;
        if (mine || Alg::HAS_SIDE) Alg::carry_init(cr, ctx, wg.k_lo, k_hi < n ? k_hi : n, p);
#pragma unroll 1
        for (int s = 0; s < nsub; ++s) {
            const int st = s % NST;
            cp_async_wait<NST - 1>();
            __syncwarp();
            const long k0 = wg.k_lo + (long)(Alg::REVERSE ? (nsub - 1 - s) : s) * LS;
            T rows_all[FETCH_AHEAD ? LS : 1][Alg::NIN][Alg::WMAX];
            if constexpr (FETCH_AHEAD) {
#pragma unroll
                for (int rr = 0; rr < LS; ++rr)
                    stream_fetch_row<Alg>(wsm + st * Lay::STAGE_BYTES, lane, Alg::REVERSE ? (LS - 1 - rr) : rr, rows_all[rr]);
                __syncwarp();
                if (s + NST < nsub) in.issue(p, n, L, wg.fast, wsm_addr + st * Lay::STAGE_BYTES);
                cp_async_commit();
            }
#pragma unroll ROW_UNROLL
            for (int rr = 0; rr < LS; ++rr) {
                const int r = Alg::REVERSE ? (LS - 1 - rr) : rr;
                const long k = k0 + r;
                T row[Alg::NIN][Alg::WMAX];
                if constexpr (FETCH_AHEAD) {
#pragma unroll
                    for (int a = 0; a < Alg::NIN; ++a)
#pragma unroll
                        for (int e = 0; e < Alg::WMAX; ++e) row[a][e] = rows_all[rr][a][e];
                } else if constexpr (Lay::OUT8) {
                    stream_fetch_row_rt<Alg>(wsm + st * Lay::STAGE_BYTES, lane, r, row);
                } else {
                    stream_fetch_row<Alg>(wsm + st * Lay::STAGE_BYTES, lane, r, row);
                }
                if (!FETCH_AHEAD && rr == LS - 1) {
                    // the stage is drained: hand it back to the copy engine before the last row's arithmetic
                    __syncwarp();
                    if (s + NST < nsub) in.issue(p, n, L, wg.fast, wsm_addr + st * Lay::STAGE_BYTES);
                    cp_async_commit();
                }
                T orow[Alg::NOUT][Alg::WMAX];
                if constexpr (Lay::OUT8) {
                    // one row at a time: step_row returns a mask of out_shift classes it produced (bit 0: outputs
                    // of row k, bit 1: outputs of row k+1, delayed by one visit)
                    int hm = 0;
                    if (mine && k < n) hm = Alg::step_row(st8, ctx, row, orow, k, p, acc, cr);
                    stream_stage_out_row8<Alg>(osm, lane, hm, orow);
                    __syncwarp();
                    out8.template store<0>(p, n, L, wg.fast, osm);
                    if (!(s == 0 && rr == 0)) out8.template store<1>(p, n, L, wg.fast, osm);
                    __syncwarp();
                } else {
                    bool has = false;
                    if (mine && k < n) has = Alg::step_row(st8, ctx, row, orow, k, p, acc, cr);
                    if (Alg::OUT_SHIFT == 0) {
                        if (has) stream_stage_out_row<Alg>(osm, lane, r, orow);
                    } else {
                        // the step of row k+1 is taken when row k is visited (reverse scans only): its outputs
                        // belong to the slot above; the top row completes the segment of the previous sub-step
                        if (has) stream_stage_out_row<Alg>(osm, lane, (r + 1) % LS, orow);
                        if (rr == 0 && s > 0) {
                            __syncwarp();
                            out.store(p, n, L, wg.fast, osm);
                            __syncwarp();
                        }
                    }
                }
            }
            if (!Lay::OUT8 && Alg::OUT_SHIFT == 0) {
                __syncwarp();
                out.store(p, n, L, wg.fast, osm);
                __syncwarp();  // staging slots are rewritten by the next sub-step
            }
        }
        cp_async_wait<0>();
        pdl_launch_dependents();
        if constexpr (Alg::FLUSH) {
            T orow[Alg::NOUT][Alg::WMAX];
            if constexpr (Lay::OUT8) {
                int hm = 0;
                if (mine) hm = Alg::step_flush(st8, ctx, orow, wg.k_lo, p, acc, cr);
                stream_stage_out_row8<Alg>(osm, lane, hm & 2, orow);
                __syncwarp();
                out8.template store<1>(p, n, L, wg.fast, osm);
            } else {
                bool has = false;
                if (mine) has = Alg::step_flush(st8, ctx, orow, wg.k_lo, p, acc, cr);
                if (has) stream_stage_out_row<Alg>(osm, lane, 0, orow);
                __syncwarp();
                out.store(p, n, L, wg.fast, osm);
            }
        }
        if constexpr (Alg::HAS_DONE) {
            if (mine) Alg::step_done(acc, cr);
        }
        if constexpr (Alg::HAS_SIDE) {
            // chunk aggregates of the scans that run in the opposite direction (fused_small.cuh)
            if (mine) Alg::side_flush(st8, k_hi < n ? k_hi : n, p, cr);
            Alg::template side_finish<NW>(p, cr, lane, wid, nCta, nChunksPad);
        }
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
