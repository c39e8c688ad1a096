// svp_kernels.cuh — sm_100a kernels of the alignment hot path (ALIGN_SPEC v1).
//
// Forward DP   svp_fwd_kernel<KP, MULTI>
//   One CTA per pair, NW warps. The diagonal band is split over threads: thread g owns the 4*KP
//   consecutive diagonals band_lo + g*4*KP ...; its even diagonals form "set A", its odd ones
//   "set B". The CTA marches anti-diagonal by anti-diagonal (step r = i + j): on an A-step every
//   thread updates its 2*KP set-A cells, on the next step its 2*KP set-B cells, so a warp-step
//   updates 64*KP cells, all independent. Two cells are packed per 32-bit register (s16x2:
//   lo = slot k, hi = slot KP+k) and updated with the DPX/video SIMD instructions
//   VIADDMNMX.S16x2 / VIMNMX3.S16x2(.RELU) / VIADD.16x2; match/mismatch comes from one HSET2.EQ
//   on fp16-coded bases. Neighbour diagonals that belong to the adjacent thread travel by
//   __shfl_up/__shfl_down (one register per state array per step), to the adjacent warp through
//   a 16-byte shared-memory mailbox and one CTA barrier per step. Sequences are staged 2-bit
//   packed in shared memory (cp.async 16-byte copies). Per step each thread streams the low byte
//   of its 2*KP H values to the traceback arena in HBM (one 8/16-byte st.global.cs).
//
// Traceback    svp_tb_kernel<KP>
//   One warp per pair. Rebuilds exact H values along the path from the stored low bytes
//   (neighbouring H differ by less than 128), 32 cells per round trip to HBM.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_pipeline.h>
#include <cstdint>

#include "../../include/svp_align.h"

namespace svp {

struct Task {               // one pair, prepared on the host (svp_align.cu: plan_pair)
    uint64_t q_off, t_off;  // byte offsets of the packed sequences
    int32_t  m, n;          // query / target length
    int32_t  band_lo, band_w;
    int32_t  r_begin;       // anti-diagonal (i + j) of step 0; (r_begin - band_lo) is even
    int32_t  n_iter;        // loop iterations; one iteration = 2*KP steps
    uint32_t flags;
    uint32_t pair_index;    // index into the caller's pair / result arrays
    uint64_t tb_off;        // byte offset of this pair's traceback rows in the arena
    uint64_t best_off;      // first entry of this pair in the per-thread best array
    uint64_t cells;         // spec cells (band ∩ matrix), copied into the result
    uint32_t cigar_off, cigar_cap;
};

struct Consts {             // scoring in the packed form the kernels use
    uint32_t mat, mis;      // s16x2: match score, mismatch score (negative)
    uint32_t ne1, ne2;      // -ext1, -ext2
    uint32_t noe1, noe2;    // -(open1+ext1), -(open2+ext2)
    int32_t  two_piece;
    // scalar copies for the traceback
    int32_t  s_match, s_mismatch, open1, ext1, open2, ext2;
    int32_t  min_signal, zdrop, zdrop_ext;
};

constexpr uint32_t TASK_INVALID = 0x80000000u;   // Task.flags: descriptor rejected by the planner
constexpr uint32_t TASK_TOO_LONG = 0x40000000u;  // ... because the score range exceeds s16
constexpr uint32_t TASK_KP8 = 0x100u;            // pair runs on the KP=8 kernels (bands wider than 16 warps x 512)
constexpr uint32_t NEGV = 0xC000C000u;   // s16x2 -16384: "minus infinity" for gap states
constexpr uint32_t QNAN16 = 0x7E00u;     // fp16 NaN: sentinel base, never equal to anything

__device__ __forceinline__ uint32_t pk(int v) { return (uint32_t)(v & 0xFFFF) * 0x00010001u; }

__device__ __forceinline__ uint32_t heq_mask(uint32_t a, uint32_t b) {
    __half2 ha = *reinterpret_cast<__half2*>(&a), hb = *reinterpret_cast<__half2*>(&b);
    return __heq2_mask(ha, hb);
}

// ---------------------------------------------------------------------------------------------
// forward DP
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pin(uint32_t v) {      // keep a constant in a register (no LDC re-loads in the loop)
    uint32_t r;
    asm volatile("mov.b32 %0, %1;" : "=r"(r) : "r"(v));
    return r;
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t saddr) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}
__device__ __forceinline__ uint4 lds_v4(uint32_t saddr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
    return v;
}
__device__ __forceinline__ void sts_v4(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// Shared-memory image of a sequence: one byte per base = the HIGH byte of its fp16 code (0x3C..0x3F),
// base p (0-based) at byte SEQ_PAD + p, with 0x7E (NaN) in the bytes before and after; positions are
// clamped into [SEQ_PAD-1, SEQ_PAD+len] so every out-of-range position reads a NaN: virtual cells
// never match and no range test is needed in the loop.
constexpr int SEQ_PAD = 16;
__host__ __device__ __forceinline__ int seq_image_bytes(int len) { return (SEQ_PAD + len + 1 + 15) & ~15; }

// 16 bases of one 2-bit packed word -> 16 code bytes (four 32-bit words)
__device__ __forceinline__ uint4 expand16(uint32_t w) {
    uint32_t o[4];
    #pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t x = (w >> (8 * k)) & 0xFFu;
        o[k] = ((x | (x << 6) | (x << 12) | (x << 18)) & 0x03030303u) + 0x3C3C3C3Cu;
    }
    return make_uint4(o[0], o[1], o[2], o[3]);
}

template <int KP, bool MULTI, bool TWO>
__global__ void __launch_bounds__(MULTI ? 512 : 32)
svp_fwd_kernel(const Task* __restrict__ tasks, const uint32_t* __restrict__ order, const uint8_t* __restrict__ seqs,
               uint8_t* __restrict__ tb, uint2* __restrict__ bests, const Consts cs) {
    extern __shared__ __align__(16) uint8_t smem8[];
    const Task tk = tasks[order[blockIdx.x]];
    const int g = threadIdx.x, lane = g & 31, warp = g >> 5, nwarp = blockDim.x >> 5;
    const int qbytes = seq_image_bytes(tk.m), tbytes = seq_image_bytes(tk.n);
    uint8_t* qimg = smem8;
    uint8_t* timg = smem8 + qbytes;
    const uint32_t s_q = (uint32_t)__cvta_generic_to_shared(qimg) + SEQ_PAD - 1;   // + clamped 1-based position
    const uint32_t s_t = (uint32_t)__cvta_generic_to_shared(timg) + SEQ_PAD - 1;
    // mailboxes (MULTI): A[w] = lane 0 of warp w after an A-step, A[nwarp] = constants;
    //                    B[w+1] = lane 31 of warp w after a B-step, B[0] = constants
    const uint32_t s_mbA = (uint32_t)__cvta_generic_to_shared(smem8 + qbytes + tbytes);
    const uint32_t s_mbB = s_mbA + (uint32_t)(nwarp + 1) * 16u;

    const uint32_t c_ne1 = pin(cs.ne1), c_noe1 = pin(cs.noe1), c_ne2 = pin(cs.ne2), c_noe2 = pin(cs.noe2);
    const uint32_t c_mat = pin(cs.mat), c_mis = pin(cs.mis), c_neg = pin(NEGV);

    // ---- stage both sequences as code bytes ------------------------------------------------------
    {
        const uint32_t* tsrc = reinterpret_cast<const uint32_t*>(seqs + tk.t_off);
        for (int w = g; w < (tk.n + 15) >> 4; w += blockDim.x)
            *reinterpret_cast<uint4*>(timg + SEQ_PAD + 16 * w) = expand16(__ldg(tsrc + w));
        const uint32_t* qsrc = reinterpret_cast<const uint32_t*>(seqs + tk.q_off);
        if (!(tk.flags & SVP_PAIR_REVCOMP)) {
            for (int w = g; w < (tk.m + 15) >> 4; w += blockDim.x)
                *reinterpret_cast<uint4*>(qimg + SEQ_PAD + 16 * w) = expand16(__ldg(qsrc + w));
        } else {
            // oriented base p = complement of original base m-1-p
            for (int w = g; w < (tk.m + 15) >> 4; w += blockDim.x) {
                uint32_t packed = 0;
                #pragma unroll
                for (int b = 0; b < 16; ++b) {
                    const int p = w * 16 + b;
                    if (p < tk.m) {
                        const unsigned o = (unsigned)(tk.m - 1 - p);
                        packed |= (3u - ((__ldg(qsrc + (o >> 4)) >> ((o & 15u) * 2u)) & 3u)) << (2 * b);
                    }
                }
                *reinterpret_cast<uint4*>(qimg + SEQ_PAD + 16 * w) = expand16(packed);
            }
        }
        __syncthreads();
        // NaN pads (after the bulk stores: the last 16-byte group may cover the trailing pad)
        for (int k = g; k < SEQ_PAD; k += blockDim.x) { qimg[k] = 0x7E; timg[k] = 0x7E; }
        for (int k = SEQ_PAD + tk.m + g; k < qbytes; k += blockDim.x) qimg[k] = 0x7E;
        for (int k = SEQ_PAD + tk.n + g; k < tbytes; k += blockDim.x) timg[k] = 0x7E;
        if (MULTI && g <= nwarp) { sts_v4(s_mbA + 16u * g, c_noe1, c_noe2, c_neg, c_neg); sts_v4(s_mbB + 16u * g, c_noe1, c_noe2, c_neg, c_neg); }
        __syncthreads();
    }

    const int n_thr_active = tk.band_w / (4 * KP);
    const bool active = g < n_thr_active;
    const bool lo_edge = (lane == 0);                           // left neighbour: mailbox / constants
    const bool hi_edge = (lane == 31) || (g == n_thr_active - 1);   // up neighbour: mailbox / constants
    const int D0 = tk.band_lo + g * 4 * KP;
    const int I0 = (tk.r_begin - D0) / 2;          // exact: r_begin - D0 is even
    const int J0 = (tk.r_begin + D0) / 2;
    const int qmax = tk.m + 1, tmax = tk.n + 1;

    auto code_at = [&](uint32_t sbase, int pos, int lim) -> uint32_t {   // fp16 code (16 bit) of a 1-based position
        return lds_u8(sbase + (uint32_t)__vimin_s32_relu(pos, lim)) << 8;
    };
    // packed state, slot k in the low half, slot KP+k in the high half
    uint32_t HA[KP], HB[KP], P1[KP], P2[KP], E1[KP], E2[KP], F1[KP], F2[KP], Q[KP], T[KP];
    #pragma unroll
    for (int k = 0; k < KP; ++k) {
        HA[k] = 0; HB[k] = 0; P1[k] = c_noe1; P2[k] = c_noe2;
        E1[k] = c_neg; E2[k] = c_neg; F1[k] = c_neg; F2[k] = c_neg;
        Q[k] = code_at(s_q, I0 - k, qmax) | (code_at(s_q, I0 - (KP + k), qmax) << 16);
        T[k] = code_at(s_t, J0 + k, tmax) | (code_at(s_t, J0 + KP + k, tmax) << 16);
    }
    int qpos = I0 + 1, tpos = J0 + 2 * KP;                   // next bases to enter
    uint32_t best = 0, snap = 0, blk_lo = 0, blk_hi = 0;
    const int pitch = tk.band_w >> 1;                        // bytes per traceback row (one step)
    uint8_t* tbp = tb + tk.tb_off + (size_t)g * (2 * KP);
    const uint32_t s_myA = s_mbA + 16u * warp, s_upA = s_mbA + 16u * (warp + 1);
    const uint32_t s_myB = s_mbB + 16u * (warp + 1), s_loB = s_mbB + 16u * warp;

    #pragma unroll 1
    for (int it = 0; it < tk.n_iter; ++it) {
        #pragma unroll
        for (int u = 0; u < KP; ++u) {
            // =============================== A-step ===========================================
            {
                uint32_t sP1 = __shfl_up_sync(0xffffffffu, P1[KP - 1], 1);
                uint32_t sE1 = __shfl_up_sync(0xffffffffu, E1[KP - 1], 1);
                uint32_t sP2 = 0, sE2 = 0;
                if (TWO) { sP2 = __shfl_up_sync(0xffffffffu, P2[KP - 1], 1); sE2 = __shfl_up_sync(0xffffffffu, E2[KP - 1], 1); }
                uint4 mb = make_uint4(c_noe1, c_noe2, c_neg, c_neg);
                if (MULTI) mb = lds_v4(s_loB);
                sP1 = lo_edge ? mb.x : sP1; sE1 = lo_edge ? mb.z : sE1;
                if (TWO) { sP2 = lo_edge ? mb.y : sP2; sE2 = lo_edge ? mb.w : sE2; }
                const uint32_t P1m = __byte_perm(sP1, P1[KP - 1], 0x5432), E1m = __byte_perm(sE1, E1[KP - 1], 0x5432);
                const uint32_t P2m = TWO ? __byte_perm(sP2, P2[KP - 1], 0x5432) : 0u, E2m = TWO ? __byte_perm(sE2, E2[KP - 1], 0x5432) : 0u;
                #pragma unroll
                for (int k = KP - 1; k >= 0; --k) {
                    const uint32_t p1l = k ? P1[k ? k - 1 : 0] : P1m, e1l = k ? E1[k ? k - 1 : 0] : E1m;
                    const uint32_t e1 = __viaddmax_s16x2(e1l, c_ne1, p1l);
                    const uint32_t f1 = __viaddmax_s16x2(F1[k], c_ne1, P1[k]);
                    const uint32_t mk = heq_mask(Q[(k - u + KP) % KP], T[(k + u) % KP]);
                    const uint32_t h = __vadd2(HA[k], (mk & c_mat) | (~mk & c_mis));
                    uint32_t H;
                    if (TWO) {
                        const uint32_t p2l = k ? P2[k ? k - 1 : 0] : P2m, e2l = k ? E2[k ? k - 1 : 0] : E2m;
                        const uint32_t e2 = __viaddmax_s16x2(e2l, c_ne2, p2l);
                        const uint32_t f2 = __viaddmax_s16x2(F2[k], c_ne2, P2[k]);
                        H = __vimax3_s16x2_relu(__vimax3_s16x2(h, e1, f1), e2, f2);
                        E2[k] = e2; F2[k] = f2; P2[k] = __vadd2(H, c_noe2);
                    } else {
                        H = __vimax3_s16x2_relu(h, e1, f1);
                    }
                    HA[k] = H; E1[k] = e1; F1[k] = f1; P1[k] = __vadd2(H, c_noe1);
                }
                #pragma unroll
                for (int k = 0; k < KP; k += 2) best = __vimax3_s16x2(best, HA[k], HA[k + 1]);
                if (active) {
                    if (KP == 4) __stcs(reinterpret_cast<uint2*>(tbp), make_uint2(__byte_perm(HA[0], HA[1], 0x6420), __byte_perm(HA[2], HA[3], 0x6420)));
                    else __stcs(reinterpret_cast<uint4*>(tbp), make_uint4(__byte_perm(HA[0], HA[1], 0x6420), __byte_perm(HA[2], HA[3], 0x6420),
                                                                          __byte_perm(HA[4 % KP], HA[5 % KP], 0x6420), __byte_perm(HA[6 % KP], HA[7 % KP], 0x6420)));
                }
                tbp += pitch;
                if (MULTI) {
                    if (lane == 0) sts_v4(s_myA, P1[0], P2[0], F1[0], F2[0]);
                    __syncthreads();
                }
                // T advances by one base: logical T[k] <- T[k+1]; the new base enters slot 2KP-1
                const uint32_t nb = lds_u8(s_t + (uint32_t)__vimin_s32_relu(tpos, tmax));
                T[u % KP] = __byte_perm(T[u % KP], nb, 0x4232);
                ++tpos;
            }
            // =============================== B-step ===========================================
            {
                uint32_t sP1 = __shfl_down_sync(0xffffffffu, P1[0], 1);
                uint32_t sF1 = __shfl_down_sync(0xffffffffu, F1[0], 1);
                uint32_t sP2 = 0, sF2 = 0;
                if (TWO) { sP2 = __shfl_down_sync(0xffffffffu, P2[0], 1); sF2 = __shfl_down_sync(0xffffffffu, F2[0], 1); }
                uint4 mb = make_uint4(c_noe1, c_noe2, c_neg, c_neg);
                if (MULTI) mb = lds_v4(s_upA);
                sP1 = hi_edge ? mb.x : sP1; sF1 = hi_edge ? mb.z : sF1;
                if (TWO) { sP2 = hi_edge ? mb.y : sP2; sF2 = hi_edge ? mb.w : sF2; }
                const uint32_t P1p = __byte_perm(P1[0], sP1, 0x5432), F1p = __byte_perm(F1[0], sF1, 0x5432);
                const uint32_t P2p = TWO ? __byte_perm(P2[0], sP2, 0x5432) : 0u, F2p = TWO ? __byte_perm(F2[0], sF2, 0x5432) : 0u;
                #pragma unroll
                for (int k = 0; k < KP; ++k) {
                    const uint32_t p1u = (k < KP - 1) ? P1[(k + 1) % KP] : P1p, f1u = (k < KP - 1) ? F1[(k + 1) % KP] : F1p;
                    const uint32_t e1 = __viaddmax_s16x2(E1[k], c_ne1, P1[k]);
                    const uint32_t f1 = __viaddmax_s16x2(f1u, c_ne1, p1u);
                    const uint32_t mk = heq_mask(Q[(k - u + KP) % KP], T[(k + u + 1) % KP]);
                    const uint32_t h = __vadd2(HB[k], (mk & c_mat) | (~mk & c_mis));
                    uint32_t H;
                    if (TWO) {
                        const uint32_t p2u = (k < KP - 1) ? P2[(k + 1) % KP] : P2p, f2u = (k < KP - 1) ? F2[(k + 1) % KP] : F2p;
                        const uint32_t e2 = __viaddmax_s16x2(E2[k], c_ne2, P2[k]);
                        const uint32_t f2 = __viaddmax_s16x2(f2u, c_ne2, p2u);
                        H = __vimax3_s16x2_relu(__vimax3_s16x2(h, e1, f1), e2, f2);
                        E2[k] = e2; F2[k] = f2; P2[k] = __vadd2(H, c_noe2);
                    } else {
                        H = __vimax3_s16x2_relu(h, e1, f1);
                    }
                    HB[k] = H; E1[k] = e1; F1[k] = f1; P1[k] = __vadd2(H, c_noe1);
                }
                #pragma unroll
                for (int k = 0; k < KP; k += 2) best = __vimax3_s16x2(best, HB[k], HB[k + 1]);
                if (active) {
                    if (KP == 4) __stcs(reinterpret_cast<uint2*>(tbp), make_uint2(__byte_perm(HB[0], HB[1], 0x6420), __byte_perm(HB[2], HB[3], 0x6420)));
                    else __stcs(reinterpret_cast<uint4*>(tbp), make_uint4(__byte_perm(HB[0], HB[1], 0x6420), __byte_perm(HB[2], HB[3], 0x6420),
                                                                          __byte_perm(HB[4 % KP], HB[5 % KP], 0x6420), __byte_perm(HB[6 % KP], HB[7 % KP], 0x6420)));
                }
                tbp += pitch;
                if (MULTI) {
                    if (lane == 31) sts_v4(s_myB, P1[KP - 1], P2[KP - 1], E1[KP - 1], E2[KP - 1]);
                    __syncthreads();
                }
                // Q advances by one base: logical Q[k] <- Q[k-1]; the new base enters slot 0
                const uint32_t nb = lds_u8(s_q + (uint32_t)__vimin_s32_relu(qpos, qmax));
                Q[(KP - 1 - u) % KP] = __byte_perm(nb, Q[(KP - 1 - u) % KP], 0x5401);
                ++qpos;
            }
        }
        // which iteration first reached the running maximum, per half (end-cell search)
        const uint32_t diff = best ^ snap;
        if (diff & 0xFFFFu) blk_lo = (uint32_t)it;
        if (diff >> 16) blk_hi = (uint32_t)it;
        snap = best;
    }
    if (active) {
        bests[tk.best_off + 2 * g]     = make_uint2(best & 0xFFFFu, blk_lo);
        bests[tk.best_off + 2 * g + 1] = make_uint2(best >> 16, blk_hi);
    }
}

// ---------------------------------------------------------------------------------------------
// traceback
// ---------------------------------------------------------------------------------------------
template <int KP>
struct TbView {
    const uint8_t* base;      // arena + tb_off
    const uint32_t* seqs32;   // packed sequences (global)
    int band_lo, band_w, r_begin, pitch, m, n;
    uint64_t q_word, t_word;  // word offsets of the packed sequences
    bool rc;

    // low byte of H(i, j); 0 for virtual cells (i < 1 or j < 1) — callers check the band
    __device__ __forceinline__ uint32_t lb(int i, int j) const {
        if (i < 1 || j < 1) return 0;
        const int x = (j - i) - band_lo, s = (i + j) - r_begin;
        if (s < 0) return 0;
        const int gthr = x / (4 * KP), y = x % (4 * KP), c = y >> 1, k = c % KP, half = c / KP;
        return base[(size_t)s * pitch + gthr * (2 * KP) + (k >> 1) * 4 + (k & 1) * 2 + half];
    }
    __device__ __forceinline__ int qbase(int i) const {       // oriented query base, 1-based
        unsigned p = (unsigned)(i - 1);
        if (rc) { p = (unsigned)(m - i); }
        const uint32_t w = __ldg(seqs32 + q_word + (p >> 4));
        const int c = (w >> ((p & 15u) * 2u)) & 3;
        return rc ? 3 - c : c;
    }
    __device__ __forceinline__ int tbase(int j) const {
        const unsigned p = (unsigned)(j - 1);
        return (__ldg(seqs32 + t_word + (p >> 4)) >> ((p & 15u) * 2u)) & 3;
    }
};

__device__ __forceinline__ int gap_cost2(int o1, int e1, int o2, int e2, int two, int k) {
    int c = o1 + k * e1;
    if (two) c = min(c, o2 + k * e2);
    return c;
}

__device__ __forceinline__ int sext8(uint32_t x) { return (int)(int8_t)(x & 0xFFu); }

template <int KP>
__global__ void __launch_bounds__(128)
svp_tb_kernel(const Task* __restrict__ tasks, int n_tasks, const uint8_t* __restrict__ seqs, const uint8_t* __restrict__ tb,
              const uint2* __restrict__ bests, svp_result* __restrict__ results, uint32_t* __restrict__ cigar, const Consts cs) {
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (wid >= n_tasks) return;
    const Task tk = tasks[wid];
    if (((tk.flags & TASK_KP8) != 0) != (KP == 8)) return;     // the other instantiation handles this pair
    if (tk.flags & TASK_INVALID) {
        if (lane == 0) {
            svp_result bad;
            memset(&bad, 0, sizeof(bad));
            bad.status = (tk.flags & TASK_TOO_LONG) ? SVP_ST_SCORE_OVF : SVP_ST_BAD_PAIR;
            bad.cigar_off = tk.cigar_off;
            results[tk.pair_index] = bad;
        }
        return;
    }
    TbView<KP> v;
    v.base = tb + tk.tb_off; v.seqs32 = reinterpret_cast<const uint32_t*>(seqs);
    v.band_lo = tk.band_lo; v.band_w = tk.band_w; v.r_begin = tk.r_begin; v.pitch = tk.band_w >> 1; v.m = tk.m; v.n = tk.n;
    v.q_word = tk.q_off >> 2; v.t_word = tk.t_off >> 2; v.rc = (tk.flags & SVP_PAIR_REVCOMP) != 0;
    const int d_hi = tk.band_lo + tk.band_w - 1;

    svp_result res;
    res.score = 0; res.status = 0; res.q_beg = res.q_end = res.t_beg = res.t_end = 0;
    res.n_match = res.n_mismatch = res.n_ins_ops = res.n_ins_bp = res.n_del_ops = res.n_del_bp = 0;
    res.cigar_off = tk.cigar_off; res.cigar_len = 0; res.sv_dist = 0.0; res.cells = tk.cells;

    // ---- 1. end cell: maximum over the per-thread-half records, earliest iteration ------------
    const int n_ent = (tk.band_w / (4 * KP)) * 2;
    int M = 0; uint32_t blk = 0xFFFFFFFFu;
    for (int e = lane; e < n_ent; e += 32) {
        const uint2 b = bests[tk.best_off + e];
        const int val = (int)(int16_t)b.x;
        if (val > M || (val == M && b.y < blk)) { M = val; blk = b.y; }
    }
    #pragma unroll
    for (int o = 16; o; o >>= 1) {
        const int M2 = __shfl_xor_sync(0xffffffffu, M, o);
        const uint32_t b2 = __shfl_xor_sync(0xffffffffu, blk, o);
        if (M2 > M || (M2 == M && b2 < blk)) { M = M2; blk = b2; }
    }
    if (M <= 0) {
        res.status = SVP_ST_NOHIT;
        if (lane == 0) results[tk.pair_index] = res;
        return;
    }
    if (M >= 32767 - cs.s_match) res.status |= SVP_ST_SCORE_OVF;
    // all (thread, half) regions that reached M in iteration blk: rebuild the 2KP x 2KP block
    int end_s = 0x7FFFFFFF, end_x = -1;
    for (int e0 = 0; e0 < n_ent; e0 += 32) {
        const int e = e0 + lane;
        bool hit = false;
        if (e < n_ent) { const uint2 b = bests[tk.best_off + e]; hit = ((int)(int16_t)b.x == M) && (b.y == blk); }
        unsigned mask = __ballot_sync(0xffffffffu, hit);
        while (mask) {
            const int src = __ffs(mask) - 1; mask &= mask - 1;
            const int ee = e0 + src;
            const int xbase = (ee >> 1) * 4 * KP + (ee & 1) * 2 * KP;     // first diagonal (band-relative) of the half
            const int s0 = (int)blk * 2 * KP;
            // cells idx = 0 .. 2*KP*KP-1: step s0 + idx/KP, diagonal xbase + 2*(idx%KP) + (step odd)
            constexpr int NC = 2 * KP * KP;
            int rel_prev_row0 = 0; uint32_t b_prev_row0 = 0;
            int rel = 0; uint32_t bprev = 0;
            int relmax = -(1 << 30);
            int cand_s = 0x7FFFFFFF, cand_x = -1;
            // sequential chain, executed uniformly by all lanes (bytes fetched 32 at a time)
            for (int c0 = 0; c0 < NC; c0 += 32) {
                const int idx = c0 + lane;
                uint32_t mine = 0;
                if (idx < NC) {
                    const int s = s0 + idx / KP, x = xbase + 2 * (idx % KP) + (s & 1);
                    // stored for every computed cell, virtual ones included: read the raw byte
                    const int gthr = x / (4 * KP), y = x % (4 * KP), c = y >> 1, k = c % KP, half = c / KP;
                    mine = v.base[(size_t)s * v.pitch + gthr * (2 * KP) + (k >> 1) * 4 + (k & 1) * 2 + half];
                }
                const int lim = min(32, NC - c0);
                for (int l = 0; l < lim; ++l) {
                    const uint32_t b = __shfl_sync(0xffffffffu, mine, l);
                    const int id = c0 + l;
                    if (id == 0) { rel = 0; }
                    else if (id % KP == 0) { rel = rel_prev_row0 + sext8(b - b_prev_row0); }
                    else { rel = rel + sext8(b - bprev); }
                    if (id % KP == 0) { rel_prev_row0 = rel; b_prev_row0 = b; }
                    bprev = b;
                    const int s = s0 + id / KP, x = xbase + 2 * (id % KP) + (s & 1);
                    if (rel > relmax) { relmax = rel; cand_s = s; cand_x = x; }
                    else if (rel == relmax && (s < cand_s || (s == cand_s && x > cand_x))) { cand_s = s; cand_x = x; }
                }
            }
            if (cand_s < end_s || (cand_s == end_s && cand_x > end_x)) { end_s = cand_s; end_x = cand_x; }
        }
    }
    int i, j;
    {
        const int r = tk.r_begin + end_s, d = tk.band_lo + end_x;
        i = (r - d) / 2; j = (r + d) / 2;
    }
    const int best_i = i, best_j = j;

    // ---- 2. walk back --------------------------------------------------------------------------
    int H = M;
    uint32_t* slot = cigar + tk.cigar_off;
    int wpos = (int)tk.cigar_cap;                    // runs are written right-aligned, backwards
    uint32_t cur_op = 0, cur_len = 0; bool ovf = false;
    uint32_t n_match = 0, n_mis = 0, n_io = 0, n_ib = 0, n_do = 0, n_db = 0;
    // snapshot at the lowest H of the walk (Z-trim, §3.5)
    int low = M + 1, low_i = i, low_j = j, low_wpos = wpos; uint32_t low_op = 0, low_len = 0; bool low_ovf = false;
    uint32_t l_match = 0, l_mis = 0, l_io = 0, l_ib = 0, l_do = 0, l_db = 0;

#define SVP_PUSH() do { if (cur_len) { if (wpos > 0) { --wpos; if (lane == 0) slot[wpos] = (cur_len << 4) | cur_op; } else ovf = true; } } while (0)
#define SVP_EMIT(op, len) do { if (cur_len && cur_op == (op)) cur_len += (len); else { SVP_PUSH(); cur_op = (op); cur_len = (len); } } while (0)
#define SVP_SNAP(hv) do { low = (hv); low_i = i; low_j = j; low_wpos = wpos; low_op = cur_op; low_len = cur_len; low_ovf = ovf; \
        l_match = n_match; l_mis = n_mis; l_io = n_io; l_ib = n_ib; l_do = n_do; l_db = n_db; } while (0)

    bool done = false;
    while (!done) {
        // ---- diagonal run: lane l looks at the cell (i-l, j-l) and its diagonal predecessor ----
        const int ci = i - lane, cj = j - lane;
        uint32_t bp = 0; int qc = 4, tc = 5;
        if (ci >= 1 && cj >= 1) {
            bp = v.lb(ci - 1, cj - 1);
            qc = v.qbase(ci); tc = v.tbase(cj);
        }
        int l = 0;
        bool gap = false;
        for (; l < 32; ++l) {
            // current cell (i, j) with exact H
            if (H < low) SVP_SNAP(H);
            if (H == 0 || i < 1 || j < 1) { done = true; break; }
            if (cs.zdrop > 0) {
                int dd = (j - i) - (low_j - low_i); dd = dd < 0 ? -dd : dd;
                if (H - low > cs.zdrop + cs.zdrop_ext * dd) { done = true; break; }
            }
            const uint32_t b = __shfl_sync(0xffffffffu, bp, l);
            const int q1 = __shfl_sync(0xffffffffu, qc, l), t1 = __shfl_sync(0xffffffffu, tc, l);
            const int Hd = (i - 1 >= 1 && j - 1 >= 1) ? H + sext8(b - (uint32_t)H) : 0;
            const int sc = (q1 == t1) ? cs.s_match : cs.s_mismatch;
            if (H == Hd + sc) {
                SVP_EMIT(SVP_CIGAR_M, 1u);
                if (q1 == t1) ++n_match; else ++n_mis;
                --i; --j; H = Hd;
            } else { gap = true; break; }
        }
        if (done || !gap) continue;
        // ---- gap: find the smallest k with H == H(i, j-k) - del(k), else H == H(i-k, j) - ins(k) ----
        const int d = j - i;
        int Hrow = H, Hcol = H; uint32_t brow = (uint32_t)H & 0xFFu, bcol = (uint32_t)H & 0xFFu;
        bool found = false;
        for (int k0 = 0; !found; k0 += 32) {
            const int k = k0 + lane + 1;
            const bool del_ok_l = (j - k >= 1) && (d - k >= tk.band_lo);
            const bool ins_ok_l = (i - k >= 1) && (d + k <= d_hi);
            const uint32_t bd = del_ok_l ? v.lb(i, j - k) : 0;
            const uint32_t bi = ins_ok_l ? v.lb(i - k, j) : 0;
            bool any = false;
            for (int l2 = 0; l2 < 32; ++l2) {
                const int kk = k0 + l2 + 1;
                const bool del_ok = (j - kk >= 1) && (d - kk >= tk.band_lo);
                const bool ins_ok = (i - kk >= 1) && (d + kk <= d_hi);
                if (!del_ok && !ins_ok) break;
                any = true;
                if (del_ok) {
                    const uint32_t b = __shfl_sync(0xffffffffu, bd, l2);
                    Hrow += sext8(b - brow); brow = b;
                    if (H == Hrow - gap_cost2(cs.open1, cs.ext1, cs.open2, cs.ext2, cs.two_piece, kk)) {
                        SVP_EMIT(SVP_CIGAR_D, (uint32_t)kk); ++n_do; n_db += kk; j -= kk; H = Hrow; found = true; break;
                    }
                }
                if (ins_ok) {
                    const uint32_t b = __shfl_sync(0xffffffffu, bi, l2);
                    Hcol += sext8(b - bcol); bcol = b;
                    if (H == Hcol - gap_cost2(cs.open1, cs.ext1, cs.open2, cs.ext2, cs.two_piece, kk)) {
                        SVP_EMIT(SVP_CIGAR_I, (uint32_t)kk); ++n_io; n_ib += kk; i -= kk; H = Hcol; found = true; break;
                    }
                }
            }
            if (!found && !any) break;
            if (!found) {
                const int kk_next = k0 + 33;
                if (!((j - kk_next >= 1) && (d - kk_next >= tk.band_lo)) && !((i - kk_next >= 1) && (d + kk_next <= d_hi))) break;
            }
        }
        if (!found) { res.status |= SVP_ST_SCORE_OVF; done = true; }
    }
    // rewind to the snapshot
    i = low_i; j = low_j; wpos = low_wpos; cur_op = low_op; cur_len = low_len; ovf = low_ovf;
    SVP_PUSH();
    res.score = M - low;
    res.q_beg = i; res.q_end = best_i; res.t_beg = j; res.t_end = best_j;
    res.n_match = l_match; res.n_mismatch = l_mis; res.n_ins_ops = l_io; res.n_ins_bp = l_ib; res.n_del_ops = l_do; res.n_del_bp = l_db;
#undef SVP_PUSH
#undef SVP_EMIT
#undef SVP_SNAP
    __syncwarp();
    if (lane == 0) {
        if (ovf) {
            res.status |= SVP_ST_CIGAR_OVF; res.cigar_len = 0;
        } else {
            res.cigar_len = tk.cigar_cap - (uint32_t)wpos;
            res.cigar_off = tk.cigar_off + (uint32_t)wpos;
            double dist = 0.0;
            for (uint32_t a = (uint32_t)wpos; a < tk.cigar_cap; ++a) {
                const uint32_t w = slot[a];
                const uint32_t op = w & 15u, len = w >> 4;
                if (op != SVP_CIGAR_M && (int)len >= cs.min_signal) {
                    const double s = (double)len;
                    dist += (1.0 - exp(-pow(s / 30.0, 1.5))) * s;
                }
            }
            res.sv_dist = dist;
        }
        results[tk.pair_index] = res;
    }
}

}  // namespace svp
