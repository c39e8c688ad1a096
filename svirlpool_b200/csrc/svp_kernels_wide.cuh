// svp_kernels_wide.cuh — the forward-DP variants beside the fast path of svp_kernels.cuh. Same decomposition
// (one CTA per pair, a thread owns consecutive diagonals, even ones = set A, odd ones = set B, one anti-diagonal
// per step, neighbours by shuffle / mailbox) and the same traceback rows, so the traceback kernels serve all of them.
//
//   svp_fwdg_kernel<MULTI, TWO>       s16x2 lanes, KP = 4 geometry, GENERAL scoring: a full 4x4 substitution
//                                     matrix (the shape of a last-train / lamassemble .mat, consensus.py:777) and
//                                     separate insertion / deletion gap costs.
//   svp_fwd32_kernel<MULTI, TWO, GEN> s32 lanes (one cell per register), KP = 4 geometry, for pairs whose score
//                                     can leave the s16 range (match * min(m, n) > 32000: reads beyond ~16 kb);
//                                     GEN selects general scoring.
//
// General scoring, substitution score of a cell = one byte out of the 16-byte matrix held in four registers
// (tab[q] = row of query base q, byte t). PRMT picks one of 8 bytes per selector nibble, so two first-level PRMTs
// look the 3-bit index (q & 1) * 4 + t up in rows {0,1} and rows {2,3}, and a third PRMT picks between the two by
// (q >> 1) while replicating the sign into the upper byte of the 16-bit lane. The third selector depends on the
// query base only and is kept per query register (Q2), updated when the query shifts. A selector nibble with bit 3
// set makes PRMT return the sign (0x00 / 0xFF) of the byte instead: the pad code of positions outside a sequence,
// so virtual cells score 0 or -1 — never positive, which is all the boundary logic needs.
#pragma once
#include "svp_kernels.cuh"

namespace svp {

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
    return r;
}

// 16 bases of one 2-bit packed word -> 16 raw codes (0..3), one per byte
__device__ __forceinline__ uint4 expand16_codes(uint32_t w) {
    uint32_t o[4];
    #pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t x = (w >> (8 * k)) & 0xFFu;
        o[k] = (x | (x << 6) | (x << 12) | (x << 18)) & 0x03030303u;
    }
    return make_uint4(o[0], o[1], o[2], o[3]);
}

// Shared-memory images for general scoring, one byte per base, base p at byte SEQ_PAD + p:
//   qsel  = (q & 1) * 4            pad 0x08      first-level selector contribution of the query
//   qsel2 = 0x80 | 0x44 * (q >> 1) pad 0x80      second-level selector byte of the low lane (the high lane adds 0x22)
//   tsel  = t                      pad 0x08
constexpr uint32_t GEN_PAD = 0x08u, GEN_PAD2 = 0x80u;
__device__ __forceinline__ void stage_sequences_gen(const Task& tk, const uint8_t* __restrict__ seqs, uint8_t* qsel, uint8_t* qsel2,
                                                    uint8_t* tsel, int qbytes, int tbytes) {
    const int g = threadIdx.x;
    const uint32_t* tsrc = reinterpret_cast<const uint32_t*>(seqs + tk.t_off);
    for (int w = g; w < (tk.n + 15) >> 4; w += blockDim.x)
        *reinterpret_cast<uint4*>(tsel + SEQ_PAD + 16 * w) = expand16_codes(__ldg(tsrc + w));
    const uint32_t* qsrc = reinterpret_cast<const uint32_t*>(seqs + tk.q_off);
    for (int w = g; w < (tk.m + 15) >> 4; w += blockDim.x) {
        uint32_t packed = 0;
        if (!(tk.flags & SVP_PAIR_REVCOMP)) {
            packed = __ldg(qsrc + w);
        } else {
            #pragma unroll
            for (int b = 0; b < 16; ++b) {
                const int p = w * 16 + b;
                if (p < tk.m) {
                    const unsigned o = (unsigned)(tk.m - 1 - p);
                    packed |= (3u - ((__ldg(qsrc + (o >> 4)) >> ((o & 15u) * 2u)) & 3u)) << (2 * b);
                }
            }
        }
        const uint4 c = expand16_codes(packed);
        const uint32_t cc[4] = { c.x, c.y, c.z, c.w };
        uint32_t s1[4], s2[4];
        #pragma unroll
        for (int k = 0; k < 4; ++k) {
            s1[k] = (cc[k] & 0x01010101u) << 2;
            s2[k] = 0x80808080u | (((cc[k] >> 1) & 0x01010101u) * 0x44u);
        }
        *reinterpret_cast<uint4*>(qsel + SEQ_PAD + 16 * w) = make_uint4(s1[0], s1[1], s1[2], s1[3]);
        *reinterpret_cast<uint4*>(qsel2 + SEQ_PAD + 16 * w) = make_uint4(s2[0], s2[1], s2[2], s2[3]);
    }
    __syncthreads();
    // letters outside ACGT: the pad codes (score(): a pad on either side scores the lowest matrix entry)
    if (tk.flags & TASK_TN) mask_to_image(seqs + tk.t_off, tk.n, tsel, (uint8_t)GEN_PAD, false);
    if (tk.flags & TASK_QN) {
        mask_to_image(seqs + tk.q_off, tk.m, qsel, (uint8_t)GEN_PAD, (tk.flags & SVP_PAIR_REVCOMP) != 0);
        mask_to_image(seqs + tk.q_off, tk.m, qsel2, (uint8_t)GEN_PAD2, (tk.flags & SVP_PAIR_REVCOMP) != 0);
    }
    for (int k = g; k < SEQ_PAD; k += blockDim.x) { qsel[k] = GEN_PAD; qsel2[k] = GEN_PAD2; tsel[k] = GEN_PAD; }
    for (int k = SEQ_PAD + tk.m + g; k < qbytes; k += blockDim.x) { qsel[k] = GEN_PAD; qsel2[k] = GEN_PAD2; }
    for (int k = SEQ_PAD + tk.n + g; k < tbytes; k += blockDim.x) tsel[k] = GEN_PAD;
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// forward DP, s16x2 lanes, general scoring (KP = 4 geometry)
// ---------------------------------------------------------------------------------------------
template <bool MULTI, bool TWO, bool CL = false, bool WALK = true>
__global__ void __launch_bounds__(MULTI ? 512 : 32)
svp_fwdg_kernel(const Task* __restrict__ tasks, const uint32_t* __restrict__ order, const uint8_t* __restrict__ seqs,
                const Arena ar, svp_result* __restrict__ results, uint32_t* __restrict__ cigar, const Consts cs) {
    static_assert(!CL || MULTI, "cluster variants are multi-warp");
    constexpr int KP = 4;
    extern __shared__ __align__(16) uint8_t smem8[];
    __shared__ unsigned long long s_slice[3];
    const uint32_t crank = CL ? cluster_ctarank() : 0u, csize = CL ? cluster_nctarank() : 1u;
    const uint32_t task_index = order[CL ? blockIdx.x / csize : blockIdx.x];
    const Task tk = tasks[task_index];
    const int gl = threadIdx.x, lane = gl & 31, warp = gl >> 5, nwarp = blockDim.x >> 5;
    const int g = (int)crank * (int)blockDim.x + gl;
    const int qbytes = seq_image_bytes(tk.m), tbytes = seq_image_bytes(tk.n);
    uint8_t* qsel = smem8;
    uint8_t* qsel2 = smem8 + qbytes;
    uint8_t* tsel = smem8 + 2 * qbytes;
    const uint32_t s_q = (uint32_t)__cvta_generic_to_shared(qsel) + SEQ_PAD - 1;    // + clamped 1-based position
    const uint32_t s_q2 = (uint32_t)__cvta_generic_to_shared(qsel2) + SEQ_PAD - 1;
    const uint32_t s_t = (uint32_t)__cvta_generic_to_shared(tsel) + SEQ_PAD - 1;
    // mailboxes (MULTI): A[w] = lane 0 of warp w after an A-step (H, -, F1, F2), A[nwarp] = constants;
    //                    B[w+1] = lane 31 of warp w after a B-step (H, -, E1, E2), B[0] = constants
    const uint32_t s_mb = (uint32_t)__cvta_generic_to_shared(smem8 + 2 * qbytes + tbytes);

    const uint32_t d_ne1 = pin(cs.d_ne1), d_noe1 = pin(cs.d_noe1), d_ne2 = pin(cs.d_ne2), d_noe2 = pin(cs.d_noe2);
    const uint32_t i_ne1 = pin(cs.i_ne1), i_noe1 = pin(cs.i_noe1), i_ne2 = pin(cs.i_ne2), i_noe2 = pin(cs.i_noe2);
    const uint32_t r0 = pin(cs.tab[0]), r1 = pin(cs.tab[1]), r2 = pin(cs.tab[2]), r3 = pin(cs.tab[3]);
    const uint32_t c_neg = pin(NEGV);

    stage_sequences_gen(tk, seqs, qsel, qsel2, tsel, qbytes, tbytes);
    if (MULTI && gl <= nwarp) { sts_v4(s_mb + MB_A + 16u * gl, 0u, 0u, c_neg, c_neg); sts_v4(s_mb + MB_B + 16u * gl, 0u, 0u, c_neg, c_neg);
                                step_link_init<CL>(s_mb, gl); }
    if (MULTI && gl == 0) sts_v4(s_mb + MB_CST, 0u, 0u, c_neg, c_neg);
    pair_begin<WALK>(ar, tk, crank, s_slice, seqs, results, cigar, cs);
    step_barrier<CL>();
    uint8_t* const rows = pair_rows<CL>(ar, crank, s_slice);
    uint2* const bests = reinterpret_cast<uint2*>(rows + tk.tb_bytes);

    const int n_thr_active = tk.band_w / (4 * KP);
    const bool active = g < n_thr_active;
    const bool lo_edge = (lane == 0);
    const bool hi_edge = (lane == 31) || (g == n_thr_active - 1);
    const int D0 = tk.band_lo + g * 4 * KP;
    const int I0 = (tk.r_begin - D0) / 2;
    const int J0 = (tk.r_begin + D0) / 2;
    const int qmax = tk.m + 1, tmax = tk.n + 1;
    auto at = [&](uint32_t sbase, int pos, int lim) -> uint32_t { return lds_u8(sbase + (uint32_t)__vimin_s32_relu(pos, lim)); };

    // packed state: slot k in the low half, slot KP+k in the high half; selector registers: low lane in byte 0, high lane in byte 1
    uint32_t HA[KP], HB[KP], E1[KP], E2[KP], F1[KP], F2[KP], Q[KP], Q2[KP], T[KP];
    #pragma unroll
    for (int k = 0; k < KP; ++k) {
        HA[k] = 0; HB[k] = 0; E1[k] = c_neg; E2[k] = c_neg; F1[k] = c_neg; F2[k] = c_neg;
        Q[k] = at(s_q, I0 - k, qmax) | (at(s_q, I0 - (KP + k), qmax) << 8);
        Q2[k] = at(s_q2, I0 - k, qmax) | ((at(s_q2, I0 - (KP + k), qmax) | 0x22u) << 8);
        T[k] = at(s_t, J0 + k, tmax) | (at(s_t, J0 + KP + k, tmax) << 8);
    }
    int qpos = I0 + 1, tpos = J0 + 2 * KP;
    uint32_t best = 0, snap = 0, blk_lo = 0, blk_hi = 0;
    const int pitch = n_thr_active * tb_slot_bytes(KP);      // bytes per traceback row (one step pair)
    uint8_t* tbp = rows + (size_t)g * tb_slot_bytes(KP);
    const StepLink link = step_link<CL>(s_mb, warp, nwarp, crank, csize);
    // the last in-band thread reads the "outside the band" constants, whatever warp / CTA follows it
    const uint32_t s_upA = (g == n_thr_active - 1) ? s_mb + MB_CST : link.in_up(), s_loB = link.in_lo();

    // a pad code (bit 3 of a selector byte: a position outside the sequence, or a letter outside ACGT) on either side -> the lowest
    // matrix entry: the byte's bit 3 becomes a 16-bit lane mask (SHL + PRMT in sign mode) and one LOP3 blends the constant in
    const uint32_t c_npk = pin(cs.n_pk);
    auto score = [&](uint32_t q, uint32_t q2, uint32_t t) -> uint32_t {
        const uint32_t sel = q | t;
        const uint32_t look = prmt(prmt(r0, r1, sel), prmt(r2, r3, sel), q2);
        const uint32_t pad = prmt(sel << 4, 0u, 0x9988u);
        return (look & ~pad) | (c_npk & pad);
    };

    #pragma unroll 1
    for (int it = 0; it < tk.n_iter; ++it) {
        #pragma unroll
        for (int u = 0; u < KP; ++u) {
            // =============================== A-step ===========================================
            {
                uint32_t sH = __shfl_up_sync(0xffffffffu, HB[KP - 1], 1);
                uint32_t sE1 = __shfl_up_sync(0xffffffffu, E1[KP - 1], 1);
                uint32_t sE2 = 0;
                if (TWO) sE2 = __shfl_up_sync(0xffffffffu, E2[KP - 1], 1);
                const uint32_t oE1 = E1[KP - 1], oE2 = E2[KP - 1];      // HB is not written in an A-step
                uint32_t Hm = 0, E1m = 0, E2m = 0;
                #pragma unroll
                for (int k = KP - 1; k >= 0; --k) {
                    if (k == 0) {              // slot 0, computed last, alone needs the lower neighbour: wait as late as possible
                        uint4 mb = make_uint4(0u, 0u, c_neg, c_neg);
                        if (MULTI) { if (it | u) link_wait_lo<CL>(link, lane, (uint32_t)((u & 1) ^ 1)); mb = lds_v4(s_loB); }
                        sH = lo_edge ? mb.x : sH; sE1 = lo_edge ? mb.z : sE1;
                        if (TWO) sE2 = lo_edge ? mb.w : sE2;
                        Hm = __byte_perm(sH, HB[KP - 1], 0x5432); E1m = __byte_perm(sE1, oE1, 0x5432);
                        if (TWO) E2m = __byte_perm(sE2, oE2, 0x5432);
                    }
                    const uint32_t hl = k ? HB[k ? k - 1 : 0] : Hm, e1l = k ? E1[k ? k - 1 : 0] : E1m;
                    const uint32_t e1 = __viaddmax_s16x2(e1l, d_ne1, __vadd2(hl, d_noe1));
                    const uint32_t f1 = __viaddmax_s16x2(F1[k], i_ne1, __vadd2(HB[k], i_noe1));
                    const uint32_t h = __vadd2(HA[k], score(Q[(k - u + KP) % KP], Q2[(k - u + KP) % KP], T[(k + u) % KP]));
                    uint32_t H;
                    if (TWO) {
                        const uint32_t e2l = k ? E2[k ? k - 1 : 0] : E2m;
                        const uint32_t e2 = __viaddmax_s16x2(e2l, d_ne2, __vadd2(hl, d_noe2));
                        const uint32_t f2 = __viaddmax_s16x2(F2[k], i_ne2, __vadd2(HB[k], i_noe2));
                        H = __vimax3_s16x2_relu(__vimax3_s16x2(h, e1, f1), e2, f2);
                        E2[k] = e2; F2[k] = f2;
                    } else {
                        H = __vimax3_s16x2_relu(h, e1, f1);
                    }
                    HA[k] = H; E1[k] = e1; F1[k] = f1;
                }
                #pragma unroll
                for (int k = 0; k < KP; k += 2) best = __vimax3_s16x2(best, HA[k], HA[k + 1]);
                if (MULTI && lane == 0) mailbox_store_A<CL>(link, HA[0], 0u, F1[0], F2[0]);
                // T advances by one base: logical T[k] <- T[k+1]; the new base enters slot 2KP-1
                const uint32_t nb = lds_u8(s_t + (uint32_t)__vimin_s32_relu(tpos, tmax));
                T[u % KP] = __byte_perm(T[u % KP], nb, 0x5541);
                ++tpos;
            }
            // =============================== B-step ===========================================
            {
                uint32_t sH = __shfl_down_sync(0xffffffffu, HA[0], 1);
                uint32_t sF1 = __shfl_down_sync(0xffffffffu, F1[0], 1);
                uint32_t sF2 = 0;
                if (TWO) sF2 = __shfl_down_sync(0xffffffffu, F2[0], 1);
                const uint32_t oF1 = F1[0], oF2 = F2[0];                // HA is not written in a B-step
                uint32_t Hp = 0, F1p = 0, F2p = 0;
                #pragma unroll
                for (int k = 0; k < KP; ++k) {
                    if (k == KP - 1) {         // the last slot alone needs the upper neighbour
                        uint4 mb = make_uint4(0u, 0u, c_neg, c_neg);
                        if (MULTI) { link_wait_up<CL>(link, lane, (uint32_t)(u & 1)); mb = lds_v4(s_upA); }
                        sH = hi_edge ? mb.x : sH; sF1 = hi_edge ? mb.z : sF1;
                        if (TWO) sF2 = hi_edge ? mb.w : sF2;
                        Hp = __byte_perm(HA[0], sH, 0x5432); F1p = __byte_perm(oF1, sF1, 0x5432);
                        if (TWO) F2p = __byte_perm(oF2, sF2, 0x5432);
                    }
                    const uint32_t hu = (k < KP - 1) ? HA[(k + 1) % KP] : Hp, f1u = (k < KP - 1) ? F1[(k + 1) % KP] : F1p;
                    const uint32_t e1 = __viaddmax_s16x2(E1[k], d_ne1, __vadd2(HA[k], d_noe1));
                    const uint32_t f1 = __viaddmax_s16x2(f1u, i_ne1, __vadd2(hu, i_noe1));
                    const uint32_t h = __vadd2(HB[k], score(Q[(k - u + KP) % KP], Q2[(k - u + KP) % KP], T[(k + u + 1) % KP]));
                    uint32_t H;
                    if (TWO) {
                        const uint32_t f2u = (k < KP - 1) ? F2[(k + 1) % KP] : F2p;
                        const uint32_t e2 = __viaddmax_s16x2(E2[k], d_ne2, __vadd2(HA[k], d_noe2));
                        const uint32_t f2 = __viaddmax_s16x2(f2u, i_ne2, __vadd2(hu, i_noe2));
                        H = __vimax3_s16x2_relu(__vimax3_s16x2(h, e1, f1), e2, f2);
                        E2[k] = e2; F2[k] = f2;
                    } else {
                        H = __vimax3_s16x2_relu(h, e1, f1);
                    }
                    HB[k] = H; E1[k] = e1; F1[k] = f1;
                }
                #pragma unroll
                for (int k = 0; k < KP; k += 2) best = __vimax3_s16x2(best, HB[k], HB[k + 1]);
                if (MULTI && lane == 31) mailbox_store_B<CL>(link, HB[KP - 1], 0u, E1[KP - 1], E2[KP - 1]);
                if (active) store_pair<KP>(tbp, HA, HB);
                tbp += pitch;
                // Q advances by one base: logical Q[k] <- Q[k-1]; the new base enters slot 0
                const uint32_t off = (uint32_t)__vimin_s32_relu(qpos, qmax);
                const uint32_t nb = lds_u8(s_q + off), nb2 = lds_u8(s_q2 + off);
                Q[(KP - 1 - u) % KP] = __byte_perm(nb, Q[(KP - 1 - u) % KP], 0x1140);
                Q2[(KP - 1 - u) % KP] = __byte_perm(nb2, Q2[(KP - 1 - u) % KP], 0x1140) | 0x2200u;
                ++qpos;
            }
        }
        const uint32_t diff = best ^ snap;
        if (diff & 0xFFFFu) blk_lo = (uint32_t)it;
        if (diff >> 16) blk_hi = (uint32_t)it;
        snap = best;
    }
    if (active) {
        bests[2 * g]     = make_uint2((uint32_t)(int)(int16_t)(best & 0xFFFFu), blk_lo);
        bests[2 * g + 1] = make_uint2((uint32_t)(int)(int16_t)(best >> 16), blk_hi);
    }
    step_barrier<CL>();                        // rows and best records complete; no CTA leaves while a neighbour may still store into its mailbox
    if (WALK && warp == 0 && crank == 0) pair_end(ar, tk, task_index, seqs, results, cigar, cs, s_slice, ar.cor_smem ? (uint32_t)__cvta_generic_to_shared(smem8) : 0u);
}

// ---------------------------------------------------------------------------------------------
// Sliding sequence windows (RING): pairs whose two sequences do not fit the shared-memory images (m + n beyond ~226 k
// bases: a 185 kb padded consensus against its reference window) keep only the part of each sequence the CTA's
// anti-diagonal currently touches. At any time the bases ENTERING the registers of a CTA lie within blockDim * K
// consecutive positions of each sequence, advancing by K positions per loop iteration, so a power-of-two ring of
// R >= blockDim*K + 4K + RING_CHUNK bytes per image, refilled with the next RING_CHUNK positions every
// RING_CHUNK / K iterations, always holds them. The ring is indexed by the VIRTUAL position (pos & (R-1), two's
// complement), positions outside [1, len] hold the pad code, so no clamping is needed.
// ---------------------------------------------------------------------------------------------
// (RING_CHUNK, ring_bytes, packed_code: svp_kernels.cuh)
template <bool GEN>
__device__ __forceinline__ void ring_fill(const Task& tk, const uint8_t* __restrict__ seqs, uint8_t* qring, uint8_t* q2ring, uint8_t* tring,
                                          int mask, int q_first, int t_first, int count) {
    const uint32_t* qsrc = reinterpret_cast<const uint32_t*>(seqs + tk.q_off);
    const uint32_t* tsrc = reinterpret_cast<const uint32_t*>(seqs + tk.t_off);
    const bool rc = (tk.flags & SVP_PAIR_REVCOMP) != 0;
    const uint32_t* qm = seq_mask(seqs, tk.q_off, tk.m, (tk.flags & TASK_QN) != 0);
    const uint32_t* tm = seq_mask(seqs, tk.t_off, tk.n, (tk.flags & TASK_TN) != 0);
    for (int x = threadIdx.x; x < count; x += blockDim.x) {
        const int qc = packed_code(qsrc, tk.m, q_first + x, rc, qm), tc = packed_code(tsrc, tk.n, t_first + x, false, tm);
        if (GEN) {
            qring[(q_first + x) & mask] = (uint8_t)(qc < 0 ? GEN_PAD : (uint32_t)(qc & 1) << 2);
            q2ring[(q_first + x) & mask] = (uint8_t)(qc < 0 ? GEN_PAD2 : 0x80u | 0x44u * (uint32_t)(qc >> 1));
            tring[(t_first + x) & mask] = (uint8_t)(tc < 0 ? GEN_PAD : (uint32_t)tc);
        } else {
            qring[(q_first + x) & mask] = (uint8_t)(qc < 0 ? 0x7E : 0x3C + qc);
            tring[(t_first + x) & mask] = (uint8_t)(tc < 0 ? 0x7D : 0x3C + tc);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// forward DP, 32-bit lanes: one cell per register, K = 8 registers per state array per set; the thread owns
// 2K = 16 diagonals = the geometry of the s16 kernels with KP = 4, and the row bytes are written in their byte
// order. One loop iteration is 2K = 16 steps = two end-cell blocks of 2*KP = 8 steps.
// ---------------------------------------------------------------------------------------------
template <bool MULTI, bool TWO, bool GEN, bool CL = false, bool RING = false, bool WALK = true>
__global__ void __launch_bounds__(MULTI ? 512 : 32)
svp_fwd32_kernel(const Task* __restrict__ tasks, const uint32_t* __restrict__ order, const uint8_t* __restrict__ seqs,
                 const Arena ar, svp_result* __restrict__ results, uint32_t* __restrict__ cigar, const Consts cs) {
    static_assert(!CL || MULTI, "cluster variants are multi-warp");
    constexpr int K = 8, KPG = K / 2;
    extern __shared__ __align__(16) uint8_t smem8[];
    __shared__ unsigned long long s_slice[3];
    const uint32_t crank = CL ? cluster_ctarank() : 0u, csize = CL ? cluster_nctarank() : 1u;
    const uint32_t task_index = order[CL ? blockIdx.x / csize : blockIdx.x];
    const Task tk = tasks[task_index];
    const int gl = threadIdx.x, lane = gl & 31, warp = gl >> 5, nwarp = blockDim.x >> 5;
    const int g = (int)crank * (int)blockDim.x + gl;
    const int ring = ring_bytes((int)blockDim.x, K), rmask = ring - 1;
    const int qbytes = RING ? ring : seq_image_bytes(tk.m), tbytes = RING ? ring : seq_image_bytes(tk.n);
    uint8_t* qimg = smem8;                                   // GEN: qsel
    uint8_t* qimg2 = smem8 + qbytes;                         // GEN only: qsel2
    uint8_t* timg = smem8 + (GEN ? 2 : 1) * qbytes;
    // whole images: base + clamped 1-based position; rings: base + (virtual position & mask)
    const uint32_t s_q = (uint32_t)__cvta_generic_to_shared(qimg) + (RING ? 0 : SEQ_PAD - 1);
    const uint32_t s_q2 = (uint32_t)__cvta_generic_to_shared(qimg2) + (RING ? 0 : SEQ_PAD - 1);
    const uint32_t s_t = (uint32_t)__cvta_generic_to_shared(timg) + (RING ? 0 : SEQ_PAD - 1);
    const uint32_t s_mb = (uint32_t)__cvta_generic_to_shared(timg + tbytes);
    const int NEG = -(1 << 29);
    const int d_ne1 = (int)pin((uint32_t)-cs.d_ext1), d_noe1 = (int)pin((uint32_t)-(cs.d_open1 + cs.d_ext1));
    const int d_ne2 = (int)pin((uint32_t)-cs.d_ext2), d_noe2 = (int)pin((uint32_t)-(cs.d_open2 + cs.d_ext2));
    // simple scoring has insertion == deletion costs: sharing the registers lets the compiler compute H + noe once per cell
    // for its two consumers (the E of the cell to the right, the F of the cell below)
    const int i_ne1 = GEN ? (int)pin((uint32_t)-cs.i_ext1) : d_ne1, i_noe1 = GEN ? (int)pin((uint32_t)-(cs.i_open1 + cs.i_ext1)) : d_noe1;
    const int i_ne2 = GEN ? (int)pin((uint32_t)-cs.i_ext2) : d_ne2, i_noe2 = GEN ? (int)pin((uint32_t)-(cs.i_open2 + cs.i_ext2)) : d_noe2;
    const int c_mat = (int)pin((uint32_t)cs.s_match), c_mis = (int)pin((uint32_t)cs.s_mismatch), c_neg = (int)pin((uint32_t)NEG);
    const uint32_t r0 = pin(cs.tab[0]), r1 = pin(cs.tab[1]), r2 = pin(cs.tab[2]), r3 = pin(cs.tab[3]);
    const int c_nsc = (int)pin((uint32_t)cs.n_score);

    // lowest virtual positions the CTA touches at iteration 0 (thread gl = 0 has the largest I0 and the smallest J0)
    const int g_first = (int)crank * (int)blockDim.x;
    const int ring_q0 = (tk.r_begin - (tk.band_lo + (g_first + (int)blockDim.x - 1) * 2 * K)) / 2 - K;
    const int ring_t0 = (tk.r_begin + (tk.band_lo + g_first * 2 * K)) / 2 - 1;
    if (RING) ring_fill<GEN>(tk, seqs, qimg, qimg2, timg, rmask, ring_q0, ring_t0, ring);
    else if (GEN) stage_sequences_gen(tk, seqs, qimg, qimg2, timg, qbytes, tbytes);
    else stage_sequences(tk, seqs, qimg, timg, qbytes, tbytes);
    if (RING) __syncthreads();
    if (MULTI && gl <= nwarp) {
        sts_v4(s_mb + MB_A + 16u * gl, 0u, 0u, (uint32_t)c_neg, (uint32_t)c_neg);
        sts_v4(s_mb + MB_B + 16u * gl, 0u, 0u, (uint32_t)c_neg, (uint32_t)c_neg);
        step_link_init<CL>(s_mb, gl);
    }
    if (MULTI && gl == 0) sts_v4(s_mb + MB_CST, 0u, 0u, (uint32_t)c_neg, (uint32_t)c_neg);
    pair_begin<WALK>(ar, tk, crank, s_slice, seqs, results, cigar, cs);
    step_barrier<CL>();
    uint8_t* const rows = pair_rows<CL>(ar, crank, s_slice);
    uint2* const bests = reinterpret_cast<uint2*>(rows + tk.tb_bytes);

    const int n_thr_active = tk.band_w / (2 * K);
    const bool active = g < n_thr_active;
    const bool lo_edge = (lane == 0);
    const bool hi_edge = (lane == 31) || (g == n_thr_active - 1);
    const int D0 = tk.band_lo + g * 2 * K;
    const int I0 = (tk.r_begin - D0) / 2, J0 = (tk.r_begin + D0) / 2;
    const int qmax = tk.m + 1, tmax = tk.n + 1;

    // second-level selector of a 32-bit lane from the qsel2 byte b (0x80 / 0xC4): nibbles (n0, 8|n0, 8|n0, 8|n0)
    auto sel2_of = [](uint32_t b) -> uint32_t { return b | ((b >> 4) * 0x1100u); };
    int HA[K], HB[K], E1[K], E2[K], F1[K], F2[K];
    uint32_t Q[K], Q2[K], T[K];
    #pragma unroll
    for (int k = 0; k < K; ++k) {
        HA[k] = 0; HB[k] = 0; E1[k] = c_neg; E2[k] = c_neg; F1[k] = c_neg; F2[k] = c_neg;
        const uint32_t qo = RING ? (uint32_t)((I0 - k) & rmask) : (uint32_t)__vimin_s32_relu(I0 - k, qmax);
        Q[k] = lds_u8(s_q + qo);
        Q2[k] = GEN ? sel2_of(lds_u8(s_q2 + qo)) : 0u;
        T[k] = lds_u8(s_t + (RING ? (uint32_t)((J0 + k) & rmask) : (uint32_t)__vimin_s32_relu(J0 + k, tmax)));
    }
    int qpos = I0 + 1, tpos = J0 + K;
    int best_lo = 0, best_hi = 0, snap_lo = 0, snap_hi = 0; uint32_t blk_lo = 0, blk_hi = 0;
    const int pitch = n_thr_active * tb_slot_bytes(KPG);     // bytes per traceback row (one step pair)
    uint8_t* tbp = rows + (size_t)g * tb_slot_bytes(KPG);
    const StepLink link = step_link<CL>(s_mb, warp, nwarp, crank, csize);
    // the last in-band thread reads the "outside the band" constants, whatever warp / CTA follows it
    const uint32_t s_upA = (g == n_thr_active - 1) ? s_mb + MB_CST : link.in_up(), s_loB = link.in_lo();

    // low bytes of a step pair in the shared row order (store_pair, KP = 4): word k = [A slot k, B slot k, A slot 4+k, B slot 4+k]
    auto store16 = [&](const int (&A)[K], const int (&B)[K]) {
        uint32_t w[KPG];
        #pragma unroll
        for (int k = 0; k < KPG; ++k)
            w[k] = __byte_perm(__byte_perm((uint32_t)A[k], (uint32_t)B[k], 0x0040), __byte_perm((uint32_t)A[KPG + k], (uint32_t)B[KPG + k], 0x0040), 0x5410);
        __stcs(reinterpret_cast<uint4*>(tbp), make_uint4(w[0], w[1], w[2], w[3]));
    };
    auto score = [&](uint32_t q, uint32_t q2, uint32_t t) -> int {
        if (GEN) {
            const uint32_t sel = q | t;
            return (sel & GEN_PAD) ? c_nsc : (int)prmt(prmt(r0, r1, sel), prmt(r2, r3, sel), q2);     // pad / letter outside ACGT: lowest entry
        }
        return (q == t) ? c_mat : c_mis;
    };
    // one cell: (left H, left E) feed E, (up H, up F) feed F, hd = diagonal H + substitution score
    auto cell = [&](int hd, int hl, int e1l, int e2l, int hu, int f1u, int f2u, int& e1, int& e2, int& f1, int& f2) -> int {
        e1 = __viaddmax_s32(e1l, d_ne1, hl + d_noe1);
        f1 = __viaddmax_s32(f1u, i_ne1, hu + i_noe1);
        if (TWO) {
            e2 = __viaddmax_s32(e2l, d_ne2, hl + d_noe2);
            f2 = __viaddmax_s32(f2u, i_ne2, hu + i_noe2);
            return __vimax3_s32_relu(__vimax3_s32(hd, e1, f1), e2, f2);
        }
        e2 = c_neg; f2 = c_neg;
        return __vimax3_s32_relu(hd, e1, f1);
    };

    #pragma unroll 1
    for (int it = 0; it < tk.n_iter; ++it) {
        if (RING && it && (it & (RING_CHUNK / K - 1)) == 0) {       // the next chunk of both windows (needed from the next iteration on)
            if (MULTI) __syncthreads();                             // the warps of a CTA drift by up to a step per warp: line them up around the refill
            ring_fill<GEN>(tk, seqs, qimg, qimg2, timg, rmask, ring_q0 + ring + (it - RING_CHUNK / K) * K, ring_t0 + ring + (it - RING_CHUNK / K) * K, RING_CHUNK);
            if (MULTI) __syncthreads(); else __syncwarp();
        }
        #pragma unroll
        for (int u = 0; u < K; ++u) {
            // =============================== A-step ===========================================
            {
                int sH = __shfl_up_sync(0xffffffffu, HB[K - 1], 1), sE1 = __shfl_up_sync(0xffffffffu, E1[K - 1], 1), sE2 = 0;
                if (TWO) sE2 = __shfl_up_sync(0xffffffffu, E2[K - 1], 1);
                #pragma unroll
                for (int k = K - 1; k >= 0; --k) {
                    if (k == 0) {              // the only cell that needs the lower neighbour: wait for its mailbox as late as possible
                        uint4 mb = make_uint4(0u, 0u, (uint32_t)c_neg, (uint32_t)c_neg);
                        if (MULTI) { if (it | u) link_wait_lo<CL>(link, lane, (uint32_t)((u & 1) ^ 1)); mb = lds_v4(s_loB); }
                        sH = lo_edge ? (int)mb.x : sH; sE1 = lo_edge ? (int)mb.z : sE1;
                        if (TWO) sE2 = lo_edge ? (int)mb.w : sE2;
                    }
                    const int hl = k ? HB[k ? k - 1 : 0] : sH, e1l = k ? E1[k ? k - 1 : 0] : sE1, e2l = k ? E2[k ? k - 1 : 0] : sE2;
                    const int hd = HA[k] + score(Q[(k - u + K) % K], Q2[(k - u + K) % K], T[(k + u) % K]);
                    int e1, e2, f1, f2;
                    HA[k] = cell(hd, hl, e1l, e2l, HB[k], F1[k], F2[k], e1, e2, f1, f2);
                    E1[k] = e1; E2[k] = e2; F1[k] = f1; F2[k] = f2;
                }
                #pragma unroll
                for (int k = 0; k < KPG; k += 2) { best_lo = __vimax3_s32(best_lo, HA[k], HA[k + 1]); best_hi = __vimax3_s32(best_hi, HA[KPG + k], HA[KPG + k + 1]); }
                if (MULTI && lane == 0) mailbox_store_A<CL>(link, (uint32_t)HA[0], 0u, (uint32_t)F1[0], (uint32_t)F2[0]);
                T[u % K] = lds_u8(s_t + (RING ? (uint32_t)(tpos & rmask) : (uint32_t)__vimin_s32_relu(tpos, tmax)));     // logical T[k] <- T[k+1]; new base in slot K-1
                ++tpos;
            }
            // =============================== B-step ===========================================
            {
                int sH = __shfl_down_sync(0xffffffffu, HA[0], 1), sF1 = __shfl_down_sync(0xffffffffu, F1[0], 1), sF2 = 0;
                if (TWO) sF2 = __shfl_down_sync(0xffffffffu, F2[0], 1);
                #pragma unroll
                for (int k = 0; k < K; ++k) {
                    if (k == K - 1) {          // the only cell that needs the upper neighbour
                        uint4 mb = make_uint4(0u, 0u, (uint32_t)c_neg, (uint32_t)c_neg);
                        if (MULTI) { link_wait_up<CL>(link, lane, (uint32_t)(u & 1)); mb = lds_v4(s_upA); }
                        sH = hi_edge ? (int)mb.x : sH; sF1 = hi_edge ? (int)mb.z : sF1;
                        if (TWO) sF2 = hi_edge ? (int)mb.w : sF2;
                    }
                    const int hu = (k < K - 1) ? HA[(k + 1) % K] : sH, f1u = (k < K - 1) ? F1[(k + 1) % K] : sF1, f2u = (k < K - 1) ? F2[(k + 1) % K] : sF2;
                    const int hd = HB[k] + score(Q[(k - u + K) % K], Q2[(k - u + K) % K], T[(k + u + 1) % K]);
                    int e1, e2, f1, f2;
                    HB[k] = cell(hd, HA[k], E1[k], E2[k], hu, f1u, f2u, e1, e2, f1, f2);
                    E1[k] = e1; E2[k] = e2; F1[k] = f1; F2[k] = f2;
                }
                #pragma unroll
                for (int k = 0; k < KPG; k += 2) { best_lo = __vimax3_s32(best_lo, HB[k], HB[k + 1]); best_hi = __vimax3_s32(best_hi, HB[KPG + k], HB[KPG + k + 1]); }
                if (MULTI && lane == 31) mailbox_store_B<CL>(link, (uint32_t)HB[K - 1], 0u, (uint32_t)E1[K - 1], (uint32_t)E2[K - 1]);
                if (active) store16(HA, HB);
                tbp += pitch;
                const uint32_t qo = RING ? (uint32_t)(qpos & rmask) : (uint32_t)__vimin_s32_relu(qpos, qmax);   // logical Q[k] <- Q[k-1]; new base in slot 0
                Q[(K - 1 - u) % K] = lds_u8(s_q + qo);
                if (GEN) Q2[(K - 1 - u) % K] = sel2_of(lds_u8(s_q2 + qo));
                ++qpos;
            }
            if (u == KPG - 1 || u == K - 1) {          // end of an 8-step block: which block first reached the running maximum
                const uint32_t blk = 2u * (uint32_t)it + (u == K - 1 ? 1u : 0u);
                if (best_lo != snap_lo) { blk_lo = blk; snap_lo = best_lo; }
                if (best_hi != snap_hi) { blk_hi = blk; snap_hi = best_hi; }
            }
        }
    }
    if (active) {
        bests[2 * g]     = make_uint2((uint32_t)best_lo, blk_lo);
        bests[2 * g + 1] = make_uint2((uint32_t)best_hi, blk_hi);
    }
    step_barrier<CL>();                        // rows and best records complete; no CTA leaves while a neighbour may still store into its mailbox
    if (WALK && warp == 0 && crank == 0) pair_end(ar, tk, task_index, seqs, results, cigar, cs, s_slice, ar.cor_smem ? (uint32_t)__cvta_generic_to_shared(smem8) : 0u);
}

}  // namespace svp
