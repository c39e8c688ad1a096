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
//   a 16-byte shared-memory mailbox with its own mbarrier (no CTA barrier in the loop; st.async + mbarrier links into the
//   neighbouring CTA's shared memory when the band spans the CTAs of a cluster). The 2-bit packed sequences are read with
//   coalesced 128-bit loads and staged in shared memory as one code byte per base (whole images or sliding windows). Per step
//   pair each thread streams the low bytes of its 4*KP H values to its slot of the pair's traceback rows in HBM (st.global.cs).
//   Every variant exists with (arena / queue scheduling) and without (WALK = false: wave scheduling) the in-kernel walk.
//
// Walk (traceback)   rebuilds exact H values along the path from the stored low bytes (neighbouring H differ by less than
//   128). svp_traceback_warp<CG, LPP>: a group of LPP lanes per pair, a diagonal run per round (the whole
//   warp keeps a corridor of the rows in shared memory); svp_traceback_groups<LPP, CG>: 32 / LPP pairs per warp on ONE
//   instruction stream (the bulk walker of serial waves, svp_tbu_kernel); svp_traceback_thread: one pair per lane.
// Scheduling   waves (the host assigns the slices, walks are kernels of their own), arena (device-side slice allocator
//   ArenaPool, the CTA walks its own pair), queue (resident walker kernels): svp_align.cu, DESIGN.md §4.8.
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
    uint64_t tb_bytes;      // bytes of this pair's traceback rows; its per-thread best records (8 B each) follow them in the arena slice
    uint64_t slice_bytes;   // rows + best records, a multiple of 256: what the pair takes from the arena
    uint64_t tb_off;        // wave scheduling: arena offset of the pair's slice, assigned by the host
    uint64_t cells;         // spec cells (band ∩ matrix), copied into the result
    uint32_t cigar_off, cigar_cap;
};

struct Consts {             // scoring in the forms the kernels use (svp_align.cu: make_consts)
    // fast path (svp_fwd_kernel): one match and one mismatch score, insertion costs == deletion costs; s16x2
    uint32_t mat, mis;      // match score, mismatch score (negative)
    uint32_t ne1, ne2;      // -ext1, -ext2
    uint32_t noe1, noe2;    // -(open1+ext1), -(open2+ext2)
    // general path (svp_fwdg_kernel): full 4x4 matrix, separate deletion (E) / insertion (F) pieces; s16x2
    uint32_t d_ne1, d_noe1, d_ne2, d_noe2;
    uint32_t i_ne1, i_noe1, i_ne2, i_noe2;
    uint32_t tab[4];        // substitution matrix as signed bytes: byte t of tab[q] = s(q, t)
    int32_t  two_piece;     // both gap kinds carry two pieces (a one-piece kind repeats its piece)
    int32_t  general;       // 0: fast path applies, 1: general path
    // scalar copies (32-bit lane kernel, traceback)
    int32_t  s_match, s_mismatch;
    int32_t  d_open1, d_ext1, d_open2, d_ext2, i_open1, i_ext1, i_open2, i_ext2;
    int32_t  a_max;         // largest matrix entry (score-range test)
    int32_t  rebase_ok;     // the rebased s16x2 kernel is exact for this scoring (make_consts)
    int32_t  min_signal, zdrop, zdrop_ext;
    int32_t  n_score;       // score of a cell with a base outside ACGT on either side: the lowest matrix entry
    uint32_t n_pk;          // ... as s16x2
};
// byte offset of a sequence's mask behind its 2-bit codes (include/svp_align.h: SVP_SEQ_CODE_BYTES)
__host__ __device__ __forceinline__ uint64_t seq_code_bytes(int len) { return (((uint64_t)len + 3) / 4 + 15) & ~(uint64_t)15; }

constexpr uint32_t TASK_INVALID = 0x80000000u;   // Task.flags: descriptor rejected by the planner
constexpr uint32_t TASK_TOO_LONG = 0x40000000u;  // ... because the two sequences do not fit the shared-memory images
constexpr uint32_t TASK_RING = 0x10000000u;      // ... on its sliding-window variant (sequences too long for whole shared-memory images)
constexpr uint32_t TASK_S32 = 0x20000000u;       // pair runs on the 32-bit lane kernel (scores beyond the s16 range)
constexpr uint32_t TASK_KP_SHIFT = 8;            // bits 8..12 of Task.flags: KP of the kernel variant the pair runs on
constexpr uint32_t TASK_KP_MASK = 0x1F00u;
constexpr uint32_t TASK_CL_SHIFT = 16;           // bits 16..19: CTAs per cluster (0/1: no cluster)
constexpr uint32_t TASK_CL_MASK = 0xF0000u;
constexpr uint32_t TASK_QN = 0x04000000u;        // the query / the target carries a mask of letters outside ACGT behind its 2-bit codes
constexpr uint32_t TASK_TN = 0x02000000u;
constexpr uint32_t TASK_REBASE = 0x08000000u;    // pair runs on the s16x2 fast path with per-warp score offsets (scores beyond the s16 range)
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

// ---- thread-block clusters: bands wider than one CTA covers (16 warps) are split over the CTAs of a cluster; the
// boundary mailboxes of neighbouring CTAs live in each other's shared memory (DSMEM) and the per-step CTA barrier
// becomes a cluster barrier.
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_nctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t mapa_shared(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t caddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared::cluster.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(caddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// one step's barrier / mailbox store, CTA-local or cluster-wide
template <bool CL> __device__ __forceinline__ void step_barrier() { if (CL) cluster_sync_all(); else __syncthreads(); }
// Step synchronisation. Inside a CTA: the mailbox stores of a step are followed by one CTA barrier. Between the CTAs of a
// cluster there is NO barrier: the two warps that face a neighbouring CTA exchange their mailbox entries point to point with
// st.async into the neighbour's shared memory, completing 16 transaction bytes on an mbarrier that lives next to the
// destination slot (STAS + SYNCS in SASS). The receiving warp arms that mbarrier (arrive.expect_tx) and waits on its phase
// only just before the one cell that reads the neighbour's entry — the last cell a step computes (slot 0 of an A-step, the
// last slot of a B-step) — so the ~215-cycle DSMEM latency hides behind the step's other cells and no thread ever executes
// a fence. (A cluster barrier per step with release/acquire semantics stalled every warp on the release fence: ncu showed
// 5.3 warps per issue slot waiting on membar, profiles/ncu_cluster_r01.json; config 3 ran at 485 instead of 8xx GCUPS.)
// A mailbox slot is safe to overwrite when its writer gets there: the exchange between two neighbours is a ping-pong (the
// A entry travels down, the B entry up), so a writer produces entry t+1 only after it consumed the entry its reader sent
// after reading entry t. One mbarrier phase per step; KP and K are even, so the phase parity of step u is u & 1.
__device__ __forceinline__ void mbar_init(uint32_t addr, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(addr), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t addr, uint32_t bytes) {
    asm volatile("{ .reg .b64 st; mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1; }" :: "r"(addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t addr, uint32_t parity) {
    asm volatile("{ .reg .pred p;\n\tSVP_MBAR_WAIT: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra SVP_MBAR_DONE;\n\tbra SVP_MBAR_WAIT;\n\tSVP_MBAR_DONE: }"
                 :: "r"(addr), "r"(parity) : "memory");
}
__device__ __forceinline__ void st_async_v4(uint32_t caddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t cmbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1,%2,%3,%4}, [%5];"
                 :: "r"(caddr), "r"(a), "r"(b), "r"(c), "r"(d), "r"(cmbar) : "memory");
}
// Per-warp view of the step links. Every warp exchanges one 16-byte mailbox entry per step with each neighbouring warp:
// its A entry (lane 0, after an A-step) travels DOWN to the warp below, its B entry (lane 31, after a B-step) UP to the warp
// above. There is NO CTA barrier in the loop: every inbound slot has its own mbarrier, the writer completes one phase per step
// pair, the reader waits for that phase only just before the one cell that consumes the entry (the last cell a step computes),
// so a warp may run up to a whole step ahead of its neighbours. (One __syncthreads() per step made every warp wait for the
// slowest of up to 16: ncu showed the multi-warp variants at 57-67 % of the ALU pipe with 1.1 warps per issue slot stalled on
// the barrier, against 89 % for single-warp CTAs, profiles/ncu_fwd_r01_v6.json.)
//   neighbour in the same CTA:        st.shared + mbarrier.arrive (release) by the writer lane, try_wait (acquire) by the reader
//   neighbour in the next CTA of a cluster: st.async ... mbarrier::complete_tx::bytes into the neighbour's shared memory; the reader
//                                     arms its mbarrier with arrive.expect_tx (STAS + SYNCS in SASS; no fence anywhere)
// A slot is single-buffered: the exchange between two neighbours is a ping-pong, a writer produces entry t+1 only after it
// consumed the entry its reader sent after reading entry t.
// Shared memory behind the sequence images of a CTA (all forward kernel families), a fixed layout for up to 16 warps so
// that every address is one per-warp register plus an immediate:
//   A[0..16], B[0..16]        16-byte mailboxes (A[w]: written by warp w, read by warp w-1; B[w]: written by warp w-1, read by
//                             warp w; A[nwarp] / B[0]: inbound slots of the cluster links)
//   barA[0..16], barB[0..16]  one mbarrier per slot
//   one 16-byte slot holding the "outside the band" constants (read by the last in-band thread when it is not lane 31)
//   wb[0..17]                 per-warp score offsets of the rebased kernel (wb[0] / wb[nwarp+1]: the neighbouring CTAs' edge warps)
constexpr uint32_t MB_SLOTS = 17, MB_A = 0, MB_B = MB_A + 16 * MB_SLOTS, MB_BARA = MB_B + 16 * MB_SLOTS, MB_BARB = MB_BARA + 8 * MB_SLOTS,
                   MB_CST = MB_BARB + 8 * MB_SLOTS, MB_WB = MB_CST + 16, MB_BYTES = MB_WB + 80;
__host__ __device__ __forceinline__ int mailbox_bytes(int) { return (int)MB_BYTES; }
__device__ __forceinline__ uint32_t smem_cst(uint32_t s_mb, int) { return s_mb + MB_CST; }
__device__ __forceinline__ uint32_t smem_wb(uint32_t s_mb, int) { return s_mb + MB_WB; }
__device__ __forceinline__ void mbar_arrive(uint32_t addr) {
    asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" :: "r"(addr) : "memory");
}
struct StepLink {
    bool lo, up;                        // this warp has a lower / upper neighbour warp
    bool rlo, rup;                      // ... and it lives in the neighbouring CTA of the cluster
    uint32_t mw, bw;                    // s_mb + 16 * warp, s_mb + 8 * warp: this warp's mailbox / mbarrier slots at fixed offsets
    uint32_t rem_A, rem_bar_A;          // lower neighbour CTA's A[nwarp] slot and its mbarrier (cluster addresses)
    uint32_t rem_B, rem_bar_B;          // upper neighbour CTA's B[0] slot and its mbarrier
    uint32_t rem_wb_lo, rem_wb_up;      // score-offset slots (rebased kernel): lower neighbour's wb[nwarp+1], upper neighbour's wb[0]
    // inbound: B[warp] (from below), A[warp + 1] (from above); outbound: A[warp], B[warp + 1]
    __device__ __forceinline__ uint32_t in_lo() const { return mw + MB_B; }
    __device__ __forceinline__ uint32_t in_up() const { return mw + MB_A + 16u; }
    __device__ __forceinline__ uint32_t out_A() const { return mw + MB_A; }
    __device__ __forceinline__ uint32_t out_B() const { return mw + MB_B + 16u; }
    __device__ __forceinline__ uint32_t bar_lo() const { return bw + MB_BARB; }
    __device__ __forceinline__ uint32_t bar_up() const { return bw + MB_BARA + 8u; }
    __device__ __forceinline__ uint32_t sig_A() const { return bw + MB_BARA; }
    __device__ __forceinline__ uint32_t sig_B() const { return bw + MB_BARB + 8u; }
};
template <bool CL>
__device__ __forceinline__ StepLink step_link(uint32_t s_mb, int warp, int nwarp, uint32_t crank, uint32_t csize) {
    StepLink L{};
    L.rlo = CL && warp == 0 && crank > 0;
    L.rup = CL && warp == nwarp - 1 && crank + 1 < csize;
    L.lo = warp > 0 || L.rlo;
    L.up = warp < nwarp - 1 || L.rup;
    L.mw = s_mb + 16u * (uint32_t)warp;
    L.bw = s_mb + 8u * (uint32_t)warp;
    if (CL) {
        if (L.rlo) { L.rem_A = mapa_shared(s_mb + MB_A + 16u * (uint32_t)nwarp, crank - 1); L.rem_bar_A = mapa_shared(s_mb + MB_BARA + 8u * (uint32_t)nwarp, crank - 1);
                     L.rem_wb_lo = mapa_shared(s_mb + MB_WB + 4u * (uint32_t)(nwarp + 1), crank - 1); }
        if (L.rup) { L.rem_B = mapa_shared(s_mb + MB_B, crank + 1); L.rem_bar_B = mapa_shared(s_mb + MB_BARB, crank + 1); L.rem_wb_up = mapa_shared(s_mb + MB_WB, crank + 1); }
    }
    return L;
}
// mbarriers of every slot (one arrival per phase: the writer lane's arrive, or — cluster links — the reader's arrive.expect_tx);
// called by the first nwarp + 1 threads before the initial barrier
template <bool CL> __device__ __forceinline__ void step_link_init(uint32_t s_mb, int idx) {
    mbar_init(s_mb + MB_BARA + 8u * (uint32_t)idx, 1); mbar_init(s_mb + MB_BARB + 8u * (uint32_t)idx, 1);
    if (CL) mbar_fence_init();
}
__device__ __forceinline__ void sts_u32(uint32_t saddr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(saddr), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t lds_u32(uint32_t saddr) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr) : "memory"); return v; }
__device__ __forceinline__ void st_cluster_u32(uint32_t caddr, uint32_t v) { asm volatile("st.shared::cluster.u32 [%0], %1;" :: "r"(caddr), "r"(v) : "memory"); }
// before the edge cell of an A-step (lower neighbour's B entry of the previous step pair) / of a B-step (upper neighbour's A
// entry of the same step pair); parity = parity of the step pair that produced the entry
template <bool CL> __device__ __forceinline__ void link_wait_lo(const StepLink& L, int lane, uint32_t parity) {
    if (L.lo) { if (CL && L.rlo && lane == 0) mbar_expect_tx(L.bar_lo(), 16u); mbar_wait(L.bar_lo(), parity); }
}
template <bool CL> __device__ __forceinline__ void link_wait_up(const StepLink& L, int lane, uint32_t parity) {
    if (L.up) { if (CL && L.rup && lane == 0) mbar_expect_tx(L.bar_up(), 16u); mbar_wait(L.bar_up(), parity); }
}
// mailbox stores of lane 0 after an A-step / of lane 31 after a B-step (called by that lane only)
template <bool CL> __device__ __forceinline__ void mailbox_store_A(const StepLink& L, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    if (CL && L.rlo) st_async_v4(L.rem_A, a, b, c, d, L.rem_bar_A);
    else if (L.lo) { sts_v4(L.out_A(), a, b, c, d); mbar_arrive(L.sig_A()); }
}
template <bool CL> __device__ __forceinline__ void mailbox_store_B(const StepLink& L, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    if (CL && L.rup) st_async_v4(L.rem_B, a, b, c, d, L.rem_bar_B);
    else if (L.up) { sts_v4(L.out_B(), a, b, c, d); mbar_arrive(L.sig_B()); }
}
// CG: the rows were written by the grid that reads them (arena scheduling: the walk runs inside the forward kernel, slices are
// recycled under it) -> ld.global.cg, L1 must not serve them; false: written by an earlier kernel (wave scheduling) -> plain loads
template <bool CG, int LPP = 32>
__device__ __noinline__ void svp_traceback_warp(const Task& tk, const uint8_t* __restrict__ rows, const uint2* __restrict__ bests,
                                                const uint8_t* __restrict__ seqs, svp_result* __restrict__ results,
                                                uint32_t* __restrict__ cigar, const Consts& cs, uint32_t s_cor);
template <bool CG>
__device__ __noinline__ void svp_traceback_thread(const Task& tk, const uint8_t* __restrict__ rows, const uint2* __restrict__ bests,
                                                  const uint8_t* __restrict__ seqs, svp_result* __restrict__ results,
                                                  uint32_t* __restrict__ cigar, const Consts& cs);
// ---------------------------------------------------------------------------------------------
// Device-side arena. The traceback rows of a pair live from the start of its forward pass to the end of its walk, both inside
// one CTA (cluster), so the arena only has to hold the pairs that are RESIDENT on the GPU: thread 0 of a pair's (first) CTA
// takes a slice when the CTA starts and returns it when the walk is done. One pool per SM (a CTA allocates from the pool of
// the SM it runs on: no contention beyond the ~16 CTAs of an SM) plus one global pool for slices too large for a per-SM
// region (long reads in wide bands, clusters). A pool is a first-fit list of live slices under a spin lock; allocation waits
// (nanosleep) when the pool is full — the holders are running CTAs that do not depend on the waiter, so it always ends. The
// host bounds the residency per SM through the shared-memory request of every launch (svp_align.cu: ballast), so that the
// slices live on an SM fit its region with room for fragmentation and the wait is the exception.
// ---------------------------------------------------------------------------------------------
#ifndef SVP_POOL_MAX
#define SVP_POOL_MAX 48
#endif
constexpr int POOL_MAX = SVP_POOL_MAX;
struct ArenaPool {
    uint32_t lock, n;                 // spin lock (0 free), live slices
    uint64_t base, bytes;             // the pool's region: arena offset and size
    uint64_t off[POOL_MAX], len[POOL_MAX];   // live slices, ascending by offset (relative to base)
    unsigned long long waits;         // allocations that had to wait (statistics)
};
// Queue scheduling: finished forward passes of a call, waiting for their walk. An entry is published by writing its task index
// last; entries are popped in order (head) — in batches of up to 32 by the call's walker kernel (one walk per LANE), one at a
// time by a CTA that is waiting for arena space (it helps instead of idling, one walk per WARP).
struct WalkEntry { uint64_t arena_off; uint32_t pool_id; uint32_t task; };
constexpr uint32_t WALK_EMPTY = 0xFFFFFFFFu;
struct WalkQueue {
    uint32_t head, tail, n, pad;      // next entry to walk, entries handed out to writers, queued pairs of the call
    WalkEntry e[1];                   // n entries, initialised to 0xFF
};
struct Arena {
    uint8_t* base;
    ArenaPool* pools;                 // pools[0 .. n_sm): per SM, pools[n_sm]: global
    uint32_t n_sm;
    uint32_t use_global;              // this launch's pairs allocate from the global pool
    unsigned long long* clocks;       // [0] summed forward ns, [1] summed inline-walk ns, [2] pairs walked, [3] walks by waiting CTAs,
                                      // [4] summed busy ns of the walker warps, [5] batches they walked (statistics)
    uint32_t static_slices;           // wave scheduling: the host assigned every pair its slice (Task.tb_off) and launches the walks itself
    uint32_t skip_walk;               // measurement aid (env SVP_SKIP_WALK): no walk at all, results are not written — the forward-only rate
    WalkQueue* queue;                 // queue scheduling: the call's queue (nullptr: none)
    const Task* tasks;                // ... the call's task array (queue entries index it)
    uint32_t inline_walk;             // ... this launch's CTAs walk their own pair after the forward pass instead of queueing it
    uint32_t walk_long;               // ... and so does every pair with m + n beyond this (a long walk would hold a walker's 31 other lanes)
    uint32_t cor_smem;                // the launch's dynamic shared memory holds at least COR_BYTES: walks inside the CTA keep their corridor there
};
__device__ __forceinline__ uint32_t smid() { uint32_t r; asm volatile("mov.u32 %0, %%smid;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned long long globaltimer_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
// one thread; one attempt to take a slice of `bytes` bytes (a multiple of 256) from pool P: arena offset, or ~0 when it does not fit now
__device__ __noinline__ uint64_t arena_try_alloc(ArenaPool* P, uint64_t bytes) {
    volatile uint64_t* off = P->off; volatile uint64_t* len = P->len;
    while (atomicCAS(&P->lock, 0u, 1u) != 0u) __nanosleep(100);
    __threadfence();
    const uint32_t n = *(volatile uint32_t*)&P->n;
    uint64_t prev_end = 0, got = ~(uint64_t)0; int at = -1;
    if (n < (uint32_t)POOL_MAX) {
        for (uint32_t k = 0; k < n; ++k) {
            if (off[k] - prev_end >= bytes) { at = (int)k; break; }
            prev_end = off[k] + len[k];
        }
        if (at < 0 && P->bytes - prev_end >= bytes) at = (int)n;
    }
    if (at >= 0) {
        for (int k = (int)n; k > at; --k) { off[k] = off[k - 1]; len[k] = len[k - 1]; }
        off[at] = prev_end; len[at] = bytes;
        *(volatile uint32_t*)&P->n = n + 1;
        got = P->base + prev_end;
    }
    __threadfence();
    atomicExch(&P->lock, 0u);
    return got;
}
__device__ __noinline__ void arena_free(const Arena& A, uint32_t pool_id, uint64_t arena_off) {
    ArenaPool* P = A.pools + pool_id;
    volatile uint64_t* off = P->off; volatile uint64_t* len = P->len;
    const uint64_t rel = arena_off - P->base;
    while (atomicCAS(&P->lock, 0u, 1u) != 0u) __nanosleep(100);
    __threadfence();
    const uint32_t n = *(volatile uint32_t*)&P->n;
    uint32_t k = 0;
    while (k < n && off[k] != rel) ++k;
    if (k < n) {
        for (; k + 1 < n; ++k) { off[k] = off[k + 1]; len[k] = len[k + 1]; }
        *(volatile uint32_t*)&P->n = n - 1;
    }
    __threadfence();
    atomicExch(&P->lock, 0u);
}
// One walk by the calling warp (all 32 lanes): rows at arena offset `o` of pool `pool_id`; the slice is returned afterwards.
__device__ __forceinline__ void walk_and_free(const Arena& A, const Task& tk, uint64_t o, uint32_t pool_id, const uint8_t* __restrict__ seqs,
                                              svp_result* __restrict__ results, uint32_t* __restrict__ cigar, const Consts& cs, int helper, uint32_t s_cor) {
    const unsigned long long t1 = globaltimer_ns();
    if (!A.skip_walk) svp_traceback_warp<true>(tk, A.base + o, reinterpret_cast<const uint2*>(A.base + o + tk.tb_bytes), seqs, results, cigar, cs, s_cor);
    __syncwarp();
    if ((threadIdx.x & 31) == 0) {
        arena_free(A, pool_id, o);
        atomicAdd(A.clocks + 1, globaltimer_ns() - t1); atomicAdd(A.clocks + 2, 1ull);
        if (helper) atomicAdd(A.clocks + 3, 1ull);
    }
}
// Try to take the next queued pair and walk it (whole warp). Returns 0: walked one, 1: nothing ready, 2: the call's queue is drained.
__device__ __noinline__ int walk_one_queued(const Arena& A, const uint8_t* __restrict__ seqs, svp_result* __restrict__ results,
                                            uint32_t* __restrict__ cigar, const Consts& cs, int helper) {
    WalkQueue* Q = A.queue;
    const int lane = threadIdx.x & 31;
    uint32_t h = 0; int state = 1;
    if (lane == 0) {
        h = *(volatile uint32_t*)&Q->head;
        if (h >= Q->n) state = 2;
        else if (*(volatile uint32_t*)&Q->e[h].task != WALK_EMPTY && atomicCAS(&Q->head, h, h + 1u) == h) state = 0;
    }
    state = __shfl_sync(0xffffffffu, state, 0);
    if (state != 0) return state;
    h = __shfl_sync(0xffffffffu, h, 0);
    __threadfence();
    const uint32_t task = *(volatile uint32_t*)&Q->e[h].task, pool_id = *(volatile uint32_t*)&Q->e[h].pool_id;
    const uint64_t o = *(volatile uint64_t*)&Q->e[h].arena_off;
    const Task tk = A.tasks[task];
    walk_and_free(A, tk, o, pool_id, seqs, results, cigar, cs, helper, 0u);     // (a helper's shared memory holds its own pair's sequences)
    return 0;
}
// Start of a pair's CTA(s). Wave scheduling: the slice is where the host put it. Arena / queue scheduling: warp 0 of the first CTA
// takes the pair's slice (lane 0 allocates) and notes it in shared memory (s_slice: uint64[3] = arena offset, pool id, start
// time). When the pool is full the warp waits — the holders are running CTAs and queued walks that do not depend on the waiter;
// with a queue it does not idle: it walks a queued pair of the call (which frees a slice) and tries again, so the call makes
// progress even while its walker kernel is not resident. After the caller's initial barrier every CTA of a cluster reads the
// offset from the first CTA's shared memory (pair_rows; remote shared memory may only be touched once the cluster barrier has
// shown that CTA running).
template <bool WALK>
__device__ __forceinline__ void pair_begin(const Arena& A, const Task& tk, uint32_t crank, unsigned long long* s_slice,
                                           const uint8_t* __restrict__ seqs, svp_result* __restrict__ results, uint32_t* __restrict__ cigar, const Consts& cs) {
    if (!WALK || A.static_slices) {
        if (threadIdx.x == 0 && crank == 0) { s_slice[0] = tk.tb_off; s_slice[1] = 0; s_slice[2] = 0; }
        return;
    }
    if (threadIdx.x < 32 && crank == 0) {
        const int lane = threadIdx.x;
        const uint32_t pool_id = A.use_global ? A.n_sm : min(smid(), A.n_sm - 1u);
        bool waited = false;
        for (;;) {
            unsigned long long o = ~0ull;
            if (lane == 0) o = arena_try_alloc(A.pools + pool_id, tk.slice_bytes);
            o = __shfl_sync(0xffffffffu, o, 0);
            if (o != ~0ull) {
                if (lane == 0) { s_slice[0] = o; s_slice[1] = pool_id; s_slice[2] = globaltimer_ns(); if (waited) atomicAdd(&A.pools[pool_id].waits, 1ull); }
                break;
            }
            waited = true;
            if (!A.queue || walk_one_queued(A, seqs, results, cigar, cs, 1) != 0) __nanosleep(20000);
        }
    }
}
template <bool CL>
__device__ __forceinline__ uint8_t* pair_rows(const Arena& A, uint32_t crank, const unsigned long long* s_slice) {
    unsigned long long o;
    if (CL && crank != 0) {
        const uint32_t ra = mapa_shared((uint32_t)__cvta_generic_to_shared(s_slice), 0u);
        asm volatile("ld.shared::cluster.u64 %0, [%1];" : "=l"(o) : "r"(ra) : "memory");
    } else o = s_slice[0];
    return A.base + o;
}
// End of a pair: called by warp 0 of the first CTA after the barrier that follows the forward pass. Wave scheduling: nothing —
// the host launches the wave's walks. Queue scheduling: the pair is queued for the call's walkers and the CTA ends, releasing its
// registers for the next forward pass (long pairs excepted). Arena scheduling: the warp walks the pair right here.
__device__ __forceinline__ void pair_end(const Arena& A, const Task& tk, uint32_t task_index, const uint8_t* __restrict__ seqs,
                                         svp_result* __restrict__ results, uint32_t* __restrict__ cigar, const Consts& cs, const unsigned long long* s_slice,
                                         uint32_t s_cor) {
    if (A.static_slices) return;
    const uint64_t o = s_slice[0];
    if ((threadIdx.x & 31) == 0) atomicAdd(A.clocks + 0, globaltimer_ns() - s_slice[2]);
    if (A.queue && !A.inline_walk && (uint32_t)tk.m + (uint32_t)tk.n <= A.walk_long) {
        if ((threadIdx.x & 31) == 0) {
            WalkQueue* Q = A.queue;
            const uint32_t slot = atomicAdd(&Q->tail, 1u);
            Q->e[slot].arena_off = o; Q->e[slot].pool_id = (uint32_t)s_slice[1];
            __threadfence();                                  // rows, best records and the entry's payload before its publication
            *(volatile uint32_t*)&Q->e[slot].task = task_index;
        }
        return;
    }
    walk_and_free(A, tk, o, (uint32_t)s_slice[1], seqs, results, cigar, cs, 0, s_cor);
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

// overwrite the image bytes of the masked positions of a sequence (include/svp_align.h: SVP_SEQ_NMASK) with `code`
__device__ __forceinline__ void mask_to_image(const uint8_t* __restrict__ seq, int len, uint8_t* img, uint8_t code, bool rc) {
    const uint32_t* mk = reinterpret_cast<const uint32_t*>(seq + seq_code_bytes(len));
    for (int w = threadIdx.x; w < (len + 31) >> 5; w += blockDim.x) {
        uint32_t bits = __ldg(mk + w);
        while (bits) {
            const int p = w * 32 + __ffs(bits) - 1;
            bits &= bits - 1;
            if (p < len) img[SEQ_PAD + (rc ? len - 1 - p : p)] = code;
        }
    }
}

// Stage both sequences of a pair as code bytes in shared memory (query reverse-complemented on the fly when
// the pair asks for it). The pads are NaN codes that differ between query (0x7E) and target (0x7D), so padded
// positions compare unequal both as fp16 (s16x2 kernel) and as integers (s32 kernel). Ends with a barrier.
__device__ __forceinline__ void stage_sequences(const Task& tk, const uint8_t* __restrict__ seqs, uint8_t* qimg, uint8_t* timg,
                                                int qbytes, int tbytes) {
    const int g = threadIdx.x;
    // 128-bit coalesced loads of the 2-bit packed reads (every sequence starts on and is readable up to a 16-byte boundary):
    // one load = 64 bases = four 16-byte groups of the image
    auto stage64 = [&](const uint8_t* src, uint8_t* img, int len) {
        const uint4* v = reinterpret_cast<const uint4*>(src);
        for (int w = g; w < (len + 63) >> 6; w += blockDim.x) {
            const uint4 p = __ldg(v + w);
            uint4* dst = reinterpret_cast<uint4*>(img + SEQ_PAD + 64 * w);
            const int groups = min(4, ((len + 15) >> 4) - 4 * w);      // do not run over the image's end
            const uint32_t word[4] = { p.x, p.y, p.z, p.w };
            #pragma unroll
            for (int k = 0; k < 4; ++k) if (k < groups) dst[k] = expand16(word[k]);
        }
    };
    stage64(seqs + tk.t_off, timg, tk.n);
    const uint32_t* qsrc = reinterpret_cast<const uint32_t*>(seqs + tk.q_off);
    if (!(tk.flags & SVP_PAIR_REVCOMP)) {
        stage64(seqs + tk.q_off, qimg, tk.m);
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
    // letters outside ACGT (mask behind the codes): the pad code, which never equals anything -> the mismatch score
    if (tk.flags & TASK_TN) mask_to_image(seqs + tk.t_off, tk.n, timg, 0x7D, false);
    if (tk.flags & TASK_QN) mask_to_image(seqs + tk.q_off, tk.m, qimg, 0x7E, (tk.flags & SVP_PAIR_REVCOMP) != 0);
    // pads (after the bulk stores: the last 16-byte group may cover the trailing pad)
    for (int k = g; k < SEQ_PAD; k += blockDim.x) { qimg[k] = 0x7E; timg[k] = 0x7D; }
    for (int k = SEQ_PAD + tk.m + g; k < qbytes; k += blockDim.x) qimg[k] = 0x7E;
    for (int k = SEQ_PAD + tk.n + g; k < tbytes; k += blockDim.x) timg[k] = 0x7D;
    __syncthreads();
}

// Sliding sequence windows (RING variants): instead of whole images, a CTA keeps only the part of each sequence its
// anti-diagonal currently touches. The bases ENTERING the registers of a CTA lie within blockDim * k consecutive
// positions of each sequence (k = positions per thread), advancing by a fixed number of positions per loop iteration,
// so a power-of-two ring of R >= blockDim*k + 4k + RING_CHUNK bytes per image, refilled with the next RING_CHUNK
// positions every RING_CHUNK / (positions per iteration) iterations, always holds them. The ring is indexed by the
// VIRTUAL position (pos & (R-1), two's complement); positions outside [1, len] hold the pad code, so no clamping is
// needed. Shared memory per CTA no longer grows with the read length (occupancy for long reads), and sequences of any
// length fit.
constexpr int RING_CHUNK = 256;
__host__ __device__ __forceinline__ int ring_bytes(int n_threads, int k) {
    int r = 1024;
    while (r < n_threads * k + 4 * k + RING_CHUNK) r *= 2;
    return r;
}
// code (0..3) of oriented query position p / target position p (1-based), or -1 outside the sequence
// (mask: the sequence's mask of letters outside ACGT, or nullptr — a masked position reads as "outside": the pad code)
__device__ __forceinline__ int packed_code(const uint32_t* __restrict__ src, int len, int p, bool rc, const uint32_t* __restrict__ mask = nullptr) {
    if (p < 1 || p > len) return -1;
    const unsigned o = rc ? (unsigned)(len - p) : (unsigned)(p - 1);
    if (mask && ((__ldg(mask + (o >> 5)) >> (o & 31u)) & 1u)) return -1;
    const int c = (int)((__ldg(src + (o >> 4)) >> ((o & 15u) * 2u)) & 3u);
    return rc ? 3 - c : c;
}
__device__ __forceinline__ const uint32_t* seq_mask(const uint8_t* __restrict__ seqs, uint64_t off, int len, bool has) {
    return has ? reinterpret_cast<const uint32_t*>(seqs + off + seq_code_bytes(len)) : nullptr;
}
// fp16-coded windows of the fast path: `count` positions from q_first / t_first on
__device__ __forceinline__ void ring_fill_fast(const Task& tk, const uint8_t* __restrict__ seqs, uint8_t* qring, uint8_t* tring,
                                               int mask, int q_first, int t_first, int count) {
    const uint32_t* qsrc = reinterpret_cast<const uint32_t*>(seqs + tk.q_off);
    const uint32_t* tsrc = reinterpret_cast<const uint32_t*>(seqs + tk.t_off);
    const bool rc = (tk.flags & SVP_PAIR_REVCOMP) != 0;
    const uint32_t* qm = seq_mask(seqs, tk.q_off, tk.m, (tk.flags & TASK_QN) != 0);
    const uint32_t* tm = seq_mask(seqs, tk.t_off, tk.n, (tk.flags & TASK_TN) != 0);
    for (int x = threadIdx.x; x < count; x += blockDim.x) {
        const int qc = packed_code(qsrc, tk.m, q_first + x, rc, qm), tc = packed_code(tsrc, tk.n, t_first + x, false, tm);
        qring[(q_first + x) & mask] = (uint8_t)(qc < 0 ? 0x7E : 0x3C + qc);
        tring[(t_first + x) & mask] = (uint8_t)(tc < 0 ? 0x7D : 0x3C + tc);
    }
}

// Traceback rows. One row = one STEP PAIR (an A-step and the B-step after it); inside a row thread g owns a slot of
// tb_slot_bytes(KP) bytes holding the low bytes of its 4*KP H values: word k = [A slot k, B slot k, A slot KP+k, B slot KP+k],
// i.e. the diagonals (2k, 2k+1, 2KP+2k, 2KP+2k+1) of the thread. A cell's left and up neighbours (the other set, one step
// earlier) and its diagonal predecessor (same set, one step pair earlier) then lie in the same 32-byte sector of the previous
// row — or, for a B cell, left / up in the cell's own sector — so a walk touches one new sector per cell instead of three
// (the former one-row-per-step layout cost the traceback 270 DRAM bytes per path cell, profiles/ncu_tbw_r01_v3.json).
// Slots are padded to a multiple of 8 bytes (KP = 5: 24, KP = 7: 32) so that the stores stay 8- or 16-byte vectors.
__host__ __device__ constexpr int tb_slot_bytes(int kp) { return kp <= 4 ? 16 : (kp <= 6 ? 24 : (kp <= 8 ? 32 : 48)); }
// low bytes of the 2*KP A values and 2*KP B values of one step pair -> this thread's slot of the row
template <int KP>
__device__ __forceinline__ void store_pair(uint8_t* p, const uint32_t (&HA)[KP], const uint32_t (&HB)[KP]) {
    constexpr int NW = tb_slot_bytes(KP) / 4;
    uint32_t w[NW];
    #pragma unroll
    for (int k = 0; k < NW; ++k) w[k] = k < KP ? __byte_perm(HA[k < KP ? k : 0], HB[k < KP ? k : 0], 0x6240) : 0u;
    if (NW % 4 == 0) {          // 16 / 32-byte slots: 16-byte aligned
        #pragma unroll
        for (int k = 0; k < NW; k += 4) __stcs(reinterpret_cast<uint4*>(p) + k / 4, make_uint4(w[k], w[k + 1], w[k + 2], w[k + 3]));
    } else {                    // 24-byte slots: 8-byte aligned
        #pragma unroll
        for (int k = 0; k < NW; k += 2) __stcs(reinterpret_cast<uint2*>(p) + k / 2, make_uint2(w[k], w[k + 1]));
    }
}

constexpr int fwd_max_threads(int kp, bool multi) { return !multi ? 32 : (kp <= 8 ? 512 : 256); }

// Rebased variant (RB = true): the s16x2 fast path for pairs whose scores leave the s16 range (reads beyond ~16 kb).
// Every warp keeps its state RELATIVE to a warp-uniform 32-bit offset `base` (a multiple of 256, so the stored low bytes
// are those of the true H). H of in-band neighbours differs by at most Lip = A_max + dearest 1-base gap,
// so the 128*KP diagonals of a warp span at most 128*KP*Lip, and between two rebases (RB_ITERS iterations = RB_ITERS*2KP
// steps) a cell's value drifts by at most Lip per step: the relative values stay inside +-RB_CLAMP, no 16-bit add ever
// wraps, and the result is bit-identical to 32-bit arithmetic (the host checks the bound for the scoring, plan_pair).
// The zero floor of local alignment becomes max(-base, -RB_CLAMP) and rides as the third operand of the last VIMNMX3
// (the RELU of the plain kernel), so the loop body costs the same. Values that arrive from the neighbouring warp
// (mailbox) are shifted by the difference of the two offsets (4 VIADD.16x2 per step); "outside the band" is a
// neighbour with offset 0. Offsets are exchanged through shared memory at every rebase (one extra barrier per
// RB_ITERS iterations; a cluster barrier when the band spans CTAs). The running maximum is kept in 32 bits per
// (thread, half) and refreshed once per iteration.
constexpr int RB_ITERS = 16;
constexpr int RB_CLAMP = 16000;

// (No minimum-CTAs bound on the single-warp variants: the sliding-window KP = 8 kernel compiles to 157 registers = 12 CTAs per
// SM, profiles/ncu_fwd_r01_v5_small.json; capping it at 128 registers / 16 CTAs per SM was measured 2 % SLOWER on configs 2
// and 4 — the kernel is bound by the ALU pipe, not by latency hiding.)
#ifndef SVP_SINGLE_MIN_CTAS
#define SVP_SINGLE_MIN_CTAS 1
#endif
// The single-warp variants of wave scheduling (WALK = false) are held to 128 registers = 16 CTAs per SM: left alone ptxas takes 110-170
// for scheduling slack; under the bound it finds 86-120 WITHOUT a spill, i.e. 17-19 CTAs per SM.
#ifndef SVP_WAVE_MIN_CTAS
#define SVP_WAVE_MIN_CTAS 16
#endif
// KP = 12: one warp covers 1536 diagonals (the bands of 1025..1536 diagonals that otherwise take three KP = 4 warps with mailboxes:
// 24 % more instructions per diagonal at 73 % instead of 86 % of the ALU pipe, profiles/ncu_fwd_r02_v2.json). 120 state registers.
#ifndef SVP_WIDE_MIN_CTAS
#define SVP_WIDE_MIN_CTAS 1
#endif
// WALK = false: the variant wave scheduling launches — no call to the walk anywhere in the kernel, so its register count is the forward
// pass's own (the walk needs ~130: with it KP = 4 goes from 72 to 124 registers, KP = 5 / 6 sliding windows from 127 to 136 / 142 = 14-15
// instead of 16 CTAs per SM)
template <int KP, bool MULTI, bool TWO, bool CL = false, bool RB = false, bool RING = false, bool WALK = true>
__global__ void __launch_bounds__(fwd_max_threads(KP, MULTI), MULTI ? 1 : (KP > 8 ? SVP_WIDE_MIN_CTAS : (!WALK ? SVP_WAVE_MIN_CTAS : SVP_SINGLE_MIN_CTAS)))
svp_fwd_kernel(const Task* __restrict__ tasks, const uint32_t* __restrict__ order, const uint8_t* __restrict__ seqs,
               const Arena ar, svp_result* __restrict__ results, uint32_t* __restrict__ cigar, const Consts cs) {
    static_assert(!CL || MULTI, "cluster variants are multi-warp");
    extern __shared__ __align__(16) uint8_t smem8[];
    __shared__ unsigned long long s_slice[3];
    const uint32_t crank = CL ? cluster_ctarank() : 0u, csize = CL ? cluster_nctarank() : 1u;
    const uint32_t task_index = order[CL ? blockIdx.x / csize : blockIdx.x];
    const Task tk = tasks[task_index];
    const int gl = threadIdx.x, lane = gl & 31, warp = gl >> 5, nwarp = blockDim.x >> 5;
    const int g = (int)crank * (int)blockDim.x + gl;          // thread index across the cluster: owner of diagonals g*4KP ...
    const int ring = ring_bytes((int)blockDim.x, 2 * KP), rmask = ring - 1;
    const int qbytes = RING ? ring : seq_image_bytes(tk.m), tbytes = RING ? ring : seq_image_bytes(tk.n);
    uint8_t* qimg = smem8;
    uint8_t* timg = smem8 + qbytes;
    // whole images: base + clamped 1-based position; rings: base + (virtual position & mask)
    const uint32_t s_q = (uint32_t)__cvta_generic_to_shared(qimg) + (RING ? 0 : SEQ_PAD - 1);
    const uint32_t s_t = (uint32_t)__cvta_generic_to_shared(timg) + (RING ? 0 : SEQ_PAD - 1);
    // mailboxes (MULTI): A[w] = lane 0 of warp w after an A-step, A[nwarp] = constants;
    //                    B[w+1] = lane 31 of warp w after a B-step, B[0] = constants
    const uint32_t s_mb = (uint32_t)__cvta_generic_to_shared(smem8 + qbytes + tbytes);
    const uint32_t s_cst = s_mb + MB_CST, s_wb = s_mb + MB_WB;

    const uint32_t c_ne1 = pin(cs.ne1), c_noe1 = pin(cs.noe1), c_ne2 = pin(cs.ne2), c_noe2 = pin(cs.noe2);
    const uint32_t c_mat = pin(cs.mat), c_mis = pin(cs.mis), c_neg = pin(NEGV);

    // lowest virtual positions this CTA touches at iteration 0 (its last thread has the smallest I0, its first the smallest J0)
    const int g_first = (int)crank * (int)blockDim.x;
    const int ring_q0 = (tk.r_begin - (tk.band_lo + (g_first + (int)blockDim.x - 1) * 4 * KP)) / 2 - 2 * KP;
    const int ring_t0 = (tk.r_begin + (tk.band_lo + g_first * 4 * KP)) / 2 - 1;
    if (RING) { ring_fill_fast(tk, seqs, qimg, timg, rmask, ring_q0, ring_t0, ring); __syncthreads(); }
    else stage_sequences(tk, seqs, qimg, timg, qbytes, tbytes);
    if (MULTI && gl <= nwarp) { sts_v4(s_mb + MB_A + 16u * gl, c_noe1, c_noe2, c_neg, c_neg); sts_v4(s_mb + MB_B + 16u * gl, c_noe1, c_noe2, c_neg, c_neg);
                                step_link_init<CL>(s_mb, gl); }
    if (MULTI && gl == 0) sts_v4(s_cst, c_noe1, c_noe2, c_neg, c_neg);
    if (MULTI && RB && gl < nwarp + 2) sts_u32(s_wb + 4u * gl, 0u);
    pair_begin<WALK>(ar, tk, crank, s_slice, seqs, results, cigar, cs);
    step_barrier<CL>();
    uint8_t* const rows = pair_rows<CL>(ar, crank, s_slice);
    uint2* const bests = reinterpret_cast<uint2*>(rows + tk.tb_bytes);

    const int n_thr_active = tk.band_w / (4 * KP);
    const bool active = g < n_thr_active;
    const bool lo_edge = (lane == 0);                           // left neighbour: mailbox / constants
    const bool hi_edge = (lane == 31) || (g == n_thr_active - 1);   // up neighbour: mailbox / constants
    const int D0 = tk.band_lo + g * 4 * KP;
    const int I0 = (tk.r_begin - D0) / 2;          // exact: r_begin - D0 is even
    const int J0 = (tk.r_begin + D0) / 2;
    const int qmax = tk.m + 1, tmax = tk.n + 1;

    auto code_at = [&](uint32_t sbase, int pos, int lim) -> uint32_t {   // fp16 code (16 bit) of a 1-based position
        return lds_u8(sbase + (RING ? (uint32_t)(pos & rmask) : (uint32_t)__vimin_s32_relu(pos, lim))) << 8;
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
    uint32_t best = RB ? 0x80008000u : 0u, snap = 0, blk_lo = 0, blk_hi = 0;
    // rebased variant: warp-uniform score offset, 32-bit running maxima, the floor (0 relative to the offset) and the
    // offset differences to the lower / upper neighbour warp as s16x2 constants
    int base = 0, best_lo = 0, best_hi = 0, nb_lo = 0;
    uint32_t c_floor = 0, dLo = 0, dLoF = 0, dUp = 0;
    const int pitch = n_thr_active * tb_slot_bytes(KP);      // bytes per traceback row (one step pair)
    uint8_t* tbp = rows + (size_t)g * tb_slot_bytes(KP);
    const StepLink link = step_link<CL>(s_mb, warp, nwarp, crank, csize);
    // the last in-band thread reads the "outside the band" constants, whatever warp / CTA follows it
    const uint32_t s_upA = (g == n_thr_active - 1) ? s_cst : link.in_up(), s_loB = link.in_lo();
    constexpr int RING_PER = RING_CHUNK / KP;                 // iterations between two window refills (RING_PER * KP positions each)

    #pragma unroll 1
    for (int it = 0; it < tk.n_iter; ++it) {
        if (RING && it && (it % RING_PER) == 0) {                   // the next chunk of both windows (needed from the next iteration on)
            if (MULTI) __syncthreads();                             // the warps of a CTA drift by up to a step per warp: line them up around the refill
            ring_fill_fast(tk, seqs, qimg, timg, rmask, ring_q0 + ring + (it - RING_PER) * KP, ring_t0 + ring + (it - RING_PER) * KP, RING_PER * KP);
            if (MULTI) __syncthreads(); else __syncwarp();
        }
        const uint32_t pit = (KP & 1) ? (uint32_t)(it & 1) : 0u;    // parity of step pair it * KP + u = (u & 1) ^ pit
        #pragma unroll
        for (int u = 0; u < KP; ++u) {
            // =============================== A-step ===========================================
            {
                uint32_t sP1 = __shfl_up_sync(0xffffffffu, P1[KP - 1], 1);
                uint32_t sE1 = __shfl_up_sync(0xffffffffu, E1[KP - 1], 1);
                uint32_t sP2 = 0, sE2 = 0;
                if (TWO) { sP2 = __shfl_up_sync(0xffffffffu, P2[KP - 1], 1); sE2 = __shfl_up_sync(0xffffffffu, E2[KP - 1], 1); }
                // slot 0 is computed last: it alone needs the lower neighbour (shuffle, or the mailbox on lane 0), so the wait for
                // the previous step's barrier sits right before it; its own-thread halves are the values from before this step
                const uint32_t oP1 = P1[KP - 1], oE1 = E1[KP - 1], oP2 = P2[KP - 1], oE2 = E2[KP - 1];
                uint32_t P1m = 0, E1m = 0, P2m = 0, E2m = 0;
                #pragma unroll
                for (int k = KP - 1; k >= 0; --k) {
                    if (k == 0) {
                        uint4 mb = make_uint4(c_noe1, c_noe2, c_neg, c_neg);
                        if (MULTI) { if (it | u) link_wait_lo<CL>(link, lane, (uint32_t)((u & 1) ^ 1) ^ pit); mb = lds_v4(s_loB); }
                        if (RB) {              // the entry was written before the last rebase when this is the first step after it
                            const uint32_t dl = (u == 0) ? dLoF : dLo;
                            mb.x = __vadd2(mb.x, dl); mb.z = __vadd2(mb.z, dl);
                            if (TWO) { mb.y = __vadd2(mb.y, dl); mb.w = __vadd2(mb.w, dl); }
                        }
                        sP1 = lo_edge ? mb.x : sP1; sE1 = lo_edge ? mb.z : sE1;
                        if (TWO) { sP2 = lo_edge ? mb.y : sP2; sE2 = lo_edge ? mb.w : sE2; }
                        P1m = __byte_perm(sP1, oP1, 0x5432); E1m = __byte_perm(sE1, oE1, 0x5432);
                        if (TWO) { P2m = __byte_perm(sP2, oP2, 0x5432); E2m = __byte_perm(sE2, oE2, 0x5432); }
                    }
                    const uint32_t p1l = k ? P1[k ? k - 1 : 0] : P1m, e1l = k ? E1[k ? k - 1 : 0] : E1m;
                    const uint32_t e1 = __viaddmax_s16x2(e1l, c_ne1, p1l);
                    const uint32_t f1 = __viaddmax_s16x2(F1[k], c_ne1, P1[k]);
                    const uint32_t mk = heq_mask(Q[(k - u + KP) % KP], T[(k + u) % KP]);
                    const uint32_t sc = (mk & c_mat) | (~mk & c_mis);
                    const uint32_t h = __vadd2(HA[k], sc);
                    uint32_t H;
                    if (TWO) {
                        const uint32_t p2l = k ? P2[k ? k - 1 : 0] : P2m, e2l = k ? E2[k ? k - 1 : 0] : E2m;
                        const uint32_t e2 = __viaddmax_s16x2(e2l, c_ne2, p2l);
                        const uint32_t f2 = __viaddmax_s16x2(F2[k], c_ne2, P2[k]);
                        if (RB) H = __vimax3_s16x2(__vimax3_s16x2(__viaddmax_s16x2(HA[k], sc, e1), f1, e2), f2, c_floor);
                        else H = __vimax3_s16x2_relu(__vimax3_s16x2(h, e1, f1), e2, f2);
                        E2[k] = e2; F2[k] = f2; P2[k] = __vadd2(H, c_noe2);
                    } else {
                        if (RB) H = __vimax3_s16x2(__viaddmax_s16x2(HA[k], sc, e1), f1, c_floor);
                        else H = __vimax3_s16x2_relu(h, e1, f1);
                    }
                    HA[k] = H; E1[k] = e1; F1[k] = f1; P1[k] = __vadd2(H, c_noe1);
                }
                #pragma unroll
                for (int k = 0; k + 1 < KP; k += 2) best = __vimax3_s16x2(best, HA[k], HA[k + 1]);
                if (KP & 1) best = __vmaxs2(best, HA[KP - 1]);
                if (MULTI && lane == 0) mailbox_store_A<CL>(link, P1[0], P2[0], F1[0], F2[0]);
                // T advances by one base: logical T[k] <- T[k+1]; the new base enters slot 2KP-1
                const uint32_t nb = lds_u8(s_t + (RING ? (uint32_t)(tpos & rmask) : (uint32_t)__vimin_s32_relu(tpos, tmax)));
                T[u % KP] = __byte_perm(T[u % KP], nb, 0x4232);
                ++tpos;
            }
            // =============================== B-step ===========================================
            {
                uint32_t sP1 = __shfl_down_sync(0xffffffffu, P1[0], 1);
                uint32_t sF1 = __shfl_down_sync(0xffffffffu, F1[0], 1);
                uint32_t sP2 = 0, sF2 = 0;
                if (TWO) { sP2 = __shfl_down_sync(0xffffffffu, P2[0], 1); sF2 = __shfl_down_sync(0xffffffffu, F2[0], 1); }
                const uint32_t oP1 = P1[0], oF1 = F1[0], oP2 = P2[0], oF2 = F2[0];
                uint32_t P1p = 0, F1p = 0, P2p = 0, F2p = 0;
                #pragma unroll
                for (int k = 0; k < KP; ++k) {
                    if (k == KP - 1) {         // the last slot alone needs the upper neighbour
                        uint4 mb = make_uint4(c_noe1, c_noe2, c_neg, c_neg);
                        if (MULTI) { link_wait_up<CL>(link, lane, (uint32_t)(u & 1) ^ pit); mb = lds_v4(s_upA); }
                        if (RB) {
                            mb.x = __vadd2(mb.x, dUp); mb.z = __vadd2(mb.z, dUp);
                            if (TWO) { mb.y = __vadd2(mb.y, dUp); mb.w = __vadd2(mb.w, dUp); }
                        }
                        sP1 = hi_edge ? mb.x : sP1; sF1 = hi_edge ? mb.z : sF1;
                        if (TWO) { sP2 = hi_edge ? mb.y : sP2; sF2 = hi_edge ? mb.w : sF2; }
                        P1p = __byte_perm(oP1, sP1, 0x5432); F1p = __byte_perm(oF1, sF1, 0x5432);
                        if (TWO) { P2p = __byte_perm(oP2, sP2, 0x5432); F2p = __byte_perm(oF2, sF2, 0x5432); }
                    }
                    const uint32_t p1u = (k < KP - 1) ? P1[(k + 1) % KP] : P1p, f1u = (k < KP - 1) ? F1[(k + 1) % KP] : F1p;
                    const uint32_t e1 = __viaddmax_s16x2(E1[k], c_ne1, P1[k]);
                    const uint32_t f1 = __viaddmax_s16x2(f1u, c_ne1, p1u);
                    const uint32_t mk = heq_mask(Q[(k - u + KP) % KP], T[(k + u + 1) % KP]);
                    const uint32_t sc = (mk & c_mat) | (~mk & c_mis);
                    const uint32_t h = __vadd2(HB[k], sc);
                    uint32_t H;
                    if (TWO) {
                        const uint32_t p2u = (k < KP - 1) ? P2[(k + 1) % KP] : P2p, f2u = (k < KP - 1) ? F2[(k + 1) % KP] : F2p;
                        const uint32_t e2 = __viaddmax_s16x2(E2[k], c_ne2, P2[k]);
                        const uint32_t f2 = __viaddmax_s16x2(f2u, c_ne2, p2u);
                        if (RB) H = __vimax3_s16x2(__vimax3_s16x2(__viaddmax_s16x2(HB[k], sc, e1), f1, e2), f2, c_floor);
                        else H = __vimax3_s16x2_relu(__vimax3_s16x2(h, e1, f1), e2, f2);
                        E2[k] = e2; F2[k] = f2; P2[k] = __vadd2(H, c_noe2);
                    } else {
                        if (RB) H = __vimax3_s16x2(__viaddmax_s16x2(HB[k], sc, e1), f1, c_floor);
                        else H = __vimax3_s16x2_relu(h, e1, f1);
                    }
                    HB[k] = H; E1[k] = e1; F1[k] = f1; P1[k] = __vadd2(H, c_noe1);
                }
                #pragma unroll
                for (int k = 0; k + 1 < KP; k += 2) best = __vimax3_s16x2(best, HB[k], HB[k + 1]);
                if (KP & 1) best = __vmaxs2(best, HB[KP - 1]);
                if (MULTI && lane == 31) mailbox_store_B<CL>(link, P1[KP - 1], P2[KP - 1], E1[KP - 1], E2[KP - 1]);
                if (active) store_pair<KP>(tbp, HA, HB);
                tbp += pitch;
                // Q advances by one base: logical Q[k] <- Q[k-1]; the new base enters slot 0
                const uint32_t nb = lds_u8(s_q + (RING ? (uint32_t)(qpos & rmask) : (uint32_t)__vimin_s32_relu(qpos, qmax)));
                Q[(KP - 1 - u) % KP] = __byte_perm(nb, Q[(KP - 1 - u) % KP], 0x5401);
                ++qpos;
            }
        }
        // which iteration first reached the running maximum, per half (end-cell search)
        if (!RB) {
            const uint32_t diff = best ^ snap;
            if (diff & 0xFFFFu) blk_lo = (uint32_t)it;
            if (diff >> 16) blk_hi = (uint32_t)it;
            snap = best;
        } else {
            const int lo = (int)(int16_t)(best & 0xFFFFu) + base, hi = ((int)best >> 16) + base;
            if (lo > best_lo) { best_lo = lo; blk_lo = (uint32_t)it; }
            if (hi > best_hi) { best_hi = hi; blk_hi = (uint32_t)it; }
            best = 0x80008000u;
            dLoF = dLo;
            if (((it + 1) & (RB_ITERS - 1)) == 0 && it + 1 < tk.n_iter) {
                // ---- rebase: the new offset is lane 0's current H of slot 0, rounded down to a multiple of 256 ----
                const int h0 = __shfl_sync(0xffffffffu, (int)(int16_t)(HA[0] & 0xFFFFu), 0);
                const int delta = h0 & ~255;
                if (delta) {
                    const uint32_t nd = pk(-delta);
                    #pragma unroll
                    for (int k = 0; k < KP; ++k) {
                        HA[k] = __vadd2(HA[k], nd); HB[k] = __vadd2(HB[k], nd); P1[k] = __vadd2(P1[k], nd);
                        E1[k] = __vadd2(E1[k], nd); F1[k] = __vadd2(F1[k], nd);
                        if (TWO) { P2[k] = __vadd2(P2[k], nd); E2[k] = __vadd2(E2[k], nd); F2[k] = __vadd2(F2[k], nd); }
                    }
                    base += delta;
                    c_floor = pk(max(-base, -RB_CLAMP));
                }
                const int nb_lo_old = nb_lo;
                int nb_up = 0;
                if (MULTI) {
                    if (lane == 0) {
                        sts_u32(s_wb + 4u * (uint32_t)(warp + 1), (uint32_t)base);
                        if (CL && link.rlo) st_cluster_u32(link.rem_wb_lo, (uint32_t)base);
                        if (CL && link.rup) st_cluster_u32(link.rem_wb_up, (uint32_t)base);
                    }
                    step_barrier<CL>();
                    nb_lo = (int)lds_u32(s_wb + 4u * (uint32_t)warp);
                    nb_up = (int)lds_u32(s_wb + 4u * (uint32_t)(warp + 2));
                }
                if (g == n_thr_active - 1) nb_up = 0;
                dLoF = pk(max(min(nb_lo_old - base, RB_CLAMP), -RB_CLAMP));
                dLo = pk(max(min(nb_lo - base, RB_CLAMP), -RB_CLAMP));
                dUp = pk(max(min(nb_up - base, RB_CLAMP), -RB_CLAMP));
            }
        }
    }
    if (active) {
        if (RB) {
            bests[2 * g]     = make_uint2((uint32_t)best_lo, blk_lo);
            bests[2 * g + 1] = make_uint2((uint32_t)best_hi, blk_hi);
        } else {
            bests[2 * g]     = make_uint2((uint32_t)(int)(int16_t)(best & 0xFFFFu), blk_lo);
            bests[2 * g + 1] = make_uint2((uint32_t)(int)(int16_t)(best >> 16), blk_hi);
        }
    }
    // rows and best records of every warp (CTA of the cluster) are complete; no CTA leaves while a neighbour may still store into its mailbox
    step_barrier<CL>();
    // the walk's corridor takes over the CTA's dynamic shared memory (sequence images / windows and mailboxes are dead now)
    if (WALK && warp == 0 && crank == 0) pair_end(ar, tk, task_index, seqs, results, cigar, cs, s_slice, ar.cor_smem ? (uint32_t)__cvta_generic_to_shared(smem8) : 0u);
}

// ---------------------------------------------------------------------------------------------
// traceback
// ---------------------------------------------------------------------------------------------
// Byte layout of the traceback rows (store_pair): row = step pair (s >> 1), s = (i + j) - r_begin; pitch = threads * slot bytes;
// inside a row thread g = x / (4*kp) owns one slot; the byte of its diagonal x' = x - g*4kp (x = (j - i) - band_lo, parity of x' =
// parity of s) sits at 4*k + 2*half + (x' & 1) with c = x' >> 1, half = c >= kp, k = c - half*kp. kp is a runtime value (4..8).
// loads of traceback rows / best records: through L2 only (CG) or cached (see svp_traceback_warp's declaration)
template <bool CG> __device__ __forceinline__ uint32_t ldrow(const uint8_t* p) { return CG ? (uint32_t)__ldcg(p) : (uint32_t)*p; }
template <bool CG> __device__ __forceinline__ uint2 ldbest(const uint2* p) { return CG ? __ldcg(p) : *p; }
constexpr int COR_W = 96, COR_RING = 128, COR_MARGIN = 16, COR_BYTES = COR_W * COR_RING;       // 12 KB per walking warp
template <bool CG>
struct TbView {
    const uint8_t* base;      // the pair's slice of the arena
    const uint32_t* seqs32;   // packed sequences (global)
    int band_lo, band_w, r_begin, pitch, m, n, kp, slot;
    uint32_t inv;             // ceil(2^32 / (4*kp)): x / (4*kp) = umulhi(x, inv), exact for x < 2^27
    uint64_t q_word, t_word;  // word offsets of the packed sequences
    const uint32_t* q_mask;   // masks of letters outside ACGT (nullptr: none)
    const uint32_t* t_mask;
    bool rc;

    // is the base at 1-based position i of the ORIENTED query / j of the target a letter outside ACGT
    __device__ __forceinline__ bool qn(int i) const {
        if (!q_mask) return false;
        const unsigned p = rc ? (unsigned)(m - i) : (unsigned)(i - 1);
        return (__ldg(q_mask + (p >> 5)) >> (p & 31u)) & 1u;
    }
    __device__ __forceinline__ bool tn(int j) const {
        if (!t_mask) return false;
        const unsigned p = (unsigned)(j - 1);
        return (__ldg(t_mask + (p >> 5)) >> (p & 31u)) & 1u;
    }
    __device__ __forceinline__ void set_geometry(int band_w_, int kp_) {
        band_w = band_w_; kp = kp_; slot = tb_slot_bytes(kp_);
        pitch = (band_w_ / (4 * kp_)) * slot;
        inv = (uint32_t)(((1ull << 32) + 4u * (uint32_t)kp_ - 1u) / (4u * (uint32_t)kp_));
    }
    // offset of diagonal x (band-relative) inside a row
    __device__ __forceinline__ int col(int x) const {
        const int g = (int)__umulhi((uint32_t)x, inv);
        const int xr = x - g * 4 * kp;
        const int c = xr >> 1;
        const int half = c >= kp ? 1 : 0, k = c - half * kp;
        return g * slot + k * 4 + half * 2 + (xr & 1);
    }
    // stored byte of step s, band-relative diagonal x (x and s have the same parity)
    __device__ __forceinline__ const uint8_t* addr_sx(int s, int x) const { return base + (ptrdiff_t)(s >> 1) * pitch + col(x); }
    // address of the stored byte of cell (i, j) (callers make sure the cell is real and inside the band)
    __device__ __forceinline__ const uint8_t* addr(int i, int j) const { return addr_sx((i + j) - r_begin, (j - i) - band_lo); }
    // ---- corridor (svp_traceback_warp) ---------------------------------------------------------------------------------------
    // A walk is a chain of dependent byte loads, ~1 us each from DRAM. But it only ever moves DOWN the rows (a predecessor lies on
    // an earlier anti-diagonal) and drifts slowly across them (one diagonal per indel), so the bytes it will need lie in a narrow
    // corridor below the current cell. The warp keeps that corridor in shared memory: a ring of COR_RING rows x COR_W bytes (a
    // sector-aligned window of every row around the path's column), filled 32 rows at a time with cp.async (LDGSTS, through L2,
    // no registers) two blocks AHEAD of the walk — so a round's loads are shared-memory loads (~30 cycles) and the DRAM latency
    // of the refills hides behind the rounds in between. A cell outside the window (the far candidates of a long gap, a neighbour
    // beyond the window's edge, a block still in flight) is loaded from global memory as before; when the path drifts to within
    // COR_MARGIN bytes of the window's edge the window is re-centred (one synchronous refill).
    uint32_t s_cor;           // shared address of the ring, 0: no corridor
    int cor_wc, cor_w;        // first byte column of the window, its width (<= COR_W)
    int cor_lo;               // lowest 32-row block issued so far (blocks are issued downwards); cor_lo + 3 is the highest still intact
    int cor_ready;            // lowest block known complete
    int cor_top;              // block of the row the corridor was (re)started at
    int n_rows;               // rows of the pair
    int cor_row_lo; unsigned cor_row_span;   // rows [cor_row_lo, cor_row_lo + cor_row_span] are complete in the ring
    __device__ __forceinline__ void cor_issue(int blk, int lane) {      // one block = 32 rows, lane = row, one commit group
        const int row = blk * 32 + lane;
        if (row < n_rows) {
            const uint8_t* src = base + (ptrdiff_t)row * pitch + cor_wc;
            const uint32_t dst = s_cor + (uint32_t)(row & (COR_RING - 1)) * COR_W;
            for (int b = 0; b < cor_w; b += 16)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst + (uint32_t)b), "l"(src + b) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    // before a round at cell (i, j): the blocks of the rows this round can touch (the current one and the one below) are complete,
    // the two below them are on their way. Warp-uniform.
    __device__ __forceinline__ void cor_advance(int i, int j, int lane) {
        if (!s_cor) return;
        const int s = (i + j) - r_begin;
        if (s < 0) return;
        const int row = s >> 1, blk = row >> 5;
        const int off = col((j - i) - band_lo) - cor_wc;
        const bool recentre = cor_top < 0 || (off < COR_MARGIN && cor_wc > 0) || (off >= cor_w - COR_MARGIN && cor_wc + cor_w < pitch);
        if (!recentre && max(blk - 1, 0) >= cor_ready && cor_lo <= max(blk - 3, 0)) return;      // nothing to do (the common case)
        if (recentre) {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncwarp();
            const int c = col((j - i) - band_lo);
            cor_wc = min(max((c & ~31) - 32, 0), max(pitch - COR_W, 0));
            cor_top = blk; cor_lo = blk + 1;
        }
        if (cor_lo > blk + 1) cor_lo = blk + 1;                        // the walk jumped over whole blocks (a long gap)
        while (cor_lo > max(blk - 3, 0)) cor_issue(--cor_lo, lane);
        // complete: everything but the blocks below blk - 1 (at most two groups)
        const int in_flight = min(max(blk - 1 - cor_lo, 0), 2);
        if (in_flight == 0) asm volatile("cp.async.wait_group 0;" ::: "memory");
        else if (in_flight == 1) asm volatile("cp.async.wait_group 1;" ::: "memory");
        else asm volatile("cp.async.wait_group 2;" ::: "memory");
        __syncwarp();
        cor_ready = cor_lo + in_flight;
        cor_row_lo = cor_ready * 32;
        cor_row_span = (unsigned)(min(cor_top, cor_lo + 3) * 32 + 31 - cor_row_lo);
    }
    // stored byte at (row, byte column): from the corridor when it holds it, else from global memory
    __device__ __forceinline__ uint32_t ld(int row, int c) const {
        const unsigned off = (unsigned)(c - cor_wc);
        if (off < (unsigned)cor_w && (unsigned)(row - cor_row_lo) <= cor_row_span)
            return lds_u8(s_cor + (uint32_t)(row & (COR_RING - 1)) * COR_W + off);
        return ldrow<CG>(base + (ptrdiff_t)row * pitch + c);
    }
    // stored byte of H(i, j); 0 for cells before the matrix or before the first stored row (callers check the band)
    __device__ __forceinline__ uint32_t lb(int i, int j) const {
        if (i < 1 || j < 1 || (i + j) < r_begin) return 0;
        return ld(((i + j) - r_begin) >> 1, col((j - i) - band_lo));
    }
    // single bases (codes 0..3) at 1-based positions of the ORIENTED query / the target
    __device__ __forceinline__ int qbase(int i) const {
        const unsigned p = rc ? (unsigned)(m - i) : (unsigned)(i - 1);
        const int c = (int)((__ldg(seqs32 + q_word + (p >> 4)) >> ((p & 15u) * 2u)) & 3u);
        return rc ? 3 - c : c;
    }
    __device__ __forceinline__ int tbase(int j) const {
        const unsigned p = (unsigned)(j - 1);
        return (int)((__ldg(seqs32 + t_word + (p >> 4)) >> ((p & 15u) * 2u)) & 3u);
    }
    // bases at 1-based positions p, p+1, ..., base t in bits 2t (positions beyond len: unspecified)
    __device__ __forceinline__ uint32_t window_asc(uint64_t word0, int len, int p) const {
        const int p0 = max(p - 1, 0);
        const int last = (len - 1) >> 4;
        const uint32_t w0 = __ldg(seqs32 + word0 + min(p0 >> 4, last));
        const uint32_t w1 = __ldg(seqs32 + word0 + min((p0 >> 4) + 1, last));
        return __funnelshift_r(w0, w1, (p0 & 15) * 2);
    }
    // bases at 1-based positions p, p-1, ..., p-n+1: base of cell k (position p-k) in bits 2k; n <= 16
    __device__ __forceinline__ uint32_t window_desc(uint64_t word0, int len, int p, int n) const {
        const int start = max(p - n + 1, 1);
        const uint32_t asc = window_asc(word0, len, start);
        uint32_t r = __brev(asc);                                   // group t -> bits 30-2t, bits swapped inside
        r = ((r & 0x55555555u) << 1) | ((r >> 1) & 0x55555555u);    // undo the swap inside each 2-bit group
        return r >> (30 - 2 * (p - start));
    }
    // bases of cells 0..n-1 of a diagonal run that starts at oriented query position i (cell k = i - k)
    __device__ __forceinline__ uint32_t qwindow(int i, int n) const {
        if (!rc) return window_desc(q_word, m, i, n);
        return ~window_asc(q_word, m, m + 1 - i);                   // oriented P = complement of original m+1-P
    }
    __device__ __forceinline__ uint32_t twindow(int j, int n) const { return window_desc(t_word, this->n, j, n); }
};

__device__ __forceinline__ int gap_cost2(int o1, int e1, int o2, int e2, int two, int k) {
    int c = o1 + k * e1;
    if (two) c = min(c, o2 + k * e2);
    return c;
}

__device__ __forceinline__ int sext8(uint32_t x) { return (int)(int8_t)(x & 0xFFu); }

// One THREAD per pair, 32 independent walks per warp. A walk is a chain of dependent byte loads, so what
// matters is the number of memory round trips and keeping the 32 lanes on ONE instruction stream:
//   * a round = every lane issues all its loads (the diagonal/left/up predecessor bytes of the next S
//     cells along its current diagonal, or — while it searches a gap longer than one — the next NSCAN
//     cells of its row and column), then the loaded cells are processed by straight-line predicated
//     code without data-dependent branches;
//   * the Z-trim start (lowest H of the walk, §3.5) is only tracked as a position; the rare walk that is
//     actually cut is replayed up to that position (same deterministic path) to rebuild CIGAR and counts.
template <bool CG>
__device__ __noinline__ void svp_traceback_thread(const Task& tk, const uint8_t* __restrict__ rows, const uint2* __restrict__ bests,
                                                  const uint8_t* __restrict__ seqs, svp_result* __restrict__ results,
                                                  uint32_t* __restrict__ cigar, const Consts& cs) {
    svp_result res;
    res.score = 0; res.status = 0; res.q_beg = res.q_end = res.t_beg = res.t_end = 0;
    res.n_match = res.n_mismatch = res.n_ins_ops = res.n_ins_bp = res.n_del_ops = res.n_del_bp = 0;
    res.cigar_off = tk.cigar_off; res.cigar_len = 0; res.sv_dist = 0.0; res.cells = tk.cells;
    const int KP = (int)((tk.flags & TASK_KP_MASK) >> TASK_KP_SHIFT);
    TbView<CG> v;
    v.base = rows; v.seqs32 = reinterpret_cast<const uint32_t*>(seqs);
    v.band_lo = tk.band_lo; v.r_begin = tk.r_begin; v.m = tk.m; v.n = tk.n;
    v.set_geometry(tk.band_w, KP);
    v.q_word = tk.q_off >> 2; v.t_word = tk.t_off >> 2; v.rc = (tk.flags & SVP_PAIR_REVCOMP) != 0;
    v.q_mask = seq_mask(seqs, tk.q_off, tk.m, (tk.flags & TASK_QN) != 0);
    v.t_mask = seq_mask(seqs, tk.t_off, tk.n, (tk.flags & TASK_TN) != 0);
    const bool any_mask = (tk.flags & (TASK_QN | TASK_TN)) != 0;
    const int d_hi = tk.band_lo + tk.band_w - 1;

    // ---- 1. end cell: maximum over the per-thread-half records, earliest iteration ------------
    const int n_ent = (tk.band_w / (4 * KP)) * 2;
    int M = 0; uint32_t blk = 0xFFFFFFFFu;
    for (int e = 0; e < n_ent; ++e) {
        const uint2 b = ldbest<CG>(bests + e);
        const int val = (int)b.x;
        if (val > M || (val == M && b.y < blk)) { M = val; blk = b.y; }
    }
    if (M <= 0) {
        res.status = SVP_ST_NOHIT;
        results[tk.pair_index] = res;
        return;
    }
    if (!(tk.flags & (TASK_S32 | TASK_REBASE)) && M >= 32767 - cs.a_max) res.status |= SVP_ST_SCORE_OVF;
    // every (thread, half) block that reached M in iteration blk: rebuild its 2KP x 2KP cells from the low bytes
    int end_s = 0x7FFFFFFF, end_x = -1;
    for (int e = 0; e < n_ent; ++e) {
        const uint2 b = ldbest<CG>(bests + e);
        if ((int)b.x != M || b.y != blk) continue;
        const int xbase = (e >> 1) * 4 * KP + (e & 1) * 2 * KP;      // first diagonal (band-relative) of the half
        const int s0 = (int)blk * 2 * KP;
        // cells id = 0 .. 2*KP*KP-1: step s0 + id/KP, diagonal xbase + 2*(id%KP) + (step odd); chained by neighbours
        int rel = 0, rel_row0 = 0, relmax = -(1 << 30), cand_s = 0x7FFFFFFF, cand_x = -1;
        uint32_t bprev = 0, b_row0 = 0;
        for (int id = 0; id < 2 * KP * KP; ++id) {
            const int s = s0 + id / KP, x = xbase + 2 * (id % KP) + (s & 1);
            const uint32_t bb = ldrow<CG>(v.addr_sx(s, x));
            if (id == 0) rel = 0;
            else if (id % KP == 0) rel = rel_row0 + sext8(bb - b_row0);
            else rel = rel + sext8(bb - bprev);
            if (id % KP == 0) { rel_row0 = rel; b_row0 = bb; }
            bprev = bb;
            if (rel > relmax) { relmax = rel; cand_s = s; cand_x = x; }
            else if (rel == relmax && (s < cand_s || (s == cand_s && x > cand_x))) { cand_s = s; cand_x = x; }
        }
        if (cand_s < end_s || (cand_s == end_s && cand_x > end_x)) { end_s = cand_s; end_x = cand_x; }
    }
    const int best_i = ((tk.r_begin + end_s) - (tk.band_lo + end_x)) / 2;
    const int best_j = ((tk.r_begin + end_s) + (tk.band_lo + end_x)) / 2;

    // ---- 2. walk back --------------------------------------------------------------------------
    constexpr int S = 8, NSCAN = 6;
    uint32_t* slot = cigar + tk.cigar_off;
    const ptrdiff_t two_rows = (ptrdiff_t)v.pitch;            // two steps = one row (step pair)
    const int cost1d = gap_cost2(cs.d_open1, cs.d_ext1, cs.d_open2, cs.d_ext2, cs.two_piece, 1);
    const int cost1i = gap_cost2(cs.i_open1, cs.i_ext1, cs.i_open2, cs.i_ext2, cs.two_piece, 1);
    const int zdrop = cs.zdrop > 0 ? cs.zdrop : 0x3FFFFFFF;

    int i, j, H, wpos, low, low_i, low_j, mrun, scan_k, Hrow, Hcol;
    uint32_t brow, bcol, n_match, n_mis, n_io, n_ib, n_do, n_db;
    bool ovf, bad = false;
    int stop_i = -1, stop_j = -1;            // replay target (cut walk): stop when the walk reaches this cell
    bool cut = false; int cut_low = 0;

#define SVP_PUSH(op, len) do { if (wpos > 0) { --wpos; slot[wpos] = ((uint32_t)(len) << 4) | (op); } else ovf = true; } while (0)

    for (int pass = 0; pass < 2; ++pass) {
        i = best_i; j = best_j; H = M; wpos = (int)tk.cigar_cap; ovf = false;
        low = M + 1; low_i = i; low_j = j; mrun = 0; scan_k = 0; Hrow = Hcol = 0; brow = bcol = 0;
        n_match = n_mis = n_io = n_ib = n_do = n_db = 0;
        bool done = false;
        while (!done) {
            const int d = j - i;
            const bool in_scan = scan_k != 0;
            // ---------------- loads of the round, all issued before any is used -----------------
            const bool del_band = (d - 1 >= tk.band_lo), ins_band = (d + 1 <= d_hi);
            const uint8_t* pd = v.addr(i - 1, j - 1);            // diagonal predecessor of cell 0 (stride: 2 rows per cell)
            const uint8_t* pl = v.addr(i, j - 1);                // left neighbour of cell 0
            const uint8_t* pu = v.addr(i - 1, j);                // up neighbour of cell 0
            uint32_t bd[S], bl[S], bu[S];
            #pragma unroll
            for (int k = 0; k < S; ++k) {
                const bool in_q = (i - k - 1 >= 1), in_t = (j - k - 1 >= 1);
                // (every load is issued — from a harmless address when its cell does not exist — and masked afterwards: a load under
                // a condition becomes a branch around a ld.global.cg, and 24 of those in a row are 24 serial round trips)
                const bool cd = !in_scan && in_q && in_t, cl = !in_scan && i - k >= 1 && in_t && del_band, cu = !in_scan && in_q && j - k >= 1 && ins_band;
                const uint32_t xd = ldrow<CG>(cd ? pd - k * two_rows : v.base), xl = ldrow<CG>(cl ? pl - k * two_rows : v.base), xu = ldrow<CG>(cu ? pu - k * two_rows : v.base);
                bd[k] = cd ? xd : 0u; bl[k] = cl ? xl : 0u; bu[k] = cu ? xu : 0u;
            }
            uint32_t br[NSCAN], bc[NSCAN]; bool okr[NSCAN], okc[NSCAN];
            #pragma unroll
            for (int t = 0; t < NSCAN; ++t) {
                const int k = scan_k + t;
                okr[t] = in_scan && (j - k >= 1) && (d - k >= tk.band_lo);
                okc[t] = in_scan && (i - k >= 1) && (d + k <= d_hi);
                const uint32_t xr = ldrow<CG>(okr[t] ? v.addr(i, j - k) : v.base), xc = ldrow<CG>(okc[t] ? v.addr(i - k, j) : v.base);
                br[t] = okr[t] ? xr : 0u; bc[t] = okc[t] ? xc : 0u;
            }
            const uint32_t qwin = v.qwindow(i, S), twin = v.twindow(j, S);   // 2 bits per cell along the diagonal
            if (!in_scan && i - S - 1 >= 1 && j - S - 1 >= 1) {              // L2 prefetch for the next round
                #pragma unroll
                for (int k = S; k < 2 * S; k += 2) {
                    const uint8_t* pk = pd - k * two_rows;      // the sector holding a cell's diagonal, left and up predecessors
                    if (pk - two_rows >= v.base) {
                        asm volatile("prefetch.global.L2 [%0];" :: "l"(pk));
                        asm volatile("prefetch.global.L2 [%0];" :: "l"(pk - two_rows));
                    }
                }
            }
            // ---------------- gap scan (lanes in the scan state) ------------------------------------
            if (in_scan) {
                bool found = false, dead = false;
                #pragma unroll
                for (int t = 0; t < NSCAN; ++t) {
                    const int k = scan_k + t;
                    const int costd = gap_cost2(cs.d_open1, cs.d_ext1, cs.d_open2, cs.d_ext2, cs.two_piece, k);
                    const int costi = gap_cost2(cs.i_open1, cs.i_ext1, cs.i_open2, cs.i_ext2, cs.two_piece, k);
                    if (!found && !dead) {
                        if (!okr[t] && !okc[t]) dead = true;
                        if (okr[t]) {
                            Hrow += sext8(br[t] - brow); brow = br[t];
                            if (H == Hrow - costd) { SVP_PUSH(SVP_CIGAR_D, k); ++n_do; n_db += k; j -= k; H = Hrow; found = true; }
                        }
                        if (!found && okc[t]) {
                            Hcol += sext8(bc[t] - bcol); bcol = bc[t];
                            if (H == Hcol - costi) { SVP_PUSH(SVP_CIGAR_I, k); ++n_io; n_ib += k; i -= k; H = Hcol; found = true; }
                        }
                    }
                }
                if (found) scan_k = 0;
                else if (dead) { bad = true; done = true; }
                else scan_k += NSCAN;
                continue;
            }
            // ---------------- H state: up to S cells along the diagonal, predicated -----------------
            bool alive = true, gap_here = false;
            uint32_t gl = 0, gu = 0;
            #pragma unroll
            for (int k = 0; k < S; ++k) {
                // stop conditions of the current cell
                int dd = d - (low_j - low_i); dd = dd < 0 ? -dd : dd;
                const bool stop = alive && (H == 0 || i < 1 || j < 1 || (i == stop_i && j == stop_j) ||
                                            (pass == 0 && H >= low && H - low > zdrop + cs.zdrop_ext * dd));
                if (stop) { if (pass == 0 && H != 0 && i >= 1 && j >= 1) cut = true; done = true; alive = false; }
                if (alive && H < low) { low = H; low_i = i; low_j = j; }
                const uint32_t qc = (qwin >> (2 * k)) & 3u, tc = (twin >> (2 * k)) & 3u;
                const bool other = any_mask && i >= 1 && j >= 1 && (v.qn(i) || v.tn(j));       // a letter outside ACGT: lowest score, never a match
                const bool same = !other && qc == tc;
                const int sub = other ? cs.n_score : (int)(int8_t)((cs.tab[qc] >> (8u * tc)) & 0xFFu);
                const int Hd = (i - 1 >= 1 && j - 1 >= 1) ? H + sext8(bd[k] - (uint32_t)H) : 0;
                const bool isdiag = alive && (H == Hd + sub);
                if (alive && !isdiag) { gap_here = true; gl = bl[k]; gu = bu[k]; alive = false; }
                if (isdiag) { ++mrun; n_match += same ? 1u : 0u; n_mis += same ? 0u : 1u; --i; --j; H = Hd; }
            }
            if (done) continue;
            if (gap_here) {
                // the M run since the previous gap, then a gap of length 1 (deletion first) or the scan state
                if (mrun) { SVP_PUSH(SVP_CIGAR_M, mrun); mrun = 0; }
                const bool del_ok = (j - 1 >= 1) && del_band, ins_ok = (i - 1 >= 1) && ins_band;
                const int Hl = H + sext8(gl - (uint32_t)H), Hu = H + sext8(gu - (uint32_t)H);
                if (!del_ok && !ins_ok) { bad = true; done = true; }
                else if (del_ok && H == Hl - cost1d) { SVP_PUSH(SVP_CIGAR_D, 1); ++n_do; ++n_db; --j; H = Hl; }
                else if (ins_ok && H == Hu - cost1i) { SVP_PUSH(SVP_CIGAR_I, 1); ++n_io; ++n_ib; --i; H = Hu; }
                else { scan_k = 2; Hrow = Hl; brow = gl; Hcol = Hu; bcol = gu; }
            }
        }
        if (mrun) { SVP_PUSH(SVP_CIGAR_M, mrun); mrun = 0; }
        if (!cut || pass == 1) break;
        stop_i = low_i; stop_j = low_j; cut_low = low;   // replay the (deterministic) walk up to the lowest point
    }
#undef SVP_PUSH
    if (bad) res.status |= SVP_ST_SCORE_OVF;
    res.score = M - (cut ? cut_low : 0);
    res.q_beg = i; res.q_end = best_i; res.t_beg = j; res.t_end = best_j;
    res.n_match = n_match; res.n_mismatch = n_mis; res.n_ins_ops = n_io; res.n_ins_bp = n_ib; res.n_del_ops = n_do; res.n_del_bp = n_db;
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
                const double sz = (double)len;
                dist += (1.0 - exp(-pow(sz / 30.0, 1.5))) * sz;
            }
        }
        res.sv_dist = dist;
    }
    results[tk.pair_index] = res;
}

// ---------------------------------------------------------------------------------------------
// Traceback, run by ONE WARP right after the forward pass of its pair, inside the same CTA (warp 0 of the pair's first CTA):
// the rows it reads are the ones this CTA (cluster) just wrote, and the pair's slice of the arena is free again as soon as the
// walk is done — there are no waves and no second kernel. Every round lane l fetches the diagonal predecessor byte, the left / up
// neighbour bytes and the base pair of cell (i-l, j-l), so 32 cells of a diagonal run cost ONE memory round trip; the run is then
// resolved by all lanes at once (prefix sums of the substitution scores). Gap search: 32 candidates of the row and of the column
// per round. The Z-trim start is kept as a snapshot of the walk at its lowest H.
// All loads of rows / best records are ld.global.cg: they were written by this grid (by other SMs when the band spans a cluster),
// and an arena slice is reused by later CTAs, so L1 must not serve them.
// ---------------------------------------------------------------------------------------------
// s_cor: shared address of COR_BYTES bytes (16-byte aligned) this warp may use for its corridor, or 0.
// LPP = lanes per pair: 32 = the whole warp walks one pair. A diagonal run between two indels is ~7 cells long on ONT-like reads
// (13 % indel columns between two reads), so a round of a whole warp uses a quarter of its lanes, and the walk is bound by its
// instruction count (~450 000 warp instructions per pair of 8 kb reads, profiles/ncu_tbw_r02.json), not by memory latency. With
// LPP = 8 / 16 a warp walks 4 / 2 pairs at once, one per group of LPP lanes: every cross-lane operation stays inside the group
// (shuffles with width LPP, ballots and reductions on the group's mask), so groups may diverge freely; `lane` below is the lane
// inside the group. No corridor unless LPP = 32.
template <bool CG, int LPP>
__device__ __noinline__ void svp_traceback_warp(const Task& tk, const uint8_t* __restrict__ rows, const uint2* __restrict__ bests,
                                                const uint8_t* __restrict__ seqs, svp_result* __restrict__ results,
                                                uint32_t* __restrict__ cigar, const Consts& cs, uint32_t s_cor) {
    static_assert(LPP == 8 || LPP == 16 || LPP == 32, "lanes per pair");
    const int lane = threadIdx.x & (LPP - 1);
    const unsigned gsh = (unsigned)(threadIdx.x & 31 & ~(LPP - 1));
    constexpr unsigned GM = LPP == 32 ? 0xffffffffu : ((1u << (LPP & 31)) - 1u);
    const unsigned gmask = GM << gsh;
    auto gballot = [&](bool p) -> unsigned { return (__ballot_sync(gmask, p) >> gsh) & GM; };
    svp_result res;
    res.score = 0; res.status = 0; res.q_beg = res.q_end = res.t_beg = res.t_end = 0;
    res.n_match = res.n_mismatch = res.n_ins_ops = res.n_ins_bp = res.n_del_ops = res.n_del_bp = 0;
    res.cigar_off = tk.cigar_off; res.cigar_len = 0; res.sv_dist = 0.0; res.cells = tk.cells;
    const int KP = (int)((tk.flags & TASK_KP_MASK) >> TASK_KP_SHIFT);
    TbView<CG> v;
    v.base = rows; v.seqs32 = reinterpret_cast<const uint32_t*>(seqs);
    v.band_lo = tk.band_lo; v.r_begin = tk.r_begin; v.m = tk.m; v.n = tk.n;
    v.set_geometry(tk.band_w, KP);
    v.q_word = tk.q_off >> 2; v.t_word = tk.t_off >> 2; v.rc = (tk.flags & SVP_PAIR_REVCOMP) != 0;
    v.q_mask = seq_mask(seqs, tk.q_off, tk.m, (tk.flags & TASK_QN) != 0);
    v.t_mask = seq_mask(seqs, tk.t_off, tk.n, (tk.flags & TASK_TN) != 0);
    v.s_cor = s_cor; v.cor_wc = 0; v.cor_w = s_cor ? min(COR_W, v.pitch) : 0; v.cor_lo = 0; v.cor_ready = 0x7FFFFFF0; v.cor_top = -1;
    v.cor_row_lo = 0; v.cor_row_span = 0;
    v.n_rows = (int)(tk.tb_bytes / (uint64_t)v.pitch);
    const int d_hi = tk.band_lo + tk.band_w - 1;

    // ---- 1. end cell ---------------------------------------------------------------------------
    const int n_ent = (tk.band_w / (4 * KP)) * 2;
    int M = 0; uint32_t blk = 0xFFFFFFFFu;
    for (int e = lane; e < n_ent; e += LPP) {
        const uint2 b = ldbest<CG>(bests + e);
        const int val = (int)b.x;
        if (val > M || (val == M && b.y < blk)) { M = val; blk = b.y; }
    }
    #pragma unroll
    for (int o = LPP / 2; o; o >>= 1) {
        const int M2 = __shfl_xor_sync(gmask, M, o, LPP);
        const uint32_t b2 = __shfl_xor_sync(gmask, blk, o, LPP);
        if (M2 > M || (M2 == M && b2 < blk)) { M = M2; blk = b2; }
    }
    if (M <= 0) {
        res.status = SVP_ST_NOHIT;
        if (lane == 0) results[tk.pair_index] = res;
        return;
    }
    if (!(tk.flags & (TASK_S32 | TASK_REBASE)) && M >= 32767 - cs.a_max) res.status |= SVP_ST_SCORE_OVF;
    int end_s = 0x7FFFFFFF, end_x = -1;
    const int NC = 2 * KP * KP;
    for (int e0 = 0; e0 < n_ent; e0 += LPP) {
        const int e = e0 + lane;
        bool hit = false;
        if (e < n_ent) { const uint2 b = ldbest<CG>(bests + e); hit = ((int)b.x == M) && (b.y == blk); }
        unsigned mask = gballot(hit);
        while (mask) {
            const int ee = e0 + __ffs(mask) - 1; mask &= mask - 1;
            const int xbase = (ee >> 1) * 4 * KP + (ee & 1) * 2 * KP;     // first diagonal (band-relative) of the half
            const int s0 = (int)blk * 2 * KP;
            int rel_row0 = 0, rel = 0, relmax = -(1 << 30), cand_s = 0x7FFFFFFF, cand_x = -1;
            uint32_t b_row0 = 0, bprev = 0;
            for (int c0 = 0; c0 < NC; c0 += LPP) {            // bytes fetched LPP at a time, chained by all lanes in lockstep
                const int idx = c0 + lane;
                uint32_t mine = 0;
                if (idx < NC) {
                    const int s = s0 + idx / KP, x = xbase + 2 * (idx % KP) + (s & 1);
                    mine = ldrow<CG>(v.addr_sx(s, x));
                }
                const int lim = min(LPP, NC - c0);
                for (int l = 0; l < lim; ++l) {
                    const uint32_t b = __shfl_sync(gmask, mine, l, LPP);
                    const int id = c0 + l;
                    if (id == 0) rel = 0;
                    else if (id % KP == 0) rel = rel_row0 + sext8(b - b_row0);
                    else rel = rel + sext8(b - bprev);
                    if (id % KP == 0) { rel_row0 = rel; b_row0 = b; }
                    bprev = b;
                    const int s = s0 + id / KP, x = xbase + 2 * (id % KP) + (s & 1);
                    if (rel > relmax) { relmax = rel; cand_s = s; cand_x = x; }
                    else if (rel == relmax && (s < cand_s || (s == cand_s && x > cand_x))) { cand_s = s; cand_x = x; }
                }
            }
            if (cand_s < end_s || (cand_s == end_s && cand_x > end_x)) { end_s = cand_s; end_x = cand_x; }
        }
    }
    int i = ((tk.r_begin + end_s) - (tk.band_lo + end_x)) / 2, j = ((tk.r_begin + end_s) + (tk.band_lo + end_x)) / 2;
    const int best_i = i, best_j = j;

    // ---- 2. walk back --------------------------------------------------------------------------
    int H = M;
    uint32_t* slot = cigar + tk.cigar_off;
    int wpos = (int)tk.cigar_cap;                    // runs are written right-aligned, backwards
    uint32_t cur_op = 0, cur_len = 0; bool ovf = false, bad = false;
    uint32_t n_match = 0, n_mis = 0, n_io = 0, n_ib = 0, n_do = 0, n_db = 0;
    // snapshot at the lowest H of the walk: where the alignment starts if the walk is cut (Z-trim, §3.5) or ends at H = 0
    int low = M + 1, low_i = i, low_j = j, low_wpos = wpos; uint32_t low_op = 0, low_len = 0; bool low_ovf = false;
    uint32_t l_match = 0, l_mis = 0, l_io = 0, l_ib = 0, l_do = 0, l_db = 0;
    const int zdrop = cs.zdrop > 0 ? cs.zdrop : 0x3FFFFFFF;

#define SVP_PUSHW() do { if (cur_len) { if (wpos > 0) { --wpos; if (lane == 0) slot[wpos] = (cur_len << 4) | cur_op; } else ovf = true; } } while (0)
#define SVP_EMITW(op, len) do { if (cur_len && cur_op == (op)) cur_len += (len); else { SVP_PUSHW(); cur_op = (op); cur_len = (len); } } while (0)


    // loop invariants in registers (tk and cs sit behind references: local / constant memory)
    const int band_lo = tk.band_lo, r_begin = tk.r_begin;
    const bool simple = cs.general == 0;                     // one match and one mismatch score (a letter outside ACGT scores the mismatch)
    const int s_match = cs.s_match, s_mismatch = cs.s_mismatch, n_score = cs.n_score, zd_ext = cs.zdrop_ext;
    const uint32_t tab0 = cs.tab[0], tab1 = cs.tab[1], tab2 = cs.tab[2], tab3 = cs.tab[3];
    const int cost1d = gap_cost2(cs.d_open1, cs.d_ext1, cs.d_open2, cs.d_ext2, cs.two_piece, 1);
    const int cost1i = gap_cost2(cs.i_open1, cs.i_ext1, cs.i_open2, cs.i_ext2, cs.two_piece, 1);
    const unsigned lt_mask = (1u << lane) - 1u;
    bool done = false;
    while (!done) {
        // ---- diagonal run: lane l holds the diagonal predecessor byte and the bases of cell (i-l, j-l) ----
        if (LPP == 32) v.cor_advance(i, j, lane);
        const int ci = i - lane, cj = j - lane;
        const int d0 = j - i;
        const bool del_band = d0 - 1 >= band_lo, ins_band = d0 + 1 <= d_hi;
        const bool inb = ci >= 1 && cj >= 1;
        // The cells of a diagonal keep their byte column; their rows go down by one per cell. Columns once per round:
        const int x0 = d0 - band_lo;
        const int c0 = v.col(x0), cL = del_band ? v.col(x0 - 1) : 0, cU = ins_band ? v.col(x0 + 1) : 0;
        const int sd = (i + j) - r_begin - 2 * lane - 2;     // step of the lane's diagonal predecessor; its left / up neighbours: sd + 1
        uint32_t bp = 0, bl = 0, bu = 0; int sub_l = 0; bool same_l = false;
        if (inb) {
            if (ci >= 2 && cj >= 2 && sd >= 0) bp = v.ld(sd >> 1, c0);
            if (del_band && cj >= 2 && sd + 1 >= 0) bl = v.ld((sd + 1) >> 1, cL);          // left / up neighbours: a 1-base gap is resolved from this round's loads
            if (ins_band && ci >= 2 && sd + 1 >= 0) bu = v.ld((sd + 1) >> 1, cU);
            const int qc = v.qbase(ci), tc = v.tbase(cj);
            const bool other = v.qn(ci) || v.tn(cj);              // a letter outside ACGT: lowest score, never a match
            same_l = !other && qc == tc;
            if (simple) sub_l = same_l ? s_match : s_mismatch;
            else {
                const uint32_t trow = qc < 2 ? (qc == 0 ? tab0 : tab1) : (qc == 2 ? tab2 : tab3);
                sub_l = other ? n_score : (int)(int8_t)((trow >> (8 * tc)) & 0xFFu);
            }
        }
        if (!s_cor && ci - LPP - 1 >= 1 && cj - LPP - 1 >= 1) asm volatile("prefetch.global.L2 [%0];" :: "l"(v.addr(ci - LPP - 1, cj - LPP - 1)));   // next round's byte
        // ---- the run is resolved by all lanes at once: lane l assumes cells 0..l-1 were diagonal moves, so its H is the
        // incoming H minus the substitution scores before it (prefix sum); the first lane that stops (H = 0, matrix edge,
        // Z-drop) or does not continue diagonally ends the run. The lowest-H snapshot (strict minimum, first occurrence)
        // is a prefix minimum. Same walk as the cell-by-cell loop of svp_tb_kernel.
        bool gap = false;
        int l_gap = 0;
        {
            const unsigned sameb = gballot(same_l);
            int Hl, cnt_same = 0;
            if (simple) {
                // two score values: the prefix sum is a population count (lanes behind the first cell outside the matrix get a
                // wrong value; the run ends at that cell or before it, and nothing behind the end of the run is used)
                const int ns = __popc(sameb & lt_mask);
                Hl = H - (ns * s_match + (lane - ns) * s_mismatch);
            } else {
                int incl = sub_l;
                #pragma unroll
                for (int o = 1; o < LPP; o <<= 1) { const int t = __shfl_up_sync(gmask, incl, o, LPP); if (lane >= o) incl += t; }
                Hl = H - (incl - sub_l);
            }
            const bool stop2 = (Hl == 0) || !inb;
            // Z-drop: Hl - min(low, prefix minimum) > zdrop + ... needs the prefix minimum only if some lane can exceed zdrop at all
            bool stop3 = false, have_pmin = false;
            int pmin = Hl;
            if (__reduce_max_sync(gmask, Hl) - min(low, __reduce_min_sync(gmask, Hl)) > zdrop) {
                #pragma unroll
                for (int o = 1; o < LPP; o <<= 1) { const int t = __shfl_up_sync(gmask, pmin, o, LPP); if (lane >= o) pmin = min(pmin, t); }
                int dd = d0 - (low_j - low_i); dd = dd < 0 ? -dd : dd;
                if (pmin < low) dd = 0;                                   // the snapshot moved onto this diagonal at or before this cell
                stop3 = Hl - min(low, pmin) > zdrop + zd_ext * dd;
                have_pmin = true;
            }
            const int Hd = (ci >= 2 && cj >= 2) ? Hl + sext8(bp - (uint32_t)Hl) : 0;
            const bool isdiag = inb && (Hl == Hd + sub_l);
            const unsigned stopb = gballot(stop2 || stop3);
            const unsigned term = stopb | gballot(!isdiag);
            const int first = term ? __ffs(term) - 1 : LPP;
            const int K = first < LPP - 1 ? first : LPP - 1;
            const int runmin = have_pmin ? __shfl_sync(gmask, pmin, K, LPP) : __reduce_min_sync(gmask, lane <= K ? Hl : 0x7FFFFFFF);
            if (runmin < low) {                                       // snapshot at the first cell of the run that reaches the new minimum
                const unsigned upto = K == 31 ? 0xffffffffu : ((2u << K) - 1u);
                const int ks = __ffs(gballot(Hl == runmin) & upto) - 1;
                low = runmin; low_i = i - ks; low_j = j - ks;
                low_wpos = wpos; low_op = cur_op; low_len = cur_len; low_ovf = ovf;
                if (ks > 0) {                                         // CIGAR state after ks more M's
                    if (low_len && low_op == SVP_CIGAR_M) low_len += (uint32_t)ks;
                    else { if (low_len) { if (low_wpos > 0) --low_wpos; else low_ovf = true; } low_op = SVP_CIGAR_M; low_len = (uint32_t)ks; }
                }
                cnt_same = __popc(sameb & ((1u << ks) - 1u));
                l_match = n_match + (uint32_t)cnt_same; l_mis = n_mis + (uint32_t)(ks - cnt_same);
                l_io = n_io; l_ib = n_ib; l_do = n_do; l_db = n_db;
            }
            const int Hnext = __shfl_sync(gmask, first < LPP ? Hl : Hl - sub_l, K, LPP);
            const bool fin = first < LPP && ((stopb >> K) & 1u);
            if (first > 0) {
                SVP_EMITW(SVP_CIGAR_M, (uint32_t)first);
                cnt_same = __popc(sameb & (first >= 32 ? 0xffffffffu : ((1u << first) - 1u)));
                n_match += (uint32_t)cnt_same; n_mis += (uint32_t)(first - cnt_same);
                i -= first; j -= first; H = Hnext;
            }
            if (fin) done = true;
            else if (first < LPP) { gap = true; l_gap = first; }
        }
        if (done || !gap) continue;
        {   // 1-base gap at the cell the run stopped at (deletion first, as in the search below)
            const uint32_t b1 = __shfl_sync(gmask, bl, l_gap, LPP), b2 = __shfl_sync(gmask, bu, l_gap, LPP);
            const int Hl = H + sext8(b1 - (uint32_t)H), Hu = H + sext8(b2 - (uint32_t)H);
            if (del_band && j - 1 >= 1 && H == Hl - cost1d) {
                SVP_EMITW(SVP_CIGAR_D, 1u); ++n_do; ++n_db; --j; H = Hl; continue;
            }
            if (ins_band && i - 1 >= 1 && H == Hu - cost1i) {
                SVP_EMITW(SVP_CIGAR_I, 1u); ++n_io; ++n_ib; --i; H = Hu; continue;
            }
        }
        // ---- gap: the smallest k with H == H(i, j-k) - del(k), else H == H(i-k, j) - ins(k) ----
        // 32 candidates of the row and of the column per round, resolved by all lanes at once: lane l holds candidate k0 + l + 1,
        // its H is the carried H plus the prefix sum of the byte differences along the row / column (5 SHFL.UP each), one ballot
        // per gap kind finds the first lane whose H explains the gap (the sequential search cost ~15 instructions per candidate).
        if (LPP == 32) v.cor_advance(i, j, lane);
        const int d = j - i;
        int Hrow = H, Hcol = H; uint32_t brow = (uint32_t)H & 0xFFu, bcol = (uint32_t)H & 0xFFu;
        bool found = false, dead = false;
        for (int k0 = 0; !found && !dead; k0 += LPP) {
            const int kk = k0 + lane + 1;
            const bool del_ok = (j - kk >= 1) && (d - kk >= tk.band_lo);      // both are true for a prefix of the candidates only
            const bool ins_ok = (i - kk >= 1) && (d + kk <= d_hi);
            const uint32_t bd = del_ok ? v.lb(i, j - kk) : 0;
            const uint32_t bi = ins_ok ? v.lb(i - kk, j) : 0;
            uint32_t pd = __shfl_up_sync(gmask, bd, 1, LPP), pi = __shfl_up_sync(gmask, bi, 1, LPP);
            if (lane == 0) { pd = brow; pi = bcol; }
            int dr = del_ok ? sext8(bd - pd) : 0, dc = ins_ok ? sext8(bi - pi) : 0;
            #pragma unroll
            for (int o = 1; o < LPP; o <<= 1) {
                const int tr = __shfl_up_sync(gmask, dr, o, LPP), tc = __shfl_up_sync(gmask, dc, o, LPP);
                if (lane >= o) { dr += tr; dc += tc; }
            }
            const int hr = Hrow + dr, hc = Hcol + dc;
            const unsigned md = gballot(del_ok && H == hr - gap_cost2(cs.d_open1, cs.d_ext1, cs.d_open2, cs.d_ext2, cs.two_piece, kk));
            const unsigned mi = gballot(ins_ok && H == hc - gap_cost2(cs.i_open1, cs.i_ext1, cs.i_open2, cs.i_ext2, cs.two_piece, kk));
            const unsigned mx = gballot(!del_ok && !ins_ok);
            const int fd = md ? __ffs(md) - 1 : LPP, fi = mi ? __ffs(mi) - 1 : LPP, fx = mx ? __ffs(mx) - 1 : LPP;
            if (min(fd, fi) < fx) {                       // at equal length the deletion comes first
                found = true;
                if (fd <= fi) {
                    const int len = k0 + fd + 1;
                    SVP_EMITW(SVP_CIGAR_D, (uint32_t)len); ++n_do; n_db += len; j -= len; H = __shfl_sync(gmask, hr, fd, LPP);
                } else {
                    const int len = k0 + fi + 1;
                    SVP_EMITW(SVP_CIGAR_I, (uint32_t)len); ++n_io; n_ib += len; i -= len; H = __shfl_sync(gmask, hc, fi, LPP);
                }
            } else if (fx < LPP) {
                dead = true;
            } else {
                Hrow = __shfl_sync(gmask, hr, LPP - 1, LPP); Hcol = __shfl_sync(gmask, hc, LPP - 1, LPP);
                brow = __shfl_sync(gmask, bd, LPP - 1, LPP); bcol = __shfl_sync(gmask, bi, LPP - 1, LPP);
            }
        }
        if (!found) { bad = true; done = true; }
    }
    if (s_cor) { asm volatile("cp.async.wait_group 0;" ::: "memory"); __syncwarp(gmask); }
    // the alignment starts at the lowest point of the walk
    const int cut_low = low;
    i = low_i; j = low_j; wpos = low_wpos; cur_op = low_op; cur_len = low_len; ovf = low_ovf;
    SVP_PUSHW();
#undef SVP_PUSHW
#undef SVP_EMITW
    if (bad) res.status |= SVP_ST_SCORE_OVF;
    res.score = M - cut_low;
    res.q_beg = i; res.q_end = best_i; res.t_beg = j; res.t_end = best_j;
    res.n_match = l_match; res.n_mismatch = l_mis; res.n_ins_ops = l_io; res.n_ins_bp = l_ib; res.n_del_ops = l_do; res.n_del_bp = l_db;
    __syncwarp(gmask);
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
                    const double sz = (double)len;
                    dist += (1.0 - exp(-pow(sz / 30.0, 1.5))) * sz;
                }
            }
            res.sv_dist = dist;
        }
        results[tk.pair_index] = res;
    }
}

// ---------------------------------------------------------------------------------------------
// The group walk with WARP-UNIFORM control flow: 32 / LPP pairs per warp, one per group of LPP lanes, all of them on one
// instruction stream. Same walk, same results as svp_traceback_warp; what differs is how the groups share the warp:
//   * every cross-lane operation is issued by all 32 lanes with the full mask at a point all groups reach together (ballots are
//     shifted and masked to the group, shuffles use width LPP, minima are butterflies inside the group) — a ballot or reduction on
//     a per-group mask makes the compiler loop over the distinct masks (MATCH.ANY), one pass per group;
//   * nothing a group decides changes the program counter of another: the diagonal-run round, the 1-base gap and the first LPP
//     candidates of a longer gap are straight-line code in every iteration, applied under per-group predicates; loops run while
//     ANY group needs them (further gap candidates, the Z-drop prefix minimum, ties of the end-cell search are the rare cases);
//   * a finished group idles until the last group of its warp is done (the pairs of a warp are neighbours in the wave's order).
// profiles/ncu_tbg_r02_v1.json (group masks, divergent groups): 524 k warp instructions per warp of four pairs, 771 per round,
// every VOTE / REDUX executed four times.
// ---------------------------------------------------------------------------------------------
// rows: the group's slice (rows, then best records); CG: loads of the slice through L2 (written by a concurrently running grid)
template <int LPP, bool CG>
__device__ __forceinline__ void svp_traceback_groups(const Task* __restrict__ tkp, const uint8_t* __restrict__ rows,
                                                  const uint8_t* __restrict__ seqs, svp_result* __restrict__ results,
                                                  uint32_t* __restrict__ cigar, const Consts& cs) {
    static_assert(LPP == 4 || LPP == 8 || LPP == 16, "lanes per pair");
    constexpr unsigned FULL = 0xffffffffu, GM = (1u << LPP) - 1u;
    const int lane = threadIdx.x & (LPP - 1);
    const unsigned gsh = (unsigned)(threadIdx.x & 31 & ~(LPP - 1));
    auto gballot = [&](bool p) -> unsigned { return (__ballot_sync(FULL, p) >> gsh) & GM; };
    auto gmin = [&](int x) -> int {
        #pragma unroll
        for (int o = LPP / 2; o; o >>= 1) x = min(x, __shfl_xor_sync(FULL, x, o, LPP));
        return x;
    };
    const bool have = tkp != nullptr;
    // the group's pair, field by field into registers (an idle group gets a harmless geometry; everything it does is predicated off)
    struct { uint64_t q_off, t_off, tb_bytes, cells; int32_t m, n, band_lo, band_w, r_begin; uint32_t flags, pair_index, cigar_off, cigar_cap; } tk;
    tk.q_off = have ? tkp->q_off : 0; tk.t_off = have ? tkp->t_off : 0; tk.tb_bytes = have ? tkp->tb_bytes : 16;
    tk.cells = have ? tkp->cells : 0; tk.m = have ? tkp->m : 1; tk.n = have ? tkp->n : 1; tk.band_lo = have ? tkp->band_lo : 0;
    tk.band_w = have ? tkp->band_w : 16; tk.r_begin = have ? tkp->r_begin : 2; tk.flags = have ? tkp->flags : (4u << TASK_KP_SHIFT);
    tk.pair_index = have ? tkp->pair_index : 0; tk.cigar_off = have ? tkp->cigar_off : 0; tk.cigar_cap = have ? tkp->cigar_cap : 0;
    const uint2* bests = reinterpret_cast<const uint2*>(rows + tk.tb_bytes);
    auto write_result = [&](uint32_t status, int score, int q_beg, int q_end, int t_beg, int t_end, uint32_t nm, uint32_t nx, uint32_t io, uint32_t ib,
                            uint32_t dop, uint32_t db, uint32_t c_off, uint32_t c_len, double dist) {
        svp_result r;
        r.score = score; r.status = status; r.q_beg = q_beg; r.q_end = q_end; r.t_beg = t_beg; r.t_end = t_end;
        r.n_match = nm; r.n_mismatch = nx; r.n_ins_ops = io; r.n_ins_bp = ib; r.n_del_ops = dop; r.n_del_bp = db;
        r.cigar_off = c_off; r.cigar_len = c_len; r.sv_dist = dist; r.cells = tk.cells;
        results[tk.pair_index] = r;
    };
    uint32_t status = 0;
    const int KP = (int)((tk.flags & TASK_KP_MASK) >> TASK_KP_SHIFT);
    TbView<CG> v;
    v.base = rows; v.seqs32 = reinterpret_cast<const uint32_t*>(seqs);
    v.band_lo = tk.band_lo; v.r_begin = tk.r_begin; v.m = tk.m; v.n = tk.n;
    v.set_geometry(tk.band_w, KP);
    v.q_word = tk.q_off >> 2; v.t_word = tk.t_off >> 2; v.rc = (tk.flags & SVP_PAIR_REVCOMP) != 0;
    v.q_mask = seq_mask(seqs, tk.q_off, tk.m, (tk.flags & TASK_QN) != 0);
    v.t_mask = seq_mask(seqs, tk.t_off, tk.n, (tk.flags & TASK_TN) != 0);
    v.s_cor = 0; v.cor_wc = 0; v.cor_w = 0; v.cor_lo = 0; v.cor_ready = 0x7FFFFFF0; v.cor_top = -1; v.cor_row_lo = 0; v.cor_row_span = 0;
    v.n_rows = (int)(tk.tb_bytes / (uint64_t)v.pitch);
    const int band_lo = tk.band_lo, r_begin = tk.r_begin;
    const int d_hi = band_lo + tk.band_w - 1;

    // ---- 1. end cell: maximum over the per-(thread, half) records, earliest block; then the exact cell inside the winning blocks ----
    const int n_ent = have ? (tk.band_w / (4 * KP)) * 2 : 0;
    const int n_ent_max = __reduce_max_sync(FULL, n_ent);
    int M = 0; uint32_t blk = 0xFFFFFFFFu;
    for (int e = lane; e < n_ent_max; e += LPP) {
        if (e < n_ent) {
            const uint2 b = ldbest<CG>(bests + e);
            const int val = (int)b.x;
            if (val > M || (val == M && b.y < blk)) { M = val; blk = b.y; }
        }
    }
    #pragma unroll
    for (int o = LPP / 2; o; o >>= 1) {
        const int M2 = __shfl_xor_sync(FULL, M, o, LPP);
        const uint32_t b2 = __shfl_xor_sync(FULL, blk, o, LPP);
        if (M2 > M || (M2 == M && b2 < blk)) { M = M2; blk = b2; }
    }
    const bool act = have && M > 0;                              // the group has a walk to do
    if (have && !act && lane == 0) write_result(SVP_ST_NOHIT, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, tk.cigar_off, 0, 0.0);
    if (act && !(tk.flags & (TASK_S32 | TASK_REBASE)) && M >= 32767 - cs.a_max) status |= SVP_ST_SCORE_OVF;
    int end_s = 0x7FFFFFFF, end_x = -1;
    const int NC = 2 * KP * KP;
    const int NC_max = __reduce_max_sync(FULL, NC);
    for (int e0 = 0; e0 < n_ent_max; e0 += LPP) {
        const int e = e0 + lane;
        bool hit = false;
        if (act && e < n_ent) { const uint2 b = ldbest<CG>(bests + e); hit = ((int)b.x == M) && (b.y == blk); }
        unsigned mask = gballot(hit);
        while (__any_sync(FULL, mask != 0u)) {
            const bool on = mask != 0u;
            const int ee = e0 + (on ? __ffs(mask) - 1 : 0);
            mask &= mask - 1u;                                   // (0 stays 0)
            const int xbase = (ee >> 1) * 4 * KP + (ee & 1) * 2 * KP;     // first diagonal (band-relative) of the half
            const int s0 = (int)blk * 2 * KP;
            int rel_row0 = 0, rel = 0, relmax = -(1 << 30), cand_s = 0x7FFFFFFF, cand_x = -1;
            uint32_t b_row0 = 0, bprev = 0;
            for (int c0 = 0; c0 < NC_max; c0 += LPP) {           // bytes fetched LPP at a time, chained by all lanes in lockstep
                const int idx = c0 + lane;
                uint32_t mine = 0;
                if (on && idx < NC) {
                    const int s = s0 + idx / KP, x = xbase + 2 * (idx % KP) + (s & 1);
                    mine = ldrow<CG>(v.addr_sx(s, x));
                }
                #pragma unroll
                for (int l = 0; l < LPP; ++l) {
                    const uint32_t b = __shfl_sync(FULL, mine, l, LPP);
                    const int id = c0 + l;
                    if (on && id < NC) {
                        if (id == 0) rel = 0;
                        else if (id % KP == 0) rel = rel_row0 + sext8(b - b_row0);
                        else rel = rel + sext8(b - bprev);
                        if (id % KP == 0) { rel_row0 = rel; b_row0 = b; }
                        bprev = b;
                        const int s = s0 + id / KP, x = xbase + 2 * (id % KP) + (s & 1);
                        if (rel > relmax) { relmax = rel; cand_s = s; cand_x = x; }
                        else if (rel == relmax && (s < cand_s || (s == cand_s && x > cand_x))) { cand_s = s; cand_x = x; }
                    }
                }
            }
            if (on && (cand_s < end_s || (cand_s == end_s && cand_x > end_x))) { end_s = cand_s; end_x = cand_x; }
        }
    }
    int i = 1, j = 1;
    if (act) { i = ((r_begin + end_s) - (band_lo + end_x)) / 2; j = ((r_begin + end_s) + (band_lo + end_x)) / 2; }
    const int best_i = i, best_j = j;

    // ---- 2. walk back --------------------------------------------------------------------------
    int H = M;
    uint32_t* slot = cigar + tk.cigar_off;
    int wpos = (int)tk.cigar_cap;                    // runs are written right-aligned, backwards
    uint32_t cur_op = 0, cur_len = 0; bool ovf = false, bad = false;
    uint32_t n_match = 0, n_mis = 0, n_io = 0, n_ib = 0, n_do = 0, n_db = 0;
    // snapshot at the lowest H of the walk: where the alignment starts if the walk is cut (Z-trim, §3.5) or ends at H = 0
    int low = M + 1, low_i = i, low_j = j, low_wpos = wpos; uint32_t low_op = 0, low_len = 0; bool low_ovf = false;
    uint32_t l_match = 0, l_mis = 0, l_io = 0, l_ib = 0, l_do = 0, l_db = 0;
    const int zdrop = cs.zdrop > 0 ? cs.zdrop : 0x3FFFFFFF;
    const bool simple = cs.general == 0;
    const int s_match = cs.s_match, s_mismatch = cs.s_mismatch, n_score = cs.n_score, zd_ext = cs.zdrop_ext, a_max = cs.a_max;
    const uint32_t tab0 = cs.tab[0], tab1 = cs.tab[1], tab2 = cs.tab[2], tab3 = cs.tab[3];
    const int d_o1 = cs.d_open1, d_e1 = cs.d_ext1, d_o2 = cs.d_open2, d_e2 = cs.d_ext2, i_o1 = cs.i_open1, i_e1 = cs.i_ext1, i_o2 = cs.i_open2, i_e2 = cs.i_ext2;
    const int two = cs.two_piece;
    const int cost1d = gap_cost2(d_o1, d_e1, d_o2, d_e2, two, 1), cost1i = gap_cost2(i_o1, i_e1, i_o2, i_e2, two, 1);
    const unsigned lt_mask = (1u << lane) - 1u;

#define SVP_PUSHG() do { if (cur_len) { if (wpos > 0) { --wpos; if (lane == 0) slot[wpos] = (cur_len << 4) | cur_op; } else ovf = true; } } while (0)
#define SVP_EMITG(op, len) do { if (cur_len && cur_op == (op)) cur_len += (len); else { SVP_PUSHG(); cur_op = (op); cur_len = (len); } } while (0)

    // byte columns of the current diagonal and its two neighbours, kept across rounds (a 1-base gap shifts them by one)
    int xc = (j - i) - band_lo;
    int c0 = v.col(xc), cL = v.col(xc - 1), cU = v.col(xc + 1);
    bool done = !act;
    while (__any_sync(FULL, !done)) {
        const bool run = !done;
        // ---- diagonal run: lane l holds the diagonal predecessor byte and the bases of cell (i-l, j-l) ----
        const int ci = i - lane, cj = j - lane;
        const int d0 = j - i;
        const bool del_band = d0 - 1 >= band_lo, ins_band = d0 + 1 <= d_hi;
        const bool inb = ci >= 1 && cj >= 1;
        const int x0 = d0 - band_lo;
        {   // columns: one new one per round in the common case (the diagonal moved by one), all three after a longer gap
            const int shift = x0 - xc;
            if (__any_sync(FULL, run && (shift > 1 || shift < -1))) {
                if (shift > 1 || shift < -1) { c0 = v.col(x0); cL = v.col(x0 - 1); cU = v.col(x0 + 1); }
                else if (shift) { const int cn = v.col(shift < 0 ? x0 - 1 : x0 + 1); if (shift < 0) { cU = c0; c0 = cL; cL = cn; } else { cL = c0; c0 = cU; cU = cn; } }
            } else if (__any_sync(FULL, shift != 0)) {
                const int cn = v.col(shift < 0 ? x0 - 1 : x0 + 1);
                if (shift < 0) { cU = c0; c0 = cL; cL = cn; } else if (shift > 0) { cL = c0; c0 = cU; cU = cn; }
            }
            xc = x0;
        }
        const int sd = (i + j) - r_begin - 2 * lane - 2;     // step of the lane's diagonal predecessor; its left / up neighbours: sd + 1
        uint32_t bp = 0, bl = 0, bu = 0; int sub_l = 0; bool same_l = false;
        if (run && inb) {
            if (ci >= 2 && cj >= 2 && sd >= 0) bp = v.ld(sd >> 1, c0);
            if (del_band && cj >= 2 && sd + 1 >= 0) bl = v.ld((sd + 1) >> 1, cL);
            if (ins_band && ci >= 2 && sd + 1 >= 0) bu = v.ld((sd + 1) >> 1, cU);
            const int qc = v.qbase(ci), tc = v.tbase(cj);
            const bool other = v.qn(ci) || v.tn(cj);              // a letter outside ACGT: lowest score, never a match
            same_l = !other && qc == tc;
            if (simple) sub_l = same_l ? s_match : s_mismatch;
            else {
                const uint32_t trow = qc < 2 ? (qc == 0 ? tab0 : tab1) : (qc == 2 ? tab2 : tab3);
                sub_l = other ? n_score : (int)(int8_t)((trow >> (8 * tc)) & 0xFFu);
            }
        }
        if (run && ci - LPP - 1 >= 1 && cj - LPP - 1 >= 1 && sd - 2 * LPP >= 0)
            asm volatile("prefetch.global.L2 [%0];" :: "l"(v.base + (ptrdiff_t)((sd - 2 * LPP) >> 1) * v.pitch + c0));     // next round's byte
        const unsigned sameb = gballot(same_l);
        int Hl;
        if (simple) {
            const int ns = __popc(sameb & lt_mask);
            Hl = H - (ns * s_match + (lane - ns) * s_mismatch);
        } else {
            int incl = sub_l;
            #pragma unroll
            for (int o = 1; o < LPP; o <<= 1) { const int t = __shfl_up_sync(FULL, incl, o, LPP); if (lane >= o) incl += t; }
            Hl = H - (incl - sub_l);
        }
        const bool stop2 = (Hl == 0) || !inb;
        // Z-drop: Hl - min(low, prefix minimum) > zdrop + ... ; the prefix minimum is at least H - lane * a_max, so the scan is only
        // needed when some lane of some group passes the test with that bound
        bool stop3 = false;
        int pmin = Hl;
        const bool scan = __any_sync(FULL, run && (Hl - min(low, H - lane * a_max) > zdrop));
        if (scan) {
            #pragma unroll
            for (int o = 1; o < LPP; o <<= 1) { const int t = __shfl_up_sync(FULL, pmin, o, LPP); if (lane >= o) pmin = min(pmin, t); }
            int dd = d0 - (low_j - low_i); dd = dd < 0 ? -dd : dd;
            if (pmin < low) dd = 0;                                   // the snapshot moved onto this diagonal at or before this cell
            stop3 = Hl - min(low, pmin) > zdrop + zd_ext * dd;
        }
        const int Hd = (ci >= 2 && cj >= 2) ? Hl + sext8(bp - (uint32_t)Hl) : 0;
        const bool isdiag = inb && (Hl == Hd + sub_l);
        const unsigned stopb = gballot(stop2 || stop3);
        const unsigned term = stopb | gballot(!isdiag);
        const int first = term ? __ffs(term) - 1 : LPP;
        const int K = first < LPP - 1 ? first : LPP - 1;
        const int runmin = scan ? __shfl_sync(FULL, pmin, K, LPP) : gmin(lane <= K ? Hl : 0x7FFFFFFF);
        const unsigned eqb = gballot(Hl == runmin);
        const int Hnext = __shfl_sync(FULL, first < LPP ? Hl : Hl - sub_l, K, LPP);
        const uint32_t b1 = __shfl_sync(FULL, bl, K, LPP), b2 = __shfl_sync(FULL, bu, K, LPP);      // neighbours of the cell the run stops at
        bool search = false;
        if (run) {
            if (runmin < low) {                                       // snapshot at the first cell of the run that reaches the new minimum
                const unsigned upto = (2u << K) - 1u;
                const int ks = __ffs(eqb & upto) - 1;
                low = runmin; low_i = i - ks; low_j = j - ks;
                low_wpos = wpos; low_op = cur_op; low_len = cur_len; low_ovf = ovf;
                if (ks > 0) {                                         // CIGAR state after ks more M's
                    if (low_len && low_op == SVP_CIGAR_M) low_len += (uint32_t)ks;
                    else { if (low_len) { if (low_wpos > 0) --low_wpos; else low_ovf = true; } low_op = SVP_CIGAR_M; low_len = (uint32_t)ks; }
                }
                const int cs_ = __popc(sameb & ((1u << ks) - 1u));
                l_match = n_match + (uint32_t)cs_; l_mis = n_mis + (uint32_t)(ks - cs_);
                l_io = n_io; l_ib = n_ib; l_do = n_do; l_db = n_db;
            }
            const bool fin = first < LPP && ((stopb >> K) & 1u);
            if (first > 0) {
                SVP_EMITG(SVP_CIGAR_M, (uint32_t)first);
                const int cs_ = __popc(sameb & ((1u << first) - 1u));
                n_match += (uint32_t)cs_; n_mis += (uint32_t)(first - cs_);
                i -= first; j -= first; H = Hnext;
            }
            if (fin) done = true;
            else if (first < LPP) {
                // 1-base gap at the cell the run stopped at (deletion first, as in the search below)
                const int Hl1 = H + sext8(b1 - (uint32_t)H), Hu1 = H + sext8(b2 - (uint32_t)H);
                if (del_band && j - 1 >= 1 && H == Hl1 - cost1d) { SVP_EMITG(SVP_CIGAR_D, 1u); ++n_do; ++n_db; --j; H = Hl1; }
                else if (ins_band && i - 1 >= 1 && H == Hu1 - cost1i) { SVP_EMITG(SVP_CIGAR_I, 1u); ++n_io; ++n_ib; --i; H = Hu1; }
                else search = true;
            }
        }
        // ---- longer gap: the smallest k with H == H(i, j-k) - del(k), else H == H(i-k, j) - ins(k); LPP candidates of the row and of the
        // column per pass, for every group that needs them (all groups go through a pass together: no group waits for another's branch)
        {
            const int d = j - i;
            int Hrow = H, Hcol = H; uint32_t brow = (uint32_t)H & 0xFFu, bcol = (uint32_t)H & 0xFFu;
            int k0 = 0;
            while (__any_sync(FULL, search)) {
                const int kk = k0 + lane + 1;
                const bool del_ok = search && (j - kk >= 1) && (d - kk >= band_lo);      // both are true for a prefix of the candidates only
                const bool ins_ok = search && (i - kk >= 1) && (d + kk <= d_hi);
                const uint32_t bd = del_ok ? v.lb(i, j - kk) : 0;
                const uint32_t bi = ins_ok ? v.lb(i - kk, j) : 0;
                uint32_t pd = __shfl_up_sync(FULL, bd, 1, LPP), pi = __shfl_up_sync(FULL, bi, 1, LPP);
                if (lane == 0) { pd = brow; pi = bcol; }
                int dr = del_ok ? sext8(bd - pd) : 0, dc = ins_ok ? sext8(bi - pi) : 0;
                #pragma unroll
                for (int o = 1; o < LPP; o <<= 1) {
                    const int tr = __shfl_up_sync(FULL, dr, o, LPP), tc = __shfl_up_sync(FULL, dc, o, LPP);
                    if (lane >= o) { dr += tr; dc += tc; }
                }
                const int hr = Hrow + dr, hc = Hcol + dc;
                const unsigned md = gballot(del_ok && H == hr - gap_cost2(d_o1, d_e1, d_o2, d_e2, two, kk));
                const unsigned mi = gballot(ins_ok && H == hc - gap_cost2(i_o1, i_e1, i_o2, i_e2, two, kk));
                const unsigned mx = gballot(!del_ok && !ins_ok);
                const int fd = md ? __ffs(md) - 1 : LPP, fi = mi ? __ffs(mi) - 1 : LPP, fx = mx ? __ffs(mx) - 1 : LPP;
                const int src = min(min(fd, fi), LPP - 1);
                const int Hfd = __shfl_sync(FULL, hr, src, LPP), Hfi = __shfl_sync(FULL, hc, src, LPP);
                const int hr_l = __shfl_sync(FULL, hr, LPP - 1, LPP), hc_l = __shfl_sync(FULL, hc, LPP - 1, LPP);
                const uint32_t bd_l = __shfl_sync(FULL, bd, LPP - 1, LPP), bi_l = __shfl_sync(FULL, bi, LPP - 1, LPP);
                if (search) {
                    if (min(fd, fi) < fx) {                       // at equal length the deletion comes first
                        search = false;
                        if (fd <= fi) { const int len = k0 + fd + 1; SVP_EMITG(SVP_CIGAR_D, (uint32_t)len); ++n_do; n_db += len; j -= len; H = Hfd; }
                        else { const int len = k0 + fi + 1; SVP_EMITG(SVP_CIGAR_I, (uint32_t)len); ++n_io; n_ib += len; i -= len; H = Hfi; }
                    } else if (fx < LPP) {
                        search = false; bad = true; done = true;      // no predecessor explains H
                    } else { Hrow = hr_l; Hcol = hc_l; brow = bd_l; bcol = bi_l; }
                }
                k0 += LPP;
            }
        }
    }
    // ---- 3. the alignment starts at the lowest point of the walk ------------------------------------------------------------
    const int cut_low = low;
    i = low_i; j = low_j; wpos = low_wpos; cur_op = low_op; cur_len = low_len; ovf = low_ovf;
    if (act) SVP_PUSHG();
#undef SVP_PUSHG
#undef SVP_EMITG
    if (bad) status |= SVP_ST_SCORE_OVF;
    __syncwarp();
    // directed distance over the CIGAR the group just wrote: the lanes of the group share the runs
    const int n_runs = (act && !ovf) ? (int)tk.cigar_cap - wpos : 0;
    const int n_runs_max = __reduce_max_sync(FULL, n_runs);
    double dist = 0.0;
    for (int a = lane; a < n_runs_max; a += LPP) {
        if (a < n_runs) {
            const uint32_t w = slot[wpos + a];
            const uint32_t op = w & 15u, len = w >> 4;
            if (op != SVP_CIGAR_M && (int)len >= cs.min_signal) {
                const double sz = (double)len;
                dist += (1.0 - exp(-pow(sz / 30.0, 1.5))) * sz;
            }
        }
    }
    #pragma unroll
    for (int o = LPP / 2; o; o >>= 1) dist += __shfl_xor_sync(FULL, dist, o, LPP);
    if (act && lane == 0) {
        if (ovf) write_result(status | SVP_ST_CIGAR_OVF, M - cut_low, i, best_i, j, best_j, l_match, l_mis, l_io, l_ib, l_do, l_db, tk.cigar_off, 0, 0.0);
        else write_result(status, M - cut_low, i, best_i, j, best_j, l_match, l_mis, l_io, l_ib, l_do, l_db, tk.cigar_off + (uint32_t)wpos, tk.cigar_cap - (uint32_t)wpos, dist);
    }
}

// ---------------------------------------------------------------------------------------------
// Wave scheduling: the walks of a wave as kernels of their own, launched by the host behind the wave's forward kernels. The rows
// were written by earlier kernels, so plain (L1-cached) loads are legal here.
//   svp_tb_kernel    one THREAD per pair: the bulk walker — 32 walks per warp on one instruction stream, ~1 % of the forward
//                    pass's issue slots; latency ~10 ms per wave, hidden under the next wave's forward pass
//   svp_tbw_kernel   one WARP per pair: a thirty-second of the throughput, a quarter of the latency — small waves, long pairs,
//                    the last wave of a blocking call
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64)
svp_tb_kernel(const Task* __restrict__ tasks, const uint32_t* __restrict__ order, int n_tasks, const uint8_t* __restrict__ seqs,
              const uint8_t* __restrict__ arena, svp_result* __restrict__ results, uint32_t* __restrict__ cigar, const Consts cs) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= n_tasks) return;
    const Task tk = tasks[order ? order[tid] : (uint32_t)tid];
    svp_traceback_thread<false>(tk, arena + tk.tb_off, reinterpret_cast<const uint2*>(arena + tk.tb_off + tk.tb_bytes), seqs, results, cigar, cs);
}
__global__ void __launch_bounds__(128)
svp_tbw_kernel(const Task* __restrict__ tasks, const uint32_t* __restrict__ order, int n_tasks, const uint8_t* __restrict__ seqs,
               const uint8_t* __restrict__ arena, svp_result* __restrict__ results, uint32_t* __restrict__ cigar, const Consts cs, const int cor) {
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (wid >= n_tasks) return;
    const Task tk = tasks[order ? order[wid] : (uint32_t)wid];
    extern __shared__ __align__(16) uint8_t cor_smem[];           // COR_BYTES per warp (or none: cor = 0)
    svp_traceback_warp<false>(tk, arena + tk.tb_off, reinterpret_cast<const uint2*>(arena + tk.tb_off + tk.tb_bytes), seqs, results, cigar, cs,
                              cor ? (uint32_t)__cvta_generic_to_shared(cor_smem) + (threadIdx.x >> 5) * (uint32_t)COR_BYTES : 0u);
}


// ---------------------------------------------------------------------------------------------
// Queue scheduling: the call's walkers. One warp per CTA, a few CTAs per SM, resident for the whole call. A walker takes up to 32
// consecutive published entries off the queue and walks them one per LANE (svp_traceback_thread: 32 walks on one instruction
// stream). A walk is a chain of dependent loads — 2-4 ms for a pair of 8 kb reads whatever runs it — so what a walker costs the
// forward passes it shares its SM with is its registers (one warp of ~5 k registers walks 8-10 pairs per millisecond; config 2
// produces 2 pairs per millisecond and SM), not issue slots. The rows were written by concurrently running grids: loads
// through L2 (CG). A partial batch waits up to `wait_ns` for more entries.
// ---------------------------------------------------------------------------------------------
template <int MIN_CTAS>                                  // 1: as many registers as the walk likes (168); 16: at most 128 (one forward CTA's worth)
__global__ void __launch_bounds__(32, MIN_CTAS)
svp_walkq_kernel(const uint8_t* __restrict__ seqs, const Arena ar, svp_result* __restrict__ results, uint32_t* __restrict__ cigar, const Consts cs,
                 const uint32_t wait_ns) {
    WalkQueue* Q = ar.queue;
    const int lane = threadIdx.x;
    const uint32_t n = Q->n;
    unsigned long long t_first = 0;
    for (;;) {
        uint32_t h = 0;
        if (lane == 0) h = *(volatile uint32_t*)&Q->head;
        h = __shfl_sync(0xffffffffu, h, 0);
        if (h >= n) return;
        const bool ready = (h + (uint32_t)lane < n) && *(volatile uint32_t*)&Q->e[h + lane].task != WALK_EMPTY;
        const unsigned m = __ballot_sync(0xffffffffu, ready);
        const int cnt = m == 0xffffffffu ? 32 : __ffs(~m) - 1;                // published entries in a row from the head
        const int want = (int)min(32u, n - h);
        if (cnt == 0) { t_first = 0; __nanosleep(2000); continue; }
        if (cnt < want) {
            unsigned long long now = globaltimer_ns();
            now = __shfl_sync(0xffffffffu, now, 0);
            if (!t_first) t_first = now;
            if (now - t_first < (unsigned long long)wait_ns) { __nanosleep(1000); continue; }
        }
        t_first = 0;
        int ok = 0;
        if (lane == 0) ok = atomicCAS(&Q->head, h, h + (uint32_t)cnt) == h;
        ok = __shfl_sync(0xffffffffu, ok, 0);
        if (!ok) continue;
        __threadfence();
        const unsigned long long t0 = globaltimer_ns();
        if (lane < cnt) {
            const uint32_t task = *(volatile uint32_t*)&Q->e[h + lane].task, pool_id = *(volatile uint32_t*)&Q->e[h + lane].pool_id;
            const uint64_t o = *(volatile uint64_t*)&Q->e[h + lane].arena_off;
            const Task tk = ar.tasks[task];
            if (!ar.skip_walk) svp_traceback_thread<true>(tk, ar.base + o, reinterpret_cast<const uint2*>(ar.base + o + tk.tb_bytes), seqs, results, cigar, cs);
        }
        __syncwarp();
        for (int l = 0; l < cnt; ++l) {                        // slices back to their pools, one lane at a time (the pools' spin locks)
            if (lane == l) arena_free(ar, *(volatile uint32_t*)&Q->e[h + l].pool_id, *(volatile uint64_t*)&Q->e[h + l].arena_off);
            __syncwarp();
        }
        if (lane == 0) { atomicAdd(ar.clocks + 4, globaltimer_ns() - t0); atomicAdd(ar.clocks + 5, 1ull); atomicAdd(ar.clocks + 2, (unsigned long long)cnt); }
    }
}


//   svp_tbg_kernel   LPP lanes per pair, 32 / LPP pairs per warp (svp_traceback_warp): the instruction count per pair of the warp
//                    walk divided by the pairs that share a warp's instruction stream
template <int LPP>
__global__ void __launch_bounds__(128)
svp_tbg_kernel(const Task* __restrict__ tasks, const uint32_t* __restrict__ order, int n_tasks, const uint8_t* __restrict__ seqs,
               const uint8_t* __restrict__ arena, svp_result* __restrict__ results, uint32_t* __restrict__ cigar, const Consts cs) {
    const int gid = (int)((blockIdx.x * blockDim.x + threadIdx.x) / LPP);
    if (gid >= n_tasks) return;                    // (every cross-lane operation of the walk names its own group only)
    const Task tk = tasks[order ? order[gid] : (uint32_t)gid];
    svp_traceback_warp<false, LPP>(tk, arena + tk.tb_off, reinterpret_cast<const uint2*>(arena + tk.tb_off + tk.tb_bytes), seqs, results, cigar, cs, 0u);
}

//   svp_tbu_kernel   the same on ONE instruction stream per warp (svp_traceback_groups): the bulk walker of serial waves
#ifndef SVP_TBU_MIN_CTAS
#define SVP_TBU_MIN_CTAS 8
#endif
template <int LPP>
__global__ void __launch_bounds__(128, SVP_TBU_MIN_CTAS)
svp_tbu_kernel(const Task* __restrict__ tasks, const uint32_t* __restrict__ order, int n_tasks, const uint8_t* __restrict__ seqs,
               const uint8_t* __restrict__ arena, svp_result* __restrict__ results, uint32_t* __restrict__ cigar, const Consts cs) {
    const int gid = (int)((blockIdx.x * blockDim.x + threadIdx.x) / LPP);
    if ((int)((blockIdx.x * blockDim.x + (threadIdx.x & ~31u)) / LPP) >= n_tasks) return;       // the whole warp is beyond the list
    // all 32 lanes stay together; a group without a pair idles
    const Task* tp = gid < n_tasks ? tasks + (order ? order[gid] : (uint32_t)gid) : nullptr;
    svp_traceback_groups<LPP, false>(tp, arena + (tp ? tp->tb_off : 0), seqs, results, cigar, cs);
}


// Queue scheduling, group walkers: one warp per CTA, resident for the whole call; a walker takes up to 32 / LPP consecutive published
// entries off the queue and walks them on one instruction stream (svp_traceback_groups), one pair per group of LPP lanes. A partial
// batch waits up to `wait_ns` for more entries. Loads of the rows through L2 (written by concurrently running grids, slices recycled).
template <int LPP>
__global__ void __launch_bounds__(32, 16)
svp_walkqg_kernel(const uint8_t* __restrict__ seqs, const Arena ar, svp_result* __restrict__ results, uint32_t* __restrict__ cigar, const Consts cs,
                  const uint32_t wait_ns) {
    constexpr int NG = 32 / LPP;
    WalkQueue* Q = ar.queue;
    const int lane = threadIdx.x, grp = lane / LPP;
    const uint32_t n = Q->n;
    unsigned long long t_first = 0;
    for (;;) {
        uint32_t h = 0;
        if (lane == 0) h = *(volatile uint32_t*)&Q->head;
        h = __shfl_sync(0xffffffffu, h, 0);
        if (h >= n) return;
        const bool ready = (h + (uint32_t)grp < n) && *(volatile uint32_t*)&Q->e[h + grp].task != WALK_EMPTY;
        const unsigned m = __ballot_sync(0xffffffffu, ready);
        int cnt = 0;                                             // published entries in a row from the head (one bit per group's first lane)
        #pragma unroll
        for (int g = 0; g < NG; ++g) if (cnt == g && ((m >> (g * LPP)) & 1u)) cnt = g + 1;
        const int want = (int)min((uint32_t)NG, n - h);
        if (cnt == 0) { t_first = 0; __nanosleep(2000); continue; }
        if (cnt < want) {
            unsigned long long now = globaltimer_ns();
            now = __shfl_sync(0xffffffffu, now, 0);
            if (!t_first) t_first = now;
            if (now - t_first < (unsigned long long)wait_ns) { __nanosleep(1000); continue; }
        }
        t_first = 0;
        int ok = 0;
        if (lane == 0) ok = atomicCAS(&Q->head, h, h + (uint32_t)cnt) == h;
        ok = __shfl_sync(0xffffffffu, ok, 0);
        if (!ok) continue;
        __threadfence();
        const unsigned long long t0 = globaltimer_ns();
        const bool mine = grp < cnt;
        const uint32_t task = mine ? *(volatile uint32_t*)&Q->e[h + grp].task : 0u;
        const uint32_t pool_id = mine ? *(volatile uint32_t*)&Q->e[h + grp].pool_id : 0u;
        const uint64_t o = mine ? *(volatile uint64_t*)&Q->e[h + grp].arena_off : 0ull;
        if (!ar.skip_walk) svp_traceback_groups<LPP, true>(mine ? ar.tasks + task : nullptr, ar.base + o, seqs, results, cigar, cs);
        __syncwarp();
        for (int g = 0; g < cnt; ++g) {                          // slices back to their pools, one group at a time (the pools' spin locks)
            if (grp == g && (lane % LPP) == 0) arena_free(ar, pool_id, o);
            __syncwarp();
        }
        if (lane == 0) { atomicAdd(ar.clocks + 4, globaltimer_ns() - t0); atomicAdd(ar.clocks + 5, 1ull); atomicAdd(ar.clocks + 2, (unsigned long long)cnt); }
    }
}

}  // namespace svp
