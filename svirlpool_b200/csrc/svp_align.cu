// svp_align.cu — C-ABI of the alignment hot path (include/svp_align.h) on top of the sm_100a
// kernels in svp_kernels.cuh. Host side: validate, plan every pair (kernel geometry, slice of traceback rows), group the
// pairs by kernel variant, schedule them (SVP_MODE: waves — the host assigns the slices, forward kernels and walk kernels
// alternate; arena / queue — device-side slice allocator, in-kernel or resident walkers), and time the call with CUDA events
// on the launching stream. No CPU fallback: without a CUDA device every entry point that computes returns SVP_ERR_NODEVICE.
#include <algorithm>
#include <chrono>
#include <thread>
#include <array>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "svp_kernels_wide.cuh"

using namespace svp;

namespace {

// Kernel widths: KP packed registers per state array per thread = 4*KP diagonals per thread, 128*KP per
// warp (KP = 4 .. 8 -> 512, 640, 768, 896, 1024). A pair runs on the (KP, warps) combination that needs
// the fewest issue slots for its band (choose_kp); single-warp CTAs need no mailbox. With only KP = 4 / 8 a band had
// to be padded to 512 diagonals per warp: on config 2 (bands |dlen| + 600, median need 768) that was 30 % more cells
// than the spec asks for.
constexpr int S16_SCORE_LIMIT = 32000;
static inline int max_warps_of(int kp) { return kp <= 8 ? 16 : 8; }
// warp-instructions per step of the inner loop (SASS counts), used only to rank variants
static inline int step_cost(int kp, int nwarp) { return 22 + kp * 25 / 2 + (nwarp > 1 ? 10 : 0); }
// latency = true (a call of few pairs, e.g. one container: the GPU is not full): the variant with the shortest step, i.e. more,
// narrower warps per pair, instead of the one with the fewest issue slots.
// kp_mask (env SVP_KP_MASK, tuning / test aid): bit (KP - 4) allows that KP; kp_multi (env SVP_KP_MULTI): KP = 5 / 6 / 7 on multi-warp CTAs too
static inline int choose_kp(uint32_t band_w, bool latency, bool two_piece, bool plain_s16, bool wide_ok, int kp_mask, bool kp_multi) {
    int best_kp = 0; long best_cost = 0;
    for (int kp : { 4, 5, 6, 7, 8, 12 }) {
        if (!(kp_mask & (1 << (kp - 4)))) continue;
        // KP = 12: single-warp CTAs of wave scheduling, plain s16 range, two-piece scoring (the only instantiations)
        if (kp == 12 && !(wide_ok && plain_s16 && two_piece && band_w <= 1536)) continue;
        if ((kp & 3) && !two_piece) continue;          // one-piece simple scoring: KP = 4 / 8 only (fewer instantiations)
        if (band_w % (uint32_t)(4 * kp)) continue;
        const int nw = (int)((band_w + 128 * kp - 1) / (128 * kp));
        if (nw > max_warps_of(kp)) continue;
        // KP = 5 / 6 / 7 for single-warp CTAs only: they exist to fit a band of 640 / 768 / 896 diagonals into one warp without padding it
        // to 1024. On multi-warp CTAs they only multiply the launch groups of a wave (config 3: 218 instead of 104 launches per 6 steps,
        // 977 instead of 1162 GCUPS; config 4: 653 instead of 696 loci/s)
        // ... and for pairs inside the s16 score range only: on the long reads of config 3 the rebased KP = 5 / 6 / 7 variants cost more
        // than they save (1058 against 1162 GCUPS with KP = 4 / 8 alone)
        if ((kp & 3) && (nw > 1 || !plain_s16) && !kp_multi) continue;
        if (kp == 12 && nw > 1) continue;
        const long cost = latency ? (long)step_cost(kp, nw) : (long)nw * step_cost(kp, nw);
        if (!best_kp || cost < best_cost) { best_kp = kp; best_cost = cost; }
    }
    return best_kp;          // 0: band too wide for every variant
}
static inline int kp_of(const Task& t) { return (int)((t.flags & TASK_KP_MASK) >> TASK_KP_SHIFT); }
static inline int cluster_of(const Task& t) { return std::max(1, (int)((t.flags & TASK_CL_MASK) >> TASK_CL_SHIFT)); }
constexpr int MAX_CLUSTER = 8;                  // portable cluster size limit
// CTAs (16 warps each) a band needs with `diag_per_warp` diagonals per warp, rounded up to a power of two; 0: too wide
static inline int cluster_for(uint32_t band_w, int diag_per_warp) {
    const int64_t per_cta = 16 * (int64_t)diag_per_warp;
    int c = 1;
    while ((int64_t)c * per_cta < (int64_t)band_w) c *= 2;
    return c <= MAX_CLUSTER ? c : 0;
}
// steps per loop iteration of the kernel the pair runs on: 2*KP (s16 kernels), 16 (32-bit lane kernel, KP = 4 geometry)
static inline int steps_per_iter_of(const Task& t) { return (t.flags & TASK_S32) ? 16 : 2 * kp_of(t); }
// traceback rows of a pair: one row per step pair, one slot per thread (svp_kernels.cuh: store_pair)
static inline uint64_t tb_bytes_of(const Task& t) {
    return (uint64_t)t.n_iter * (steps_per_iter_of(t) / 2) * (uint64_t)(t.band_w / (4 * kp_of(t))) * (uint64_t)tb_slot_bytes(kp_of(t));
}
static inline uint64_t best_entries_of(const Task& t) { return (uint64_t)(t.band_w / (4 * kp_of(t))) * 2; }

// dynamic shared memory of a pair's CTA (nw warps): sequence images (whole, or the sliding windows of the RING variants) + mailboxes
static inline size_t smem_need(const Task& t, const Consts& cs, int nw) {
    const size_t n_qimg = cs.general ? 2 : 1;              // general scoring stages two query images
    if (t.flags & TASK_RING) {
        const int per_thread = (t.flags & TASK_S32) || cs.general ? 8 : 2 * kp_of(t);      // positions per thread
        return (n_qimg + 1) * (size_t)ring_bytes(nw * 32, per_thread) + (size_t)mailbox_bytes(nw);
    }
    return n_qimg * (size_t)seq_image_bytes(t.m) + (size_t)seq_image_bytes(t.n) + (size_t)mailbox_bytes(nw);
}

struct DevBuf {
    void* p = nullptr; size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

}  // namespace

// Per-call state. Two sets alternate, so a submitted call (svp_submit_batch_device) can still be running while the next call
// is planned and its kernels are enqueued behind / beside it.
struct CallSet {
    DevBuf d_tasks, d_order, d_queue;
    uint32_t n_walkers = 0;                                     // queue scheduling: walker warps of this call
    Task* h_tasks = nullptr; size_t h_tasks_cap = 0;            // pinned staging of the task / order arrays
    uint32_t* h_order = nullptr; size_t h_order_cap = 0;
    std::vector<cudaEvent_t> events;                            // pool, grown on demand (stream joins)
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;               // first enqueue on the main stream / last kernel done
    bool pending = false;                                       // enqueued, times not collected yet
    size_t n_waves = 0;                                         // wave scheduling: waves of this call (their events: wave_events)
    std::vector<std::pair<size_t, size_t>> wave_ranges;        // ... their [first, last) task ranges (input order)
    std::vector<cudaEvent_t> wave_events;                       // timing events, EV_PER_WAVE per wave
    svp_stats stats{};                                          // counts filled at submit, times at collection
};

constexpr int N_FWD_STREAMS = 8;
constexpr size_t EV_PER_WAVE = 4 + (N_FWD_STREAMS - 1);         // [fwd start, fwd end, walks end, long walks end, stream joins ...]

struct svp_handle {
    int device = 0;
    svp_scoring scoring{};
    Consts cs{};
    cudaStream_t stream = nullptr;              // copies, joins, the first kernel group of a call
    cudaStream_t stream_aux[N_FWD_STREAMS - 1] = {};   // the other groups of a call run beside it
    // Scheduling (env SVP_MODE):
    //   waves (default)  the host gives every pair of a wave its slice of the arena (halves alternate from wave to wave) and launches
    //                    the wave's walks as bulk kernels under the next wave's forward pass: highest throughput, arena = two waves
    //   arena            every CTA takes its slice from the device-side pools and walks its pair itself: no waves, a bounded arena
    //                    (the pairs resident on the GPU), ~15 % slower on config 2
    //   queue            arena scheduling, but a CTA ends with its forward pass and queues its pair for the call's resident walker
    //                    kernel (one walk per lane)
    bool waves = true;
    bool queue = false;
    cudaStream_t stream_copy = nullptr;         // host-buffer calls: results and CIGARs of a finished wave go home under the next wave
    cudaStream_t stream_walk = nullptr;         // queue: the call's walker kernel (highest priority: resident before the forward CTAs flood the SMs)
    int walk_ctas_per_sm = 4;                   // queue: walker CTAs (one warp each) per SM and call (env SVP_WALK_CTAS_PER_SM)
    long walk_wait_ns = 100000;                 // queue: how long a walker waits for a full batch of 32 (env SVP_WALK_WAIT_NS)
    long queue_min_pairs = 2048;                // queue: calls of fewer pairs walk inside their CTAs (a warp per walk: a third of the latency)
    cudaStream_t stream_tb = nullptr;           // waves: the walk kernels (high priority; they overlap the next wave's forward pass)
    cudaStream_t stream_tb2 = nullptr;          // ... the warp-per-pair walks of a wave's long pairs, beside the bulk kernel
    long tbw_long = 24000;                      // waves: pairs with m + n beyond this are walked by a warp each (env SVP_TBW_LONG)
    int tbw_max = 2048;                         // waves of at most this many pairs use the warp-per-pair walk (env SVP_TBW_MAX)
    uint64_t wave_idx = 0;                      // waves enqueued so far: wave w uses arena half (w & 1), across calls
    cudaEvent_t half_free[2] = { nullptr, nullptr };   // walks-done event of the last wave that used each arena half
    int sm_count = 0;
    cudaEvent_t marks[2] = { nullptr, nullptr };   // svp_timer_mark
    // arena: n_sm per-SM pools of sm_region bytes each, then one global pool (svp_kernels.cuh: ArenaPool)
    uint8_t* arena = nullptr; size_t arena_bytes = 0;
    ArenaPool* d_pools = nullptr;
    uint32_t n_sm = 0;                          // %nsmid of the device
    uint64_t sm_region = 0, global_region = 0;
    double global_frac = -1.0;                  // share of the arena given to the global pool (set_partition)
    unsigned long long* d_clocks = nullptr;     // summed forward ns / traceback ns / pairs over all CTAs
    unsigned long long clocks_seen[6] = { 0, 0, 0, 0, 0, 0 };
    int smem_per_sm = 233472;
    CallSet sets[2];
    uint64_t call_idx = 0;                      // calls planned so far (set = call_idx & 1)
    DevBuf d_seqs, d_results, d_cigar, d_dense, d_dense_off;
    svp_stats stats{};                          // the last collected call
    svp_stats stats_total{};                    // sums over collected calls since the last reset (svp_get_stats_total)
    std::string err;
    int max_smem_optin = 0;
    long ring_min = 4096;                       // fast path: sliding windows when whole images exceed this many bytes per warp of the CTA (env SVP_RING_MIN)
    long ring_min_single = 4096;                // ... of a single-warp CTA (env SVP_RING_MIN_SINGLE). Whole images for 8 kb pairs (20000) were measured
                                                // within +-2 % on config 2 and 6 % slower on config 4 (more launch groups): windows stay
    long small_call_pairs = 1024;               // calls of at most this many pairs use the KP = 4 / 8 widths only (env SVP_SMALL_CALL_PAIRS; plan_pair)
    long latency_pairs = 0;                     // calls of at most this many pairs are planned for short steps (more, narrower warps) instead of fewest
                                                // issue slots (env SVP_LATENCY_PAIRS). Off by default: measured SLOWER for one container per call
    int kp_mask = 31; bool kp_multi = false;      // kernel widths the planner may use (env SVP_KP_MASK, SVP_KP_MULTI: choose_kp); KP = 12 is off by default
                                                  // (bit 8): one 220-register warp for a band of 1025..1536 diagonals measured SLOWER than three KP = 4 warps (config 2: 538 against 565 regions/s)
    bool force_ring = false;                    // test aid (env SVP_FORCE_RING): every pair on a sliding-window variant
};

static thread_local std::string g_err;

static int fail(svp_handle* h, int code, const std::string& msg) {
    if (h) h->err = msg; else g_err = msg;
    return code;
}
#define CUDA_TRY(h, x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return fail(h, e_ == cudaErrorMemoryAllocation ? SVP_ERR_NOMEM : SVP_ERR_CUDA, std::string(#x) + ": " + cudaGetErrorString(e_)); } while (0)

extern "C" int svp_abi_version(void) { return SVP_ABI_VERSION; }

extern "C" int svp_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" int svp_scoring_preset(const char* name, svp_scoring* s) {
    if (!name || !s) return SVP_ERR_ARG;
    memset(s, 0, sizeof(*s));
    s->min_signal_size = 12;
    const std::string nm(name);
    if (nm == "map-ont" || nm == "ava-ont") {          // minimap2 defaults -A2 -B4 -O4,24 -E2,1 -z400
        for (int a = 0; a < 4; ++a) for (int b = 0; b < 4; ++b) s->matrix[a * 4 + b] = (a == b) ? 2 : -4;
        s->del_open1 = s->ins_open1 = 4;  s->del_ext1 = s->ins_ext1 = 2;
        s->del_open2 = s->ins_open2 = 24; s->del_ext2 = s->ins_ext2 = 1;
        s->zdrop = 400; s->zdrop_ext = 1;
        return SVP_OK;
    }
    if (nm == "last") {                                // shape of a last-train .mat; stand-in numbers
        static const int m[16] = { 5, -8, -4, -9,   -8, 6, -9, -4,   -4, -9, 6, -8,   -9, -4, -8, 5 };
        memcpy(s->matrix, m, sizeof(m));
        s->del_open1 = 10; s->del_ext1 = 3; s->ins_open1 = 12; s->ins_ext1 = 2;
        s->del_open2 = s->ins_open2 = -1;
        s->zdrop = 400; s->zdrop_ext = 2;
        return SVP_OK;
    }
    return SVP_ERR_ARG;
}

// Scoring the kernels implement. Fast path (svp_fwd_kernel): the matrix is {A on the diagonal, -B elsewhere} and
// insertion costs == deletion costs. Everything else — a full 4x4 matrix (|entry| <= 127), separate insertion /
// deletion costs — runs on the general kernels. Common limit: the traceback stores one byte per cell (H mod 256),
// so H of cells one or two steps apart must differ by less than 128: 2 * (max entry + dearest 1-base gap) <= 127
// (horizontally / vertically adjacent in-band cells differ by at most max entry + 1-base gap).
static int make_consts(const svp_scoring& s, Consts& c, std::string& why) {
    memset(&c, 0, sizeof(c));
    int a_max = -(1 << 30);
    for (int k = 0; k < 16; ++k) {
        if (s.matrix[k] < -127 || s.matrix[k] > 127) { why = "substitution scores must fit a signed byte"; return SVP_ERR_ARG; }
        a_max = std::max(a_max, s.matrix[k]);
    }
    if (a_max <= 0) { why = "the substitution matrix needs a positive entry"; return SVP_ERR_ARG; }
    const bool two_d = s.del_open2 >= 0, two_i = s.ins_open2 >= 0;
    if (s.del_open1 < 0 || s.del_ext1 <= 0 || s.ins_open1 < 0 || s.ins_ext1 <= 0 || (two_d && s.del_ext2 <= 0) || (two_i && s.ins_ext2 <= 0)) {
        why = "gap costs must be open >= 0, ext > 0"; return SVP_ERR_ARG; }
    // a kind with one piece repeats it as its second piece when the other kind has two
    c.two_piece = (two_d || two_i) ? 1 : 0;
    c.d_open1 = s.del_open1; c.d_ext1 = s.del_ext1; c.d_open2 = two_d ? s.del_open2 : s.del_open1; c.d_ext2 = two_d ? s.del_ext2 : s.del_ext1;
    c.i_open1 = s.ins_open1; c.i_ext1 = s.ins_ext1; c.i_open2 = two_i ? s.ins_open2 : s.ins_open1; c.i_ext2 = two_i ? s.ins_ext2 : s.ins_ext1;
    const int g1d = std::min(c.d_open1 + c.d_ext1, c.d_open2 + c.d_ext2), g1i = std::min(c.i_open1 + c.i_ext1, c.i_open2 + c.i_ext2);
    if (2 * (a_max + std::max(g1d, g1i)) > 127) {
        why = "scores too large for the one-byte traceback (2*(max matrix entry + dearest 1-base gap) must be <= 127)"; return SVP_ERR_UNSUPPORTED; }
    if (std::max(std::max(c.d_open1 + c.d_ext1, c.d_open2 + c.d_ext2), std::max(c.i_open1 + c.i_ext1, c.i_open2 + c.i_ext2)) > 8000) {
        why = "gap costs too large for 16-bit lanes"; return SVP_ERR_UNSUPPORTED; }
    const int A = s.matrix[0], B = -s.matrix[1];
    bool simple = A > 0 && B > 0 && s.del_open1 == s.ins_open1 && s.del_ext1 == s.ins_ext1 && two_d == two_i &&
                  (!two_d || (s.del_open2 == s.ins_open2 && s.del_ext2 == s.ins_ext2));
    for (int a = 0; a < 4; ++a) for (int b = 0; b < 4; ++b) if (s.matrix[a * 4 + b] != (a == b ? A : -B)) simple = false;
    if (const char* e = getenv("SVP_FORCE_GENERAL")) if (atoi(e)) simple = false;     // test aid: run simple scoring on the general kernels
    c.general = simple ? 0 : 1;
    auto pk16 = [](int v) { return (uint32_t)(v & 0xFFFF) * 0x00010001u; };
    c.mat = pk16(A); c.mis = pk16(-B); c.s_match = A; c.s_mismatch = -B;
    c.ne1 = pk16(-c.d_ext1); c.noe1 = pk16(-(c.d_open1 + c.d_ext1));
    c.ne2 = c.two_piece ? pk16(-c.d_ext2) : 0; c.noe2 = c.two_piece ? pk16(-(c.d_open2 + c.d_ext2)) : NEGV;
    c.d_ne1 = pk16(-c.d_ext1); c.d_noe1 = pk16(-(c.d_open1 + c.d_ext1)); c.d_ne2 = pk16(-c.d_ext2); c.d_noe2 = pk16(-(c.d_open2 + c.d_ext2));
    c.i_ne1 = pk16(-c.i_ext1); c.i_noe1 = pk16(-(c.i_open1 + c.i_ext1)); c.i_ne2 = pk16(-c.i_ext2); c.i_noe2 = pk16(-(c.i_open2 + c.i_ext2));
    for (int q = 0; q < 4; ++q) for (int t = 0; t < 4; ++t) c.tab[q] |= (uint32_t)(s.matrix[q * 4 + t] & 0xFF) << (8 * t);
    c.a_max = a_max;
    c.n_score = s.matrix[0];                     // a base outside ACGT scores the lowest entry of the matrix against everything
    for (int k = 1; k < 16; ++k) c.n_score = std::min(c.n_score, s.matrix[k]);
    c.n_pk = pk16(c.n_score);
    // rebased s16x2 kernel (scores beyond the s16 range on the fast path): the values of a warp, relative to its offset, must
    // stay inside +-RB_CLAMP between two rebases: (128*KP diagonals + RB_ITERS*2*KP steps) * Lip + 256 (svp_kernels.cuh)
    const int lip = a_max + std::max(g1d, g1i);
    c.rebase_ok = (simple && (128 * 8 + RB_ITERS * 16) * lip + 256 <= RB_CLAMP - 400) ? 1 : 0;
    if (const char* e = getenv("SVP_NO_REBASE")) if (atoi(e)) c.rebase_ok = 0;          // test / tuning aid: long reads on the 32-bit lane kernel
    c.min_signal = s.min_signal_size; c.zdrop = s.zdrop; c.zdrop_ext = s.zdrop_ext;
    return SVP_OK;
}

// tuning / test knobs read from the environment when a handle (or a planning-only pseudo handle) is set up
static void read_env_knobs(svp_handle* h) {
    if (const char* e = getenv("SVP_RING_MIN")) h->ring_min = h->ring_min_single = atol(e);
    if (const char* e = getenv("SVP_RING_MIN_SINGLE")) h->ring_min_single = atol(e);
    if (const char* e = getenv("SVP_FORCE_RING")) h->force_ring = atoi(e) != 0;
    if (const char* e = getenv("SVP_LATENCY_PAIRS")) h->latency_pairs = atol(e);
    if (const char* e = getenv("SVP_SMALL_CALL_PAIRS")) h->small_call_pairs = atol(e);
    if (const char* e = getenv("SVP_KP_MASK")) { h->kp_mask = atoi(e); if (!(h->kp_mask & 0x11F)) h->kp_mask = 31; }
    if (const char* e = getenv("SVP_KP_MULTI")) h->kp_multi = atoi(e) != 0;
    if (const char* e = getenv("SVP_MODE")) { const std::string m(e); h->waves = m == "waves"; h->queue = m == "queue"; }
    if (const char* e = getenv("SVP_WALK_CTAS_PER_SM")) h->walk_ctas_per_sm = std::min(std::max(atoi(e), 1), 8);
    if (const char* e = getenv("SVP_WALK_WAIT_NS")) h->walk_wait_ns = std::max(atol(e), 0L);
    if (const char* e = getenv("SVP_QUEUE_MIN_PAIRS")) h->queue_min_pairs = atol(e);
    if (const char* e = getenv("SVP_TBW_LONG")) h->tbw_long = atol(e);
    if (const char* e = getenv("SVP_TBW_MAX")) h->tbw_max = atoi(e);
}

// kernel family of a pair
enum { FAM_FAST = 0, FAM_GEN = 1, FAM_S32 = 2, FAM_S32_GEN = 3, FAM_FAST_RB = 4 };

// Every forward kernel has the same signature; the variant a group of pairs runs on is looked up here (svp_create sets the
// attributes of every entry, run_plan launches through it). nullptr: no such variant.
constexpr int SMEM_STATIC = 64;          // static shared memory of the forward kernels (the pair's arena slice note), rounded up
using FwdKernel = void (*)(const Task*, const uint32_t*, const uint8_t*, const Arena, svp_result*, uint32_t*, const Consts);
template <bool W>
static FwdKernel fwd_kernel_w(int fam, int kp, bool multi, bool cluster, bool ring, bool two) {
    if (fam == FAM_FAST || fam == FAM_FAST_RB) {        // s16x2 fast path: plain / rebased, whole images / sliding windows, one CTA / cluster
        const bool rb = fam == FAM_FAST_RB;
        if (cluster) {
            if (kp != 8) return nullptr;
#define SVP_FCL(T, R, G) if (two == T && rb == R && ring == G) return svp_fwd_kernel<8, true, T, true, R, G, W>;
            SVP_FCL(true, false, false) SVP_FCL(true, false, true) SVP_FCL(true, true, false) SVP_FCL(true, true, true)
            SVP_FCL(false, false, false) SVP_FCL(false, false, true) SVP_FCL(false, true, false) SVP_FCL(false, true, true)
#undef SVP_FCL
            return nullptr;
        }
#define SVP_F(K, M, T, R, G) if (kp == K && multi == M && two == T && rb == R && ring == G) return svp_fwd_kernel<K, M, T, false, R, G, W>;
#define SVP_F4(K, M, T) SVP_F(K, M, T, false, false) SVP_F(K, M, T, false, true) SVP_F(K, M, T, true, false) SVP_F(K, M, T, true, true)
        SVP_F4(4, false, true) SVP_F4(4, true, true) SVP_F4(8, false, true) SVP_F4(8, true, true)
        SVP_F4(5, false, true) SVP_F4(5, true, true) SVP_F4(6, false, true) SVP_F4(6, true, true) SVP_F4(7, false, true) SVP_F4(7, true, true)
        SVP_F4(4, false, false) SVP_F4(4, true, false) SVP_F4(8, false, false) SVP_F4(8, true, false)
        if constexpr (!W) { SVP_F(12, false, true, false, false) SVP_F(12, false, true, false, true) }
#undef SVP_F4
#undef SVP_F
        return nullptr;
    }
    if (kp != 4) return nullptr;                        // general scoring and 32-bit lanes: KP = 4 geometry
    if (fam == FAM_GEN) {                               // whole images only
        if (ring) return nullptr;
#define SVP_G(M, T, C) if (multi == M && two == T && cluster == C) return svp_fwdg_kernel<M, T, C, W>;
        SVP_G(false, true, false) SVP_G(true, true, false) SVP_G(false, false, false) SVP_G(true, false, false) SVP_G(true, true, true) SVP_G(true, false, true)
#undef SVP_G
        return nullptr;
    }
    const bool gen = fam == FAM_S32_GEN;
#define SVP_W(M, T, G, C, R) if (multi == M && two == T && gen == G && cluster == C && ring == R) return svp_fwd32_kernel<M, T, G, C, R, W>;
#define SVP_W2(M, C, R) SVP_W(M, true, false, C, R) SVP_W(M, false, false, C, R) SVP_W(M, true, true, C, R) SVP_W(M, false, true, C, R)
    SVP_W2(false, false, false) SVP_W2(true, false, false) SVP_W2(true, true, false)
    SVP_W2(false, false, true) SVP_W2(true, false, true) SVP_W2(true, true, true)
#undef SVP_W2
#undef SVP_W
    return nullptr;
}
// walk = false: the variants without the in-kernel walk (wave scheduling)
static FwdKernel fwd_kernel(int fam, int kp, bool multi, bool cluster, bool ring, bool two, bool walk) {
    return walk ? fwd_kernel_w<true>(fam, kp, multi, cluster, ring, two) : fwd_kernel_w<false>(fam, kp, multi, cluster, ring, two);
}

__global__ void svp_queue_init_kernel(WalkQueue* q, uint32_t n) { q->head = 0; q->tail = 0; q->n = n; q->pad = 0; }

__global__ void svp_nsmid_kernel(uint32_t* out) { uint32_t n; asm volatile("mov.u32 %0, %%nsmid;" : "=r"(n)); *out = n; }

// result rows of the pairs the planner rejected (no CTA runs for them)
__global__ void svp_invalid_kernel(const Task* __restrict__ tasks, const uint32_t* __restrict__ order, int n, svp_result* __restrict__ results) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const Task tk = tasks[order[k]];
    svp_result res;
    memset(&res, 0, sizeof(res));
    res.cigar_off = tk.cigar_off;
    res.status = (tk.flags & TASK_TOO_LONG) ? SVP_ST_TOO_LONG : SVP_ST_BAD_PAIR;
    results[tk.pair_index] = res;
}

// Split the arena into n_sm per-SM regions and one global region of `global_frac` of it. The device must be idle.
static int set_partition(svp_handle* h, double global_frac) {
    const uint64_t total = h->arena_bytes & ~(uint64_t)255;
    uint64_t glob = (uint64_t)((double)total * global_frac) & ~(uint64_t)255;
    const uint64_t region = ((total - glob) / h->n_sm) & ~(uint64_t)255;
    glob = total - region * h->n_sm;
    std::vector<ArenaPool> pools(h->n_sm + 1);
    memset(pools.data(), 0, pools.size() * sizeof(ArenaPool));
    for (uint32_t k = 0; k < h->n_sm; ++k) { pools[k].base = (uint64_t)k * region; pools[k].bytes = region; }
    pools[h->n_sm].base = (uint64_t)h->n_sm * region; pools[h->n_sm].bytes = glob;
    CUDA_TRY(h, cudaMemcpy(h->d_pools, pools.data(), pools.size() * sizeof(ArenaPool), cudaMemcpyHostToDevice));
    h->sm_region = region; h->global_region = glob; h->global_frac = global_frac;
    return SVP_OK;
}

extern "C" int svp_create(int device, const svp_scoring* scoring, size_t max_tb_bytes, svp_handle** out) {
    if (!scoring || !out) return fail(nullptr, SVP_ERR_ARG, "null argument");
    *out = nullptr;
    int ndev = svp_device_count();
    if (ndev <= 0) return fail(nullptr, SVP_ERR_NODEVICE, "no CUDA device visible; svirlpool_b200 has no CPU fallback");
    if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) device = 0; }
    if (device >= ndev) return fail(nullptr, SVP_ERR_ARG, "device index out of range");
    Consts cs{};
    std::string why;
    int rc = make_consts(*scoring, cs, why);
    if (rc != SVP_OK) return fail(nullptr, rc, why);
    svp_handle* h = new svp_handle();
    h->device = device; h->scoring = *scoring; h->cs = cs;
    auto bail = [&](int code, const std::string& m) { g_err = m; svp_destroy(h); return code; };
    if (cudaSetDevice(device) != cudaSuccess) return bail(SVP_ERR_CUDA, "cudaSetDevice failed");
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) return bail(SVP_ERR_CUDA, "cudaStreamCreate failed");
    for (cudaStream_t& sx : h->stream_aux)
        if (cudaStreamCreateWithFlags(&sx, cudaStreamNonBlocking) != cudaSuccess) return bail(SVP_ERR_CUDA, "cudaStreamCreate failed");
    {
        int prio_lo = 0, prio_hi = 0;
        cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
        if (cudaStreamCreateWithPriority(&h->stream_tb, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
            cudaStreamCreateWithPriority(&h->stream_tb2, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
            cudaStreamCreateWithPriority(&h->stream_walk, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
            cudaStreamCreateWithFlags(&h->stream_copy, cudaStreamNonBlocking) != cudaSuccess) return bail(SVP_ERR_CUDA, "cudaStreamCreate failed");
    }
    read_env_knobs(h);
    cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device);
    cudaDeviceGetAttribute(&h->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    cudaDeviceGetAttribute(&h->smem_per_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, device);
    // every variant: the largest dynamic shared memory (sequence images, or the residency ballast of run_plan) and the whole
    // L1 / shared-memory array as shared memory (the kernels stream their stores and read the rows through L2)
    for (int fam = 0; fam < 5; ++fam) for (int kp = 4; kp <= 12; ++kp) for (int m = 0; m < 2; ++m) for (int c = 0; c < 2; ++c)
        for (int r = 0; r < 2; ++r) for (int t = 0; t < 2; ++t) for (int w = 0; w < 2; ++w)
            if (FwdKernel k = fwd_kernel(fam, kp, m != 0, c != 0, r != 0, t != 0, w != 0)) {
                // (dynamic + the kernels' few bytes of static shared memory must stay within the per-block limit)
                if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_smem_optin - SMEM_STATIC) != cudaSuccess ||
                    cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared) != cudaSuccess) {
                    cudaGetLastError();
                    return bail(SVP_ERR_CUDA, "cudaFuncSetAttribute failed");
                }
            }
    // The walkers share their SMs with forward CTAs, and an SM's L1 / shared-memory split cannot change while a CTA is resident:
    // a walker launched with the default split (all L1) would keep every forward CTA that needs shared memory off its SM —
    // with a walker on all 148 SMs nothing would ever run. Same split for both.
    if (cudaFuncSetAttribute(svp_walkq_kernel<1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared) != cudaSuccess ||
        cudaFuncSetAttribute(svp_walkq_kernel<16>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared) != cudaSuccess ||
        cudaFuncSetAttribute(svp_walkqg_kernel<8>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared) != cudaSuccess) {
        cudaGetLastError();
        return bail(SVP_ERR_CUDA, "cudaFuncSetAttribute failed");
    }
    // %nsmid: the range of the SM ids the per-SM arena pools are indexed with
    uint32_t* d_n = nullptr;
    if (cudaMalloc((void**)&d_n, sizeof(uint32_t)) != cudaSuccess) return bail(SVP_ERR_NOMEM, "cudaMalloc failed");
    svp_nsmid_kernel<<<1, 1, 0, h->stream>>>(d_n);
    uint32_t nsm = 0;
    const cudaError_t e_n = cudaMemcpyAsync(&nsm, d_n, sizeof(nsm), cudaMemcpyDeviceToHost, h->stream);
    const cudaError_t e_s = cudaStreamSynchronize(h->stream);
    cudaFree(d_n);
    if (e_n != cudaSuccess || e_s != cudaSuccess || nsm == 0) return bail(SVP_ERR_CUDA, "cannot query the SM count");
    h->n_sm = nsm;
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return bail(SVP_ERR_CUDA, "cudaMemGetInfo failed");
    // Default arena: 64 GB, at most 60 % of the free memory — bounded, so that several handles / processes share a device (the
    // reference runs several consensus jobs per node, main.smk:649-707). Wave scheduling runs faster with more: a wave is what fits
    // the arena (serial walks) or half of it, and every wave boundary drains the machine (config 2, 96 regions per call: 64 / 107 /
    // 150 / 165 GB -> 6 / 4 / 3 / 2 waves -> ~520 / 545 / 565 / 577 regions/s); a caller that owns the GPU passes max_tb_bytes
    // (bench.py: 92 % of the free memory). Arena / queue scheduling hold the rows of the pairs RESIDENT on the GPU only.
    size_t want = max_tb_bytes ? max_tb_bytes : std::min<size_t>((size_t)64 << 30, free_b / 10 * 6);
    if (!max_tb_bytes) if (const char* e = getenv("SVP_ARENA_GB")) want = std::min<size_t>((size_t)std::max(atoi(e), 1) << 30, free_b / 20 * 19);   // tuning aid
    want = (want + 255) & ~(size_t)255;
    // the default size shrinks until the allocation succeeds (fragmentation, another process on the device); an explicit size does not
    while (cudaMalloc((void**)&h->arena, want) != cudaSuccess) {
        cudaGetLastError();
        h->arena = nullptr;
        if (max_tb_bytes || want < ((size_t)1 << 30)) return bail(SVP_ERR_NOMEM, "cannot allocate the traceback arena");
        want = (want / 4 * 3 + 255) & ~(size_t)255;
    }
    h->arena_bytes = want;
    if (cudaMalloc((void**)&h->d_pools, (h->n_sm + 1) * sizeof(ArenaPool)) != cudaSuccess) return bail(SVP_ERR_NOMEM, "cudaMalloc failed");
    if (cudaMalloc((void**)&h->d_clocks, 8 * sizeof(unsigned long long)) != cudaSuccess) return bail(SVP_ERR_NOMEM, "cudaMalloc failed");
    cudaMemset(h->d_clocks, 0, 8 * sizeof(unsigned long long));
    if (set_partition(h, 0.25) != SVP_OK) return bail(SVP_ERR_CUDA, "cannot initialise the arena pools");
    *out = h;
    return SVP_OK;
}

extern "C" void svp_destroy(svp_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    if (h->arena) cudaFree(h->arena);
    if (h->d_pools) cudaFree(h->d_pools);
    if (h->d_clocks) cudaFree(h->d_clocks);
    for (DevBuf* b : { &h->d_seqs, &h->d_results, &h->d_cigar, &h->d_dense, &h->d_dense_off }) b->release();
    for (CallSet& c : h->sets) {
        for (DevBuf* b : { &c.d_tasks, &c.d_order, &c.d_queue }) b->release();
        for (cudaEvent_t e : c.wave_events) cudaEventDestroy(e);
        if (c.h_tasks) cudaFreeHost(c.h_tasks);
        if (c.h_order) cudaFreeHost(c.h_order);
        for (cudaEvent_t e : c.events) cudaEventDestroy(e);
        for (cudaEvent_t e : { c.ev_t0, c.ev_t1 }) if (e) cudaEventDestroy(e);
    }
    for (cudaEvent_t e : h->marks) if (e) cudaEventDestroy(e);
    for (cudaStream_t sx : h->stream_aux) if (sx) cudaStreamDestroy(sx);
    if (h->stream) cudaStreamDestroy(h->stream);
    for (cudaStream_t sx : { h->stream_tb, h->stream_tb2, h->stream_walk, h->stream_copy }) if (sx) cudaStreamDestroy(sx);
    delete h;
}

extern "C" void* svp_pinned_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}

extern "C" void svp_pinned_free(void* p) { if (p) cudaFreeHost(p); }

extern "C" const char* svp_last_error(const svp_handle* h) { return h ? h->err.c_str() : g_err.c_str(); }

// wait for a submitted call and collect its times
static int collect_call(svp_handle* h, CallSet& c) {
    if (!c.pending) return SVP_OK;
    c.pending = false;
    // wait for the call with a deadline (env SVP_CALL_TIMEOUT_S, default 900): kernels that cannot finish — a bug, or an arena far
    // too small for the batch — are reported with the state of the walk queue and the arena pools instead of blocking forever
    {
        static const double limit_s = [] { const char* e = getenv("SVP_CALL_TIMEOUT_S"); return e ? atof(e) : 900.0; }();
        const auto t0 = std::chrono::steady_clock::now();
        for (;;) {
            const cudaError_t q = cudaEventQuery(c.ev_t1);
            if (q == cudaSuccess) break;
            if (q != cudaErrorNotReady) return fail(h, SVP_ERR_CUDA, std::string("cudaEventQuery: ") + cudaGetErrorString(q));
            const double waited = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (waited > limit_s) {
                std::string msg = "call did not finish within " + std::to_string((int)limit_s) + " s";
                cudaStream_t dbg = nullptr;
                if (cudaStreamCreateWithFlags(&dbg, cudaStreamNonBlocking) == cudaSuccess) {
                    std::vector<ArenaPool> pools(h->n_sm + 1);
                    if (cudaMemcpyAsync(pools.data(), h->d_pools, pools.size() * sizeof(ArenaPool), cudaMemcpyDeviceToHost, dbg) == cudaSuccess && cudaStreamSynchronize(dbg) == cudaSuccess) {
                        uint64_t live = 0, bytes = 0, waits = 0, locked = 0;
                        for (const ArenaPool& p : pools) { live += p.n; waits += p.waits; locked += p.lock; for (uint32_t k = 0; k < p.n && k < (uint32_t)POOL_MAX; ++k) bytes += p.len[k]; }
                        msg += "; arena: " + std::to_string(live) + " live slices, " + std::to_string(bytes >> 20) + " MB, " + std::to_string(waits) + " waits, " +
                               std::to_string(locked) + " locks held, global pool " + std::to_string(pools[h->n_sm].n) + " slices of " + std::to_string(pools[h->n_sm].bytes >> 20) + " MB";
                    }
                    unsigned long long clk[4] = { 0, 0, 0, 0 };
                    if (cudaMemcpyAsync(clk, h->d_clocks, sizeof(clk), cudaMemcpyDeviceToHost, dbg) == cudaSuccess && cudaStreamSynchronize(dbg) == cudaSuccess)
                        msg += "; pairs walked " + std::to_string(clk[2]);
                }
                return fail(h, SVP_ERR_CUDA, msg);
            }
            if (waited > 0.002) std::this_thread::sleep_for(std::chrono::microseconds(waited > 0.05 ? 200 : 20));
        }
    }
    float tot = 0; cudaEventElapsedTime(&tot, c.ev_t0, c.ev_t1);
    c.stats.total_ms = tot;
    if (c.n_waves) {
        // wave scheduling: forward time = length of the union of the waves' forward intervals; walk time = sum of the walk kernels' spans
        auto EV = [&](size_t wv, int which) { return c.wave_events[wv * EV_PER_WAVE + which]; };
        float covered_to = 0;
        for (size_t wv = 0; wv < c.n_waves; ++wv) {
            float f0 = 0, f1 = 0, b = 0;
            cudaEventElapsedTime(&f0, c.ev_t0, EV(wv, 0));
            cudaEventElapsedTime(&f1, c.ev_t0, EV(wv, 1));
            cudaEventElapsedTime(&b, EV(wv, 1), EV(wv, 2));
            const float from = std::max(f0, covered_to);
            if (f1 > from) { c.stats.fwd_ms += f1 - from; covered_to = f1; }
            c.stats.tb_ms += b;
        }
        for (cudaEvent_t& hf : h->half_free)                     // this call's waves are done: the halves they used last are free
            for (cudaEvent_t e : c.wave_events) if (hf == e) { hf = nullptr; break; }
        h->stats = c.stats;
        svp_stats& T = h->stats_total;
        T.fwd_ms += c.stats.fwd_ms; T.tb_ms += c.stats.tb_ms; T.total_ms += c.stats.total_ms; T.cells += c.stats.cells; T.tb_bytes += c.stats.tb_bytes;
        T.fwd_launches += c.stats.fwd_launches; T.tb_launches += c.stats.tb_launches; T.waves += c.stats.waves;
        CUDA_TRY(h, cudaGetLastError());
        return SVP_OK;
    }
    // arena scheduling: forward pass and walk of a pair run inside one kernel; the kernel time of the call is split by the CTAs' own clocks
    // (summed ns of the two phases since the last collection; a pipelined neighbour call's CTAs are in the same sums)
    unsigned long long clk[6] = { 0, 0, 0, 0, 0, 0 };
    CUDA_TRY(h, cudaMemcpy(clk, h->d_clocks, sizeof(clk), cudaMemcpyDeviceToHost));
    const double f_ns = (double)(clk[0] - h->clocks_seen[0]), t_ns = (double)(clk[1] - h->clocks_seen[1]), w_ns = (double)(clk[4] - h->clocks_seen[4]);
    for (int k = 0; k < 6; ++k) h->clocks_seen[k] = clk[k];
    c.stats.fwd_ms = tot;
    // walks inside the CTAs: the call's time x the walks' share of the CTAs' summed phase times; walker kernel: the mean busy time of its warps
    c.stats.tb_ms = (f_ns + t_ns) > 0 ? tot * t_ns / (f_ns + t_ns) : 0.0;
    if (c.n_walkers) c.stats.tb_ms += w_ns / 1e6 / (double)c.n_walkers;
    h->stats = c.stats;
    svp_stats& T = h->stats_total;
    T.fwd_ms += c.stats.fwd_ms; T.tb_ms += c.stats.tb_ms; T.total_ms += c.stats.total_ms; T.cells += c.stats.cells; T.tb_bytes += c.stats.tb_bytes;
    T.fwd_launches += c.stats.fwd_launches; T.tb_launches += c.stats.tb_launches; T.waves += c.stats.waves;
    CUDA_TRY(h, cudaGetLastError());
    return SVP_OK;
}

extern "C" int svp_sync(svp_handle* h) {
    if (!h) return SVP_ERR_ARG;
    CUDA_TRY(h, cudaSetDevice(h->device));
    // older call first
    for (int k = 0; k < 2; ++k) { int rc = collect_call(h, h->sets[(h->call_idx + k) & 1]); if (rc != SVP_OK) return rc; }
    return SVP_OK;
}

extern "C" int svp_get_stats(const svp_handle* h, svp_stats* out) {
    if (!h || !out) return SVP_ERR_ARG;
    int rc = svp_sync(const_cast<svp_handle*>(h));
    if (rc != SVP_OK) return rc;
    *out = h->stats;
    return SVP_OK;
}

extern "C" int svp_get_stats_total(svp_handle* h, svp_stats* out, int reset) {
    if (!h) return SVP_ERR_ARG;
    int rc = svp_sync(h);
    if (rc != SVP_OK) return rc;
    if (out) *out = h->stats_total;
    if (reset) h->stats_total = svp_stats{};
    return SVP_OK;
}

extern "C" int svp_timer_mark(svp_handle* h, int which) {
    if (!h || which < 0 || which > 1) return SVP_ERR_ARG;
    CUDA_TRY(h, cudaSetDevice(h->device));
    if (!h->marks[which]) CUDA_TRY(h, cudaEventCreate(&h->marks[which]));
    // submitted calls end on the traceback stream: the mark comes after them
    for (CallSet& c : h->sets) if (c.pending) CUDA_TRY(h, cudaStreamWaitEvent(h->stream, c.ev_t1, 0));
    CUDA_TRY(h, cudaEventRecord(h->marks[which], h->stream));
    return SVP_OK;
}

extern "C" double svp_timer_elapsed_ms(svp_handle* h) {
    if (!h || !h->marks[0] || !h->marks[1]) return (double)SVP_ERR_ARG;
    if (cudaEventSynchronize(h->marks[1]) != cudaSuccess) return (double)SVP_ERR_CUDA;
    float ms = 0;
    if (cudaEventElapsedTime(&ms, h->marks[0], h->marks[1]) != cudaSuccess) return (double)SVP_ERR_CUDA;
    return (double)ms;
}

// #{(i, j) : 1<=i<=m, 1<=j<=n, j - i <= x}
static int64_t cells_below(int64_t m, int64_t n, int64_t x) {
    // row i contributes clamp(i + x, 0, n)
    int64_t total = 0;
    const int64_t i0 = std::max<int64_t>(1, 1 - x);            // first row with i + x >= 1
    const int64_t i1 = std::min<int64_t>(m, n - x - 1);        // last row with i + x <= n - 1
    if (i1 >= i0) total += (i1 - i0 + 1) * (i0 + i1) / 2 + x * (i1 - i0 + 1);
    const int64_t i2 = std::max<int64_t>(1, n - x);            // first row with i + x >= n
    if (i2 <= m) total += n * (m - i2 + 1);
    return total;
}

// one kernel = every pair of a call that runs on one variant (grid = pairs x CTAs per pair)
static cudaError_t launch_group(FwdKernel kern, int cluster, unsigned n_pairs, unsigned thr, size_t smem, cudaStream_t st, const Task* tasks,
                                const uint32_t* order, const uint8_t* seqs, const Arena& ar, svp_result* results, uint32_t* cigar, const Consts& cs) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(n_pairs * (unsigned)cluster, 1, 1);
    cfg.blockDim = dim3(thr, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = cluster > 1 ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, tasks, order, seqs, ar, results, cigar, cs);
}

// CIGAR compaction for the host-buffer entry point: one warp per pair copies its (right-aligned) slot
// content to the dense position computed on the host, so only the used words cross PCIe.
__global__ void svp_cigar_gather(const svp_result* __restrict__ res, const uint32_t* __restrict__ dense_off,
                                 const uint32_t* __restrict__ slots, uint32_t* __restrict__ dense, int n) {
    const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (k >= n) return;
    const uint32_t len = res[k].cigar_len;
    const uint32_t* src = slots + res[k].cigar_off;
    uint32_t* dst = dense + dense_off[k];
    for (uint32_t i = lane; i < len; i += 32) dst[i] = src[i];
}

struct Plan {
    std::vector<Task> tasks;                     // all pairs, input order
    std::vector<std::pair<size_t, size_t>> waves;   // [first, last) task ranges
};

static void plan_pair(const svp_handle* h, int call_kind, const svp_pair& p, uint32_t index, const uint64_t* seq_off, const uint32_t* seq_len, uint32_t n_seqs,
                      size_t packed_bytes, size_t cigar_words, bool need_padding, Task& t) {
    memset(&t, 0, sizeof(t));
    t.pair_index = index; t.cigar_off = p.cigar_off; t.cigar_cap = p.cigar_cap; t.flags = p.flags & SVP_PAIR_REVCOMP;
    bool ok = p.q < n_seqs && p.t < n_seqs && p.band_w > 0 && (p.band_w % 32u) == 0 && (size_t)p.cigar_off + p.cigar_cap <= cigar_words;
    if (ok) {
        for (uint32_t s : { p.q, p.t }) {
            const uint64_t off = seq_off[s] & ~(uint64_t)15;
            // codes, and — flagged in the low bits of seq_off — the mask of the letters outside ACGT behind them (SVP_SEQ_NMASK)
            const uint64_t bytes = (seq_off[s] & SVP_SEQ_NMASK) ? SVP_SEQ_CODE_BYTES(seq_len[s]) + ((uint64_t)seq_len[s] + 7) / 8 : ((uint64_t)seq_len[s] + 3) / 4;
            const uint64_t end = need_padding ? ((off + bytes + 15) & ~(uint64_t)15) : off + bytes;
            if ((seq_off[s] & 14u) || end > packed_bytes || seq_len[s] == 0 || seq_len[s] > (1u << 28)) ok = false;
        }
    }
    if (!ok) { t.flags |= TASK_INVALID; return; }
    t.q_off = seq_off[p.q] & ~(uint64_t)15; t.t_off = seq_off[p.t] & ~(uint64_t)15;
    if (seq_off[p.q] & SVP_SEQ_NMASK) t.flags |= TASK_QN;
    if (seq_off[p.t] & SVP_SEQ_NMASK) t.flags |= TASK_TN;
    t.m = (int32_t)seq_len[p.q]; t.n = (int32_t)seq_len[p.t];
    t.band_lo = p.band_lo; t.band_w = (int32_t)p.band_w;
    const Consts& cs = h->cs;
    const size_t max_smem = (size_t)h->max_smem_optin;
    const bool force_ring = h->force_ring;
    const long ring_min = h->ring_min;
    // kernel variant. General scoring and 32-bit lanes use the KP = 4 geometry. Sequences that do not fit whole shared-memory
    // images (with the widest CTA's mailboxes) run on a sliding-window variant.
    const size_t image_need = (size_t)(cs.general ? 2 : 1) * seq_image_bytes(t.m) + (size_t)seq_image_bytes(t.n) + (size_t)mailbox_bytes(16) + 64;
    // sliding windows: always when whole images do not fit; on the fast path also when whole images would cost occupancy
    // (more than SVP_RING_MIN bytes of images per warp of the CTA)
    const bool beyond_s16 = (int64_t)std::min(t.m, t.n) * cs.a_max > S16_SCORE_LIMIT;
    // scores beyond the s16 range: the rebased s16x2 fast path when the scoring allows it (simple scoring, bounded neighbour
    // differences); else 32-bit lanes
    const bool rebase = beyond_s16 && !cs.general && cs.rebase_ok && (int64_t)std::min(t.m, t.n) * cs.a_max < (1 << 30);
    const bool wide = beyond_s16 && !rebase;
    bool ring = force_ring || image_need > max_smem;
    const bool wide32 = wide || (ring && cs.general);        // general scoring has no s16 sliding-window kernel: 32-bit lanes
    // bands beyond one CTA (16 warps) run on a thread-block cluster of 2, 4 or 8 CTAs
    int kp, cl = 1;
    if (wide32 || cs.general) { kp = 4; cl = cluster_for(p.band_w, 128 * 4); }
    else {
        // a small call (one or two containers: fewer warps than the GPU has schedulers) runs on KP = 4 / 8 only: a container of 435 pairs
        // x 8 kb takes 3.6 ms on them against 5.6 ms on its four single-warp widths (measured; a lone KP = 5 / 6 / 7 warp per scheduler
        // is latency-bound, and four half-empty launches replace two)
        const int kp_mask = (call_kind & 2) && (h->kp_mask & 0x11) ? (h->kp_mask & 0x11) : h->kp_mask;
        kp = choose_kp(p.band_w, (call_kind & 1) != 0, cs.two_piece != 0, !beyond_s16, h->waves, kp_mask, h->kp_multi);
        if (kp == 0) { kp = 8; cl = cluster_for(p.band_w, 128 * 8); }
        if (!ring && cl) {
            const int nw = cl > 1 ? 16 : (int)((p.band_w + 128 * kp - 1) / (128 * kp));
            ring = ((long)seq_image_bytes(t.m) + (long)seq_image_bytes(t.n)) / nw > (nw == 1 ? h->ring_min_single : ring_min);
        }
    }
    if (cl == 0) { t.flags |= TASK_INVALID; return; }                         // band too wide for every variant
    if (cl > 1) t.flags |= (uint32_t)cl << TASK_CL_SHIFT;
    if (wide32) t.flags |= TASK_S32;
    if (rebase && !wide32) t.flags |= TASK_REBASE;
    if (ring) t.flags |= TASK_RING;
    t.flags |= (uint32_t)kp << TASK_KP_SHIFT;
    const int steps_per_iter = steps_per_iter_of(t);
    const int64_t m = t.m, n = t.n, lo = p.band_lo, hi = lo + (int64_t)p.band_w - 1;
    if (lo > n - 1 || hi < -(m - 1)) { t.flags |= TASK_INVALID; return; }     // band misses the matrix
    t.cells = (uint64_t)(cells_below(m, n, hi) - cells_below(m, n, lo - 1));
    // first / last anti-diagonal holding a real in-band cell
    int64_t rmin;
    if (lo <= 0 && hi >= 0) rmin = 2; else rmin = 2 + std::min(std::llabs(lo), std::llabs(hi));
    const int64_t dpeak = std::min(std::max<int64_t>(n - m, lo), hi);        // diagonal whose last cell is furthest
    const int64_t rmax = (dpeak <= n - m) ? 2 * m + dpeak : 2 * n - dpeak;
    int64_t rb = rmin;
    if (((rb - lo) & 1) != 0) rb -= 1;                                       // step 0 must be an A-step
    const int64_t steps = rmax - rb + 1;
    t.r_begin = (int32_t)rb;
    t.n_iter = (int32_t)((steps + steps_per_iter - 1) / steps_per_iter);
    t.tb_bytes = (tb_bytes_of(t) + 15) & ~(uint64_t)15;
    t.slice_bytes = (t.tb_bytes + best_entries_of(t) * sizeof(uint2) + 255) & ~(uint64_t)255;
}

// pinned staging for the next call's task / order arrays; waits for the call that used this set two calls ago
static int begin_call(svp_handle* h, size_t n, CallSet** out) {
    CallSet& c = h->sets[h->call_idx & 1];
    int rc = collect_call(h, c);
    if (rc != SVP_OK) return rc;
    if (n > c.h_tasks_cap) {
        if (c.h_tasks) cudaFreeHost(c.h_tasks);
        c.h_tasks = nullptr; c.h_tasks_cap = 0;
        const size_t cap = n + n / 4 + 16;
        CUDA_TRY(h, cudaHostAlloc((void**)&c.h_tasks, cap * sizeof(Task), cudaHostAllocDefault));
        c.h_tasks_cap = cap;
    }
    if (2 * n > c.h_order_cap) {                 // forward launch order + the traceback lists of every wave
        if (c.h_order) cudaFreeHost(c.h_order);
        c.h_order = nullptr; c.h_order_cap = 0;
        const size_t cap = 2 * n + n / 2 + 16;
        CUDA_TRY(h, cudaHostAlloc((void**)&c.h_order, cap * sizeof(uint32_t), cudaHostAllocDefault));
        c.h_order_cap = cap;
    }
    if (!c.ev_t0) { CUDA_TRY(h, cudaEventCreate(&c.ev_t0)); CUDA_TRY(h, cudaEventCreate(&c.ev_t1)); }
    *out = &c;
    return SVP_OK;
}

// ---- launch groups: the pairs of a call (or of one wave of it) that run on one kernel variant ------------------------------------
struct Group { int fam, nwarp, kp, cluster; bool ring, huge, inline_walk; size_t first, count, smem; uint64_t work, max_slice; };

static inline int warps_of(const Task& t) { const int cl = cluster_of(t); return cl > 1 ? 16 : (t.band_w + 128 * kp_of(t) - 1) / (128 * kp_of(t)); }
static inline int family_of(const Task& t, const Consts& cs) {
    return (t.flags & TASK_S32) ? (cs.general ? FAM_S32_GEN : FAM_S32) : (cs.general ? FAM_GEN : (t.flags & TASK_REBASE) ? FAM_FAST_RB : FAM_FAST);
}
// Groups of the valid tasks among `idx` (task indices), appended to `groups`; their order lists go to order[pos...]. A group = one
// kernel variant x CTA width x shared-memory class (<= 32 KB, then powers of two, so that a few long pairs do not cut the occupancy
// of the short ones sharing their launch; sliding-window variants have a fixed size) [x slice-size class, arena scheduling].
// Inside a group the longest pairs come first (the tail of the launch is short pairs); the groups with the longest-running CTAs
// come first (+6 % over biggest-group-first: the many short single-warp CTAs fill the tail).
static void make_groups(const svp_handle* h, const Task* tasks, const uint32_t* idx, size_t n_idx, bool slice_classes, uint32_t* order, size_t& pos,
                        std::vector<Group>& groups) {
    const Consts& cs = h->cs;
    // one 64-bit sort key per pair: [group key: 31 bits][0xFFFFFFFF - n_iter: 32 bits] + the task index beside it; a single sort puts the
    // pairs of a group next to each other, longest first (a map of vectors and a sort per group cost ~4 ms per 35 000 pairs — GPU idle
    // time of a blocking call)
    struct Item { uint64_t key; uint32_t task; };
    std::vector<Item> items;
    items.reserve(n_idx);
    for (size_t a = 0; a < n_idx; ++a) {
        const uint32_t k = idx ? idx[a] : (uint32_t)a;
        const Task& t = tasks[k];
        if (t.flags & TASK_INVALID) continue;
        const int nw = warps_of(t);
        const bool ring = (t.flags & TASK_RING) != 0;
        int smem_cls = 0;
        if (!ring) while (((size_t)32768 << smem_cls) < smem_need(t, cs, nw)) ++smem_cls;
        const bool huge = slice_classes && t.slice_bytes * 3 > h->sm_region;
        // slice-size classes (x1.5 steps) keep the ballast of a regular group close to what each of its pairs needs
        int size_cls = 0;
        if (slice_classes && !huge) for (uint64_t b = (uint64_t)1 << 20; b < t.slice_bytes; b += b / 2) ++size_cls;
        const uint64_t gkey = ((uint64_t)family_of(t, cs) << 28) | ((uint64_t)kp_of(t) << 23) | ((uint64_t)nw << 18) | ((uint64_t)cluster_of(t) << 14) |
                              ((uint64_t)ring << 13) | ((uint64_t)smem_cls << 8) | ((uint64_t)huge << 7) | (uint64_t)std::min(size_cls, 127);
        items.push_back({ (gkey << 32) | (uint64_t)(0xFFFFFFFFu - (uint32_t)t.n_iter), k });
    }
    std::sort(items.begin(), items.end(), [](const Item& a, const Item& b) { return a.key != b.key ? a.key < b.key : a.task < b.task; });
    const size_t g0 = groups.size();
    for (size_t a = 0; a < items.size();) {
        const uint64_t gkey = items[a].key >> 32;
        size_t e = a;
        while (e < items.size() && (items[e].key >> 32) == gkey) ++e;
        Group g{ (int)((gkey >> 28) & 7), (int)((gkey >> 18) & 31), (int)((gkey >> 23) & 31), (int)((gkey >> 14) & 15), ((gkey >> 13) & 1) != 0, ((gkey >> 7) & 1) != 0,
                 false, pos, e - a, 0, 0, 0 };
        const int lane_cost = (g.fam == FAM_S32 || g.fam == FAM_S32_GEN) ? 2 : 1;
        for (size_t x = a; x < e; ++x) {
            const Task& t = tasks[items[x].task];
            g.smem = std::max(g.smem, smem_need(t, cs, g.nwarp));
            g.max_slice = std::max(g.max_slice, t.slice_bytes);
            g.work += (uint64_t)t.n_iter * g.nwarp * g.kp * lane_cost * (steps_per_iter_of(t) / (2 * g.kp)) * g.cluster;
            order[pos++] = items[x].task;
        }
        groups.push_back(g);
        a = e;
    }
    std::sort(groups.begin() + g0, groups.end(), [](const Group& a, const Group& b) { return a.work * b.count > b.work * a.count; });
}

static const bool g_corridor = [] { const char* e = getenv("SVP_CORRIDOR"); return e ? atoi(e) != 0 : true; }();  // A/B aid: 0 = the walks load every byte from global memory
static const bool g_skip_walk = [] { const char* e = getenv("SVP_SKIP_WALK"); return e && atoi(e) != 0; }();      // measurement aid: forward passes only

// ---------------------------------------------------------------------------------------------------------------------------------
// Wave scheduling. A call is cut into waves of consecutive pairs whose slices fit HALF the arena; the halves alternate from wave to
// wave (across calls), the host assigns every pair its slice (Task.tb_off), and the walks of a wave run as kernels of their own
// (svp_tb_kernel / svp_tbw_kernel) on a high-priority stream under the forward pass of the next wave.
// ---------------------------------------------------------------------------------------------------------------------------------
static int run_waves(svp_handle* h, CallSet& c, size_t n, const uint8_t* d_seqs, svp_result* d_results, uint32_t* d_cigar, bool wait) {
    const Consts& cs = h->cs;
    cudaStream_t st = h->stream;
    Task* tasks = c.h_tasks;
    const uint64_t half = (h->arena_bytes / 2) & ~(uint64_t)255;
    uint64_t cap = half;                        // bytes a wave may use
    bool serial = false;
    std::vector<std::pair<size_t, size_t>> waves;
    {
        uint64_t total_need = 0, total_len = 0, max_need = 0;
        for (size_t k = 0; k < n; ++k)
            if (!(tasks[k].flags & TASK_INVALID)) { total_need += tasks[k].slice_bytes; total_len += (uint64_t)tasks[k].m + tasks[k].n; max_need = std::max(max_need, tasks[k].slice_bytes); }
        // Two ways to run the walks. CONCURRENT: waves of half the arena, the walks of a wave run under the forward pass of the next
        // one. SERIAL: waves of the whole arena, the walks of a wave run alone on the warp-per-pair kernel (5-6 ms for 10 000 pairs)
        // before the next forward pass starts. Concurrent walks slow the forward pass they share the SMs with by 12 % (config 2) to
        // 18-50 % (config 4: many short pairs), more than they take alone, and whole-arena waves halve the number of wave
        // boundaries; but waves of few, long pairs (config 3) have latency-bound walks that need the overlap, and on config 2 (8 kb
        // reads) the two measure the same. Hence: serial when a half-arena wave would hold at least serial_min_pairs pairs of, on
        // average, less than serial_max_len bases (m + n). Env SVP_TB_SERIAL = 0 / 1 forces the choice.
        static const int serial_mode = [] { const char* e = getenv("SVP_TB_SERIAL"); return e ? atoi(e) : -1; }();
        static const long serial_min_pairs = [] { const char* e = getenv("SVP_TB_SERIAL_PAIRS"); return e ? atol(e) : 4096L; }();
        static const long serial_max_len = [] { const char* e = getenv("SVP_TB_SERIAL_LEN"); return e ? atol(e) : 20000L; }();
        const uint64_t n_w_half = std::max<uint64_t>(1, (total_need + half - 1) / half);
        serial = serial_mode >= 0 ? serial_mode != 0 : ((long)(n / n_w_half) >= serial_min_pairs && (long)(total_len / std::max<size_t>(n, 1)) < serial_max_len);
        if (max_need > half) serial = true;      // a pair beyond half the arena: whole-arena waves, whatever the preference
        if (serial) cap = h->arena_bytes & ~(uint64_t)255;
        // a call that needs several waves gets waves of equal size (not full halves and a remainder: the short wave's forward pass
        // would wait for the long wave's walks)
        const uint64_t n_w = std::max<uint64_t>(1, (total_need + cap - 1) / cap);
        const uint64_t target = std::min(cap, total_need / n_w + total_need / (n_w * 50) + 4096);
        size_t first = 0; uint64_t used = 0;
        for (size_t k = 0; k < n; ++k) {
            Task& t = tasks[k];
            uint64_t need = 0;
            if (!(t.flags & TASK_INVALID)) {
                need = t.slice_bytes;
                if (need > cap) return fail(h, SVP_ERR_NOMEM, "one pair's traceback exceeds the arena; create the handle with a larger max_tb_bytes");
            }
            if (used + need > cap || (used >= target && need)) { waves.emplace_back(first, k); first = k; used = 0; }
            t.tb_off = (serial ? 0 : ((h->wave_idx + waves.size()) & 1) * half) + used;
            used += need;
        }
        waves.emplace_back(first, n);
    }
    // per wave: launch groups, then the walk lists — a wave that is not walked by the warp-per-pair kernel as a whole hands its
    // long pairs (m + n beyond tbw_long: a 40-90 kb consensus against its reference window takes the thread-per-pair kernel 10 000
    // dependent rounds) to warps, the rest to the bulk kernel
    uint32_t* order = c.h_order;
    std::vector<std::vector<Group>> wave_groups(waves.size());
    struct TbList { size_t first_short, n_short, first_long, n_long; };
    std::vector<TbList> tb_lists(waves.size());
    size_t pos = 0;
    std::vector<uint32_t> idx;
    for (size_t wv = 0; wv < waves.size(); ++wv) {
        idx.resize(waves[wv].second - waves[wv].first);
        for (size_t a = 0; a < idx.size(); ++a) idx[a] = (uint32_t)(waves[wv].first + a);
        make_groups(h, tasks, idx.data(), idx.size(), false, order, pos, wave_groups[wv]);
    }
    size_t n_invalid_total = 0;
    for (size_t wv = 0; wv < waves.size(); ++wv) {
        TbList& L = tb_lists[wv];
        L.first_short = pos;
        for (size_t k = waves[wv].first; k < waves[wv].second; ++k) {
            if (tasks[k].flags & TASK_INVALID) { ++n_invalid_total; continue; }
            if ((long)tasks[k].m + tasks[k].n <= h->tbw_long) order[pos++] = (uint32_t)k;
        }
        L.n_short = pos - L.first_short;
        // the groups of a warp walk in lockstep until the longest of them is done (svp_traceback_groups): neighbours in the list
        // should be equally long; longest first, so that the kernel's tail is short walks
        {
            std::vector<uint64_t> keyed(pos - L.first_short);
            for (size_t a = 0; a < keyed.size(); ++a) {
                const uint32_t k = order[L.first_short + a];
                keyed[a] = ((uint64_t)(0xFFFFFFFFu - (uint32_t)(tasks[k].m + tasks[k].n)) << 32) | k;
            }
            std::sort(keyed.begin(), keyed.end());
            for (size_t a = 0; a < keyed.size(); ++a) order[L.first_short + a] = (uint32_t)keyed[a];
        }
        L.first_long = pos;
        for (size_t k = waves[wv].first; k < waves[wv].second; ++k)
            if (!(tasks[k].flags & TASK_INVALID) && (long)tasks[k].m + tasks[k].n > h->tbw_long) order[pos++] = (uint32_t)k;
        L.n_long = pos - L.first_long;
    }
    const size_t inv_first = pos;
    if (n_invalid_total) for (size_t k = 0; k < n; ++k) if (tasks[k].flags & TASK_INVALID) order[pos++] = (uint32_t)k;

    CUDA_TRY(h, c.d_tasks.reserve(n * sizeof(Task)));
    CUDA_TRY(h, c.d_order.reserve(std::max<size_t>(pos, 1) * sizeof(uint32_t)));
    CUDA_TRY(h, cudaEventRecord(c.ev_t0, st));
    CUDA_TRY(h, cudaMemcpyAsync(c.d_tasks.p, tasks, n * sizeof(Task), cudaMemcpyHostToDevice, st));
    if (pos) CUDA_TRY(h, cudaMemcpyAsync(c.d_order.p, order, pos * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    const Task* d_tasks = (const Task*)c.d_tasks.p;
    const uint32_t* d_order = (const uint32_t*)c.d_order.p;
    if (n_invalid_total)
        svp_invalid_kernel<<<(unsigned)((n_invalid_total + 127) / 128), 128, 0, st>>>(d_tasks, d_order + inv_first, (int)n_invalid_total, d_results);

    const size_t n_ev = waves.size() * EV_PER_WAVE;
    while (c.wave_events.size() < n_ev) { cudaEvent_t e; CUDA_TRY(h, cudaEventCreate(&e)); c.wave_events.push_back(e); }
    auto EV = [&](size_t wv, int which) { return c.wave_events[wv * EV_PER_WAVE + which]; };
    // Forward passes of consecutive waves stay in stream order. Letting them overlap was measured slower four times — on separate
    // stream sets (1810 vs 1915 GCUPS) and on priority-tiered stream sets (1810-1909 vs 1923-1962): the union of the forward
    // intervals gets longer, not shorter, and the walks are delayed. Inside a wave the groups run on up to four streams (2..6
    // measure the same, one stream 10 % slower).
    cudaStream_t sF[N_FWD_STREAMS] = { st };
    for (int a = 1; a < N_FWD_STREAMS; ++a) sF[a] = h->stream_aux[a - 1];
    static const int max_streams = [] { const char* e = getenv("SVP_FWD_STREAMS"); int v = e ? atoi(e) : 4; return std::min(std::max(v, 1), N_FWD_STREAMS); }();
    cudaStream_t sT = h->stream_tb;
    // Waves of few, long-running CTAs (long reads, wide bands: config 3) end in a tail of a few CTAs on an otherwise idle machine;
    // there the forward pass of the next wave starts at once on the other half of the streams (it writes the other arena half
    // anyway) instead of waiting in stream order. SVP_WAVE_OVERLAP = 0 / 1 forces the choice, default: waves averaging fewer than
    // overlap_max_pairs pairs.
    static const int overlap_mode = [] { const char* e = getenv("SVP_WAVE_OVERLAP"); return e ? atoi(e) : -1; }();
    static const long overlap_max_pairs = [] { const char* e = getenv("SVP_WAVE_OVERLAP_PAIRS"); return e ? atol(e) : 4096L; }();
    const bool overlap = !serial && waves.size() > 1 && (overlap_mode >= 0 ? overlap_mode != 0 : (long)(n / waves.size()) < overlap_max_pairs);
    while (c.events.size() < 1) { cudaEvent_t e; CUDA_TRY(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); c.events.push_back(e); }
    cudaEvent_t ev_copied = c.events[0];                 // the copies above precede every wave, whatever stream leads it
    CUDA_TRY(h, cudaEventRecord(ev_copied, st));
    Arena ar{ h->arena, h->d_pools, h->n_sm, 0u, h->d_clocks, 1u, 0u, nullptr, nullptr, 0u, 0u, 0u };
    for (size_t wv = 0; wv < waves.size(); ++wv) {
        const int half_id = (int)((h->wave_idx + wv) & 1);
        cudaStream_t* sW = overlap ? sF + half_id * (N_FWD_STREAMS / 2) : sF;         // this wave's streams
        const int avail = overlap ? N_FWD_STREAMS / 2 : max_streams;
        const std::vector<Group>& groups = wave_groups[wv];
        const int n_streams = (int)std::max<size_t>(1, std::min<size_t>((size_t)avail, groups.size()));
        // join = true: the wave starts on all its streams at once, after the previous wave's forward pass has ended on all of them.
        // join = false (env SVP_WAVE_JOIN=0): every stream goes on with this wave's groups as soon as its own groups of the previous
        // wave are done and the arena half is free — no drain and ramp between waves
        static const bool join = [] { const char* e = getenv("SVP_WAVE_JOIN"); return e ? atoi(e) != 0 : true; }();
        for (int a = 0; a < (join ? 1 : n_streams); ++a) {
            if (sW[a] != st) CUDA_TRY(h, cudaStreamWaitEvent(sW[a], ev_copied, 0));
            if (h->half_free[half_id]) CUDA_TRY(h, cudaStreamWaitEvent(sW[a], h->half_free[half_id], 0));      // this arena half is free again
            if (serial && h->half_free[half_id ^ 1]) CUDA_TRY(h, cudaStreamWaitEvent(sW[a], h->half_free[half_id ^ 1], 0));   // whole-arena wave
        }
        CUDA_TRY(h, cudaEventRecord(EV(wv, 0), sW[0]));
        if (join) for (int a = 1; a < n_streams; ++a) CUDA_TRY(h, cudaStreamWaitEvent(sW[a], EV(wv, 0), 0));
        for (size_t gi = 0; gi < groups.size(); ++gi) {
            const Group& g = groups[gi];
            FwdKernel kern = fwd_kernel(g.fam, g.kp, g.nwarp > 1, g.cluster > 1, g.ring, cs.two_piece != 0, false);
            if (!kern) return fail(h, SVP_ERR_UNSUPPORTED, "internal: no kernel variant for a planned group");
            CUDA_TRY(h, launch_group(kern, g.cluster, (unsigned)g.count, (unsigned)g.nwarp * 32u, g.smem, sW[gi % (size_t)n_streams], d_tasks, d_order + g.first,
                                     d_seqs, ar, d_results, d_cigar, cs));
            c.stats.fwd_launches++;
        }
        if (join) {
            for (int a = 1; a < n_streams; ++a) {
                CUDA_TRY(h, cudaEventRecord(EV(wv, 3 + a), sW[a]));
                CUDA_TRY(h, cudaStreamWaitEvent(sW[0], EV(wv, 3 + a), 0));
            }
            CUDA_TRY(h, cudaEventRecord(EV(wv, 1), sW[0]));
            CUDA_TRY(h, cudaStreamWaitEvent(sT, EV(wv, 1), 0));
        } else {                                 // only the walks wait for the whole wave
            for (int a = 1; a < n_streams; ++a) {
                CUDA_TRY(h, cudaEventRecord(EV(wv, 3 + a), sW[a]));
                CUDA_TRY(h, cudaStreamWaitEvent(sT, EV(wv, 3 + a), 0));
            }
            CUDA_TRY(h, cudaEventRecord(EV(wv, 1), sW[0]));
            CUDA_TRY(h, cudaStreamWaitEvent(sT, EV(wv, 1), 0));
            if (h->stream_tb2) for (int a = 1; a < n_streams; ++a) CUDA_TRY(h, cudaStreamWaitEvent(h->stream_tb2, EV(wv, 3 + a), 0));
        }
        const TbList& L = tb_lists[wv];
        const int nt = (int)(L.n_short + L.n_long);
        if (nt > 0 && !g_skip_walk) {
            // few pairs in the wave (one container per call; long tandem-repeat pairs): the warp-per-pair kernel walks 32 cells per
            // memory round trip; in bulk the thread-per-pair kernel costs fewer issue slots under the next wave's forward pass.
            // The last wave of a blocking call has no forward pass left to share the machine with: its walks are pure latency.
            static const bool last_wave_warp = [] { const char* e = getenv("SVP_TBW_LAST"); return e ? atoi(e) != 0 : true; }();
            if (nt <= h->tbw_max || serial || (last_wave_warp && wait && wv + 1 == waves.size())) {
                // (the short and the long list of a wave are adjacent in the order array). Lanes per pair (env SVP_TB_LPP): 8 — four
                // pairs share a warp's instruction stream; 32 — the whole warp, with its corridor in shared memory
                static const int lpp = [] { const char* e = getenv("SVP_TB_LPP"); const int v = e ? atoi(e) : 8; return v == 4 || v == 16 || v == 32 ? v : 8; }();
                static const int tbg_min = [] { const char* e = getenv("SVP_TBG_MIN"); return e ? atoi(e) : -1; }();
                const bool whole = lpp == 32 || nt < (tbg_min >= 0 ? tbg_min : 4 * h->sm_count);         // few pairs: latency counts, not issue slots
                if (whole)
                    svp_tbw_kernel<<<(nt * 32 + 127) / 128, 128, g_corridor ? 4 * COR_BYTES : 0, sT>>>(d_tasks, d_order + L.first_short, nt, d_seqs, h->arena, d_results, d_cigar, cs, g_corridor ? 1 : 0);
                else {
                    if (L.n_long) svp_tbw_kernel<<<((int)L.n_long * 32 + 127) / 128, 128, 0, sT>>>(d_tasks, d_order + L.first_long, (int)L.n_long, d_seqs, h->arena, d_results, d_cigar, cs, 0);
                    // SVP_TBG_UNIFORM=0: the groups of a warp walk independently (group masks, divergent) instead of on one instruction stream
                    static const int uni = [] { const char* e = getenv("SVP_TBG_UNIFORM"); return e ? atoi(e) : 1; }();
#define SVP_TBG_ARGS d_tasks, d_order + L.first_short, (int)L.n_short, d_seqs, h->arena, d_results, d_cigar, cs
                    const unsigned grid = (unsigned)(((size_t)L.n_short * (size_t)lpp + 127) / 128);
                    if (L.n_short && lpp == 16) { if (uni) svp_tbu_kernel<16><<<grid, 128, 0, sT>>>(SVP_TBG_ARGS); else svp_tbg_kernel<16><<<grid, 128, 0, sT>>>(SVP_TBG_ARGS); }
                    if (L.n_short && lpp == 8) { if (uni) svp_tbu_kernel<8><<<grid, 128, 0, sT>>>(SVP_TBG_ARGS); else svp_tbg_kernel<8><<<grid, 128, 0, sT>>>(SVP_TBG_ARGS); }
                    if (L.n_short && lpp == 4) svp_tbu_kernel<4><<<grid, 128, 0, sT>>>(SVP_TBG_ARGS);
#undef SVP_TBG_ARGS
                }
                c.stats.tb_launches++;
            } else {
                if (L.n_long) {              // the long walks first, on their own stream beside the bulk kernel
                    CUDA_TRY(h, cudaStreamWaitEvent(h->stream_tb2, EV(wv, 1), 0));
                    svp_tbw_kernel<<<((int)L.n_long * 32 + 127) / 128, 128, g_corridor ? 4 * COR_BYTES : 0, h->stream_tb2>>>(d_tasks, d_order + L.first_long, (int)L.n_long, d_seqs,
                                                                                                                       h->arena, d_results, d_cigar, cs, g_corridor ? 1 : 0);
                    CUDA_TRY(h, cudaEventRecord(EV(wv, 3), h->stream_tb2));
                    c.stats.tb_launches++;
                }
                if (L.n_short) {
                    svp_tb_kernel<<<((int)L.n_short + 63) / 64, 64, 0, sT>>>(d_tasks, d_order + L.first_short, (int)L.n_short, d_seqs, h->arena, d_results, d_cigar, cs);
                    c.stats.tb_launches++;
                }
                if (L.n_long) CUDA_TRY(h, cudaStreamWaitEvent(sT, EV(wv, 3), 0));
            }
        }
        CUDA_TRY(h, cudaEventRecord(EV(wv, 2), sT));
        h->half_free[half_id] = EV(wv, 2);
        if (serial) h->half_free[half_id ^ 1] = EV(wv, 2);
        c.stats.waves++;
    }
    h->wave_idx += waves.size();
    CUDA_TRY(h, cudaStreamWaitEvent(sT, ev_copied, 0));  // (a call of invalid pairs only: their rows are written on the main stream)
    CUDA_TRY(h, cudaEventRecord(c.ev_t1, sT));           // walks run in wave order on one stream: the last one ends the call
    c.n_waves = waves.size();
    c.wave_ranges = waves;
    return SVP_OK;
}

// ---------------------------------------------------------------------------------------------------------------------------------
// Arena scheduling. There are no waves. Every pair's CTA takes its traceback rows from the device-side arena when it starts, walks
// its pair itself after the forward pass and returns the slice (svp_kernels.cuh: ArenaPool), so the arena bounds the pairs in
// flight, not the batch. What the host does:
//   * classify the pairs: "regular" slices (at most a third of a per-SM region) come from the pool of the SM the CTA runs on,
//     larger ones from the global pool; the split of the arena between the two follows the demand of the call;
//   * group the pairs by kernel variant, CTA width, shared-memory class and slice-size class; one launch per group;
//   * bound the residency: a regular group asks for at least 1.5 x slice / region of an SM's shared memory per CTA (ballast), so
//     the block scheduler itself keeps the slices live on an SM within its region, with room for fragmentation — a CTA that still
//     finds its pool full waits for a neighbour to finish (pair_begin);
//   * launch the groups with the longest-running CTAs first, round-robin over the streams; consecutive calls rotate the streams
//     and overlap freely.
// ---------------------------------------------------------------------------------------------------------------------------------
static int run_arena(svp_handle* h, CallSet& c, size_t n, const uint8_t* d_seqs, svp_result* d_results, uint32_t* d_cigar) {
    const Consts& cs = h->cs;
    cudaStream_t st = h->stream;
    Task* tasks = c.h_tasks;
    // ---- arena partition: which share goes to the global pool ----------------------------------------------------------
    {
        uint64_t max_slice = 0;
        for (size_t k = 0; k < n; ++k) if (!(tasks[k].flags & TASK_INVALID)) max_slice = std::max(max_slice, tasks[k].slice_bytes);
        auto demand = [&](double gf, double& share, bool& fits) {
            const uint64_t total = h->arena_bytes & ~(uint64_t)255;
            const uint64_t glob = (uint64_t)((double)total * gf), region = (total - glob) / h->n_sm;
            double huge = 0, all = 0;
            for (size_t k = 0; k < n; ++k) {
                if (tasks[k].flags & TASK_INVALID) continue;
                all += (double)tasks[k].slice_bytes;
                if (tasks[k].slice_bytes * 3 > region) huge += (double)tasks[k].slice_bytes;
            }
            share = all > 0 ? huge / all : 0.0;
            fits = max_slice * 3 <= region || max_slice + 4096 <= glob;
        };
        double share = 0; bool fits = false;
        demand(h->global_frac, share, fits);
        if (!fits || std::fabs(std::min(std::max(share, 0.1), 0.9) - h->global_frac) > 0.3) {
            double best_gf = -1, best_d = 1e9;
            for (double gf : { 0.1, 0.25, 0.5, 0.75, 0.9 }) {
                double sh; bool ok;
                demand(gf, sh, ok);
                const double d = std::fabs(std::min(std::max(sh, 0.1), 0.9) - gf);
                if (ok && d < best_d) { best_d = d; best_gf = gf; }
            }
            if (best_gf < 0) return fail(h, SVP_ERR_NOMEM, "one pair's traceback exceeds the arena; create the handle with a larger max_tb_bytes");
            if (best_gf != h->global_frac) {
                for (CallSet& o : h->sets) { int rc = collect_call(h, o); if (rc != SVP_OK) return rc; }      // the device must be idle
                CUDA_TRY(h, cudaDeviceSynchronize());
                int rc = set_partition(h, best_gf);
                if (rc != SVP_OK) return rc;
            }
        }
    }
    std::vector<Group> groups;
    uint32_t* order = c.h_order;
    size_t pos = 0;
    make_groups(h, tasks, nullptr, n, true, order, pos, groups);
    for (Group& g : groups)
        if (!g.huge) {
            // ballast: CTAs per SM <= smem_per_sm / (smem + 1 KB the system reserves per CTA); keep 1.5 x their slices within the SM's region
            static const double slack = [] { const char* e = getenv("SVP_BALLAST_SLACK"); return e ? atof(e) : 1.5; }();
            const double per_cta = slack * (double)g.max_slice / (double)h->sm_region * (double)h->smem_per_sm;
            const size_t ballast = per_cta > 1024.0 ? (size_t)per_cta - 1024 + 16 : 0;
            g.smem = std::min<size_t>(std::max(g.smem, ballast), (size_t)h->max_smem_optin - SMEM_STATIC);
        }
    if (g_corridor) for (Group& g : groups) g.smem = std::max<size_t>(g.smem, COR_BYTES);       // the walking warp's corridor (svp_kernels.cuh: TbView)
    const size_t inv_first = pos;
    for (size_t k = 0; k < n; ++k) if (tasks[k].flags & TASK_INVALID) order[pos++] = (uint32_t)k;
    const size_t n_invalid = pos - inv_first;
    // ---- queue scheduling: which pairs are queued for the walker kernel ---------------------------------------------------------
    // Not the pairs of a small call (a warp per walk inside the CTAs has a third of the latency), not long pairs (Arena.walk_long),
    // and not the groups whose CTA needs (nearly) a whole SM — 14+ warps at up to 128 registers, or sequence images filling the
    // shared memory: such a CTA cannot become resident beside a walker CTA, and the walkers stay until every queued pair is
    // walked, so it must not depend on them.
    constexpr int WALKER_REGS = 32 * 128;
    size_t n_queued = 0;
    const bool use_queue = h->queue && (long)(n - n_invalid) >= h->queue_min_pairs;
    for (Group& g : groups) {
        g.inline_walk = !use_queue || g.nwarp * 32 * 128 + h->walk_ctas_per_sm * WALKER_REGS > 65536 ||
                        g.smem + 1024 + (size_t)h->walk_ctas_per_sm * 1024 > (size_t)h->smem_per_sm;
        if (!g.inline_walk)
            for (size_t a = 0; a < g.count; ++a) { const Task& t = tasks[order[g.first + a]]; if ((long)t.m + t.n <= h->tbw_long) ++n_queued; }
    }

    CUDA_TRY(h, c.d_tasks.reserve(n * sizeof(Task)));
    CUDA_TRY(h, c.d_order.reserve(std::max<size_t>(pos, 1) * sizeof(uint32_t)));
    CUDA_TRY(h, cudaEventRecord(c.ev_t0, st));
    CUDA_TRY(h, cudaMemcpyAsync(c.d_tasks.p, tasks, n * sizeof(Task), cudaMemcpyHostToDevice, st));
    if (pos) CUDA_TRY(h, cudaMemcpyAsync(c.d_order.p, order, pos * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    const Task* d_tasks = (const Task*)c.d_tasks.p;
    const uint32_t* d_order = (const uint32_t*)c.d_order.p;
    WalkQueue* d_queue = nullptr;
    if (n_queued) {                                     // one entry per queued pair, unpublished entries are all-ones
        const size_t qbytes = sizeof(WalkQueue) + n_queued * sizeof(WalkEntry);
        CUDA_TRY(h, c.d_queue.reserve(qbytes));
        d_queue = (WalkQueue*)c.d_queue.p;
        CUDA_TRY(h, cudaMemsetAsync(d_queue, 0xFF, qbytes, st));
        svp_queue_init_kernel<<<1, 1, 0, st>>>(d_queue, (uint32_t)n_queued);
    }

    static const int max_streams = [] { const char* e = getenv("SVP_FWD_STREAMS"); int v = e ? atoi(e) : N_FWD_STREAMS; return std::min(std::max(v, 1), N_FWD_STREAMS); }();
    cudaStream_t sF[N_FWD_STREAMS] = { st };
    for (int a = 1; a < N_FWD_STREAMS; ++a) sF[a] = h->stream_aux[a - 1];
    while (c.events.size() < (size_t)N_FWD_STREAMS + 2) { cudaEvent_t e; CUDA_TRY(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); c.events.push_back(e); }
    const int n_streams = (int)std::min<size_t>((size_t)max_streams, std::max<size_t>(groups.size(), 1));
    cudaEvent_t ev_copied = c.events[N_FWD_STREAMS];
    CUDA_TRY(h, cudaEventRecord(ev_copied, st));
    // consecutive calls start on different streams: the heavy first group of call k+1 does not queue behind the one of call k
    const int rot = (int)(h->call_idx % (uint64_t)N_FWD_STREAMS);
    bool used[N_FWD_STREAMS] = {};
    Arena ar{ h->arena, h->d_pools, h->n_sm, 0u, h->d_clocks, 0u, g_skip_walk ? 1u : 0u, d_queue, d_tasks, 0u, (uint32_t)std::min<long>(h->tbw_long, 0x7FFFFFFF), g_corridor ? 1u : 0u };
    // The walkers first, on the highest-priority stream: one warp per SM that stays for the whole call and walks the pairs the
    // forward CTAs queue. (A forward CTA ends with its forward pass; its ~130 registers per thread go to the next pair instead of
    // idling through a latency-bound walk — walking inside the forward CTA was measured at 37 % of the CTA's life, 482 regions/s
    // on config 2 where the forward passes alone run at 688.)
    c.n_walkers = 0;
    if (d_queue) {
        const unsigned ctas = (unsigned)std::min<size_t>((size_t)h->sm_count * (size_t)h->walk_ctas_per_sm, (n_queued + 3) / 4);
        CUDA_TRY(h, cudaStreamWaitEvent(h->stream_walk, ev_copied, 0));
        // walkers (env SVP_WALKER): groups (default) — 4 pairs per warp on one instruction stream, 8 lanes each; lanes — one pair per lane
        static const bool lane_walkers = [] { const char* e = getenv("SVP_WALKER"); return e && std::string(e) == "lanes"; }();
        static const bool lean = [] { const char* e = getenv("SVP_WALK_LEAN"); return e && atoi(e) != 0; }();
        if (!lane_walkers) svp_walkqg_kernel<8><<<ctas, 32, 0, h->stream_walk>>>(d_seqs, ar, d_results, d_cigar, cs, (uint32_t)h->walk_wait_ns);
        else if (lean) svp_walkq_kernel<16><<<ctas, 32, 0, h->stream_walk>>>(d_seqs, ar, d_results, d_cigar, cs, (uint32_t)h->walk_wait_ns);
        else svp_walkq_kernel<1><<<ctas, 32, 0, h->stream_walk>>>(d_seqs, ar, d_results, d_cigar, cs, (uint32_t)h->walk_wait_ns);
        c.n_walkers = ctas;
        c.stats.tb_launches = 1;
    }
    for (size_t gi = 0; gi < groups.size(); ++gi) {
        const Group& g = groups[gi];
        FwdKernel kern = fwd_kernel(g.fam, g.kp, g.nwarp > 1, g.cluster > 1, g.ring, cs.two_piece != 0, true);
        if (!kern) return fail(h, SVP_ERR_UNSUPPORTED, "internal: no kernel variant for a planned group");
        const int si = (int)((gi + (size_t)rot) % (size_t)n_streams);
        if (sF[si] != st && !used[si]) CUDA_TRY(h, cudaStreamWaitEvent(sF[si], ev_copied, 0));
        used[si] = true;
        ar.use_global = g.huge ? 1u : 0u;
        ar.inline_walk = g.inline_walk ? 1u : 0u;
        CUDA_TRY(h, launch_group(kern, g.cluster, (unsigned)g.count, (unsigned)g.nwarp * 32u, g.smem, sF[si], d_tasks, d_order + g.first, d_seqs, ar,
                                 d_results, d_cigar, cs));
        c.stats.fwd_launches++;
    }
    if (n_invalid)
        svp_invalid_kernel<<<(unsigned)((n_invalid + 127) / 128), 128, 0, st>>>(d_tasks, d_order + inv_first, (int)n_invalid, d_results);
    for (int a = 0; a < N_FWD_STREAMS; ++a) {
        if (!used[a] || sF[a] == st) continue;
        CUDA_TRY(h, cudaEventRecord(c.events[a], sF[a]));
        CUDA_TRY(h, cudaStreamWaitEvent(st, c.events[a], 0));
    }
    if (c.n_walkers) {
        CUDA_TRY(h, cudaEventRecord(c.events[N_FWD_STREAMS + 1], h->stream_walk));
        CUDA_TRY(h, cudaStreamWaitEvent(st, c.events[N_FWD_STREAMS + 1], 0));
    }
    CUDA_TRY(h, cudaEventRecord(c.ev_t1, st));
    c.stats.waves = (uint32_t)groups.size();
    c.n_waves = 0;
    return SVP_OK;
}

// Enqueue the planned pairs of call set `c` (c.h_tasks[0..n)). wait = true: return when the results are complete and the
// call's times are collected; false: return after the enqueue (svp_submit_batch_device) — the next call is planned while this
// one runs, and its kernels start as CTA slots and arena space free up.
static int run_plan(svp_handle* h, CallSet& c, size_t n, const uint8_t* d_seqs, svp_result* d_results, uint32_t* d_cigar, bool wait) {
    const Consts& cs = h->cs;
    Task* tasks = c.h_tasks;
    c.stats = svp_stats{};
    c.n_waves = 0;
    if (n == 0) { h->stats = c.stats; return SVP_OK; }
    // pairs whose sequence images do not fit the shared memory of a CTA are reported per pair (SVP_ST_TOO_LONG)
    for (size_t k = 0; k < n; ++k) {
        Task& t = tasks[k];
        if (t.flags & (TASK_INVALID | TASK_RING)) continue;
        if (smem_need(t, cs, warps_of(t)) + SMEM_STATIC > (size_t)h->max_smem_optin) t.flags |= TASK_INVALID | TASK_TOO_LONG;      // (+ the kernels' static shared memory)
    }
    const int rc = h->waves ? run_waves(h, c, n, d_seqs, d_results, d_cigar, wait) : run_arena(h, c, n, d_seqs, d_results, d_cigar);
    if (rc != SVP_OK) return rc;
    for (size_t k = 0; k < n; ++k) if (!(tasks[k].flags & TASK_INVALID)) { c.stats.cells += tasks[k].cells; c.stats.tb_bytes += tasks[k].tb_bytes; }
    c.pending = true;
    h->call_idx++;
    CUDA_TRY(h, cudaGetLastError());
    if (!wait) return SVP_OK;
    return collect_call(h, c);
}

static int plan_and_run(svp_handle* h, const uint8_t* d_seqs, size_t packed_bytes, const uint64_t* seq_off, const uint32_t* seq_len, uint32_t n_seqs,
                        const svp_pair* pairs, uint32_t n_pairs, svp_result* d_results, uint32_t* d_cigar, size_t cigar_words, bool need_padding, bool wait) {
    CallSet* c = nullptr;
    int rc = begin_call(h, n_pairs, &c);
    if (rc != SVP_OK) return rc;
    // few pairs: optionally favour short steps (see latency_pairs)
    const int latency = ((long)n_pairs <= h->latency_pairs ? 1 : 0) | ((long)n_pairs <= h->small_call_pairs ? 2 : 0);
    // planning is the GPU's idle time in a blocking call: large batches are planned by several host threads (every pair is independent)
    static const int plan_threads = [] { const char* e = getenv("SVP_PLAN_THREADS"); int v = e ? atoi(e) : 4; return std::min(std::max(v, 1), 16); }();
    const int nt = n_pairs >= 8192 ? std::min<int>(plan_threads, (int)std::max(1u, std::thread::hardware_concurrency())) : 1;
    auto plan_range = [&](uint32_t lo, uint32_t hi) {
        for (uint32_t k = lo; k < hi; ++k) plan_pair(h, latency, pairs[k], k, seq_off, seq_len, n_seqs, packed_bytes, cigar_words, need_padding, c->h_tasks[k]);
    };
    if (nt <= 1) plan_range(0, n_pairs);
    else {
        std::vector<std::thread> workers;
        const uint32_t per = (n_pairs + (uint32_t)nt - 1) / (uint32_t)nt;
        for (int w = 1; w < nt; ++w) workers.emplace_back(plan_range, std::min(n_pairs, (uint32_t)w * per), std::min(n_pairs, (uint32_t)(w + 1) * per));
        plan_range(0, std::min(n_pairs, per));
        for (std::thread& w : workers) w.join();
    }
    return run_plan(h, *c, n_pairs, d_seqs, d_results, d_cigar, wait);
}

// Host-only view of the planner: how the pairs of a batch would run (no device needed; B200's shared-memory limit assumed).
extern "C" int svp_plan_pairs(const svp_scoring* scoring, size_t packed_bytes, const uint64_t* seq_off, const uint32_t* seq_len,
                              uint32_t n_seqs, const svp_pair* pairs, uint32_t n_pairs, size_t cigar_words, svp_pair_plan* out) {
    if (!scoring || !seq_off || !seq_len || (!pairs && n_pairs) || (!out && n_pairs)) return fail(nullptr, SVP_ERR_ARG, "null argument");
    svp_handle pseudo;                                   // never touches CUDA: only the fields plan_pair() reads
    std::string why;
    int rc = make_consts(*scoring, pseudo.cs, why);
    if (rc != SVP_OK) return fail(nullptr, rc, why);
    pseudo.max_smem_optin = 232448;                      // 227 KB: cudaDevAttrMaxSharedMemoryPerBlockOptin of sm_100
    read_env_knobs(&pseudo);
    const int latency = ((long)n_pairs <= pseudo.latency_pairs ? 1 : 0) | ((long)n_pairs <= pseudo.small_call_pairs ? 2 : 0);
    for (uint32_t k = 0; k < n_pairs; ++k) {
        Task t;
        plan_pair(&pseudo, latency, pairs[k], k, seq_off, seq_len, n_seqs, packed_bytes, cigar_words, false, t);
        svp_pair_plan& o = out[k];
        memset(&o, 0, sizeof(o));
        if (t.flags & TASK_INVALID) { o.flags = SVP_PLAN_INVALID; continue; }
        const int kp = kp_of(t), cl = cluster_of(t);
        const int nw = cl > 1 ? 16 : (t.band_w + 128 * kp - 1) / (128 * kp);
        o.kp = (uint32_t)kp; o.warps = (uint32_t)nw; o.cluster = (uint32_t)cl;
        o.flags = ((t.flags & TASK_S32) ? SVP_PLAN_S32 : 0u) | ((t.flags & TASK_REBASE) ? SVP_PLAN_REBASE : 0u) |
                  ((t.flags & TASK_RING) ? SVP_PLAN_WINDOWS : 0u) | (pseudo.cs.general ? SVP_PLAN_GENERAL : 0u);
        o.smem_bytes = (uint32_t)smem_need(t, pseudo.cs, nw);
        if (!(t.flags & TASK_RING) && o.smem_bytes + 64 > (uint32_t)pseudo.max_smem_optin) o.flags |= SVP_PLAN_TOO_LONG;
        o.cells = t.cells;
        o.tb_bytes = t.tb_bytes;
    }
    return SVP_OK;
}

extern "C" int svp_align_batch_device(svp_handle* h, const uint8_t* d_packed_seqs, size_t packed_bytes,
                                      const uint64_t* seq_off, const uint32_t* seq_len, uint32_t n_seqs,
                                      const svp_pair* pairs, uint32_t n_pairs, svp_result* d_results,
                                      uint32_t* d_cigar_buf, size_t cigar_words) {
    if (!h) return SVP_ERR_ARG;
    if (!d_packed_seqs || !seq_off || !seq_len || (!pairs && n_pairs) || !d_results || (!d_cigar_buf && cigar_words))
        return fail(h, SVP_ERR_ARG, "null argument");
    CUDA_TRY(h, cudaSetDevice(h->device));
    return plan_and_run(h, d_packed_seqs, packed_bytes, seq_off, seq_len, n_seqs, pairs, n_pairs, d_results, d_cigar_buf, cigar_words, true, true);
}

extern "C" int svp_submit_batch_device(svp_handle* h, const uint8_t* d_packed_seqs, size_t packed_bytes,
                                       const uint64_t* seq_off, const uint32_t* seq_len, uint32_t n_seqs,
                                       const svp_pair* pairs, uint32_t n_pairs, svp_result* d_results,
                                       uint32_t* d_cigar_buf, size_t cigar_words) {
    if (!h) return SVP_ERR_ARG;
    if (!d_packed_seqs || !seq_off || !seq_len || (!pairs && n_pairs) || !d_results || (!d_cigar_buf && cigar_words))
        return fail(h, SVP_ERR_ARG, "null argument");
    CUDA_TRY(h, cudaSetDevice(h->device));
    return plan_and_run(h, d_packed_seqs, packed_bytes, seq_off, seq_len, n_seqs, pairs, n_pairs, d_results, d_cigar_buf, cigar_words, true, false);
}

static int align_host(svp_handle* h, const uint8_t* packed_seqs, size_t packed_bytes, const uint64_t* seq_off, const uint32_t* seq_len,
                      uint32_t n_seqs, const svp_pair* pairs, uint32_t n_pairs, svp_result* results, uint32_t* cigar_buf, size_t cigar_words,
                      bool want_cigars) {
    CUDA_TRY(h, cudaSetDevice(h->device));
    const size_t padded = ((packed_bytes + 15) & ~(size_t)15) + 64;
    CUDA_TRY(h, h->d_seqs.reserve(padded));
    CUDA_TRY(h, h->d_results.reserve(std::max<size_t>(n_pairs, 1) * sizeof(svp_result)));
    CUDA_TRY(h, h->d_cigar.reserve(std::max<size_t>(cigar_words, 1) * sizeof(uint32_t)));
    CUDA_TRY(h, cudaMemsetAsync((uint8_t*)h->d_seqs.p + (packed_bytes & ~(size_t)15), 0, padded - (packed_bytes & ~(size_t)15), h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(h->d_seqs.p, packed_seqs, packed_bytes, cudaMemcpyHostToDevice, h->stream));
    // a submitted device-resident call may still read the handle's sequence / result buffers only if it came through this
    // function, which always waits: nothing of ours is in flight on them here
    static const bool stream_back = [] { const char* e = getenv("SVP_STREAM_BACK"); return e ? atoi(e) != 0 : true; }();
    if (want_cigars && h->waves && stream_back && n_pairs) {
        // Wave scheduling, several waves: the rows and CIGARs of a finished wave are copied back while the next wave computes (a wave is
        // a range of consecutive pairs, so its part of the dense CIGAR layout follows the previous wave's). The dense staging buffer is
        // sized before the kernels start: growing it later would wait for the device.
        CUDA_TRY(h, h->d_dense.reserve(std::max<size_t>(cigar_words / 2, 1) * sizeof(uint32_t)));
        CUDA_TRY(h, h->d_dense_off.reserve(n_pairs * sizeof(uint32_t)));
        int rc = plan_and_run(h, (const uint8_t*)h->d_seqs.p, padded, seq_off, seq_len, n_seqs, pairs, n_pairs, (svp_result*)h->d_results.p,
                              (uint32_t*)h->d_cigar.p, cigar_words, false, false);
        if (rc != SVP_OK) return rc;
        CallSet& c = h->sets[(h->call_idx - 1) & 1];
        if (c.n_waves >= 2 && c.pending) {
            cudaStream_t sc = h->stream_copy;
            std::vector<uint32_t> dense_off(n_pairs), rel(n_pairs);
            uint64_t total = 0;
            bool too_small = false;
            for (size_t wv = 0; wv < c.n_waves && !too_small; ++wv) {
                const size_t first = c.wave_ranges[wv].first, last = c.wave_ranges[wv].second, cnt = last - first;
                if (!cnt) continue;
                CUDA_TRY(h, cudaStreamWaitEvent(sc, c.wave_events[wv * EV_PER_WAVE + 2], 0));            // this wave's walks are done
                CUDA_TRY(h, cudaMemcpyAsync(results + first, (const svp_result*)h->d_results.p + first, cnt * sizeof(svp_result), cudaMemcpyDeviceToHost, sc));
                CUDA_TRY(h, cudaStreamSynchronize(sc));
                uint64_t wave_total = 0;
                for (size_t k = first; k < last; ++k) { rel[k] = (uint32_t)wave_total; dense_off[k] = (uint32_t)(total + wave_total); wave_total += results[k].cigar_len; }
                if (total + wave_total > cigar_words) { too_small = true; break; }
                if (wave_total) {
                    if (wave_total * sizeof(uint32_t) > h->d_dense.cap) {          // (rare: more than half of the slots used) — wait, then grow
                        CUDA_TRY(h, cudaDeviceSynchronize());
                        CUDA_TRY(h, h->d_dense.reserve(wave_total * sizeof(uint32_t)));
                    }
                    CUDA_TRY(h, cudaMemcpyAsync((uint32_t*)h->d_dense_off.p + first, rel.data() + first, cnt * sizeof(uint32_t), cudaMemcpyHostToDevice, sc));
                    const int wpb = 8;
                    svp_cigar_gather<<<(unsigned)((cnt + wpb - 1) / wpb), wpb * 32, 0, sc>>>((const svp_result*)h->d_results.p + first, (const uint32_t*)h->d_dense_off.p + first,
                                                                                              (const uint32_t*)h->d_cigar.p, (uint32_t*)h->d_dense.p, (int)cnt);
                    CUDA_TRY(h, cudaMemcpyAsync(cigar_buf + total, h->d_dense.p, wave_total * sizeof(uint32_t), cudaMemcpyDeviceToHost, sc));
                }
                total += wave_total;
            }
            CUDA_TRY(h, cudaStreamSynchronize(sc));
            rc = collect_call(h, c);
            if (rc != SVP_OK) return rc;
            if (too_small) return fail(h, SVP_ERR_ARG, "cigar_buf smaller than the CIGARs produced");
            for (uint32_t k = 0; k < n_pairs; ++k) results[k].cigar_off = dense_off[k];
            h->stats.reserved = 1;
            CUDA_TRY(h, cudaStreamSynchronize(h->stream));
            return SVP_OK;
        }
        rc = collect_call(h, c);             // one wave: nothing to overlap with — the plain copy-back below
        if (rc != SVP_OK) return rc;
    } else {
        int rc = plan_and_run(h, (const uint8_t*)h->d_seqs.p, padded, seq_off, seq_len, n_seqs, pairs, n_pairs, (svp_result*)h->d_results.p,
                              (uint32_t*)h->d_cigar.p, cigar_words, false, true);
        if (rc != SVP_OK) return rc;
    }
    if (n_pairs) {
        CUDA_TRY(h, cudaMemcpyAsync(results, h->d_results.p, n_pairs * sizeof(svp_result), cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        if (want_cigars) {
            // dense CIGAR layout in pair order; only the used words are copied back
            std::vector<uint32_t> dense_off(n_pairs);
            uint64_t total = 0;
            for (uint32_t k = 0; k < n_pairs; ++k) { dense_off[k] = (uint32_t)total; total += results[k].cigar_len; }
            if (total > cigar_words) return fail(h, SVP_ERR_ARG, "cigar_buf smaller than the CIGARs produced");
            if (total) {
                CUDA_TRY(h, h->d_dense.reserve(total * sizeof(uint32_t)));
                CUDA_TRY(h, h->d_dense_off.reserve(n_pairs * sizeof(uint32_t)));
                CUDA_TRY(h, cudaMemcpyAsync(h->d_dense_off.p, dense_off.data(), n_pairs * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
                const int wpb = 8;
                svp_cigar_gather<<<(n_pairs + wpb - 1) / wpb, wpb * 32, 0, h->stream>>>((const svp_result*)h->d_results.p, (const uint32_t*)h->d_dense_off.p,
                                                                                      (const uint32_t*)h->d_cigar.p, (uint32_t*)h->d_dense.p, (int)n_pairs);
                CUDA_TRY(h, cudaMemcpyAsync(cigar_buf, h->d_dense.p, total * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
            }
            CUDA_TRY(h, cudaStreamSynchronize(h->stream));
            for (uint32_t k = 0; k < n_pairs; ++k) results[k].cigar_off = dense_off[k];
        } else {
            for (uint32_t k = 0; k < n_pairs; ++k) results[k].cigar_off = 0;
        }
        h->stats.reserved = 1;
    }
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return SVP_OK;
}

extern "C" int svp_align_batch(svp_handle* h, const uint8_t* packed_seqs, size_t packed_bytes,
                               const uint64_t* seq_off, const uint32_t* seq_len, uint32_t n_seqs,
                               const svp_pair* pairs, uint32_t n_pairs, svp_result* results,
                               uint32_t* cigar_buf, size_t cigar_words) {
    if (!h) return SVP_ERR_ARG;
    if (!packed_seqs || !seq_off || !seq_len || (!pairs && n_pairs) || !results || (!cigar_buf && cigar_words))
        return fail(h, SVP_ERR_ARG, "null argument");
    return align_host(h, packed_seqs, packed_bytes, seq_off, seq_len, n_seqs, pairs, n_pairs, results, cigar_buf, cigar_words, true);
}

extern "C" int svp_align_batch_summary(svp_handle* h, const uint8_t* packed_seqs, size_t packed_bytes,
                                       const uint64_t* seq_off, const uint32_t* seq_len, uint32_t n_seqs,
                                       const svp_pair* pairs, uint32_t n_pairs, svp_result* results, size_t cigar_words) {
    if (!h) return SVP_ERR_ARG;
    if (!packed_seqs || !seq_off || !seq_len || (!pairs && n_pairs) || !results) return fail(h, SVP_ERR_ARG, "null argument");
    return align_host(h, packed_seqs, packed_bytes, seq_off, seq_len, n_seqs, pairs, n_pairs, results, nullptr, cigar_words, false);
}
