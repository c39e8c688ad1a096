/*
 * svp_align.h — C-ABI of the B200 pairwise-alignment hot path (libsvp_align.so).
 *
 * The reference (bihealth/svirlpool) has no FFI for this path: every alignment is a
 * subprocess call to minimap2 through
 *     util.align_reads_with_minimap()          src/svirlpool/util/util.py:162-238
 * issued at
 *     AVA  (read-vs-read, one container)       src/svirlpool/localassembly/consensus.py:1378-1386
 *     reads -> consensus                       src/svirlpool/localassembly/consensus.py:851-857
 *     consensus -> reference                   src/svirlpool/localassembly/consensus_align.py:1977-1987
 * and its SAM output is consumed by
 *     consensus_lib.parse_directed_signals_from_ava()      consensus_lib.py:279-345
 *     alignments_to_rafs.parse_SVsignals_from_alignment()  alignments_to_rafs.py:174-258
 * The entry points below are what a ctypes/cffi binding added at that seam would call
 * (see INTEGRATION.md). Plain pointers and sizes only; no torch types; never throws.
 *
 * The same declarations are implemented twice:
 *   svp_*   — CUDA (sm_100a) implementation, svirlpool_b200/csrc/  -> libsvp_align.so
 *   svpo_*  — CPU oracle (test infrastructure only), oracle/svp_oracle.c -> liboracle
 * The alignment semantics both must satisfy bit-for-bit are in DESIGN.md §3 (ALIGN_SPEC).
 */
#ifndef SVP_ALIGN_H
#define SVP_ALIGN_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SVP_ABI_VERSION 1

/* return codes */
#define SVP_OK               0
#define SVP_ERR_ARG         -1   /* invalid argument (see svp_last_error) */
#define SVP_ERR_CUDA        -2   /* CUDA runtime failure */
#define SVP_ERR_NOMEM       -3   /* device or host allocation failed */
#define SVP_ERR_UNSUPPORTED -4   /* scoring/limits outside what the kernels implement */
#define SVP_ERR_NODEVICE    -5   /* no CUDA device: there is NO CPU fallback */

/* per-pair status bits (svp_result.status) */
#define SVP_ST_OK            0u
#define SVP_ST_NOHIT         1u  /* best local score == 0: no alignment ("unmapped") */
#define SVP_ST_CIGAR_OVF     2u  /* CIGAR longer than the pair's slot; counts/coords still valid */
#define SVP_ST_SCORE_OVF     4u  /* score range exceeded the kernel's lane type (must not happen) */
#define SVP_ST_BAD_PAIR      8u  /* pair descriptor invalid (index, band), or band wider than a cluster covers */
#define SVP_ST_TOO_LONG     16u  /* the pair's sequences do not fit the kernels' shared-memory images (m + n beyond ~226 k bases) */

/* CIGAR op codes: BAM encoding, word = len << 4 | op */
#define SVP_CIGAR_M 0u
#define SVP_CIGAR_I 1u
#define SVP_CIGAR_D 2u

/* Scoring profile. Two presets are provided by svp_scoring_preset():
 *   "map-ont"/"ava-ont": minimap2 defaults A=2 B=4 O=4,24 E=2,1 (verified on the 20 frozen
 *       reference alignments)
 *   "last": single-piece affine with a 4x4 matrix and separate ins/del costs, as trained by
 *       last-train into a lamassemble .mat (consensus.py:777). */
typedef struct svp_scoring {
    int32_t matrix[16];   /* substitution score s[q*4+t], q,t in 0..3 (A,C,G,T) */
    int32_t del_open1, del_ext1;   /* deletion (target-only, CIGAR D) piece 1: cost(k) = open + k*ext */
    int32_t ins_open1, ins_ext1;   /* insertion (query-only, CIGAR I) piece 1 */
    int32_t del_open2, del_ext2;   /* piece 2 (long gaps); open2 < 0 disables the second piece */
    int32_t ins_open2, ins_ext2;
    int32_t min_signal_size;       /* indel runs >= this enter sv_dist (consensus.py:1441-1445: 12) */
    int32_t zdrop;                 /* Z-trim threshold (minimap2 -z, default 400); 0 disables */
    int32_t zdrop_ext;             /* per-diagonal allowance of the Z-trim test (minimap2: the long-gap extension e2) */
    int32_t reserved;
} svp_scoring;

/* One alignment problem: query sequence index, target sequence index (into the batch's
 * sequence table), strand, and the diagonal band d = t_pos - q_pos in [band_lo, band_lo+band_w).
 * band_w must be a positive multiple of 32. */
typedef struct svp_pair {
    uint32_t q;          /* query sequence index */
    uint32_t t;          /* target sequence index */
    int32_t  band_lo;    /* lowest diagonal of the band */
    uint32_t band_w;     /* number of diagonals (multiple of 32) */
    uint32_t flags;      /* bit0: align the reverse complement of the query */
    uint32_t cigar_off;  /* first word of this pair's slot in the cigar buffer */
    uint32_t cigar_cap;  /* slot size in words */
    uint32_t reserved;
} svp_pair;
#define SVP_PAIR_REVCOMP 1u

typedef struct svp_result {
    int32_t  score;
    uint32_t status;
    int32_t  q_beg, q_end;   /* aligned interval on the ORIENTED query (after revcomp), 0-based half-open */
    int32_t  t_beg, t_end;   /* aligned interval on the target */
    uint32_t n_match, n_mismatch;
    uint32_t n_ins_ops, n_ins_bp;
    uint32_t n_del_ops, n_del_bp;
    uint32_t cigar_off;      /* first word of the CIGAR (inside the pair's slot) */
    uint32_t cigar_len;      /* number of words */
    double   sv_dist;        /* sum over indel runs >= min_signal_size of (1-exp(-(s/30)^1.5))*s
                                (consensus_lib.py:24-37, 353-396 with unit weights) */
    uint64_t cells;          /* DP cells of the spec (band ∩ matrix) for this pair */
} svp_result;

typedef struct svp_handle svp_handle;

int         svp_abi_version(void);
int         svp_device_count(void);
int         svp_scoring_preset(const char* name, svp_scoring* out);
/* device < 0: current device. max_tb_bytes: size of the device traceback arena (0 = default: 64 GB, at most 60 % of the free
 * device memory, so that several handles / processes share a device). With the default wave scheduler a call is cut into waves of
 * pairs whose traceback rows fit the arena (or half of it), so a caller that owns the GPU and submits large batches should pass more
 * (INTEGRATION.md "Memory and scheduling"); with SVP_MODE=arena / queue the arena only holds the rows of the pairs resident on the
 * GPU. A pair whose rows exceed the arena makes the call fail with SVP_ERR_NOMEM. */
int         svp_create(int device, const svp_scoring* scoring, size_t max_tb_bytes, svp_handle** out);
void        svp_destroy(svp_handle* h);
const char* svp_last_error(const svp_handle* h);

/* Page-locked host memory for the buffers handed to svp_align_batch() (optional; pageable memory
 * works, pinned memory makes the H2D/D2H copies asynchronous and fast). NULL on failure. */
void*       svp_pinned_alloc(size_t bytes);
void        svp_pinned_free(void* p);

/* Letters outside A, C, G, T (N, IUPAC codes). The 2-bit alphabet has no fifth symbol; a sequence that contains such letters
 * carries a 1-bit-per-base mask behind its 2-bit codes ("+1 bit/base mask only if N present"):
 *     seq_off[s] | SVP_SEQ_NMASK      the sequence has a mask: ceil(len/8) bytes starting at byte offset
 *                                     (seq_off[s] & ~15) + SVP_SEQ_CODE_BYTES(len); bit k%8 of its byte k/8 = base k is masked
 * (seq_off entries are multiples of 16, so their low bits are free). A cell with a masked base on either side scores the LOWEST
 * entry of the substitution matrix and counts as a mismatch — LAST's treatment of letters outside its matrix; with minimap2's
 * scoring that is the mismatch score — so an N never matches anything. The 2-bit code stored for a masked
 * base is ignored. svp_pack_2bit() writes the masks. */
#define SVP_SEQ_NMASK 1ull
#define SVP_SEQ_CODE_BYTES(len) ((((uint64_t)(len) + 3) / 4 + 15) & ~(uint64_t)15)

/* Align n_pairs pairs. Sequences are 2-bit packed (A=0,C=1,G=2,T=3; base k of a sequence lives
 * in bits 2*(k%4).. of byte seq_off[s] + k/4); seq_off[] entries must be multiples of 16 (plus the flag above).
 * All buffers are HOST memory; results and CIGAR words are written to caller-allocated arrays.
 * On return the CIGARs are packed densely from cigar_buf[0] in pair order and result.cigar_off
 * points at each of them (the per-pair slots of svp_pair only size the device scratch; a pair whose
 * CIGAR does not fit its slot gets SVP_ST_CIGAR_OVF). Blocking; one call at a time per handle. */
int svp_align_batch(svp_handle* h,
                    const uint8_t* packed_seqs, size_t packed_bytes,
                    const uint64_t* seq_off, const uint32_t* seq_len, uint32_t n_seqs,
                    const svp_pair* pairs, uint32_t n_pairs,
                    svp_result* results,
                    uint32_t* cigar_buf, size_t cigar_words);

/* Summary variant for consumers that need no CIGARs (the consensus_lib similarity matrix with unit density weights:
 * score, intervals, counts and sv_dist are all it reads; consensus_lib.py:353-413, consensus.py:1276-1305): identical
 * to svp_align_batch() but nothing except the result rows is copied back. cigar_words still sizes the device scratch
 * the traceback writes (svp_pair.cigar_off / cigar_cap); result.cigar_off is meaningless on return. */
int svp_align_batch_summary(svp_handle* h,
                            const uint8_t* packed_seqs, size_t packed_bytes,
                            const uint64_t* seq_off, const uint32_t* seq_len, uint32_t n_seqs,
                            const svp_pair* pairs, uint32_t n_pairs,
                            svp_result* results, size_t cigar_words);

/* Device-resident variant: the bulk data — packed sequences in, results and CIGAR words out — are
 * DEVICE pointers on the handle's device (e.g. torch tensor data_ptr()); the launch plan (sequence
 * table and pair descriptors) stays in HOST memory. Every sequence must be readable up to the next
 * 16-byte boundary. Used by bench.py for the HBM-resident `value` measurement. Blocking. */
int svp_align_batch_device(svp_handle* h,
                           const uint8_t* d_packed_seqs, size_t packed_bytes,
                           const uint64_t* seq_off, const uint32_t* seq_len, uint32_t n_seqs,
                           const svp_pair* pairs, uint32_t n_pairs,
                           svp_result* d_results,
                           uint32_t* d_cigar_buf, size_t cigar_words);

/* Timing / accounting of the last svp_align_batch*() call on this handle (CUDA events on the
 * stream the kernels ran on). */
/* ---- host-side batch builders (no device needed) ---------------------------------------------------------------
 * What the reference does by writing all_reads.fasta for minimap2 (consensus.py:1362-1369) and letting
 * `minimap2 -x ava-ont -X` choose the read pairs (util/util.py:181-183) is two calls here; their outputs are the
 * arguments of svp_align_batch().
 *
 * svp_pack_bytes   bytes svp_pack_2bit() needs for these sequences in the worst case (every sequence starts on a 16-byte boundary
 *                  and may carry a mask, 16 readable bytes follow the last one).
 * svp_pack_2bit    ASCII (ACGT, either case) -> 2 bits per base, base k of a sequence in bits 2*(k%4) of its byte k/4; a sequence
 *                  with any other letter gets a mask behind its codes (SVP_SEQ_NMASK above; the letter's code is 0); fills
 *                  seq_off[k] with the byte offset of sequence k (| SVP_SEQ_NMASK).
 * svp_ava_pairs    pair list of one container's all-vs-all alignment: reads seq_base .. seq_base+n_reads-1 of the sequence
 *                  table, every unordered pair once, query = the read that comes first in name_order (n_reads indices,
 *                  the reads sorted by name; NULL = table order). strand: +1 / -1 per read when the strands are known
 *                  (one orientation per pair), NULL or 0 = unknown (both orientations). offsets: start of each read on a
 *                  common coordinate (band starts at their difference), or NULL. Bands as in DESIGN.md §3.1: around the
 *                  line from the start diagonal to n - m, +-slack, rounded up to `granule` (multiple of 32); full != 0:
 *                  the whole matrix. CIGAR slots are laid out from *cigar_words on, which is advanced. Returns
 *                  SVP_ERR_NOMEM with *n_pairs = the number of pairs when pairs_cap is too small. */
size_t svp_pack_bytes(const uint32_t* seq_len, uint32_t n_seqs);
int    svp_pack_2bit(const char* const* seqs, const uint32_t* seq_len, uint32_t n_seqs,
                     uint8_t* packed, size_t packed_bytes, uint64_t* seq_off);
int    svp_ava_pairs(const uint32_t* seq_len, const uint32_t* name_order, const int8_t* strand, const int32_t* offsets,
                     uint32_t seq_base, uint32_t n_reads, int32_t slack, int32_t granule, int full,
                     svp_pair* pairs, uint32_t pairs_cap, uint32_t* n_pairs, uint64_t* cigar_words);

/* Pipelined form of svp_align_batch_device: returns as soon as the batch is enqueued. Results (and the caller's buffers)
 * belong to the library until svp_sync() returns. Consecutive submits overlap on the device: the next batch is planned on
 * the host and its CTAs start as those of the previous batch retire (two batches in flight at most; a third submit first
 * waits for the oldest). svp_sync waits for every submitted batch. */
int svp_submit_batch_device(svp_handle* h,
                            const uint8_t* d_packed_seqs, size_t packed_bytes,
                            const uint64_t* seq_off, const uint32_t* seq_len, uint32_t n_seqs,
                            const svp_pair* pairs, uint32_t n_pairs,
                            svp_result* d_results, uint32_t* d_cigar_buf, size_t cigar_words);
int svp_sync(svp_handle* h);

/* Host-only view of the planner (no device needed): how each pair of a batch would run — for capacity planning (traceback bytes
 * = arena use, shared memory per CTA) and for tests of the planning rules. Same arguments as svp_align_batch. */
#define SVP_PLAN_S32      1u   /* 32-bit lanes */
#define SVP_PLAN_REBASE   2u   /* 16-bit lanes relative to per-warp score offsets (scores beyond the s16 range) */
#define SVP_PLAN_WINDOWS  4u   /* sliding sequence windows instead of whole shared-memory images */
#define SVP_PLAN_GENERAL  8u   /* general scoring kernels (4x4 matrix, separate insertion / deletion costs) */
#define SVP_PLAN_INVALID 16u   /* descriptor rejected (SVP_ST_BAD_PAIR) */
#define SVP_PLAN_TOO_LONG 32u  /* sequences do not fit the shared-memory images (SVP_ST_TOO_LONG) */
typedef struct svp_pair_plan {
    uint32_t flags;        /* SVP_PLAN_* */
    uint32_t kp;           /* packed registers per state array: a thread owns 4*kp diagonals */
    uint32_t warps;        /* warps per CTA */
    uint32_t cluster;      /* CTAs per pair (thread-block cluster), 1 = none */
    uint32_t smem_bytes;   /* dynamic shared memory of the CTA */
    uint32_t reserved;
    uint64_t cells;        /* DP cells of the spec (band ∩ matrix) */
    uint64_t tb_bytes;     /* traceback bytes this pair occupies in the arena */
} svp_pair_plan;
int svp_plan_pairs(const svp_scoring* scoring, size_t packed_bytes, const uint64_t* seq_off, const uint32_t* seq_len,
                   uint32_t n_seqs, const svp_pair* pairs, uint32_t n_pairs, size_t cigar_words, svp_pair_plan* out);

typedef struct svp_stats {
    double   fwd_ms;        /* wave scheduling: length of the union of the waves' forward intervals; arena / queue: the call's kernel time */
    double   tb_ms;         /* wave scheduling: summed spans of the walk kernels; arena: fwd_ms x the walks' share of the CTAs' phase times;
                             * queue: mean busy time of the walker warps (informational) */
    double   total_ms;      /* first enqueue -> last kernel end */
    uint64_t cells;         /* spec cells aligned */
    uint64_t tb_bytes;      /* traceback bytes written to HBM */
    uint32_t fwd_launches;  /* kernels launched */
    uint32_t tb_launches;   /* walk kernels launched (0 when every CTA walks its own pair) */
    uint32_t waves;         /* wave scheduling: waves; arena / queue: launch groups (kernel variant x CTA width x size class) */
    uint32_t reserved;
} svp_stats;
int svp_get_stats(const svp_handle* h, svp_stats* out);
/* sums over all completed batches since svp_create or the last call with reset != 0 (waits for submitted batches first) */
int svp_get_stats_total(svp_handle* h, svp_stats* out, int reset);

/* Device-side stopwatch for callers that time several calls as one region (bench.py): mark 0 and mark 1 are
 * CUDA events recorded on the handle's main stream, which every kernel and copy of a call is ordered before
 * when the call returns. svp_timer_elapsed_ms() waits for mark 1 and returns the device time between the marks
 * (host gaps between calls included), or a negative SVP_ERR_* code. */
int    svp_timer_mark(svp_handle* h, int which);
double svp_timer_elapsed_ms(svp_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* SVP_ALIGN_H */
