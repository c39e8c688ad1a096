/*
 * svp_oracle.c — CPU restatement of the alignment spec (ALIGN_SPEC v1).
 *
 * TEST INFRASTRUCTURE ONLY. Nothing under svirlpool_b200/ may import, link or execute this
 * file; it is used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs as the checker and the timed CPU arm.
 *
 * Provenance / parity status:
 *   The reference (bihealth/svirlpool) contains no alignment arithmetic: the path shells out
 *   to minimap2 2.31 (util/util.py:162-238; call sites consensus.py:1378-1386, :851-857,
 *   consensus_align.py:1977-1987), whose source is not vendored. This file restates the
 *   published algorithm those calls rely on — affine-gap local alignment with minimap2's
 *   default two-piece gap cost (A=2 B=4 O=4,24 E=2,1) — as an explicit spec, and is pinned
 *   against the reference's own frozen minimap2 outputs (tests/golden/, derived from
 *   /root/reference/tests/data by tests/golden/make_golden.py):
 *     - AS == score of the reported CIGAR under this scoring: all 20 frozen records
 *     - CIGAR, strand, coordinates: the 14 records whose inputs are available
 *   AVA mode (-x ava-ont) and lamassemble .mat scoring have no fixture in the reference:
 *   PARITY UNPINNED for those.
 *
 * Plain C11, int32 scores, full band matrix in memory, pthreads over pairs.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#include "../include/svp_align.h"

#define NEG (-(1 << 29))

static int g_threads = 1;

int svpo_abi_version(void) { return SVP_ABI_VERSION; }

void svpo_set_threads(int n) { g_threads = n < 1 ? 1 : n; }

/* minimap2 presets: the defaults of `-x map-ont` / `-x ava-ont` (-A2 -B4 -O4,24 -E2,1), the
 * scoring of every aligner call on the path (verified on fixtures §8c).
 * "last": single-piece affine, separate ins/del costs, full 4x4 matrix — the shape of a
 * last-train .mat (consensus.py:777). The numbers here are a LABELLED STAND-IN (the real
 * promethion.mat is a Git-LFS pointer in the reference checkout), not trained values. */
int svpo_scoring_preset(const char* name, svp_scoring* s) {
    if (!name || !s) return SVP_ERR_ARG;
    memset(s, 0, sizeof(*s));
    s->min_signal_size = 12;
    if (!strcmp(name, "map-ont") || !strcmp(name, "ava-ont")) {
        for (int a = 0; a < 4; ++a)
            for (int b = 0; b < 4; ++b) s->matrix[a * 4 + b] = (a == b) ? 2 : -4;
        s->del_open1 = s->ins_open1 = 4;  s->del_ext1 = s->ins_ext1 = 2;
        s->del_open2 = s->ins_open2 = 24; s->del_ext2 = s->ins_ext2 = 1;
        s->zdrop = 400; s->zdrop_ext = 1;
        return SVP_OK;
    }
    if (!strcmp(name, "last")) {
        static const int m[16] = { 5, -8, -4, -9,   -8, 6, -9, -4,   -4, -9, 6, -8,   -9, -4, -8, 5 };
        memcpy(s->matrix, m, sizeof(m));
        s->del_open1 = 10; s->del_ext1 = 3;
        s->ins_open1 = 12; s->ins_ext1 = 2;
        s->del_open2 = s->ins_open2 = -1; s->del_ext2 = s->ins_ext2 = 0;
        s->zdrop = 400; s->zdrop_ext = 2;
        return SVP_OK;
    }
    return SVP_ERR_ARG;
}

static inline int base_at(const uint8_t* p, uint32_t k) { return (p[k >> 2] >> ((k & 3) * 2)) & 3; }

/* Letters outside A, C, G, T (N, IUPAC codes) are flagged by the optional 1-bit-per-base mask that follows a sequence's 2-bit
 * codes (include/svp_align.h: SVP_SEQ_NMASK); unpacked, such a base has code 4. A cell with code 4 on either side scores the
 * LOWEST entry of the substitution matrix and counts as a mismatch (LAST gives letters outside its matrix the minimum score;
 * for minimap2's scoring the minimum is the mismatch score). DESIGN.md §3.1. */
static inline int sub_score(const svp_scoring* sc, int n_score, int qi, int tj) {
    return (qi > 3 || tj > 3) ? n_score : sc->matrix[qi * 4 + tj];
}
static inline int mask_at(const uint8_t* mask, uint32_t k) { return (mask[k >> 3] >> (k & 7)) & 1; }

static inline int gap_cost(int open1, int ext1, int open2, int ext2, int k) {
    int c = open1 + k * ext1;
    if (open2 >= 0) { int c2 = open2 + k * ext2; if (c2 < c) c = c2; }
    return c;
}

/* ALIGN_SPEC v1, one pair. q,t: unpacked base codes (q already oriented). */
static void align_one(const svp_scoring* sc, const uint8_t* q, int m, const uint8_t* t, int n,
                      int band_lo, int W, svp_result* res, uint32_t* cigar_slot, uint32_t slot_off,
                      uint32_t cap) {
    const int d_hi = band_lo + W - 1;
    memset(res, 0, sizeof(*res));
    res->cigar_off = slot_off;
    const int two_d = sc->del_open2 >= 0, two_i = sc->ins_open2 >= 0;
    const int doe1 = sc->del_open1 + sc->del_ext1, de1 = sc->del_ext1;
    const int ioe1 = sc->ins_open1 + sc->ins_ext1, ie1 = sc->ins_ext1;
    const int doe2 = sc->del_open2 + sc->del_ext2, de2 = sc->del_ext2;
    const int ioe2 = sc->ins_open2 + sc->ins_ext2, ie2 = sc->ins_ext2;
    int n_score = sc->matrix[0];
    for (int k = 1; k < 16; ++k) if (sc->matrix[k] < n_score) n_score = sc->matrix[k];

    /* H stored for rows 0..m over the band; row i holds j = i+band_lo .. i+d_hi */
    size_t rows = (size_t)m + 1;
    int32_t* H = (int32_t*)calloc(rows * (size_t)W, sizeof(int32_t));
    int32_t* F1 = (int32_t*)malloc(((size_t)n + 2) * sizeof(int32_t));
    int32_t* F2 = (int32_t*)malloc(((size_t)n + 2) * sizeof(int32_t));
    if (!H || !F1 || !F2) { res->status = SVP_ST_BAD_PAIR; free(H); free(F1); free(F2); return; }
#define HAT(i, j) H[(size_t)(i) * W + ((j) - (i) - band_lo)]

    int best = 0, best_i = 0, best_j = 0;
    uint64_t cells = 0;
    for (int i = 1; i <= m; ++i) {
        long jl = (long)i + band_lo, jh = (long)i + d_hi;
        int jlo = jl < 1 ? 1 : (int)jl;
        int jhi = jh > n ? n : (int)jh;
        if (jlo > jhi) continue;
        cells += (uint64_t)(jhi - jlo + 1);
        int E1 = NEG, E2 = NEG, Hleft = 0;      /* cell left of jlo is boundary or out of band */
        const int qi = q[i - 1];
        for (int j = jlo; j <= jhi; ++j) {
            const int d = j - i;
            /* up neighbour (i-1, j): diagonal d+1 */
            int Hup = 0, F1up = NEG, F2up = NEG;
            if (i - 1 >= 1 && d + 1 <= d_hi) { Hup = HAT(i - 1, j); F1up = F1[j]; F2up = F2[j]; }
            /* diagonal neighbour (i-1, j-1): same diagonal; boundary row/col is 0 */
            int Hd = (i - 1 >= 1 && j - 1 >= 1) ? HAT(i - 1, j - 1) : 0;
            int e1 = E1 - de1; { int o = Hleft - doe1; if (o > e1) e1 = o; }
            int f1 = F1up - ie1; { int o = Hup - ioe1; if (o > f1) f1 = o; }
            int h = Hd + sub_score(sc, n_score, qi, t[j - 1]);
            if (e1 > h) h = e1;
            if (f1 > h) h = f1;
            int e2 = NEG, f2 = NEG;
            if (two_d) { e2 = E2 - de2; int o = Hleft - doe2; if (o > e2) e2 = o; if (e2 > h) h = e2; }
            if (two_i) { f2 = F2up - ie2; int o = Hup - ioe2; if (o > f2) f2 = o; if (f2 > h) h = f2; }
            if (h < 0) h = 0;
            HAT(i, j) = h;
            E1 = e1; E2 = e2; F1[j] = f1; F2[j] = f2; Hleft = h;
            if (h > best || (h == best && h > 0 && (i + j < best_i + best_j))) {
                /* row-major scan: for equal r the smaller i was seen first, so strict '<' on r */
                best = h; best_i = i; best_j = j;
            }
        }
    }
    res->cells = cells;
    res->score = best;
    if (best <= 0) {
        res->status = SVP_ST_NOHIT;
        free(H); free(F1); free(F2);
        return;
    }

    /* traceback (ALIGN_SPEC §3.4) with Z-trim (§3.5): runs emitted end -> start, then reversed */
    uint32_t* ops = NULL; size_t nops = 0, ops_cap = 0;
    uint32_t cur_op = 0, cur_len = 0;
    int i = best_i, j = best_j;
    uint32_t status = 0;
    /* snapshot of the walk at the lowest H seen so far (the alignment start if the walk is cut) */
    int low = best + 1, low_i = best_i, low_j = best_j;
    size_t low_nops = 0; uint32_t low_op = 0, low_len = 0;
    uint32_t low_cnt[6] = {0, 0, 0, 0, 0, 0};
#define PUSH_RUN() do { if (cur_len) { if (nops == ops_cap) { ops_cap = ops_cap ? ops_cap * 2 : 256; ops = (uint32_t*)realloc(ops, ops_cap * sizeof(uint32_t)); } ops[nops++] = (cur_len << 4) | cur_op; } } while (0)
#define EMIT(op, len) do { if (cur_len && cur_op == (op)) cur_len += (len); else { PUSH_RUN(); cur_op = (op); cur_len = (len); } } while (0)
#define SNAPSHOT(hval) do { low = (hval); low_i = i; low_j = j; low_nops = nops; low_op = cur_op; low_len = cur_len; \
        low_cnt[0] = res->n_match; low_cnt[1] = res->n_mismatch; low_cnt[2] = res->n_ins_ops; low_cnt[3] = res->n_ins_bp; \
        low_cnt[4] = res->n_del_ops; low_cnt[5] = res->n_del_bp; } while (0)
    for (;;) {
        const int h = (i >= 1 && j >= 1) ? HAT(i, j) : 0;
        if (h < low) SNAPSHOT(h);
        if (h == 0) break;
        const int d = j - i;
        if (sc->zdrop > 0) {
            int dd = d - (low_j - low_i); if (dd < 0) dd = -dd;
            if (h - low > sc->zdrop + sc->zdrop_ext * dd) break;      /* Z-trim: cut at the snapshot */
        }
        const int Hd = (i - 1 >= 1 && j - 1 >= 1) ? HAT(i - 1, j - 1) : 0;
        const int qi = q[i - 1], tj = t[j - 1];
        if (h == Hd + sub_score(sc, n_score, qi, tj)) {
            EMIT(SVP_CIGAR_M, 1);
            if (qi == tj && qi <= 3) res->n_match++; else res->n_mismatch++;
            --i; --j;
            continue;
        }
        int found = 0;
        for (int k = 1;; ++k) {
            int del_ok = (j - k >= 1) && (d - k >= band_lo);
            int ins_ok = (i - k >= 1) && (d + k <= d_hi);
            if (!del_ok && !ins_ok) break;
            if (del_ok && h == HAT(i, j - k) - gap_cost(sc->del_open1, sc->del_ext1, sc->del_open2, sc->del_ext2, k)) {
                EMIT(SVP_CIGAR_D, (uint32_t)k); res->n_del_ops++; res->n_del_bp += (uint32_t)k; j -= k; found = 1; break;
            }
            if (ins_ok && h == HAT(i - k, j) - gap_cost(sc->ins_open1, sc->ins_ext1, sc->ins_open2, sc->ins_ext2, k)) {
                EMIT(SVP_CIGAR_I, (uint32_t)k); res->n_ins_ops++; res->n_ins_bp += (uint32_t)k; i -= k; found = 1; break;
            }
        }
        if (!found) { status |= SVP_ST_SCORE_OVF; break; }   /* cannot happen for a consistent matrix */
    }
    /* rewind to the snapshot: the reported alignment starts at the lowest point of the walk */
    i = low_i; j = low_j; nops = low_nops; cur_op = low_op; cur_len = low_len;
    res->n_match = low_cnt[0]; res->n_mismatch = low_cnt[1]; res->n_ins_ops = low_cnt[2]; res->n_ins_bp = low_cnt[3];
    res->n_del_ops = low_cnt[4]; res->n_del_bp = low_cnt[5];
    res->score = best - low;
    PUSH_RUN();
    res->q_beg = i; res->q_end = best_i;
    res->t_beg = j; res->t_end = best_j;

    /* forward-order CIGAR + sv_dist */
    double dist = 0.0;
    for (size_t a = 0; a < nops; ++a) {
        uint32_t w = ops[nops - 1 - a];
        uint32_t op = w & 15u, len = w >> 4;
        if (op != SVP_CIGAR_M && (int)len >= sc->min_signal_size) {
            double s = (double)len;
            dist += (1.0 - exp(-pow(s / 30.0, 1.5))) * s;
        }
    }
    res->sv_dist = dist;
    if (nops > cap) {
        status |= SVP_ST_CIGAR_OVF;
        res->cigar_len = 0;
    } else {
        res->cigar_len = (uint32_t)nops;
        res->cigar_off = slot_off + cap - (uint32_t)nops;     /* right-aligned in the slot, like the GPU */
        for (size_t a = 0; a < nops; ++a) cigar_slot[cap - nops + a] = ops[nops - 1 - a];
    }
    res->status = status;
    free(ops); free(H); free(F1); free(F2);
#undef HAT
#undef PUSH_RUN
#undef EMIT
#undef SNAPSHOT
}

typedef struct {
    const svp_scoring* sc;
    const uint8_t* packed_seqs;
    const uint64_t* seq_off;
    const uint32_t* seq_len;
    uint32_t n_seqs;
    const svp_pair* pairs;
    uint32_t n_pairs;
    svp_result* results;
    uint32_t* cigar_buf;
    size_t cigar_words;
    volatile long next;      /* work counter, advanced with __sync_fetch_and_add */
} batch_ctx;

static void* batch_worker(void* arg) {
    batch_ctx* c = (batch_ctx*)arg;
    for (;;) {
        long p = __sync_fetch_and_add(&c->next, 1);
        if (p >= (long)c->n_pairs) break;
        const svp_pair* pr = &c->pairs[p];
        svp_result* r = &c->results[p];
        if (pr->q >= c->n_seqs || pr->t >= c->n_seqs || pr->band_w == 0 || (pr->band_w & 31u) ||
            (size_t)pr->cigar_off + pr->cigar_cap > c->cigar_words) {
            memset(r, 0, sizeof(*r)); r->status = SVP_ST_BAD_PAIR; continue;
        }
        const int m = (int)c->seq_len[pr->q], n = (int)c->seq_len[pr->t];
        uint8_t* q = (uint8_t*)malloc((size_t)m + 1);
        uint8_t* t = (uint8_t*)malloc((size_t)n + 1);
        const uint64_t oq = c->seq_off[pr->q] & ~(uint64_t)15, ot = c->seq_off[pr->t] & ~(uint64_t)15;
        const uint8_t* pq = c->packed_seqs + oq;
        const uint8_t* pt = c->packed_seqs + ot;
        if (pr->flags & SVP_PAIR_REVCOMP) for (int k = 0; k < m; ++k) q[k] = (uint8_t)(3 - base_at(pq, (uint32_t)(m - 1 - k)));
        else for (int k = 0; k < m; ++k) q[k] = (uint8_t)base_at(pq, (uint32_t)k);
        for (int k = 0; k < n; ++k) t[k] = (uint8_t)base_at(pt, (uint32_t)k);
        if (c->seq_off[pr->q] & SVP_SEQ_NMASK) {
            const uint8_t* mq = pq + SVP_SEQ_CODE_BYTES(m);
            for (int k = 0; k < m; ++k) if (mask_at(mq, (uint32_t)((pr->flags & SVP_PAIR_REVCOMP) ? m - 1 - k : k))) q[k] = 4;
        }
        if (c->seq_off[pr->t] & SVP_SEQ_NMASK) {
            const uint8_t* mt = pt + SVP_SEQ_CODE_BYTES(n);
            for (int k = 0; k < n; ++k) if (mask_at(mt, (uint32_t)k)) t[k] = 4;
        }
        align_one(c->sc, q, m, t, n, pr->band_lo, (int)pr->band_w, r, c->cigar_buf + pr->cigar_off, pr->cigar_off, pr->cigar_cap);
        free(q); free(t);
    }
    return NULL;
}

int svpo_align_batch(const svp_scoring* sc,
                     const uint8_t* packed_seqs, size_t packed_bytes,
                     const uint64_t* seq_off, const uint32_t* seq_len, uint32_t n_seqs,
                     const svp_pair* pairs, uint32_t n_pairs,
                     svp_result* results, uint32_t* cigar_buf, size_t cigar_words) {
    if (!sc || !packed_seqs || !seq_off || !seq_len || (!pairs && n_pairs) || !results) return SVP_ERR_ARG;
    for (uint32_t s = 0; s < n_seqs; ++s) {
        const uint64_t o = seq_off[s] & ~(uint64_t)15;
        const uint64_t need = (seq_off[s] & SVP_SEQ_NMASK) ? SVP_SEQ_CODE_BYTES(seq_len[s]) + ((uint64_t)seq_len[s] + 7) / 8 : ((uint64_t)seq_len[s] + 3) / 4;
        if (o + need > packed_bytes) return SVP_ERR_ARG;
    }
    batch_ctx c = { sc, packed_seqs, seq_off, seq_len, n_seqs, pairs, n_pairs, results, cigar_buf, cigar_words, 0 };
    int nt = g_threads;
    if (nt > (int)n_pairs) nt = (int)n_pairs;
    if (nt <= 1) { batch_worker(&c); return SVP_OK; }
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nt);
    int started = 0;
    for (int k = 0; k < nt; ++k) { if (pthread_create(&th[k], NULL, batch_worker, &c) == 0) started++; else break; }
    if (started == 0) batch_worker(&c);
    for (int k = 0; k < started; ++k) pthread_join(th[k], NULL);
    free(th);
    return SVP_OK;
}
