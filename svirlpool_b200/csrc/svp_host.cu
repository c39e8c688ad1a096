// svp_host.cu — host-only entry points of libsvp_align.so (include/svp_align.h): 2-bit packing of reads and the pair list
// of a container's all-vs-all alignment, laid out the way svp_align_batch() takes them. They replace what the reference
// does by writing FASTA files for minimap2 (all_reads.fasta, consensus.py:1362-1369) and letting `minimap2 -x ava-ont -X`
// pick the pairs (util/util.py:181-183). No device is touched: these run (and are tested) without a GPU.
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>

#include "../../include/svp_align.h"

namespace {

inline int64_t round_up64(int64_t x, int64_t g) { return (x + g - 1) / g * g; }

struct CodeLut {
    uint8_t v[256];
    CodeLut() {
        memset(v, 4, sizeof(v));                 // 4: a letter outside ACGT (N, IUPAC) — code 0 plus a bit in the sequence's mask (SVP_SEQ_NMASK)
        v[(int)'A'] = v[(int)'a'] = 0; v[(int)'C'] = v[(int)'c'] = 1; v[(int)'G'] = v[(int)'g'] = 2; v[(int)'T'] = v[(int)'t'] = 3;
    }
};
const CodeLut g_lut;

}  // namespace

extern "C" size_t svp_pack_bytes(const uint32_t* seq_len, uint32_t n_seqs) {
    size_t total = 16;                           // readable slack behind the last sequence
    if (!seq_len) return total;
    for (uint32_t k = 0; k < n_seqs; ++k) total += (size_t)SVP_SEQ_CODE_BYTES(seq_len[k]) + (((size_t)seq_len[k] + 7) / 8 + 15) / 16 * 16;   // codes + a mask
    return total;
}

extern "C" int svp_pack_2bit(const char* const* seqs, const uint32_t* seq_len, uint32_t n_seqs, uint8_t* packed, size_t packed_bytes,
                             uint64_t* seq_off) {
    if ((!seqs || !seq_len || !seq_off) && n_seqs) return SVP_ERR_ARG;
    if (!packed || packed_bytes < svp_pack_bytes(seq_len, n_seqs)) return SVP_ERR_ARG;
    size_t pos = 0;
    for (uint32_t k = 0; k < n_seqs; ++k) {
        const uint8_t* s = reinterpret_cast<const uint8_t*>(seqs[k]);
        if (!s && seq_len[k]) return SVP_ERR_ARG;
        const size_t len = seq_len[k], bytes = (len + 3) / 4, slot = (bytes + 15) / 16 * 16;
        seq_off[k] = pos;
        uint8_t* out = packed + pos;
        const size_t whole = len / 4;
        unsigned other = 0;                      // any letter outside ACGT
        for (size_t b = 0; b < whole; ++b) {
            const uint8_t* c = s + 4 * b;
            const unsigned c0 = g_lut.v[c[0]], c1 = g_lut.v[c[1]], c2 = g_lut.v[c[2]], c3 = g_lut.v[c[3]];
            other |= c0 | c1 | c2 | c3;
            out[b] = (uint8_t)((c0 & 3) | ((c1 & 3) << 2) | ((c2 & 3) << 4) | ((c3 & 3) << 6));
        }
        if (len & 3) {
            uint8_t x = 0;
            for (size_t r = 0; r < (len & 3); ++r) { const unsigned cr = g_lut.v[s[4 * whole + r]]; other |= cr; x |= (uint8_t)((cr & 3) << (2 * r)); }
            out[whole] = x;
        }
        memset(out + bytes, 0, slot - bytes);
        pos += slot;
        if (other & 4) {                         // the mask follows the codes: bit k % 8 of byte k / 8 = base k is not A, C, G or T
            const size_t mbytes = (len + 7) / 8, mslot = (mbytes + 15) / 16 * 16;
            uint8_t* mk = packed + pos;
            memset(mk, 0, mslot);
            for (size_t b = 0; b < len; ++b) if (g_lut.v[s[b]] & 4) mk[b >> 3] |= (uint8_t)(1u << (b & 7));
            seq_off[k] |= SVP_SEQ_NMASK;
            pos += mslot;
        }
    }
    memset(packed + pos, 0, packed_bytes - pos);
    return SVP_OK;
}

// band covering the whole m x n matrix / a band around the line from diagonal d_start to n - m (ava.full_band, ava.offset_band)
static void band_for(int64_t m, int64_t n, int64_t d_start, int64_t slack, int64_t granule, bool full, int32_t* lo_out, uint32_t* w_out) {
    const int64_t flo = -(m - 1), fw = round_up64(m + n - 1, granule);
    int64_t lo = flo, w = fw;
    if (!full) {
        const int64_t d_end = n - m;
        const int64_t l = std::min(d_start, d_end) - slack, h = std::max(d_start, d_end) + slack;
        const int64_t bw = round_up64(h - l + 1, granule);
        if (bw < fw) { lo = l; w = bw; }
    }
    *lo_out = (int32_t)lo; *w_out = (uint32_t)w;
}

extern "C" int svp_ava_pairs(const uint32_t* seq_len, const uint32_t* name_order, const int8_t* strand, const int32_t* offsets,
                             uint32_t seq_base, uint32_t n_reads, int32_t slack, int32_t granule, int full,
                             svp_pair* pairs, uint32_t pairs_cap, uint32_t* n_pairs, uint64_t* cigar_words) {
    if (!n_pairs || !cigar_words || (!seq_len && n_reads) || granule <= 0 || (granule % 32) != 0 || slack < 0) return SVP_ERR_ARG;
    uint32_t count = 0;
    uint64_t pos = *cigar_words;
    for (uint32_t a = 0; a < n_reads; ++a) {
        for (uint32_t b = a + 1; b < n_reads; ++b) {
            // every unordered pair once, query = the read that sorts first by name (minimap2 -X keeps query < target)
            const uint32_t q = name_order ? name_order[a] : a, t = name_order ? name_order[b] : b;
            if (q >= n_reads || t >= n_reads) return SVP_ERR_ARG;
            const int64_t m = seq_len[seq_base + q], n = seq_len[seq_base + t];
            if (m <= 0 || n <= 0) return SVP_ERR_ARG;
            // unknown strands: both orientations are aligned; known: the one that brings both reads onto the same strand
            const bool known = strand && strand[q] != 0 && strand[t] != 0;
            const int first_rc = known ? (strand[q] != strand[t] ? 1 : 0) : 0, last_rc = known ? first_rc : 1;
            for (int rc = first_rc; rc <= last_rc; ++rc) {
                const int64_t d0 = (offsets && !rc) ? (int64_t)offsets[q] - (int64_t)offsets[t] : 0;
                if (count < pairs_cap && pairs) {
                    svp_pair& p = pairs[count];
                    memset(&p, 0, sizeof(p));
                    p.q = seq_base + q; p.t = seq_base + t;
                    band_for(m, n, d0, slack, granule, full != 0, &p.band_lo, &p.band_w);
                    p.flags = rc ? SVP_PAIR_REVCOMP : 0u;
                    const uint64_t cap = std::min<uint64_t>((uint64_t)(m + n), (uint64_t)(m + n) / 3 + 64);   // ava.cigar_cap_for
                    if (pos + cap >= ((uint64_t)1 << 32)) return SVP_ERR_ARG;   // svp_pair.cigar_off is 32-bit: split the batch
                    p.cigar_off = (uint32_t)pos; p.cigar_cap = (uint32_t)cap;
                    pos += cap;
                }
                ++count;
            }
        }
    }
    *n_pairs = count;
    if (count > pairs_cap || (!pairs && count)) return SVP_ERR_NOMEM;       // *n_pairs tells the caller how many slots to provide
    *cigar_words = pos;
    return SVP_OK;
}
