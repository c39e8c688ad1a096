"""Pair batches of the BASELINE.json workloads (configs 2-5) from the seeded generators in synth.py — shared by
bench.py, the parity tests and the multi-GPU sharding driver, so that every consumer aligns exactly the same pairs.

    config 2   10k regions x 30 reads x 8 kb, all-vs-all                      batch_config2()
    config 3   tandem-repeat regions, reads 10-50 kb, band |dlen| + 512      batch_config3()
    config 4   HG002-shape loci: all-vs-all + consensus -> reference window   batch_config4() / consensus_objects()
    config 5   trio, three config-4-shaped sets with a Pareto length tail     batch_config4(seed_base=40/50/60 000, skew=True)
"""
from __future__ import annotations

import numpy as np

from . import synth
from ._abi import PAIR_DTYPE, PAIR_REVCOMP, pack_codes
from .ava import cigar_cap_for, round_up
from .consensus_class import Consensus, ConsensusPadding


def assemble(seqs: list[np.ndarray], specs: list[tuple[int, int, int, int, int]], n_regions: int) -> dict:
    """Code arrays + (q, t, band_lo, band_w, flags) specs -> the buffers svp_align_batch() takes."""
    packed, off, lens = pack_codes(seqs)
    pairs = np.zeros(len(specs), dtype=PAIR_DTYPE)
    if specs:
        arr = np.array(specs, dtype=np.int64)
        pairs["q"], pairs["t"], pairs["band_lo"], pairs["band_w"], pairs["flags"] = arr[:, 0], arr[:, 1], arr[:, 2], arr[:, 3], arr[:, 4]
        caps = np.array([cigar_cap_for(int(lens[q]), int(lens[t])) for q, t in arr[:, :2]], dtype=np.int64)
        pairs["cigar_off"], pairs["cigar_cap"] = np.concatenate(([0], np.cumsum(caps)[:-1])), caps
        words = int(caps.sum())
        if words >= 1 << 32:
            raise ValueError("CIGAR slots of one batch exceed 2^32 words (svp_pair.cigar_off is 32-bit): split the batch")
    else:
        words = 1
    return {"packed": packed, "off": off, "lens": lens, "pairs": pairs, "cigar_words": words, "n_regions": n_regions,
            "seq_bytes": int(((lens.astype(np.int64) + 3) // 4).sum())}


def batch_config2(region_ids: list[int], slack: int = 300, granule: int = 128) -> dict:
    seqs, specs = [], []
    for rid in region_ids:
        reg = synth.region_config2(rid)
        base = len(seqs)
        seqs += reg.reads
        specs += [(base + q, base + t, lo, w, fl) for (q, t, lo, w, fl) in synth.region_pairs(reg, slack, granule)]
    return assemble(seqs, specs, len(region_ids))


def batch_config3(region_ids: list[int], n_reads: int = 20, scale: float = 1.0, slack: int = 256, granule: int = 32) -> dict:
    seqs, specs = [], []
    for rid in region_ids:
        reg = synth.region_config3(rid, n_reads=n_reads, scale=scale)
        base = len(seqs)
        seqs += reg.reads
        specs += [(base + q, base + t, lo, w, fl) for (q, t, lo, w, fl) in synth.region_pairs_config3(reg, slack, granule)]
    return assemble(seqs, specs, len(region_ids))


def consensus_window_specs(locus: synth.Locus, q_index: list[int], t_index: int, band_w: int = 4096, granule: int = 32
                           ) -> list[tuple[int, int, int, int, int]]:
    """Pair specs of a locus' padded consensuses vs its reference window, both strands, band around the expected
    diagonal (core start <-> candidate-region start), as consensus_align.align_consensuses_to_reference() builds them."""
    specs = []
    n = len(locus.ref_window)
    r_start = locus.core_interval[0] - locus.window_start
    for c, q in zip(locus.consensus, q_index):
        m = len(c["padded"])
        lo_c, hi_c = c["core"]
        w = min(round_up(band_w, granule), round_up(m + n - 1, granule))
        for rc, d in ((0, r_start - lo_c), (1, r_start - (m - hi_c))):
            specs.append((q, t_index, d - w // 2, w, PAIR_REVCOMP if rc else 0))
    return specs


def batch_config4(locus_ids: list[int], seed_base: int = 30_000, skew: bool = False, scale: float = 1.0, max_reads: int = 60,
                  slack: int = 300, granule: int = 128, band_w: int = 4096) -> dict:
    """All-vs-all of every locus' reads plus its consensus -> reference-window pairs (both strands)."""
    seqs, specs = [], []
    for lid in locus_ids:
        loc = synth.locus_config4(lid, seed_base=seed_base, skew=skew, scale=scale, max_reads=max_reads)
        base = len(seqs)
        seqs += loc.region.reads
        specs += [(base + q, base + t, lo, w, fl) for (q, t, lo, w, fl) in synth.region_pairs(loc.region, slack, granule)]
        t_index = len(seqs)
        seqs.append(loc.ref_window)
        q_index = []
        for c in loc.consensus:
            q_index.append(len(seqs))
            seqs.append(c["padded"])
        specs += consensus_window_specs(loc, q_index, t_index, band_w)
    return assemble(seqs, specs, len(locus_ids))


def batch_loci(items: list[tuple[int, int, bool]], slack: int = 300, granule: int = 128, band_w: int = 4096) -> dict:
    """batch_config4() over loci of several sets: items = (seed_base, locus id, skew). Config 5 = the three trio sets
    (seed bases 40 000 / 50 000 / 60 000, skew) in one pool."""
    seqs, specs, bounds = [], [], [0]
    for seed_base, lid, skew in items:
        loc = synth.locus_config4(lid, seed_base=seed_base, skew=skew)
        base = len(seqs)
        seqs += loc.region.reads
        specs += [(base + q, base + t, lo, w, fl) for (q, t, lo, w, fl) in synth.region_pairs(loc.region, slack, granule)]
        t_index = len(seqs)
        seqs.append(loc.ref_window)
        q_index = []
        for c in loc.consensus:
            q_index.append(len(seqs))
            seqs.append(c["padded"])
        specs += consensus_window_specs(loc, q_index, t_index, band_w)
        bounds.append(len(specs))
    out = assemble(seqs, specs, len(items))
    out["pair_bounds"] = np.array(bounds, dtype=np.int64)        # pairs of locus k: [pair_bounds[k], pair_bounds[k + 1])
    return out


def consensus_objects(locus: synth.Locus) -> dict[str, Consensus]:
    """The locus' padded consensuses as Consensus records (lower-case flanks, UPPER core), for the consensus_align chain."""
    out = {}
    for c in locus.consensus:
        s = synth.codes_to_str(c["padded"])
        lo, hi = c["core"]
        padded = s[:lo].lower() + s[lo:hi].upper() + s[hi:].lower()
        pad = ConsensusPadding(sequence=padded, readname_left="", readname_right="", padding_size_left=lo,
                               padding_size_right=len(s) - hi, consensus_interval_on_sequence_with_padding=(lo, hi))
        out[c["id"]] = Consensus(ID=c["id"], crIDs=[locus.region.region_id],
                                 original_regions=[(locus.chrom, locus.core_interval[0], locus.core_interval[1])],
                                 consensus_sequence=padded[lo:hi], consensus_padding=pad)
    return out
