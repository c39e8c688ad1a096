"""consensus -> reference alignment and its consumer chain, on the CUDA aligner — the alignment-path part of
svirlpool.localassembly.consensus_align (/root/reference/src/svirlpool/localassembly/consensus_align.py):

    minimap2 call (padded consensus vs WHOLE genome)        :1977-1987   -> align_consensuses_to_reference()
    write_partitioned_alignments (qstart/qend annotation)   :661-735     -> annotate_alignments() / write_alignments_tsv()
    _process_alignment_file_for_core_intervals              :851-1040    -> core_intervals_on_reference()
    parse_sv_signals_from_consensus per alignment           :1745-1790   -> consensus_sv_signals()

The reference seeds against the whole genome; here every padded consensus is aligned to a WINDOW of the reference
around the candidate regions it was assembled from (Consensus.original_regions +- padding + margin), on both strands,
inside a diagonal band around the expected position. Split alignments (inversions, large indels) come from the
iterative flank re-alignment of ava.map_jobs(). What is not reproduced: hits outside the window (translocations to
other chromosomes), MAPQ.
"""
from __future__ import annotations

import json
from pathlib import Path
from typing import Callable

from . import ava, consensus_align_lib, consensus_class, util
from .consensus_class import Consensus
from .datatypes import AlignedRecord, Alignment, MergedSVSignal

MAX_BAND = 32768         # first-round band around the expected diagonal (clusters cover up to 131 072 diagonals); unaligned flanks are
                         # then searched over the whole window (ava.map_jobs), so SVs larger than half this band still split properly


def reference_windows(cons: Consensus, chrom_lengths: dict[str, int], margin_frac: float = 0.10, min_margin: int = 500
                      ) -> list[tuple[str, int, int]]:
    """One window per chromosome the consensus' candidate regions lie on: their hull, widened on both sides by the larger
    padding of the consensus plus a margin (10 % of it, at least min_margin), clipped to the chromosome."""
    pad = 0
    if cons.consensus_padding is not None:
        pad = max(cons.consensus_padding.padding_size_left, cons.consensus_padding.padding_size_right)
    ext = pad + max(min_margin, int(margin_frac * pad))
    hull: dict[str, list[int]] = {}
    for chrom, start, end in cons.original_regions:
        h = hull.setdefault(chrom, [start, end])
        h[0], h[1] = min(h[0], start), max(h[1], end)
    return [(c, max(0, s - ext), min(chrom_lengths.get(c, e + ext), e + ext)) for c, (s, e) in hull.items()]


def align_consensuses_to_reference(engine, consensuses: dict[str, Consensus], fetch: Callable[[str, int, int], str],
                                   chrom_lengths: dict[str, int], band_w: int = 4096, min_score: int = ava.MIN_DP_SCORE,
                                   margin_frac: float = 0.10, min_margin: int = 500) -> list[AlignedRecord]:
    """Align every padded consensus to its reference window(s). `fetch(chrom, start, end)` returns reference bases.
    Records carry chromosome names and chromosome coordinates, sorted by (chromosome, position) like the
    coordinate-sorted BAM of the reference's call."""
    target_names, targets, offsets, jobs = [], [], [], []
    for cid, cons in consensuses.items():
        if cons.consensus_padding is None:
            raise ValueError(f"Consensus {cid} has no consensus_padding. This should not happen. Please check the input data.")
        seq = cons.consensus_padding.sequence
        core_lo, core_hi = cons.consensus_padding.consensus_interval_on_sequence_with_padding
        for chrom, w_start, w_end in reference_windows(cons, chrom_lengths, margin_frac, min_margin):
            regions = [r for r in cons.original_regions if r[0] == chrom]
            r_start, r_end = min(r[1] for r in regions), max(r[2] for r in regions)
            target_names.append(chrom)
            targets.append(fetch(chrom, w_start, w_end))
            offsets.append(w_start)
            # expected diagonal: core start <-> region start (forward); core end <-> region start on the reverse strand
            d_fwd = (r_start - w_start) - core_lo
            d_rev = (r_start - w_start) - (len(seq) - core_hi)
            jobs.append(ava.MapJob(name=cid, seq=seq, target=len(targets) - 1, diag_fwd=d_fwd, diag_rev=d_rev,
                                   band_w=min(MAX_BAND, max(band_w, 2 * (r_end - r_start)))))
    uniq = [f"{n}#{k}" for k, n in enumerate(target_names)]            # one entry per window, also on the same chromosome
    recs = ava.map_jobs(engine, jobs, uniq, targets, min_score=min_score)
    back = {u: (n, o) for u, n, o in zip(uniq, target_names, offsets)}
    for r in recs:
        chrom, off = back[r.reference_name]
        r.reference_name = chrom
        r.reference_start += off
    recs.sort(key=lambda r: (r.reference_name, r.reference_start, r.query_name))
    return recs


def annotate_alignments(records: list[AlignedRecord], samplename: str) -> list[tuple[str, int, int, Alignment]]:
    """(consensusID, qstart, qend, Alignment) per mapped record, with the annotations write_partitioned_alignments adds
    (interval on the consensus covered by the alignment and its midpoint)."""
    out = []
    for rec in records:
        if rec.is_unmapped:
            continue
        if rec.reference_end is None:
            raise ValueError(f"Alignment for consensusID {rec.query_name} has no reference_end.")
        qstart, qend = util.get_interval_on_read_in_region(a=rec, start=rec.reference_start, end=rec.reference_end)
        ann = {"qstart": qstart, "qend": qend, "qmid": (qstart + qend) // 2}
        out.append((rec.query_name, qstart, qend, Alignment.from_record(rec, samplename=samplename, annotations=ann)))
    return out


def write_alignments_tsv(path: Path | str, annotated: list[tuple[str, int, int, Alignment]]) -> None:
    """consensusID \\t qstart \\t qend \\t JSON(Alignment), sorted by (consensusID, qstart, qend) — the partition file
    format of consensus_align.py:661-735 / sort_partitioned_alignments :1875-1890."""
    with open(path, "w") as f:
        for cid, qs, qe, aln in sorted(annotated, key=lambda t: (t[0], t[1], t[2])):
            print(cid, qs, qe, json.dumps(aln.unstructure()), sep="\t", file=f)


def write_partitioned_alignments(input_bam, output_paths: dict[int, Path], index_consensusIDs: dict[int, set[str]],
                                 samplename: str) -> None:
    """The reference's entry point of this step (consensus_align.py:661-735), same signature and file format: every mapped
    consensus alignment -> one line `consensusID \t qstart \t qend \t JSON(datatypes.Alignment)` in the TSV of the partition
    that owns its consensusID (so all alignments of a consensus land in one file). `input_bam` is a BAM/SAM path (read by
    sam.read_alignments, no htslib) or an iterable of AlignedRecord straight from align_consensuses_to_reference().
    Raises ValueError for a consensusID outside index_consensusIDs or a record without reference_end, like the reference."""
    owner = {cid: k for k, ids in index_consensusIDs.items() for cid in ids}
    if isinstance(input_bam, (str, Path)):
        from . import sam
        records = list(sam.read_alignments(input_bam))
    else:
        records = list(input_bam)
    writers = {k: open(path, "w") for k, path in output_paths.items()}
    try:
        for rec in records:
            if rec.is_unmapped:
                continue
            if rec.query_name not in owner:
                raise ValueError(f"consensusID {rec.query_name} not found in index_consensusIDs. "
                                 "This should not happen. Please check the input data.")
        for cid, qstart, qend, aln in annotate_alignments(records, samplename):
            print(cid, qstart, qend, json.dumps(aln.unstructure()), sep="\t", file=writers[owner[cid]])
    finally:
        for w in writers.values():
            w.close()


def core_intervals_on_reference(consensuses: dict[str, Consensus], annotated: list[tuple[str, int, int, Alignment]]
                                ) -> dict[str, list[tuple[int, str, int, int, int, int]]]:
    """consensusID -> [(alignment index in (qstart, qend) order, chromosome, core start on ref, core end on ref,
    core start on consensus, core end on consensus)]; alignments that do not touch the core are skipped."""
    per: dict[str, list[tuple[int, int, Alignment]]] = {}
    for cid, qs, qe, aln in annotated:
        per.setdefault(cid, []).append((qs, qe, aln))
    out: dict[str, list[tuple[int, str, int, int, int, int]]] = {}
    for cid, lst in per.items():
        cons = consensuses[cid]
        lst.sort(key=lambda t: (t[0], t[1]))
        for idx, (_qs, _qe, aln) in enumerate(lst):
            ref_name, a, b = consensus_class.get_consensus_core_alignment_interval_on_reference(cons, aln.to_record())
            if a == b:
                continue
            lo, hi = cons.consensus_padding.consensus_interval_on_sequence_with_padding
            out.setdefault(cid, []).append((idx, ref_name, min(a, b), max(a, b), lo, hi))
    return out


def consensus_sv_signals(samplename: str, consensuses: dict[str, Consensus], annotated: list[tuple[str, int, int, Alignment]],
                         trf_intervals: dict[str, list[tuple[int, int, int]]] | None = None, min_signal_size: int = 12,
                         min_bnd_size: int = 50) -> dict[str, list[list[MergedSVSignal]]]:
    """consensusID -> per alignment (in (qstart, qend) order) the MergedSVSignals inside the core
    (thresholds of consensus_align.py:2199-2209)."""
    per: dict[str, list[tuple[int, int, Alignment]]] = {}
    for cid, qs, qe, aln in annotated:
        per.setdefault(cid, []).append((qs, qe, aln))
    out: dict[str, list[list[MergedSVSignal]]] = {}
    for cid, lst in per.items():
        cons = consensuses[cid]
        lst.sort(key=lambda t: (t[0], t[1]))
        core = cons.consensus_padding.consensus_interval_on_sequence_with_padding
        for _qs, _qe, aln in lst:
            rec = aln.to_record()
            trf = (trf_intervals or {}).get(rec.reference_name, [])
            out.setdefault(cid, []).append(consensus_align_lib.parse_sv_signals_from_consensus(
                samplename=samplename, reference_name=rec.reference_name, consensus_sequence=cons.consensus_sequence,
                pysam_alignment=rec, interval_core=core, trf_intervals=trf, min_signal_size=min_signal_size, min_bnd_size=min_bnd_size))
    return out
