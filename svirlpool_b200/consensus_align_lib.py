"""SV signals of a consensus-to-reference alignment — host mirror of svirlpool.localassembly.consensus_align_lib
(/root/reference/src/svirlpool/localassembly/consensus_align_lib.py):

    ref_and_read_start_from_merging_sv_signals   :25-33
    alt_sequence_for_MergedSVSignal              :36-80
    assign_repeat_ids                            :95-109
    parse_sv_signals_from_consensus              :112-197

The alignment consumed here is a record produced by the CUDA aligner (consensus_align.align_consensuses_to_reference)
instead of a pysam.AlignedSegment read back from minimap2's BAM; the arithmetic downstream of the CIGAR is the
reference's.
"""
from __future__ import annotations

import logging

from . import consensus
from .datatypes import MergedSVSignal
from .util import reverse_complement

log = logging.getLogger(__name__)


def ref_and_read_start_from_merging_sv_signals(sv_signals: list[MergedSVSignal]) -> tuple[int, int]:
    first = sorted(sv_signals, key=lambda s: s.ref_start)[0]
    return (first.ref_start, first.read_start)


def alt_sequence_for_MergedSVSignal(merged_signal: MergedSVSignal, consensus_sequence: str, reverse: bool,
                                    interval_core: tuple[int, int]) -> str:
    """ALT allele of one signal, cut out of the CORE consensus sequence (read coordinates are on the padded
    sequence, hence the shift by the core start): the inserted bases for INS, one base at the break for BNDs,
    empty for DEL; reverse-complemented for reverse-strand alignments."""
    assert len(consensus_sequence) > 0, "consensus_sequence must have a length > 0"
    readstart = merged_signal.read_start - interval_core[0]
    readend = merged_signal.read_end - interval_core[0]
    if merged_signal.sv_type == 0:
        seq = consensus_sequence[readstart:readend]
        return reverse_complement(seq) if reverse else seq
    if merged_signal.sv_type in (3, 4):
        base = consensus_sequence[readstart]
        return reverse_complement(base) if reverse else base
    if merged_signal.sv_type == 1:
        return ""
    raise ValueError(f"merged signal type must be 0 (INS), 1 (DEL), 3 (BND) or 4 (BND). Got {merged_signal.sv_type}")


def _overlaps(a0: int, a1: int, b0: int, b1: int) -> bool:
    """intervaltree.Interval.overlaps for half-open intervals."""
    return a0 < b1 and b0 < a1


def assign_repeat_ids(sv_signals: list[MergedSVSignal], trf_intervals: list[tuple[int, int, int]]) -> None:
    """repeatIDs of every tandem-repeat interval (start, end, repeatID) a signal's reference interval overlaps."""
    for sv in sv_signals:
        sv.repeatIDs = list({rid for start, end, rid in trf_intervals if _overlaps(sv.ref_start, sv.ref_end, start, end)})


def parse_sv_signals_from_consensus(samplename: str, reference_name: str, consensus_sequence: str, pysam_alignment,
                                    interval_core: tuple[int, int] | None, trf_intervals: list[tuple[int, int, int]],
                                    min_signal_size: int = 5, min_bnd_size: int = 50) -> list[MergedSVSignal]:
    """Signals of one consensus alignment that lie inside the core interval, as MergedSVSignals with repeat ids and
    ALT sequences, sorted by their position on the consensus."""
    sv_signals = consensus.parse_ReadAlignmentSignals_from_alignment(
        samplename=samplename, alignment=pysam_alignment, min_signal_size=min_signal_size, min_bnd_size=min_bnd_size).SV_signals
    if interval_core is not None:
        dropped = [s for s in sv_signals if not (s.read_start >= interval_core[0] and s.read_end <= interval_core[1])]
        sv_signals = [s for s in sv_signals if s.read_start >= interval_core[0] and s.read_end <= interval_core[1]]
        if dropped:
            log.debug("parse_sv_signals_from_consensus: dropped %d signals outside of interval_core %s of %s",
                      len(dropped), interval_core, pysam_alignment.query_name)
    merged = [MergedSVSignal(ref_start=s.ref_start, ref_end=s.ref_end, read_start=s.read_start, read_end=s.read_end, size=s.size,
                             sv_type=s.sv_type, chr=reference_name, repeatIDs=[], original_alt_sequences=[], original_ref_sequences=[])
              for s in sv_signals]
    assign_repeat_ids(merged, trf_intervals)
    for m in merged:
        m.original_alt_sequences = [alt_sequence_for_MergedSVSignal(m, consensus_sequence, pysam_alignment.is_reverse, interval_core)]
        if m.sv_type >= 3 and len(m.get_alt_sequence()) == 0:
            raise ValueError(f"merged signal {m} has no alt_sequence. All breakends need to have an alt_sequence!")
    return sorted(merged, key=lambda s: s.read_start)
