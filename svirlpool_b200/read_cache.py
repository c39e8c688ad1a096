"""Read ingest for the consensus stage — host mirror of svirlpool.localassembly.read_cache.ReadSequenceCache
(/root/reference/src/svirlpool/localassembly/read_cache.py:58-264) on the htslib-free BAM reader of sam.py, with one
addition for the GPU path: `packed_for_cr()` hands the reads of a candidate region over as 2-bit packed arrays with
known strands, so the aligner is fed without FASTA temp files.

Semantics kept from the reference: one open BAM handle per batch of chromosomally adjacent candidate regions; every
region triggers one indexed fetch; secondary alignments are skipped; an alignment already seen for a read (same flag and
position) is not stored twice; sequences are kept in ORIGINAL read orientation (reverse-strand records are
reverse-complemented back); `advance(window_start)` evicts reads whose alignments end before the next region; a chromosome
switch flushes everything. Reads whose fetched record is hard-clipped are resolved through their SA tag (the reference
defers to consensus.get_full_read_sequences_of_alignments, consensus.py:382-580).
"""
from __future__ import annotations

import logging
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

from . import sam
from ._abi import pack_2bit
from .datatypes import CIGAR_H, AlignedRecord
from .util import reverse_complement

log = logging.getLogger(__name__)


def _full_read_length(aln: AlignedRecord) -> int:
    """pysam's infer_read_length(): query bases incl. soft AND hard clips."""
    return sum(n for op, n in (aln.cigartuples or ()) if op in (0, 1, 4, 5, 7, 8))


def _original_orientation(aln: AlignedRecord) -> str:
    return reverse_complement(aln.query_sequence) if aln.is_reverse else aln.query_sequence


@dataclass
class _CachedRead:
    """Everything the cache knows about one read on the current chromosome."""

    records: dict[tuple[int, int], AlignedRecord] = field(default_factory=dict)   # (flag, reference_start) -> the record kept for it
    right_end: int = -1                 # rightmost reference end over its records: the read stays until a window starts beyond it
    sequence: str | None = None         # full read, original orientation; None while unresolved


class ReadSequenceCache:
    def __init__(self, path_alignments: Path):
        self.path_alignments = Path(path_alignments)
        self._bam = sam.BamFile(self.path_alignments)
        self._reads: dict[str, _CachedRead] = {}
        self._chrom: str | None = None
        self.fetch_calls = 0
        self.cache_hits_reads = 0
        self.cache_misses_reads = 0

    # ------------------------------------------------------------------ public API
    def fetch_for_cr(self, cr) -> tuple[list[AlignedRecord], dict[str, str]]:
        """(alignments overlapping the candidate region, full sequences of their reads). `cr` needs .chr,
        .referenceStart, .referenceEnd."""
        if cr.chr != self._chrom:                  # the cache never spans chromosomes
            self._reads = {}
            self._chrom = cr.chr
        self.fetch_calls += 1
        region_records: list[AlignedRecord] = []
        unresolved: set[str] = set()               # reads without a sequence yet: first met in this fetch, or not resolvable so far
        for aln in self._bam.fetch(cr.chr, max(int(cr.referenceStart), 0), int(cr.referenceEnd)):
            if aln.is_secondary:
                continue
            read = self._reads.get(aln.query_name)
            if read is None:
                read = self._reads[aln.query_name] = _CachedRead()
            if read.sequence is None:
                unresolved.add(aln.query_name)
            # an alignment met again through an overlapping region is answered with the record already held
            kept = read.records.setdefault((int(aln.flag), int(aln.reference_start)), aln)
            region_records.append(kept)
            if kept is aln:
                read.right_end = max(read.right_end, aln.reference_end or 0)
        for name in unresolved:
            self._reads[name].sequence = self._full_sequence(name, list(self._reads[name].records.values()))
        names = {r.query_name for r in region_records}
        self.cache_misses_reads += len(unresolved)
        self.cache_hits_reads += len(names) - len(unresolved)
        return region_records, {n: self._reads[n].sequence for n in names if self._reads[n].sequence is not None}

    def packed_for_cr(self, cr) -> tuple[list[str], np.ndarray, np.ndarray, np.ndarray, dict[str, str]]:
        """The region's reads ready for svp_align_batch(): (names, packed 2-bit bytes, seq_off, seq_len, strand per read).
        The strand is that of the read's primary (else first) alignment in the region: known orientations halve the
        all-vs-all work (ava.ava_pair_specs)."""
        alns, seqs = self.fetch_for_cr(cr)
        names = sorted(seqs)
        strands: dict[str, str] = {}
        for a in alns:
            if a.query_name in seqs and (a.query_name not in strands or not a.is_supplementary):
                strands[a.query_name] = "-" if a.is_reverse else "+"
        packed, off, lens = pack_2bit([seqs[n] for n in names])
        return names, packed, off, lens, strands

    def advance(self, window_start) -> None:
        """Drop every read none of whose alignments reaches window_start (the start of the next container on this chromosome)."""
        self._reads = {name: read for name, read in self._reads.items() if read.right_end >= window_start}

    def cached_readnames(self) -> set[str]:
        """Names of the reads whose sequence is held right now."""
        return {name for name, read in self._reads.items() if read.sequence is not None}

    def current_size_reads(self) -> int:
        return sum(1 for read in self._reads.values() if read.sequence is not None)

    def current_size_bp(self) -> int:
        return sum(len(read.sequence) for read in self._reads.values() if read.sequence is not None)

    def close(self) -> None:
        self._bam.close()
        self._reads = {}

    def __enter__(self) -> "ReadSequenceCache":
        return self

    def __exit__(self, exc_type, exc, tb) -> None:
        self.close()

    # ------------------------------------------------------------------ internals
    def _full_sequence(self, name: str, records: list[AlignedRecord]) -> str | None:
        """The read's full sequence in original orientation: from a record of this fetch that carries it (no hard clip), else
        from the record its SA tag points to."""
        for aln in records:
            if aln.query_sequence and _full_read_length(aln) == len(aln.query_sequence):
                return _original_orientation(aln)
        for aln in records:                      # hard-clipped records only: the full SEQ sits on another record of the read
            full = self._find_full_record(aln)
            if full is not None:
                return _original_orientation(full)
        log.warning("ReadSequenceCache: no record with the full sequence of read %s", name)
        return None

    def _find_full_record(self, aln: AlignedRecord) -> AlignedRecord | None:
        """Follow the SA tag (rname,pos,strand,CIGAR,mapQ,NM;...) to the read's other records until one carries the
        full sequence (no hard clip)."""
        for entry in str(aln.tags.get("SA", "")).split(";"):
            parts = entry.split(",")
            if len(parts) < 4 or parts[0] not in self._bam.references:
                continue
            pos = int(parts[1]) - 1
            for cand in self._bam.fetch(parts[0], pos, pos + 1):
                clipped = any(op == CIGAR_H for op, _n in cand.cigartuples or ())
                if cand.query_name == aln.query_name and cand.reference_start == pos and cand.query_sequence and not clipped:
                    return cand
        return None
