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


class ReadSequenceCache:
    def __init__(self, path_alignments: Path):
        self.path_alignments = Path(path_alignments)
        self._samfile = sam.BamFile(self.path_alignments)
        self._seqs: dict[str, str] = {}                       # readname -> full sequence, original read orientation
        self._alns: dict[str, list[AlignedRecord]] = {}
        self._max_ref_end: dict[str, int] = {}
        self._aln_keys: dict[str, set[tuple[int, int]]] = {}
        self._current_chr: str | None = None
        self.fetch_calls = 0
        self.cache_hits_reads = 0
        self.cache_misses_reads = 0

    # ------------------------------------------------------------------ public API
    def fetch_for_cr(self, cr) -> tuple[list[AlignedRecord], dict[str, str]]:
        """(alignments overlapping the candidate region, full sequences of their reads). `cr` needs .chr,
        .referenceStart, .referenceEnd."""
        if self._current_chr != cr.chr:
            self._flush()
            self._current_chr = cr.chr
        cr_alns: list[AlignedRecord] = []
        new_readnames: set[str] = set()
        new_alns: list[AlignedRecord] = []
        self.fetch_calls += 1
        for aln in self._samfile.fetch(cr.chr, max(int(cr.referenceStart), 0), int(cr.referenceEnd)):
            if aln.is_secondary:
                continue
            qname = aln.query_name
            key = (int(aln.flag), int(aln.reference_start))
            seen = self._aln_keys.setdefault(qname, set())
            if key in seen:
                cr_alns.append(_pick_existing_aln(self._alns[qname], key) or aln)
                continue
            seen.add(key)
            self._alns.setdefault(qname, []).append(aln)
            ref_end = aln.reference_end if aln.reference_end is not None else 0
            if ref_end > self._max_ref_end.get(qname, -1):
                self._max_ref_end[qname] = ref_end
            cr_alns.append(aln)
            if qname not in self._seqs:
                new_readnames.add(qname)
                new_alns.append(aln)
        if new_readnames:
            self.cache_misses_reads += len(new_readnames)
            self._resolve_new_read_sequences(new_alns, new_readnames)
        cr_readnames = {a.query_name for a in cr_alns}
        self.cache_hits_reads += len(cr_readnames - new_readnames)
        return cr_alns, {rn: self._seqs[rn] for rn in cr_readnames if rn in self._seqs}

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
        for rn in [rn for rn, end in self._max_ref_end.items() if end < window_start]:
            self._seqs.pop(rn, None)
            self._alns.pop(rn, None)
            self._max_ref_end.pop(rn, None)
            self._aln_keys.pop(rn, None)

    def current_size_reads(self) -> int:
        return len(self._seqs)

    def current_size_bp(self) -> int:
        return sum(len(s) for s in self._seqs.values())

    def close(self) -> None:
        self._samfile.close()
        self._flush()

    def __enter__(self) -> "ReadSequenceCache":
        return self

    def __exit__(self, exc_type, exc, tb) -> None:
        self.close()

    # ------------------------------------------------------------------ internals
    def _flush(self) -> None:
        self._seqs.clear()
        self._alns.clear()
        self._max_ref_end.clear()
        self._aln_keys.clear()

    def _resolve_new_read_sequences(self, new_alns: list[AlignedRecord], new_readnames: set[str]) -> None:
        deferred = []
        for aln in new_alns:
            if aln.query_name in self._seqs:
                continue
            if aln.query_sequence and _full_read_length(aln) == len(aln.query_sequence):
                self._seqs[aln.query_name] = _original_orientation(aln)
            else:
                deferred.append(aln)
        for aln in deferred:                     # hard-clipped record: the full SEQ sits on another record of the read
            if aln.query_name in self._seqs:
                continue
            full = self._find_full_record(aln)
            if full is not None:
                self._seqs[aln.query_name] = _original_orientation(full)
            else:
                log.warning("ReadSequenceCache: no record with the full sequence of read %s", aln.query_name)

    def _find_full_record(self, aln: AlignedRecord) -> AlignedRecord | None:
        """Follow the SA tag (rname,pos,strand,CIGAR,mapQ,NM;...) to the read's other records until one carries the
        full sequence (no hard clip)."""
        for entry in str(aln.tags.get("SA", "")).split(";"):
            parts = entry.split(",")
            if len(parts) < 4 or parts[0] not in self._samfile.references:
                continue
            pos = int(parts[1]) - 1
            for cand in self._samfile.fetch(parts[0], pos, pos + 1):
                clipped = any(op == CIGAR_H for op, _n in cand.cigartuples or ())
                if cand.query_name == aln.query_name and cand.reference_start == pos and cand.query_sequence and not clipped:
                    return cand
        return None


def _pick_existing_aln(alns: list[AlignedRecord], key: tuple[int, int]) -> AlignedRecord | None:
    flag, ref_start = key
    for a in alns:
        if int(a.flag) == flag and int(a.reference_start) == ref_start:
            return a
    return None
