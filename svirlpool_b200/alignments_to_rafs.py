"""CIGAR -> SV signals: host mirror of the three functions of
svirlpool.signalprocessing.alignments_to_rafs that sit on the hot path
(/root/reference/src/svirlpool/signalprocessing/alignments_to_rafs.py):
    get_indels                       :101-136
    get_start_end                    :139-168
    parse_SVsignals_from_alignment   :174-258
They take any record with pysam.AlignedSegment's attribute names (datatypes.AlignedRecord).
"""
from __future__ import annotations

import numpy as np

from . import util
from .datatypes import SVsignal

Interval = tuple[int, int]


def _ops(alignment) -> tuple[np.ndarray, np.ndarray]:
    ct = alignment.cigartuples
    t = np.fromiter((op for op, _ in ct), dtype=np.int64, count=len(ct))
    x = np.fromiter((n for _, n in ct), dtype=np.int64, count=len(ct))
    return t, x


def get_indels(alignment) -> tuple[list[Interval], list[Interval], list[Interval], list[Interval]]:
    """(deletions on ref, deletions on read, insertions on ref, insertions on read), CIGAR order."""
    t, x = _ops(alignment)
    rs, re_, fs, fe = util.get_starts_ends(t=t, x=x, is_reverse=alignment.is_reverse,
                                           reference_start=alignment.reference_start)
    d, i = t == 2, t == 1
    pairs = lambda a, b, m: list(zip(a[m].tolist(), b[m].tolist()))  # noqa: E731
    return pairs(fs, fe, d), pairs(rs, re_, d), pairs(fs, fe, i), pairs(rs, re_, i)


def get_start_end(alignment) -> tuple[int, int, int, int]:
    """(ref_start, ref_end, read_start, read_end); read_start is the clip (S or H) that precedes the
    alignment in ORIGINAL read orientation: the leading clip on forward, the trailing clip on reverse."""
    if alignment.is_unmapped:
        raise ValueError("Alignment is unmapped. Make sure unly mapped alignments are passed.")
    ct = alignment.cigartuples
    aligned = sum(n for op, n in ct if op in (0, 1, 7, 8))
    edge = ct[-1] if alignment.is_reverse else ct[0]
    read_start = edge[1] if edge[0] in (4, 5) else 0
    return alignment.reference_start, alignment.reference_end, read_start, read_start + aligned


def parse_SVsignals_from_alignment(alignment, ref_start: int, ref_end: int, read_start: int, read_end: int,
                                   min_signal_size: int, min_bnd_size: int) -> list[SVsignal]:
    """SV signals of one alignment, ordered [BNDL?, BNDR?, DELs..., INSs...] (CIGAR order inside each
    group). BNDL(3)/BNDR(4): first/last op is a clip (S or H) of at least min_bnd_size; DEL(1)/INS(0):
    every D/I block of at least min_signal_size. On the reverse strand the caller's read_start/read_end
    are swapped before use (:184-185)."""
    if alignment.is_reverse:
        read_start, read_end = read_end, read_start
    ct = alignment.cigartuples
    out: list[SVsignal] = []
    if ct[0][0] in (4, 5) and ct[0][1] >= min_bnd_size:
        out.append(SVsignal(int(ref_start), int(ref_start), int(read_start), int(read_start), int(ct[0][1]), 3))
    if ct[-1][0] in (4, 5) and ct[-1][1] >= min_bnd_size:
        out.append(SVsignal(int(ref_end), int(ref_end), int(read_end), int(read_end), int(ct[-1][1]), 4))
    del_ref, del_read, ins_ref, ins_read = get_indels(alignment)
    for (a, b), (ra, rb) in zip(del_ref, del_read):
        if abs(b - a) >= min_signal_size:
            out.append(SVsignal(int(a), int(b), int(ra), int(rb), int(abs(b - a)), 1))
    for (a, b), (ra, rb) in zip(ins_ref, ins_read):
        if abs(rb - ra) >= min_signal_size:
            out.append(SVsignal(int(a), int(b), int(ra), int(rb), int(abs(rb - ra)), 0))
    return out
