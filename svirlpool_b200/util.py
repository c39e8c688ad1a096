"""CIGAR block coordinates — host mirror of svirlpool.util.util.get_starts_ends
(/root/reference/src/svirlpool/util/util.py:1336-1361) plus small sequence helpers.
"""
from __future__ import annotations

import numpy as np

_READ_OPS = np.array([1, 1, 0, 0, 0, 0, 0, 1, 1], dtype=np.int64)        # M I = X consume aligned read bases
_READLEN_OPS = np.array([1, 1, 0, 0, 1, 1, 0, 1, 1], dtype=np.int64)     # + S and H: full read length
_REF_OPS = np.array([1, 0, 1, 1, 0, 0, 0, 1, 1], dtype=np.int64)         # M D N = X consume reference

_COMP = str.maketrans("ACGTNacgtn", "TGCANtgcan")


def get_starts_ends(t: np.ndarray, x: np.ndarray, reference_start: int, is_reverse: bool
                    ) -> tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    """Per-CIGAR-block (read_starts, read_ends, ref_starts, ref_ends).

    t: op codes, x: lengths. Read coordinates count from the start of the ORIGINAL read
    including hard clips; for reverse-strand alignments they are mirrored so that they are
    in original read orientation (util.py:1353-1359). Reference coordinates are unaffected.
    """
    t = np.asarray(t, dtype=np.int64)
    x = np.asarray(x, dtype=np.int64)
    lead = int(x[0]) if t[0] in (4, 5) else 0
    read_len = int(np.dot(x, _READLEN_OPS[t]))
    read_ends = np.cumsum(x * _READ_OPS[t]) + lead
    read_starts = np.concatenate(([lead], read_ends[:-1]))
    ref_ends = np.cumsum(x * _REF_OPS[t]) + reference_start
    ref_starts = np.concatenate(([reference_start], ref_ends[:-1]))
    if is_reverse:
        return read_len - read_ends, read_len - read_starts, ref_starts, ref_ends
    return read_starts, read_ends, ref_starts, ref_ends


def reverse_complement(seq: str) -> str:
    return seq.translate(_COMP)[::-1]


def read_fasta(path) -> dict[str, str]:
    """Minimal FASTA reader (header = first word), for the file seam of align_reads_with_minimap."""
    seqs: dict[str, list[str]] = {}
    name = None
    with open(path) as f:
        for line in f:
            line = line.strip()
            if not line:
                continue
            if line[0] == ">":
                name = line[1:].split()[0]
                seqs[name] = []
            elif name is not None:
                seqs[name].append(line)
    return {k: "".join(v) for k, v in seqs.items()}


def read_fasta_records(path):
    """(id, description, sequence) per FASTA record with Bio.SeqIO's conventions (what svirltile._add_consensus_sequences_to_db
    stores, localassembly/svirltile.py:78-110): the title is the header line without '>' and trailing white space, id = its first
    word ("" for an empty title), description = the WHOLE title (it repeats the id); the sequence is the record's lines joined,
    blanks and carriage returns removed."""
    title, chunks = None, []
    with open(path) as f:
        for line in f:
            if line.startswith(">"):
                if title is not None:
                    yield (title.split(None, 1)[0] if title.strip() else ""), title, "".join(chunks).replace(" ", "").replace("\r", "")
                title, chunks = line[1:].rstrip(), []
            elif title is not None:
                chunks.append(line.rstrip("\n"))
    if title is not None:
        yield (title.split(None, 1)[0] if title.strip() else ""), title, "".join(chunks).replace(" ", "").replace("\r", "")


def align_reads_with_minimap(reference, reads, bamout, tech: str = "map-ont", threads: int = 3, aln_args: str = "",
                             logfile=None, timeout: float | None = None, engine=None) -> None:
    """Drop-in for svirlpool.util.util.align_reads_with_minimap (util/util.py:162-238): FASTA path(s) in,
    coordinate-sorted BAM out, computed by the CUDA aligner instead of `minimap2 | samtools sort`.

    tech "ava-ont": all-vs-all of the reads (`reference` and `reads` are the same FASTA at the reference's
    AVA call site, consensus.py:1378-1386); otherwise every read is mapped against every sequence of `reference` and keeps
    the records on the sequence of its best primary alignment (primary + supplementary; what `--secondary=no` leaves of a
    multi-sequence reference, consensus.py:2089-2096). Of `aln_args` only `-Y` has a counterpart (supplementary records
    soft-clipped with the full read sequence instead of hard-clipped, consensus.py:851-857); `--sam-hit-only` and
    `--secondary=no` are what the engine does anyway, seeding options (`-U`, `-H`, `-z`, `-r`) have none. `threads` and
    `logfile` are accepted for signature compatibility.

    `timeout` (seconds) is honoured the way the reference's callers rely on (util/util.py:205-226 kills the child and raises
    TimeoutError; consensus.py:1387-1415 catches it and retries with half the reads): the work of the call is estimated from the
    planner BEFORE anything is launched (aligner.estimate_seconds: spec cells at a conservative rate) and TimeoutError is raised
    when it exceeds the budget — a pathological container (400x coverage, alignments_to_rafs.py:377) is refused, not run; a call
    that nevertheless overruns raises TimeoutError afterwards and leaves no BAM. Other failures raise aligner.AlignerError, a
    subprocess.CalledProcessError, which the reference's callers already catch. `engine`: an Aligner to use; default: the
    process-wide handle of aligner.shared_aligner() (bounded arena, created once).
    Like the reference (`samtools index`, util.py:232-238) a `.bai` is written next to the BAM."""
    import time
    from . import ava, sam
    from .aligner import estimate_seconds, shared_aligner
    t_start = time.monotonic()
    paths = reads if isinstance(reads, list) else [reads]
    queries: dict[str, str] = {}
    for p in paths:
        queries.update(read_fasta(p))
    refs = read_fasta(reference)
    if engine is None:
        engine = shared_aligner("ava-ont" if tech == "ava-ont" else "map-ont")
    if timeout is not None:
        names = list(queries)
        if tech == "ava-ont":
            first = ava.PairBatch.build_ava(names, [queries[n] for n in names], None, None, 300, 32, False)[0] if len(names) > 1 else None
        else:
            tn = list(refs)
            specs = [(len(tn) + qi, ti, *ava.full_band(len(queries[q]), len(refs[t])), fl)
                     for qi, q in enumerate(names) for ti, t in enumerate(tn) for fl in (0, ava.PAIR_REVCOMP)]
            first = ava.PairBatch.build(tn + names, [refs[t] for t in tn] + [queries[q] for q in names], specs) if specs else None
        need = estimate_seconds(engine.scoring, first) if first is not None else 0.0
        if need > timeout:
            raise TimeoutError(f"align_reads_with_minimap: estimated {need:.1f} s of alignment for {len(queries)} reads exceeds the timeout of {timeout} s")
    if tech == "ava-ont":
        records = ava.ava_align(engine, queries)
        lengths = {n: len(s) for n, s in queries.items()}
    else:
        records, _chosen = ava.map_queries_to_best_target(engine, queries, refs)
        lengths = {n: len(s) for n, s in refs.items()}
    if timeout is not None and time.monotonic() - t_start > timeout:
        raise TimeoutError(f"align_reads_with_minimap: exceeded the timeout of {timeout} s")
    soft_supplementary = "-Y" in aln_args.split()
    for r in records:
        seq = queries[r.query_name]
        r.query_sequence = reverse_complement(seq) if r.is_reverse else seq
        if r.is_supplementary and soft_supplementary:
            r.cigartuples = [(4 if op == 5 else op, n) for op, n in r.cigartuples]
        elif r.is_supplementary:                    # hard-clipped records carry only the aligned part
            a = r.cigartuples[0][1] if r.cigartuples[0][0] == 5 else 0
            r.query_sequence = r.query_sequence[a:a + r.query_alignment_length]
    cmd = f"svirlpool_b200 -x {tech} {aln_args}".strip()
    sam.write_bam(bamout, sam.sort_records(records, lengths), lengths, command=cmd, index=True)     # + <bamout>.bai


# ---------------------------------------------------------------------------------------------
# reference position <-> read position through a CIGAR (consensus_align consumer chain)
# ---------------------------------------------------------------------------------------------
import enum
import logging

log = logging.getLogger(__name__)

_MATCH_OPS = (0, 7, 8)


class Direction(enum.Enum):
    """util/util.py:1329-1333."""
    NONE = 0
    LEFT = 1
    RIGHT = 2
    BOTH = 3


def _cigar_arrays(alignment) -> tuple[np.ndarray, np.ndarray]:
    ct = alignment.cigartuples
    t = np.fromiter((op for op, _ in ct), dtype=np.int64, count=len(ct))
    x = np.fromiter((n for _, n in ct), dtype=np.int64, count=len(ct))
    return t, x


def _walk_to_match(t: np.ndarray, block: int, step: int) -> int:
    """Next block in direction `step` (+1 / -1) whose op is M/=/X; stops at the CIGAR's ends."""
    if step > 0:
        while t[block] not in _MATCH_OPS and block < len(t) - 1:
            block += 1
    else:
        while t[block] not in _MATCH_OPS and block > 0:
            block -= 1
    return block


def _swap_for_strand(direction: Direction, is_reverse: bool) -> Direction:
    """Read RIGHT (increasing read coordinate) is reference LEFT for a reverse-strand alignment."""
    if not is_reverse:
        return direction
    if direction == Direction.RIGHT:
        return Direction.LEFT
    if direction == Direction.LEFT:
        return Direction.RIGHT
    return direction


def get_ref_pitx_on_read(alignment, position: int, direction: Direction, buffer_clipped_length: int = 0
                         ) -> tuple[int, int, np.ndarray, np.ndarray]:
    """Read position (ORIGINAL read orientation, hard clips counted) that corresponds to reference `position`,
    the CIGAR block it falls into, and the op/length arrays — util/util.py:1365-1498, same conventions:
    positions left/right of the aligned reference span snap to the clip boundary (optionally reaching
    `buffer_clipped_length` into the clip); on an insertion or deletion block the position moves to the next
    match block in `direction` (given in READ orientation)."""
    is_reverse = alignment.is_reverse
    ref_start, ref_end = alignment.reference_start, alignment.reference_end
    t, x = _cigar_arrays(alignment)
    aln_length = int(np.dot(x, _READLEN_OPS[t]))
    left_clipped = int(x[0]) if t[0] in (4, 5) else 0
    right_clipped = int(x[-1]) if t[-1] in (4, 5) else 0
    if position < ref_start:
        buffer = max(0, left_clipped - buffer_clipped_length)
        return (aln_length - buffer, -1, t, x) if is_reverse else (buffer, 0, t, x)
    if position >= ref_end:
        last_ref_block = int(np.flatnonzero(_REF_OPS[t] == 1)[-1])
        buffer = max(0, right_clipped - (buffer_clipped_length if position > ref_end else 0))
        return (buffer, last_ref_block, t, x) if is_reverse else (aln_length - buffer, last_ref_block, t, x)
    read_starts, read_ends, ref_starts, ref_ends = get_starts_ends(t, x, ref_start, is_reverse)
    walk = _swap_for_strand(direction, is_reverse)
    if direction == Direction.RIGHT:
        block = int(np.flatnonzero((ref_starts <= position) & (position < ref_ends))[0])
    elif direction == Direction.LEFT:
        block = int(np.flatnonzero((ref_starts < position) & (position <= ref_ends))[0])
    else:
        block = int(np.flatnonzero((ref_starts <= position) & (position <= ref_ends))[0])
    final_pos = -1
    if t[block] in _MATCH_OPS:
        final_pos = read_starts[block] + ((ref_ends[block] - position) if is_reverse else (position - ref_starts[block]))
    if t[block] == 1:                      # on an insertion (only reachable with the closed-interval lookup)
        if walk == Direction.RIGHT:
            block = _walk_to_match(t, block, +1)
            final_pos = read_starts[block]
        else:
            block = _walk_to_match(t, block, -1)
            final_pos = read_ends[block]
    if t[block] in (2, 3):                 # on a deletion / skipped region
        if walk == Direction.LEFT:
            block = _walk_to_match(t, block, -1)
            final_pos = read_starts[block] if is_reverse else read_ends[block]
        if walk == Direction.RIGHT:
            block = _walk_to_match(t, block, +1)
            final_pos = read_ends[block] if is_reverse else read_starts[block]
        if direction in (Direction.NONE, Direction.BOTH):   # the nearer of the two flanking match blocks
            bl, br = _walk_to_match(t, block, -1), _walk_to_match(t, block, +1)
            if position - ref_ends[bl] < ref_starts[br] - position:
                final_pos, block = (read_starts[bl] if is_reverse else read_ends[bl]), bl
            else:
                final_pos, block = (read_ends[br] if is_reverse else read_starts[br]), br
    return int(final_pos), int(block), t, x


def get_ref_position_on_read(alignment, position: int, direction: Direction, buffer_clipped_length: int = 0) -> int:
    """util/util.py:1501-1510."""
    return get_ref_pitx_on_read(alignment, position, direction, buffer_clipped_length)[0]


def is_ref_position_on_aligned_read(alignment, position: int) -> bool:
    """util/util.py:1513-1519."""
    return alignment.reference_start <= position < alignment.reference_end


def get_read_pitx_on_ref(alignment, position: int, direction: Direction) -> tuple[int, int, np.ndarray, np.ndarray]:
    """Reference position that corresponds to read `position` (ORIGINAL read orientation) — util/util.py:1524-1648.
    Positions in the clipped part before / after the aligned read span return the alignment's reference start / end
    (swapped for reverse-strand alignments). Only LEFT and RIGHT are supported; anything else is treated as RIGHT
    (with a warning), as the reference does."""
    is_reverse = alignment.is_reverse
    ref_start, ref_end = alignment.reference_start, alignment.reference_end
    t, x = _cigar_arrays(alignment)
    read_start = 0
    if is_reverse:
        if t[-1] in (4, 5):
            read_start = int(x[-1])
    elif t[0] in (4, 5):
        read_start = int(x[0])
    read_end = read_start + int(np.dot(x, _READ_OPS[t]))
    read_blocks = np.flatnonzero(_READ_OPS[t] == 1)
    if position <= read_start:
        return (int(ref_end), int(read_blocks[-1]), t, x) if is_reverse else (int(ref_start), 0, t, x)
    if position >= read_end:
        return (int(ref_start), 0, t, x) if is_reverse else (int(ref_end), int(read_blocks[-1]), t, x)
    read_starts, read_ends, ref_starts, ref_ends = get_starts_ends(t, x, ref_start, is_reverse)
    if direction not in (Direction.RIGHT, Direction.LEFT):
        log.warning("Other directions than LEFT or RIGHT are not supported. Using RIGHT as default.")
        direction = Direction.RIGHT
    walk = _swap_for_strand(direction, is_reverse)
    if direction == Direction.RIGHT:
        block = int(np.flatnonzero((read_starts <= position) & (position < read_ends))[0])
    else:
        block = int(np.flatnonzero((read_starts < position) & (position <= read_ends))[0])
    final_pos = -1
    if t[block] in _MATCH_OPS:
        if is_reverse:
            final_pos = ref_ends[block] - (position - read_starts[block])
        else:
            final_pos = ref_starts[block] + position - read_starts[block]
    if t[block] == 2:
        if walk == Direction.RIGHT:
            block = _walk_to_match(t, block, +1)
            final_pos = ref_starts[block]
        else:
            block = _walk_to_match(t, block, -1)
            final_pos = ref_ends[block]
    if t[block] == 1:
        if walk == Direction.LEFT:
            if block > 0:                  # (an insertion in the very first block keeps -1, as in the reference)
                block = _walk_to_match(t, block, -1)
                final_pos = ref_starts[block] if is_reverse else ref_ends[block]
        elif walk == Direction.RIGHT:
            block = _walk_to_match(t, block, +1)
            final_pos = ref_ends[block] if is_reverse else ref_starts[block]
    return int(final_pos), int(block), t, x


def get_read_position_on_ref(alignment, position: int, direction: Direction) -> int:
    """util/util.py:1651-1656."""
    return get_read_pitx_on_ref(alignment, position, direction)[0]


def get_interval_on_read_in_region(a, start: int, end: int, buffer_clipped_length: int = 0) -> tuple[int, int]:
    """Read interval covering reference [start, end) — util/util.py:716-736."""
    start, end = (end, start) if start > end else (start, end)
    istart = get_ref_position_on_read(a, start, Direction.RIGHT, buffer_clipped_length)
    iend = get_ref_position_on_read(a, end, Direction.RIGHT, buffer_clipped_length)
    return (iend, istart) if a.is_reverse else (istart, iend)


def get_interval_on_ref_in_region(a, start: int, end: int) -> tuple[int, int]:
    """Reference interval covering read [start, end) — util/util.py:739-749."""
    start, end = (end, start) if a.is_reverse else (start, end)
    istart = get_read_position_on_ref(a, start, Direction.RIGHT)
    iend = get_read_position_on_ref(a, end, Direction.LEFT)
    return min(iend, istart), max(iend, istart)
