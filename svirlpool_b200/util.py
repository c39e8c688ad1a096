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


def align_reads_with_minimap(reference, reads, bamout, tech: str = "map-ont", threads: int = 3, aln_args: str = "",
                             logfile=None, timeout: float | None = None, engine=None) -> None:
    """Drop-in for svirlpool.util.util.align_reads_with_minimap (util/util.py:162-238): FASTA path(s) in,
    coordinate-sorted BAM out, computed by the CUDA aligner instead of `minimap2 | samtools sort`.

    tech "ava-ont": all-vs-all of the reads (`reference` and `reads` are the same FASTA at the reference's
    AVA call site, consensus.py:1378-1386); otherwise every read is mapped to every sequence of `reference`
    (primary + supplementary records, `--secondary=no`). `threads`, `aln_args`, `logfile` and `timeout` are
    accepted for signature compatibility (seeding options have no counterpart; the call is bounded by the
    kernels, not by a child process). Failures raise aligner.AlignerError, a subprocess.CalledProcessError,
    which the reference's callers already catch. `engine`: an Aligner to reuse; default one per call.
    A `.bai` index is not written."""
    from . import ava, sam
    from .aligner import Aligner
    paths = reads if isinstance(reads, list) else [reads]
    queries: dict[str, str] = {}
    for p in paths:
        queries.update(read_fasta(p))
    refs = read_fasta(reference)
    own = engine is None
    if own:
        engine = Aligner(preset="ava-ont" if tech == "ava-ont" else "map-ont")
    try:
        if tech == "ava-ont":
            records = ava.ava_align(engine, queries)
            lengths = {n: len(s) for n, s in queries.items()}
        else:
            records = [r for name, seq in refs.items() for r in ava.map_queries(engine, queries, name, seq)]
            lengths = {n: len(s) for n, s in refs.items()}
        for r in records:
            seq = queries[r.query_name]
            r.query_sequence = reverse_complement(seq) if r.is_reverse else seq
            if r.is_supplementary:                      # hard-clipped records carry only the aligned part
                a = r.cigartuples[0][1] if r.cigartuples[0][0] == 5 else 0
                r.query_sequence = r.query_sequence[a:a + r.query_alignment_length]
        cmd = f"svirlpool_b200 -x {tech} {aln_args}".strip()
        sam.write_bam(bamout, sam.sort_records(records, lengths), lengths, command=cmd)
    finally:
        if own:
            engine.close()
