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
