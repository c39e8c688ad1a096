"""last-train `.mat` files -> svp_scoring.

The reference hands a trained LAST scoring file to lamassemble (`--lamassemble-mat`, consensus.py:764-813:
`lamassemble ... -s 2 -g 67 {lamassemble_mat} {reads}`; shipped files: data/lamassemble-mats/{promethion,promethion-rna,
sequel-II-CLR}.mat, stored through Git LFS). The file is the output of `last-train`: a log of the training rounds as comment
lines, and the final parameters as

    #last -Q 0
    #last -t4.40198
    #last -a 15          gap existence cost     (deletion: letters of the reference missing from the query)
    #last -A 16          insertion existence cost
    #last -b 3           gap extension cost
    #last -B 3           insertion extension cost
    #last -S 1
    # score matrix (query letters = columns, reference letters = rows):
           A      C      G      T
    A      5    -38    -16    -35
    C    -40      6    -40    -14
    G    -15    -38      5    -39
    T    -37    -14    -37      6

(the numbers above only illustrate the layout). Later `#last` lines override earlier ones, the last matrix block wins. A gap of
k letters costs existence + k * extension, which is svp_scoring's open + k * ext. LAST gives letters outside the matrix (N) the
lowest score of the matrix; so does this library.

A checkout without `git lfs pull` holds a three-line pointer instead of the matrix; that is reported as such, not parsed.
"""
from __future__ import annotations

import re
from pathlib import Path

import numpy as np

from ._abi import SCORING_DTYPE

LFS_MAGIC = "version https://git-lfs.github.com/spec/"


class LastMatError(ValueError):
    """The file is not a usable last-train scoring file."""


def parse_last_mat(path: Path | str, min_signal_size: int = 12, zdrop: int = 400) -> np.ndarray:
    """Read a last-train `.mat` and return the svp_scoring record (numpy, SCORING_DTYPE) for Aligner(scoring=...): the 4x4
    matrix as s[query][target], deletion costs from -a / -b, insertion costs from -A / -B (-A / -B default to -a / -b as in
    lastal), one gap piece. Raises LastMatError for a Git-LFS pointer, a missing matrix or missing gap costs."""
    path = Path(path)
    text = path.read_text(errors="replace")
    if text.lstrip().startswith(LFS_MAGIC):
        raise LastMatError(f"{path} is a Git-LFS pointer, not a matrix: run `git lfs pull` (or download the file) to get the "
                           "last-train scoring file")
    opts: dict[str, float] = {}
    blocks: list[tuple[list[str], dict[str, list[int]]]] = []
    cols: list[str] | None = None
    rows: dict[str, list[int]] = {}
    for raw in text.splitlines():
        line = raw.strip()
        m = re.match(r"^#last\s+-([A-Za-z])\s*(\S+)", line)
        if m:
            try:
                opts[m.group(1)] = float(m.group(2))
            except ValueError:
                pass
            continue
        if not line or line.startswith("#"):
            if cols is not None and rows:
                blocks.append((cols, rows))
            cols, rows = None, {}
            continue
        tok = line.split()
        if cols is None:
            if all(len(t) == 1 and t.isalpha() for t in tok) and len(tok) >= 4:
                cols = [t.upper() for t in tok]
            continue
        if len(tok) == len(cols) + 1 and len(tok[0]) == 1 and tok[0].isalpha():
            try:
                rows[tok[0].upper()] = [int(round(float(x))) for x in tok[1:]]
                continue
            except ValueError:
                pass
        if rows:
            blocks.append((cols, rows))
        cols, rows = None, {}
    if cols is not None and rows:
        blocks.append((cols, rows))
    blocks = [b for b in blocks if all(x in b[0] for x in "ACGT") and all(x in b[1] for x in "ACGT")]
    if not blocks:
        raise LastMatError(f"{path}: no score matrix with rows and columns A, C, G, T found")
    if "a" not in opts or "b" not in opts:
        raise LastMatError(f"{path}: gap costs (#last -a / -b) not found")
    cols, rows = blocks[-1]
    s = np.zeros(1, dtype=SCORING_DTYPE)
    for qi, q in enumerate("ACGT"):                       # matrix[q * 4 + t]; file: rows = reference (target) letters, columns = query letters
        for ti, t in enumerate("ACGT"):
            s["matrix"][0][qi * 4 + ti] = rows[t][cols.index(q)]
    s["del_open1"], s["del_ext1"] = int(opts["a"]), int(opts["b"])
    s["ins_open1"], s["ins_ext1"] = int(opts.get("A", opts["a"])), int(opts.get("B", opts["b"]))
    s["del_open2"] = s["ins_open2"] = -1
    s["min_signal_size"], s["zdrop"] = min_signal_size, zdrop
    s["zdrop_ext"] = max(int(s["del_ext1"][0]), int(s["ins_ext1"][0]))
    return s


def stand_in_mat_text() -> str:
    """A labelled STAND-IN in last-train's format for tests and benchmarks: the layout is real, the numbers are not a trained
    model (the shipped .mat files are LFS pointers in this checkout)."""
    return "\n".join([
        "# STAND-IN scoring file in last-train format; numbers are illustrative, not trained",
        "#last -Q 0", "#last -t4.4", "#last -a 21", "#last -A 25", "#last -b 9", "#last -B 7", "#last -S 1",
        "# score matrix (query letters = columns, reference letters = rows):",
        "       A      C      G      T",
        "A      6    -18     -9    -19",
        "C    -18      5    -20     -8",
        "G     -9    -20      5    -18",
        "T    -19     -8    -18      6", ""])
