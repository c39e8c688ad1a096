"""ctypes binding of the CPU oracle (oracle/libsvp_oracle.so) with the same engine interface as
svirlpool_b200.aligner.Aligner. TEST INFRASTRUCTURE: importable from tests/, __graft_entry__.smoke()
and bench.py's CPU arms only."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

from svirlpool_b200._abi import RESULT_DTYPE, SCORING_DTYPE, SVP_OK, ptr

ORACLE_DIR = Path(__file__).resolve().parent.parent / "oracle"
LIB = ORACLE_DIR / "libsvp_oracle.so"


def build_oracle(force: bool = False) -> Path:
    src = ORACLE_DIR / "svp_oracle.c"
    if force or not LIB.exists() or LIB.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(ORACLE_DIR)], check=True, capture_output=True)
    return LIB


class OracleEngine:
    def __init__(self, preset: str = "map-ont", scoring: np.ndarray | None = None, threads: int = 1):
        self.lib = C.CDLL(str(build_oracle()))
        self.lib.svpo_align_batch.restype = C.c_int
        self.lib.svpo_scoring_preset.restype = C.c_int
        if scoring is None:
            scoring = np.zeros(1, dtype=SCORING_DTYPE)
            rc = self.lib.svpo_scoring_preset(preset.encode(), ptr(scoring))
            if rc != SVP_OK:
                raise ValueError(f"unknown scoring preset {preset!r}")
        self.scoring = scoring
        self.set_threads(threads)

    def set_threads(self, n: int) -> None:
        self.lib.svpo_set_threads(C.c_int(int(n)))

    def align_batch(self, packed, seq_off, seq_len, pairs, cigar_words):
        res = np.zeros(len(pairs), dtype=RESULT_DTYPE)
        cig = np.zeros(max(int(cigar_words), 1), dtype=np.uint32)
        rc = self.lib.svpo_align_batch(ptr(self.scoring), ptr(packed), C.c_size_t(packed.nbytes), ptr(seq_off), ptr(seq_len),
                                       C.c_uint32(len(seq_len)), ptr(pairs), C.c_uint32(len(pairs)), ptr(res), ptr(cig),
                                       C.c_size_t(len(cig)))
        if rc != SVP_OK:
            raise RuntimeError(f"svpo_align_batch failed: {rc}")
        return res, cig
