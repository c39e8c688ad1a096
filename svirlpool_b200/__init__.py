"""svirlpool_b200 — B200-native pairwise-alignment hot path for Svirlpool.

Only what the path named by BASELINE.json:north_star needs lives here:
  csrc/               CUDA kernels (sm_100a) + the C-ABI (include/svp_align.h) -> libsvp_align.so
  aligner.py          ctypes binding of the C-ABI (fails loudly without the CUDA library/GPU)
  ava.py              read-vs-read (AVA) and query-vs-window drivers: batching, band policy, records
  consensus_lib.py    host mirror of svirlpool.localassembly.consensus_lib (signals -> similarity)
  alignments_to_rafs.py / util.py / datatypes.py   the CIGAR -> SVsignal consumers the path feeds
  crs_to_batches.py   batching API kept compatible with svirlpool.candidateregions.crs_to_batches
  scheduler.py        size-balanced (LPT) region -> GPU assignment
  synth.py            seeded synthetic ONT-like regions for the BASELINE configs
"""
__version__ = "0.1.0"
