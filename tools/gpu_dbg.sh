#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
export SVP_CALL_TIMEOUT_S=30
timeout 280 python - > $O/dbg1.log 2>&1 <<'PY'
import sys, time, os
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
from svirlpool_b200 import workloads
from svirlpool_b200._abi import cigartuples
from svirlpool_b200.aligner import Aligner, AlignerError
from oracle_engine import OracleEngine
b = workloads.batch_config2([0, 1])
ores, ocig = OracleEngine("map-ont", threads=os.cpu_count()).align_batch(b["packed"], b["off"], b["lens"], b["pairs"], b["cigar_words"])
with Aligner(device=0, preset="map-ont") as g:
    res, cig = g.align_batch(b["packed"], b["off"], b["lens"], b["pairs"], b["cigar_words"])
    bad = sum(1 for k in range(len(res)) if any(res[f][k] != ores[f][k] for f in ("score", "status", "q_beg", "q_end", "t_beg", "t_end", "n_match", "n_mismatch", "cigar_len"))
              or cigartuples(cig, res[k]) != cigartuples(ocig, ores[k]))
    print("parity 2 regions: mismatches", bad, flush=True)
    for n in (8, 24, 96):
        bb = workloads.batch_config2(list(range(100, 100 + n)))
        for rep in range(3):
            t0 = time.time()
            g.align_batch_summary(bb["packed"], bb["off"], bb["lens"], bb["pairs"], bb["cigar_words"], copy=False)
            dt = time.time() - t0
            st = g.stats()
        print(n, "regions: wall", round(dt * 1e3, 1), "ms device", round(st["total_ms"], 1), "ms ->", round(n / st["total_ms"] * 1e3, 1), "regions/s",
              round(st["cells"] / st["total_ms"] / 1e6), "GCUPS walk share", round(st["tb_ms"] / st["total_ms"], 2), "groups", st["waves"], flush=True)
PY
echo "exit $?" >> $O/dbg1.log; cat $O/dbg1.log
