#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
export SVP_CALL_TIMEOUT_S=20
timeout 150 python - > $O/dbg1.log 2>&1 <<'PY'
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
from svirlpool_b200 import workloads
from svirlpool_b200.aligner import Aligner, AlignerError
for regions, arena in (([0], 24), ([0, 1], 24), (list(range(8)), 48)):
    b = workloads.batch_config2(regions)
    with Aligner(device=0, preset="map-ont", tb_bytes=arena << 30) as g:
        t0 = time.time()
        try:
            res, cig = g.align_batch(b["packed"], b["off"], b["lens"], b["pairs"], b["cigar_words"])
            print(len(regions), "regions ok", round(time.time() - t0, 3), "s", g.stats(), flush=True)
        except AlignerError as e:
            print(len(regions), "regions FAILED:", e, flush=True)
            break
PY
echo "exit $?" >> $O/dbg1.log; cat $O/dbg1.log
