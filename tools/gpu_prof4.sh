#!/bin/bash
# ncu --set full of the uniform group walk kernel (8 regions of config 2, serial waves), per-SASS-line executed instruction counts
cd "$(dirname "$0")/.."
O=gpurun_out
cat > /tmp/one.py <<'PY'
import sys, os
sys.path.insert(0, ".")
os.environ["SVP_SMALL_CALL_PAIRS"] = "0"
from svirlpool_b200 import workloads
from svirlpool_b200.aligner import Aligner
b = workloads.batch_config2(list(range(100, 108)))
with Aligner(device=0, preset="map-ont", tb_bytes=40 << 30) as g:
    for _ in range(2):
        g.align_batch_summary(b["packed"], b["off"], b["lens"], b["pairs"], b["cigar_words"], copy=False)
    print(g.stats())
PY
export SVP_MODE=waves SVP_TB_SERIAL=1 SVP_TBG_MIN=0
timeout 120 python /tmp/one.py > $O/p4_plain.log 2>&1; echo "plain exit $?"; tail -1 $O/p4_plain.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:svp_tbu -s 1 -c 1 -o /tmp/prof_tbu -f python /tmp/one.py > $O/p4_ncu.log 2>&1
echo "ncu exit $?"; tail -2 $O/p4_ncu.log
python tools/ncu_summary.py /tmp/prof_tbu.ncu-rep "SVP_MODE=waves SVP_TB_SERIAL=1 ncu --set full --import-source on -k regex:svp_tbu -s 1 -c 1 (8 regions of config 2 = 3480 pairs, one call)" > $O/ncu_tbu_r02_v2.json 2> $O/p4_sum.err
ncu -i /tmp/prof_tbu.ncu-rep --page source --csv > $O/p4_source.csv 2>/dev/null
cat $O/ncu_tbu_r02_v2.json | head -36
