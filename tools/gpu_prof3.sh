#!/bin/bash
# ncu --set full of the warp-per-pair walk kernel alone (wave scheduling, serial walks), per-SASS-line executed instruction counts
cd "$(dirname "$0")/.."
O=gpurun_out
cat > /tmp/one.py <<'PY'
import sys
sys.path.insert(0, ".")
from svirlpool_b200 import workloads
from svirlpool_b200.aligner import Aligner
b = workloads.batch_config2(list(range(100, 108)))
with Aligner(device=0, preset="map-ont") as g:
    for _ in range(2):
        g.align_batch_summary(b["packed"], b["off"], b["lens"], b["pairs"], b["cigar_words"], copy=False)
    print(g.stats())
PY
export SVP_MODE=waves SVP_TB_SERIAL=1 SVP_CORRIDOR=${COR:-1}
timeout 120 python /tmp/one.py > $O/p3_plain.log 2>&1; echo "plain exit $?"; tail -1 $O/p3_plain.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:svp_tbw -s 1 -c 1 -o /tmp/prof_tbw -f python /tmp/one.py > $O/p3_ncu.log 2>&1
echo "ncu exit $?"; tail -2 $O/p3_ncu.log
python tools/ncu_summary.py /tmp/prof_tbw.ncu-rep "SVP_MODE=waves SVP_TB_SERIAL=1 ncu --set full --import-source on -k regex:svp_tbw -s 1 -c 1 (8 regions of config 2, one call, corridor=$SVP_CORRIDOR)" > $O/p3_tbw.json 2> $O/p3_sum.err
ncu -i /tmp/prof_tbw.ncu-rep --page source --csv > $O/p3_source.csv 2>/dev/null
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/p3_source.csv')))
h=rows[0]; print(h)
PY
cp /tmp/prof_tbw.ncu-rep $O/p3_tbw.ncu-rep
cat $O/p3_tbw.json | head -40
