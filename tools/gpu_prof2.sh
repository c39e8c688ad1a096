#!/bin/bash
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
timeout 300 ncu --set full --clock-control none --import-source on -k regex:svp_walk -s 1 -c 1 -o /tmp/prof_walk -f python /tmp/one.py > $O/p2_ncu.log 2>&1
echo "ncu exit $?"; tail -3 $O/p2_ncu.log
python tools/ncu_summary.py /tmp/prof_walk.ncu-rep "ncu --set full -k regex:svp_walk (8 regions of config 2, one call)" > $O/p2_walk.json 2> $O/p2_sum.err
ncu -i /tmp/prof_walk.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; u=rows[1]; r=rows[2]
want=['smsp__inst_executed.sum','smsp__thread_inst_executed.sum','lts__t_requests_srcunit_tex.sum','lts__t_sectors_srcunit_tex_op_read.sum','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','smsp__warps_active.avg.per_cycle_active','sm__cycles_elapsed.max','smsp__cycles_active.avg','l1tex__m_xbar2l1tex_read_sectors.sum','smsp__inst_executed_op_global_ld.sum','smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio','lts__average_t_sector_hit_rate_srcunit_tex.pct' ]
for k,n in enumerate(h):
    if n in want or 'stalled' in n and 'per_issue_active' in n: print(n, r[k], u[k])
" > $O/p2_raw.txt 2>&1
ncu -i /tmp/prof_walk.ncu-rep --page source --csv 2>/dev/null > /tmp/src.csv; python - <<'PY' > gpurun_out/p2_source_top.txt 2>&1
import csv
rows=list(csv.reader(open('/tmp/src.csv')))
h=rows[0]
print(h[:12])
ci={n:i for i,n in enumerate(h)}
key=[n for n in h if 'Sampling' in n and 'All' in n][:1] or [n for n in h if 'Samples' in n][:1]
print(key)
if key:
    k=ci[key[0]]
    def val(r):
        try: return float(r[k].replace(',',''))
        except: return 0.0
    top=sorted(rows[1:], key=val, reverse=True)[:40]
    for r in top: print(val(r), r[ci.get('Source',1)][:110] if 'Source' in ci else r[:3])
PY
cat $O/p2_raw.txt | head -40; head -60 $O/p2_source_top.txt
