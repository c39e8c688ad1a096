#!/bin/bash
# --set full of one wave's forward launches (24 regions per step: every variant's grid still covers the GPU several times)
cd "$(dirname "$0")/.."
O=gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --regions 24 --distinct-batches 1 --no-cpu-baseline --scaling-pool 0 --no-pipeline"
timeout 100 $CMD > $O/k1_plain.json 2>$O/k1_plain.err; echo "plain exit $?"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:svp_fwd -s 8 -c 8 -o /tmp/prof_fwd -f $CMD > $O/k1_ncu_f.log 2>&1; echo "fwd capture exit $?"
python tools/ncu_summary.py /tmp/prof_fwd.ncu-rep "ncu --set full --clock-control none --import-source on -k regex:svp_fwd -s 8 -c 8 $CMD" > $O/ncu_fwd_r02_v2.json 2> $O/k1_sum_f.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/ncu_fwd_r02_v2.json'))
for l in d['launches']: print(l['kernel'], l['block'], l['grid'], 'ms', l['duration_ms'], 'alu', l.get('alu_pipe_pct'), 'issue', l.get('issue_slots_pct'), 'occ', l.get('achieved_occupancy_pct'), 'regs', l.get('registers'), 'instr', l.get('warp_instructions'), 'dramW', l.get('dram_write_gbyte'), l['top_stalls_warps_per_issue'])
PY
