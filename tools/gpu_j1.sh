#!/bin/bash
# KP rule A/B, then the r02 profiles: launch list of the default bench command, --set full of a wave's forward launches at full grid and of the
# group walk kernel, and the whole 10 000-region config-2 set once
cd "$(dirname "$0")/.."
O=gpurun_out
export SVP_CALL_TIMEOUT_S=120
run() { name=$1; w=$2; shift; shift; env "$@" timeout 300 python bench.py --workload $w --no-cpu-baseline --scaling-pool 0 > $O/j1_$name.json 2>$O/j1_$name.err; python -c "
import json,sys
try:
    d=json.load(open('gpurun_out/j1_$name.json')); print('$name: regions/s', round(d['regions_per_s'],1), 'GCUPS', round(d['value']), 'e2e r/s', round(d['e2e']['regions_per_s'],1), {k:(round(v,1) if isinstance(v,float) else v) for k,v in d['kernel_ms'].items() if k!='note'}, 'launches', d['roofline']['launches'])
except Exception as e: print('$name failed', e); print(open('gpurun_out/j1_$name.err').read()[-600:])"; }
run c2 config2 A=1
run c3 config3 A=1
run c4 config4 A=1
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --scaling-pool 0"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/ncu_launches_r02_v1.csv $CMD > $O/j1_ncu_l.log 2>&1; echo "launch list exit $?"
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --scaling-pool 0 --no-pipeline"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:svp_fwd -s 14 -c 14 -o /tmp/prof_fwd -f $CMD > $O/j1_ncu_f.log 2>&1; echo "fwd capture exit $?"
python tools/ncu_summary.py /tmp/prof_fwd.ncu-rep "ncu --set full --clock-control none --import-source on -k regex:svp_fwd -s 14 -c 14 $CMD (one full-size wave of the default bench: 96 regions per step)" > $O/ncu_fwd_r02_v1.json 2> $O/j1_sum_f.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:svp_tbg -s 1 -c 1 -o /tmp/prof_tbg -f $CMD > $O/j1_ncu_t.log 2>&1; echo "walk capture exit $?"
python tools/ncu_summary.py /tmp/prof_tbg.ncu-rep "ncu --set full --clock-control none --import-source on -k regex:svp_tbg -s 1 -c 1 $CMD" > $O/ncu_tbg_r02_v1.json 2> $O/j1_sum_t.err
ncu -i /tmp/prof_tbg.ncu-rep --page source --csv > $O/j1_tbg_source.csv 2>/dev/null
timeout 500 python tools/full_config2.py > $O/j1_full_config2.json 2> $O/j1_full_config2.err; echo "full config2 exit $?"; tail -c 1500 $O/j1_full_config2.json
du -sh $O
