#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
export SVP_CALL_TIMEOUT_S=60
run() { name=$1; w=$2; shift; shift; env "$@" timeout 300 python bench.py --workload $w --no-cpu-baseline --scaling-pool 0 > $O/y1_$name.json 2>$O/y1_$name.err; python -c "
import json,sys
try:
    d=json.load(open('gpurun_out/y1_$name.json')); print('$name: regions/s', round(d['regions_per_s'],1), 'GCUPS', round(d['value']), 'e2e r/s', round(d['e2e']['regions_per_s'],1), {k:(round(v,1) if isinstance(v,float) else v) for k,v in d['kernel_ms'].items() if k!='note'})
except Exception as e: print('$name failed', e); print(open('gpurun_out/y1_$name.err').read()[-600:])"; }
run base config2 A=1
run wm1 config2 SVP_LIB=$PWD/svirlpool_b200/libsvp_align_wm1.so
run wm12 config2 SVP_LIB=$PWD/svirlpool_b200/libsvp_align_wm12.so
run wm20 config2 SVP_LIB=$PWD/svirlpool_b200/libsvp_align_wm20.so
