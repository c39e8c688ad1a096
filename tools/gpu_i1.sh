#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
export SVP_CALL_TIMEOUT_S=120
run() { name=$1; w=$2; shift; shift; env "$@" timeout 300 python bench.py --workload $w --no-cpu-baseline --scaling-pool 0 > $O/i1_$name.json 2>$O/i1_$name.err; python -c "
import json,sys
try:
    d=json.load(open('gpurun_out/i1_$name.json')); print('$name: regions/s', round(d['regions_per_s'],1), 'GCUPS', round(d['value']), 'e2e r/s', round(d['e2e']['regions_per_s'],1), {k:(round(v,1) if isinstance(v,float) else v) for k,v in d['kernel_ms'].items() if k!='note'}, 'launches', d['roofline']['launches'])
except Exception as e: print('$name failed', e); print(open('gpurun_out/i1_$name.err').read()[-600:])"; }
run c3_kp48 config3 SVP_KP_MASK=17
run c3_nocap config3 SVP_LIB=$PWD/svirlpool_b200/libsvp_align_nocap.so
run c3_nocap_kp48 config3 SVP_KP_MASK=17 SVP_LIB=$PWD/svirlpool_b200/libsvp_align_nocap.so
run c4_kp48 config4 SVP_KP_MASK=17
run c4_nocap config4 SVP_LIB=$PWD/svirlpool_b200/libsvp_align_nocap.so
run c4_lpp32 config4 SVP_TB_LPP=32
run c2_nocap config2 SVP_LIB=$PWD/svirlpool_b200/libsvp_align_nocap.so
run c2_kp48 config2 SVP_KP_MASK=17
