#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
export SVP_CALL_TIMEOUT_S=120
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "band_widths or schedulers or arena_pressure" > $O/l1_tests.log 2>&1; echo "tests exit $?" >> $O/l1_tests.log
tail -5 $O/l1_tests.log
run() { name=$1; w=$2; shift; shift; env "$@" timeout 300 python bench.py --workload $w --no-cpu-baseline --scaling-pool 0 > $O/l1_$name.json 2>$O/l1_$name.err; python -c "
import json,sys
try:
    d=json.load(open('gpurun_out/l1_$name.json')); print('$name: regions/s', round(d['regions_per_s'],1), 'GCUPS', round(d['value']), 'e2e r/s', round(d['e2e']['regions_per_s'],1), {k:(round(v,1) if isinstance(v,float) else v) for k,v in d['kernel_ms'].items() if k!='note'}, 'launches', d['roofline']['launches'])
except Exception as e: print('$name failed', e); print(open('gpurun_out/l1_$name.err').read()[-600:])"; }
run c2_kp12 config2 A=1
run c2_kp12_w12 config2 SVP_LIB=$PWD/svirlpool_b200/libsvp_align_w12.so
run c2_nokp12 config2 SVP_KP_MASK=31
run c3 config3 A=1
run c4_kp12 config4 A=1
