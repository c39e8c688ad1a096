#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
export SVP_CALL_TIMEOUT_S=60
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_workloads.py -q -m gpu -x > $O/t1_tests.log 2>&1; echo "tests exit $?" >> $O/t1_tests.log
tail -3 $O/t1_tests.log
run() { name=$1; w=$2; shift; shift; env "$@" timeout 300 python bench.py --workload $w --no-cpu-baseline --scaling-pool 0 > $O/t1_$name.json 2>$O/t1_$name.err; python -c "
import json,sys
try:
    d=json.load(open('gpurun_out/t1_$name.json')); print('$name: regions/s', round(d['regions_per_s'],1), 'GCUPS', round(d['value']), 'e2e r/s', round(d['e2e']['regions_per_s'],1), {k:(round(v,1) if isinstance(v,float) else v) for k,v in d['kernel_ms'].items() if k!='note'}, 'launches', d['roofline']['launches'])
except Exception as e: print('$name failed', e); print(open('gpurun_out/t1_$name.err').read()[-600:])"; }
run c2 config2 A=1
run c2_old config2 SVP_RING_MIN_SINGLE=4096
run c3 config3 A=1
run c4 config4 A=1
run c4_old config4 SVP_RING_MIN_SINGLE=4096
