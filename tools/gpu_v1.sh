#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
export SVP_CALL_TIMEOUT_S=60
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_chain.py -q -m gpu -x > $O/v1_tests.log 2>&1; echo "tests exit $?" >> $O/v1_tests.log
tail -3 $O/v1_tests.log
run() { name=$1; w=$2; shift; shift; env "$@" timeout 300 python bench.py --workload $w --no-cpu-baseline --scaling-pool 0 > $O/v1_$name.json 2>$O/v1_$name.err; python -c "
import json,sys
try:
    d=json.load(open('gpurun_out/v1_$name.json')); print('$name: regions/s', round(d['regions_per_s'],1), 'GCUPS', round(d['value']), 'e2e r/s', round(d['e2e']['regions_per_s'],1), 'e2e wall ms/step', round(d['e2e']['wall_ms_per_step'],1), 'dev ms/step', round(d['ms_per_step'],1), {k:(round(v,1) if isinstance(v,float) else v) for k,v in d['kernel_ms'].items() if k!='note'})
except Exception as e: print('$name failed', e); print(open('gpurun_out/v1_$name.err').read()[-600:])"; }
run c2 config2 A=1
run c2_1thread config2 SVP_PLAN_THREADS=1
