#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
export SVP_CALL_TIMEOUT_S=60
timeout 900 python -m pytest tests -q -m gpu -x > $O/w1_tests.log 2>&1; echo "tests exit $?" >> $O/w1_tests.log
tail -3 $O/w1_tests.log
run() { name=$1; w=$2; shift; shift; env "$@" timeout 300 python bench.py --workload $w --no-cpu-baseline --scaling-pool 0 > $O/w1_$name.json 2>$O/w1_$name.err; python -c "
import json,sys
try:
    d=json.load(open('gpurun_out/w1_$name.json')); print('$name: regions/s', round(d['regions_per_s'],1), 'GCUPS', round(d['value']), 'e2e r/s', round(d['e2e']['regions_per_s'],1), 'e2e wall ms/step', round(d['e2e']['wall_ms_per_step'],1), 'dev ms/step', round(d['ms_per_step'],1))
except Exception as e: print('$name failed', e); print(open('gpurun_out/w1_$name.err').read()[-600:])"; }
run c2 config2 A=1
run c2_nostream config2 SVP_STREAM_BACK=0
