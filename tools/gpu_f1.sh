#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
export SVP_CALL_TIMEOUT_S=60
SVP_MODE=waves timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x > $O/f1_tests_waves.log 2>&1; echo "tests exit $?" >> $O/f1_tests_waves.log
tail -4 $O/f1_tests_waves.log
run() { name=$1; shift; env "$@" timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --scaling-pool 0 $EXTRA > $O/f1_$name.json 2>$O/f1_$name.err; python -c "
import json,sys
try:
    d=json.load(open('gpurun_out/f1_$name.json')); print('$name: regions/s', round(d['regions_per_s'],1), 'GCUPS', round(d['value']), 'e2e r/s', round(d['e2e']['regions_per_s'],1), d['kernel_ms'])
except Exception as e: print('$name failed', e); print(open('gpurun_out/f1_$name.err').read()[-600:])"; }
run waves_150 SVP_MODE=waves SVP_ARENA_GB=150
run waves_def SVP_MODE=waves
run waves_150_tbw SVP_MODE=waves SVP_ARENA_GB=150 SVP_TBW_MAX=1000000
run waves_150_skip SVP_MODE=waves SVP_ARENA_GB=150 SVP_SKIP_WALK=1
run waves_150_serial SVP_MODE=waves SVP_ARENA_GB=150 SVP_TB_SERIAL=1 SVP_CORRIDOR=0
