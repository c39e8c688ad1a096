#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
export SVP_CALL_TIMEOUT_S=60
SVP_MODE=arena timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x > $O/e1_tests_arena.log 2>&1; echo "tests exit $?" >> $O/e1_tests_arena.log
tail -4 $O/e1_tests_arena.log
run() { name=$1; shift; env "$@" timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --scaling-pool 0 $EXTRA > $O/e1_$name.json 2>$O/e1_$name.err; python -c "
import json,sys
try:
    d=json.load(open('gpurun_out/e1_$name.json')); print('$name: regions/s', round(d['regions_per_s'],1), 'GCUPS', round(d['value']), 'e2e r/s', round(d['e2e']['regions_per_s'],1), d['kernel_ms'])
except Exception as e: print('$name failed', e); print(open('gpurun_out/e1_$name.err').read()[-600:])"; }
EXTRA="--regions 24 --no-pipeline" run serial_cor SVP_MODE=waves SVP_TB_SERIAL=1
EXTRA="--regions 24 --no-pipeline" run serial_nocor SVP_MODE=waves SVP_TB_SERIAL=1 SVP_CORRIDOR=0
EXTRA="" run arena SVP_MODE=arena
EXTRA="" run arena_nocor SVP_MODE=arena SVP_CORRIDOR=0
EXTRA="" run waves_tbw SVP_MODE=waves SVP_TBW_MAX=1000000 SVP_ARENA_GB=150
