#!/bin/bash
# shared-memory corridor of the warp walk: parity tests in arena and waves mode, then A/B bench runs
cd "$(dirname "$0")/.."
O=gpurun_out
export SVP_CALL_TIMEOUT_S=60
SVP_MODE=arena timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_chain.py -q -m gpu -x > $O/c1_tests_arena.log 2>&1; echo "tests exit $?" >> $O/c1_tests_arena.log
tail -4 $O/c1_tests_arena.log
SVP_MODE=waves SVP_TBW_MAX=1000000 timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x > $O/c1_tests_waves.log 2>&1; echo "tests exit $?" >> $O/c1_tests_waves.log
tail -4 $O/c1_tests_waves.log
run() { name=$1; shift; env "$@" timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --scaling-pool 0 > $O/c1_$name.json 2>$O/c1_$name.err; python -c "
import json,sys
try:
    d=json.load(open('gpurun_out/c1_$name.json')); print('$name: regions/s', round(d['regions_per_s'],1), 'GCUPS', round(d['value']), 'e2e r/s', round(d['e2e']['regions_per_s'],1), d['kernel_ms'])
except Exception as e: print('$name failed', e); print(open('gpurun_out/c1_$name.err').read()[-600:])"; }
run arena SVP_MODE=arena
run arena_nocor SVP_MODE=arena SVP_CORRIDOR=0
run waves_tbw SVP_MODE=waves SVP_TBW_MAX=1000000
run queue SVP_MODE=queue
