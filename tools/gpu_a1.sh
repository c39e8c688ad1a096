#!/bin/bash
# wave scheduling restored: parity tests, then A/B bench runs
cd "$(dirname "$0")/.."
O=gpurun_out
export SVP_CALL_TIMEOUT_S=60
timeout 420 python -m pytest tests/test_gpu_parity.py tests/test_gpu_chain.py -q -m gpu -x --deselect tests/test_gpu_parity.py::test_arena_pressure_and_stream_counts > $O/a1_tests.log 2>&1; echo "tests exit $?" >> $O/a1_tests.log
tail -5 $O/a1_tests.log
run() { name=$1; shift; env "$@" timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --scaling-pool 0 > $O/a1_$name.json 2>$O/a1_$name.err; python -c "
import json,sys
try:
    d=json.load(open('gpurun_out/a1_$name.json')); print('$name: regions/s', round(d['regions_per_s'],1), 'GCUPS', round(d['value']), 'e2e r/s', round(d['e2e']['regions_per_s'],1), d['kernel_ms'])
except Exception as e: print('$name failed', e)"; }
run waves SVP_MODE=waves
run waves_skip SVP_MODE=waves SVP_SKIP_WALK=1
run waves_s8 SVP_MODE=waves SVP_FWD_STREAMS=8
run arena SVP_MODE=arena
