#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
export SVP_CALL_TIMEOUT_S=60
SVP_MODE=queue SVP_QUEUE_MIN_PAIRS=16 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_chain.py -q -m gpu -x > $O/r1_tests_queue.log 2>&1; echo "tests exit $?" >> $O/r1_tests_queue.log
tail -4 $O/r1_tests_queue.log
run() { name=$1; w=$2; shift; shift; env "$@" timeout 300 python bench.py --workload $w --no-cpu-baseline --scaling-pool 0 > $O/r1_$name.json 2>$O/r1_$name.err; python -c "
import json,sys
try:
    d=json.load(open('gpurun_out/r1_$name.json')); print('$name: regions/s', round(d['regions_per_s'],1), 'GCUPS', round(d['value']), 'e2e r/s', round(d['e2e']['regions_per_s'],1), {k:(round(v,1) if isinstance(v,float) else v) for k,v in d['kernel_ms'].items() if k!='note'}, 'launches', d['roofline']['launches'])
except Exception as e: print('$name failed', e); print(open('gpurun_out/r1_$name.err').read()[-600:])"; }
run q_w4_64 config2 SVP_MODE=queue SVP_ARENA_GB=64
run q_w2_64 config2 SVP_MODE=queue SVP_ARENA_GB=64 SVP_WALK_CTAS_PER_SM=2
run q_w6_64 config2 SVP_MODE=queue SVP_ARENA_GB=64 SVP_WALK_CTAS_PER_SM=6
run q_w4_32 config2 SVP_MODE=queue SVP_ARENA_GB=32
