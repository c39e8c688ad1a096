#!/bin/bash
# queue scheduling (resident walker kernel, one walk per lane): parity tests, then A/B bench runs
cd "$(dirname "$0")/.."
O=gpurun_out
export SVP_CALL_TIMEOUT_S=60
SVP_MODE=queue SVP_QUEUE_MIN_PAIRS=16 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_chain.py -q -m gpu -x > $O/b1_tests.log 2>&1; echo "tests exit $?" >> $O/b1_tests.log
tail -15 $O/b1_tests.log
run() { name=$1; shift; env "$@" timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --scaling-pool 0 > $O/b1_$name.json 2>$O/b1_$name.err; python -c "
import json,sys
try:
    d=json.load(open('gpurun_out/b1_$name.json')); print('$name: regions/s', round(d['regions_per_s'],1), 'GCUPS', round(d['value']), 'e2e r/s', round(d['e2e']['regions_per_s'],1), d['kernel_ms'])
except Exception as e: print('$name failed', e); print(open('gpurun_out/b1_$name.err').read()[-600:])"; }
run queue SVP_MODE=queue
run queue_lean SVP_MODE=queue SVP_WALK_LEAN=1
run queue_w2 SVP_MODE=queue SVP_WALK_CTAS_PER_SM=2
run queue_skip SVP_MODE=queue SVP_SKIP_WALK=1
run queue_wait20 SVP_MODE=queue SVP_WALK_WAIT_NS=20000
