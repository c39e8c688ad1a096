#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
export SVP_CALL_TIMEOUT_S=60
run() { name=$1; shift; env "$@" timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --scaling-pool 0 > $O/d1_$name.json 2>$O/d1_$name.err; python -c "
import json,sys
try:
    d=json.load(open('gpurun_out/d1_$name.json')); print('$name: regions/s', round(d['regions_per_s'],1), 'GCUPS', round(d['value']), 'e2e r/s', round(d['e2e']['regions_per_s'],1), d['kernel_ms'])
except Exception as e: print('$name failed', e); print(open('gpurun_out/d1_$name.err').read()[-600:])"; }
run waves_nojoin SVP_MODE=waves SVP_WAVE_JOIN=0
run waves_nojoin_150 SVP_MODE=waves SVP_WAVE_JOIN=0 SVP_ARENA_GB=150
run waves_join_150 SVP_MODE=waves SVP_ARENA_GB=150
run waves_nojoin_150_s8 SVP_MODE=waves SVP_WAVE_JOIN=0 SVP_ARENA_GB=150 SVP_FWD_STREAMS=8
run queue_p160_w2 SVP_MODE=queue SVP_ARENA_GB=140 SVP_WALK_CTAS_PER_SM=2 SVP_LIB=$PWD/svirlpool_b200/libsvp_align_pool160.so
run queue_p160_w3 SVP_MODE=queue SVP_ARENA_GB=140 SVP_WALK_CTAS_PER_SM=3 SVP_LIB=$PWD/svirlpool_b200/libsvp_align_pool160.so
