#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
run() { name=$1; shift; env "$@" timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --scaling-pool 0 > $O/exp_$name.json 2>$O/exp_$name.err; python -c "
import json,sys
try:
    d=json.load(open('gpurun_out/exp_$name.json')); print('$name: regions/s', round(d['regions_per_s'],1), 'GCUPS', round(d['value']), 'e2e r/s', round(d['e2e']['regions_per_s'],1), 'walk share', round(d['kernel_ms']['walk_share_ms']/d['kernel_ms']['device_total'],2))
except Exception as e: print('$name failed', e)"; }
run fwdonly SVP_FUSED_WALK=1 SVP_SKIP_WALK=1
run fwdonly_kp8 SVP_FUSED_WALK=1 SVP_SKIP_WALK=1 SVP_KP_MASK=17
run fwdonly_1stream SVP_FUSED_WALK=1 SVP_SKIP_WALK=1 SVP_FWD_STREAMS=1
