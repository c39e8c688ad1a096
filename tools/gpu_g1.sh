#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
export SVP_CALL_TIMEOUT_S=60
for lpp in 8 16; do
SVP_MODE=waves SVP_TB_SERIAL=1 SVP_TB_LPP=$lpp SVP_TBG_MIN=0 timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x > $O/g1_tests_lpp$lpp.log 2>&1; echo "tests exit $?" >> $O/g1_tests_lpp$lpp.log
tail -4 $O/g1_tests_lpp$lpp.log
done
run() { name=$1; shift; env "$@" timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --scaling-pool 0 $EXTRA > $O/g1_$name.json 2>$O/g1_$name.err; python -c "
import json,sys
try:
    d=json.load(open('gpurun_out/g1_$name.json')); print('$name: regions/s', round(d['regions_per_s'],1), 'GCUPS', round(d['value']), 'e2e r/s', round(d['e2e']['regions_per_s'],1), d['kernel_ms'])
except Exception as e: print('$name failed', e); print(open('gpurun_out/g1_$name.err').read()[-600:])"; }
run serial_lpp8 SVP_MODE=waves SVP_ARENA_GB=150 SVP_TB_SERIAL=1 SVP_TB_LPP=8
run serial_lpp16 SVP_MODE=waves SVP_ARENA_GB=150 SVP_TB_SERIAL=1 SVP_TB_LPP=16
run serial_lpp32 SVP_MODE=waves SVP_ARENA_GB=150 SVP_TB_SERIAL=1 SVP_TB_LPP=32 SVP_CORRIDOR=0
run conc_lpp8 SVP_MODE=waves SVP_ARENA_GB=150 SVP_TB_LPP=8 SVP_TBW_MAX=1000000
