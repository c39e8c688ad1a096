#!/bin/bash
# GPU pass of the fused (wave-free) build: sanitizer on a small case first, then tests, then bench
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 300 python tools/sanitize_cases.py > $O/r2_cases.log 2>&1; echo "cases exit $?" >> $O/r2_cases.log
tail -6 $O/r2_cases.log
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > $O/r2_parity.log 2>&1; echo "parity exit $?" >> $O/r2_parity.log
tail -15 $O/r2_parity.log
timeout 1500 python -m pytest tests -q -m gpu --deselect tests/test_gpu_parity.py > $O/r2_rest.log 2>&1; echo "rest exit $?" >> $O/r2_rest.log
tail -15 $O/r2_rest.log
timeout 900 python bench.py --steps 6 --warmup 3 > $O/r2_bench.json 2> $O/r2_bench.err; echo "bench exit $?" >> $O/r2_bench.err
tail -3 $O/r2_bench.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2_bench.json'))
    print('value',round(d['value']),'regions/s',round(d['regions_per_s'],1),'e2e',round(d['e2e']['value']),'e2e r/s',round(d['e2e']['regions_per_s'],1),d['kernel_ms'])
    print('parity',d.get('parity_checked'))
    for k,v in d.get('strong_scaling',{}).items(): print(k,{x:v[x] for x in ('loci','seconds','loci_per_s','gcups','int_pipe_frac','lpt_imbalance_actual_cells','bad_status_pairs')})
except Exception as e: print('no bench json',e)
PY
