#!/bin/bash
# GPU pass with hang protection: every test has a 120 s limit (thread method: the process exits, which tears the context down),
# the risky geometries (whole-SM CTAs, clusters, tiny arenas) run first, the bench only when the tests are green.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
PT="python -m pytest -q -m gpu -p no:cacheprovider --timeout 120 --timeout-method thread"
timeout 400 $PT -x tests/test_gpu_parity.py -k "band_widths_kp_5_6_7 or cluster or arena or pipelined or two_handles or sliding" > $O/r3_risky.log 2>&1; rc=$?
echo "risky exit $rc" >> $O/r3_risky.log; tail -4 $O/r3_risky.log
if [ $rc -ne 0 ]; then tail -40 $O/r3_risky.log; exit 0; fi
timeout 900 $PT tests > $O/r3_all.log 2>&1; rc=$?
echo "all exit $rc" >> $O/r3_all.log; tail -12 $O/r3_all.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout 900 python bench.py --steps 6 --warmup 3 > $O/r3_bench.json 2> $O/r3_bench.err; echo "bench exit $?" >> $O/r3_bench.err
tail -3 $O/r3_bench.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r3_bench.json'))
    print('value',round(d['value']),'regions/s',round(d['regions_per_s'],1),'e2e',round(d['e2e']['value']),'e2e r/s',round(d['e2e']['regions_per_s'],1),d['kernel_ms'])
    print('parity',d.get('parity_checked',{}).get('mismatches'), 'g32', d.get('cells_granule32'))
    for k,v in d.get('strong_scaling',{}).items(): print(k,{x:v[x] for x in ('loci','seconds','loci_per_s','gcups','int_pipe_frac','lpt_imbalance_actual_cells','bad_status_pairs')})
except Exception as e: print('no bench json',e)
PY
for w in 2 4; do SVP_WALK_CTAS_PER_SM=$w timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --scaling-pool 0 > $O/r3_bench_w$w.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r3_bench_w$w.json')); print('walk ctas/sm $w: regions/s', round(d['regions_per_s'],1), d['kernel_ms']['device_total'])"; done
SVP_FUSED_WALK=1 timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --scaling-pool 0 > $O/r3_bench_fused.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r3_bench_fused.json')); print('fused: regions/s', round(d['regions_per_s'],1))"
