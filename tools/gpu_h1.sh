#!/bin/bash
# default build (wave scheduling, serial group walks): whole GPU suite, default bench (with CPU arm, parity, scaling pools), configs 3 / 4
cd "$(dirname "$0")/.."
O=gpurun_out
export SVP_CALL_TIMEOUT_S=120
timeout 900 python -m pytest tests -q -m gpu -x > $O/h1_tests.log 2>&1; echo "tests exit $?" >> $O/h1_tests.log
tail -6 $O/h1_tests.log
timeout 900 python bench.py > $O/h1_bench.json 2> $O/h1_bench.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/h1_bench.json'))
print('regions/s', round(d['regions_per_s'],1), 'GCUPS', round(d['value']), 'e2e r/s', round(d['e2e']['regions_per_s'],1), d['kernel_ms'], d.get('parity_checked'), d.get('cpu_baseline',{}).get('value'))
for k,v in d.get('strong_scaling',{}).items(): print(k, {a:v[a] for a in ('loci','seconds','loci_per_s','gcups','int_pipe_frac','lpt_imbalance_actual_cells','bad_status_pairs')})
PY
for w in config3 config4; do
timeout 300 python bench.py --workload $w --no-cpu-baseline > $O/h1_$w.json 2> $O/h1_$w.err; python -c "
import json
d=json.load(open('gpurun_out/h1_$w.json')); print('$w: regions/s', round(d['regions_per_s'],1), 'GCUPS', round(d['value']), 'e2e', round(d['e2e']['value']), d['kernel_ms'])"
SVP_TB_SERIAL=0 timeout 300 python bench.py --workload $w --no-cpu-baseline > $O/h1_${w}_conc.json 2> $O/h1_${w}_conc.err; python -c "
import json
d=json.load(open('gpurun_out/h1_${w}_conc.json')); print('$w conc: regions/s', round(d['regions_per_s'],1), 'GCUPS', round(d['value']), 'e2e', round(d['e2e']['value']), d['kernel_ms'])"
done
