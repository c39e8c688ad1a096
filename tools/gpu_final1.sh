#!/bin/bash
# round-2 evidence run, 1 GPU: the whole GPU suite, smoke, the default bench line, configs 3 / 4, the CPU arm, one container per call,
# and the launch list of the default bench command
cd "$(dirname "$0")/.."
O=gpurun_out
export SVP_CALL_TIMEOUT_S=120
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/fin_smi.txt
timeout 900 python -m pytest tests -q -m gpu -x > $O/fin_tests.log 2>&1; echo "tests exit $?" >> $O/fin_tests.log; tail -3 $O/fin_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/fin_smoke.log 2>&1; tail -1 $O/fin_smoke.log
timeout 600 python bench.py > $O/fin_bench_c2.json 2> $O/fin_bench_c2.err; echo "bench exit $?"
timeout 300 python bench.py --workload config3 --no-cpu-baseline > $O/fin_bench_c3.json 2> $O/fin_bench_c3.err
timeout 300 python bench.py --workload config4 --no-cpu-baseline > $O/fin_bench_c4.json 2> $O/fin_bench_c4.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/fin_bench_ref.json 2> $O/fin_bench_ref.err
timeout 100 python tools/quick_perf.py 1 > $O/fin_quick1.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/ncu_launches_r02_v2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --scaling-pool 0 > $O/fin_ncu_l.log 2>&1; echo "launch list exit $?"
python - <<'PY'
import json
for w in ("c2","c3","c4"):
    d=json.load(open(f'gpurun_out/fin_bench_{w}.json'))
    print(w, 'regions/s', round(d['regions_per_s'],1), 'GCUPS', round(d['value']), 'e2e', round(d['e2e']['value']), round(d['e2e']['regions_per_s'],1), {k:(round(v,1) if isinstance(v,float) else v) for k,v in d['kernel_ms'].items() if k!='note'}, d.get('parity_checked',{}).get('mismatches'), 'intpipe', round(d['roofline']['int_pipe']['frac'],3), round(d['roofline']['int_pipe']['alu_pipe_14op']['frac'],3), 'hbm', round(d['roofline']['frac'],3))
    for k,v in d.get('strong_scaling',{}).items(): print('  ', k, {a:v[a] for a in ('loci','seconds','loci_per_s','gcups','int_pipe_frac','bad_status_pairs')})
print(open('gpurun_out/fin_bench_ref.json').read()[:400])
print(open('gpurun_out/fin_quick1.log').read()[-400:])
PY
