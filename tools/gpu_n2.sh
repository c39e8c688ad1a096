#!/bin/bash
# two ranks on one box: the default bench command under torchrun (weak scaling of config 2 + the strong-scaling sub-records)
cd "$(dirname "$0")/.."
O=gpurun_out
export SVP_CALL_TIMEOUT_S=120
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 > $O/fin_bench_n2.json 2> $O/fin_bench_n2.err; echo "n2 exit $?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/fin_bench_n2.json') if l.startswith('{')][-1])
print('N=2 regions/s', round(d['regions_per_s'],1), 'GCUPS', round(d['value']), 'e2e', round(d['e2e']['value']), 'n_gpus', d['n_gpus'])
for k,v in d.get('strong_scaling',{}).items(): print('  ', k, {a:v[a] for a in ('loci','seconds','loci_per_s','gcups','int_pipe_frac','lpt_imbalance_estimated','lpt_imbalance_actual_cells','rank_busy_frac','bad_status_pairs')})
PY
tail -3 $O/fin_bench_n2.err
