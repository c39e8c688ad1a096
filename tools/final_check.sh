python -m pytest tests -q -m gpu -x > gpurun_out/t_final.log 2>&1; tail -3 gpurun_out/t_final.log
python bench.py --no-cpu-baseline --workload config4 > gpurun_out/final3_c4.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/final3_c4.json')); print('c4', round(d['value']), round(d['e2e']['value']), d['kernel_ms'])"
python bench.py --no-cpu-baseline --workload config4 --no-pipeline > gpurun_out/final3_c4np.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/final3_c4np.json')); print('c4 nopipe', round(d['value']), round(d['e2e']['value']), d['kernel_ms'])"
python bench.py --no-cpu-baseline --workload config3 > gpurun_out/final3_c3.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/final3_c3.json')); print('c3', round(d['value']), round(d['e2e']['value']), d['kernel_ms'])"
