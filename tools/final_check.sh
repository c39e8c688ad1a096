# Round-end sanity on the GPU box: the -m gpu suite, smoke(), and one default bench line.
python -m pytest tests -q -m gpu > gpurun_out/t_final.log 2>&1; tail -2 gpurun_out/t_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/final6_c2.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/final6_c2.json')); print('c2', round(d['value']), round(d['e2e']['value']), d['kernel_ms'])"
