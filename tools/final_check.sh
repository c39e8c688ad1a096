python -m pytest tests -q -m gpu > gpurun_out/t_final.log 2>&1; tail -2 gpurun_out/t_final.log
python bench.py > gpurun_out/final4_c2.json 2>/dev/null
python bench.py --no-cpu-baseline --workload config4 > gpurun_out/final4_c4.json 2>/dev/null
python bench.py --no-cpu-baseline --workload config3 > gpurun_out/final4_c3.json 2>/dev/null
python - <<'PY'
import json
for n in ("c2","c4","c3"):
    d=json.load(open(f"gpurun_out/final4_{n}.json")); print(n, round(d["value"]), round(d["e2e"]["value"]), round(d["regions_per_s"]), {k: round(v,1) for k,v in d["kernel_ms"].items()})
PY
