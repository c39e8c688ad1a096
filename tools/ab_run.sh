#!/bin/bash
# scratch A/B driver for gpurun: wave overlap on configs 2/3/4 + tests
for ov in 0 1; do
  for c in config3 config4; do SVP_WAVE_OVERLAP=$ov python bench.py --workload $c --steps 3 --warmup 3 > gpurun_out/bench_${c}_ov$ov.json 2>/dev/null; done
  SVP_WAVE_OVERLAP=$ov python bench.py --steps 4 --warmup 3 > gpurun_out/bench_config2_ov$ov.json 2>/dev/null
done
SVP_WAVE_OVERLAP=0 SVP_FWD_STREAMS=8 python bench.py --workload config3 --steps 3 --warmup 3 > gpurun_out/bench_config3_ov0s8.json 2>/dev/null
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_config?_ov*.json")):
    d=json.load(open(f)); print(f, round(d["value"]), round(d["e2e"]["value"]), round(d["regions_per_s"]), d["kernel_ms"])
PY
python -m pytest tests -q -m gpu -x > gpurun_out/t_all3.log 2>&1; tail -3 gpurun_out/t_all3.log
SVP_WAVE_OVERLAP=1 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "multiple_waves or noisy" > gpurun_out/t_ov.log 2>&1; tail -3 gpurun_out/t_ov.log
