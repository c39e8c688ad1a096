#!/bin/bash
python -m pytest tests -q -m gpu -x > gpurun_out/t_all5.log 2>&1; tail -3 gpurun_out/t_all5.log
SVP_TB_SERIAL=1 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "multiple_waves or pipelined or noisy or both_traceback" > gpurun_out/t_serial.log 2>&1; tail -2 gpurun_out/t_serial.log
b() { name=$1; shift; python bench.py --no-cpu-baseline "$@" > gpurun_out/ab5_$name.json 2>gpurun_out/ab5_$name.err; }
b c2_pipe --steps 6 --warmup 3
b c2_nopipe --steps 6 --warmup 3 --no-pipeline
SVP_TB_SERIAL=0 b c2_pipe_conc --steps 6 --warmup 3
for c in config3 config4; do
  b ${c}_pipe --workload $c --steps 4 --warmup 3
  b ${c}_nopipe --workload $c --steps 4 --warmup 3 --no-pipeline
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/ab5_*.json")):
    try:
        d=json.load(open(f)); print(f[15:-5], round(d["value"]), round(d["e2e"]["value"]), round(d["regions_per_s"]), {k: round(v,1) for k,v in d["kernel_ms"].items()})
    except Exception as e:
        print(f, "FAILED", e)
PY
