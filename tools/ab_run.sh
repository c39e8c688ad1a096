#!/bin/bash
python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "serial or pipelined or multiple_waves" > gpurun_out/t_ser2.log 2>&1; tail -2 gpurun_out/t_ser2.log
b() { name=$1; shift; python bench.py --no-cpu-baseline "$@" > gpurun_out/ab6_$name.json 2>gpurun_out/ab6_$name.err; }
b c2_pipe --steps 6 --warmup 3
b config4_pipe --workload config4 --steps 4 --warmup 3
SVP_TB_SERIAL_LEN=1000000 b c2_pipe_serial2 --steps 6 --warmup 3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ab6_c2_n2.json 2> gpurun_out/ab6_c2_n2.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/ab6_*.json")):
    try:
        d=json.load(open(f)); print(f[15:-5], d["n_gpus"], round(d["value"]), round(d["e2e"]["value"]), round(d["regions_per_s"]), {k: round(v,1) for k,v in d["kernel_ms"].items()})
    except Exception as e:
        print(f, "FAILED", e)
PY
