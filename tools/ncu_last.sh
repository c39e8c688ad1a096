CMD="python bench.py --steps 1 --warmup 1 --regions 8 --distinct-batches 1 --no-cpu-baseline --no-pipeline"
SVP_ARENA_GB=40 timeout 170 ncu --set full --clock-control none --import-source on -k regex:svp_fwd -c 10 -o /tmp/prof_fwd -f $CMD > gpurun_out/ncu_f6.log 2>&1
tail -2 gpurun_out/ncu_f6.log | cut -c1-300
python tools/ncu_summary.py /tmp/prof_fwd.ncu-rep "SVP_ARENA_GB=40 ncu --set full --clock-control none --import-source on -k regex:svp_fwd -c 10 $CMD" > gpurun_out/ncu_fwd_r01_v6.json 2> gpurun_out/ncu_fwd_sum.err
grep -c '"kernel"' gpurun_out/ncu_fwd_r01_v6.json
