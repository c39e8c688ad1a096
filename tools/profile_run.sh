#!/bin/bash
# Round-end evidence run (under gpurun): bench lines first (no profiler), then the ncu launch list and the --set full captures
# of the same commands; everything lands in gpurun_out/ and is summarised into profiles/ by tools/ncu_summary.py.
set -x
python bench.py > gpurun_out/final_bench_c2.json 2> gpurun_out/final_bench_c2.err
python bench.py --workload config3 --no-cpu-baseline > gpurun_out/final_bench_c3.json 2> gpurun_out/final_bench_c3.err
python bench.py --workload config4 --no-cpu-baseline > gpurun_out/final_bench_c4.json 2> gpurun_out/final_bench_c4.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err
python tools/quick_perf.py 1 > gpurun_out/final_quick1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01_v5.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_l5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:svp_fwd -c 6 -o gpurun_out/prof_fwd_r01_v4 -f \
    python bench.py --steps 1 --warmup 1 --regions 16 --distinct-batches 1 --no-cpu-baseline > gpurun_out/ncu_f4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:svp_tbw -c 1 -o gpurun_out/prof_tbw_r01_v3 -f \
    python bench.py --steps 1 --warmup 1 --regions 16 --distinct-batches 1 --no-cpu-baseline > gpurun_out/ncu_t3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:svp_fwd -c 8 -o gpurun_out/prof_c3_r01_v2 -f \
    python bench.py --workload config3 --steps 1 --warmup 1 --regions 4 --distinct-batches 1 --no-cpu-baseline > gpurun_out/ncu_c3v2.log 2>&1
ls -la gpurun_out/*.ncu-rep
