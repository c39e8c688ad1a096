#!/bin/bash
# Round-end evidence run (under gpurun): bench lines first (no profiler), then the ncu launch list and --set full captures of the
# same commands. The reports are summarised on the box (tools/ncu_summary.py) and deleted: gpurun_out/ must stay below 64 MiB.
set -x
O=gpurun_out
timeout 300 python bench.py > $O/final_bench_c2.json 2> $O/final_bench_c2.err
timeout 200 python bench.py --workload config3 --no-cpu-baseline > $O/final_bench_c3.json 2> $O/final_bench_c3.err
timeout 200 python bench.py --workload config4 --no-cpu-baseline > $O/final_bench_c4.json 2> $O/final_bench_c4.err
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > $O/final_bench_ref.json 2> $O/final_bench_ref.err
timeout 100 python tools/quick_perf.py 1 > $O/final_quick1.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_r01_v5.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_l5.log 2>&1
CMD="python bench.py --steps 1 --warmup 1 --regions 16 --distinct-batches 1 --no-cpu-baseline"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:svp_fwd -c 6 -o /tmp/prof_fwd -f $CMD > $O/ncu_f4.log 2>&1
python tools/ncu_summary.py /tmp/prof_fwd.ncu-rep "ncu --set full --clock-control none --import-source on -k regex:svp_fwd -c 6 $CMD" > $O/ncu_fwd_r01_v4.json 2> $O/ncu_fwd_sum.err
timeout 200 ncu --set full --clock-control none --import-source on -k regex:svp_tbw -c 1 -o /tmp/prof_tbw -f $CMD > $O/ncu_t3.log 2>&1
python tools/ncu_summary.py /tmp/prof_tbw.ncu-rep "ncu --set full --clock-control none --import-source on -k regex:svp_tbw -c 1 $CMD" > $O/ncu_tbw_r01_v3.json 2> $O/ncu_tbw_sum.err
# top source lines of the warp traceback by stall samples (is it the loads, the scans or the end-cell search?)
ncu -i /tmp/prof_tbw.ncu-rep --page source --csv 2>/dev/null | head -400 > $O/ncu_tbw_source_head.csv
du -sh $O
