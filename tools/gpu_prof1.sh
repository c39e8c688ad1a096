#!/bin/bash
# per-variant efficiency of a build: ncu launch list (timing only) + --set full of the forward / traceback kernels of one small step
cd "$(dirname "$0")/.."
O=gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --regions 24 --distinct-batches 1 --no-cpu-baseline --no-pipeline"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/p1_launches.csv $CMD > $O/p1_ncu_l.log 2>&1
echo "launch list exit $?"
CMD2="python bench.py --steps 1 --warmup 1 --regions 8 --distinct-batches 1 --no-cpu-baseline --no-pipeline"
SVP_ARENA_GB=40 timeout 400 ncu --set full --clock-control none --import-source on -k regex:svp_fwd -c 14 -o /tmp/prof_fwd -f $CMD2 > $O/p1_ncu_f.log 2>&1
echo "fwd full exit $?"
python tools/ncu_summary.py /tmp/prof_fwd.ncu-rep "SVP_ARENA_GB=40 ncu --set full --clock-control none --import-source on -k regex:svp_fwd -c 14 $CMD2" > $O/p1_ncu_fwd.json 2> $O/p1_sum.err
SVP_ARENA_GB=40 timeout 300 ncu --set full --clock-control none --import-source on -k regex:svp_tb -c 2 -o /tmp/prof_tb -f $CMD2 > $O/p1_ncu_t.log 2>&1
echo "tb full exit $?"
python tools/ncu_summary.py /tmp/prof_tb.ncu-rep "SVP_ARENA_GB=40 ncu --set full --clock-control none --import-source on -k regex:svp_tb -c 2 $CMD2" > $O/p1_ncu_tb.json 2>> $O/p1_sum.err
du -sh $O
