#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
for lp in 0 4096; do for mask in 31 17 1; do
echo "== latency_pairs $lp kp_mask $mask"
SVP_LATENCY_PAIRS=$lp SVP_KP_MASK=$mask timeout 100 python tools/quick_perf.py 1 128 2>&1 | grep -E "rep[12]"
done; done
echo "== 4 containers, default"; timeout 100 python tools/quick_perf.py 4 128 2>&1 | grep -E "rep[12]"
echo "== 4 containers, latency 4096"; SVP_LATENCY_PAIRS=4096 timeout 100 python tools/quick_perf.py 4 128 2>&1 | grep -E "rep[12]"
