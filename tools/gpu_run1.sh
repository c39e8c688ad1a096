#!/bin/bash
# first GPU pass of a build: new-geometry tests first, then the whole -m gpu suite, then the bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/r1_smi.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "band_widths or noisy_pairs or identical" > gpurun_out/r1_new_tests.log 2>&1
echo "new tests exit $?" >> gpurun_out/r1_new_tests.log
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r1_all_tests.log 2>&1
echo "all tests exit $?" >> gpurun_out/r1_all_tests.log
timeout 600 python bench.py --steps 6 --warmup 3 > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err
echo "bench exit $?" >> gpurun_out/r1_bench.err
tail -3 gpurun_out/r1_new_tests.log; tail -5 gpurun_out/r1_all_tests.log; cat gpurun_out/r1_bench.json | head -c 3000
