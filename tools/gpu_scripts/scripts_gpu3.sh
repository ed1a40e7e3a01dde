#!/bin/bash
mkdir -p gpurun_out
P="timeout 1200 python -m pytest -q --timeout=600 -p no:cacheprovider"
$P tests -m gpu > gpurun_out/t_all.log 2>&1; echo "tests $?" > gpurun_out/summary.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_first.json 2> gpurun_out/bench_first.err; echo "bench $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -n 15 gpurun_out/t_all.log
tail -n 5 gpurun_out/bench_first.err
cat gpurun_out/bench_first.json
