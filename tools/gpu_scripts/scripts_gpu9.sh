#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/tc_probe.py --time > gpurun_out/tc_probe5.log 2>&1
echo "probe $?" > gpurun_out/summary.txt
if grep -q '"name": "tap1_f16", "nan": false' gpurun_out/tc_probe5.log; then
timeout 600 python -m pytest -q --timeout=120 -p no:cacheprovider tests -m gpu > gpurun_out/t_all.log 2>&1; echo "tests $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v5.json 2> gpurun_out/bench_v5.err; echo "bench $?" >> gpurun_out/summary.txt
fi
cat gpurun_out/summary.txt
cat gpurun_out/tc_probe5.log | cut -c1-420
tail -5 gpurun_out/t_all.log
cat gpurun_out/bench_v5.json | cut -c1-1800
