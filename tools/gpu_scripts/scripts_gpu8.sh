#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest -q --timeout=120 -p no:cacheprovider tests -m gpu > gpurun_out/t_all.log 2>&1; echo "tests $?" > gpurun_out/summary.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_v4.json 2> gpurun_out/bench_v4.err; echo "bench $?" >> gpurun_out/summary.txt
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_step_v4.csv python tools/profile_step.py > gpurun_out/profile_step.log 2>&1
echo "ncu $?" >> gpurun_out/summary.txt
python tools/summarize_launches.py gpurun_out/launches_step_v4.csv > gpurun_out/step_breakdown_v4.md 2>&1
cat gpurun_out/summary.txt
tail -n 8 gpurun_out/t_all.log
cat gpurun_out/step_breakdown_v4.md
tail -3 gpurun_out/bench_v4.err
cat gpurun_out/bench_v4.json
