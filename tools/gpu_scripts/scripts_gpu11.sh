#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest -q --timeout=120 -p no:cacheprovider tests -m gpu > gpurun_out/t_all.log 2>&1; echo "tests $?" > gpurun_out/summary.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_v7.json 2> gpurun_out/bench_v7.err; echo "bench $?" >> gpurun_out/summary.txt
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_step_v7.csv python tools/profile_step.py > gpurun_out/profile_step.log 2>&1
echo "ncu $?" >> gpurun_out/summary.txt
python tools/summarize_launches.py gpurun_out/launches_step_v7.csv > gpurun_out/step_breakdown_v7.md 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_conv_tc -s 3 -c 1 -o gpurun_out/conv_full_v7 -f python tools/profile_conv.py > gpurun_out/ncu_full.log 2>&1
echo "ncufull $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -4 gpurun_out/t_all.log
cat gpurun_out/step_breakdown_v7.md
cat gpurun_out/bench_v7.json | cut -c1-2600
tail -3 gpurun_out/ncu_full.log
