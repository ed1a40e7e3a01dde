#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest -q --timeout=120 -p no:cacheprovider tests -m gpu > gpurun_out/t_all.log 2>&1; echo "tests $?" > gpurun_out/summary.txt
timeout 300 python tools/determinism_probe.py tiny fp16 > gpurun_out/determinism.log 2>&1; echo "det $?" >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_v3.json 2> gpurun_out/bench_v3.err; echo "bench $?" >> gpurun_out/summary.txt
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_step_v3.csv python tools/profile_step.py > gpurun_out/profile_step.log 2>&1
echo "ncu $?" >> gpurun_out/summary.txt
python tools/summarize_launches.py gpurun_out/launches_step_v3.csv > gpurun_out/step_breakdown_v3.md 2>&1
cat gpurun_out/summary.txt
tail -n 8 gpurun_out/t_all.log
grep -E "nondet|end-to-end|calls:" gpurun_out/determinism.log
cat gpurun_out/step_breakdown_v3.md
tail -3 gpurun_out/bench_v3.err
cat gpurun_out/bench_v3.json
