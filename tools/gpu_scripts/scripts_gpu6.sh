#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/tc_probe.py --time > gpurun_out/tc_probe2.log 2>&1
echo "probe $?" > gpurun_out/summary.txt
if grep -q '"name": "tap1_f16", "nan": false' gpurun_out/tc_probe2.log; then
timeout 600 python -m pytest -q --timeout=120 -p no:cacheprovider tests/test_kernels_gpu.py tests/test_unet_gpu.py -m gpu > gpurun_out/t_all.log 2>&1; echo "tests $?" >> gpurun_out/summary.txt
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_step_v2.csv python tools/profile_step.py > gpurun_out/profile_step.log 2>&1
echo "ncu $?" >> gpurun_out/summary.txt
python tools/summarize_launches.py gpurun_out/launches_step_v2.csv > gpurun_out/step_breakdown_v2.md 2>&1
fi
cat gpurun_out/summary.txt
cat gpurun_out/tc_probe2.log
tail -n 12 gpurun_out/t_all.log
cat gpurun_out/step_breakdown_v2.md
tail -5 gpurun_out/profile_step.log
