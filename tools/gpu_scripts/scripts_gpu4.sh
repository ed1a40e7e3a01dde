#!/bin/bash
mkdir -p gpurun_out
timeout 1500 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_step.csv python tools/profile_step.py > gpurun_out/profile_step.log 2>&1
echo "ncu $?"
python tools/summarize_launches.py gpurun_out/launches_step.csv | tee gpurun_out/step_breakdown.md
