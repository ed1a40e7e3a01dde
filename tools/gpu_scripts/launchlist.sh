#!/bin/bash
# ncu launch list of one eager production step (per-kernel durations; cold-cache, serialised)
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/launches_cur.csv python tools/profile_step.py > gpurun_out/launchlist.log 2>&1; echo "ncu $?"
python tools/summarize_launches.py gpurun_out/launches_cur.csv | tee gpurun_out/launches_cur.md
