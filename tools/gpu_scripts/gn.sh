#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/profile_gn.py > gpurun_out/gn_time.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gn_apply -s 3 -c 1 -o gpurun_out/gn_full -f python tools/profile_gn.py > gpurun_out/ncu_gn.log 2>&1
cat gpurun_out/gn_time.log
tail -2 gpurun_out/ncu_gn.log
