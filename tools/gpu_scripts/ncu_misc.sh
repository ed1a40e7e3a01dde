#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_gn_apply -s 3 -c 1 -o gpurun_out/gn_apply_full_v10 -f python tools/profile_gn.py > gpurun_out/ncu_gn.log 2>&1; echo "gn $?"; tail -2 gpurun_out/ncu_gn.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_attention_mma -s 3 -c 1 -o gpurun_out/attention_full_v10 -f python tools/profile_attention.py > gpurun_out/ncu_att.log 2>&1; echo "att $?"; tail -2 gpurun_out/ncu_att.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_render_tc -s 2 -c 1 -o gpurun_out/render_full_v10 -f python tools/profile_render.py > gpurun_out/ncu_render.log 2>&1; echo "render $?"; tail -2 gpurun_out/ncu_render.log
timeout 100 python tools/profile_attention.py; timeout 100 python tools/profile_gn.py; timeout 200 python tools/profile_render.py
ls -la gpurun_out/*.ncu-rep
