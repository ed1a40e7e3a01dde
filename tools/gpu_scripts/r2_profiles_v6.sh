#!/bin/bash
# round-2 final evidence pass (tag $1): everything r2_profiles_v3.sh collects + the canonical-space probe and its ncu capture
V=${1:-v6}
bash tools/gpu_scripts/r2_profiles_v3.sh $V
timeout 300 python tools/canon_probe.py > gpurun_out/r2_canon_probe_$V.log 2>&1; echo "canon probe $?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_render_tc5 -s 1 -c 1 -f -o gpurun_out/r2_canon_tc5_$V python tools/canon_ncu.py > gpurun_out/r2_ncu_canon.log 2>&1; echo "ncu canon $?"
tail -12 gpurun_out/r2_canon_probe_$V.log
