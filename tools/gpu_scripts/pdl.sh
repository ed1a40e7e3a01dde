#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest -q -x --timeout=100 -p no:cacheprovider tests/test_unet_gpu.py -m gpu > gpurun_out/t_unet.log 2>&1; rc=$?; echo "unet tests $rc"
tail -5 gpurun_out/t_unet.log
[ $rc -ne 0 ] && exit 1
HL_ABLATE_FULL_ONLY=1 HL_PDL=0 timeout 200 python tools/ablate_step.py 2>&1 | head -1
HL_ABLATE_FULL_ONLY=1 HL_PDL=1 timeout 200 python tools/ablate_step.py 2>&1 | head -1
./tools/gpu_scripts/step.sh
