#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest -q -x --timeout=60 -p no:cacheprovider tests/test_kernels_gpu.py -m gpu -k "split_k" > gpurun_out/t_split.log 2>&1; rc=$?; echo "split tests $rc"
tail -15 gpurun_out/t_split.log
[ $rc -ne 0 ] && exit 1
./tools/gpu_scripts/step.sh
HL_ABLATE_FULL_ONLY=1 timeout 200 python tools/ablate_step.py 2>&1 | tail -1
