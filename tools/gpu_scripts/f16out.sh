#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest -q -x --timeout=60 -p no:cacheprovider tests/test_kernels_gpu.py -m gpu -k "fp16_output or attention" > gpurun_out/t_f16.log 2>&1; rc=$?; echo "f16 tests $rc"
tail -15 gpurun_out/t_f16.log
[ $rc -ne 0 ] && exit 1
./tools/gpu_scripts/step.sh
timeout 300 python tools/profile_step.py > gpurun_out/profile_step.log 2>&1; tail -40 gpurun_out/profile_step.log
