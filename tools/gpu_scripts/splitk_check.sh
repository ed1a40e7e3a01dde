#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest -q -x --timeout=60 -p no:cacheprovider tests/test_kernels_gpu.py tests/test_fullsize_gpu.py -m gpu -k "split_k or spot_check" > gpurun_out/t_split.log 2>&1; rc=$?; echo "split tests $rc"
tail -6 gpurun_out/t_split.log
[ $rc -ne 0 ] && exit 1
./tools/gpu_scripts/step.sh
