#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest -q -x --timeout=60 -p no:cacheprovider tests/test_kernels_gpu.py -m gpu -k "split_k" > gpurun_out/t_split.log 2>&1; rc=$?; echo "split tests $rc"
tail -4 gpurun_out/t_split.log
[ $rc -ne 0 ] && exit 1
timeout 600 python -m pytest -q --timeout=200 -p no:cacheprovider tests -m gpu > gpurun_out/t_all.log 2>&1; echo "tests $?"
tail -3 gpurun_out/t_all.log
timeout 300 python tools/layered_demo.py 50 1 64 2>&1 | head -1
