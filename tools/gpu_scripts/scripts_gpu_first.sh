#!/bin/bash
# first GPU contact: isolate groups in separate processes so one trapped kernel does not poison the rest
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
P="timeout 900 python -m pytest -q --timeout=300 -p no:cacheprovider"
$P tests/test_kernels_gpu.py -m gpu -k "not tensor_core and not conv_tc" > gpurun_out/t1_kernels.log 2>&1; echo "t1 $?" >> gpurun_out/summary.txt
$P tests/test_kernels_gpu.py -m gpu -k "tensor_core or conv_tc" > gpurun_out/t2_tc.log 2>&1; echo "t2 $?" >> gpurun_out/summary.txt
$P tests/test_unet_gpu.py -m gpu -k "fp32 or injected or unconditional" > gpurun_out/t3_unet_fp32.log 2>&1; echo "t3 $?" >> gpurun_out/summary.txt
$P tests/test_unet_gpu.py -m gpu -k "tf32" > gpurun_out/t4_unet_tf32.log 2>&1; echo "t4 $?" >> gpurun_out/summary.txt
$P tests/test_render_gpu.py -m gpu > gpurun_out/t5_render.log 2>&1; echo "t5 $?" >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_first.json 2> gpurun_out/bench_first.err; echo "bench $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -3 gpurun_out/t*.log
cat gpurun_out/bench_first.json
