#!/bin/bash
# round-1 session-3 first contact: tcgen05 conv v2 probe (parity per tiling variant + timing), then tests + bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python tools/tc_probe.py --time > gpurun_out/tc_probe.log 2>&1
echo "probe $?" > gpurun_out/summary.txt
P="timeout 900 python -m pytest -q --timeout=600 -p no:cacheprovider -x"
$P tests/test_kernels_gpu.py -m gpu -k "not tensor_core and not conv_tc" > gpurun_out/t1_kernels.log 2>&1; echo "t1 $?" >> gpurun_out/summary.txt
timeout 900 python -m pytest -q --timeout=600 -p no:cacheprovider tests/test_kernels_gpu.py -m gpu -k "tensor_core or conv_tc" > gpurun_out/t2_tc.log 2>&1; echo "t2 $?" >> gpurun_out/summary.txt
timeout 900 python -m pytest -q --timeout=600 -p no:cacheprovider tests/test_unet_gpu.py -m gpu > gpurun_out/t3_unet.log 2>&1; echo "t3 $?" >> gpurun_out/summary.txt
$P tests/test_render_gpu.py -m gpu > gpurun_out/t5_render.log 2>&1; echo "t5 $?" >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_v2.json 2> gpurun_out/bench_v2.err; echo "bench $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
cat gpurun_out/tc_probe.log
tail -n 25 gpurun_out/t2_tc.log
tail -n 25 gpurun_out/t3_unet.log
tail -n 5 gpurun_out/t1_kernels.log gpurun_out/t5_render.log gpurun_out/bench_v2.err
cat gpurun_out/bench_v2.json
