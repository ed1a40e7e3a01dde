#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest -q --timeout=120 -p no:cacheprovider tests/test_render_gpu.py -m gpu > gpurun_out/t_render.log 2>&1; echo "tests $?"
tail -6 gpurun_out/t_render.log
timeout 300 python tools/render_probe.py > gpurun_out/render_probe.log 2>&1
cat gpurun_out/render_probe.log | cut -c1-330
