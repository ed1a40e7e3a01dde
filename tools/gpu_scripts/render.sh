#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest -q --timeout=120 -p no:cacheprovider tests/test_render_gpu.py -m gpu > gpurun_out/t_render.log 2>&1; echo "tests $?" > gpurun_out/summary.txt
timeout 300 python - > gpurun_out/render_bench.log 2>&1 <<'PY'
import sys, json, torch
sys.path.insert(0, ".")
import bench
from humanliff_b200 import renderer as R
dev = torch.device("cuda:0")
for prec in ("fp32", "fp16"):
    orig = R.Renderer.__init__
    def init(self, *a, **k):
        k["precision"] = prec
        orig(self, *a, **k)
    R.Renderer.__init__ = init
    try:
        print(prec, json.dumps(bench.render_throughput(dev, n_rays=65536 if prec == "fp32" else 262144)))
    finally:
        R.Renderer.__init__ = orig
PY
echo "bench $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -15 gpurun_out/t_render.log
cat gpurun_out/render_bench.log
