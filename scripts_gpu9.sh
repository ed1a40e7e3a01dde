#!/bin/bash
mkdir -p gpurun_out
timeout 600 python - > gpurun_out/tc_probe4.log 2>&1 <<'PY'
import sys
sys.argv = ["tc_probe.py"]
sys.path.insert(0, "tools")
import tc_probe
tc_probe.drive("timing")
PY
echo "probe $?" > gpurun_out/summary.txt
timeout 600 python -m pytest -q --timeout=120 -p no:cacheprovider tests/test_kernels_gpu.py -m gpu -k "groupnorm or attention" > gpurun_out/t_k.log 2>&1; echo "tests $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
cat gpurun_out/tc_probe4.log
tail -3 gpurun_out/t_k.log
