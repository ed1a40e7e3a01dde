#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/determinism_probe.py tiny fp16 > gpurun_out/determinism.log 2>&1
echo "det $?" > gpurun_out/summary.txt
timeout 600 python tools/tc_probe.py --time > gpurun_out/tc_probe3.log 2>&1
echo "probe $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -20 gpurun_out/determinism.log
grep -v '"nan": false' gpurun_out/tc_probe3.log
