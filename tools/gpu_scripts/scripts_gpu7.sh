#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/determinism_probe.py tiny fp16 > gpurun_out/determinism.log 2>&1
echo "det $?" > gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -30 gpurun_out/determinism.log
