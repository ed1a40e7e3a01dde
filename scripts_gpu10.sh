#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/tc_probe.py --time > gpurun_out/tc_probe6.log 2>&1
echo "probe $?" > gpurun_out/summary.txt
cat gpurun_out/summary.txt
cat gpurun_out/tc_probe6.log | cut -c1-330
