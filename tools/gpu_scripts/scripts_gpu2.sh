#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/tc_debug.py > gpurun_out/tc_debug.log 2>&1
cat gpurun_out/tc_debug.log
