#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/tc_probe.py > gpurun_out/tc_cases.log 2>&1
grep -c '"nan": false' gpurun_out/tc_cases.log
grep -v '"nan": false' gpurun_out/tc_cases.log | head -5
grep -c '"stats": "ok"' gpurun_out/tc_cases.log
./tools/gpu_scripts/probe_timing.sh
timeout 300 python tools/determinism_probe.py tiny fp16 2>&1 | grep -E "nondet|end-to-end|calls:"
