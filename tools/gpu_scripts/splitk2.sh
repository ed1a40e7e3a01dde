#!/bin/bash
mkdir -p gpurun_out
export HL_ABLATE_FULL_ONLY=1
for i in 1 2; do
for sk in 0 1; do
  echo "splitk=$sk: $(HL_SPLITK=$sk timeout 200 python tools/ablate_step.py 2>&1 | tail -1)"
done
done
./tools/gpu_scripts/step.sh
