#!/bin/bash
mkdir -p gpurun_out
export HL_ABLATE_FULL_ONLY=1
for tag in "" _t0 _t2; do
  for pdl in 0 1; do
    echo "lib$tag pdl=$pdl: $(HL_LIB=$PWD/humanliff_b200/libhumanliff_b200$tag.so HL_PDL=$pdl timeout 200 python tools/ablate_step.py 2>&1 | tail -1)"
  done
done
