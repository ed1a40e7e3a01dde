#!/bin/bash
# same-box A/B of the unroll factor of the nearest-vertex search loops (builds: libhumanliff_b200{,_cu1,_cu4}.so)
mkdir -p gpurun_out
timeout 300 python -m pytest -q --timeout=200 -p no:cacheprovider tests -m gpu -k canon 2>&1 | tail -2
: > gpurun_out/r2_canon_unroll_ab.log
for rep in 1 2; do for tag in "" _cu1 _cu4; do
  echo "== lib$tag (rep $rep)" >> gpurun_out/r2_canon_unroll_ab.log
  HL_LIB=$PWD/humanliff_b200/libhumanliff_b200$tag.so timeout 200 python tools/canon_probe.py 2>&1 | grep -v "fp32\]" >> gpurun_out/r2_canon_unroll_ab.log
done; done
grep -E "==|512 x 512|per call" gpurun_out/r2_canon_unroll_ab.log
