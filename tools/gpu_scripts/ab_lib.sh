#!/bin/bash
# same-box A/B of two builds of the library: $1 = tag of the alternative build (libhumanliff_b200_<tag>.so)
mkdir -p gpurun_out
timeout 600 python -m pytest -q --timeout=200 -p no:cacheprovider tests -m gpu > gpurun_out/t_all.log 2>&1; echo "tests $?"
tail -3 gpurun_out/t_all.log
export HL_ABLATE_FULL_ONLY=1
for i in 1 2 3; do for tag in "_$1" ""; do echo "lib$tag: $(HL_LIB=$PWD/humanliff_b200/libhumanliff_b200$tag.so timeout 200 python tools/ablate_step.py 2>&1 | tail -1)"; done; done
