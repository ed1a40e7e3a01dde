#!/bin/bash
# round-2 closing pass (tag v8) after the canonical-search changes: smoke, all GPU tests, default bench, ncu --set full of
# the canonical-space render launch
V=${1:-v8}
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke_$V.log 2>&1; echo "smoke $?"; tail -2 gpurun_out/r2_smoke_$V.log
timeout 900 python -m pytest -q --timeout=200 -p no:cacheprovider tests -m gpu > gpurun_out/r2_tests_$V.log 2>&1; echo "tests $?"
tail -2 gpurun_out/r2_tests_$V.log
timeout 1500 python bench.py > gpurun_out/r2_bench_n1_$V.json 2> gpurun_out/r2_bench_n1_$V.err; echo "bench $?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_render_tc5 -s 1 -c 1 -f -o gpurun_out/r2_canon_tc5_$V python tools/canon_ncu.py > gpurun_out/r2_ncu_canon.log 2>&1; echo "ncu canon $?"
cut -c1-700 gpurun_out/r2_bench_n1_$V.json
