#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke $?"
tail -2 gpurun_out/smoke.log
timeout 600 python tools/loop_check.py 50 > gpurun_out/loop.log 2>&1; echo "loop $?"
tail -3 gpurun_out/loop.log
timeout 600 python -m pytest -q --timeout=120 -p no:cacheprovider tests -m gpu > gpurun_out/t_all.log 2>&1; echo "tests $?"
tail -3 gpurun_out/t_all.log
