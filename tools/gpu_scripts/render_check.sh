#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest -q --timeout=200 -p no:cacheprovider tests -m gpu > gpurun_out/t_all.log 2>&1; echo "tests $?"
tail -4 gpurun_out/t_all.log
timeout 300 python tools/render_probe.py 2>&1 | tail -3
timeout 100 python tools/profile_attention.py; timeout 100 python tools/profile_gn.py
