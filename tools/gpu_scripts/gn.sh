#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/profile_gn.py > gpurun_out/gn_time.log 2>&1
cat gpurun_out/gn_time.log
timeout 600 python -m pytest -q --timeout=120 -p no:cacheprovider tests -m gpu > gpurun_out/t_all.log 2>&1; echo "tests $?"
tail -4 gpurun_out/t_all.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v8.json 2> gpurun_out/bench_v8.err; echo "bench $?"
cat gpurun_out/bench_v8.json | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/bench_v8.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved'], d['render'])"
