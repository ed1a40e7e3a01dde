#!/bin/bash
# full validation + headline bench
mkdir -p gpurun_out
timeout 600 python -m pytest -q --timeout=120 -p no:cacheprovider tests -m gpu > gpurun_out/t_all.log 2>&1; echo "tests $?"
tail -4 gpurun_out/t_all.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-render > gpurun_out/bench_cur.json 2> gpurun_out/bench_cur.err; echo "bench $?"
tail -3 gpurun_out/bench_cur.err
python -c "
import json; d=json.load(open('gpurun_out/bench_cur.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved'], d['gpu_launches'])"
