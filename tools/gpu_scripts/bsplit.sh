#!/bin/bash
mkdir -p gpurun_out
for sp in 1 2 1 2; do
  HL_BATCH_SPLIT=$sp timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-render | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('split=$sp', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"
done
HL_BATCH_SPLIT=2 timeout 300 python -m pytest -q -x --timeout=200 -p no:cacheprovider tests/test_fullsize_gpu.py tests/test_unet_gpu.py -m gpu 2>&1 | tail -3
