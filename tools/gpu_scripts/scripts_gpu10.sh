#!/bin/bash
mkdir -p gpurun_out
timeout 900 python - > gpurun_out/tc_sweep.log 2>&1 <<'PY'
import sys
sys.argv = ["tc_probe.py"]
sys.path.insert(0, "tools")
import tc_probe
tc_probe.drive("timing")
PY
echo "sweep $?" > gpurun_out/summary.txt
timeout 600 python -m pytest -q --timeout=120 -p no:cacheprovider tests -m gpu > gpurun_out/t_all.log 2>&1; echo "tests $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v6.json 2> gpurun_out/bench_v6.err; echo "bench $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
python - <<'PY'
import json
for l in open("gpurun_out/tc_sweep.log"):
    try:
        d = json.loads(l)
        print(d["name"], d.get("ms"), d.get("tflops"), d.get("FAILED", ""))
    except Exception:
        print(l.strip()[:200])
PY
tail -5 gpurun_out/t_all.log
cat gpurun_out/bench_v6.json | cut -c1-2200
