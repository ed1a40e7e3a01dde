#!/bin/bash
mkdir -p gpurun_out
timeout 600 python - > gpurun_out/tc_timing.log 2>&1 <<'PY'
import sys
sys.argv = ["tc_probe.py"]
sys.path.insert(0, "tools")
import tc_probe
tc_probe.drive("timing")
PY
python - <<'PY'
import json
for l in open("gpurun_out/tc_timing.log"):
    try:
        d = json.loads(l)
        p = d.get("prof_kcycles", {})
        print("%-34s %.4f ms %7.1f TF | mma waits A %5.1f B %5.1f tmem %5.1f | epi full %6.1f res %5.1f bar %5.1f | prod A %6.1f B %6.1f / total %6.1f" % (
            d["name"], d.get("ms", 0), d.get("tflops", 0), p.get("mma_wait_A", 0), p.get("mma_wait_B", 0), p.get("mma_wait_tmem_empty", 0),
            p.get("epi_wait_tmem_full", 0), p.get("epi_wait_res", 0), p.get("epi_barrier", 0), p.get("prodA_wait_empty", 0),
            p.get("prodB_wait_empty", 0), p.get("total", 0)))
    except Exception:
        print(l.strip()[:200])
PY
