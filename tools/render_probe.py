"""Phase breakdown (cycles of CTA 0 / group 0) and throughput of the render kernels on the 512x512 synthetic camera.
    python tools/render_probe.py [fp16|fp16_mma]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from humanliff_b200 import _lib  # noqa: E402

precision = sys.argv[1] if len(sys.argv) > 1 else "fp16"
dev = torch.device("cuda:0")
print(json.dumps(bench.render_throughput(dev, precision=precision)))
prof = torch.zeros(8, device=dev, dtype=torch.int64)
lib = _lib.load()
hook = lib.hl_render5_set_profile if precision == "fp16" else lib.hl_render_set_profile
hook(prof.data_ptr())
bench.render_throughput(dev, reps=1, precision=precision)
torch.cuda.synchronize()
hook(None)
p = prof.cpu().tolist()
per_cta = 262144 // 148 + 1
rays = 2 * (per_cta // 2 if precision == "fp16" else per_cta)      # warm-up launch + 1 timed launch; tc5: 2 groups share a CTA's rays
names = ["setup", "gather", "mlp", "resample+sort", "composite", "total", "mlp:barrier+issue", "mlp:wait_mma"]
print({n: round(v / rays) for n, v in zip(names, p)}, "cycles per ray (CTA 0%s)" % (", group 0" if precision == "fp16" else ""))
