"""Phase breakdown (cycles of CTA 0) and throughput of the tensor-core render kernel on the 512x512 synthetic camera."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from humanliff_b200 import _lib  # noqa: E402

dev = torch.device("cuda:0")
print(json.dumps(bench.render_throughput(dev)))
prof = torch.zeros(8, device=dev, dtype=torch.int64)
lib = _lib.load()
lib.hl_render_set_profile(prof.data_ptr())
bench.render_throughput(dev, reps=1)
torch.cuda.synchronize()
lib.hl_render_set_profile(None)
p = prof.cpu().tolist()
rays_cta0 = 2 * (262144 // 148 + 1)          # warm-up launch + 1 timed launch
names = ["setup", "gather", "mlp", "resample+sort", "composite", "total"]
print({n: round(v / rays_cta0) for n, v in zip(names, p)}, "cycles per ray (CTA 0)")
