"""One launch of the render kernel on a bounded ray block (for `ncu --set full`; a number printed under ncu is
never a bench value).  python tools/render_ncu.py [fp16|fp16_mma] [n_rays]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from humanliff_b200 import synth  # noqa: E402
from humanliff_b200.renderer import Renderer  # noqa: E402

precision = sys.argv[1] if len(sys.argv) > 1 else "fp16"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 148 * 2 * 24
dev = torch.device("cuda:0")
r = Renderer(triplane_ch=27, test=True, precision=precision)
synth.randomize_(r, seed=3, weight_gain=1.5)
r = r.to(dev)
planes = synth.synth_triplane(256, seed=7)[0].to(dev)
bounds = torch.tensor(synth.WORLD_BOUNDS)
ro, rd, near, far, _ = synth.synth_camera_rays(512, 512, focal=600.0, azimuth_deg=30.0)
sel = slice(512 * 192, 512 * 192 + n)
ro, rd, near, far = (t[sel].contiguous().to(dev) for t in (ro, rd, near, far))
for i in range(2):
    r.render_rays(planes, bounds, ro, rd, near, far, u=None, seed=1 + i)
torch.cuda.synchronize()
print("rendered", n, "rays with", precision)
