"""k_render_tc on one 16,384-ray chunk of the 512x512 synthetic camera (in-kernel uniforms) for ncu."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
from common import renderer_state_dict  # noqa: E402
from humanliff_b200 import synth  # noqa: E402

dev = torch.device("cuda:0")
r, _ = renderer_state_dict(3, "fp16")
r = r.to(dev)
planes = synth.synth_triplane(256, seed=7).to(dev)
bounds = torch.tensor(synth.WORLD_BOUNDS)
ro, rd, near, far, hit = synth.synth_camera_rays(512, 512, focal=600.0, azimuth_deg=45.0)
n = 65536
sl = slice(96 * 512, 96 * 512 + n)           # a band through the middle of the image (hits and misses)
args = [t[sl].to(dev) for t in (ro, rd, near, far)]
for _ in range(3):
    out = r.render_rays(planes[0], bounds, *args, u=None, seed=1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
out = r.render_rays(planes[0], bounds, *args, u=None, seed=1)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print("render %d rays: %.2f ms, %.2f M rays/s" % (n, ms, n / ms / 1e3))
