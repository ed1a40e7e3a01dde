"""Per-phase cycles of ONE ray group running alone on its SM (n_rays = number of SMs: the second group of every CTA
idles) vs both groups busy: separates what a group loses to its own latencies from what it loses to sharing the SM."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from humanliff_b200 import _lib, synth  # noqa: E402
from humanliff_b200.renderer import Renderer  # noqa: E402

dev = torch.device("cuda:0")
r = Renderer(triplane_ch=27, test=True, precision="fp16")
synth.randomize_(r, seed=3, weight_gain=1.5)
r = r.to(dev)
planes = synth.synth_triplane(256, seed=7)[0].to(dev)
bounds = torch.tensor(synth.WORLD_BOUNDS)
ro, rd, near, far, _ = synth.synth_camera_rays(512, 512, focal=600.0, azimuth_deg=30.0)
lib = _lib.load()
names = ["setup", "gather", "mlp", "resample+sort", "composite", "total"]
for label, n, per in (("one group per SM", 148, 1), ("two groups per SM", 296, 1), ("steady state", 148 * 2 * 40, 40)):
    sel = slice(512 * 200, 512 * 200 + n)
    a = [t[sel].contiguous().to(dev) for t in (ro, rd, near, far)]
    r.render_rays(planes, bounds, *a, u=None, seed=1)
    prof = torch.zeros(8, device=dev, dtype=torch.int64)
    lib.hl_render5_set_profile(prof.data_ptr())
    r.render_rays(planes, bounds, *a, u=None, seed=2)
    torch.cuda.synchronize()
    lib.hl_render5_set_profile(None)
    p = prof.cpu().tolist()
    print(label, {k: round(v / per) for k, v in zip(names, p)})
