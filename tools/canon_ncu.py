"""One canonical-space 256 x 256 frame through the tcgen05 render kernel (for `ncu -k regex:k_render_tc5`)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from humanliff_b200 import synth  # noqa: E402
from humanliff_b200.renderer import Renderer, render  # noqa: E402

dev = torch.device("cuda:0")
asset = synth.synth_smpl(5)
r = Renderer(use_canonical_space=True, triplane_ch=27, test=True, smpl=asset, precision="fp16")
synth.randomize_(r, seed=3, weight_gain=1.5)
r.to(dev)
tp = synth.synth_canonical_frame(asset, 21)
mv = lambda v: {k: mv(x) for k, x in v.items()} if isinstance(v, dict) else v.to(dev)
tpd = mv(tp)
planes = synth.synth_triplane(256, seed=7).to(dev)
ro, rd, near, far, hit = synth.synth_camera_rays(256, 256, focal=300.0, azimuth_deg=30.0, bounds=tp["world_bounds"][0].tolist())
for _ in range(2):
    render(rays_o=ro[None].to(dev), rays_d=rd[None].to(dev), near=near[None].to(dev), far=far[None].to(dev),
           tri_planes=planes, tp_input=tpd, renderer=r, n_samples=128, n_importance=128)
torch.cuda.synchronize()
