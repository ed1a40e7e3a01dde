"""SURVEY.md 8(d) config 4 end to end on one GPU with synthetic weights: four clothing layers sampled in one
process (x_cond_k = sample_{k-1} kept in HBM, y = k), every finished tri-plane rendered at 512x512 from
`views` azimuths and its density grid evaluated -- the work triplane_sample_layered.py does per human, minus
file output and marching cubes.  Prints wall-clock per stage."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
from common import renderer_state_dict  # noqa: E402
from humanliff_b200 import factory, render, sample_all_layers, synth  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
views = int(sys.argv[2]) if len(sys.argv) > 2 else 4
grid = int(sys.argv[3]) if len(sys.argv) > 3 else 128
dev = torch.device("cuda:0")
model, diffusion = factory.create_model_and_diffusion(**factory.production_flags(str(steps)))
sd = synth.synth_state_dict({k: v.shape for k, v in model.state_dict().items()}, seed=0)
model.load_state_dict(sd)
model = model.to(dev).eval()
nerf, _ = renderer_state_dict(3, "fp16")
nerf = nerf.to(dev)
bounds = torch.tensor(synth.WORLD_BOUNDS)
tp = {"world_bounds": bounds[None].to(dev)}
torch.manual_seed(0)


def tick():
    torch.cuda.synchronize()
    return time.perf_counter()


sample_all_layers(model, diffusion, 1, num_layers=1)                  # warm-up: plan + graph for B = 1
t0 = tick()
layers = sample_all_layers(model, diffusion, 1)
t1 = tick()
print("sampling: 4 layers x %d steps, B=1: %.2f s (%.2f ms/step)" % (steps, t1 - t0, 1e3 * (t1 - t0) / (4 * steps)))
for k, (sample, labels) in enumerate(layers):
    assert torch.isfinite(sample).all() and labels.tolist() == [k]
    tri = sample[0:1].reshape(1, 3, -1, *sample.shape[-2:])           # triplane_sample_layered.py:158
    t2 = tick()
    for v in range(views):
        ro, rd, near, far, hit = synth.synth_camera_rays(512, 512, focal=600.0, azimuth_deg=360.0 * v / views)
        rgb, acc, normal, depth = render(chunk=512 * 512 // 16, rays_o=ro[None].to(dev), rays_d=rd[None].to(dev),
                                         near=near[None].to(dev), far=far[None].to(dev), tri_planes=tri, tp_input=tp,
                                         renderer=nerf, n_samples=128, perturb=0., n_importance=128)
        assert rgb.shape == (1, 512 * 512, 3) and torch.isfinite(rgb).all()
    t3 = tick()
    u = nerf.density_grid(tp, tri, resolution=grid)
    t4 = tick()
    print("layer %d: %d views 512^2 in %.3f s (%.1f ms/view incl. ray set-up + H2D), density grid %d^3 in %.1f ms, "
          "occupied %.1f %%" % (k, views, t3 - t2, 1e3 * (t3 - t2) / views, grid, 1e3 * (t4 - t3),
                                100.0 * float((u < 0).float().mean())))
