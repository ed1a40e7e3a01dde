import sys, time, torch
sys.path.insert(0, '/root/repo')
import bench
from humanliff_b200 import synth
dev = torch.device("cuda:0")
r, planes, bounds, rays = bench._render_setup(dev)
tp = {"world_bounds": bounds[None].to(dev)}
pl = planes.to(dev)
for res in (128, 256, 512, 512, 512):
    torch.cuda.synchronize(); t0 = time.time()
    g = r.density_grid(tp, pl, resolution=res)
    torch.cuda.synchronize(); dt = time.time() - t0
    print("density grid %d^3: %.1f ms, %.2f G points/s" % (res, dt * 1e3, res ** 3 / dt / 1e9), flush=True)
    del g
