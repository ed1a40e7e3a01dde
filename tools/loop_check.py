"""p_sample_loop at production size through the public API (BASELINE configs[1] shape, short respacing):
wall-clock per step including everything the loop does (timestep tensors, randn_like, graph replay)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from humanliff_b200 import factory, synth  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
dev = torch.device("cuda:0")
model, diffusion = factory.create_model_and_diffusion(**factory.production_flags(str(steps)))
sd = synth.synth_state_dict({k: v.shape for k, v in model.state_dict().items()}, seed=0)
model.load_state_dict(sd)
model = model.to(dev).eval()
B = 4
y = (torch.arange(B) % 4).to(dev)
xc = torch.zeros(B, 27, 256, 256, device=dev)
torch.manual_seed(0)
out = diffusion.p_sample_loop(model, (B, 27, 256, 256), x_cond=xc, clip_denoised=True, model_kwargs={"y": y})   # warm-up (plan + graph)
torch.cuda.synchronize()
t0 = time.perf_counter()
out = diffusion.p_sample_loop(model, (B, 27, 256, 256), x_cond=xc, clip_denoised=True, model_kwargs={"y": y})
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print("p_sample_loop %d steps, B=%d: %.3f s wall, %.2f ms/step, %.1f sample-steps/s; finite=%s, |x|max=%.3f, mem %.1f GB" % (
    steps, B, dt, 1e3 * dt / steps, B * steps / dt, bool(torch.isfinite(out).all()), float(out.abs().max()),
    torch.cuda.max_memory_allocated() / 2**30))
t0 = time.perf_counter()
out = diffusion.ddim_sample_loop(model, (B, 27, 256, 256), x_cond=xc, clip_denoised=True, model_kwargs={"y": y}, eta=0.0)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print("ddim_sample_loop %d steps: %.2f ms/step; finite=%s" % (steps, 1e3 * dt / steps, bool(torch.isfinite(out).all())))
