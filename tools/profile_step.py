"""One production p_sample step (B=4, 27x256x256) bracketed by cudaProfilerStart/Stop, for
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv ...
Never a bench number: ncu serialises launches and runs them cold-cache."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

B = int(os.environ.get("HL_PROFILE_BATCH", "4"))
dev = torch.device("cuda:0")
model, diffusion, _ = bench.build_model(dev)
g = torch.Generator().manual_seed(0)
x = torch.randn(B, 27, 256, 256, generator=g).to(dev)
xc = torch.zeros_like(x)
z = torch.randn(B, 27, 256, 256, generator=g).to(dev)
y = (torch.arange(B) % 4).to(dev)
t = torch.full((B,), 500, dtype=torch.int64, device=dev)
for _ in range(2):
    diffusion.p_sample(model, x, xc, t, model_kwargs={"y": y}, noise=z)
torch.cuda.synchronize()
torch.cuda.profiler.start()
diffusion.p_sample(model, x, xc, t, model_kwargs={"y": y}, noise=z)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
