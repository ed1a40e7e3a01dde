"""One production p_sample step (B=4, 27x256x256) between cudaProfilerStart/Stop, for
`ncu --profile-from-start off`.  The step is launched EAGERLY (no CUDA graph) so every kernel is visible."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from bench import build_model  # noqa: E402

dev = torch.device("cuda:0")
B = int(os.environ.get("HL_BATCH", "4"))
model, diffusion, _ = build_model(dev, os.environ.get("HL_PRECISION", "fp16"))
model.use_cuda_graph = False
g = torch.Generator().manual_seed(0)
x = torch.randn(B, 27, 256, 256, generator=g).to(dev)
xc = torch.zeros_like(x)
z = torch.randn(B, 27, 256, 256, generator=g).to(dev)
y = (torch.arange(B) % 4).to(dev)
t = torch.full((B,), 500, dtype=torch.int64, device=dev)
for _ in range(2):
    diffusion.p_sample(model, x, xc, t, model_kwargs={"y": y}, noise=z)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
diffusion.p_sample(model, x, xc, t, model_kwargs={"y": y}, noise=z)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
