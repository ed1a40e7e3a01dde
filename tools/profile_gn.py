"""k_gn_apply on the dominant shape (192 channels @ 256x256, B=4, fp16 operand out) for ncu."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from humanliff_b200._lib import call  # noqa: E402

dev = torch.device("cuda:0")
B, HW, C = 4, 256 * 256, 192
x = torch.randn(B, HW, C, device=dev)
y = torch.empty(B, HW, C, device=dev, dtype=torch.float16)
gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)
film = torch.zeros(B, 2 * C, device=dev)
stats = torch.zeros(B * C * 2, device=dev, dtype=torch.float64)
st = torch.cuda.current_stream(dev).cuda_stream
call("hl_gn_stats", x.data_ptr(), C, B, HW, C, stats.data_ptr(), C, st)
for _ in range(5):
    call("hl_gn_apply", x.data_ptr(), C, stats.data_ptr(), C, gamma.data_ptr(), beta.data_ptr(), film.data_ptr(), 2 * C,
         y.data_ptr(), 1, C, None, 0, B, HW, C, 32, 1e-5, 1, 0, st)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    call("hl_gn_apply", x.data_ptr(), C, stats.data_ptr(), C, gamma.data_ptr(), beta.data_ptr(), film.data_ptr(), 2 * C,
         y.data_ptr(), 1, C, None, 0, B, HW, C, 32, 1e-5, 1, 0, st)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print("gn_apply 192@256^2 B=4: %.1f us, %.0f GB/s" % (ms * 1e3, B * HW * C * 6 / ms / 1e6))
