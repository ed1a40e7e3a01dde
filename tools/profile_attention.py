"""k_attention_mma on the dominant attention shape (C=384, T=1024, 4 heads, B=4, fp16 qkv) for ncu."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from humanliff_b200._lib import call  # noqa: E402

dev = torch.device("cuda:0")
B, T, C, heads = 4, 1024, 384, 4
qkv = torch.randn(B, T, 3 * C, device=dev).half()
out = torch.empty(B, T, C, device=dev, dtype=torch.float16)
st = torch.cuda.current_stream(dev).cuda_stream
for _ in range(6):
    call("hl_attention", qkv.data_ptr(), 1, 3 * C, out.data_ptr(), 1, C, B, T, C, heads, 0, st)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    call("hl_attention", qkv.data_ptr(), 1, 3 * C, out.data_ptr(), 1, C, B, T, C, heads, 0, st)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print("attention C=384 T=1024 B=4: %.1f us, %.1f TFLOP/s" % (ms * 1e3, 4.0 * B * T * T * C / ms / 1e9))
