"""The dominant launch of the denoise step in isolation (conv3x3 192->192 @ 256x256, B=4, residual + GroupNorm
statistics, automatic tiling) for `ncu --set full -k regex:k_conv_tc -s 3 -c 1`."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from humanliff_b200._lib import call  # noqa: E402
from humanliff_b200.unet import pack_conv  # noqa: E402

dev = torch.device("cuda:0")
B, HW, Cin, Cout = 4, 256, 192, 192
g = torch.Generator().manual_seed(0)
x = torch.randn(B, HW, HW, Cin, device=dev).half()
r = torch.randn(B, HW, HW, Cout, device=dev)
y = torch.empty(B, HW, HW, Cout, device=dev)
st = torch.zeros(B * Cout * 2, device=dev, dtype=torch.float64)
w = torch.randn(Cout, Cin, 3, 3, generator=g) / 41.6
wpk, bpk = pack_conv(w, torch.zeros(Cout), Cin, "fp16", dev)
stream = torch.cuda.current_stream(dev).cuda_stream
for _ in range(6):
    call("hl_conv2d", x.data_ptr(), 1, Cin, wpk.data_ptr(), bpk.data_ptr(), r.data_ptr(), Cout, y.data_ptr(), Cout,
         st.data_ptr(), Cout, B, HW, HW, Cin, Cout, 3, 1, 0, stream)
torch.cuda.synchronize()
