"""Debug harness for the tcgen05 conv: each shape in its own process; prints error structure."""
import subprocess
import sys
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHAPES = [(1, 16, 8, 32, 32, 1), (1, 16, 8, 32, 32, 3), (1, 64, 64, 192, 192, 3), (2, 32, 32, 384, 384, 3),
          (4, 8, 8, 768, 768, 3), (1, 128, 128, 32, 192, 3), (1, 256, 256, 64, 64, 3), (2, 16, 16, 384, 192, 1),
          (1, 64, 64, 192, 27, 3), (3, 8, 8, 64, 32, 3)]

CHILD = r'''
import sys, math, torch
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/tests")
import torch.nn.functional as F
from humanliff_b200 import _lib
from humanliff_b200._lib import call
from humanliff_b200.unet import pack_conv
B, H, W, Cin, Cout, k = map(int, sys.argv[2:8])
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(7)
x = torch.randn(B, H, W, Cin, generator=g).to(dev)
w = torch.randn(Cout, Cin, k, k, generator=g) / math.sqrt(Cin * k * k)
b = torch.randn(Cout, generator=g) * 0.1
st = torch.cuda.current_stream().cuda_stream
call("hl_round_tf32", x.data_ptr(), Cin, x.data_ptr(), Cin, Cin, B * H * W, st)
wpk, bpk = pack_conv(w, b, Cin, True, dev)
y1 = torch.full((B, H, W, Cout), float("nan"), device=dev)
y2 = torch.full((B, H, W, Cout), float("nan"), device=dev)
uses = _lib.load().hl_conv2d_uses_tensor_cores(B, H, W, Cin, Cout, k, 1, Cin, 0)
call("hl_conv2d", x.data_ptr(), Cin, wpk.data_ptr(), bpk.data_ptr(), None, 0, y1.data_ptr(), Cout, B, H, W, Cin, Cout, k, 1, 0, st)
call("hl_conv2d", x.data_ptr(), Cin, wpk.data_ptr(), bpk.data_ptr(), None, 0, y2.data_ptr(), Cout, B, H, W, Cin, Cout, k, 1, 1, st)
torch.cuda.synchronize()
err = float((y1 - y2).double().norm() / y2.double().norm())
nan = int(torch.isnan(y1).sum())
print(f"shape {(B,H,W,Cin,Cout,k)} tc={uses} rel_err={err:.3e} nan={nan}", flush=True)
if not (err < 1e-4):
    d = (y1 - y2).abs().nan_to_num(1e3).reshape(-1, Cout)
    print("  err by pixel row (first 16):", [round(float(v), 3) for v in d.mean(1)[:16]])
    print("  err by pixel row%8:", [round(float(d[i::8].mean()), 3) for i in range(8)])
    print("  err by channel (first 16):", [round(float(v), 3) for v in d.mean(0)[:16]])
    print("  y1[0,:8]:", [round(float(v), 3) for v in y1.reshape(-1, Cout)[0, :8]])
    print("  y2[0,:8]:", [round(float(v), 3) for v in y2.reshape(-1, Cout)[0, :8]])
    # does y1 equal the conv with only a subset of taps / channels?
    xn = x.permute(0, 3, 1, 2).cpu()
    wr = wpk[:, :Cout].reshape(k, k, Cout, Cin).permute(2, 3, 0, 1).cpu()
    for kc in range(0, Cin, 32):
        part = F.conv2d(xn[:, kc:kc+32], wr[:, kc:kc+32], padding=k // 2).permute(0, 2, 3, 1)
        r = float(((y1.cpu() - b) - part).double().norm() / part.double().norm())
        print(f"  vs channels {kc}..{kc+32} only: {r:.3e}")
        if kc >= 64: break
'''

if __name__ == "__main__":
    for s in SHAPES:
        p = subprocess.run([sys.executable, "-c", CHILD, ROOT] + [str(v) for v in s], capture_output=True, text=True, timeout=300)
        print(p.stdout.strip())
        if p.returncode != 0:
            print("  EXIT", p.returncode, p.stderr.strip()[-600:])
