"""hl_gn_skip (fused GroupNorm-1 operand pass + 1x1 skip conv) against the two launches it replaces, on the decoder's
concat inputs of the production step (B = 4): time per launch and HBM rate."""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from humanliff_b200 import _lib  # noqa: E402
from humanliff_b200._lib import call  # noqa: E402
from humanliff_b200.unet import pack_conv  # noqa: E402

dev = torch.device("cuda:0")
st = torch.cuda.current_stream(dev).cuda_stream
B = int(os.environ.get("HL_B", "4"))
for H, Cin, Cout in ((256, 384, 192), (128, 576, 192), (128, 384, 192), (64, 768, 384), (64, 576, 384), (32, 768, 384), (32, 1152, 384), (64, 192, 384)):
    HW = H * H
    x = torch.randn(B, HW, Cin, device=dev)
    gamma, beta = torch.ones(Cin, device=dev), torch.zeros(Cin, device=dev)
    w = torch.randn(Cout, Cin, 1, 1) / math.sqrt(Cin)
    wpk, bpk = pack_conv(w, torch.zeros(Cout), Cin, "fp16", dev, mode="split")
    stats = torch.zeros(B * Cin * 2, device=dev, dtype=torch.float64)
    call("hl_gn_stats", x.data_ptr(), Cin, B, HW, Cin, stats.data_ptr(), Cin, st)
    act = torch.empty(B, HW, Cin, device=dev, dtype=torch.float16)
    raw = torch.empty(B, HW, 2 * Cin, device=dev, dtype=torch.float16)
    skip = torch.empty(B, HW, Cout, device=dev)
    mode = (_lib.OP_SPLIT | _lib.OP_SCALED) << _lib.OP_RAW_SHIFT

    def two():
        call("hl_gn_apply", x.data_ptr(), Cin, stats.data_ptr(), Cin, gamma.data_ptr(), beta.data_ptr(), None, 0, act.data_ptr(), 1,
             Cin, raw.data_ptr(), 2 * Cin, B, HW, Cin, 32, 1e-5, 1, mode, st)
        call("hl_conv2d", raw.data_ptr(), 1, 2 * Cin, wpk.data_ptr(), bpk.data_ptr(), None, 0, skip.data_ptr(), Cout, None, 0, B, H, H,
             Cin, Cout, 1, 1, _lib.CONV_SPLIT3, st)

    def fused():
        call("hl_gn_skip", x.data_ptr(), Cin, stats.data_ptr(), Cin, gamma.data_ptr(), beta.data_ptr(), act.data_ptr(), Cin,
             wpk.data_ptr(), bpk.data_ptr(), skip.data_ptr(), Cout, B, HW, Cin, Cout, 32, 1e-5, st)
    res = {}
    for name, fn in (("two launches", two), ("fused", fused)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / 20 * 1e3
    prof = torch.zeros(16, device=dev, dtype=torch.int64)
    lib = _lib.load()
    lib.hl_gn_skip_set_profile(prof.data_ptr())
    fused()
    torch.cuda.synchronize()
    lib.hl_gn_skip_set_profile(None)
    pr = prof.cpu().tolist()
    names = ["bookkeeping", "wait_x", "act", "wait_A", "split", "barrier", "mma_issue", "wait_last_mma", "epilogue"]
    tiles = max(pr[9], 1)
    print("      cycles per tile (CTA 0, thread 0): " + ", ".join("%s %d" % (n_, v // tiles) for n_, v in zip(names, pr)) + ", tiles %d" % tiles)
    nb = B * HW * (Cin * 6 + Cout * 4)
    print("%3d^2 %4d->%3d B=%d: two launches %6.1f us, fused %6.1f us (%5.0f GB/s of %d MB compulsory)" % (
        H, Cin, Cout, B, res["two launches"], res["fused"], nb / res["fused"] / 1e3, nb >> 20), flush=True)
