import math, sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from humanliff_b200 import _lib
from humanliff_b200._lib import call
from humanliff_b200.unet import pack_conv
dev = torch.device("cuda:0"); st = torch.cuda.current_stream(dev).cuda_stream
B, H, W, Cin, Cout = 2, 64, 64, 384, 192
HW = H * W
g = torch.Generator().manual_seed(1)
x = (torch.randn(B, HW, Cin, generator=g) * 3 + 0.7).to(dev)
gamma, beta = (1 + 0.1 * torch.randn(Cin, generator=g)).to(dev), (0.1 * torch.randn(Cin, generator=g)).to(dev)
w = torch.randn(Cout, Cin, 1, 1, generator=g) / math.sqrt(Cin); b = torch.randn(Cout, generator=g) * 0.1
wpk, bpk = pack_conv(w, b, Cin, "fp16", dev, mode="split")
stats = torch.zeros(B * Cin * 2, device=dev, dtype=torch.float64)
call("hl_gn_stats", x.data_ptr(), Cin, B, HW, Cin, stats.data_ptr(), Cin, st)
act_ref = torch.empty(B, HW, Cin, device=dev, dtype=torch.float16); raw = torch.empty(B, HW, 2 * Cin, device=dev, dtype=torch.float16)
mode = (_lib.OP_SPLIT | _lib.OP_SCALED) << _lib.OP_RAW_SHIFT
call("hl_gn_apply", x.data_ptr(), Cin, stats.data_ptr(), Cin, gamma.data_ptr(), beta.data_ptr(), None, 0, act_ref.data_ptr(), 1, Cin, raw.data_ptr(), 2 * Cin, B, HW, Cin, 32, 1e-5, 1, mode, st)
skip_ref = torch.empty(B, HW, Cout, device=dev)
call("hl_conv2d", raw.data_ptr(), 1, 2 * Cin, wpk.data_ptr(), bpk.data_ptr(), None, 0, skip_ref.data_ptr(), Cout, None, 0, B, H, W, Cin, Cout, 1, 1, _lib.CONV_SPLIT3, st)
bad = 0
for it in range(40):
    act = torch.full((B, HW, Cin), float("nan"), device=dev, dtype=torch.float16); skip = torch.full((B, HW, Cout), float("nan"), device=dev)
    call("hl_gn_skip", x.data_ptr(), Cin, stats.data_ptr(), Cin, gamma.data_ptr(), beta.data_ptr(), act.data_ptr(), Cin, wpk.data_ptr(), bpk.data_ptr(), skip.data_ptr(), Cout, B, HW, Cin, Cout, 32, 1e-5, st)
    torch.cuda.synchronize()
    d = (act.float() - act_ref.float()).abs()
    big = (d > 0.01).nonzero()
    ds = (skip - skip_ref).abs()
    bigs = (ds > 1e-3).nonzero()
    if len(big) or len(bigs):
        bad += 1
        if bad <= 4:
            print("iter", it, "act bad", len(big), big[:12].tolist(), "skip bad", len(bigs), bigs[:8].tolist())
            if len(big):
                i = big[0]; print("   act", act[i[0], i[1], i[2]].item(), "ref", act_ref[i[0], i[1], i[2]].item())
print("bad iterations", bad, "of 40")
