"""k_gn_apply on the step's dominant shapes (256^2, B = 4): time per launch and achieved HBM rate for a sweep of the
blocks-per-SM launch parameter (hl_gn_set_tuning); also the driver for `ncu -k regex:k_gn_apply`."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from humanliff_b200 import _lib  # noqa: E402
from humanliff_b200._lib import call  # noqa: E402

dev = torch.device("cuda:0")
B, HW = 4, 256 * 256
st = torch.cuda.current_stream(dev).cuda_stream
lib = _lib.load()
sweep = [int(v) for v in os.environ.get("HL_GN_SWEEP", "0").split(",")]   # 0 = the default (one wave by occupancy)
# (C, input fp16?, FiLM?, raw split copy?)
cases = [(192, False, True, False), (192, True, True, False), (384, False, False, True), (192, False, False, False)]
for C, xf16, film_on, raw_on in cases:
    x = torch.randn(B, HW, C, device=dev)
    xin = x.half() if xf16 else x
    y = torch.empty(B, HW, C, device=dev, dtype=torch.float16)
    raw = torch.empty(B, HW, 2 * C, device=dev, dtype=torch.float16) if raw_on else None
    gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    film = torch.zeros(B, 2 * C, device=dev)
    stats = torch.zeros(B * C * 2, device=dev, dtype=torch.float64)
    call("hl_gn_stats", x.data_ptr(), C, B, HW, C, stats.data_ptr(), C, st)
    mode = (_lib.OP_X_F16 if xf16 else 0) | (((_lib.OP_SPLIT | _lib.OP_SCALED) << _lib.OP_RAW_SHIFT) if raw_on else 0)
    nbytes = B * HW * C * ((2 if xf16 else 4) + 2 + (4 if raw_on else 0))

    def launch():
        call("hl_gn_apply", xin.data_ptr(), C, stats.data_ptr(), C, gamma.data_ptr(), beta.data_ptr(),
             film.data_ptr() if film_on else None, 2 * C, y.data_ptr(), 1, C, raw.data_ptr() if raw_on else None,
             2 * C if raw_on else 0, B, HW, C, 32, 1e-5, 1, mode, st)
    for bps in sweep:
        lib.hl_gn_set_tuning(bps)
        for _ in range(3):
            launch()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            launch()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print("gn_apply C=%d @256^2 B=4 x_f16=%d film=%d raw=%d blocks/SM=%2d: %6.1f us, %5.0f GB/s (%d MB)"
              % (C, xf16, film_on, raw_on, bps, ms * 1e3, nbytes / ms / 1e6, nbytes >> 20), flush=True)
lib.hl_gn_set_tuning(0)
