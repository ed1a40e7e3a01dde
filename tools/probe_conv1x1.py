"""The HBM-bound 1x1 convolutions of the step at 256^2 (B = 4): time per launch and achieved HBM rate, for the
production operand modes (three-pass hi + lo, HL_CONV_SPLIT3) against a one-pass fp16 launch of the same shape,
over a few tilings (hl_conv_set_tuning / hl_conv_set_tuning2)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from humanliff_b200 import _lib  # noqa: E402
from humanliff_b200._lib import call  # noqa: E402
from humanliff_b200.unet import pack_conv  # noqa: E402

dev = torch.device("cuda:0")
lib = _lib.load()
B, HW = 4, int(os.environ.get("HL_HW", "256"))
stream = torch.cuda.current_stream(dev).cuda_stream
g = torch.Generator().manual_seed(0)
# (Cin, Cout, residual, stats, split3)
cases = [(384, 192, False, False, True), (192, 192, True, True, True), (192, 192, False, True, True),
         (384, 192, False, False, False), (192, 192, True, True, False), (192, 192, True, False, False),
         (192, 192, False, True, False), (192, 192, False, False, False)]
tunings = [("auto", (-1, -1, -1, -1, -1), (-1, -1, -1)), ("nbuf=2", (-1, -1, -1, -1, -1), (-1, 2, -1)),
           ("nbuf=3", (-1, -1, -1, -1, -1), (-1, 3, -1)), ("1cta", (-1, -1, -1, -1, -1), (-1, -1, 0))]
if os.environ.get("HL_ONLY_AUTO"):
    tunings = tunings[:1]
for Cin, Cout, res, st_on, split in cases:
    x = torch.randn(B, HW, HW, 2 * Cin, device=dev).half()
    r = torch.randn(B, HW, HW, Cout, device=dev) if res else None
    y = torch.empty(B, HW, HW, Cout, device=dev)
    st = torch.zeros(B * Cout * 2, device=dev, dtype=torch.float64) if st_on else None
    w = torch.randn(Cout, Cin, 1, 1, generator=g) / Cin ** 0.5
    wpk, bpk = pack_conv(w, torch.zeros(Cout), Cin, "fp16", dev, mode="split" if split else None)
    flags = _lib.CONV_SPLIT3 if split else 0
    nbytes = B * HW * HW * (Cin * (4 if split else 2) + Cout * 4 * (2 if res else 1))

    def launch():
        call("hl_conv2d", x.data_ptr(), 1, 2 * Cin, wpk.data_ptr(), bpk.data_ptr(), r.data_ptr() if res else None, Cout,
             y.data_ptr(), Cout, st.data_ptr() if st_on else None, Cout, B, HW, HW, Cin, Cout, 1, 1, flags, stream)
    for name, t1, t2 in tunings:
        lib.hl_conv_set_tuning(*t1)
        lib.hl_conv_set_tuning2(*t2)
        try:
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
            if name == "auto":       # one more launch with the in-kernel wait counters of CTA 0 switched on
                prof = torch.zeros(16, device=dev, dtype=torch.int64)
                lib.hl_conv_set_profile(prof.data_ptr())
                launch()
                torch.cuda.synchronize()
                lib.hl_conv_set_profile(None)
                pr = prof.cpu().tolist()
                keys = ["total", "mma_wait_A", "mma_wait_tmem_empty", "mma_wait_B", "epi_wait_tmem_full", "epi_wait_res",
                        "epi_wait_named_bar", "prodA_wait_empty", "prodB_wait_empty", "e0_wait_store_read", "tiles"]
                tiles = max(pr[10], 1)
                print("      cycles per tile (CTA 0): " + ", ".join("%s %d" % (k, v // tiles) for k, v in zip(keys[:10], pr[:10])) +
                      ", tiles %d" % tiles, flush=True)
            print("1x1 %d->%d @%d^2 res=%d stats=%d %s %-10s: %6.1f us  %5.0f GB/s" % (
                Cin, Cout, HW, res, st_on, "split3" if split else "1-pass", name, ms * 1e3, nbytes / ms / 1e6), flush=True)
        except RuntimeError as e:
            print("1x1 %d->%d %s %s: %s" % (Cin, Cout, "split3" if split else "1-pass", name, str(e)[:80]))
        finally:
            lib.hl_conv_set_tuning(-1, -1, -1, -1, -1)
            lib.hl_conv_set_tuning2(-1, -1, -1)
