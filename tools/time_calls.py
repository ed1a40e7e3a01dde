"""Per-launch timing of the production step: every distinct C-ABI call of the step plan is timed ALONE (CUDA events,
`reps` back-to-back launches on its real buffers) and reported with its count in the step -- the map of where an
isolated-kernel optimisation pays.  Sum(alone x count) over-counts what the two-stream graph overlaps; the in-situ
figure stays tools/ablate_step.py."""
import os
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from bench import build_model  # noqa: E402
from humanliff_b200 import _lib  # noqa: E402
from humanliff_b200._lib import call  # noqa: E402

dev = torch.device("cuda:0")
B = int(os.environ.get("HL_B", "4"))
reps = int(os.environ.get("HL_REPS", "20"))
model, diffusion, _ = build_model(dev, "fp16")
g = torch.Generator().manual_seed(0)
x = torch.randn(B, 27, 256, 256, generator=g).to(dev)
xc = torch.randn(B, 27, 256, 256, generator=g).to(dev)
y = (torch.arange(B) % 4).to(dev)
t = torch.full((B,), 500, dtype=torch.int64, device=dev)
for _ in range(3):
    model(x, t, x_cond=xc, y=y)
plan = next(iter(model._plans.values()))
lib = _lib.load()
stream = torch.cuda.current_stream(dev).cuda_stream
ws = torch.empty(plan.SPLITK_BYTES // 4, device=dev)
lib.hl_conv_set_workspace(ws.data_ptr(), plan.SPLITK_BYTES, stream)


def key_of(name, a):
    if name == "hl_conv2d":
        Bn, H, W, Cin, Cout, k, s, flags = a[11:19]
        return (name, H, Cin, Cout, k, s, "res" if a[5] else "-", "st" if a[9] else "-", "f%d" % flags)
    if name == "hl_conv2d_dual":
        Bn, H, W, Cin, Cout, k, s, flags = a[15:23]
        return (name, H, Cin, Cout, k, s, "f%d" % flags)
    if name == "hl_gn_skip":
        return (name, a[13], a[14], a[15])
    if name == "hl_gn_apply":
        return (name, a[14], a[15], "film" if a[6] else "-", "raw" if a[11] else "-", "m%d" % a[19])
    if name == "hl_attention":
        return (name, a[7], a[8])
    if name == "hl_cast_operand":
        return (name, a[5], a[6], a[7])
    if name == "hl_upsample2x":
        return (name, a[6], a[8])
    return (name,)


groups = OrderedDict()
for name, a, br in plan.calls:
    if name[0] == "#":
        continue
    k = key_of(name, a)
    if k not in groups:
        groups[k] = [name, a, 0, [0, 0]]
    groups[k][2] += 1
    groups[k][3][br] += 1

e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
rows = []
for k, (name, a, n, brs) in groups.items():
    for _ in range(3):
        call(name, *a, stream)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        call(name, *a, stream)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    tf = ""
    if name == "hl_conv2d":
        Bn, H, W, Cin, Cout, ks, s, flags = a[11:19]
        npass = 3 if flags & 16 else 2 if flags & (32 | 128) else 1
        fl = 2.0 * Bn * (H // s) * (W // s) * Cout * Cin * ks * ks
        tf = "%7.1f TF/s (x%d passes)" % (fl / us * 1e-6, npass)
    rows.append((us * n, us, n, brs, k, tf))
tot = sum(r[0] for r in rows)
print("alone-us  count(main/side)  total-ms  key")
for totus, us, n, brs, k, tf in sorted(rows, key=lambda r: -r[0]):
    print("%8.1f  %3d (%3d/%3d)  %7.3f  %s  %s" % (us, n, brs[0], brs[1], totus * 1e-3, " ".join(str(v) for v in k), tf))
print("sum of alone x count: %.3f ms over %d launches" % (tot * 1e-3, sum(r[2] for r in rows)))
