"""In-situ cost of kernel groups: replay the production step's CUDA graph with a subset of its launches
removed (results are garbage, timing is valid: no kernel has data-dependent control flow) and report the
difference to the full step.  Unlike an ncu launch list this sees warm L2 and the two-stream overlap."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from bench import build_model  # noqa: E402

dev = torch.device("cuda:0")
B = int(os.environ.get("HL_B", "4"))
model, diffusion, _ = build_model(dev, "fp16")
g = torch.Generator().manual_seed(0)
x = torch.randn(B, 27, 256, 256, generator=g).to(dev)
xc = torch.zeros_like(x)
y = (torch.arange(B) % 4).to(dev)
t = torch.full((B,), 500, dtype=torch.int64, device=dev)
for _ in range(3):
    model(x, t, x_cond=xc, y=y)
plan = next(iter(model._plans.values()))
full = list(plan.calls)


def conv_h(a):
    return a[12]


def is_conv(n):
    return n in ("hl_conv2d", "hl_conv2d_dual")


def H_of(n, a):
    return a[16] if n == "hl_conv2d_dual" else a[12]


def k_of(n, a):
    return a[20] if n == "hl_conv2d_dual" else a[16]


def pix(name, a):
    if name == "hl_conv2d":
        return a[12] * a[13]
    if name == "hl_conv2d_dual":
        return a[16] * a[17]
    if name == "hl_gn_apply":
        return a[14]
    if name == "hl_gn_skip":
        return a[13]
    if name == "hl_attention":
        return a[7]
    return None


def timed(calls, reps=10):
    plan.calls = calls
    plan.graph = None
    plan.run(x, t, xc, y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        plan.graph.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


base = timed(full)
print(f"full step (graph replay, no I/O copies): {base:.3f} ms, {len(full)} entries", flush=True)
if os.environ.get("HL_ABLATE_FULL_ONLY"):
    sys.exit(0)
cases = {
    "conv H<=8": lambda n, a: is_conv(n) and H_of(n, a) <= 8,
    "conv H==16": lambda n, a: is_conv(n) and H_of(n, a) == 16,
    "conv H==32": lambda n, a: is_conv(n) and H_of(n, a) == 32,
    "conv H==64": lambda n, a: is_conv(n) and H_of(n, a) == 64,
    "conv H==128": lambda n, a: is_conv(n) and H_of(n, a) == 128,
    "conv H==256": lambda n, a: is_conv(n) and H_of(n, a) == 256,
    "conv 1x1 H>=128": lambda n, a: is_conv(n) and H_of(n, a) >= 128 and k_of(n, a) == 1,
    "gn_apply all": lambda n, a: n == "hl_gn_apply",
    "gn_apply HW>=128^2": lambda n, a: n == "hl_gn_apply" and a[14] >= 128 * 128,
    "gn_apply HW<=32^2": lambda n, a: n == "hl_gn_apply" and a[14] <= 32 * 32,
    "gn_skip (fused GN1 + 1x1 skip)": lambda n, a: n == "hl_gn_skip",
    "attention": lambda n, a: n == "hl_attention",
    "cast/upsample/stats": lambda n, a: n in ("hl_cast_operand", "hl_upsample2x", "hl_gn_stats"),
    "everything at H<=32": lambda n, a: (pix(n, a) or 1 << 30) <= 32 * 32,
    "side stream (ControlNet encoder)": None,
}
for label, pred in cases.items():
    if pred is None:
        calls = [c for c in full if c[2] == 0 or c[0][0] == "#"]
    else:
        calls = [c for c in full if not (c[0][0] != "#" and pred(c[0], c[1]))]
    ms = timed(calls)
    print(f"without {label:34s}: {ms:7.3f} ms  (delta {base - ms:6.3f} ms, {len(full) - len(calls)} launches removed)")
# serial: everything on one stream
plan.calls = full
