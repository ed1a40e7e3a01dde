"""Run the tiny UNet plan eagerly several times on identical inputs and report, call by call, the first
workspace buffers whose contents differ between runs (race detector for the kernels)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
from common import CASES, load_golden, model_state_dict  # noqa: E402
from humanliff_b200._lib import call  # noqa: E402

case = sys.argv[1] if len(sys.argv) > 1 else "tiny"
prec = sys.argv[2] if len(sys.argv) > 2 else "fp16"
fname, flags, seed, heads = CASES[case]
model, diffusion, sd = model_state_dict(dict(flags, precision=prec), seed)
model.load_state_dict(sd)
dev = torch.device("cuda:0")
model = model.to(dev)
model.use_cuda_graph = False
g = load_golden(fname)
x, xc, y = g["x"].to(dev), g["x_cond"].to(dev), g["y"].to(dev)
ts = torch.full((x.shape[0],), 400, device=dev)
model(x, ts, xc, y=y)
plan = next(iter(model._plans.values()))
stream = torch.cuda.current_stream(dev).cuda_stream


def snapshot():
    d = {k: v.clone() for k, v in plan.bufs.items()}
    d["stats"] = plan.stats.clone()
    return d


def restore(d):
    for k, v in plan.bufs.items():
        v.copy_(d[k])
    plan.stats.copy_(d["stats"])


# every call executed twice from the same state: per-kernel nondeterminism, isolated
for t in plan.bufs.values():
    t.zero_()
plan.stats.zero_()
worst = {}
for i, (name, args, _br) in enumerate([c for c in plan.calls if c[0][0] != '#']):
    s0 = snapshot()
    call(name, *args, stream)
    torch.cuda.synchronize()
    o1 = snapshot()
    restore(s0)
    call(name, *args, stream)
    torch.cuda.synchronize()
    o2 = snapshot()
    for k in o1:
        va, vb = torch.nan_to_num(o1[k].double()), torch.nan_to_num(o2[k].double())
        if not torch.equal(va, vb):
            rel = float((va - vb).abs().max() / va.abs().max().clamp_min(1e-30))
            key = (name, k)
            if rel > worst.get(key, (0, 0))[0]:
                worst[key] = (rel, i)
for (name, k), (rel, i) in sorted(worst.items(), key=lambda kv: -kv[1][0]):
    print(f"nondeterministic: {name:18s} buffer {k:10s} max-rel diff {rel:.3e} (call {i})")
print("calls:", len(plan.calls), "nondeterministic (kernel, buffer) pairs:", len(worst))

# end-to-end: the whole plan twice
outs = []
for rep in range(3):
    outs.append(model(x, ts, xc, y=y).double())
print("end-to-end eps rel-L2 between runs:", [float((o - outs[0]).norm() / outs[0].norm()) for o in outs[1:]])

# ---- part 2: single convolutions, 6 launches each, bitwise comparison of the outputs -------------------
from test_kernels_gpu import _conv_case  # noqa: E402
for shape in [(2, 32, 32, 64, 64, 3, 1), (2, 16, 16, 128, 128, 3, 1), (2, 8, 8, 128, 128, 3, 1), (2, 4, 4, 256, 256, 3, 1),
              (2, 32, 32, 64, 64, 1, 1), (2, 32, 32, 64, 64, 3, 2), (2, 8, 8, 256, 128, 1, 1)]:
    B, H, W, Cin, Cout, k, s = shape
    outs = []
    for rep in range(6):
        yv, ref, ref_r, st = _conv_case(dev, B, H, W, Cin, Cout, k, s, mode="fp16", residual=True, stats=True, seed=11)
        outs.append((yv, st))
    dy = max(float((o[0] - outs[0][0]).abs().max()) for o in outs[1:])
    ds = max(float((o[1] - outs[0][1]).abs().max()) for o in outs[1:])
    print("conv", shape, "max |dy| over 5 repeats:", dy, " max |dstats|:", ds, " err vs rounded ref:",
          float((outs[0][0] - ref_r).norm() / ref_r.norm()))
