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


def run_logged():
    """Execute the launch list call by call, snapshotting every workspace buffer after each call."""
    snaps = []
    for name, args in plan.calls:
        call(name, *args, stream)
        torch.cuda.synchronize()
        snaps.append({k: v.clone() for k, v in plan.bufs.items()} | {"stats": plan.stats.clone()})
    return snaps


a = run_logged()
for rep in range(3):
    b = run_logged()
    first = None
    for i, (sa, sb) in enumerate(zip(a, b)):
        bad = []
        for k in sa:
            va, vb = sa[k].float() if sa[k].dtype != torch.float64 else sa[k], sb[k].float() if sb[k].dtype != torch.float64 else sb[k]
            same = torch.equal(torch.nan_to_num(va), torch.nan_to_num(vb))
            if not same:
                d = (torch.nan_to_num(va) - torch.nan_to_num(vb)).abs().max().item()
                bad.append((k, d, float(torch.nan_to_num(va).abs().max())))
        if bad:
            first = (i, plan.calls[i][0], bad[:4])
            break
    print("rep", rep, "first differing call:", first)
    if first:
        i = first[0]
        name, args = plan.calls[i]
        print("   args:", [a_ if not isinstance(a_, int) or a_ < 1 << 20 else hex(a_) for a_ in args])
