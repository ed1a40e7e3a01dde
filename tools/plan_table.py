"""Host-only: the tiling / split-K decision of every conv launch of the production step (hl_conv2d_plan_info)."""
import ctypes
import os
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from humanliff_b200 import _lib, factory  # noqa: E402
from humanliff_b200.unet import _StepPlan  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
m, _ = factory.create_model_and_diffusion(**factory.production_flags(""))
cpu = torch.device("cpu")
m._pack(cpu)
plan = _StepPlan(m, cpu, B, 256, 256)
lib = _lib.load()
rows = OrderedDict()
for name, a, _br in plan.calls:
    if name != "hl_conv2d":
        continue
    Bn, H, W, Cin, Cout, k, s = a[11:18]
    key = (H, Cin, Cout, k, s, a[5] is not None, a[9] is not None)
    rows[key] = rows.get(key, 0) + 1
print("H Cin Cout k s res stats | n | pair mh N halo Aslots Bslots nbuf acc tmem smemKB grid tiles S kc/S epi_stats")
for key, n in rows.items():
    H, Cin, Cout, k, s, res, st = key
    out = (ctypes.c_int * 16)()
    lib.hl_conv2d_plan_info(1, B, H, H, Cin, Cout, k, s, int(res), int(st), plan.SPLITK_BYTES, out)
    o = list(out)
    print(key, n, "|", o[1], o[2], o[3], o[4], o[5], o[6], o[7], o[8], o[9], o[10] // 1024, o[11], o[12], o[13], o[14], o[15])
