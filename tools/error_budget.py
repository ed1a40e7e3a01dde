"""Where does the operand-rounding error of the fp16 plan come from?  CPU emulation (oracle) on the production
architecture at 64^2: every contraction rounds its operands to fp16 except one group of layers, which runs
exact -- the drop in the final rel-L2 of epsilon is that group's share of the error budget.  Also evaluates
candidate mixed plans (hi+lo operands for a few layers).  TEST / ANALYSIS TOOL, not product code."""
import json
import os
import re
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import CASES, load_golden, model_state_dict, rel_l2, rel_max   # noqa: E402
from oracle import unet_oracle   # noqa: E402


def main():
    torch.set_num_threads(int(os.environ.get("THREADS", "4")))
    fname, flags, seed, heads = CASES["prod64"]
    _, diffusion, sd = model_state_dict(flags, seed)
    g = load_golden(fname)
    t = int(os.environ.get("T", "100"))
    ts = torch.tensor(diffusion.timestep_map)[torch.full((g["x"].shape[0],), t)]
    run = lambda pol: unet_oracle.unet_forward(sd, g["x"], ts, g["x_cond"], g["y"], num_heads=heads, operand_round=pol)
    ref = run(None)
    groups = {
        "none (all fp16)": r"^$",
        "out.2": r"^out\.2$",
        "dec 20-23 (256^2)": r"^output_blocks\.(2[0-3])\.",
        "dec 16-19 (128^2)": r"^output_blocks\.(1[6-9])\.",
        "dec 8-15 (64^2,32^2)": r"^output_blocks\.([89]|1[0-5])\.",
        "dec 0-7 (16^2,8^2)": r"^output_blocks\.[0-7]\.",
        "middle": r"^middle_block\.",
        "enc 0-8 (256^2,128^2)": r"^input_blocks\.[0-8]\.",
        "enc 9-23": r"^input_blocks\.(9|1[0-9]|2[0-3])\.",
        "cond enc 0-8": r"^input_blocks_cond\.[0-8]\.",
        "cond enc 9-23": r"^input_blocks_cond\.(9|1[0-9]|2[0-3])\.",
        "proj": r"^input_blocks_proj_cond\.",
        "skip convs": r"skip_connection$",
        "all convs of ResBlocks' second conv (out_layers.3)": r"out_layers\.3$",
        "all first convs (in_layers.2)": r"in_layers\.2$",
        "attention blocks": r"\.1$|middle_block\.1$",
    }
    rows = []
    for name, rx in groups.items():
        cre = re.compile(rx)
        out = run(lambda p: None if cre.search(p) else "fp16")
        rows.append({"exact_group": name, "rel_l2": rel_l2(out, ref), "rel_max": rel_max(out, ref)})
        print(json.dumps(rows[-1]), flush=True)
    plans = {
        "out.2 hi+lo": lambda p: "fp16x2" if p == "out.2" else "fp16",
        "out.2 + dec23 hi+lo": lambda p: "fp16x2" if (p == "out.2" or p.startswith("output_blocks.23.")) else "fp16",
        "raw convs tf32-trunc": "tf32_trunc_raw_fp16",
    }
    for name, pol in plans.items():
        if isinstance(pol, str):
            continue
        out = run(pol)
        print(json.dumps({"plan": name, "rel_l2": rel_l2(out, ref), "rel_max": rel_max(out, ref)}), flush=True)


if __name__ == "__main__":
    main()
