"""Shared test helpers: configs, golden loading, synthetic state dicts."""
import os

import numpy as np
import torch

from humanliff_b200 import factory, synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

PROD = factory.production_flags("250")
TINY = dict(PROD, image_size=32, num_channels=64, num_res_blocks=1, num_heads=2, attention_resolutions="16,8")

CASES = {"tiny": ("unet_tiny_32.npz", TINY, 11, 2), "prod64": ("unet_prod_64.npz", PROD, 0, 4)}


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name))
    return {k: torch.from_numpy(z[k]) if z[k].ndim else z[k].item() for k in z.files}


def model_state_dict(flags, seed):
    model, diffusion = factory.create_model_and_diffusion(**flags)
    sd = synth.synth_state_dict({k: v.shape for k, v in model.state_dict().items()}, seed=seed)
    return model, diffusion, sd


def renderer_state_dict(seed=3, precision="fp16"):
    from humanliff_b200.renderer import Renderer
    r = Renderer(triplane_ch=27, test=True, precision=precision)
    shapes = {k: v.shape for k, v in r.state_dict().items() if not k.startswith("view_enc")}
    sd = synth.synth_state_dict(shapes, seed=seed, weight_gain=1.5)
    r.load_state_dict(sd, strict=False)
    return r, sd


def rel_l2(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def rel_max(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
