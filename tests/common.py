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


def sq_dist_separate(q, v):
    """Squared distance of two fp32 points with every operation rounded: ((dx^2 + dy^2) + dz^2) -- the arithmetic of the
    oracle's nearest_vertex (and of the torch stand-in for pytorch3d's knn_points in oracle/make_goldens.py)."""
    f32 = np.float32
    d = (q - v).astype(f32)
    s = (d * d).astype(f32)
    return f32(f32(s[0] + s[1]) + s[2])


def sq_dist_fused(q, v):
    """The same with fused multiply-adds, fma(dz, dz, fma(dy, dy, dx^2)): what the packed filter of the search computes."""
    f32 = np.float32
    d = (q - v).astype(f32).astype(np.float64)
    r = f32(d[0] * d[0])
    r = f32(d[1] * d[1] + np.float64(r))
    return f32(d[2] * d[2] + np.float64(r))


def nearest_vertex_near_tie_cases(n_adversarial, n_ties, seed=11):
    """(query, v1, v2) triples of fp32 points, each in its own region 3 units from the next.
    adversarial: the separately rounded and the fused squared distance DISAGREE about which of v1, v2 is nearer (found by
    a seeded search: v2's z is solved so that the real distances agree to about one ulp); ties: v = q +- delta with
    everything exactly representable, so both distances are identical under any rounding."""
    f32 = np.float32
    rs = np.random.RandomState(seed)
    centre = lambda k: np.array([3.0 * k - 90.0, 0.37, -0.21], dtype=f32)
    cases, k = [], 0
    while len(cases) < n_adversarial:
        c = centre(len(cases))
        q = (c + rs.uniform(-0.2, 0.2, 3)).astype(f32)
        off = rs.normal(size=3); off *= rs.uniform(0.05, 0.4) / np.linalg.norm(off)
        v1 = (q + off).astype(f32)
        D1 = float((((q - v1).astype(f32).astype(np.float64)) ** 2).sum())
        v2 = (q + rs.uniform(-0.6, 0.6, 3) * np.sqrt(D1)).astype(f32)
        dxy = (q - v2).astype(f32).astype(np.float64)[:2]
        v2[2] = f32(np.float64(q[2]) + rs.choice([-1.0, 1.0]) * np.sqrt(D1 - float((dxy ** 2).sum())))
        s = np.sign(float(sq_dist_separate(q, v1)) - float(sq_dist_separate(q, v2)))
        if s != 0 and s == -np.sign(float(sq_dist_fused(q, v1)) - float(sq_dist_fused(q, v2))):      # strict reversal
            cases.append((q, v1, v2))
        k += 1
        assert k < 200000
    ties = []
    for j in range(n_ties):
        q = (np.round(centre(100 + j) * 256) / 256 + rs.randint(-64, 64, 3) / 256.0).astype(f32)
        dl = (rs.randint(-1500, 1500, 3) / 4096.0).astype(f32)
        v1, v2 = (q + dl).astype(f32), (q - dl).astype(f32)
        assert sq_dist_separate(q, v1) == sq_dist_separate(q, v2) and np.all(q - v1 == -dl)
        ties.append((q, v1, v2))
    return cases, ties
