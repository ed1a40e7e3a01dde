"""Deterministic synthetic weights / inputs (no datasets or checkpoints ship with the reference).

Every tensor is drawn from its own ``torch.Generator`` seeded by a hash of (seed, tensor name), so a
state dict is reproducible regardless of construction order and can be loaded, unchanged, into either
the reference model or this package's model.  The reference zero-initialises 118 weight tensors
(``zero_module``); a fresh model therefore outputs exactly 0 -- every parity test uses these
randomised weights instead."""
import hashlib

import torch


def _gen(seed, name):
    h = hashlib.sha256(f"{seed}:{name}".encode()).digest()
    g = torch.Generator()
    g.manual_seed(int.from_bytes(h[:8], "little") & ((1 << 63) - 1))
    return g


def synth_state_dict(shapes, seed=0, weight_gain=1.0):
    """``shapes``: mapping name -> shape (e.g. from ``model.state_dict()``).  Returns name -> fp32 CPU
    tensor.  Conv / linear weights ~ N(0, gain^2 / fan_in); norm gains ~ 1 + 0.1 N; biases ~ 0.05 N;
    embeddings ~ 0.5 N; non-float buffers are left to the caller."""
    out = {}
    for name, shape in shapes.items():
        shape = tuple(shape)
        g = _gen(seed, name)
        r = torch.randn(shape, generator=g, dtype=torch.float32)
        is_norm = (name.endswith(("in_layers.0.weight", "out_layers.0.weight", "norm.weight"))
                   or name == "out.0.weight")
        if name.endswith("tri_planes"):
            t = 0.1 * r
        elif is_norm:
            t = 1.0 + 0.1 * r
        elif name.endswith(".bias"):
            t = 0.05 * r
        elif name.startswith("label_emb"):
            t = 0.5 * r
        elif len(shape) >= 2:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            t = r * (weight_gain / fan_in ** 0.5)
        else:
            t = r
        out[name] = t
    return out


def randomize_(module, seed=0, weight_gain=1.0):
    """Fill every floating parameter of ``module`` in place with ``synth_state_dict`` values."""
    shapes = {k: v.shape for k, v in module.state_dict().items() if v.is_floating_point()
              and not k.endswith(("_freqs", "_phases"))}
    sd = synth_state_dict(shapes, seed, weight_gain)
    module.load_state_dict(sd, strict=False)
    return sd


def synth_denoise_inputs(B, C, H, W, seed=1234):
    """x ~ N(0,1), x_cond = clamp(0.3 N, -1, 1), per-step noise generator."""
    g = torch.Generator()
    g.manual_seed(seed)
    x = torch.randn(B, C, H, W, generator=g)
    x_cond = (0.3 * torch.randn(B, C, H, W, generator=g)).clamp(-1, 1)
    return x, x_cond, g


def synth_triplane(R=256, seed=7):
    g = torch.Generator()
    g.manual_seed(seed)
    return (0.3 * torch.randn(1, 3, 9, R, R, generator=g)).clamp(-1, 1)


WORLD_BOUNDS = [[-0.55, -1.10, -0.35], [0.55, 0.95, 0.35]]


def synth_camera_rays(H=512, W=512, focal=600.0, dist=3.0, azimuth_deg=0.0, bounds=WORLD_BOUNDS):
    """Pinhole camera orbiting the box centre about the y axis.  Returns rays_o, rays_d (unnormalised:
    pixel_world - origin, as recon_NeRF/lib/if_nerf_data_utils.py:5-18 produces), near, far from the
    slab test against ``bounds`` (miss rays get near 0 / far 1, if_nerf_data_utils.py:50-85,180-185)."""
    import math
    bmin = torch.tensor(bounds[0], dtype=torch.float64)
    bmax = torch.tensor(bounds[1], dtype=torch.float64)
    centre = 0.5 * (bmin + bmax)
    az = math.radians(azimuth_deg)
    cam = centre + dist * torch.tensor([math.sin(az), 0.0, math.cos(az)], dtype=torch.float64)
    fwd = (centre - cam) / (centre - cam).norm()
    up = torch.tensor([0.0, -1.0, 0.0], dtype=torch.float64)    # image y points down
    right = torch.linalg.cross(fwd, up)
    right = right / right.norm()
    down = torch.linalg.cross(fwd, right)
    j, i = torch.meshgrid(torch.arange(H, dtype=torch.float64), torch.arange(W, dtype=torch.float64),
                          indexing="ij")
    dirs = ((i - W / 2) / focal)[..., None] * right + ((j - H / 2) / focal)[..., None] * down + fwd
    rays_d = dirs.reshape(-1, 3)
    rays_o = cam.expand_as(rays_d)
    inv = 1.0 / torch.where(rays_d.abs() < 1e-9, torch.full_like(rays_d, 1e-9), rays_d)
    t0 = (bmin - rays_o) * inv
    t1 = (bmax - rays_o) * inv
    tn = torch.minimum(t0, t1).amax(-1)
    tf = torch.maximum(t0, t1).amin(-1)
    hit = tf > torch.clamp(tn, min=0.0)
    near = torch.where(hit, torch.clamp(tn, min=0.0), torch.zeros_like(tn))
    far = torch.where(hit, tf, torch.ones_like(tf))
    return rays_o.float().contiguous(), rays_d.float().contiguous(), near.float(), far.float(), hit
