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


# --------------------------------------------------------------------------------------------------------------
# Synthetic SMPL-shaped body model.  The reference's canonical-space path (use_canonical_space=True,
# human_diffusion/NeRF/renderer.py:41-50) loads assets/SMPL_NEUTRAL.pkl, which is licensed and ships neither with the
# reference nor here.  Everything the deformation reads from it is a table of a fixed shape -- v_template [6890, 3],
# shapedirs [6890, 3, 10], posedirs [6890, 3, 207], J_regressor [24, 6890], weights [6890, 24], kintree_table [2, 24] --
# so a seeded stand-in with the same shapes, a plausible skeleton and smooth skinning weights exercises the same code.
SMPL_PARENTS = [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21]
_SMPL_JOINTS = [
    (0.00, -0.22, 0.00), (0.07, -0.31, 0.00), (-0.07, -0.31, 0.00), (0.00, -0.10, 0.00), (0.10, -0.68, 0.00),
    (-0.10, -0.68, 0.00), (0.00, 0.03, 0.00), (0.09, -1.06, -0.03), (-0.09, -1.06, -0.03), (0.00, 0.09, 0.00),
    (0.11, -1.11, 0.08), (-0.11, -1.11, 0.08), (0.00, 0.30, -0.02), (0.08, 0.20, 0.00), (-0.08, 0.20, 0.00),
    (0.00, 0.38, 0.02), (0.18, 0.23, 0.00), (-0.18, 0.23, 0.00), (0.43, 0.22, 0.00), (-0.43, 0.22, 0.00),
    (0.68, 0.22, 0.00), (-0.68, 0.22, 0.00), (0.76, 0.21, 0.00), (-0.76, 0.21, 0.00)]
_SMPL_RADIUS = [0.13, 0.09, 0.09, 0.13, 0.07, 0.07, 0.14, 0.05, 0.05, 0.15, 0.04, 0.04, 0.06, 0.08, 0.08, 0.10,
                0.07, 0.07, 0.05, 0.05, 0.04, 0.04, 0.03, 0.03]


def synth_smpl(seed=5, n_verts=6890, n_betas=10):
    """-> dict of numpy arrays with the keys / shapes / dtypes SMPL_to_tensor (human_diffusion/NeRF/renderer.py:340-352)
    converts: a tube of vertices around every bone of a 24-joint skeleton."""
    import numpy as np
    rs = np.random.RandomState(seed)
    J = np.array(_SMPL_JOINTS, dtype=np.float64)
    par = np.array(SMPL_PARENTS)
    bone = rs.randint(1, 24, size=n_verts)
    s = rs.rand(n_verts, 1)
    centre = J[par[bone]] * (1 - s) + J[bone] * s
    dirs = rs.randn(n_verts, 3)
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    rad = np.array(_SMPL_RADIUS)
    r = (rad[par[bone]] * (1 - s[:, 0]) + rad[bone] * s[:, 0]) * (0.9 + 0.2 * rs.rand(n_verts))
    v = centre + dirs * r[:, None]
    d2 = ((v[:, None, :] - J[None]) ** 2).sum(-1)                       # [V, 24]
    # skinning as on a real body: the segment between joint p = parent(j) and joint j moves with p, blends into j over
    # its last 30 % and into p's own parent over its first 30 % (at most three joints per vertex, smooth along the limb)
    smooth = lambda x: np.clip(x, 0.0, 1.0) ** 2 * (3.0 - 2.0 * np.clip(x, 0.0, 1.0))
    w = np.zeros((n_verts, 24))
    pj = par[bone]
    w_child = 0.5 * smooth((s[:, 0] - 0.7) / 0.3)
    w_grand = np.where(par[pj] >= 0, 0.5 * smooth((0.3 - s[:, 0]) / 0.3), 0.0)
    rows = np.arange(n_verts)
    np.add.at(w, (rows, bone), w_child)
    np.add.at(w, (rows, np.maximum(par[pj], 0)), w_grand)
    np.add.at(w, (rows, pj), 1.0 - w_child - w_grand)
    jr = np.exp(-d2.T / (2 * 0.05 ** 2)) + 1e-12                         # [24, V]
    jr /= jr.sum(1, keepdims=True)
    # blend shapes are smooth displacement fields on a real body: affine functions of the rest position here
    hom = np.concatenate([v, np.ones((n_verts, 1))], axis=1)                       # [V, 4]
    shapedirs = np.einsum("vh,hck->vck", hom, 0.02 * rs.randn(4, 3, n_betas))
    posedirs = np.einsum("vh,hcf->vcf", hom, 0.004 * rs.randn(4, 3, 207))
    kin = np.stack([np.where(par < 0, 4294967295, par), np.arange(24)]).astype(np.int64)
    faces = rs.randint(0, n_verts, size=(13776, 3)).astype(np.int64)
    return {"v_template": v, "shapedirs": shapedirs, "posedirs": posedirs, "J_regressor": jr, "weights": w,
            "kintree_table": kin, "f": faces}


def _rodrigues_np(rv):
    """[N, 3] axis-angle -> [N, 3, 3] (the formula of batch_rodrigues, renderer.py:435-462, in float64)."""
    import numpy as np
    angle = np.linalg.norm(rv + 1e-8, axis=1, keepdims=True)
    k = rv / angle
    K = np.zeros((rv.shape[0], 3, 3))
    K[:, 0, 1], K[:, 0, 2], K[:, 1, 0] = -k[:, 2], k[:, 1], k[:, 2]
    K[:, 1, 2], K[:, 2, 0], K[:, 2, 1] = -k[:, 0], -k[:, 1], k[:, 0]
    s, c = np.sin(angle)[:, :, None], np.cos(angle)[:, :, None]
    return np.eye(3)[None] + s * K + (1 - c) * (K @ K)


def smpl_pose_vertices(smpl, poses, betas):
    """Forward linear blend skinning in float64 numpy (test-data generation only): -> posed vertices [V, 3] in SMPL
    space.  The inverse of what deform_target2c_op undoes."""
    import numpy as np
    poses = np.asarray(poses, dtype=np.float64).reshape(24, 3)
    betas = np.asarray(betas, dtype=np.float64).reshape(-1)
    v_shaped = smpl["v_template"] + smpl["shapedirs"][..., :betas.size] @ betas
    rot = _rodrigues_np(poses)
    v_posed = v_shaped + smpl["posedirs"] @ (rot[1:] - np.eye(3)[None]).reshape(-1)
    Jp = smpl["J_regressor"] @ v_shaped
    G = [None] * 24
    for j in range(24):
        T = np.eye(4)
        T[:3, :3] = rot[j]
        T[:3, 3] = Jp[j] - (Jp[SMPL_PARENTS[j]] if j else 0)
        G[j] = T if j == 0 else G[SMPL_PARENTS[j]] @ T
    A = np.stack(G)
    A[:, :3, 3] -= np.einsum("jab,jb->ja", A[:, :3, :3], Jp)
    Av = np.einsum("vj,jab->vab", smpl["weights"], A)
    return np.einsum("vab,vb->va", Av[:, :3, :3], v_posed) + Av[:, :3, 3]


def big_pose():
    """The canonical 'big pose' of the reference (recon_NeRF/lib/renderer.py:50-58)."""
    import math
    import numpy as np
    p = np.zeros(72)
    p[5], p[8], p[23], p[26] = math.radians(45), -math.radians(45), -math.radians(30), math.radians(30)
    return p


def synth_canonical_frame(smpl, seed=21):
    """One posed frame in the layout TightCapView_datasets.py:351-356 hands to the renderer: ``tp_input`` with
    params {poses [1,1,72], shapes [1,1,10], R [1,3,3], Th [1,1,3]}, vertices [1,V,3] (world space), world_bounds,
    t_params (big pose, zero shape), t_world_bounds (bounds of the big-pose vertices, padded as the dataset does)."""
    import numpy as np
    rs = np.random.RandomState(seed)
    poses = 0.25 * rs.randn(72)
    poses[:3] = 0.0
    betas = 0.8 * rs.randn(10)
    Rm = _rodrigues_np(np.array([[0.15, -0.4, 0.1]]))[0]
    Th = np.array([0.12, 0.05, -0.08])
    v_smpl = smpl_pose_vertices(smpl, poses, betas)
    v_world = v_smpl @ Rm.T + Th
    wb = np.stack([v_world.min(0) - 0.05, v_world.max(0) + 0.05])
    v_big = smpl_pose_vertices(smpl, big_pose(), np.zeros(10))
    lo, hi = v_big.min(0) - 0.05, v_big.max(0) + 0.05
    lo[1] -= 0.1
    hi[1] += 0.1
    f = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float32))
    return {
        "params": {"poses": f(poses).view(1, 1, 72), "shapes": f(betas).view(1, 1, 10), "R": f(Rm).view(1, 3, 3),
                   "Th": f(Th).view(1, 1, 3)},
        "vertices": f(v_world).view(1, -1, 3),
        "world_bounds": f(wb).view(1, 2, 3),
        "t_params": {"poses": f(big_pose()).view(1, 1, 72), "shapes": torch.zeros(1, 1, 10),
                     "R": torch.eye(3).view(1, 3, 3), "Th": torch.zeros(1, 1, 3)},
        "t_world_bounds": f(np.stack([lo, hi])).view(1, 2, 3),
    }


def synth_canonical_rays(tp, n_rays, seed=99):
    """Rays of a 256 x 256 orbit camera aimed at the posed body's world box: half hit the box, half are drawn from the
    whole image (border / miss rays included); plus the [n_rays, 128] uniforms of sample_pdf."""
    wb = tp["world_bounds"][0].tolist()
    ro, rd, near, far, hit = synth_camera_rays(256, 256, focal=300.0, azimuth_deg=30.0, bounds=wb)
    g = torch.Generator()
    g.manual_seed(seed)
    hit_idx = torch.nonzero(hit)[:, 0]
    sel = torch.cat([hit_idx[torch.randperm(hit_idx.numel(), generator=g)[:n_rays // 2]],
                     torch.randint(0, ro.shape[0], (n_rays - n_rays // 2,), generator=g)])
    u = torch.rand(n_rays, 128, generator=g)
    return ro[sel], rd[sel], near[sel], far[sel], u
