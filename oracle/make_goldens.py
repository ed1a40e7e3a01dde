"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference)
on this container's CPU with seeded synthetic weights / inputs.  TEST INFRASTRUCTURE ONLY.

    python -m oracle.make_goldens            # writes tests/golden/

Weights are never stored: ``humanliff_b200.synth.synth_state_dict`` regenerates them bit-identically
from (seed, tensor name) and they are loaded into the reference with ``strict=True`` -- which also
proves the product's state-dict key set / shapes equal the reference's."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from humanliff_b200 import synth                      # noqa: E402
from oracle import ref_shims                          # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

PROD = dict(image_size=256, in_channels=27, num_channels=192, out_channels=27, num_res_blocks=3,
            num_heads=4, num_heads_upsample=-1, attention_resolutions="32,16,8", dropout=0.0,
            learn_sigma=False, sigma_small=False, class_cond=True, diffusion_steps=1000,
            noise_schedule="linear", timestep_respacing="250", use_kl=False, predict_xstart=False,
            rescale_timesteps=False, rescale_learned_sigmas=False, use_checkpoint=False,
            use_scale_shift_norm=True, cond_type="controlnet", use_3d_aware=False)
TINY = dict(PROD, image_size=32, num_channels=64, num_res_blocks=1, num_heads=2,
            attention_resolutions="16,8")


def build_ref(flags, seed):
    su = ref_shims.import_diffusion()
    model, diffusion = su.create_model_and_diffusion(**flags)
    sd = synth.synth_state_dict({k: v.shape for k, v in model.state_dict().items()}, seed=seed)
    model.load_state_dict(sd, strict=True)
    model.eval()
    return model, diffusion


def inject_noise(noises):
    """Patch torch.randn_like so the reference's p_sample consumes our pre-drawn noise."""
    it = iter(noises)
    orig = torch.randn_like
    torch.randn_like = lambda x, *a, **k: next(it).to(x)
    return orig


def unet_case(name, flags, B, HW, ts, seed_w, loop_steps):
    t0 = time.time()
    model, diffusion = build_ref(flags, seed_w)
    x, x_cond, g = synth.synth_denoise_inputs(B, 27, HW, HW, seed=1234)
    y = torch.arange(B) % 4
    out = {"x": x.numpy(), "x_cond": x_cond.numpy(), "y": y.numpy(), "ts": np.array(ts)}
    with torch.no_grad():
        for t in ts:
            tt = torch.full((B,), t, dtype=torch.int64)
            noise = torch.randn(B, 27, HW, HW, generator=g)
            orig = inject_noise([noise])
            try:
                r = diffusion.p_sample(model, x, x_cond, tt, clip_denoised=True, model_kwargs={"y": y})
            finally:
                torch.randn_like = orig
            # epsilon itself (model called with the ORIGINAL timestep index, respace.py:117-122)
            eps = model(x, torch.tensor(diffusion.timestep_map)[tt], x_cond, y=y)
            out[f"noise_{t}"] = noise.numpy()
            out[f"eps_{t}"] = eps.numpy()
            out[f"sample_{t}"] = r["sample"].numpy()
            out[f"x0_{t}"] = r["pred_xstart"].numpy()
        # short free-running loop: last `loop_steps` of the chain starting from x as x_T
        T = diffusion.num_timesteps
        noises = [torch.randn(B, 27, HW, HW, generator=g) for _ in range(loop_steps)]
        img = x
        orig = inject_noise(noises)
        try:
            for k, i in enumerate(range(T - 1, T - 1 - loop_steps, -1)):
                tt = torch.full((B,), i, dtype=torch.int64)
                img = diffusion.p_sample(model, img, x_cond, tt, clip_denoised=True,
                                         model_kwargs={"y": y})["sample"]
        finally:
            torch.randn_like = orig
        out["loop_steps"] = np.array(loop_steps)
        out["loop_noise"] = torch.stack(noises).numpy()
        out["loop_final"] = img.numpy()
    np.savez(os.path.join(OUT, name), **out)
    print(name, "done in %.1fs" % (time.time() - t0), flush=True)


def ddim_case(name, flags, B, HW, ts, etas, seed_w):
    """ddim_sample of the unmodified reference (gaussian_diffusion.py:484-529) with injected noise; inputs and
    weights are those of the matching unet_case, so epsilon is already in that golden."""
    t0 = time.time()
    model, diffusion = build_ref(flags, seed_w)
    x, x_cond, g = synth.synth_denoise_inputs(B, 27, HW, HW, seed=1234)
    y = torch.arange(B) % 4
    noise = torch.randn(B, 27, HW, HW, generator=torch.Generator().manual_seed(4321))
    out = {"ts": np.array(ts), "etas": np.array(etas), "noise": noise.numpy()}
    with torch.no_grad():
        for t in ts:
            tt = torch.full((B,), t, dtype=torch.int64)
            for eta in etas:
                orig = inject_noise([noise])
                try:
                    r = diffusion.ddim_sample(model, x, tt, x_cond=x_cond, clip_denoised=True,
                                              model_kwargs={"y": y}, eta=eta)
                finally:
                    torch.randn_like = orig
                out[f"sample_{t}_{eta}"] = r["sample"].numpy()
                out[f"x0_{t}_{eta}"] = r["pred_xstart"].numpy()
    np.savez(os.path.join(OUT, name), **out)
    print(name, "done in %.1fs" % (time.time() - t0), flush=True)


def render_case(name, n_rays=1024, seed_w=3):
    t0 = time.time()
    hd = ref_shims.import_hd_renderer()
    torch.manual_seed(0)
    r = hd.Renderer(use_canonical_space=False, triplane_ch=27, smpl_type=None, test=True)
    shapes = {k: v.shape for k, v in r.state_dict().items() if not k.startswith("view_enc")}
    sd = synth.synth_state_dict(shapes, seed=seed_w, weight_gain=1.5)
    r.load_state_dict(sd, strict=False)
    planes = synth.synth_triplane(256, seed=7)
    bounds = torch.tensor(synth.WORLD_BOUNDS)
    ro, rd, near, far, hit = synth.synth_camera_rays(512, 512, azimuth_deg=30.0)
    g = torch.Generator(); g.manual_seed(99)
    # half hitting rays, a quarter random, plus image-border (miss) rays
    hit_idx = torch.nonzero(hit)[:, 0]
    sel = torch.cat([hit_idx[torch.randperm(hit_idx.numel(), generator=g)[:n_rays // 2]],
                     torch.randint(0, ro.shape[0], (n_rays // 2,), generator=g)])
    ro, rd, near, far = ro[sel], rd[sel], near[sel], far[sel]
    u = torch.rand(n_rays, 128, generator=g)
    orig = torch.rand
    torch.rand = lambda *a, **k: u.clone()
    try:
        t = torch.linspace(0., 1., steps=128)
        z = near[None, :, None] * (1. - t) + far[None, :, None] * t
        pts = ro[None, :, None, :] + rd[None, :, None, :] * z[..., :, None]
        tp = {"world_bounds": bounds[None]}
        with torch.no_grad():
            ret = r.render(tp, pts.reshape(1, -1, 3), z, ro[None], rd[None], near[None, :, None],
                           far[None, :, None], planes, 128, False)
    finally:
        torch.rand = orig
    np.savez(os.path.join(OUT, name), rays_o=ro.numpy(), rays_d=rd.numpy(), near=near.numpy(),
             far=far.numpy(), u=u.numpy(), rgb=ret["rgb_map"][0].numpy(), acc=ret["acc_map"][0].numpy(),
             depth=ret["depth_map"][0].numpy(), seed_w=np.array(seed_w))
    print(name, "done in %.1fs" % (time.time() - t0), flush=True)


def render_full_image_case(name, seed_w=3, azimuth=45.0, chunk=16384, seed_u=99, max_chunks=None):
    """A whole 512 x 512 image through the reference renderer in its own chunk order
    (run_nerf_batch.py:29-67: 16 chunks of 16,384 rays, one torch.rand([chunk, 128]) per chunk).  Only the three
    output maps are stored (5 MB); rays and uniforms are regenerated from (azimuth, seed_u)."""
    t0 = time.time()
    hd = ref_shims.import_hd_renderer()
    torch.manual_seed(0)
    r = hd.Renderer(use_canonical_space=False, triplane_ch=27, smpl_type=None, test=True)
    shapes = {k: v.shape for k, v in r.state_dict().items() if not k.startswith("view_enc")}
    sd = synth.synth_state_dict(shapes, seed=seed_w, weight_gain=1.5)
    r.load_state_dict(sd, strict=False)
    planes = synth.synth_triplane(256, seed=7)
    bounds = torch.tensor(synth.WORLD_BOUNDS)
    ro, rd, near, far, hit = synth.synth_camera_rays(512, 512, focal=600.0, azimuth_deg=azimuth)
    n = ro.shape[0]
    g = torch.Generator(); g.manual_seed(seed_u)
    tp = {"world_bounds": bounds[None]}
    t = torch.linspace(0., 1., steps=128)
    outs = {"rgb": [], "acc": [], "depth": []}
    orig = torch.rand
    n_chunks = n // chunk if max_chunks is None else max_chunks
    try:
        for c in range(n_chunks):
            sl = slice(c * chunk, (c + 1) * chunk)
            u = orig(chunk, 128, generator=g)
            torch.rand = lambda *a, **k: u.clone()
            z = near[None, sl, None] * (1. - t) + far[None, sl, None] * t
            pts = ro[None, sl, None, :] + rd[None, sl, None, :] * z[..., :, None]
            with torch.no_grad():
                ret = r.render(tp, pts.reshape(1, -1, 3), z, ro[None, sl], rd[None, sl], near[None, sl, None],
                               far[None, sl, None], planes, 128, False)
            outs["rgb"].append(ret["rgb_map"][0]); outs["acc"].append(ret["acc_map"][0]); outs["depth"].append(ret["depth_map"][0])
            print("  chunk", c, "%.1fs" % (time.time() - t0), flush=True)
    finally:
        torch.rand = orig
    np.savez(os.path.join(OUT, name), rgb=torch.cat(outs["rgb"]).numpy(), acc=torch.cat(outs["acc"]).numpy(),
             depth=torch.cat(outs["depth"]).numpy(), seed_w=np.array(seed_w), seed_u=np.array(seed_u),
             azimuth=np.array(azimuth), chunk=np.array(chunk), n_rays=np.array(n_chunks * chunk))
    print(name, "done in %.1fs" % (time.time() - t0), flush=True)


def render_rn_case(name, n_rays=256, seed_w=3, layer=2):
    """recon_NeRF/lib/renderer.py (the reconstruction-side Renderer that owns tri_planes [N,4,3,9,256,256] and does
    NOT clamp depth, renderer.py:244-295) on the first rays of the render_1024 golden's ray set."""
    t0 = time.time()
    rn = ref_shims.import_rn_renderer()
    torch.manual_seed(0)
    r = rn.Renderer(use_canonical_space=False, num_instances=1, triplane_dim=256, triplane_ch=27, test=True)
    shapes = {k: v.shape for k, v in r.state_dict().items() if not k.startswith("view_enc") and k != "tri_planes"}
    sd = synth.synth_state_dict(shapes, seed=seed_w, weight_gain=1.5)
    r.load_state_dict(sd, strict=False)
    planes = synth.synth_triplane(256, seed=7)
    with torch.no_grad():
        r.tri_planes[0, layer] = planes[0]
    gold = np.load(os.path.join(OUT, "render_1024.npz"))
    ro, rd = torch.from_numpy(gold["rays_o"][:n_rays]), torch.from_numpy(gold["rays_d"][:n_rays])
    near, far = torch.from_numpy(gold["near"][:n_rays]), torch.from_numpy(gold["far"][:n_rays])
    u = torch.from_numpy(gold["u"][:n_rays])
    bounds = torch.tensor(synth.WORLD_BOUNDS)
    orig = torch.rand
    torch.rand = lambda *a, **k: u.clone()
    try:
        t = torch.linspace(0., 1., steps=128)
        z = near[None, :, None] * (1. - t) + far[None, :, None] * t
        pts = ro[None, :, None, :] + rd[None, :, None, :] * z[..., :, None]
        tp = {"world_bounds": bounds[None], "instance_idx": torch.tensor([0]), "cloth_layer_index": torch.tensor([layer])}
        with torch.no_grad():
            ret = r.render(tp, pts.reshape(1, -1, 3), z, ro[None], rd[None], near[None, :, None], far[None, :, None],
                           128, False)
    finally:
        torch.rand = orig
    np.savez(os.path.join(OUT, name), n_rays=np.array(n_rays), layer=np.array(layer), seed_w=np.array(seed_w),
             rgb=ret["rgb_map"][0].numpy(), acc=ret["acc_map"][0].numpy(), depth=ret["depth_map"][0].numpy())
    print(name, "done in %.1fs" % (time.time() - t0), flush=True)


def density_grid_case(name, res=24, seed_w=3):
    """The field of Renderer.extract_geometry (human_diffusion/NeRF/renderer.py:290-318) run by the reference itself:
    the absent `mcubes` is stubbed so that smooth() hands back the raw grid `u` (= -sigma) it was given."""
    hd = ref_shims.import_hd_renderer()
    grabbed = {}
    hd.mcubes.smooth = lambda u: grabbed.setdefault("u", u.copy())
    hd.mcubes.marching_cubes = lambda u, thr: (np.zeros((1, 3), np.float32), np.zeros((1, 3), np.int64))
    torch.manual_seed(0)
    r = hd.Renderer(use_canonical_space=False, triplane_ch=27, smpl_type=None, test=True)
    shapes = {k: v.shape for k, v in r.state_dict().items() if not k.startswith("view_enc")}
    r.load_state_dict(synth.synth_state_dict(shapes, seed=seed_w, weight_gain=1.5), strict=False)
    planes = synth.synth_triplane(256, seed=7)
    tp = {"world_bounds": torch.tensor(synth.WORLD_BOUNDS)[None]}
    r.extract_geometry(tp, tri_planes=planes, resolution=res, threshold=0.0)
    np.savez(os.path.join(OUT, name), u=grabbed["u"], res=np.array(res), seed_w=np.array(seed_w))
    print(name, "done", flush=True)


def contract_case(name):
    """state_dict key -> shape of the reference modules (the checkpoint contract of SURVEY.md 8(b)), as JSON."""
    import json
    su = ref_shims.import_diffusion()
    out = {}
    model, _ = su.create_model_and_diffusion(**PROD)
    out["unet_production"] = [[k, list(v.shape)] for k, v in model.state_dict().items()]
    model, _ = su.create_model_and_diffusion(**dict(TINY, cond_type="", class_cond=False))
    out["unet_tiny_unconditional"] = [[k, list(v.shape)] for k, v in model.state_dict().items()]
    hd = ref_shims.import_hd_renderer()
    r = hd.Renderer(use_canonical_space=False, triplane_ch=27, smpl_type=None, test=True)
    out["renderer_hd"] = [[k, list(v.shape)] for k, v in r.state_dict().items()]
    rn = ref_shims.import_rn_renderer()
    r = rn.Renderer(use_canonical_space=False, num_instances=2, triplane_dim=256, triplane_ch=27, test=True)
    out["renderer_rn_2_instances"] = [[k, list(v.shape)] for k, v in r.state_dict().items()]
    from improved_diffusion import respace
    cases = []
    for T, sec in [(1000, "ddim50"), (1000, "ddim333"), (1000, "ddim7"), (300, "10,10,10"), (10, "20"), (1000, "1000"),
                   (1000, "999"), (100, "5,200"), (1000, "250"), (1000, "1"), (50, "ddim25")]:
        try:
            cases.append([T, sec, sorted(respace.space_timesteps(T, sec))])
        except Exception as e:                                                       # respace.py:30,47
            cases.append([T, sec, type(e).__name__])
    out["space_timesteps_cases"] = cases
    out["model_and_diffusion_defaults"] = su.model_and_diffusion_defaults()          # script_util.py:11-39
    out["NUM_CLASSES"] = su.NUM_CLASSES
    with open(os.path.join(OUT, name), "w") as f:
        json.dump(out, f)
    print(name, "done", {k: (len(v) if hasattr(v, "__len__") else v) for k, v in out.items()}, flush=True)


def variants_case(name):
    """Flag-envelope variants on the tiny model (B = 1, 27 x 32 x 32), straight from the reference:
    (a) unconditional UNet (cond_type='', class_cond=False);  (b) p_sample with rescale_timesteps=True on a 500-step
    schedule respaced to 50 (timesteps reach the model as floats x 1000/500, respace.py:117-122) with FIXED_SMALL."""
    out = {}
    g = torch.Generator(); g.manual_seed(4321)
    x = torch.randn(1, 27, 32, 32, generator=g)
    xc = (0.3 * torch.randn(1, 27, 32, 32, generator=g)).clamp(-1, 1)
    noise = torch.randn(1, 27, 32, 32, generator=g)
    out.update(x=x.numpy(), x_cond=xc.numpy(), noise=noise.numpy())
    with torch.no_grad():
        model, _ = build_ref(dict(TINY, cond_type="", class_cond=False), 3)
        out["a_t"] = np.array(17)
        out["a_eps"] = model(x, torch.tensor([17])).numpy()
        flags = dict(TINY, diffusion_steps=500, timestep_respacing="50", rescale_timesteps=True, sigma_small=True)
        model, diffusion = build_ref(flags, 11)
        tt = torch.tensor([20])
        orig = inject_noise([noise])
        try:
            r = diffusion.p_sample(model, x, xc, tt, clip_denoised=True, model_kwargs={"y": torch.tensor([3])})
        finally:
            torch.randn_like = orig
        out["b_t"] = np.array(20)
        out["b_sample"], out["b_x0"] = r["sample"].numpy(), r["pred_xstart"].numpy()
        out["b_timestep_map"] = np.array(diffusion.timestep_map)
    np.savez(os.path.join(OUT, name), **out)
    print(name, "done", flush=True)


SCHEDULE_GRID = [(ns, ss, rs) for ns in ("linear", "cosine") for ss in (False, True)
                 for rs in ("", "250", "ddim50", "10,15,20")]


def schedule_case(name):
    """The float64 tables of GaussianDiffusion / SpacedDiffusion (gaussian_diffusion.py:118-169, respace.py:7-86)
    for the whole in-scope flag envelope, straight from the reference's objects."""
    su = ref_shims.import_diffusion()
    out = {}
    for k, (ns, ss, rs) in enumerate(SCHEDULE_GRID):
        d = su.create_gaussian_diffusion(steps=1000, learn_sigma=False, sigma_small=ss, noise_schedule=ns,
                                         use_kl=False, predict_xstart=False, rescale_timesteps=False,
                                         rescale_learned_sigmas=False, timestep_respacing=rs)
        out[f"{k}_timestep_map"] = np.array(d.timestep_map)
        for attr in ("betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_recip_alphas_cumprod",
                     "sqrt_recipm1_alphas_cumprod", "posterior_variance", "posterior_log_variance_clipped",
                     "posterior_mean_coef1", "posterior_mean_coef2"):
            out[f"{k}_{attr}"] = np.asarray(getattr(d, attr), dtype=np.float64)
        # the (variance, log-variance) pair p_mean_variance selects (gaussian_diffusion.py:278-291)
        if ss:
            out[f"{k}_model_var"], out[f"{k}_model_logvar"] = d.posterior_variance, d.posterior_log_variance_clipped
        else:
            v = np.append(d.posterior_variance[1], d.betas[1:])
            out[f"{k}_model_var"], out[f"{k}_model_logvar"] = v, np.log(v)
    np.savez(os.path.join(OUT, name), **out)
    print(name, "done", flush=True)


def unet_full_size_case(name, flags, HW, t, seed_w, seed_in=1234):
    """One forward of the production model at the BASELINE resolution (B = 1, 27 x HW x HW).  Only epsilon is
    stored (7 MB at 256^2); inputs are regenerated from `seed_in` by synth.synth_denoise_inputs."""
    t0 = time.time()
    model, diffusion = build_ref(flags, seed_w)
    x, x_cond, _ = synth.synth_denoise_inputs(1, 27, HW, HW, seed=seed_in)
    y = torch.tensor([2])
    with torch.no_grad():
        tt = torch.full((1,), t, dtype=torch.int64)
        eps = model(x, torch.tensor(diffusion.timestep_map)[tt], x_cond, y=y)
    np.savez(os.path.join(OUT, name), eps=eps.numpy(), t=np.array(t), y=y.numpy(), seed_in=np.array(seed_in),
             x_checksum=np.array(float(x.double().sum())), xc_checksum=np.array(float(x_cond.double().sum())))
    print(name, "done in %.1fs" % (time.time() - t0), flush=True)


def render_noimp_case(name, n_rays=256, seed_w=3):
    """Renderer.render with n_importance=0 (human_diffusion/NeRF/renderer.py: the `if n_importance > 0` block is
    skipped, render_core composites the 128 coarse samples) on the first rays of the render_1024 golden's ray set."""
    t0 = time.time()
    hd = ref_shims.import_hd_renderer()
    torch.manual_seed(0)
    r = hd.Renderer(use_canonical_space=False, triplane_ch=27, smpl_type=None, test=True)
    shapes = {k: v.shape for k, v in r.state_dict().items() if not k.startswith("view_enc")}
    r.load_state_dict(synth.synth_state_dict(shapes, seed=seed_w, weight_gain=1.5), strict=False)
    g = np.load(os.path.join(OUT, "render_1024.npz"))
    ro, rd, near, far = (torch.from_numpy(g[k][:n_rays]) for k in ("rays_o", "rays_d", "near", "far"))
    planes = synth.synth_triplane(256, seed=7)
    bounds = torch.tensor(synth.WORLD_BOUNDS)
    t = torch.linspace(0., 1., steps=128)
    z = near[None, :, None] * (1. - t) + far[None, :, None] * t
    pts = ro[None, :, None, :] + rd[None, :, None, :] * z[..., :, None]
    with torch.no_grad():
        ret = r.render({"world_bounds": bounds[None]}, pts.reshape(1, -1, 3), z, ro[None], rd[None], near[None, :, None],
                       far[None, :, None], planes, 0, False)
    np.savez(os.path.join(OUT, name), n_rays=np.array(n_rays), rgb=ret["rgb_map"][0].numpy(),
             acc=ret["acc_map"][0].numpy(), depth=ret["depth_map"][0].numpy(), seed_w=np.array(seed_w))
    print(name, "done in %.1fs" % (time.time() - t0), flush=True)


def knn_points_bruteforce(p1, p2, K=1):
    """Stand-in for pytorch3d.ops.knn.knn_points (pytorch3d 0.7.x, absent from this image; the reference calls it with
    K=1, renderer.py:62).  Published behaviour restated: squared Euclidean distances in fp32, the K smallest per query
    in ascending order, lowest index on ties -> (dists [N,P1,K], idx [N,P1,K] int64, None)."""
    assert K == 1
    ds, ids = [], []
    for s in range(0, p1.shape[1], 8192):
        q = p1[:, s:s + 8192]
        d = ((q[:, :, None, :] - p2[:, None, :, :]) ** 2).sum(-1)
        m, i = d.min(-1)
        ds.append(m[..., None]); ids.append(i[..., None])
    return torch.cat(ds, 1), torch.cat(ids, 1), None


def render_canon_case(name, n_rays=384, seed_w=3, seed_smpl=5, seed_pose=21):
    """use_canonical_space=True (the TightCap branch of triplane_sample_layered.py:73-76): the unmodified
    human_diffusion/NeRF/renderer.py -- deform_target2c / deform_target2c_op (:52-133), get_transform_params_torch,
    get_rigid_transformation_torch, batch_rodrigues, SMPL_to_tensor -- on a seeded SMPL-shaped asset
    (synth.synth_smpl; the licensed SMPL_NEUTRAL.pkl ships with neither repo) and a brute-force stand-in for
    pytorch3d's knn_points.  Also stores the per-point intermediates of the first 8 rays' coarse points."""
    t0 = time.time()
    hd = ref_shims.import_hd_renderer()
    smpl = synth.synth_smpl(seed_smpl)
    saved = (hd.read_pickle, hd.SMPL_to_tensor, hd.knn_points, torch.Tensor.cuda)
    to_tensor = hd.SMPL_to_tensor
    hd.read_pickle = lambda path: {k: v.copy() for k, v in smpl.items()}
    hd.SMPL_to_tensor = lambda params, device=None: to_tensor(params, torch.device("cpu"))
    hd.knn_points = knn_points_bruteforce
    torch.Tensor.cuda = lambda self, *a, **k: self
    if not torch.cuda.is_available():
        torch.cuda.current_device = lambda: 0
    orig_rand = torch.rand
    try:
        torch.manual_seed(0)
        r = hd.Renderer(use_canonical_space=True, triplane_ch=27, smpl_type="smpl", test=True)
        shapes = {k: v.shape for k, v in r.state_dict().items() if not k.startswith("view_enc")}
        r.load_state_dict(synth.synth_state_dict(shapes, seed=seed_w, weight_gain=1.5), strict=False)
        planes = synth.synth_triplane(256, seed=7)
        tp = synth.synth_canonical_frame(smpl, seed_pose)
        ro, rd, near, far, u = synth.synth_canonical_rays(tp, n_rays)
        torch.rand = lambda *a, **k: u.clone()
        t = torch.linspace(0., 1., steps=128)
        z = near[None, :, None] * (1. - t) + far[None, :, None] * t
        pts = ro[None, :, None, :] + rd[None, :, None, :] * z[..., :, None]
        with torch.no_grad():
            ret = r.render(tp, pts.reshape(1, -1, 3), z, ro[None], rd[None], near[None, :, None], far[None, :, None],
                           planes, 128, False)
            vd = (rd / rd.norm(dim=-1, keepdim=True))[:8, None].expand(8, 128, 3).reshape(1, -1, 3)
            cpts, cdirs, _ = r.deform_target2c(tp, pts[:, :8].reshape(1, -1, 3), vd)
    finally:
        torch.rand = orig_rand
        hd.read_pickle, hd.SMPL_to_tensor, hd.knn_points, torch.Tensor.cuda = saved
    np.savez(os.path.join(OUT, name), n_rays=np.array(n_rays), seed_w=np.array(seed_w), seed_smpl=np.array(seed_smpl),
             seed_pose=np.array(seed_pose), rgb=ret["rgb_map"][0].numpy(), acc=ret["acc_map"][0].numpy(),
             depth=ret["depth_map"][0].numpy(), canonical_pts=cpts[0].numpy(), canonical_dirs=cdirs[0].numpy())
    print(name, "done in %.1fs" % (time.time() - t0), "acc mean %.3f" % float(ret["acc_map"].mean()), flush=True)


def loop_noise(k, shape, seed=9000):
    """Per-step Gaussian of the long free-running goldens: regenerated from (seed + k), never stored."""
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed + k))


def unet_long_loop_case(name, flags, HW, steps, seed_w):
    """>= 50 free-running p_sample steps of the PRODUCTION architecture (SURVEY 8(d) parity metric (ii)), B = 1 at
    HW x HW: the first `steps` steps of the chain (t = T-1 ... T-steps) from x_T = synth x.  Only the final sample is
    stored; x_T, x_cond and the per-step noise are regenerated from their seeds (loop_noise)."""
    t0 = time.time()
    model, diffusion = build_ref(flags, seed_w)
    x, x_cond, _ = synth.synth_denoise_inputs(1, 27, HW, HW, seed=1234)
    y = torch.tensor([2])
    T = diffusion.num_timesteps
    noises = [loop_noise(k, x.shape) for k in range(steps)]
    img = x
    orig = inject_noise(noises)
    try:
        with torch.no_grad():
            for i in range(T - 1, T - 1 - steps, -1):
                img = diffusion.p_sample(model, img, x_cond, torch.full((1,), i, dtype=torch.int64), clip_denoised=True,
                                         model_kwargs={"y": y})["sample"]
    finally:
        torch.randn_like = orig
    np.savez(os.path.join(OUT, name), steps=np.array(steps), y=y.numpy(), loop_final=img.numpy(),
             respacing=np.array(flags["timestep_respacing"]))
    print(name, "done in %.1fs" % (time.time() - t0), flush=True)


SUB = (slice(None), slice(None), slice(1, None, 4), slice(2, None, 4))      # the stored sub-lattice of a 256^2 epsilon


def unet_full_size_sweep_case(name, flags, HW, ts, seed_w, seed_in=1234):
    """The production model at the BASELINE resolution on the 1000-step ("") schedule at several timesteps (the
    sweep of SURVEY 8(d) config 1, at config 2's size).  To keep the fixture small only the sub-lattice
    eps[:, :, 1::4, 2::4] is stored (1/16 of the pixels, 27 x 64 x 64 per timestep): an unbiased sample of the
    rel-L2 / max-norm statistics the GPU test evaluates on the same sub-lattice."""
    t0 = time.time()
    model, diffusion = build_ref(flags, seed_w)
    x, x_cond, _ = synth.synth_denoise_inputs(1, 27, HW, HW, seed=seed_in)
    y = torch.tensor([2])
    out = {"ts": np.array(ts), "seed_in": np.array(seed_in), "y": y.numpy()}
    with torch.no_grad():
        for t in ts:
            eps = model(x, torch.tensor([t]), x_cond, y=y)
            out[f"eps_sub_{t}"] = eps[SUB].numpy()
            out[f"eps_norm_{t}"] = np.array(float(eps.double().norm()))
    np.savez(os.path.join(OUT, name), **out)
    print(name, "done in %.1fs" % (time.time() - t0), flush=True)


if __name__ == "__main__":
    assert ref_shims.available(), "reference tree not found"
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    which = sys.argv[1:] or ["tiny", "render", "prod64", "ddim", "schedules", "variants", "render_rn", "density", "contract",
                             "render_noimp"]
    if "tiny" in which:
        unet_case("unet_tiny_32.npz", TINY, B=2, HW=32, ts=[0, 100, 249], seed_w=11, loop_steps=12)
    if "render" in which:
        render_case("render_1024.npz")
    if "ddim" in which:
        ddim_case("ddim_tiny_32.npz", TINY, B=2, HW=32, ts=[0, 100, 249], etas=[0.0, 0.5], seed_w=11)
    if "prod64" in which:
        unet_case("unet_prod_64.npz", PROD, B=1, HW=64, ts=[0, 100, 249], seed_w=0, loop_steps=6)
    if "contract" in which:
        contract_case("state_dict_contract.json")
    if "density" in which:
        density_grid_case("density_grid_24.npz")
    if "render_rn" in which:
        render_rn_case("render_rn_256.npz")
    if "variants" in which:
        variants_case("unet_variants_32.npz")
    if "schedules" in which:
        schedule_case("schedules.npz")
    if "render_noimp" in which:
        render_noimp_case("render_noimp_256.npz")
    if "render_canon" in which:   # not in the default list: ~1 min of CPU (brute-force nearest vertex)
        render_canon_case("render_canon_384.npz")
    if "loop50" in which:         # not in the default list: ~40 s of CPU
        unet_long_loop_case("unet_prod_64_loop50.npz", PROD, HW=64, steps=50, seed_w=0)
    if "prod256sweep" in which:   # not in the default list: ~1 min of CPU, 1.8 MB
        unet_full_size_sweep_case("unet_prod_256_sweep.npz", dict(PROD, timestep_respacing=""), HW=256,
                                  ts=[0, 1, 500, 999], seed_w=0)
    if "render512" in which:      # not in the default list: several minutes of CPU, 5 MB
        render_full_image_case("render_512x512.npz")
    if "prod256" in which:        # not in the default list: ~1 min of CPU, 7 MB
        unet_full_size_case("unet_prod_256_eps.npz", PROD, HW=256, t=100, seed_w=0)
