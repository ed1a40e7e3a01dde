"""GPU parity of the fused render kernel vs the golden frozen from the unmodified reference
(human_diffusion/NeRF/renderer.py run on CPU) and vs the oracle on a fresh ray set.
Tolerance 1e-3 rel-L2 per output map (north_star); observed error is reported in the assertion."""
import pytest
import torch

from common import load_golden, renderer_state_dict, rel_l2, rel_max
from humanliff_b200 import synth

pytestmark = pytest.mark.gpu


def _setup(precision="fp16"):
    g = load_golden("render_1024.npz")
    r, sd = renderer_state_dict(int(g["seed_w"]), precision)
    return r.to("cuda:0"), sd, g, synth.synth_triplane(256, seed=7), torch.tensor(synth.WORLD_BOUNDS)


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("fp16", 3e-4), ("fp16_mma", 3e-4)])
def test_render_vs_reference_golden(precision, tol):
    """fp32 = exact CUDA-core MLP; fp16 = tcgen05 MLP (activations in tensor memory), fp16_mma = the mma.sync MLP
    (operand rounding averaged over 256 samples per ray: the CPU emulation of the same numerics gives rgb 1.2e-5 /
    depth 4e-5).  north_star bar: 1e-3."""
    r, sd, g, planes, bounds = _setup(precision)
    dev = torch.device("cuda:0")
    rgb, acc, depth = r.render_rays(planes[0].to(dev), bounds, g["rays_o"].to(dev), g["rays_d"].to(dev),
                                    g["near"].to(dev), g["far"].to(dev), u=g["u"].to(dev))
    for name, a, b in (("rgb", rgb, g["rgb"]), ("acc", acc, g["acc"]), ("depth", depth, g["depth"])):
        e = rel_l2(a, b)
        assert e < tol, f"{name}: rel-L2 {e:.3e} max {rel_max(a, b):.3e}"
    assert float((acc.cpu() - 1).abs().max()) < 1e-3      # reference quirk: acc ~ 1.00002 on every ray


@pytest.mark.parametrize("precision", ["fp16", "fp16_mma"])
def test_render_tensor_core_equals_cuda_core_kernel(precision):
    """Tensor-core and exact kernels on a full-resolution ray block, in-kernel uniforms with the same seed: the only
    difference is the fp16 rounding of the MLP operands.  An odd ray count exercises the ragged last CTA / group."""
    dev = torch.device("cuda:0")
    r16, _, g, planes, bounds = _setup(precision)
    r32, _, _, _, _ = _setup("fp32")
    ro, rd, near, far, hit = synth.synth_camera_rays(128, 128, focal=150.0, azimuth_deg=70.0)
    ro, rd, near, far = ro[:16383], rd[:16383], near[:16383], far[:16383]
    a = r16.render_rays(planes[0].to(dev), bounds, ro.to(dev), rd.to(dev), near.to(dev), far.to(dev), u=None, seed=5)
    b = r32.render_rays(planes[0].to(dev), bounds, ro.to(dev), rd.to(dev), near.to(dev), far.to(dev), u=None, seed=5)
    for name, x, y in zip(("rgb", "acc", "depth"), a, b):
        assert not torch.isnan(x).any()
        assert rel_l2(x, y) < 3e-4, (name, rel_l2(x, y))


def test_tcgen05_and_mma_kernels_agree_closely():
    """Same operand rounding (fp16 features / activations / weights, fp32 accumulate) in both tensor-core kernels:
    only the accumulation order inside the MMA differs."""
    dev = torch.device("cuda:0")
    r5, _, g, planes, bounds = _setup("fp16")
    rm, _, _, _, _ = _setup("fp16_mma")
    args = (planes[0].to(dev), bounds, g["rays_o"].to(dev), g["rays_d"].to(dev), g["near"].to(dev), g["far"].to(dev))
    a = r5.render_rays(*args, u=g["u"].to(dev))
    b = rm.render_rays(*args, u=g["u"].to(dev))
    for name, x, y in zip(("rgb", "acc", "depth"), a, b):
        assert rel_l2(x, y) < 5e-5, (name, rel_l2(x, y))


def test_two_different_triplanes_in_a_row():
    """a texel cache keyed on the tensor address served plane A's texels for plane B when the
    allocator reused the address.  Render A, free it, render a same-shaped B: outputs must differ and match the oracle."""
    from oracle import render_oracle
    r, sd, g, _, bounds = _setup("fp16")
    dev = torch.device("cuda:0")
    n = 128
    ray = [g[k][:n] for k in ("rays_o", "rays_d", "near", "far")]
    outs = []
    for seed in (7, 8):
        planes = synth.synth_triplane(256, seed=seed)
        pd = planes[0].to(dev)
        rgb, acc, dep = r.render_rays(pd, bounds, *[t.to(dev) for t in ray], u=g["u"][:n].to(dev))
        ref = render_oracle.render_rays(sd, planes[0], bounds, *ray, g["u"][:n])
        assert rel_l2(rgb, ref[0]) < 1e-4 and rel_l2(dep, ref[2]) < 3e-4, seed
        outs.append(rgb.clone())
        del pd, rgb, acc, dep                                   # the next plane may land on the same address
    assert rel_l2(outs[0], outs[1]) > 1e-3      # colours saturate near 1 on this MLP: 2.5e-3 between the planes, 1e-5 to each oracle


def test_n_importance_zero_vs_reference_golden():
    """`if n_importance > 0` (recon_NeRF/lib/renderer.py:258): with 0 the coarse pass and sample_pdf are skipped and
    the 128 coarse samples are composited.  Golden from the unmodified reference + the oracle on the same rays."""
    from humanliff_b200 import render
    from oracle import render_oracle
    r, sd, g, planes, bounds = _setup("fp16")
    gz = load_golden("render_noimp_256.npz")
    dev = torch.device("cuda:0")
    n = int(gz["n_rays"])
    ray = [g[k][:n] for k in ("rays_o", "rays_d", "near", "far")]
    ref = render_oracle.render_rays(sd, planes[0], bounds, *ray, None, n_importance=0)
    for name, a, b in zip(("rgb", "acc", "depth"), ref, (gz["rgb"], gz["acc"], gz["depth"])):
        assert rel_l2(a, b) < 2e-6, (name, rel_l2(a, b))        # the oracle restates the reference here too
    tp = {"world_bounds": bounds[None].to(dev)}
    lst = render(rays_o=ray[0][None].to(dev), rays_d=ray[1][None].to(dev), near=ray[2][None].to(dev),
                 far=ray[3][None].to(dev), tri_planes=planes.to(dev), tp_input=tp, renderer=r, n_samples=128,
                 perturb=0., n_importance=0, white_bkgd=False)
    for name, a, b in (("rgb", lst[0][0], gz["rgb"]), ("acc", lst[1][0], gz["acc"]), ("depth", lst[3][0], gz["depth"])):
        assert rel_l2(a, b) < 3e-4, (name, rel_l2(a, b))


def test_render_reference_shaped_api_and_script_helper():
    """Renderer.render(...) dict (HD signature) and the script-level render() list, vs the oracle."""
    from humanliff_b200 import render
    from oracle import render_oracle
    r, sd, g, planes, bounds = _setup()
    dev = torch.device("cuda:0")
    ro, rd, near, far, hit = synth.synth_camera_rays(64, 64, focal=75.0, azimuth_deg=200.0)
    n = ro.shape[0]
    u = torch.rand(n, 128, generator=torch.Generator().manual_seed(4))
    ref_rgb, ref_acc, ref_dep = render_oracle.render_rays(sd, planes[0], bounds, ro, rd, near, far, u, clamp_depth=True)
    tp = {"world_bounds": bounds[None].to(dev)}
    t = torch.linspace(0., 1., 128)
    z = near[None, :, None] * (1 - t) + far[None, :, None] * t
    out = r.render(tp, None, z.to(dev), ro[None].to(dev), rd[None].to(dev), near[None, :, None].to(dev),
                   far[None, :, None].to(dev), planes.to(dev), 128, False, u=u.to(dev))
    assert out["rgb_map"].shape == (1, n, 3) and out["acc_map"].shape == (1, n) and out["depth_map"].shape == (1, n)
    assert rel_l2(out["rgb_map"][0], ref_rgb) < 1e-3
    assert rel_l2(out["depth_map"][0], ref_dep) < 1e-3
    assert out["normal_map"].data_ptr() == out["rgb_map"].data_ptr() or torch.equal(out["normal_map"], out["rgb_map"])
    lst = render(chunk=64 * 64 // 16, rays_o=ro[None].to(dev), rays_d=rd[None].to(dev), near=near[None].to(dev),
                 far=far[None].to(dev), tri_planes=planes.to(dev), tp_input=tp, renderer=r, n_samples=128,
                 perturb=0., n_importance=128, white_bkgd=False, u=u.to(dev))
    assert len(lst) == 4 and rel_l2(lst[0][0], ref_rgb) < 1e-3 and rel_l2(lst[1][0], ref_acc) < 1e-3
    assert rel_l2(lst[3][0], ref_dep) < 1e-3
    # rays that miss the box see only empty space: density is softplus(bias-only MLP) everywhere, and the
    # reference's 1e10 last interval still makes the final sample opaque -> acc ~ 1
    assert float((lst[1][0] - 1).abs().max()) < 1e-3
    with pytest.raises(NotImplementedError):
        render(rays_o=ro[None].to(dev), rays_d=rd[None].to(dev), near=near[None].to(dev), far=far[None].to(dev),
               tri_planes=planes.to(dev), tp_input=tp, renderer=r, n_samples=64, n_importance=64)


def test_render_recon_variant_no_depth_clamp():
    from humanliff_b200.renderer import ReconRenderer
    from oracle import render_oracle
    r, sd, g, planes, bounds = _setup()
    rr = ReconRenderer(triplane_ch=27, triplane_dim=256, num_instances=1, test=True)
    rr.load_state_dict(sd, strict=False)
    with torch.no_grad():
        rr.tri_planes[0, 2] = planes[0]
    rr = rr.to("cuda:0")
    dev = torch.device("cuda:0")
    n = 256
    tp = {"world_bounds": bounds[None].to(dev), "instance_idx": torch.tensor([0]), "cloth_layer_index": torch.tensor([2])}
    t = torch.linspace(0., 1., 128)
    z = g["near"][None, :n, None] * (1 - t) + g["far"][None, :n, None] * t
    out = rr.module.render(tp, None, z.to(dev), g["rays_o"][None, :n].to(dev), g["rays_d"][None, :n].to(dev),
                           g["near"][None, :n, None].to(dev), g["far"][None, :n, None].to(dev), 128, False,
                           u=g["u"][:n].to(dev))
    ref = render_oracle.render_rays(sd, planes[0], bounds, g["rays_o"][:n], g["rays_d"][:n], g["near"][:n],
                                    g["far"][:n], g["u"][:n], clamp_depth=False)
    assert rel_l2(out["rgb_map"][0], ref[0]) < 1e-3 and rel_l2(out["depth_map"][0], ref[2]) < 1e-3
    # and against the reconstruction-side reference renderer itself (recon_NeRF/lib/renderer.py, golden render_rn_256)
    gr = load_golden("render_rn_256.npz")
    assert int(gr["n_rays"]) == n and int(gr["layer"]) == 2
    for name, a, b in (("rgb", out["rgb_map"][0], gr["rgb"]), ("acc", out["acc_map"][0], gr["acc"]),
                       ("depth", out["depth_map"][0], gr["depth"])):
        assert rel_l2(a, b) < 3e-4, (name, rel_l2(a, b))


@pytest.mark.parametrize("precision", ["fp16", "fp16_mma"])
def test_density_grid_vs_oracle(precision):
    """extract_geometry's field (human_diffusion/NeRF/renderer.py:290-318): -sigma on linspace(min, max, res)^3."""
    from oracle import render_oracle
    r, sd, g, planes, bounds = _setup(precision)
    dev = torch.device("cuda:0")
    res = 20
    u = r.density_grid({"world_bounds": bounds[None].to(dev)}, planes.to(dev), resolution=res)
    assert u.shape == (res, res, res)
    xs, ys, zs = (torch.linspace(float(bounds[0, i]), float(bounds[1, i]), res) for i in range(3))
    xx, yy, zz = torch.meshgrid(xs, ys, zs, indexing="ij")
    pts = torch.stack([xx.reshape(-1), yy.reshape(-1), zz.reshape(-1)], -1)
    ref = -render_oracle.mlp(sd, render_oracle.plane_features(planes[0], pts, bounds[0], bounds[1])).reshape(res, res, res)
    assert rel_l2(u, ref) < 1e-3, rel_l2(u, ref)
    assert rel_max(u, ref) < 2e-3


def test_density_grid_vs_reference_golden():
    """Renderer.density_grid against the grid the reference's own extract_geometry evaluated (mcubes stubbed to hand
    the raw field back): golden density_grid_24, index order [x, y, z], values -sigma (pre-softplus)."""
    gd = load_golden("density_grid_24.npz")
    r, sd = renderer_state_dict(int(gd["seed_w"]), "fp16")
    r = r.to("cuda:0")
    dev = torch.device("cuda:0")
    planes = synth.synth_triplane(256, seed=7)
    bounds = torch.tensor(synth.WORLD_BOUNDS)
    res = int(gd["res"])
    u = r.density_grid({"world_bounds": bounds[None].to(dev)}, planes.to(dev), resolution=res)
    assert u.shape == (res, res, res)
    assert rel_l2(u, gd["u"]) < 1e-3, rel_l2(u, gd["u"])
    assert rel_max(u, gd["u"]) < 2e-3


# ------------------------------------------------------------------------------------------------------------------
# use_canonical_space=True (the TightCap branch of triplane_sample_layered.py:73-76)
def _canon_setup(precision="fp16"):
    from humanliff_b200.renderer import Renderer
    gz = load_golden("render_canon_384.npz")
    asset = synth.synth_smpl(int(gz["seed_smpl"]))
    r = Renderer(use_canonical_space=True, triplane_ch=27, test=True, smpl=asset, precision=precision)
    shapes = {k: v.shape for k, v in r.state_dict().items() if not k.startswith("view_enc")}
    sd = synth.synth_state_dict(shapes, seed=int(gz["seed_w"]), weight_gain=1.5)
    r.load_state_dict(sd, strict=False)
    tp = synth.synth_canonical_frame(asset, int(gz["seed_pose"]))
    return r.to("cuda:0"), sd, gz, asset, tp


def _to_dev(tp, dev):
    mv = lambda v: {k: mv(x) for k, x in v.items()} if isinstance(v, dict) else v.to(dev)
    return mv(tp)


def test_canonical_deformation_vs_reference_golden():
    """Renderer.deform_target2c (hl_smpl_vertex_tables + hl_canonical_points) against the per-point canonical positions
    and directions the unmodified reference computed for 1024 coarse points; then 200k random points in the posed box
    against the oracle: the nearest vertex is discontinuous, so a point within rounding of a cell boundary may pick the
    other vertex -- those are counted (must be < 1e-4 of the points) and the rest must agree to 2e-5 of the box size."""
    from oracle import render_oracle
    r, _, gz, asset, tp = _canon_setup()
    dev = torch.device("cuda:0")
    n = int(gz["n_rays"])
    ro, rd, near, far, _ = synth.synth_canonical_rays(tp, n)
    t = torch.linspace(0., 1., steps=128)
    z = near[:8, None] * (1. - t) + far[:8, None] * t
    pts = (ro[:8, None] + rd[:8, None] * z[..., None]).reshape(1, -1, 3)
    vd = (rd / rd.norm(dim=-1, keepdim=True))[:8, None].expand(8, 128, 3).reshape(1, -1, 3)
    c, cd, box = r.deform_target2c(_to_dev(tp, dev), pts.to(dev), vd.to(dev))
    assert torch.equal(box.cpu(), tp["t_world_bounds"])
    ep = (c[0].cpu() - gz["canonical_pts"]).abs().max(-1).values
    ed = (cd[0].cpu() - gz["canonical_dirs"]).abs().max(-1).values
    assert int((ep > 2e-5).sum()) == 0 and int((ed > 2e-5).sum()) == 0, (float(ep.max()), float(ed.max()))
    # bulk check vs the oracle
    g = torch.Generator(); g.manual_seed(4)
    wb = tp["world_bounds"][0]
    P = wb[0] + (wb[1] - wb[0]) * torch.rand(200_000, 3, generator=g)
    c2, _, _ = r.deform_target2c(_to_dev(tp, dev), P[None].to(dev))
    ref, _ = render_oracle.deform_to_canonical(render_oracle.smpl_tensors(asset), tp, P)
    err = (c2[0].cpu() - ref).abs().max(-1).values
    flips = int((err > 4e-5).sum())
    assert flips < 20, (flips, float(err.max()))


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("fp16", 3e-4)])
def test_canonical_render_vs_reference_golden(precision, tol):
    """Rendered maps of 384 rays with every sample deformed to the canonical space (golden: the unmodified
    human_diffusion/NeRF/renderer.py with use_canonical_space=True on the seeded SMPL-shaped asset), through the
    reference-shaped Renderer.render and the script-level render().  fp16 = the tcgen05 kernel (per-sample view encoding
    in the constant tile), fp32 = the exact CUDA-core kernel."""
    from humanliff_b200.renderer import render as script_render
    r, _, gz, asset, tp = _canon_setup(precision)
    dev = torch.device("cuda:0")
    n = int(gz["n_rays"])
    ro, rd, near, far, u = synth.synth_canonical_rays(tp, n)
    t = torch.linspace(0., 1., steps=128)
    z = near[None, :, None] * (1. - t) + far[None, :, None] * t
    tpd = _to_dev(tp, dev)
    out = r.render(tpd, None, z.to(dev), ro[None].to(dev), rd[None].to(dev), near[None, :, None].to(dev),
                   far[None, :, None].to(dev), synth.synth_triplane(256, seed=7).to(dev), 128, False, u=u.to(dev))
    for name, key in (("rgb", "rgb_map"), ("acc", "acc_map"), ("depth", "depth_map")):
        e = rel_l2(out[key][0], gz[name])
        assert e < tol, f"{name}: rel-L2 {e:.3e} max {rel_max(out[key][0], gz[name]):.3e}"
    lst = script_render(rays_o=ro[None].to(dev), rays_d=rd[None].to(dev), near=near[None].to(dev), far=far[None].to(dev),
                        tri_planes=synth.synth_triplane(256, seed=7).to(dev), tp_input=tpd, renderer=r, n_samples=128,
                        n_importance=128, u=u.to(dev))
    assert rel_l2(lst[0][0], gz["rgb"]) < tol and rel_l2(lst[3][0], gz["depth"]) < tol


def test_canonical_density_grid_vs_oracle():
    """extract_geometry's field with use_canonical_space=True (renderer.py:290-318): grid points of the posed box are
    deformed, then looked up inside t_world_bounds."""
    from oracle import render_oracle
    r, sd, gz, asset, tp = _canon_setup("fp32")
    dev = torch.device("cuda:0")
    res = 20
    planes = synth.synth_triplane(256, seed=7)
    got = r.density_grid(_to_dev(tp, dev), planes.to(dev), resolution=res).cpu()
    wb = tp["world_bounds"][0]
    ax = [torch.linspace(float(wb[0, i]), float(wb[1, i]), res) for i in range(3)]
    P = torch.stack(torch.meshgrid(*ax, indexing="ij"), -1).reshape(-1, 3)
    c, _ = render_oracle.deform_to_canonical(render_oracle.smpl_tensors(asset), tp, P)
    tb = tp["t_world_bounds"][0]
    want = -render_oracle.mlp(sd, render_oracle.plane_features(planes[0], c, tb[0], tb[1])).reshape(res, res, res)
    bad = ((got - want).abs() > 1e-3 * want.abs().max()).sum()
    assert int(bad) <= 2, (int(bad), rel_l2(got, want))


def test_canonical_recon_variant_and_n_importance_zero_vs_oracle():
    """The reconstruction-side renderer (recon_NeRF/lib/renderer.py: owns tri_planes, no depth clamp) in canonical-space
    mode, with n_importance = 128 and 0 (coarse samples only), against the oracle; two frames in one batch."""
    from humanliff_b200.renderer import ReconRenderer
    from oracle import render_oracle
    gz = load_golden("render_canon_384.npz")
    asset = synth.synth_smpl(int(gz["seed_smpl"]))
    r = ReconRenderer(use_canonical_space=True, num_instances=1, triplane_dim=256, triplane_ch=27, test=True, smpl=asset)
    shapes = {k: v.shape for k, v in r.state_dict().items() if not k.startswith("view_enc") and k != "tri_planes"}
    sd = synth.synth_state_dict(shapes, seed=int(gz["seed_w"]), weight_gain=1.5)
    r.load_state_dict(sd, strict=False)
    planes = synth.synth_triplane(256, seed=7)
    with torch.no_grad():
        r.tri_planes[0, 1] = planes[0]
    dev = torch.device("cuda:0")
    r = r.to(dev)
    frames = [synth.synth_canonical_frame(asset, s) for s in (21, 22)]
    cat = lambda key: torch.cat([f[key] for f in frames], 0)
    tp = {"params": {k: torch.cat([f["params"][k] for f in frames], 0) for k in frames[0]["params"]},
          "t_params": {k: torch.cat([f["t_params"][k] for f in frames], 0) for k in frames[0]["t_params"]},
          "vertices": cat("vertices"), "world_bounds": cat("world_bounds"), "t_world_bounds": cat("t_world_bounds"),
          "instance_idx": torch.tensor([0, 0]), "cloth_layer_index": torch.tensor([1, 1])}
    n = 96
    rays = [synth.synth_canonical_rays(f, n, seed=5 + i) for i, f in enumerate(frames)]
    st = lambda j: torch.stack([rr[j] for rr in rays], 0)
    ro, rd, near, far, u = st(0), st(1), st(2), st(3), torch.cat([rr[4] for rr in rays], 0)
    t = torch.linspace(0., 1., steps=128)
    z = near[:, :, None] * (1. - t) + far[:, :, None] * t
    mv = lambda v: {k: mv(x) for k, x in v.items()} if isinstance(v, dict) else v.to(dev)
    smpl = render_oracle.smpl_tensors(asset)
    for n_imp in (128, 0):
        out = r.render(mv(tp), None, z.to(dev), ro.to(dev), rd.to(dev), near[..., None].to(dev), far[..., None].to(dev),
                       n_importance=n_imp, u=u.to(dev))
        for b, f in enumerate(frames):
            ref = render_oracle.render_rays(sd, planes[0], f["t_world_bounds"][0], ro[b], rd[b], near[b], far[b],
                                            u[b * n:(b + 1) * n], clamp_depth=False, n_importance=n_imp, canon=(smpl, f))
            for name, key, want in (("rgb", "rgb_map", ref[0]), ("acc", "acc_map", ref[1]), ("depth", "depth_map", ref[2])):
                e = rel_l2(out[key][b], want)
                assert e < 3e-4, f"n_importance={n_imp} frame {b} {name}: rel-L2 {e:.3e}"


def _pair_layout(t):
    """[n, 4] (x, y, z, w) rows -> the two-entries-per-pair-of-float4 layout of the search tables (canon.cuh)."""
    import numpy as np
    t = np.asarray(t, dtype=np.float32).reshape(-1, 2, 4)             # [pairs, entry, xyzw]
    a = np.stack([t[:, 0, 0], t[:, 1, 0], t[:, 0, 1], t[:, 1, 1]], -1)
    b = np.stack([t[:, 0, 2], t[:, 1, 2], t[:, 0, 3], t[:, 1, 3]], -1)
    return np.stack([a, b], 1).reshape(-1, 4)


def test_nearest_vertex_decisions_are_separately_rounded_with_ties_to_the_lower_index():
    """hl_canonical_points on a hand-built table (the C ABI takes the tables as plain pointers).  Every query has exactly two
    candidate vertices at (nearly) the same distance, either in one cluster or in two; everything else is 3 units away.
    * adversarial pairs: found by a seeded search so that the separately rounded distance ((dx^2 + dy^2) + dz^2, each
      operation rounded: the oracle's / the torch stand-in's arithmetic) and the fused one (fma(dz, dz, fma(dy, dy, dx^2)):
      what the packed filter computes, and what ptxas makes of packed mul + add) DISAGREE about the nearer vertex -- the kernel must side with the separately rounded one;
    * exact ties (v = q +- delta, exactly representable): the lower vertex index wins although the higher one is scanned
      first (earlier slot / earlier cluster).
    The per-vertex affine is M = 0, c = (index, 0, 0), so the returned canonical x IS the chosen vertex."""
    import ctypes
    import numpy as np
    from common import nearest_vertex_near_tie_cases, sq_dist_fused, sq_dist_separate
    from humanliff_b200._lib import call
    f32 = np.float32
    d_sep, d_fma = sq_dist_separate, sq_dist_fused
    NC, CL = 64, 4
    n_same_adv, n_same_tie, n_cross_adv, n_cross_tie = 24, 8, 12, 4
    adv_all, ties = nearest_vertex_near_tie_cases(n_same_adv + n_cross_adv, n_same_tie + n_cross_tie, seed=11)
    adv, cross_adv = adv_all[:n_same_adv], adv_all[n_same_adv:]
    verts = np.full((NC * CL, 4), 1e18, dtype=f32)
    idx = np.full(NC * CL, 0x7fffffff, dtype=np.int32)
    queries, expect, fused_would_differ = [], [], 0

    def place(case, slot_a, slot_b, ia, ib):
        nonlocal fused_would_differ
        q, va, vb = case
        verts[slot_a, :3], verts[slot_b, :3] = va, vb
        idx[slot_a], idx[slot_b] = ia, ib
        key = lambda v, i, dist: (float(dist(q, v)), i)
        want = min((key(va, ia, d_sep), ia), (key(vb, ib, d_sep), ib))[1]
        fused = min((key(va, ia, d_fma), ia), (key(vb, ib, d_fma), ib))[1]
        fused_would_differ += int(want != fused)
        queries.append(q); expect.append(want)

    # same cluster: slot 0 (scanned first) carries the HIGHER index
    for c, case in enumerate(adv + ties[:n_same_tie]):
        place(case, c * CL, c * CL + 1, 2 * c + 1, 2 * c)
    # two clusters: the earlier cluster carries the HIGHER index
    for j, case in enumerate(cross_adv + ties[n_same_tie:]):
        # (odd j: the first vertex sits in a cluster of the OTHER half of the table, whose bounding sphere it blows up to
        # ~100 units -- any partition is legal -- and whose half is searched by the other thread of the point)
        ca, cb = (32 + j, 48 + j) if j % 2 == 0 else (j, 48 + j)
        place(case, ca * CL + 2, cb * CL + 2, 1000 + 2 * j + 1, 1000 + 2 * j)
    assert fused_would_differ == n_same_adv + n_cross_adv, fused_would_differ      # the test has teeth
    spheres = np.zeros((NC, 4), dtype=f32)
    for c in range(NC):
        m = verts[c * CL:(c + 1) * CL]
        m = m[m[:, 0] < 1e17, :3].astype(np.float64)
        if len(m) == 0:
            spheres[c] = (1e18, 1e18, 1e18, 0.0)
            continue
        cen = m.mean(0).astype(f32)
        rad = np.sqrt(((m - cen.astype(np.float64)) ** 2).sum(-1)).max()
        spheres[c] = (*cen, f32(rad * (1 + 1e-5) + 1e-6))
    verts[:, 3] = idx.view(f32)
    dev = torch.device("cuda:0")
    knn = torch.from_numpy(np.concatenate([_pair_layout(spheres), _pair_layout(verts)]).reshape(-1)).to(dev)
    n_idx = 1000 + 2 * 16
    aff = np.zeros((n_idx, 3, 4), dtype=f32)
    aff[:, 0, 3] = np.arange(n_idx, dtype=f32)
    aff = torch.from_numpy(aff.reshape(-1)).to(dev)
    rot = (ctypes.c_float * 9)(1, 0, 0, 0, 1, 0, 0, 0, 1)
    trans = (ctypes.c_float * 3)(0, 0, 0)
    pts = torch.from_numpy(np.stack(queries)).to(dev).contiguous()
    out = torch.empty_like(pts)
    call("hl_canonical_points", pts.data_ptr(), None, pts.shape[0], knn.data_ptr(), aff.data_ptr(), NC, CL,
         ctypes.cast(rot, ctypes.c_void_p), ctypes.cast(trans, ctypes.c_void_p), out.data_ptr(), None,
         torch.cuda.current_stream(dev).cuda_stream)
    torch.cuda.synchronize()
    got = out[:, 0].cpu().numpy().astype(np.int64)
    assert np.array_equal(got, np.asarray(expect)), [(i, int(g), int(e)) for i, (g, e) in enumerate(zip(got, expect)) if g != e]
    assert float(out[:, 1:].abs().max()) == 0.0
