"""CPU restatement of the tri-plane render path (TEST INFRASTRUCTURE ONLY).

Follows recon_NeRF/run_nerf_batch.py:29-67 (render(): coarse z, points), recon_NeRF/lib/renderer.py
:142-164 (NeRF_network), :166-178 (up_sample), :180-241 (render_core), :244-295 (render), :504-549
(plane projection + nine-plane grid_sample), :551-581 (sample_pdf), lib/fields.py:69-85 (view
encoding); ``clamp_depth`` selects the human_diffusion/NeRF/renderer.py:273-274 variant.
SURVEY.md Appendix A.5-A.8 state the formulas.  `sd` is the renderer state dict, `u` the uniforms."""
import math

import torch
import torch.nn.functional as F


def plane_features(tri_planes, pts, bmin, bmax):
    """tri_planes [3, 9, R, R]; pts [M, 3] -> [M, 27] (feature j = plane*9 + sub*3 + ch)."""
    R = tri_planes.shape[-1]
    c = 2 * (pts - bmin) / (bmax - bmin) - 1
    uv = [c[:, [0, 1]], c[:, [0, 2]], c[:, [2, 1]]]
    feats = []
    for plane in range(3):
        for sub in range(3):
            g = uv[plane].clone()
            if sub == 1:
                g[:, 0] = g[:, 0] + 1 / R
            if sub == 2:
                g[:, 1] = g[:, 1] + 1 / R
            img = tri_planes[plane, sub * 3:sub * 3 + 3][None]            # [1, 3, R, R]
            s = F.grid_sample(img, g[None, None], mode="bilinear", padding_mode="zeros",
                              align_corners=False)                         # [1, 3, 1, M]
            feats.append(s[0, :, 0].t())
    return torch.cat(feats, -1)


def view_encoding(d):
    out = [d]
    for k in range(4):
        f = float(2 ** k)
        out.append(torch.sin(d * f))
        out.append(torch.sin(torch.addcmul(torch.tensor(math.pi * 0.5), d, torch.tensor(f))))
    return torch.cat(out, -1)


def mlp(sd, x, viewdir=None):
    lin = lambda n, v: F.linear(v, sd[n + ".weight"], sd[n + ".bias"])
    h0 = F.softplus(lin("pts_linears.0", x))
    h1 = F.softplus(lin("pts_linears.1", h0))
    h2 = F.softplus(lin("pts_linears.2", torch.cat([x, h1], -1)))
    sigma = lin("alpha_linear", h2)[:, 0]
    if viewdir is None:
        return sigma
    feat = lin("feature_linear", h2)
    hv = F.softplus(lin("views_linear", torch.cat([feat, view_encoding(viewdir)], -1)))
    return lin("rgb_linear", hv), sigma


# ------------------------------------------------------------------------------------------------------------------
# Canonical-space deformation (use_canonical_space=True): human_diffusion/NeRF/renderer.py:52-133 (deform_target2c_op,
# deform_target2c), :354-420 (get_transform_params_torch, get_rigid_transformation_torch, batch_rodrigues_torch),
# :435-462 (batch_rodrigues).  `smpl` = the asset as fp32 tensors (v_template [V,3], shapedirs [V,3,S], posedirs
# [V,3,207], J_regressor [24,V], weights [V,24], parents [24]); `frame` = tp_input of one sample (batch 1).
# pytorch3d's knn_points (K=1; third-party, pinned by the reference's environment to pytorch3d 0.7, absent here) is
# restated as: fp32 squared distances, smallest wins, lowest index on ties.

def smpl_tensors(asset):
    f = lambda a: torch.as_tensor(a, dtype=torch.float64).float()
    out = {k: f(asset[k]) for k in ("v_template", "shapedirs", "posedirs", "J_regressor", "weights")}
    out["parents"] = [int(v) for v in asset["kintree_table"][0]]
    return out


def axis_angle_matrices(rv):
    """[N,3] -> [N,3,3]: I + sin(a) K + (1 - cos(a)) K^2 with a = |rv + 1e-8| (both batch_rodrigues variants)."""
    a = torch.norm(rv + 1e-8, dim=1, keepdim=True)
    k = rv / a
    z = torch.zeros_like(k[:, 0])
    K = torch.stack([z, -k[:, 2], k[:, 1], k[:, 2], z, -k[:, 0], -k[:, 1], k[:, 0], z], 1).view(-1, 3, 3)
    return torch.eye(3)[None] + torch.sin(a)[:, None] * K + (1 - torch.cos(a))[:, None] * torch.matmul(K, K)


def joint_transforms(smpl, poses, shapes):
    """get_transform_params_torch: poses [72], shapes [S] -> A [24,4,4] (posed joint frames with the rest joints
    removed, i.e. the matrices linear blend skinning applies to rest-pose points)."""
    v_shaped = smpl["v_template"] + (smpl["shapedirs"][..., :shapes.numel()] * shapes.view(1, 1, -1)).sum(-1)
    rot = axis_angle_matrices(poses.reshape(-1, 3))
    J = smpl["J_regressor"] @ v_shaped
    par = smpl["parents"]
    rel = J.clone()
    rel[1:] = J[1:] - J[par[1:]]
    local = torch.zeros(24, 4, 4)
    local[:, :3, :3] = rot
    local[:, :3, 3] = rel
    local[:, 3, 3] = 1
    chain = [local[0]]
    for j in range(1, 24):
        chain.append(chain[par[j]] @ local[j])
    A = torch.stack(chain)
    Jh = torch.cat([J, torch.zeros(24, 1)], -1)
    A[..., 3] = A[..., 3] - (A * Jh[:, None, :]).sum(-1)
    return A


def nearest_vertex(q, v):
    out = []
    for s in range(0, q.shape[0], 8192):
        out.append(((q[s:s + 8192, None, :] - v[None]) ** 2).sum(-1).argmin(-1))
    return torch.cat(out)


def deform_to_canonical(smpl, frame, pts, viewdir=None):
    """deform_target2c + deform_target2c_op for batch 1: pts [M,3] world -> canonical big-pose space; viewdir [M,3] ->
    canonical directions (the reference subtracts Th from the *direction* too, renderer.py:125 -- reproduced)."""
    prm, tprm = frame["params"], frame["t_params"]
    R, Th = prm["R"][0], prm["Th"][0]
    poses, shapes = prm["poses"].reshape(-1), prm["shapes"].reshape(-1)
    q = (pts - Th) @ R
    qv = (viewdir - Th) @ R if viewdir is not None else None
    verts = (frame["vertices"][0] - Th) @ R
    vid = nearest_vertex(q.float(), verts.float())
    bw = smpl["weights"][vid]                                          # [M,24]
    A = (bw @ joint_transforms(smpl, poses, shapes).reshape(24, 16)).view(-1, 4, 4)
    Rinv = torch.inverse(A[:, :3, :3].float())
    c = (Rinv @ (q - A[:, :3, 3])[..., None])[..., 0]
    if qv is not None:
        qv = (Rinv @ qv[..., None])[..., 0]
    eye = torch.eye(3)
    pdirs = smpl["posedirs"].reshape(-1, 207).t()                     # [207, V*3]
    feat = (axis_angle_matrices(poses.view(-1, 3))[1:] - eye).reshape(1, -1)
    c = c - (feat @ pdirs).view(-1, 3)[vid]
    c = c - (smpl["shapedirs"][..., :shapes.numel()] @ shapes.view(-1, 1))[..., 0][vid]
    tposes = tprm["poses"].reshape(-1)
    tfeat = (axis_angle_matrices(tposes.view(-1, 3))[1:] - eye).reshape(1, -1)
    c = c + (tfeat @ pdirs).view(-1, 3)[vid]
    Ab = (bw @ joint_transforms(smpl, tposes, torch.zeros_like(shapes)).reshape(24, 16)).view(-1, 4, 4)
    c = (Ab[:, :3, :3] @ c[..., None])[..., 0] + Ab[:, :3, 3]
    if qv is not None:
        qv = (Ab[:, :3, :3] @ qv[..., None])[..., 0]
    return c, qv


@torch.no_grad()
def render_rays(sd, tri_planes, bounds, rays_o, rays_d, near, far, u, clamp_depth=True, n=128, n_importance=128,
                canon=None):
    """tri_planes [3,9,R,R]; bounds [2,3]; rays [N,3]; near/far [N]; u [N,128] -> rgb, acc, depth.
    ``n_importance=0``: the `if n_importance > 0` block of Renderer.render (recon_NeRF/lib/renderer.py:258-270) is
    skipped and render_core composites the n coarse samples.
    ``canon=(smpl, frame)``: use_canonical_space=True -- every sample point (and, in render_core, its view direction) goes
    through deform_to_canonical and ``bounds`` must be the frame's t_world_bounds."""
    N = rays_o.shape[0]
    bmin, bmax = bounds[0], bounds[1]
    t = torch.linspace(0., 1., steps=n)
    z = near[:, None] * (1. - t) + far[:, None] * t
    if n_importance == 0:
        return _render_core(sd, tri_planes, bmin, bmax, rays_o, rays_d, near, far, z, clamp_depth, canon)
    pts = (rays_o[:, None] + rays_d[:, None] * z[..., None]).reshape(-1, 3)
    if canon is not None:
        pts, _ = deform_to_canonical(canon[0], canon[1], pts)
    sigma = mlp(sd, plane_features(tri_planes, pts, bmin, bmax)).reshape(N, n)
    # up_sample
    dists = torch.cat([z[:, 1:] - z[:, :-1], torch.full((N, 1), 1e10)], -1) * rays_d.norm(dim=-1, keepdim=True)
    alpha = 1. - torch.exp(-F.softplus(sigma) * dists)
    w = alpha * torch.cumprod(torch.cat([torch.ones(N, 1), 1. - alpha + 1e-10], -1), -1)[:, :-1]
    bins = .5 * (z[:, 1:] + z[:, :-1])
    ww = w[:, 1:-1] + 1e-5
    pdf = ww / ww.sum(-1, keepdim=True)
    cdf = torch.cat([torch.zeros(N, 1), torch.cumsum(pdf, -1)], -1)
    inds = torch.searchsorted(cdf, u.contiguous(), right=True)
    below = (inds - 1).clamp(min=0)
    above = inds.clamp(max=cdf.shape[-1] - 1)
    cb, ca = torch.gather(cdf, 1, below), torch.gather(cdf, 1, above)
    bb, ba = torch.gather(bins, 1, below), torch.gather(bins, 1, above)
    den = ca - cb
    den = torch.where(den < 1e-5, torch.ones_like(den), den)
    z_new = bb + (u - cb) / den * (ba - bb)
    zf, _ = torch.sort(torch.cat([z, z_new], -1), -1)
    return _render_core(sd, tri_planes, bmin, bmax, rays_o, rays_d, near, far, zf, clamp_depth, canon)


def _render_core(sd, tri_planes, bmin, bmax, rays_o, rays_d, near, far, zf, clamp_depth, canon=None):
    """render_core (recon_NeRF/lib/renderer.py:180-241) + the depth normalisation of render (:283-286)."""
    N = rays_o.shape[0]
    m = zf.shape[1]
    pts = (rays_o[:, None] + rays_d[:, None] * zf[..., None]).reshape(-1, 3)
    vd = (rays_d / rays_d.norm(dim=-1, keepdim=True))[:, None].expand(N, m, 3).reshape(-1, 3)
    if canon is not None:
        pts, vd = deform_to_canonical(canon[0], canon[1], pts, vd)
    rgb_raw, sigma = mlp(sd, plane_features(tri_planes, pts, bmin, bmax), vd)
    sigma = sigma.reshape(N, m)
    d2 = torch.cat([zf[:, 1:] - zf[:, :-1], torch.full((N, 1), 1e10)], -1)      # not scaled by |d|
    alpha = 1. - torch.exp(-F.softplus(sigma) * d2)
    col = torch.sigmoid(rgb_raw).reshape(N, m, 3)
    w = alpha * torch.cumprod(torch.cat([torch.ones(N, 1), 1. - alpha + 1e-7], -1), -1)[:, :-1]
    acc = w.sum(-1)
    rgb = (col * w[..., None]).sum(1)
    depth = ((w * zf).sum(-1) - near) / (far - near + 1e-5)
    if clamp_depth:
        depth = depth.clamp(0, 1)
    return rgb, acc, depth
