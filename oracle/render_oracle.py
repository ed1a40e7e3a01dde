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


@torch.no_grad()
def render_rays(sd, tri_planes, bounds, rays_o, rays_d, near, far, u, clamp_depth=True, n=128, n_importance=128):
    """tri_planes [3,9,R,R]; bounds [2,3]; rays [N,3]; near/far [N]; u [N,128] -> rgb, acc, depth.
    ``n_importance=0``: the `if n_importance > 0` block of Renderer.render (recon_NeRF/lib/renderer.py:258-270) is
    skipped and render_core composites the n coarse samples."""
    N = rays_o.shape[0]
    bmin, bmax = bounds[0], bounds[1]
    t = torch.linspace(0., 1., steps=n)
    z = near[:, None] * (1. - t) + far[:, None] * t
    if n_importance == 0:
        return _render_core(sd, tri_planes, bmin, bmax, rays_o, rays_d, near, far, z, clamp_depth)
    pts = rays_o[:, None] + rays_d[:, None] * z[..., None]
    sigma = mlp(sd, plane_features(tri_planes, pts.reshape(-1, 3), bmin, bmax)).reshape(N, n)
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
    return _render_core(sd, tri_planes, bmin, bmax, rays_o, rays_d, near, far, zf, clamp_depth)


def _render_core(sd, tri_planes, bmin, bmax, rays_o, rays_d, near, far, zf, clamp_depth):
    """render_core (recon_NeRF/lib/renderer.py:180-241) + the depth normalisation of render (:283-286)."""
    N = rays_o.shape[0]
    m = zf.shape[1]
    pts = rays_o[:, None] + rays_d[:, None] * zf[..., None]
    vd = (rays_d / rays_d.norm(dim=-1, keepdim=True))[:, None].expand(N, m, 3).reshape(-1, 3)
    rgb_raw, sigma = mlp(sd, plane_features(tri_planes, pts.reshape(-1, 3), bmin, bmax), vd)
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
