"""GPU tests of the sampling-loop plumbing: in-kernel Philox noise, the one-graph-per-step loop, p_mean_variance,
denoised_fn, input validation (gaussian_diffusion.py:232-326,356-482; respace.py:88-122)."""
import numpy as np
import pytest
import torch

from common import CASES, load_golden, model_state_dict, rel_l2
from humanliff_b200 import _lib
from humanliff_b200._lib import call

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _tiny(precision="fp32", **over):
    fname, flags, seed, heads = CASES["tiny"]
    model, diffusion, sd = model_state_dict(dict(flags, precision=precision, **over), seed)
    model.load_state_dict(sd, strict=True)
    return model.to(DEV).eval(), diffusion, load_golden(fname), sd, heads


def test_philox_randn_moments_and_reproducibility():
    n = 1 << 22
    st = torch.cuda.current_stream().cuda_stream
    a, b, c = (torch.empty(n, device=DEV) for _ in range(3))
    call("hl_randn", a.data_ptr(), n, None, 1234, 0, 0, st)
    call("hl_randn", b.data_ptr(), n, None, 1234, 0, 0, st)
    call("hl_randn", c.data_ptr(), n, None, 1234, 1, 0, st)
    assert torch.equal(a, b)                                   # counter-based: same (seed, draw) -> same stream
    ad, cd = a.double(), c.double()
    assert abs(float(ad.mean())) < 3e-3 and abs(float(ad.var()) - 1.0) < 5e-3
    assert abs(float((ad ** 4).mean()) - 3.0) < 0.05           # kurtosis of a Gaussian
    assert abs(float((ad * cd).mean())) < 3e-3                 # successive draws are uncorrelated
    assert abs(float((ad[:-1] * ad[1:]).mean())) < 3e-3        # neighbouring elements are uncorrelated
    assert float(ad.abs().max()) < 7.0 and torch.isfinite(a).all()
    # device-resident state is read instead of the by-value pair
    state = torch.tensor([1234, 1], dtype=torch.int64, device=DEV)
    call("hl_randn", b.data_ptr(), n, state.data_ptr(), 0, 0, 0, st)
    assert torch.equal(b, c)
    # sharding invariance: the second half drawn on its own with its element offset equals the second half of the whole
    call("hl_randn", b.data_ptr(), n // 2, None, 1234, 0, n // 2, st)
    assert torch.equal(b[:n // 2], a[n // 2:])


def test_ddpm_step_rng_bit_exact_with_injected_noise_and_gaussian_without():
    from oracle.diffusion_oracle import DiffusionOracle
    _, diffusion, _, _, _ = _tiny()
    orc = DiffusionOracle(1000, "250")
    g = torch.Generator().manual_seed(3)
    B, n = 3, 27 * 32 * 32
    x, eps, z = (torch.randn(B, 27, 32, 32, generator=g) for _ in range(3))
    t = torch.tensor([0, 100, 249])
    ref_s, ref_x0 = orc.posterior(x, eps, t, z)
    s, x0 = diffusion._fused_step(x.to(DEV), eps.to(DEV), z.to(DEV), t.to(DEV), True)
    assert torch.equal(s.cpu(), ref_s) and torch.equal(x0.cpu(), ref_x0)
    # float64 / non-contiguous / CPU inputs are coerced instead of being reinterpreted
    s2, _ = diffusion._fused_step(x.to(DEV), eps.double().to(DEV), z.double(), t, True)
    assert torch.equal(s2, s)
    # in-kernel noise: (sample - mean) / sigma_t must be a standard normal; t = 0 adds no noise
    mean, _ = diffusion._fused_step(x.to(DEV), eps.to(DEV), torch.zeros_like(x).to(DEV), t.to(DEV), True)
    torch.manual_seed(7)
    s3, _ = diffusion._fused_step(x.to(DEV), eps.to(DEV), None, t.to(DEV), True)
    s4, _ = diffusion._fused_step(x.to(DEV), eps.to(DEV), None, t.to(DEV), True)
    assert torch.equal(s3[0], mean[0])
    sig = torch.exp(0.5 * torch.from_numpy(orc.logvar)[t[1:]].float()).view(-1, 1, 1, 1).to(DEV)
    zz = ((s3[1:] - mean[1:]) / sig).double()
    assert abs(float(zz.mean())) < 2e-2 and abs(float(zz.var()) - 1.0) < 3e-2
    assert not torch.equal(s3[1:], s4[1:])                     # the draw counter advanced
    torch.manual_seed(7)                                       # re-seeding restarts the sequence
    s5, _ = diffusion._fused_step(x.to(DEV), eps.to(DEV), None, t.to(DEV), True)
    assert torch.equal(s5, s3)
    # an out-of-range timestep poisons that sample's row instead of reading outside the tables
    bad, _ = diffusion._fused_step(x.to(DEV), eps.to(DEV), z.to(DEV), torch.tensor([5, 250, 7]).to(DEV), True)
    assert torch.isnan(bad[1]).all() and torch.isfinite(bad[0]).all() and torch.isfinite(bad[2]).all()


def test_p_mean_variance_entry_point_vs_oracle():
    """SpacedDiffusion.p_mean_variance(model, x, t, x_cond, ...) -- NB t BEFORE x_cond (gaussian_diffusion.py:232)."""
    from oracle.diffusion_oracle import DiffusionOracle
    model, diffusion, g, sd, heads = _tiny("fp32")
    orc = DiffusionOracle(1000, "250", num_heads=heads)
    x, xc, y = g["x"], g["x_cond"], g["y"]
    for t in (0, 100, 249):
        tt = torch.full((x.shape[0],), t, dtype=torch.int64)
        out = diffusion.p_mean_variance(model, x.to(DEV), tt.to(DEV), xc.to(DEV), clip_denoised=True,
                                        model_kwargs={"y": y.to(DEV)})
        assert set(out) >= {"mean", "variance", "log_variance", "pred_xstart"}
        ref = orc.p_sample(sd, x, xc, tt, y, torch.zeros_like(x))          # noise 0 -> sample == mean
        assert rel_l2(out["mean"], ref["sample"]) < 5e-5 and rel_l2(out["pred_xstart"], ref["pred_xstart"]) < 5e-5
        lv = torch.from_numpy(orc.logvar)[tt].float().view(-1, 1, 1, 1).expand_as(x)
        assert out["log_variance"].shape == x.shape and torch.equal(out["log_variance"].cpu(), lv)
        assert out["variance"].shape == x.shape and rel_l2(out["variance"], torch.exp(lv.double())) < 1e-6
        # reference goldens for the same call (sample with noise 0 is not stored; pred_xstart is)
        assert rel_l2(out["pred_xstart"], g[f"x0_{t}"]) < 5e-5 * max(1.0, float(diffusion.sqrt_recipm1_alphas_cumprod[t]))


def test_denoised_fn_is_applied_before_the_clamp():
    """gaussian_diffusion.py:293-298: x0 = denoised_fn(c0 x - c1 eps), THEN clamp(-1, 1), then the posterior mean."""
    from oracle.diffusion_oracle import DiffusionOracle
    model, diffusion, g, sd, heads = _tiny("fp32")
    orc = DiffusionOracle(1000, "250", num_heads=heads)
    x, xc, y = g["x"], g["x_cond"], g["y"]
    t = torch.full((x.shape[0],), 100, dtype=torch.int64)
    z = g["noise_100"]
    fn = lambda v: 0.25 * v + 0.1
    out = diffusion.p_sample(model, x.to(DEV), xc.to(DEV), t.to(DEV), clip_denoised=True, denoised_fn=fn,
                             model_kwargs={"y": y.to(DEV)}, noise=z.to(DEV))
    eps = orc.p_sample(sd, x, xc, t, y, z)["eps"]
    ex = lambda a: torch.from_numpy(a)[t].float().view(-1, 1, 1, 1)
    x0 = fn(ex(orc.sqrt_recip) * x - ex(orc.sqrt_recipm1) * eps).clamp(-1, 1)
    ref = ex(orc.c1) * x0 + ex(orc.c2) * x + torch.exp(0.5 * ex(orc.logvar)) * z
    assert rel_l2(out["pred_xstart"], x0) < 5e-5 and rel_l2(out["sample"], ref) < 5e-5


@pytest.mark.parametrize("precision", ["fp32", "fp16"])
def test_graph_loop_equals_per_step_path(precision):
    """p_sample_loop's one-CUDA-graph-per-step body (UNet + posterior + on-device t / RNG advance) against the
    same loop driven step by step through p_sample: identical inputs and injected noise -> same result."""
    model, diffusion, g, _, _ = _tiny(precision, timestep_respacing="8")
    x_T, xc, y = g["x"].to(DEV), g["x_cond"].to(DEV), g["y"].to(DEV)
    gen = torch.Generator().manual_seed(11)
    noises = {i: torch.randn(x_T.shape, generator=gen).to(DEV) for i in range(8)}
    n0 = _lib.launch_count
    out = diffusion.p_sample_loop(model, tuple(x_T.shape), x_cond=xc, noise=x_T.clone(), model_kwargs={"y": y},
                                  step_noise=lambda i: noises[i])
    assert _lib.launch_count - n0 >= 8 * 100            # every replay is accounted for in the launch counter
    plan = next(iter(model._plans.values()))
    assert plan.__dict__.get("_loops"), "the fused loop graph was not used"
    img = x_T.clone()
    for i in range(7, -1, -1):
        tt = torch.full((img.shape[0],), i, dtype=torch.int64, device=DEV)
        img = diffusion.p_sample(model, img, xc, tt, model_kwargs={"y": y}, noise=noises[i])["sample"]
    assert rel_l2(out, img) < 1e-6, rel_l2(out, img)
    # the progressive generator yields fresh tensors per step (reference semantics), pred_xstart included
    steps = list(diffusion.p_sample_loop_progressive(model, tuple(x_T.shape), x_cond=xc, noise=x_T.clone(),
                                                     model_kwargs={"y": y}, step_noise=lambda i: noises[i]))
    assert len(steps) == 8 and rel_l2(steps[-1]["sample"], img) < 1e-6
    assert steps[0]["sample"].data_ptr() != steps[1]["sample"].data_ptr() and steps[-1]["pred_xstart"] is not None
    # free-running with the in-kernel generator: finite, reproducible under torch.manual_seed, different otherwise
    torch.manual_seed(5)
    a = diffusion.p_sample_loop(model, tuple(x_T.shape), x_cond=xc, model_kwargs={"y": y})
    torch.manual_seed(5)
    b = diffusion.p_sample_loop(model, tuple(x_T.shape), x_cond=xc, model_kwargs={"y": y})
    c = diffusion.p_sample_loop(model, tuple(x_T.shape), x_cond=xc, model_kwargs={"y": y})
    assert torch.isfinite(a).all() and rel_l2(a, b) < 1e-6 and rel_l2(c, a) > 1e-2
    assert float(a.abs().max()) <= 1.0 + 1e-6            # the last step (t = 0) returns the clipped-x0 posterior mean


def test_rescale_timesteps_through_the_graph_loop():
    """rescale_timesteps=True + respacing: the on-device advance must hand the model 1000/T_orig * map[t]."""
    model, diffusion, g, _, _ = _tiny("fp32", timestep_respacing="5", rescale_timesteps=True)
    x_T, xc, y = g["x"].to(DEV), g["x_cond"].to(DEV), g["y"].to(DEV)
    gen = torch.Generator().manual_seed(2)
    noises = {i: torch.randn(x_T.shape, generator=gen).to(DEV) for i in range(5)}
    out = diffusion.p_sample_loop(model, tuple(x_T.shape), x_cond=xc, noise=x_T.clone(), model_kwargs={"y": y},
                                  step_noise=lambda i: noises[i])
    img = x_T.clone()
    for i in range(4, -1, -1):
        tt = torch.full((img.shape[0],), i, dtype=torch.int64, device=DEV)
        img = diffusion.p_sample(model, img, xc, tt, model_kwargs={"y": y}, noise=noises[i])["sample"]
    assert rel_l2(out, img) < 1e-6
