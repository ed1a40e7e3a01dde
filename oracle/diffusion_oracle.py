"""CPU restatement of the DDPM sampling maths (TEST INFRASTRUCTURE ONLY).

Follows human_diffusion/improved_diffusion/gaussian_diffusion.py:18-42 (beta schedule), :118-169
(tables), :232-326 (p_mean_variance, FIXED_LARGE / FIXED_SMALL, epsilon prediction), :356-388
(p_sample), :434-482 (loop), :484-529 (ddim_sample) and respace.py:7-60,72-86,117-122 (timestep respacing).  Tables are
float64 numpy; per-step arithmetic is fp32 torch exactly in the reference's operation order
(SURVEY.md Appendix A.1).  Noise is always injected by the caller."""
import numpy as np
import torch

from .unet_oracle import unet_forward


def linear_betas(T):
    scale = 1000 / T
    return np.linspace(scale * 1e-4, scale * 2e-2, T, dtype=np.float64)


def kept_timesteps(T, respacing):
    """respace.py:7-60 for the plain 'N' / '' forms and comma lists."""
    if not respacing:
        return list(range(T))
    counts = [int(v) for v in str(respacing).split(",")]
    per, extra = T // len(counts), T % len(counts)
    out, start = [], 0
    for i, c in enumerate(counts):
        size = per + (1 if i < extra else 0)
        stride = 1 if c <= 1 else (size - 1) / (c - 1)
        cur = 0.0
        for _ in range(c):
            out.append(start + round(cur))
            cur += stride
        start += size
    return sorted(set(out))


class DiffusionOracle:
    def __init__(self, T=1000, respacing="", sigma_small=False, num_heads=4):
        self.num_heads = num_heads
        abar_full = np.cumprod(1.0 - linear_betas(T))
        self.timestep_map = kept_timesteps(T, respacing)
        kept = abar_full[self.timestep_map]
        betas = 1.0 - kept / np.append(1.0, kept[:-1])          # respace.py:79-83
        abar = np.cumprod(1.0 - betas)                           # re-derived by GaussianDiffusion.__init__
        abar_prev = np.append(1.0, abar[:-1])
        self.num_timesteps = len(betas)
        self.betas, self.abar, self.abar_prev = betas, abar, abar_prev
        self.sqrt_recip = np.sqrt(1.0 / abar)
        self.sqrt_recipm1 = np.sqrt(1.0 / abar - 1)
        self.post_var = betas * (1.0 - abar_prev) / (1.0 - abar)
        self.c1 = betas * np.sqrt(abar_prev) / (1.0 - abar)
        self.c2 = (1.0 - abar_prev) * np.sqrt(1.0 - betas) / (1.0 - abar)
        if sigma_small:
            self.logvar = np.log(np.append(self.post_var[1], self.post_var[1:]))
        else:
            self.logvar = np.log(np.append(self.post_var[1], betas[1:]))

    @staticmethod
    def _ex(arr, t, x):
        return torch.from_numpy(arr)[t].float().view(-1, *([1] * (x.dim() - 1)))

    def posterior(self, x, eps, t, noise, clip=True):
        """(sample, pred_xstart) given the model output -- gaussian_diffusion.py:293-314,383-387."""
        x0 = self._ex(self.sqrt_recip, t, x) * x - self._ex(self.sqrt_recipm1, t, x) * eps
        if clip:
            x0 = x0.clamp(-1, 1)
        mean = self._ex(self.c1, t, x) * x0 + self._ex(self.c2, t, x) * x
        mask = (t != 0).float().view(-1, *([1] * (x.dim() - 1)))
        sample = mean + mask * torch.exp(0.5 * self._ex(self.logvar, t, x)) * noise
        return sample, x0

    def ddim_posterior(self, x, eps, t, noise, eta=0.0, clip=True):
        """(sample, pred_xstart) of ddim_sample given the model output -- gaussian_diffusion.py:500-529."""
        c0, c1 = self._ex(self.sqrt_recip, t, x), self._ex(self.sqrt_recipm1, t, x)
        x0 = c0 * x - c1 * eps
        if clip:
            x0 = x0.clamp(-1, 1)
        eps2 = (c0 * x - x0) / c1                                  # _predict_eps_from_xstart (:335-339)
        ab, abp = self._ex(self.abar, t, x), self._ex(self.abar_prev, t, x)
        sigma = eta * torch.sqrt((1 - abp) / (1 - ab)) * torch.sqrt(1 - ab / abp)
        mean = x0 * torch.sqrt(abp) + torch.sqrt(1 - abp - sigma ** 2) * eps2
        mask = (t != 0).float().view(-1, *([1] * (x.dim() - 1)))
        return mean + mask * sigma * noise, x0

    @torch.no_grad()
    def p_sample(self, sd, x, x_cond, t, y, noise, clip=True, operand_round=None):
        ts = torch.tensor(self.timestep_map, dtype=torch.int64)[t]
        eps = unet_forward(sd, x, ts, x_cond, y, num_heads=self.num_heads, operand_round=operand_round)
        sample, x0 = self.posterior(x, eps, t, noise, clip)
        return {"sample": sample, "pred_xstart": x0, "eps": eps}

    @torch.no_grad()
    def p_sample_loop(self, sd, x_T, x_cond, y, step_noise, steps=None, operand_round=None):
        """Free-running loop from x_T over the LAST `steps` timesteps' worth of indices T-1..T-steps."""
        img = x_T
        idx = list(range(self.num_timesteps))[::-1]
        if steps is not None:
            idx = idx[:steps]
        for i in idx:
            t = torch.full((x_T.shape[0],), i, dtype=torch.int64)
            img = self.p_sample(sd, img, x_cond, t, y, step_noise(i), operand_round=operand_round)["sample"]
        return img
