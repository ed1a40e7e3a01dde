"""Functional CPU restatement of ``UNetModel.forward`` (TEST INFRASTRUCTURE ONLY).

Follows human_diffusion/improved_diffusion/unet.py:550-615 with
``cond_type in {"controlnet", ""}``, ``use_scale_shift_norm=True``,
``use_3d_aware=False``.  The block structure is re-derived from the state-dict
keys (no architecture table shared with the product):

* ``<blk>.0.in_layers.2.weight``  -> ResBlock            (unet.py:198-219)
* ``<blk>.N.qkv.weight``          -> AttentionBlock      (unet.py:244-274)
* ``<blk>.0.op.weight``           -> Downsample conv s2  (unet.py:104-106)
* ``<blk>.N.conv.weight``         -> Upsample + conv     (unet.py:70-80)

``operand_round`` optionally emulates the product's numerics: conv / conv1d
operands rounded to TF32 (10-bit mantissa, round-to-nearest, ties away) with
fp32 accumulation.  It exists only to predict the parity margin in tests.
"""
import math

import torch
import torch.nn.functional as F


def round_tf32(x: torch.Tensor) -> torch.Tensor:
    """cvt.rna.tf32.f32: round-to-nearest (ties away from zero) to 10 mantissa bits."""
    i = x.contiguous().view(torch.int32)
    i = (i + 0x1000) & ~0x1FFF
    return i.view(torch.float32)


def trunc_tf32(x: torch.Tensor) -> torch.Tensor:
    i = x.contiguous().view(torch.int32)
    return (i & ~0x1FFF).view(torch.float32)


class _Numerics:
    """``mode``: one operand-rounding mode for every contraction, or a callable ``layer_prefix -> mode`` (a
    per-layer policy: which layers the product runs in which operand format; used for the error budget and to
    emulate mixed-precision plans)."""

    def __init__(self, mode):
        self.policy = mode if callable(mode) else None
        self.mode = None if callable(mode) else mode

    def at(self, name):
        """The numerics of layer ``name``."""
        if self.policy is None:
            return self
        return _Numerics(self.policy(name))

    def op(self, x, raw=False):
        if self.mode is None:
            return x
        if self.mode == "tf32":
            return round_tf32(x)
        if self.mode == "tf32_trunc_raw":
            return trunc_tf32(x) if raw else round_tf32(x)
        if self.mode == "fp16_scaled":
            # raw-stream operand carried as fp16(x * 2^-4) (weights as fp16(w * 2^4)): range 1.0e6 instead of 65504
            return (x * 2.0 ** -4).to(torch.float16).to(torch.float32) * 2.0 ** 4
        if self.mode == "fp16x2":
            # fp16 hi + lo pair for BOTH operands of the layer (three MMAs: hi.hi + hi.lo + lo.hi): ~22 bits
            hi = x.to(torch.float16).to(torch.float32)
            return hi + (x - hi).to(torch.float16).to(torch.float32)
        if self.mode == "fp16":
            # cvt.rn.f16.f32 operands (11-bit significand like TF32, 5-bit exponent), fp32 accumulate
            return x.to(torch.float16).to(torch.float32)
        if self.mode.startswith("fp16_rawsplit"):
            # raw residual-stream operands carried as an fp16 hi + lo pair (K doubled): ~22 bits.
            # "fp16_rawsplit" splits every raw conv; "fp16_rawsplit:skip,proj" only the named kinds.
            kinds = self.mode.split(":")[1].split(",") if ":" in self.mode else None
            if raw and (kinds is None or raw in kinds):
                hi = x.to(torch.float16).to(torch.float32)
                return hi + (x - hi).to(torch.float16).to(torch.float32)
            return x.to(torch.float16).to(torch.float32)
        if self.mode == "bf16":
            return x.to(torch.bfloat16).to(torch.float32)
        raise ValueError(self.mode)


def product_fp16_plan(prefix):
    """Per-layer policy emulating the product's ``precision="fp16"`` plan: every contraction rounds
    its operands to fp16 except the convs that read the raw residual stream (1x1 skip, ControlNet projection,
    the stems), which carry hi + lo fp16 pairs for both operands in three tensor-core passes, and the output conv
    (weight pair only, stacked along Cout: one pass); the Upsample / Downsample convs read a 2^-4-scaled fp16 operand.  Pass as ``operand_round=product_fp16_plan``."""
    import re
    if prefix.endswith(".h"):
        return "fp16"                            # ResBlock intermediate (in_layers output) stored as fp16
    if (prefix.endswith("skip_connection") or prefix.startswith("input_blocks_proj_cond.")
            or re.fullmatch(r"input_blocks(_cond)?\.0\.0", prefix)):
        return "fp16_split3"
    if prefix == "out.2":
        return "fp16_split2w"                    # weight pair only (activations plain fp16), one tensor-core pass
    if prefix.endswith(".conv") or prefix.endswith(".op"):
        return "fp16_scaled"
    return "fp16"


def timestep_embedding(t, dim, max_period=10000):
    """nn.py:103-121 -- [cos | sin] blocks, fp32."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32) / half)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def _gn(x, sd, prefix):
    return F.group_norm(x.float(), 32, sd[prefix + ".weight"], sd[prefix + ".bias"], eps=1e-5)


def _silu(x):
    return x * torch.sigmoid(x)


def _conv(x, sd, prefix, nm, stride=1, raw=False):
    w = sd[prefix + ".weight"]
    pad = w.shape[-1] // 2
    nm = nm.at(prefix)
    if nm.mode == "fp16_split2a":
        f16 = lambda t: t.to(torch.float16).to(torch.float32)
        xh = f16(x)
        return F.conv2d(xh + f16(x - xh), f16(w), sd[prefix + ".bias"], stride=stride, padding=pad)
    if nm.mode == "fp16_split2w":
        f16 = lambda t: t.to(torch.float16).to(torch.float32)
        wh = f16(w)
        return F.conv2d(f16(x), wh, sd[prefix + ".bias"], stride=stride, padding=pad) + \
            F.conv2d(f16(x), f16(w - wh), None, stride=stride, padding=pad)
    if nm.mode == "fp16_split3":
        # the product's high-precision conv: activations * 2^-4 and weights * 2^4 each carried as an fp16 hi + lo
        # pair, three tensor-core passes hi.hi + lo.hi + hi.lo (the lo.lo term, 2^-22 relative, is dropped)
        f16 = lambda t: t.to(torch.float16).to(torch.float32)
        xs, wsc = x * 2.0 ** -4, w * 2.0 ** 4
        xh, wh = f16(xs), f16(wsc)
        xl, wl = f16(xs - xh), f16(wsc - wh)
        y = F.conv2d(xh, wh, None, stride=stride, padding=pad) + F.conv2d(xl, wh, None, stride=stride, padding=pad) \
            + F.conv2d(xh, wl, None, stride=stride, padding=pad)
        return y + sd[prefix + ".bias"][None, :, None, None]
    return F.conv2d(nm.op(x, raw), nm.op(w), sd[prefix + ".bias"], stride=stride, padding=pad)


def _resblock(x, emb, sd, p, nm):
    """unet.py:198-219 (scale-shift norm)."""
    h = _conv(_silu(_gn(x, sd, p + ".in_layers.0")), sd, p + ".in_layers.2", nm)
    if nm.at(p + ".h").mode == "fp16":
        # the product stores the tensor between the two convs (read only by out_layers' GroupNorm) as fp16
        h = h.to(torch.float16).to(torch.float32)
    e = F.linear(_silu(emb), sd[p + ".emb_layers.1.weight"], sd[p + ".emb_layers.1.bias"])
    cout = h.shape[1]
    scale, shift = e[:, :cout, None, None], e[:, cout:, None, None]
    h = _gn(h, sd, p + ".out_layers.0") * (1 + scale) + shift
    h = _conv(_silu(h), sd, p + ".out_layers.3", nm)
    if (p + ".skip_connection.weight") in sd:
        x = _conv(x, sd, p + ".skip_connection", nm, raw="skip")
    return x + h


def _attention(x, sd, p, heads, nm):
    """unet.py:244-274 -- head-major qkv channel order."""
    b, c, hh, ww = x.shape
    xf = x.reshape(b, c, -1)
    nm = nm.at(p)
    n = F.group_norm(xf.float(), 32, sd[p + ".norm.weight"], sd[p + ".norm.bias"], eps=1e-5)
    qkv = F.conv1d(nm.op(n), nm.op(sd[p + ".qkv.weight"]), sd[p + ".qkv.bias"])
    qkv = qkv.reshape(b * heads, -1, qkv.shape[2])
    ch = qkv.shape[1] // 3
    q, k, v = torch.split(qkv, ch, dim=1)
    s = 1.0 / math.sqrt(math.sqrt(ch))
    w = torch.einsum("bct,bcs->bts", nm.op(q * s), nm.op(k * s))
    w = torch.softmax(w.float(), dim=-1)
    a = torch.einsum("bts,bcs->bct", nm.op(w), nm.op(v)).reshape(b, -1, xf.shape[-1])
    a = F.conv1d(nm.op(a), nm.op(sd[p + ".proj_out.weight"]), sd[p + ".proj_out.bias"])
    return (xf + a).reshape(b, c, hh, ww)


def _run_block(h, emb, sd, p, heads, nm):
    """One TimestepEmbedSequential (unet.py:41-49): walk sub-indices 0,1,2."""
    j = 0
    while True:
        q = f"{p}.{j}"
        if (q + ".in_layers.2.weight") in sd:
            h = _resblock(h, emb, sd, q, nm)
        elif (q + ".qkv.weight") in sd:
            h = _attention(h, sd, q, heads, nm)
        elif (q + ".op.weight") in sd:
            h = _conv(h, sd, q + ".op", nm, stride=2, raw="down")
        elif (q + ".conv.weight") in sd:
            h = F.interpolate(h, scale_factor=2, mode="nearest")
            h = _conv(h, sd, q + ".conv", nm, raw="up")
        elif (q + ".weight") in sd and sd[q + ".weight"].dim() == 4:
            h = _conv(h, sd, q, nm, raw="stem")  # stem conv
        else:
            break
        j += 1
    assert j > 0, p
    return h


def _count(sd, prefix):
    idx = set()
    for k in sd:
        if k.startswith(prefix + "."):
            idx.add(int(k[len(prefix) + 1:].split(".")[0]))
    return (max(idx) + 1) if idx else 0


@torch.no_grad()
def unet_forward(sd, x, timesteps, x_cond=None, y=None, num_heads=4, operand_round=None):
    """ε = UNetModel.forward(x, timesteps, x_cond, y)   (unet.py:550-615)."""
    nm = _Numerics(operand_round)
    model_ch = sd["time_embed.0.weight"].shape[1]
    emb = timestep_embedding(timesteps, model_ch)
    emb = F.linear(emb, sd["time_embed.0.weight"], sd["time_embed.0.bias"])
    emb = F.linear(_silu(emb), sd["time_embed.2.weight"], sd["time_embed.2.bias"])
    if "label_emb.weight" in sd:
        assert y is not None and y.shape == (x.shape[0],)
        emb = emb + sd["label_emb.weight"][y]

    n_in = _count(sd, "input_blocks")
    hs = []
    h = x.float()
    for i in range(n_in):
        h = _run_block(h, emb, sd, f"input_blocks.{i}", num_heads, nm)
        hs.append(h)
    h = _run_block(h, emb, sd, "middle_block", num_heads, nm)  # Res, Attn, Res (unet.py:416-438)

    controlnet = "input_blocks_cond.0.0.weight" in sd
    hs_cond = []
    if controlnet:
        hc = x.float() + x_cond.float()
        for i in range(n_in):
            hc = _run_block(hc, emb, sd, f"input_blocks_cond.{i}", num_heads, nm)
            # NB unet.py:599-601 -- the projection REPLACES h_cond and feeds the next block
            hc = _conv(hc, sd, f"input_blocks_proj_cond.{i}", nm, raw="proj")
            hs_cond.append(hc)

    for i in range(_count(sd, "output_blocks")):
        skip = hs.pop() + hs_cond.pop() if controlnet else hs.pop()
        h = _run_block(torch.cat([h, skip], dim=1), emb, sd, f"output_blocks.{i}", num_heads, nm)
    h = _silu(_gn(h, sd, "out.0"))
    return _conv(h, sd, "out.2", nm)

