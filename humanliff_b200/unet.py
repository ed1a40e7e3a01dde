"""B200-native ``UNetModel`` -- the drop-in for human_diffusion/improved_diffusion/unet.py:300-615.

Same constructor flags, same ``forward(x, timesteps, x_cond=None, y=None)`` (NCHW fp32 in / out) and the
same 953-tensor ``state_dict()`` key set / shapes, so reference checkpoints load unchanged.  Nothing
here computes with PyTorch: torch owns device memory and the stream, every FLOP is a call into
``libhumanliff_b200.so`` (``include/humanliff_b200.h``).

Internals
  * the residual stream and every conv result are NHWC fp32; conv OPERANDS (normalised / activated
    tensors and the packed weights ``[tap][Cout_pad][Cin_pad]``) are fp16 (``precision="fp16"``,
    tcgen05 kind::f16 -- the same 11-bit significand as TF32 at twice the MMA rate and half the
    bytes), TF32-rounded fp32 (``"tf32"``) or exact fp32 on the CUDA cores (``"fp32"``);
  * per (batch, H, W) the forward pass is compiled once into a flat list of C-ABI launches over a
    fixed workspace, run eagerly once and then replayed as ONE CUDA graph;
  * GroupNorm statistics are per-channel fp64 sums produced by the epilogue of the conv that
    writes the tensor; one fused pass applies GroupNorm32 + FiLM + SiLU and writes the next conv's
    operand (plus, for ResBlocks with a 1x1 skip, the raw operand copy of the input);
  * the decoder's ``cat([h, hs.pop() + hs_cond.pop()], 1)`` (unet.py:606) is never materialised by a
    copy: producers write straight into channel slices of the concat buffer, and the ControlNet
    projection conv adds ``hs`` as its residual while writing the skip half;
  * all 62 ResBlock ``emb_layers`` Linear layers are stacked into one matrix and evaluated by a
    single streaming GEMV per step (unet.py:151-157,200).
"""
import math
import os

import torch
import torch.nn as nn

from . import _lib
from ._lib import call

_DT = {"fp16": (_lib.DT_F16, torch.float16, 64), "tf32": (_lib.DT_F32, torch.float32, 32),
       "fp32": (_lib.DT_F32, torch.float32, 4)}


def _ptr(t):
    return t.data_ptr() if t is not None else None


class _Node(nn.Module):
    """Plain container; lets parameters carry the reference's dotted names."""


def _set_param(root, dotted, shape):
    parts = dotted.split(".")
    node = root
    for p in parts[:-1]:
        if not hasattr(node, p):
            node.add_module(p, _Node())
        node = getattr(node, p)
    node.register_parameter(parts[-1], nn.Parameter(torch.zeros(*shape), requires_grad=False))


HP_SCALE = 16.0      # raw residual-stream operands are stored as value * 2^-4; their weights carry 2^4 (HL_OP_SCALED)


def pack_conv(weight, bias, cin_pad=None, precision="fp16", device=None, mode=None):
    """OIHW (or Conv1d [O, I, 1]) weight -> packed ``[kh*kw][Cout_pad][Cin_pad]`` operand (fp16, TF32-rounded
    fp32 or exact fp32) + padded fp32 bias.  One-time host-side re-layout at load time.

    ``mode`` (fp16 only; the high-precision operand passes of DESIGN.md 3, HL_CONV_SPLIT3 / SPLIT2P in the header):
    ``"scaled"`` weights * 2^4; ``"split"`` two slabs ``{W_hi, W_lo}`` of the scaled weights, ``W_lo = fp16(W - W_hi)``;
    ``"split_a"`` plain weights for an operand that is an (unscaled) hi | lo pair; ``"split_w"`` the weights as an fp16
    hi + lo pair stacked along Cout -- rows ``[0, Cout)`` = W_hi, rows ``[Cout_pad, Cout_pad + Cout)`` = W_lo of a
    ``2 * Cout_pad``-row operand whose two result halves the caller sums (plain fp16 activations); ``"split_packed"`` (stem: 2 * Cin <= Cin_pad) slabs ``{[W_hi | W_hi], [W_lo | 0]}`` for
    an operand row ``[hi(Cin) 0.. | lo(Cin) 0..]`` with the lo half at channel Cin_pad / 2."""
    lib = _lib.load()
    device = device if device is not None else weight.device
    w = weight.detach().to(device=device, dtype=torch.float32)
    if w.dim() == 3:
        w = w[..., None]
    cout, cin, kh, kw = w.shape
    cin_pad = cin_pad or cin
    cout_pad = lib.hl_conv_cout_pad(cout)
    pk = torch.zeros(kh * kw, cout_pad, cin_pad, device=device, dtype=torch.float32)
    pk[:, :cout, :cin] = w.permute(2, 3, 0, 1).reshape(kh * kw, cout, cin)
    if precision == "fp16" and mode == "split_w":
        hi = pk.to(torch.float16)
        lo = (pk - hi.float()).to(torch.float16)
        b = torch.zeros(2 * cout_pad, device=device, dtype=torch.float32)
        if bias is not None:
            b[:cout] = bias.detach().to(device=device, dtype=torch.float32)
        return torch.cat([hi, lo], 1).contiguous(), b                 # [taps][2 * Cout_pad][Cin_pad]
    if precision == "fp16":
        if mode in ("scaled", "split"):
            pk = pk * HP_SCALE
        hi = pk.to(torch.float16)          # round-to-nearest-even, as cvt.rn.f16.f32
        if mode in ("split", "split_packed"):
            lo = (pk - hi.float()).to(torch.float16)
            if mode == "split_packed":
                half = cin_pad // 2
                assert cin <= half, "split_packed needs 2 * Cin <= Cin_pad"
                s0, s1 = hi.clone(), torch.zeros_like(hi)
                s0[:, :, half:half + cin] = hi[:, :, :cin]
                s1[:, :, :cin] = lo[:, :, :cin]
                pk = torch.cat([s0, s1], 0)
            else:
                pk = torch.cat([hi, lo], 0)
        else:
            assert mode in (None, "scaled", "split_a"), mode
            pk = hi
    elif precision == "tf32":
        flat = pk.view(-1, 4)
        call("hl_cast_operand", _ptr(flat), 4, _ptr(flat), _lib.DT_F32, 4, 4, flat.shape[0], 1,
             torch.cuda.current_stream(device).cuda_stream)
    b = torch.zeros(cout_pad, device=device, dtype=torch.float32)
    if bias is not None:
        b[:cout] = bias.detach().to(device=device, dtype=torch.float32)
    return pk.contiguous(), b


def _hp_mode(name, cin, cin_pad):
    """Which convs of the fp16 plan run the high-precision operand passes: the ones that read the RAW
    residual stream -- 1x1 skip, ControlNet projection, Downsample, stem -- and the output conv carry hi + lo fp16
    pairs (error budget: tools/error_budget.py); the Upsample conv reads a 2^-4-scaled operand (range only)."""
    if name.endswith("skip_connection") or name.startswith("input_blocks_proj_cond."):
        return "split"
    if name == "out.2":
        # weight pair only, stacked along Cout (N = 64 costs the tensor core what N = 32 does: at this width the MMA
        # is bound by fetching its 128 activation rows): ONE pass, the same emulated error as the activation pair in
        # two passes (4.75e-4 either way, tools/error_budget.py), -0.13 ms/step
        return os.environ.get("HL_OUT_CONV", "split_w")
    if name in ("input_blocks.0.0", "input_blocks_cond.0.0"):
        return "split_packed" if 2 * cin <= cin_pad else None
    if name.endswith(".conv") or name.endswith(".op"):
        return "scaled"           # Upsample / Downsample convs: range only (3 % of the error variance, tools/error_budget.py)
    return None


class _Conv:
    """One packed convolution / conv1d / linear-over-pixels."""

    def __init__(self, name, cin, cout, ksize, stride=1, cin_pad=None):
        self.name, self.cin, self.cout, self.ksize, self.stride = name, cin, cout, ksize, stride
        self.cout_launch = cout  # output channels of the launch (split_w: 2 * Cout_pad, the hi and lo halves)
        self.cin_pad = cin_pad or cin
        self.w = None
        self.b = None
        self.hp = None           # high-precision operand mode of the fp16 plan (_hp_mode), set by UNetModel


class UNetModel(nn.Module):
    """See module docstring.  Supported envelope: ``cond_type`` in {"controlnet", ""},
    ``use_scale_shift_norm=True``, ``use_3d_aware=False``, ``dims=2``, ``conv_resample=True``,
    ``dropout`` ignored at inference; anything else raises ``NotImplementedError``."""

    def __init__(self, in_channels, model_channels, out_channels, num_res_blocks, attention_resolutions,
                 dropout=0, channel_mult=(1, 2, 4, 8), conv_resample=True, dims=2, num_classes=None,
                 use_checkpoint=False, num_heads=1, num_heads_upsample=-1, use_scale_shift_norm=False,
                 cond_type="", use_3d_aware=False, transformer_depth=1, context_dim=None,
                 precision="fp16"):
        super().__init__()
        if cond_type not in ("controlnet", ""):
            raise NotImplementedError(f"cond_type={cond_type!r}: only 'controlnet' and '' are built")
        if not use_scale_shift_norm:
            raise NotImplementedError("use_scale_shift_norm=False is outside the production envelope")
        if use_3d_aware or dims != 2 or not conv_resample:
            raise NotImplementedError("use_3d_aware / dims != 2 / conv_resample=False are not built")
        if precision not in _DT:
            raise ValueError("precision must be 'fp16', 'tf32' or 'fp32'")
        if num_heads_upsample == -1:
            num_heads_upsample = num_heads
        if num_heads_upsample != num_heads:
            raise NotImplementedError("num_heads_upsample != num_heads")
        self.in_channels = in_channels
        self.model_channels = model_channels
        self.out_channels = out_channels
        self.num_res_blocks = num_res_blocks
        self.attention_resolutions = tuple(attention_resolutions)
        self.dropout = dropout
        self.channel_mult = tuple(channel_mult)
        self.conv_resample = conv_resample
        self.num_classes = num_classes
        self.use_checkpoint = use_checkpoint
        self.num_heads = num_heads
        self.num_heads_upsample = num_heads_upsample
        self.cond_type = cond_type
        self.use_3d_aware = use_3d_aware
        self.precision = precision
        self.emb_dim = model_channels * 4
        chunk = _DT[precision][2]
        self.cin_pad = (in_channels + chunk - 1) // chunk * chunk   # stem operand: one whole K chunk

        self._convs = {}
        self._film = []          # (prefix, cout, offset) in stacking order
        self._film_rows = 0
        # hi + lo operand passes for the raw-stream convs and the output conv (fp16 plan only; HL_HIPREC=0 = plain fp16)
        self.hi_precision = precision == "fp16" and os.environ.get("HL_HIPREC", "1") != "0"
        self.h_f16 = precision == "fp16" and os.environ.get("HL_H_F16", "1") != "0"     # ResBlock intermediate as fp16
        self.dual_proj = os.environ.get("HL_DUAL_PROJ", "1") != "0"    # ControlNet projection: one launch, two results
        self.side_skip = os.environ.get("HL_SIDE_SKIP", "1") != "0"    # decoder <= 32^2: 1x1 skip conv on the side stream
        self.fused_gn_skip = os.environ.get("HL_FUSED_GN_SKIP", "1") != "0"   # GroupNorm-1 + 1x1 skip conv in one kernel (hl_gn_skip)
        self.split_reduce_in_kernel = os.environ.get("HL_SPLIT_RED", "0") == "1"   # experiment: split-K second pass inside the conv kernel (measured slower)
        self._build_plan()
        if self.hi_precision:
            for c in self._convs.values():
                c.hp = _hp_mode(c.name, c.cin, c.cin_pad)
                if c.hp == "split_w":
                    c.cout_launch = 2 * (32 * ((c.cout + 31) // 32))
        self._packed_key = None
        self._plans = {}
        self.use_cuda_graph = True
        self.concurrent_encoders = True     # ControlNet encoder on a side stream (see _StepPlan._build)
        self.batch_split = int(os.environ.get("HL_BATCH_SPLIT", "1"))     # independent sub-batch chains in one graph (_SplitPlan)
        self.split_k = os.environ.get("HL_SPLITK", "1") != "0"   # split-K for the 8^2 / 16^2 3x3 layers (hl_conv_set_workspace)
        self.programmatic_launch = os.environ.get("HL_PDL", "0") == "1"    # PDL: each kernel's prologue overlaps its predecessor's tail (hl_set_pdl)

    # ------------------------------------------------------------------ architecture / parameters
    def _conv(self, name, cin, cout, k, stride=1, cin_pad=None):
        shape = (cout, cin, k, k) if not name.endswith(("qkv", "proj_out")) else (cout, cin, 1)
        _set_param(self, name + ".weight", shape)
        _set_param(self, name + ".bias", (cout,))
        self._convs[name] = _Conv(name, cin, cout, k, stride, cin_pad)
        return name

    def _norm(self, name, c):
        if c % 32:
            raise NotImplementedError(f"GroupNorm32 over {c} channels")
        _set_param(self, name + ".weight", (c,))
        _set_param(self, name + ".bias", (c,))
        return name

    def _res(self, prefix, cin, cout):
        blk = {"kind": "res", "p": prefix, "cin": cin, "cout": cout,
               "n1": self._norm(prefix + ".in_layers.0", cin),
               "c1": self._conv(prefix + ".in_layers.2", cin, cout, 3)}
        _set_param(self, prefix + ".emb_layers.1.weight", (2 * cout, self.emb_dim))
        _set_param(self, prefix + ".emb_layers.1.bias", (2 * cout,))
        blk["film_off"] = self._film_rows
        self._film.append((prefix, cout, self._film_rows))
        self._film_rows += 2 * cout
        blk["n2"] = self._norm(prefix + ".out_layers.0", cout)
        blk["c2"] = self._conv(prefix + ".out_layers.3", cout, cout, 3)
        blk["skip"] = self._conv(prefix + ".skip_connection", cin, cout, 1) if cin != cout else None
        return blk

    def _attn(self, prefix, c):
        return {"kind": "attn", "p": prefix, "c": c, "n": self._norm(prefix + ".norm", c),
                "qkv": self._conv(prefix + ".qkv", c, 3 * c, 1),
                "proj": self._conv(prefix + ".proj_out", c, c, 1)}

    def _encoder(self, root, with_proj):
        """input_blocks / input_blocks_cond (unet.py:375-415, 477-518).  Returns (blocks, channels)."""
        mc = self.model_channels
        blocks = [[{"kind": "stem", "c": self._conv(f"{root}.0.0", self.in_channels, mc, 3,
                                                    cin_pad=self.cin_pad), "cout": mc}]]
        chans = [mc]
        ch, ds = mc, 1
        for level, mult in enumerate(self.channel_mult):
            for _ in range(self.num_res_blocks):
                i = len(blocks)
                layers = [self._res(f"{root}.{i}.0", ch, mult * mc)]
                ch = mult * mc
                if ds in self.attention_resolutions:
                    layers.append(self._attn(f"{root}.{i}.1", ch))
                blocks.append(layers)
                chans.append(ch)
            if level != len(self.channel_mult) - 1:
                i = len(blocks)
                blocks.append([{"kind": "down", "c": self._conv(f"{root}.{i}.0.op", ch, ch, 3, stride=2),
                                "ch": ch}])
                chans.append(ch)
                ds *= 2
        if with_proj:
            for i, c in enumerate(chans):
                self._conv(f"input_blocks_proj_cond.{i}", c, c, 1)
        return blocks, chans, ch, ds

    def _build_plan(self):
        mc, ed = self.model_channels, self.emb_dim
        _set_param(self, "time_embed.0.weight", (ed, mc))
        _set_param(self, "time_embed.0.bias", (ed,))
        _set_param(self, "time_embed.2.weight", (ed, ed))
        _set_param(self, "time_embed.2.bias", (ed,))
        if self.num_classes is not None:
            _set_param(self, "label_emb.weight", (self.num_classes, ed))
        self._enc, chans, ch, ds = self._encoder("input_blocks", False)
        self._enc_chans = list(chans)
        self._mid = [self._res("middle_block.0", ch, ch), self._attn("middle_block.1", ch),
                     self._res("middle_block.2", ch, ch)]
        self._dec = []
        stack = list(chans)
        for level, mult in list(enumerate(self.channel_mult))[::-1]:
            for i in range(self.num_res_blocks + 1):
                j = len(self._dec)
                skip_c = stack.pop()
                layers = [self._res(f"output_blocks.{j}.0", ch + skip_c, mc * mult)]
                layers[0]["cat"] = (ch, skip_c)
                ch = mc * mult
                if ds in self.attention_resolutions:
                    layers.append(self._attn(f"output_blocks.{j}.1", ch))
                if level and i == self.num_res_blocks:
                    layers.append({"kind": "up", "ch": ch,
                                   "c": self._conv(f"output_blocks.{j}.{len(layers)}.conv", ch, ch, 3)})
                    ds //= 2
                self._dec.append(layers)
        self._norm("out.0", ch)
        self._conv("out.2", mc, self.out_channels, 3)
        self._enc_cond = None
        if self.cond_type == "controlnet":
            self._enc_cond, _, _, _ = self._encoder("input_blocks_cond", True)

    def convert_to_fp16(self):
        raise NotImplementedError("fp16 torso (training option) is outside the inference hot path")

    def convert_to_fp32(self):
        return None

    @property
    def inner_dtype(self):
        return torch.float32

    # ------------------------------------------------------------------ weight packing
    def _p(self, name):
        node = self
        for part in name.split("."):
            node = getattr(node, part)
        return node

    def _pack(self, device):
        key = (str(device), tuple(p._version for p in self.parameters()),
               tuple(p.data_ptr() for p in self.parameters()))
        if key == self._packed_key:
            return
        self._plans = {}          # plans hold pointers into the packed weights
        for c in self._convs.values():
            c.w, c.b = pack_conv(self._p(c.name + ".weight"), self._p(c.name + ".bias"), c.cin_pad,
                                 self.precision, device, mode=c.hp)
        ws, bs = [], []
        for prefix, cout, off in self._film:
            ws.append(self._p(prefix + ".emb_layers.1.weight").detach().to(device, torch.float32))
            bs.append(self._p(prefix + ".emb_layers.1.bias").detach().to(device, torch.float32))
        self._film_w = torch.cat(ws, 0).contiguous()
        self._film_b = torch.cat(bs, 0).contiguous()
        self._small = {n: self._p(n).detach().to(device, torch.float32).contiguous()
                       for n in ("time_embed.0.weight", "time_embed.0.bias", "time_embed.2.weight",
                                 "time_embed.2.bias")}
        if self.num_classes is not None:
            self._small["label_emb.weight"] = self._p("label_emb.weight").detach().to(
                device, torch.float32).contiguous()
        half = self.model_channels // 2
        self._freqs = torch.exp(-math.log(10000) * torch.arange(start=0, end=half, dtype=torch.float32)
                                / half).to(device)     # host fp32, as nn.py:114-116
        self._norm_p = {}
        for name, p in self.named_parameters():
            if p.dim() == 1 and (name.endswith("in_layers.0.weight") or name.endswith("in_layers.0.bias")
                                 or name.endswith("out_layers.0.weight") or name.endswith("out_layers.0.bias")
                                 or ".norm." in name or name.startswith("out.0.")):
                self._norm_p[name] = p.detach().to(device, torch.float32).contiguous()
        if torch.device(device).type == "cuda":
            torch.cuda.current_stream(device).synchronize()
        self._packed_key = key

    def _uses_tc(self, cname, B, H, W, ldx=None, ldy=None, flags=0):
        """True if conv ``cname`` on a B x H x W input would run on the tcgen05 kernel."""
        c = self._convs[cname]
        if self.precision == "fp32":
            return False
        if self.precision == "tf32":
            flags |= _lib.CONV_TF32
        return bool(_lib.load().hl_conv2d_uses_tensor_cores(
            _DT[self.precision][0], B, H, W, c.cin_pad, c.cout, c.ksize, c.stride, ldx or c.cin_pad,
            ldy or _lib.load().hl_conv_cout_pad(c.cout), flags))

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def forward(self, x, timesteps, x_cond=None, y=None):
        """eps-prediction.  x: [B, C, H, W] fp32 CUDA, timesteps: [B] (int or float), x_cond like x
        (required for cond_type='controlnet'), y: [B] int64 labels iff class-conditional."""
        if not x.is_cuda:
            raise RuntimeError("humanliff_b200.UNetModel runs on CUDA (sm_100a) only -- no CPU fallback")
        if self.num_classes is not None:
            assert y is not None and y.shape == (x.shape[0],)
        if self.cond_type == "controlnet" and x_cond is None:
            raise ValueError("cond_type='controlnet' needs x_cond")
        B, Cx, H, W = x.shape
        assert Cx == self.in_channels
        nlev = len(self.channel_mult) - 1
        if H % (1 << nlev) or W % (1 << nlev):
            raise ValueError(f"H, W must be divisible by {1 << nlev}")
        device = x.device
        with torch.cuda.device(device):
            return self.plan_for(device, B, H, W).run(x, timesteps, x_cond, y)

    def plan_for(self, device, B, H, W):
        """The compiled launch plan (workspace + CUDA graph) of a forward pass at (B, H, W) on ``device``; call
        under ``torch.cuda.device(device)``.  Keyed on the tuning switches too, so toggling one after the first
        forward takes effect."""
        self._pack(device)
        key = (str(device), B, H, W, self.use_cuda_graph, self.concurrent_encoders, self.batch_split, self.split_k,
               self.programmatic_launch, self.h_f16, self.dual_proj, self.side_skip, self.split_reduce_in_kernel, self.fused_gn_skip)
        plan = self._plans.get(key)
        if plan is None:
            parts = self.batch_split
            if parts > 1 and B % parts == 0 and B // parts >= 2:
                plan = _SplitPlan(self, device, B, H, W, parts)
            else:
                plan = _StepPlan(self, device, B, H, W)
            self._plans[key] = plan
        return plan


class _Ref:
    """A residual-stream tensor inside the workspace: fp32 NHWC ``ptr`` with pixel pitch ``ld``, ``C``
    channels at H x W, and the per-channel statistics row (``st`` pointer, ``st_ld`` doubles-pairs per
    sample) its producer fills."""
    __slots__ = ("ptr", "ld", "C", "H", "W", "st", "st_ld", "f16")

    def __init__(self, ptr, ld, C, H, W, st, st_ld, f16=False):
        self.ptr, self.ld, self.C, self.H, self.W, self.st, self.st_ld = ptr, ld, C, H, W, st, st_ld
        # an fp16 OPERAND buffer written directly by the producing conv: 1 = HL_CONV_OUT_F16, 2 = HL_CONV_OUT_F16_SPLIT
        # (scaled hi | lo pair, ld >= 2 C)
        self.f16 = int(f16)


class _StepPlan:
    """One UNet forward at a fixed (B, H, W): workspace + flat launch list (+ its CUDA graph)."""
    SPLITK_BYTES = 16 << 20      # >= S * B*H*W*Cout*4 of every layer that splits (8^2: 4.7 MB, 16^2: 9.4 MB at B = 4)

    def __init__(self, model, device, B, H, W):
        self.m, self.device, self.B, self.H, self.W = model, device, B, H, W
        self.dt, self.tdt, self.chunk = _DT[model.precision]
        self.rnd = 1 if model.precision == "tf32" else 0
        self.calls = []
        self.bufs = {}
        self.branch = 0              # 0 = main stream, 1 = side stream (ControlNet encoder)
        self.n_events = 0
        self.graph = None
        self.side = None
        self.splitk_ws = None
        self.kernels_per_run = 0
        self.SPLITK_BYTES = (4 << 20) * max(4, B)
        self.events = None
        self.runs = 0
        self._stats_off = 0
        self._stats_reqs = []
        self._build()

    # ---------------------------------------------------------------- workspace
    def buf(self, name, numel, dtype=torch.float32):
        t = self.bufs.get(name)
        if t is None:
            t = torch.empty(int(numel), device=self.device, dtype=dtype)
            self.bufs[name] = t
        assert t.numel() >= numel and t.dtype == dtype, name
        return t

    def opbuf(self, name, numel):
        return self.buf(name, numel, self.tdt)

    def scratch(self, name, numel, op=False):
        """Scratch reused in stream order; each concurrent branch (stream) owns its own copy."""
        full = name if self.branch == 0 else f"{name}.b{self.branch}"
        return self.opbuf(full, numel) if op else self.buf(full, numel)

    def stats_row(self, C):
        """Reserve a [B, C, 2] fp64 statistics block; returns its offset in doubles (resolved to a
        pointer once the arena is allocated)."""
        off = self._stats_off
        self._stats_off += self.B * C * 2
        return off

    def new_ref(self, name, C, H, W):
        """A fresh fp32 tensor [B, H, W, C] with its own statistics row."""
        t = self.buf(name, self.B * H * W * C)
        return _Ref(_ptr(t), C, C, H, W, self.stats_row(C), C)

    def emit(self, name, *args):
        self.calls.append((name, args, self.branch))

    def emit_sync(self, kind, arg=None):
        """Stream-dependency markers interpreted by _launch_all: ("fork",) side stream starts after everything
        issued so far on the main stream; ("join",) main waits for the side stream; ("signal", k) main records
        event k; ("wait", k) side stream waits for event k."""
        self.calls.append(("#" + kind, (arg,), self.branch))

    # ---------------------------------------------------------------- op emitters
    def conv(self, cname, x_ptr, ldx, res, dst, H, W, flags=0, want_stats=True):
        """dst: _Ref (its ptr/ld/st are used); res: _Ref or None.  H, W: input size."""
        m = self.m
        c = m._convs[cname]
        if m.precision == "fp32":
            flags |= _lib.CONV_FORCE_SIMT
        elif m.precision == "tf32":
            flags |= _lib.CONV_TF32
        st = ("stats", dst.st) if (want_stats and dst.st is not None) else None
        if dst.f16:
            assert st is None or dst.f16 == 1          # an fp16 result may carry statistics (of the rounded values)
            flags |= _lib.CONV_OUT_F16_SPLIT if dst.f16 == 2 else _lib.CONV_OUT_F16
        if c.hp == "split":
            flags |= _lib.CONV_SPLIT3          # x_ptr = [hi(Cin) | lo(Cin)] rows, weights {W_hi, W_lo}
            assert ldx >= 2 * c.cin_pad
        elif c.hp == "split_a":
            flags |= _lib.CONV_SPLIT2A         # x_ptr = [hi | lo] rows, one weight slab
            assert ldx >= 2 * c.cin_pad
        elif c.hp == "split_packed":
            flags |= _lib.CONV_SPLIT2P
        self.emit("hl_conv2d", x_ptr, self.dt, ldx, _ptr(c.w), _ptr(c.b), res.ptr if res else None,
                  res.ld if res else 0, dst.ptr, dst.ld, st, dst.st_ld if st else 0, self.B, H, W, c.cin_pad,
                  c.cout_launch, c.ksize, c.stride, flags)

    def conv_dual(self, cname, x_ptr, ldx, res, dst, dst2, H, W):
        """One launch, two results (hl_conv2d_dual): dst = conv + res, dst2 = conv; both fp32 with statistics."""
        m = self.m
        c = m._convs[cname]
        flags = 0
        if m.precision == "fp32":
            flags |= _lib.CONV_FORCE_SIMT
        elif m.precision == "tf32":
            flags |= _lib.CONV_TF32
        assert not dst.f16 and not dst2.f16 and dst.st is not None and dst2.st is not None
        if c.hp == "split":
            flags |= _lib.CONV_SPLIT3
            assert ldx >= 2 * c.cin_pad
        else:
            assert c.hp is None, c.hp
        self.emit("hl_conv2d_dual", x_ptr, self.dt, ldx, _ptr(c.w), _ptr(c.b), res.ptr, res.ld, dst.ptr, dst.ld,
                  ("stats", dst.st), dst.st_ld, dst2.ptr, dst2.ld, ("stats", dst2.st), dst2.st_ld, self.B, H, W, c.cin_pad,
                  c.cout, c.ksize, c.stride, flags)

    def gn(self, nname, x, out_ptr, ldo, silu, film=None, raw_ptr=None, ldraw=0, out_mode=0, raw_mode=0):
        """out_mode / raw_mode: HL_OP_* flags of the normalised output / the raw operand copy (a split output has
        its lo half at channel C of a row of pitch >= 2 C)."""
        m = self.m
        mode = self.rnd | out_mode | (raw_mode << _lib.OP_RAW_SHIFT)
        if out_mode & _lib.OP_SPLIT:
            mode |= x.C << 8
        if x.f16:
            assert x.f16 == 1 and raw_ptr is None
            mode |= _lib.OP_X_F16
        self.emit("hl_gn_apply", x.ptr, x.ld, ("stats", x.st), x.st_ld, _ptr(m._norm_p[nname + ".weight"]),
                  _ptr(m._norm_p[nname + ".bias"]), film, m._film_rows if film is not None else 0, out_ptr,
                  self.dt, ldo, raw_ptr, ldraw, self.B, x.H * x.W, x.C, 32, 1e-5, 1 if silu else 0, mode)

    def raw_operand(self, cname, C):
        """(pitch in elements, HL_OP_* mode) of the operand buffer a raw-stream conv reads."""
        hp = self.m._convs[cname].hp
        if hp == "split":
            return 2 * C, _lib.OP_SPLIT | _lib.OP_SCALED
        if hp == "scaled":
            return C, _lib.OP_SCALED
        return C, 0

    def cast(self, x, dst_ptr, ldd, mode=0):
        if mode & _lib.OP_SPLIT:
            mode |= x.C << 8
        self.emit("hl_cast_operand", x.ptr, x.ld, dst_ptr, self.dt, ldd, x.C, self.B * x.H * x.W, self.rnd | mode)

    # ---------------------------------------------------------------- blocks
    def res_block(self, blk, x, dst):
        """unet.py:198-219.  x, dst: _Ref."""
        m, B = self.m, self.B
        cin, cout, H, W = blk["cin"], blk["cout"], x.H, x.W
        assert x.C == cin and dst.C == cout
        # the shared scratch buffers are sized by _scratch_sizes' bound: check the real footprint of every block
        assert B * H * W * max(cin, cout) <= self.max_act, ("scratch bound", blk["p"], cin, cout, H, W)
        act = _ptr(self.scratch("act", self.max_act, op=True))
        # the tensor between the two convs is read by out_layers' GroupNorm only: in the fp16 plan conv1's epilogue
        # rounds it once to fp16 (statistics of the rounded values), halving its write, its read and the epilogue's
        # shared-memory traffic (emulated cost 4.3e-4 -> 4.7e-4 on the production architecture)
        hf16 = m.h_f16 and self.dt == _lib.DT_F16
        h = _Ref(_ptr(self.scratch("h", self.max_act, op=hf16)), cout, cout, H, W, self.stats_row(cout), cout, f16=hf16)
        raw, ldraw, raw_mode = None, 0, 0
        if blk["skip"] is not None:
            raw = _ptr(self.scratch("raw", self.max_raw, op=True))
            ldraw, raw_mode = self.raw_operand(blk["skip"], cin)
            assert B * H * W * ldraw <= self.max_raw, ("scratch bound (raw)", blk["p"])
        # GroupNorm-1 + the 1x1 skip conv in ONE kernel (hl_gn_skip): x is read once, the hi | lo operand pair of the skip
        # conv is built in shared memory instead of travelling through HBM.  Used where the launch fills the GPU (>= 96
        # tiles of 128 pixels); below that the two-launch form (skip conv on the side stream) is faster.
        fused_skip = (blk["skip"] is not None and m.fused_gn_skip and self.dt == _lib.DT_F16
                      and m._convs[blk["skip"]].hp == "split" and B * H * W >= 96 * 128 and x.ld % 4 == 0
                      and bool(_lib.load().hl_gn_skip_supported(B, H * W, cin, cout)))
        if fused_skip:
            sc = m._convs[blk["skip"]]
            s = _Ref(_ptr(self.scratch("skipbuf", self.max_act)), cout, cout, H, W, None, 0)
            self.emit("hl_gn_skip", x.ptr, x.ld, ("stats", x.st), x.st_ld, _ptr(m._norm_p[blk["n1"] + ".weight"]),
                      _ptr(m._norm_p[blk["n1"] + ".bias"]), act, cin, _ptr(sc.w), _ptr(sc.b), s.ptr, s.ld, B, H * W, cin, cout,
                      32, 1e-5)
            self.conv(blk["c1"], act, cin, None, h, H, W)
            self.gn(blk["n2"], h, act, cout, True, film=("film", blk["film_off"]))
            self.conv(blk["c2"], act, cout, s, dst, H, W)
            return
        self.gn(blk["n1"], x, act, cin, True, raw_ptr=raw, ldraw=ldraw, raw_mode=raw_mode)
        # The 1x1 skip conv needs only GroupNorm-1's raw copy: in the decoder's low-resolution stretch (one stream, every
        # kernel a fraction of a wave, the chain bound by launch latency) it runs on the side stream next to conv1 / GroupNorm-2
        overlap_skip = (blk["skip"] is not None and self.side_skip and self.branch == 0 and self.decoding
                        and H * W <= 32 * 32)
        s = None
        if blk["skip"] is not None:
            s = _Ref(_ptr(self.scratch("skipbuf", self.max_act)), cout, cout, H, W, None, 0)
        if overlap_skip:
            self.emit_sync("fork")
            self.branch = 1
            self.conv(blk["skip"], raw, ldraw, None, s, H, W, want_stats=False)
            self.branch = 0
        self.conv(blk["c1"], act, cin, None, h, H, W)
        film = ("film", blk["film_off"])
        self.gn(blk["n2"], h, act, cout, True, film=film)
        if blk["skip"] is not None:
            if overlap_skip:
                self.emit_sync("join")
            else:
                self.conv(blk["skip"], raw, ldraw, None, s, H, W, want_stats=False)
            self.conv(blk["c2"], act, cout, s, dst, H, W)
        else:
            self.conv(blk["c2"], act, cout, x, dst, H, W)

    def attn_block(self, blk, x, dst):
        """unet.py:244-274."""
        m, B = self.m, self.B
        C, H, W = blk["c"], x.H, x.W
        assert B * H * W * C <= self.max_act and B * H * W * 3 * C <= self.max_qkv, ("scratch bound", blk["p"])
        act = _ptr(self.scratch("act", self.max_act, op=True))
        # fp16 mode: qkv is only ever an attention operand -> the conv writes it as fp16
        f16 = self.dt == _lib.DT_F16
        qkv = _Ref(_ptr(self.scratch("qkv", self.max_qkv, op=f16)), 3 * C, 3 * C, H, W, None, 0, f16=f16)
        att = _ptr(self.scratch("att", self.max_act, op=True))
        self.gn(blk["n"], x, act, C, False)
        self.conv(blk["qkv"], act, C, None, qkv, H, W, want_stats=False)
        self.emit("hl_attention", qkv.ptr, self.dt if f16 else _lib.DT_F32, 3 * C, att, self.dt, C, B, H * W, C,
                  m.num_heads, self.rnd)
        self.conv(blk["proj"], att, C, x, dst, H, W)

    def layers(self, layers, x, dst_name, dst=None):
        """One TimestepEmbedSequential (unet.py:41-49).  The last layer writes ``dst`` (a _Ref made by
        the caller, e.g. a slice of a concat buffer) or a fresh tensor named ``dst_name``."""
        n = len(layers)
        for li, blk in enumerate(layers):
            last = li == n - 1
            kind = blk["kind"]
            if kind == "res":
                C, H, W = blk["cout"], x.H, x.W
            elif kind == "attn":
                C, H, W = blk["c"], x.H, x.W
            elif kind == "down":
                C, H, W = blk["ch"], x.H // 2, x.W // 2
            else:
                C, H, W = blk["ch"], x.H * 2, x.W * 2
            if last and dst is not None:
                out = dst
                assert (out.C, out.H, out.W) == (C, H, W), (dst_name, out.C, C, out.H, H)
            else:
                out = self.new_ref(dst_name if last else f"{dst_name}.t{li}", C, H, W)
            if kind == "res":
                self.res_block(blk, x, out)
            elif kind == "attn":
                self.attn_block(blk, x, out)
            elif kind == "down":
                op = _ptr(self.scratch("raw", self.max_raw, op=True))
                ldop, mode = self.raw_operand(blk["c"], x.C)
                assert self.B * x.H * x.W * ldop <= self.max_raw, ("scratch bound (raw)", dst_name)
                self.cast(x, op, ldop, mode)
                self.conv(blk["c"], op, ldop, None, out, x.H, x.W)
            elif kind == "up":
                op = _ptr(self.scratch("upbuf", self.max_act, op=True))
                _, mode = self.raw_operand(blk["c"], x.C)
                assert self.B * H * W * x.C <= self.max_act, ("scratch bound", dst_name)
                self.emit("hl_upsample2x", x.ptr, x.ld, op, self.dt, x.C, self.B, x.H, x.W, x.C, self.rnd | mode)
                self.conv(blk["c"], op, x.C, None, out, H, W)
            else:
                raise AssertionError(kind)
            x = out
        return x

    def encoder(self, enc, xin_ptr, tag, cats, controlnet_branch):
        """input_blocks / input_blocks_cond (unet.py:589-602).  ``cats[i]``: the _Ref of the skip half
        of the concat buffer that block i's skip tensor is written to.
        main encoder, no ControlNet : block i writes its output straight into cats[i]
        main encoder, ControlNet    : block i keeps hs[i] (returned)
        ControlNet encoder          : hc_i = proj_i(raw_i) feeds block i+1; the same projection run
                                      with residual hs[i] writes hs[i] + hc_i into cats[i]"""
        m, B, H, W = self.m, self.B, self.H, self.W
        mc = m.model_channels
        outs = []
        x = None
        f16 = self.dt == _lib.DT_F16
        for i, layers in enumerate(enc):
            direct = (not controlnet_branch) and not self.keep_hs     # unconditional: write into cat
            dst = cats[i] if direct else None
            if controlnet_branch and f16:
                # a ControlNet block's output is consumed only by its projection conv: write the fp16
                # operand straight from the producing conv's epilogue (no fp32 tensor, no cast pass)
                C_, H_, W_ = self.geo[i]
                ldp, pmode = self.raw_operand(f"input_blocks_proj_cond.{i}", C_)
                dst = _Ref(_ptr(self.scratch("hcop", self.max_raw, op=True)), ldp, C_, H_, W_, None, 0,
                           f16=2 if pmode & _lib.OP_SPLIT else 1)
            if i == 0:
                out = dst if dst is not None else self.new_ref(f"{tag}{i}", mc, H, W)
                self.conv(layers[0]["c"], xin_ptr, m.cin_pad, None, out, H, W)
                x = out
            else:
                if i == 1 and controlnet_branch and self.concurrent:
                    self.emit_sync("wait", self.film_event)                     # FiLM scale / shift of the ResBlocks
                x = self.layers(layers, x, f"{tag}{i}", dst=dst)
            if controlnet_branch:
                cname = f"input_blocks_proj_cond.{i}"
                if x.f16:
                    op, ldop = x.ptr, x.ld
                else:
                    op = _ptr(self.scratch("raw", self.max_raw, op=True))
                    ldop, pmode = self.raw_operand(cname, x.C)
                    self.cast(x, op, ldop, pmode)
                hc = self.new_ref(f"{tag}p{i}", x.C, x.H, x.W)
                if m.dual_proj:
                    # h_cond (unet.py:600) and hs + hs_cond (unet.py:606) from one pass over the operands
                    if self.concurrent:
                        self.emit_sync("wait", i)                                # hs[i] of the main encoder
                    self.conv_dual(cname, op, ldop, self.hs[i], cats[i], hc, x.H, x.W)
                else:
                    self.conv(cname, op, ldop, None, hc, x.H, x.W)               # h_cond (unet.py:600)
                    if self.concurrent:
                        self.emit_sync("wait", i)                                # hs[i] of the main encoder
                    self.conv(cname, op, ldop, self.hs[i], cats[i], x.H, x.W)    # hs + hs_cond (unet.py:606)
                x = hc
            elif self.keep_hs and self.concurrent:
                self.emit_sync("signal", i)                                      # hs[i] complete
            outs.append(x)
        return outs

    # ---------------------------------------------------------------- whole step
    def _scratch_sizes(self):
        m, B, H, W = self.m, self.B, self.H, self.W
        mc, cm = m.model_channels, m.channel_mult
        act, qkv = B * H * W * max(m.cin_pad, mc), 4
        for l, mult in enumerate(cm):
            pix = B * (H >> l) * (W >> l)
            below = cm[min(l + 1, len(cm) - 1)]
            act = max(act, pix * mc * (mult + max(mult, below)))
            if (1 << l) in m.attention_resolutions or l == len(cm) - 1:   # middle_block.1 always attends
                qkv = max(qkv, pix * 3 * mc * max(mult, below))
        return act, qkv

    def _build(self):
        m, B, H, W = self.m, self.B, self.H, self.W
        mc, ed = m.model_channels, m.emb_dim
        dev = self.device
        self.max_act, self.max_qkv = self._scratch_sizes()
        self.max_raw = self.max_act * (2 if m.hi_precision else 1)       # split raw operands are hi | lo pairs
        # static inputs / outputs of the graph
        self.x_in = torch.zeros(B, m.in_channels, H, W, device=dev)
        self.xc_in = torch.zeros(B, m.in_channels, H, W, device=dev) if m.cond_type == "controlnet" else None
        self.t_in = torch.zeros(B, device=dev)
        self.y_in = torch.zeros(B, device=dev, dtype=torch.int64)
        self.out = torch.empty(B, m.out_channels, H, W, device=dev)

        # --- concat buffers of the decoder: cat_j = [ h (hC) | skip (sC) ] at the skip's resolution ---
        nblk = len(m._enc)
        geo = []                       # per encoder block i: (C, H, W) of its output
        h_, w_ = H, W
        for i, layers in enumerate(m._enc):
            if layers[0]["kind"] == "down":
                h_, w_ = h_ // 2, w_ // 2
            geo.append((m._enc_chans[i], h_, w_))
        self.geo = geo
        self.cat_full, self.cat_skip = [None] * nblk, [None] * nblk
        hC = m._mid[-1]["cout"]
        self.cat_h = [None] * nblk     # indexed by decoder block j
        for j, layers in enumerate(m._dec):
            i = nblk - 1 - j
            sC, sH, sW = geo[i]
            assert layers[0]["cat"] == (hC, sC), (j, layers[0]["cat"], hC, sC)
            ld = hC + sC
            t = self.buf(f"cat{j}", B * sH * sW * ld)
            st = self.stats_row(ld)
            self.cat_h[j] = _Ref(_ptr(t), ld, hC, sH, sW, st, ld)
            self.cat_skip[i] = _Ref(_ptr(t) + 4 * hC, ld, sC, sH, sW, st + 2 * hC, ld)
            self.cat_full[j] = _Ref(_ptr(t), ld, ld, sH, sW, st, ld)
            hC = layers[0]["cout"]

        self.emit("hl_zero", ("stats", 0), ("stats_bytes",))
        xin = self.opbuf("xin", B * H * W * m.cin_pad)
        stem_mode = self.rnd
        if m._convs["input_blocks.0.0"].hp == "split_packed":
            stem_mode |= _lib.OP_SPLIT | ((m.cin_pad // 2) << 8)      # row = [hi(27) 0.. | lo(27) 0..]
        self.emit("hl_nchw_to_nhwc", _ptr(self.x_in), None, _ptr(xin), self.dt, B, m.in_channels, H * W,
                  m.cin_pad, stem_mode)

        # --- encoder, middle (unet.py:589-592) and ControlNet encoder (unet.py:594-602) ---
        # The two encoders are independent until the decoder.  With `concurrent` the ControlNet branch is
        # issued on a side stream (own scratch buffers): wherever a kernel of one branch leaves SMs idle
        # (the 8^2 .. 32^2 layers launch 24-96 CTAs) the other branch fills them.  The only cross
        # dependency -- the projection that adds hs[i] -- waits on an event recorded after main block i.
        controlnet = m._enc_cond is not None
        self.concurrent = controlnet and m.concurrent_encoders
        self.hs, self.keep_hs = None, controlnet
        self.decoding = False                  # set once the two encoders have joined
        self.side_skip = self.concurrent and m.side_skip
        self.n_events = len(m._enc) + 1 if self.concurrent else 0
        self.film_event = len(m._enc)          # recorded on the main stream once the FiLM table is complete
        if controlnet:
            xcin = self.opbuf("xcin", B * H * W * m.cin_pad)
            self.emit("hl_nchw_to_nhwc", _ptr(self.x_in), _ptr(self.xc_in), _ptr(xcin), self.dt, B,
                      m.in_channels, H * W, m.cin_pad, stem_mode)
        if self.concurrent:
            self.emit_sync("fork")             # the ControlNet stem conv needs only xcin: it overlaps the embedding GEMVs
        # --- embeddings (unet.py:564,584-586; all ResBlock emb_layers in one GEMV) ---
        temb, e1, emb = self.buf("temb", B * mc), self.buf("e1", B * ed), self.buf("emb", B * ed)
        self.film = self.buf("film", B * m._film_rows)
        self.emit("hl_timestep_embedding", _ptr(self.t_in), _ptr(m._freqs), B, mc, _ptr(temb))
        self.emit("hl_linear_small", _ptr(temb), _ptr(m._small["time_embed.0.weight"]),
                  _ptr(m._small["time_embed.0.bias"]), _ptr(e1), B, mc, ed, 0, None, None)
        if m.num_classes is not None:
            self.emit("hl_linear_small", _ptr(e1), _ptr(m._small["time_embed.2.weight"]),
                      _ptr(m._small["time_embed.2.bias"]), _ptr(emb), B, ed, ed, 1,
                      _ptr(m._small["label_emb.weight"]), _ptr(self.y_in))
        else:
            self.emit("hl_linear_small", _ptr(e1), _ptr(m._small["time_embed.2.weight"]),
                      _ptr(m._small["time_embed.2.bias"]), _ptr(emb), B, ed, ed, 1, None, None)
        self.emit("hl_linear_small", _ptr(emb), _ptr(m._film_w), _ptr(m._film_b), _ptr(self.film), B, ed,
                  m._film_rows, 1, None, None)

        if self.concurrent:
            self.emit_sync("signal", self.film_event)
        hs = self.encoder(m._enc, _ptr(xin), "hs", self.cat_skip, False)
        if controlnet:
            self.hs = hs
        x = self.layers(m._mid, hs[-1], "mid", dst=self.cat_h[0])
        if controlnet:
            self.branch = 1 if self.concurrent else 0
            self.encoder(m._enc_cond, _ptr(xcin), "hc", self.cat_skip, True)
            self.branch = 0
        if self.concurrent:
            self.emit_sync("join")

        # --- decoder (unet.py:604-609) ---
        self.decoding = True
        ndec = len(m._dec)
        for j, layers in enumerate(m._dec):
            dst = self.cat_h[j + 1] if j + 1 < ndec else None
            x = self.layers(layers, self.cat_full[j], f"dec{j}", dst=dst)

        # --- out: GN -> SiLU -> conv3x3 (unet.py:471-475,612) ---
        act = _ptr(self.scratch("act", self.max_act, op=True))
        ldo, omode = (2 * x.C, _lib.OP_SPLIT) if m._convs["out.2"].hp == "split_a" else (x.C, 0)
        assert B * H * W * ldo <= self.max_act
        self.gn("out.0", x, act, ldo, True, out_mode=omode)
        co_pad = 32 * ((m.out_channels + 31) // 32)
        if m._convs["out.2"].hp == "split_w":
            # result = [conv with W_hi | conv with W_lo]: the halves are summed on the way to NCHW
            eps = _Ref(_ptr(self.buf("eps_nhwc", B * H * W * 2 * co_pad)), 2 * co_pad, 2 * co_pad, H, W, None, 0)
            self.conv("out.2", act, ldo, None, eps, H, W, want_stats=False)
            self.emit("hl_nhwc_to_nchw_sum2", eps.ptr, 2 * co_pad, co_pad, _ptr(self.out), B, m.out_channels, H * W)
        else:
            eps = _Ref(_ptr(self.buf("eps_nhwc", B * H * W * co_pad)), co_pad, m.out_channels, H, W, None, 0)
            self.conv("out.2", act, ldo, None, eps, H, W, want_stats=False)
            self.emit("hl_nhwc_to_nchw", eps.ptr, co_pad, _ptr(self.out), B, m.out_channels, H * W)

        # --- resolve the symbolic statistics / FiLM pointers ---
        self.stats = torch.zeros(max(self._stats_off, 2), device=dev, dtype=torch.float64)
        sbase, fbase = _ptr(self.stats), _ptr(self.film)

        def fix(a):
            if isinstance(a, tuple):
                if a[0] == "stats":
                    return sbase + 8 * a[1]
                if a[0] == "film":
                    return fbase + 4 * a[1]
                if a[0] == "stats_bytes":
                    return 8 * self.stats.numel()
            return a
        self.calls = [(name, tuple(fix(a) for a in args), br) for name, args, br in self.calls]

    # ---------------------------------------------------------------- execution
    def _launch_all(self):
        main = torch.cuda.current_stream(self.device)
        if self.n_events and self.side is None:
            self.side = torch.cuda.Stream(self.device)
            self.events = [torch.cuda.Event() for _ in range(self.n_events)]
        streams = (main.cuda_stream, self.side.cuda_stream if self.side is not None else main.cuda_stream)
        lib = _lib.load()
        prev = lib.hl_set_pdl(1 if self.m.programmatic_launch else 0)
        serialize = [True, True]        # first launch of a stream / after a cross-stream wait: a normal launch
        # split-K partial sums of the 8^2 / 16^2 layers: one workspace per concurrently running stream
        if self.splitk_ws is None:
            n = 2 if self.side is not None else 1
            self.splitk_ws = [torch.empty(self.SPLITK_BYTES // 4, device=self.device) for _ in range(n)]
        for ws, st in zip(self.splitk_ws if self.m.split_k else [], streams):
            _lib.check(lib.hl_conv_set_workspace(_ptr(ws), self.SPLITK_BYTES, st), "hl_conv_set_workspace")
        lib.hl_conv_set_split_reduce(1 if self.m.split_reduce_in_kernel else 0)
        try:
            for name, args, br in self.calls:
                if name[0] != "#":
                    if serialize[br]:
                        lib.hl_pdl_barrier()
                        serialize[br] = False
                    call(name, *args, streams[br])
                elif name == "#fork":
                    self.side.wait_stream(main)
                    serialize[1] = True
                elif name == "#join":
                    main.wait_stream(self.side)
                    serialize[0] = True
                elif name == "#signal":
                    self.events[args[0]].record(main)
                elif name == "#wait":
                    self.side.wait_event(self.events[args[0]])
                    serialize[1] = True
        finally:
            lib.hl_set_pdl(prev)
            for st in set(streams):
                lib.hl_conv_set_workspace(None, 0, st)     # the registry must not outlive this plan's memory

    def load_inputs(self, x, timesteps, x_cond, y):
        self.x_in.copy_(x)
        self.t_in.copy_(timesteps)
        if self.xc_in is not None:
            self.xc_in.copy_(x_cond)
        if self.m.num_classes is not None:
            self.y_in.copy_(y)

    def result(self):
        return self.out.clone()

    def run(self, x, timesteps, x_cond, y):
        m = self.m
        self.load_inputs(x, timesteps, x_cond, y)
        if not m.use_cuda_graph or self.runs == 0:
            lib, n0 = _lib.load(), _lib.launch_count
            k0 = lib.hl_launch_count()
            self._launch_all()                       # first run is eager: function attributes, driver entry points
            self.kernels_per_run = lib.hl_launch_count() - k0     # a split-K conv is two kernels
            _lib.launch_count = n0 + self.kernels_per_run
        else:
            if self.graph is None:
                n0 = _lib.launch_count
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._launch_all()
                self.graph = g
                _lib.launch_count = n0               # capture records launches, it does not run them
            self.graph.replay()
            _lib.launch_count += self.n_launches
        self.runs += 1
        return self.result()

    @property
    def n_launches(self):
        if self.kernels_per_run:
            return self.kernels_per_run
        return sum(1 for name, _, _ in self.calls if name[0] != "#")


class _SplitPlan(_StepPlan):
    """The batch cut into ``parts`` independent sub-batches, each with its own _StepPlan (workspace, statistics
    arena, split-K workspaces), all launched inside ONE CUDA graph on separate streams.  Samples never interact
    (GroupNorm is per sample), so the chains are independent for the whole forward: where one chain sits in a
    low-resolution, latency-bound stretch (8^2 .. 32^2: a few dozen CTAs per kernel) the other chains' kernels
    fill the idle SMs -- including the decoder, which has no ControlNet branch to overlap with."""

    def __init__(self, model, device, B, H, W, parts):        # noqa: super().__init__ builds a plan; this only wraps
        self.m, self.device, self.B, self.H, self.W = model, device, B, H, W
        self.parts, self.b = parts, B // parts
        self.subs = [_StepPlan(model, device, self.b, H, W) for _ in range(parts)]
        self.streams = None
        self.graph, self.runs, self.kernels_per_run = None, 0, 0
        self.calls = [c for p in self.subs for c in p.calls]

    def load_inputs(self, x, timesteps, x_cond, y):
        b = self.b
        timesteps = torch.as_tensor(timesteps, device=x.device)
        for k, p in enumerate(self.subs):
            sl = slice(k * b, (k + 1) * b)
            p.load_inputs(x[sl], timesteps[sl], x_cond[sl] if x_cond is not None else None,
                          y[sl] if y is not None else None)

    def result(self):
        return torch.cat([p.out for p in self.subs], 0)

    def _launch_all(self):
        main = torch.cuda.current_stream(self.device)
        if self.streams is None:
            self.streams = [torch.cuda.Stream(self.device) for _ in self.subs]
        for p, s in zip(self.subs, self.streams):
            s.wait_stream(main)
            with torch.cuda.stream(s):
                p._launch_all()
        for s in self.streams:
            main.wait_stream(s)
