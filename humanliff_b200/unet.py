"""B200-native ``UNetModel`` -- the drop-in for human_diffusion/improved_diffusion/unet.py:300-615.

Same constructor flags, same ``forward(x, timesteps, x_cond=None, y=None)`` (NCHW fp32 in / out) and the
same 953-tensor ``state_dict()`` key set / shapes, so reference checkpoints load unchanged.  Nothing
here computes with PyTorch: torch owns device memory and the stream, every FLOP is a call into
``libhumanliff_b200.so`` (``include/humanliff_b200.h``).

Internals
  * activations are NHWC fp32 buffers owned by a per-shape workspace (no allocation after the first
    call of a shape -> the whole step is CUDA-graph capturable);
  * conv / conv1d weights are re-packed once to ``[tap][Cout_pad][Cin_pad]`` (TMA / tcgen05 friendly),
    TF32-rounded in ``precision="tf32"`` mode;
  * all 62 ResBlock ``emb_layers`` Linear layers are stacked into one matrix and evaluated by a
    single streaming GEMV per step (unet.py:151-157,200);
  * GroupNorm32 + SiLU + FiLM (unet.py:204-206) is one stats pass + one fused apply pass that writes
    the next conv's TF32 operand.
"""
import math

import torch
import torch.nn as nn

from . import _lib
from ._lib import call


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


def pack_conv(weight, bias, cin_pad=None, round_tf32=True, device=None, stream=None):
    """OIHW (or Conv1d [O, I, 1]) weight -> packed ``[kh*kw][Cout_pad][Cin_pad]`` fp32 + padded bias.
    ``round_tf32`` applies cvt.rna.tf32 on the device (B operand of tcgen05 kind::tf32)."""
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
    if round_tf32:
        if stream is None:
            stream = torch.cuda.current_stream(device).cuda_stream
        flat = pk.view(-1, 4)
        call("hl_round_tf32", _ptr(flat), 4, _ptr(flat), 4, 4, flat.shape[0], stream)
    b = torch.zeros(cout_pad, device=device, dtype=torch.float32)
    if bias is not None:
        b[:cout] = bias.detach().to(device=device, dtype=torch.float32)
    return pk, b


class _Conv:
    """One packed convolution / conv1d / linear-over-pixels."""

    def __init__(self, name, cin, cout, ksize, stride=1, cin_pad=None):
        self.name, self.cin, self.cout, self.ksize, self.stride = name, cin, cout, ksize, stride
        self.cin_pad = cin_pad or cin
        self.w = None
        self.b = None


class UNetModel(nn.Module):
    """See module docstring.  Supported envelope: ``cond_type`` in {"controlnet", ""},
    ``use_scale_shift_norm=True``, ``use_3d_aware=False``, ``dims=2``, ``conv_resample=True``,
    ``dropout`` ignored at inference; anything else raises ``NotImplementedError``."""

    def __init__(self, in_channels, model_channels, out_channels, num_res_blocks, attention_resolutions,
                 dropout=0, channel_mult=(1, 2, 4, 8), conv_resample=True, dims=2, num_classes=None,
                 use_checkpoint=False, num_heads=1, num_heads_upsample=-1, use_scale_shift_norm=False,
                 cond_type="", use_3d_aware=False, transformer_depth=1, context_dim=None,
                 precision="tf32"):
        super().__init__()
        if cond_type not in ("controlnet", ""):
            raise NotImplementedError(f"cond_type={cond_type!r}: only 'controlnet' and '' are built")
        if not use_scale_shift_norm:
            raise NotImplementedError("use_scale_shift_norm=False is outside the production envelope")
        if use_3d_aware or dims != 2 or not conv_resample:
            raise NotImplementedError("use_3d_aware / dims != 2 / conv_resample=False are not built")
        if precision not in ("tf32", "fp32"):
            raise ValueError("precision must be 'tf32' or 'fp32'")
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
        self.cin_pad = (in_channels + 31) // 32 * 32

        self._convs = {}
        self._film = []          # (prefix, cout, offset) in stacking order
        self._film_rows = 0
        self._build_plan()
        self._packed_key = None
        self._ws = {}

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
        rnd = self.precision == "tf32"
        lib = _lib.load()
        stream = torch.cuda.current_stream(device).cuda_stream
        for c in self._convs.values():
            c.w, c.b = pack_conv(self._p(c.name + ".weight"), self._p(c.name + ".bias"), c.cin_pad, rnd,
                                 device, stream)
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
        self._packed_key = key

    # ------------------------------------------------------------------ workspace
    class _WS:
        def __init__(self, device):
            self.device = device
            self.bufs = {}

        def get(self, name, *shape, dtype=torch.float32):
            t = self.bufs.get(name)
            if t is None or tuple(t.shape) != tuple(shape):
                t = torch.empty(*shape, device=self.device, dtype=dtype)
                self.bufs[name] = t
            return t

    def _workspace(self, device, B, H, W):
        key = (str(device), B, H, W)
        ws = self._ws.get(key)
        if ws is None:
            ws = UNetModel._WS(device)
            self._ws[key] = ws
        return ws

    def _scratch_sizes(self, B, H, W):
        """Largest [pixels x channels] activation (concat inputs included) and largest qkv tensor."""
        mc, cm = self.model_channels, self.channel_mult
        act, qkv = B * H * W * max(self.cin_pad, mc), 4
        for l, m in enumerate(cm):
            pix = B * (H >> l) * (W >> l)
            below = cm[min(l + 1, len(cm) - 1)]
            act = max(act, pix * mc * (m + max(m, below)))
            if (1 << l) in self.attention_resolutions:
                qkv = max(qkv, pix * 3 * mc * max(m, below))
        return act, qkv

    # ------------------------------------------------------------------ op helpers
    def _conv_call(self, cname, x, ldx, res, ldr, y, ldy, B, H, W, flags=0):
        """x / res / y are raw device pointers (ints); res may be None."""
        c = self._convs[cname]
        if self.precision == "fp32":
            flags |= _lib.CONV_FORCE_SIMT
        call("hl_conv2d", x, ldx, _ptr(c.w), _ptr(c.b), res, ldr, y, ldy, B, H, W, c.cin_pad, c.cout,
             c.ksize, c.stride, flags, self._stream)

    def _uses_tc(self, cname, B, H, W, ldx, flags=0):
        c = self._convs[cname]
        if self.precision == "fp32":
            return False
        return bool(_lib.load().hl_conv2d_uses_tensor_cores(B, H, W, c.cin_pad, c.cout, c.ksize, c.stride,
                                                            ldx, flags))

    def _gn(self, nname, x, ldx, C, B, HW, out, ldo, silu, film=None):
        ws = self._cur_ws
        sums = ws.get("gn_sums", B * 32 * 2, dtype=torch.float64)
        call("hl_gn_stats", x, ldx, B, HW, C, 32, _ptr(sums), self._stream)
        call("hl_gn_apply", x, ldx, _ptr(sums), _ptr(self._norm_p[nname + ".weight"]),
             _ptr(self._norm_p[nname + ".bias"]), film, self._film_rows if film is not None else 0,
             out, ldo, B, HW, C, 32, 1e-5, 1 if silu else 0, 1 if self.precision == "tf32" else 0,
             self._stream)

    def _raw_operand(self, cname, x, ldx, C, B, H, W):
        """A-operand staging for convs that read the raw residual stream: the tensor-core path wants
        round-to-nearest TF32 operands (tcgen05 would otherwise truncate the low mantissa bits)."""
        if not self._uses_tc(cname, B, H, W, ldx):
            return x, ldx
        ws = self._cur_ws
        buf = ws.get("stage", self._max_act)
        call("hl_round_tf32", x, ldx, _ptr(buf), C, C, B * H * W, self._stream)
        return _ptr(buf), C

    def _run_res(self, blk, x, ldx, B, H, W, out, ldo):
        ws = self._cur_ws
        HW = H * W
        cin, cout = blk["cin"], blk["cout"]
        act = _ptr(ws.get("act", self._max_act))
        h = _ptr(ws.get("h", self._max_act))
        self._gn(blk["n1"], x, ldx, cin, B, HW, act, cin, True)
        self._conv_call(blk["c1"], act, cin, None, 0, h, cout, B, H, W)
        film = self._film_ptr + 4 * blk["film_off"]
        self._gn(blk["n2"], h, cout, cout, B, HW, act, cout, True, film=film)
        if blk["skip"] is not None:
            xs, lds = self._raw_operand(blk["skip"], x, ldx, cin, B, H, W)
            s = _ptr(ws.get("skipbuf", self._max_act))
            self._conv_call(blk["skip"], xs, lds, None, 0, s, cout, B, H, W)
            self._conv_call(blk["c2"], act, cout, s, cout, out, ldo, B, H, W)
        else:
            self._conv_call(blk["c2"], act, cout, x, ldx, out, ldo, B, H, W)

    def _run_attn(self, blk, x, ldx, B, H, W, out, ldo):
        ws = self._cur_ws
        T, C = H * W, blk["c"]
        act = _ptr(ws.get("act", self._max_act))
        qkv = _ptr(ws.get("qkv", self._max_qkv))
        att = _ptr(ws.get("h", self._max_act))
        self._gn(blk["n"], x, ldx, C, B, T, act, C, False)
        self._conv_call(blk["qkv"], act, C, None, 0, qkv, 3 * C, B, H, W)
        call("hl_attention", qkv, 3 * C, att, C, B, T, C, self.num_heads,
             1 if self.precision == "tf32" else 0, self._stream)
        self._conv_call(blk["proj"], att, C, x, ldx, out, ldo, B, H, W)

    def _run_layers(self, layers, x, ldx, B, H, W, out_name):
        """Run one TimestepEmbedSequential (unet.py:41-49).  Returns (ptr, channels, H, W)."""
        ws = self._cur_ws
        n = len(layers)
        C = None
        for li, blk in enumerate(layers):
            last = li == n - 1
            kind = blk["kind"]
            if kind == "res":
                C = blk["cout"]
                dst = ws.get(out_name if last else f"{out_name}.t{li}", B * H * W * C)
                self._run_res(blk, x, ldx, B, H, W, _ptr(dst), C)
            elif kind == "attn":
                C = blk["c"]
                dst = ws.get(out_name if last else f"{out_name}.t{li}", B * H * W * C)
                self._run_attn(blk, x, ldx, B, H, W, _ptr(dst), C)
            elif kind == "down":
                C = blk["ch"]
                dst = ws.get(out_name, B * (H // 2) * (W // 2) * C)
                self._conv_call(blk["c"], x, ldx, None, 0, _ptr(dst), C, B, H, W)
                H, W = H // 2, W // 2
            elif kind == "up":
                C = blk["ch"]
                up = ws.get("upbuf", self._max_act)
                call("hl_upsample2x", x, ldx, _ptr(up), C, B, H, W, C,
                     1 if self.precision == "tf32" else 0, self._stream)
                H, W = 2 * H, 2 * W
                dst = ws.get(out_name, B * H * W * C)
                self._conv_call(blk["c"], _ptr(up), C, None, 0, _ptr(dst), C, B, H, W)
            else:
                raise AssertionError(kind)
            x, ldx = _ptr(dst), C
        return x, C, H, W

    def _run_encoder(self, enc, xin, B, H, W, tag, proj):
        """Returns the list of skip tensors [(ptr, C, H, W)] (hs / hs_cond of unet.py:589-602)."""
        ws = self._cur_ws
        outs = []
        mc = self.model_channels
        name = f"{tag}0" if not proj else f"{tag}raw0"
        h0 = ws.get(name, B * H * W * mc)
        self._conv_call(enc[0][0]["c"], _ptr(xin), self.cin_pad, None, 0, _ptr(h0), mc, B, H, W)
        x, C = _ptr(h0), mc
        if proj:
            x = self._proj(0, x, C, B, H, W, f"{tag}0")
        outs.append((x, C, H, W))
        for i in range(1, len(enc)):
            name = f"{tag}{i}" if not proj else f"{tag}raw{i}"
            x, C, H, W = self._run_layers(enc[i], x, C, B, H, W, name)
            if proj:
                x = self._proj(i, x, C, B, H, W, f"{tag}{i}")
            outs.append((x, C, H, W))
        return outs

    def _proj(self, i, x, C, B, H, W, out_name):
        """h_cond = input_blocks_proj_cond[i](h_cond)   (unet.py:600; replaces h_cond)."""
        cname = f"input_blocks_proj_cond.{i}"
        dst = self._cur_ws.get(out_name, B * H * W * C)
        xs, lds = self._raw_operand(cname, x, C, C, B, H, W)
        self._conv_call(cname, xs, lds, None, 0, _ptr(dst), C, B, H, W)
        return _ptr(dst)

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def forward(self, x, timesteps, x_cond=None, y=None):
        """ε-prediction.  x: [B, C, H, W] fp32 CUDA, timesteps: [B] (int or float), x_cond like x
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
            self._pack(device)
            self._stream = torch.cuda.current_stream(device).cuda_stream
            ws = self._workspace(device, B, H, W)
            self._cur_ws = ws
            self._max_act, self._max_qkv = self._scratch_sizes(B, H, W)
            x = x.contiguous().float()
            mc, ed = self.model_channels, self.emb_dim

            # --- embeddings (unet.py:564,584-586; all ResBlock emb_layers in one GEMV) ---
            tf = ws.get("t", B)
            tf.copy_(timesteps)
            temb = ws.get("temb", B * mc)
            e1 = ws.get("e1", B * ed)
            emb = ws.get("emb", B * ed)
            film = ws.get("film", B * self._film_rows)
            call("hl_timestep_embedding", _ptr(tf), _ptr(self._freqs), B, mc, _ptr(temb), self._stream)
            call("hl_linear_small", _ptr(temb), _ptr(self._small["time_embed.0.weight"]),
                 _ptr(self._small["time_embed.0.bias"]), _ptr(e1), B, mc, ed, 0, None, None, self._stream)
            if self.num_classes is not None:
                yb = ws.get("y", B, dtype=torch.int64)
                yb.copy_(y)
                call("hl_linear_small", _ptr(e1), _ptr(self._small["time_embed.2.weight"]),
                     _ptr(self._small["time_embed.2.bias"]), _ptr(emb), B, ed, ed, 1,
                     _ptr(self._small["label_emb.weight"]), _ptr(yb), self._stream)
            else:
                call("hl_linear_small", _ptr(e1), _ptr(self._small["time_embed.2.weight"]),
                     _ptr(self._small["time_embed.2.bias"]), _ptr(emb), B, ed, ed, 1, None, None,
                     self._stream)
            call("hl_linear_small", _ptr(emb), _ptr(self._film_w), _ptr(self._film_b), _ptr(film), B, ed,
                 self._film_rows, 1, None, None, self._stream)
            self._film_ptr = _ptr(film)

            rnd = 1 if self.precision == "tf32" else 0
            xin = ws.get("xin", B * H * W * self.cin_pad)
            call("hl_nchw_to_nhwc", _ptr(x), None, _ptr(xin), B, Cx, H * W, self.cin_pad, rnd, self._stream)

            # --- encoder, middle (unet.py:589-592) ---
            hs = self._run_encoder(self._enc, xin, B, H, W, "hs", False)
            hx, hC, hH, hW = hs[-1]
            hx, hC, hH, hW = self._run_layers(self._mid, hx, hC, B, hH, hW, "mid")

            # --- ControlNet encoder (unet.py:594-602) ---
            hs_cond = None
            if self._enc_cond is not None:
                xc = x_cond.contiguous().float()
                xcin = ws.get("xcin", B * H * W * self.cin_pad)
                call("hl_nchw_to_nhwc", _ptr(x), _ptr(xc), _ptr(xcin), B, Cx, H * W, self.cin_pad, rnd,
                     self._stream)
                hs_cond = self._run_encoder(self._enc_cond, xcin, B, H, W, "hc", True)

            # --- decoder (unet.py:604-609) ---
            for j, layers in enumerate(self._dec):
                sx, sC, sH, sW = hs.pop()
                assert (sH, sW) == (hH, hW) and layers[0]["cat"] == (hC, sC)
                cat = ws.get(f"cat{j}", B * hH * hW * (hC + sC))
                cx = hs_cond.pop()[0] if hs_cond is not None else None
                call("hl_concat_add", hx, hC, hC, sx, sC, cx, sC, sC, _ptr(cat), hC + sC, B * hH * hW,
                     self._stream)
                hx, hC, hH, hW = self._run_layers(layers, _ptr(cat), hC + sC, B, hH, hW, f"dec{j}")

            # --- out: GN -> SiLU -> conv3x3 (unet.py:471-475,612) ---
            act = _ptr(ws.get("act", self._max_act))
            self._gn("out.0", hx, hC, hC, B, hH * hW, act, hC, True)
            co_pad = 32 * ((self.out_channels + 31) // 32)
            eps_nhwc = ws.get("eps_nhwc", B * H * W * co_pad)
            self._conv_call("out.2", act, hC, None, 0, _ptr(eps_nhwc), co_pad, B, H, W)
            out = torch.empty(B, self.out_channels, H, W, device=device, dtype=torch.float32)
            call("hl_nhwc_to_nchw", _ptr(eps_nhwc), co_pad, _ptr(out), B, self.out_channels, H * W,
                 self._stream)
        return out
