"""GPU: every C-ABI kernel against a plain torch-CPU statement of the same op (the oracle's arithmetic
library) on seeded inputs.  Tolerances: 1e-5 (fp32-accumulate kernels on identical operands: accumulation-
order differences only), 1e-3 rel-L2 for the reduced-precision-operand tensor-core convolution against the
exact fp32 conv (north_star bar), bit-exact for layout / cast / DDPM update."""
import math

import pytest
import torch
import torch.nn.functional as F

from common import rel_l2, rel_max

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _call(name, *a):
    from humanliff_b200._lib import call
    call(name, *a)


def test_layout_roundtrip(dev):
    g = torch.Generator().manual_seed(0)
    for B, C, H, W, ld in [(2, 27, 16, 16, 32), (1, 192, 8, 8, 192), (3, 5, 7, 9, 8)]:
        x = torch.randn(B, C, H, W, generator=g)
        x2 = torch.randn(B, C, H, W, generator=g)
        xd, x2d = x.to(dev), x2.to(dev)
        nhwc = torch.full((B, H * W, ld), 7.0, device=dev)
        _call("hl_nchw_to_nhwc", xd.data_ptr(), x2d.data_ptr(), nhwc.data_ptr(), 0, B, C, H * W, ld, 0, _stream())
        ref = (x + x2).permute(0, 2, 3, 1).reshape(B, H * W, C)
        assert torch.equal(nhwc[:, :, :C].cpu(), ref)
        assert float(nhwc[:, :, C:].abs().max() if ld > C else 0) == 0.0, "channel padding must be zero"
        back = torch.empty(B, C, H, W, device=dev)
        _call("hl_nhwc_to_nchw", nhwc.data_ptr(), ld, back.data_ptr(), B, C, H * W, _stream())
        assert torch.equal(back.cpu(), x + x2)


def test_layout_fp16_operand(dev):
    g = torch.Generator().manual_seed(0)
    B, C, H, W, ld = 2, 27, 8, 8, 64
    x = torch.randn(B, C, H, W, generator=g)
    nhwc = torch.full((B, H * W, ld), 7.0, device=dev, dtype=torch.float16)
    _call("hl_nchw_to_nhwc", x.to(dev).data_ptr(), None, nhwc.data_ptr(), 1, B, C, H * W, ld, 0, _stream())
    ref = x.permute(0, 2, 3, 1).reshape(B, H * W, C).half()
    assert torch.equal(nhwc[:, :, :C].cpu(), ref) and float(nhwc[:, :, C:].abs().max()) == 0.0


def test_cast_operand_matches_emulation(dev):
    from oracle.unet_oracle import round_tf32
    x = torch.randn(1000, 8) * torch.logspace(-6, 4, 1000)[:, None]
    xd = x.to(dev)
    out = torch.empty_like(xd)
    _call("hl_cast_operand", xd.data_ptr(), 8, out.data_ptr(), 0, 8, 8, 1000, 1, _stream())
    assert torch.equal(out.cpu(), round_tf32(x))
    _call("hl_cast_operand", xd.data_ptr(), 8, out.data_ptr(), 0, 8, 8, 1000, 0, _stream())
    assert torch.equal(out.cpu(), x)
    outh = torch.empty(1000, 8, device=dev, dtype=torch.float16)
    _call("hl_cast_operand", xd.data_ptr(), 8, outh.data_ptr(), 1, 8, 8, 1000, 0, _stream())
    assert torch.equal(outh.cpu(), x.half())          # round-to-nearest-even, like torch


def test_upsample_and_summed_layout(dev):
    g = torch.Generator().manual_seed(1)
    # hl_nhwc_to_nchw_sum2: the two halves of a split-weight conv result summed on the way to NCHW
    B, C, HW, ld, off2 = 2, 27, 70, 64, 32
    src = torch.randn(B, HW, ld, generator=g)
    out = torch.empty(B, C, HW, device=dev)
    _call("hl_nhwc_to_nchw_sum2", src.to(dev).data_ptr(), ld, off2, out.data_ptr(), B, C, HW, _stream())
    assert torch.equal(out.cpu(), (src[..., :C] + src[..., off2:off2 + C]).permute(0, 2, 1))
    x = torch.randn(2, 3, 5, 8, generator=g)                       # NHWC [B,H,W,C]
    ref = F.interpolate(x.permute(0, 3, 1, 2), scale_factor=2, mode="nearest").permute(0, 2, 3, 1)
    xd = x.to(dev)
    for code, tdt in ((0, torch.float32), (1, torch.float16)):
        up = torch.empty(2, 6, 10, 8, device=dev, dtype=tdt)
        _call("hl_upsample2x", xd.data_ptr(), 8, up.data_ptr(), code, 8, 2, 3, 5, 8, 0, _stream())
        assert torch.equal(up.cpu(), ref.to(tdt))


def test_embeddings(dev):
    from oracle.unet_oracle import timestep_embedding
    t = torch.tensor([0., 1., 401., 999., 250.5])
    out = torch.empty(5, 192, device=dev)
    td = t.to(dev)
    freqs = torch.exp(-math.log(10000) * torch.arange(96, dtype=torch.float32) / 96).to(dev)
    _call("hl_timestep_embedding", td.data_ptr(), freqs.data_ptr(), 5, 192, out.data_ptr(), _stream())
    assert rel_max(out, timestep_embedding(t, 192)) < 2e-6
    g = torch.Generator().manual_seed(2)
    for B in (1, 4, 11):
        x, W, b = torch.randn(B, 768, generator=g), torch.randn(1000, 768, generator=g) / 27, torch.randn(1000, generator=g)
        tab, idx = torch.randn(4, 1000, generator=g), torch.randint(0, 4, (B,), generator=g)
        y = torch.empty(B, 1000, device=dev)
        xd, Wd, bd, tabd, idxd = x.to(dev), W.to(dev), b.to(dev), tab.to(dev), idx.to(dev)
        _call("hl_linear_small", xd.data_ptr(), Wd.data_ptr(), bd.data_ptr(), y.data_ptr(), B, 768, 1000, 1,
              tabd.data_ptr(), idxd.data_ptr(), _stream())
        ref = F.linear(x * torch.sigmoid(x), W, b) + tab[idx]
        assert rel_l2(y, ref) < 1e-6, B


@pytest.mark.parametrize("C,HW,B", [(192, 64 * 64, 2), (384, 256, 1), (576, 1024, 1), (1152, 64, 3), (1536, 16, 2), (64, 4, 2)])
def test_groupnorm_silu_film(dev, C, HW, B):
    g = torch.Generator().manual_seed(C + HW)
    x = torch.randn(B, HW, C, generator=g) * 2 + 0.5
    gamma, beta = 1 + 0.1 * torch.randn(C, generator=g), 0.1 * torch.randn(C, generator=g)
    film = 0.3 * torch.randn(B, 2 * C + 10, generator=g)
    xd, gd, bd, fd = x.to(dev), gamma.to(dev), beta.to(dev), film.to(dev)
    ld_st = C + 8                                                   # statistics row wider than the tensor
    stats = torch.zeros(B * ld_st * 2, device=dev, dtype=torch.float64)
    _call("hl_gn_stats", xd.data_ptr(), C, B, HW, C, stats.data_ptr(), ld_st, _stream())
    st = stats.cpu().reshape(B, ld_st, 2)
    assert rel_l2(st[:, :C, 0], x.double().sum(1)) < 1e-6 and rel_l2(st[:, :C, 1], (x.double() ** 2).sum(1)) < 1e-6
    assert float(st[:, C:].abs().max()) == 0.0
    xn = F.group_norm(x.permute(0, 2, 1), 32, gamma, beta, eps=1e-5)          # [B, C, HW]
    for use_film, silu in [(False, True), (True, True), (False, False)]:
        r = xn
        if use_film:
            r = r * (1 + film[:, :C, None]) + film[:, C:2 * C, None]
        if silu:
            r = r * torch.sigmoid(r)
        for code, tdt, tol in ((0, torch.float32, 2e-6), (1, torch.float16, 6e-4)):
            y = torch.empty(B, HW, C, device=dev, dtype=tdt)
            raw = torch.empty(B, HW, C, device=dev, dtype=tdt)
            _call("hl_gn_apply", xd.data_ptr(), C, stats.data_ptr(), ld_st, gd.data_ptr(), bd.data_ptr(),
                  fd.data_ptr() if use_film else None, 2 * C + 10, y.data_ptr(), code, C, raw.data_ptr(), C, B, HW,
                  C, 32, 1e-5, 1 if silu else 0, 0, _stream())
            assert rel_l2(y.float().permute(0, 2, 1), r) < tol, (use_film, silu, code)
            assert torch.equal(raw.cpu(), x.to(tdt)), "raw operand copy of the input"
            if code == 1:   # the fp16 operand is the rounded fp32 result; SiLU uses ex2/rcp.approx here (2e-7), so a
                #             value sitting on an fp16 rounding boundary may land one fp16 ulp away
                y32 = torch.empty(B, HW, C, device=dev)
                _call("hl_gn_apply", xd.data_ptr(), C, stats.data_ptr(), ld_st, gd.data_ptr(), bd.data_ptr(),
                      fd.data_ptr() if use_film else None, 2 * C + 10, y32.data_ptr(), 0, C, None, 0, B, HW, C, 32,
                      1e-5, 1 if silu else 0, 0, _stream())
                exact = y32.cpu().half()
                assert float((y.cpu() != exact).float().mean()) < 2e-3
                assert rel_l2(y.float(), exact.float()) < 2e-5


@pytest.mark.parametrize("C,HW,B", [(192, 64 * 64, 2), (768, 64, 3), (384, 1024, 1), (36, 16, 2)])
def test_groupnorm_fp16_input(dev, C, HW, B):
    """HL_OP_X_F16: the GroupNorm input is an fp16 tensor (a conv's HL_CONV_OUT_F16 result): same output as the fp32
    path fed the same (exactly representable) values -- both the 8-channel fast path and the generic one (C = 36)."""
    from humanliff_b200 import _lib
    g = torch.Generator().manual_seed(C + HW)
    x = (torch.randn(B, HW, C, generator=g) * 2 + 0.5).half()
    groups = 32 if C % 32 == 0 else 4
    gamma, beta = 1 + 0.1 * torch.randn(C, generator=g), 0.1 * torch.randn(C, generator=g)
    film = 0.3 * torch.randn(B, 2 * C, generator=g)
    x16, x32, gd, bd, fd = x.to(dev), x.float().to(dev), gamma.to(dev), beta.to(dev), film.to(dev)
    stats = torch.zeros(B * C * 2, device=dev, dtype=torch.float64)
    _call("hl_gn_stats", x32.data_ptr(), C, B, HW, C, stats.data_ptr(), C, _stream())
    outs = []
    for xin, flag in ((x32, 0), (x16, _lib.OP_X_F16)):
        y = torch.empty(B, HW, C, device=dev, dtype=torch.float16)
        _call("hl_gn_apply", xin.data_ptr(), C, stats.data_ptr(), C, gd.data_ptr(), bd.data_ptr(), fd.data_ptr(), 2 * C,
              y.data_ptr(), 1, C, None, 0, B, HW, C, groups, 1e-5, 1, flag, _stream())
        outs.append(y.cpu())
    assert torch.equal(outs[0], outs[1])
    xn = F.group_norm(x.float().permute(0, 2, 1), groups, gamma, beta, eps=1e-5) * (1 + film[:, :C, None]) + film[:, C:, None]
    assert rel_l2(outs[1].float().permute(0, 2, 1), xn * torch.sigmoid(xn)) < 6e-4


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 64, 64, 384, 192), (1, 32, 32, 576, 192), (2, 32, 32, 768, 384),
                                             (1, 16, 8, 192, 384), (4, 16, 16, 384, 256), (3, 128, 128, 384, 192),
                                             (4, 256, 256, 384, 192), (4, 128, 128, 576, 192), (4, 64, 64, 768, 384),
                                             (4, 64, 64, 192, 384)])          # the last four: launches of the production step
def test_gn_skip_fused(dev, B, H, W, Cin, Cout):
    """hl_gn_skip (GroupNorm-1 + SiLU operand pass fused with the 1x1 skip conv, hi + lo operand pair built in shared
    memory) against the two launches it replaces: hl_gn_apply with the raw hi | lo copy, then hl_conv2d(HL_CONV_SPLIT3).
    act: equal but for rare one-ulp fp16 differences; skip: fp32 accumulation order only (5e-6)."""
    from humanliff_b200 import _lib
    from humanliff_b200.unet import pack_conv
    lib = _lib.load()
    HW = H * W
    assert lib.hl_gn_skip_supported(B, HW, Cin, Cout) == 1
    g = torch.Generator().manual_seed(Cin + Cout + HW)
    x = (torch.randn(B, HW, Cin, generator=g) * 3 + 0.7).to(dev)
    gamma, beta = (1 + 0.1 * torch.randn(Cin, generator=g)).to(dev), (0.1 * torch.randn(Cin, generator=g)).to(dev)
    w = torch.randn(Cout, Cin, 1, 1, generator=g) / math.sqrt(Cin)
    b = torch.randn(Cout, generator=g) * 0.1
    wpk, bpk = pack_conv(w, b, Cin, "fp16", dev, mode="split")
    stats = torch.zeros(B * Cin * 2, device=dev, dtype=torch.float64)
    _call("hl_gn_stats", x.data_ptr(), Cin, B, HW, Cin, stats.data_ptr(), Cin, _stream())
    # reference: the two-kernel form
    act_ref = torch.empty(B, HW, Cin, device=dev, dtype=torch.float16)
    raw = torch.empty(B, HW, 2 * Cin, device=dev, dtype=torch.float16)
    mode = (_lib.OP_SPLIT | _lib.OP_SCALED) << _lib.OP_RAW_SHIFT
    _call("hl_gn_apply", x.data_ptr(), Cin, stats.data_ptr(), Cin, gamma.data_ptr(), beta.data_ptr(), None, 0,
          act_ref.data_ptr(), 1, Cin, raw.data_ptr(), 2 * Cin, B, HW, Cin, 32, 1e-5, 1, mode, _stream())
    skip_ref = torch.empty(B, HW, Cout, device=dev)
    _call("hl_conv2d", raw.data_ptr(), 1, 2 * Cin, wpk.data_ptr(), bpk.data_ptr(), None, 0, skip_ref.data_ptr(), Cout, None, 0,
          B, H, W, Cin, Cout, 1, 1, _lib.CONV_SPLIT3, _stream())
    # fused
    act = torch.full((B, HW, Cin), float("nan"), device=dev, dtype=torch.float16)
    skip = torch.full((B, HW, Cout), float("nan"), device=dev)
    _call("hl_gn_skip", x.data_ptr(), Cin, stats.data_ptr(), Cin, gamma.data_ptr(), beta.data_ptr(), act.data_ptr(), Cin,
          wpk.data_ptr(), bpk.data_ptr(), skip.data_ptr(), Cout, B, HW, Cin, Cout, 32, 1e-5, _stream())
    torch.cuda.synchronize()
    assert not torch.isnan(skip).any() and not torch.isnan(act.float()).any()
    # about ten elements per million land one fp16 ulp away (the two kernels' SiLU instruction sequences differ in the last bit)
    assert float((act != act_ref).float().mean()) < 1e-4 and rel_max(act.float(), act_ref.float()) < 2e-3
    assert rel_l2(skip, skip_ref) < 5e-6, rel_l2(skip, skip_ref)
    # and against the exact fp32 statement (the hi + lo pair carries ~22 bits: far inside the fp16 operand error)
    ref = x.cpu().double() @ w.reshape(Cout, Cin).double().T + b.double()
    assert rel_l2(skip, ref) < 2e-5, rel_l2(skip, ref)


def _operand(t, mode):
    """fp32 tensor -> (device-ready operand tensor, dtype code, fp32 view of the rounded values)."""
    from oracle.unet_oracle import round_tf32
    if mode == "fp16":
        h = t.half()
        return h, 1, h.float()
    if mode == "tf32":
        r = round_tf32(t)
        return r, 0, r
    return t, 0, t


def _conv_case(dev, B, H, W, Cin, Cout, k, stride, mode="fp32", flags=0, residual=True, stats=False, cin_pad=None,
               seed=0, ldy=None, tuning=None, tuning2=None, out_f16=False, split=None):
    """Runs hl_conv2d on seeded inputs; returns (y NCHW cpu, fp32 reference, reference on the ROUNDED operands,
    stats cpu or None)."""
    from humanliff_b200.unet import pack_conv
    from humanliff_b200 import _lib
    g = torch.Generator().manual_seed(seed)
    cin_pad = cin_pad or Cin
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) / math.sqrt(Cin * k * k)
    b = torch.randn(Cout, generator=g) * 0.1
    ups = bool(flags & _lib.CONV_UPSAMPLE2X)
    xr = F.interpolate(x, scale_factor=2, mode="nearest") if ups else x
    ref = F.conv2d(xr, w, b, stride=stride, padding=k // 2)
    Ho, Wo = ref.shape[2:]
    res = torch.randn(B, Cout, Ho, Wo, generator=g) if residual else None
    xn = torch.zeros(B, H, W, cin_pad)
    xn[..., :Cin] = x.permute(0, 2, 3, 1)
    xop, code, xround = _operand(xn, mode)
    wr = _operand(w, mode)[2]
    ref_r = F.conv2d(F.interpolate(xround[..., :Cin].permute(0, 3, 1, 2), scale_factor=2, mode="nearest") if ups
                     else xround[..., :Cin].permute(0, 3, 1, 2), wr, b, stride=stride, padding=k // 2)
    if residual:
        ref, ref_r = ref + res, ref_r + res
    if mode == "tf32":
        flags |= _lib.CONV_TF32
    xd = xop.to(dev)
    wpk, bpk = pack_conv(w, b, cin_pad, mode, dev)
    ldy = ldy or Cout
    y = torch.full((B, Ho, Wo, ldy), float("nan"), device=dev, dtype=torch.float16 if out_f16 else torch.float32)
    if out_f16:
        flags |= _lib.CONV_OUT_F16
    rd = res.permute(0, 2, 3, 1).contiguous().to(dev) if residual else None
    st = torch.zeros(B * ldy * 2, device=dev, dtype=torch.float64) if stats else None
    lib = _lib.load()
    if tuning is not None:
        lib.hl_conv_set_tuning(*tuning)
    if tuning2 is not None:
        lib.hl_conv_set_tuning2(*tuning2)
    if split is not None:                                    # split-K: partial-sum workspace on the launching stream
        ws = torch.full((64 << 20,), float("nan"), device=dev)
        _call("hl_conv_set_workspace", ws.data_ptr(), ws.numel() * 4, _stream())
        lib.hl_conv_set_split(split)
    try:
        _call("hl_conv2d", xd.data_ptr(), code, cin_pad, wpk.data_ptr(), bpk.data_ptr(),
              rd.data_ptr() if residual else None, Cout, y.data_ptr(), ldy, st.data_ptr() if stats else None, ldy,
              B, H, W, cin_pad, Cout, k, stride, flags, _stream())
        torch.cuda.synchronize()
    finally:
        lib.hl_conv_set_tuning(-1, -1, -1, -1, -1)
        lib.hl_conv_set_tuning2(-1, -1, -1)
        if split is not None:
            lib.hl_conv_set_split(-1)
            lib.hl_conv_set_workspace(None, 0, _stream())
    return y[..., :Cout].permute(0, 3, 1, 2).float().cpu(), ref, ref_r, (st.cpu().reshape(B, ldy, 2) if stats else None)


def _check_stats(st, y, Cout):
    yy = y.double()                                                    # [B, C, H, W]
    assert rel_l2(st[:, :Cout, 0], yy.sum((2, 3))) < 1e-5, "channel sums from the conv epilogue"
    assert rel_l2(st[:, :Cout, 1], (yy * yy).sum((2, 3))) < 1e-5, "channel sums of squares from the conv epilogue"


@pytest.mark.parametrize("shape", [
    (2, 16, 16, 32, 48, 3, 1), (1, 9, 7, 27, 20, 3, 1), (2, 16, 16, 64, 64, 3, 2), (1, 8, 8, 96, 40, 1, 1),
    (2, 5, 5, 33, 7, 3, 2), (1, 4, 4, 768, 768, 3, 1)])
@pytest.mark.parametrize("mode", ["fp32", "fp16"])
def test_conv_simt_exact(dev, shape, mode):
    from humanliff_b200 import _lib
    B, H, W, Cin, Cout, k, s = shape
    want_stats = Cout % 4 == 0                     # the statistics kernel works on channel quads
    y, ref, ref_r, st = _conv_case(dev, B, H, W, Cin, Cout, k, s, mode=mode, flags=_lib.CONV_FORCE_SIMT,
                                   stats=want_stats)
    assert rel_l2(y, ref_r) < 2e-6 and rel_max(y, ref_r) < 2e-5
    if want_stats:
        _check_stats(st, y, Cout)


def test_conv_simt_upsample_folded(dev):
    from humanliff_b200 import _lib
    y, ref, _, _ = _conv_case(dev, 2, 8, 8, 32, 32, 3, 1, flags=_lib.CONV_FORCE_SIMT | _lib.CONV_UPSAMPLE2X)
    assert rel_l2(y, ref) < 2e-6


TC_SHAPES = [
    # B, H, W, Cin, Cout, k, stride
    (1, 64, 64, 192, 192, 3, 1),     # box (64,2,1)
    (2, 32, 32, 384, 384, 3, 1),     # box (32,4,1), 2+ n-tiles
    (4, 8, 8, 768, 768, 3, 1),       # box (8,8,2): two samples per box
    (1, 128, 128, 64, 192, 3, 1),    # stem-like (Cin padded to one chunk); HALO eligible
    (1, 256, 256, 64, 64, 3, 1),     # W > 128 : 2 boxes per row; HALO eligible
    (2, 128, 128, 192, 192, 3, 1),   # HALO, mh = 2, several tiles per CTA
    (2, 16, 16, 384, 192, 1, 1),     # 1x1 skip conv
    (1, 16, 16, 384, 1152, 1, 1),    # qkv GEMM
    (1, 64, 64, 192, 27, 3, 1),      # out conv: Cout 27 (TMA store clipped at the channel edge)
    (3, 8, 8, 64, 32, 3, 1),         # bn = 2 with B = 3: out-of-range batch rows zero-filled / clipped
    (3, 8, 8, 768, 768, 1, 1),       # 1x1 at 8^2, B odd: statistics of two samples per box from the epilogue, padded sample skipped
    (1, 16, 8, 128, 64, 3, 1),       # non-square
    (2, 64, 64, 192, 192, 3, 2),     # Downsample conv: stride 2 through TMA element strides
    (1, 256, 256, 64, 128, 3, 2),    # stride 2, 256-wide traversal box
]


@pytest.mark.parametrize("shape", TC_SHAPES)
@pytest.mark.parametrize("mode", ["fp16", "tf32"])
def test_conv_tensor_core(dev, shape, mode):
    """tcgen05 kernel (automatic tiling) vs the fp32 reference conv evaluated on the SAME rounded
    operands (only the accumulation order differs -> 2e-5) and vs the exact fp32 conv (operand rounding
    -> 1e-3, the north_star bar); residual add and in-epilogue GroupNorm statistics included."""
    from humanliff_b200 import _lib
    B, H, W, Cin, Cout, k, s = shape
    lib = _lib.load()
    code = 1 if mode == "fp16" else 0
    fl = _lib.CONV_TF32 if mode == "tf32" else 0
    ldy = (Cout + 3) // 4 * 4                      # output pitch: 16-byte rows for the TMA store
    assert lib.hl_conv2d_uses_tensor_cores(code, B, H, W, Cin, Cout, k, s, Cin, ldy, fl) == 1, "must take the tcgen05 path"
    y, ref, ref_r, st = _conv_case(dev, B, H, W, Cin, Cout, k, s, mode=mode, stats=Cout % 4 == 0, seed=7, ldy=ldy,
                                   residual=Cout % 4 == 0)
    assert not torch.isnan(y).any()
    assert rel_l2(y, ref_r) < 2e-5, f"vs fp32 conv on identical (rounded) operands: {rel_l2(y, ref_r)}"
    assert rel_l2(y, ref) < 1e-3, f"vs fp32 reference: {rel_l2(y, ref)}"
    if st is not None:
        _check_stats(st, y, Cout)


@pytest.mark.parametrize("tuning", [(1, -1, 0, -1, -1), (2, -1, 0, -1, -1), (1, -1, 1, -1, -1), (2, 192, 1, -1, -1),
                                    (2, 96, 1, -1, -1), (2, 64, 1, 0, -1), (1, 32, 0, -1, -1)])
@pytest.mark.parametrize("residual", [False, True])
@pytest.mark.parametrize("cta2", [0, 1])
def test_conv_tensor_core_tilings(dev, tuning, residual, cta2):
    """Every tiling variant of the kernel (halves per CTA, N tile, TAP vs HALO operand path, statistics in
    the epilogue or by the separate kernel, single CTA vs CTA-pair MMA) must give the same numbers."""
    B, H, W, Cin, Cout = 2, 128, 128, 192, 192
    y, ref, ref_r, st = _conv_case(dev, B, H, W, Cin, Cout, 3, 1, mode="fp16", stats=True, residual=residual, seed=3,
                                   tuning=tuning, tuning2=(-1, -1, cta2))
    assert not torch.isnan(y).any()
    assert rel_l2(y, ref_r) < 2e-5, (tuning, rel_l2(y, ref_r))
    _check_stats(st, y, Cout)


@pytest.mark.parametrize("shape", [(1, 64, 64, 192, 192, 3, 1), (2, 128, 128, 192, 192, 3, 1), (2, 16, 16, 384, 192, 1, 1),
                                   (1, 16, 16, 384, 1152, 1, 1), (2, 64, 64, 192, 192, 3, 2), (4, 8, 8, 768, 768, 3, 1),
                                   (1, 8, 8, 96, 40, 1, 1)])
@pytest.mark.parametrize("residual", [False, True])
@pytest.mark.parametrize("cta2", [0, 1])
def test_conv_fp16_output(dev, shape, residual, cta2):
    """HL_CONV_OUT_F16: the same fp32 arithmetic rounded once -> bit-identical to the fp32 output's .half()
    (tcgen05 path with / without residual, CTA pairs, 1x1 staging ring; the last shape takes the CUDA-core path)."""
    B, H, W, Cin, Cout, k, s = shape
    kw = dict(mode="fp16", residual=residual, seed=11, tuning2=(-1, -1, cta2))
    y32 = _conv_case(dev, B, H, W, Cin, Cout, k, s, **kw)[0]
    y16, _, _, st = _conv_case(dev, B, H, W, Cin, Cout, k, s, out_f16=True, stats=Cout % 4 == 0, **kw)
    assert not torch.isnan(y16).any()
    assert torch.equal(y16, y32.half().float())
    if st is not None:        # statistics of an fp16 result = those of the ROUNDED values (epilogue, stats kernel, CUDA-core path)
        _check_stats(st, y16, Cout)


@pytest.mark.parametrize("shape,split", [
    ((4, 8, 8, 768, 768, 3, 1), -1),      # automatic: 24 CTAs -> 6 slices
    ((4, 16, 16, 768, 768, 3, 1), -1),    # automatic: 48 CTAs -> 3 slices
    ((4, 16, 16, 1536, 768, 3, 1), -1),
    ((4, 32, 32, 768, 768, 3, 2), -1),    # stride-2 Downsample conv onto 16^2
    ((2, 16, 16, 384, 192, 1, 1), 3),     # forced: 1x1
    ((3, 8, 8, 128, 64, 3, 1), 2),        # bn = 2 with B = 3: the last batch box is half empty (slice padded to 4 samples)
    ((1, 8, 8, 768, 768, 3, 1), -1),      # B = 1 (layer-by-layer generation): automatic split with a padded slice
    ((1, 128, 128, 192, 192, 3, 1), 3),   # forced on the HALO operand path
    ((2, 32, 32, 384, 384, 3, 1), 2)])
@pytest.mark.parametrize("residual", [False, True])
@pytest.mark.parametrize("cta2", [0, 1])
def test_conv_split_k(dev, shape, split, residual, cta2):
    """K cut into slices over more CTAs + fixed-order second pass: same numbers as the one-pass kernel
    (fp32 accumulation order differs -> 2e-5), statistics and fp16 output included."""
    B, H, W, Cin, Cout, k, s = shape
    kw = dict(mode="fp16", residual=residual, seed=13, tuning2=(-1, -1, cta2), split=split)
    y, ref, ref_r, st = _conv_case(dev, B, H, W, Cin, Cout, k, s, stats=True, **kw)
    assert not torch.isnan(y).any()
    assert rel_l2(y, ref_r) < 2e-5, rel_l2(y, ref_r)
    _check_stats(st, y, Cout)
    y2 = _conv_case(dev, B, H, W, Cin, Cout, k, s, stats=True, **kw)[0]
    assert torch.equal(y, y2), "fixed-order reduction: bit-reproducible"
    y16, _, _, st16 = _conv_case(dev, B, H, W, Cin, Cout, k, s, out_f16=True, stats=True, **kw)
    assert torch.equal(y16, y.half().float())
    _check_stats(st16, y16, Cout)
    # the experimental second pass inside the conv kernel (the last CTA of a tile reduces) vs the separate reduction
    # launch: the same slice order, so the same bits
    from humanliff_b200 import _lib
    lib = _lib.load()
    lib.hl_conv_set_split_reduce(1)
    try:
        ys, _, _, sts = _conv_case(dev, B, H, W, Cin, Cout, k, s, stats=True, **kw)
        ys16 = _conv_case(dev, B, H, W, Cin, Cout, k, s, out_f16=True, **kw)[0]
    finally:
        lib.hl_conv_set_split_reduce(0)
    assert torch.equal(y, ys) and torch.equal(y16, ys16)
    assert torch.allclose(st, sts, rtol=2e-5, atol=1e-4)      # fp32 quarter sums (epilogue) vs fp64 (reduction kernel)


@pytest.mark.parametrize("shape", [(2, 64, 64, 192, 192, 1), (4, 32, 32, 384, 384, 1), (4, 8, 8, 768, 768, 1),   # ControlNet projections
                                   (1, 128, 128, 192, 192, 1), (2, 32, 32, 192, 192, 3), (2, 8, 8, 96, 40, 1)])
@pytest.mark.parametrize("split3", [False, True])
def test_conv_dual_output(dev, shape, split3):
    """hl_conv2d_dual: y = conv + residual and y2 = conv from one launch, each with its statistics row -- bit-identical
    to the two hl_conv2d launches it replaces (epilogue statistics, the separate statistics kernel at 8^2, the 3x3
    shape, and the last shape's CUDA-core fallback)."""
    from humanliff_b200.unet import pack_conv
    from humanliff_b200 import _lib
    B, H, W, Cin, Cout, k = shape
    if split3 and Cin % 64:
        pytest.skip("the hi + lo operand passes exist on the tensor-core path only")
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B, H, W, Cin, generator=g) * 3
    w = torch.randn(Cout, Cin, k, k, generator=g) / math.sqrt(Cin * k * k)
    b = torch.randn(Cout, generator=g) * 0.1
    res = torch.randn(B, H, W, Cout, generator=g).to(dev)
    if split3:
        xs = x * 2.0 ** -4
        hi = xs.half()
        xop = torch.cat([hi, (xs - hi.float()).half()], -1).contiguous().to(dev)
        ldx, flags, mode = 2 * Cin, _lib.CONV_SPLIT3, "split"
    else:
        xop, ldx, flags, mode = x.half().to(dev), Cin, 0, None
    wpk, bpk = pack_conv(w, b, Cin, "fp16", dev, mode=mode)
    ldy, ldy2 = Cout + 32, Cout                                    # y: a slice of a wider (concat) buffer

    def run(dual):
        y = torch.zeros(B, H, W, ldy, device=dev)
        y2 = torch.zeros(B, H, W, ldy2, device=dev)
        st = torch.zeros(B * ldy * 2, device=dev, dtype=torch.float64)
        st2 = torch.zeros(B * ldy2 * 2, device=dev, dtype=torch.float64)
        want = Cout % 4 == 0
        if dual:
            _call("hl_conv2d_dual", xop.data_ptr(), 1, ldx, wpk.data_ptr(), bpk.data_ptr(), res.data_ptr(), Cout, y.data_ptr(), ldy,
                  st.data_ptr() if want else None, ldy, y2.data_ptr(), ldy2, st2.data_ptr() if want else None, ldy2, B, H, W, Cin,
                  Cout, k, 1, flags, _stream())
        else:
            _call("hl_conv2d", xop.data_ptr(), 1, ldx, wpk.data_ptr(), bpk.data_ptr(), None, 0, y2.data_ptr(), ldy2,
                  st2.data_ptr() if want else None, ldy2, B, H, W, Cin, Cout, k, 1, flags, _stream())
            _call("hl_conv2d", xop.data_ptr(), 1, ldx, wpk.data_ptr(), bpk.data_ptr(), res.data_ptr(), Cout, y.data_ptr(), ldy,
                  st.data_ptr() if want else None, ldy, B, H, W, Cin, Cout, k, 1, flags, _stream())
        torch.cuda.synchronize()
        return y.cpu(), y2.cpu(), st.cpu().reshape(B, ldy, 2), st2.cpu().reshape(B, ldy2, 2)
    y, y2, st, st2 = run(True)
    ry, ry2, rst, rst2 = run(False)
    assert torch.equal(y, ry) and torch.equal(y2, ry2)
    assert float(y[..., Cout:].abs().max()) == 0.0, "channels beyond Cout of the wider buffer stay untouched"
    assert torch.allclose(st, rst, rtol=1e-12, atol=1e-9) and torch.allclose(st2, rst2, rtol=1e-12, atol=1e-9)
    ref2 = F.conv2d(x.half().float().permute(0, 3, 1, 2), w.half().float(), b, padding=k // 2).permute(0, 2, 3, 1)
    assert rel_l2(y2, ref2) < (2e-5 if not split3 else 1e-3)      # split3 carries MORE bits than the fp16 reference here
    assert rel_l2(y[..., :Cout] - y2, res.cpu()) < 1e-6
    if Cout % 4 == 0:
        _check_stats(st2, y2.permute(0, 3, 1, 2), Cout)
        _check_stats(st, y[..., :Cout].permute(0, 3, 1, 2), Cout)


def test_conv_tc_strided_output_and_input(dev):
    """Operands living inside wider (concat) buffers: ldx > Cin, ldy > Cout, statistics row at an offset."""
    from humanliff_b200.unet import pack_conv
    g = torch.Generator().manual_seed(5)
    B, H, W, Cin, Cout = 1, 32, 32, 64, 64
    big = torch.randn(B, H, W, 192, generator=g).half().to(dev)
    w = torch.randn(Cout, Cin, 3, 3, generator=g) / 24
    b = torch.zeros(Cout)
    wpk, bpk = pack_conv(w, b, Cin, "fp16", dev)
    out = torch.zeros(B, H, W, 96, device=dev)
    st = torch.zeros(B, 96, 2, device=dev, dtype=torch.float64)
    x_off, y_off = 64, 16
    _call("hl_conv2d", big.data_ptr() + 2 * x_off, 1, 192, wpk.data_ptr(), bpk.data_ptr(), None, 0,
          out.data_ptr() + 4 * y_off, 96, st.data_ptr() + 16 * y_off, 96, B, H, W, Cin, Cout, 3, 1, 0, _stream())
    ref = F.conv2d(big[..., x_off:x_off + Cin].float().permute(0, 3, 1, 2).cpu(), w.half().float(), b, padding=1)
    got = out[..., y_off:y_off + Cout].permute(0, 3, 1, 2).cpu()
    assert rel_l2(got, ref) < 2e-5
    assert float(out[..., :y_off].abs().max()) == 0 and float(out[..., y_off + Cout:].abs().max()) == 0
    assert rel_l2(st[:, y_off:y_off + Cout, 0].cpu(), got.double().sum((2, 3))) < 1e-5
    assert float(st[:, :y_off].abs().max()) == 0 and float(st[:, y_off + Cout:].abs().max()) == 0


@pytest.mark.parametrize("B,T,C,heads", [(2, 64, 384, 4), (1, 1024, 384, 4), (2, 256, 768, 4), (1, 16, 128, 2), (3, 4, 64, 2), (1, 100, 256, 4)])
def test_attention(dev, B, T, C, heads):
    g = torch.Generator().manual_seed(T + C)
    qkv = torch.randn(B, 3 * C, T, generator=g)
    ch = C // heads
    r = qkv.reshape(B * heads, 3 * ch, T)
    q, k, v = torch.split(r, ch, dim=1)
    s = 1 / math.sqrt(math.sqrt(ch))
    w = torch.softmax(torch.einsum("bct,bcs->bts", q * s, k * s), -1)
    ref = torch.einsum("bts,bcs->bct", w, v).reshape(B, C, T)
    qd = qkv.permute(0, 2, 1).contiguous().to(dev)             # [B, T, 3C]
    out = torch.empty(B, T, C, device=dev)
    _call("hl_attention", qd.data_ptr(), 0, 3 * C, out.data_ptr(), 0, C, B, T, C, heads, 0, _stream())
    assert rel_l2(out.permute(0, 2, 1), ref) < 5e-6
    assert rel_max(out.permute(0, 2, 1), ref) < 5e-5
    outh = torch.empty(B, T, C, device=dev, dtype=torch.float16)
    _call("hl_attention", qd.data_ptr(), 0, 3 * C, outh.data_ptr(), 1, C, B, T, C, heads, 2, _stream())   # CUDA cores
    assert torch.equal(outh.cpu(), out.cpu().half())
    # tensor-core kernel (mma.sync, fp16 operands / fp32 accumulate): vs the same attention evaluated in fp32 on
    # the fp16-ROUNDED q, k, v (isolates the kernel from the operand rounding), and vs the exact result
    outm = torch.full((B, T, C), float("nan"), device=dev, dtype=torch.float16)
    _call("hl_attention", qd.data_ptr(), 0, 3 * C, outm.data_ptr(), 1, C, B, T, C, heads, 0, _stream())
    rh = qkv.half().float().reshape(B * heads, 3 * ch, T)
    qh, kh, vh = torch.split(rh, ch, dim=1)
    wh = torch.softmax(torch.einsum("bct,bcs->bts", qh, kh) / math.sqrt(ch), -1)
    refh = torch.einsum("bts,bcs->bct", wh, vh).reshape(B, C, T)
    assert not torch.isnan(outm.float()).any()
    assert rel_l2(outm.float().permute(0, 2, 1), refh) < 6e-4      # P and the output rounded to fp16
    assert rel_l2(outm.float().permute(0, 2, 1), ref) < 2e-3
    # fp16 qkv (as written by a HL_CONV_OUT_F16 conv): cp.async double-buffered mma.sync variant, same arithmetic
    # (flag 4 keeps the call on the mma.sync kernel where the tcgen05 kernel would serve the shape)
    qh16 = qd.half()
    outf = torch.full((B, T, C), float("nan"), device=dev, dtype=torch.float16)
    _call("hl_attention", qh16.data_ptr(), 1, 3 * C, outf.data_ptr(), 1, C, B, T, C, heads, 4, _stream())
    assert torch.equal(outf.cpu(), outm.cpu())


@pytest.mark.parametrize("B,T,C,heads", [(4, 1024, 384, 4), (4, 256, 768, 4), (4, 64, 768, 4),      # the production blocks
                                         (2, 128, 256, 4), (1, 192, 512, 4), (3, 64, 192, 2)])
def test_attention_tcgen05(dev, B, T, C, heads):
    """The tcgen05 / TMEM / TMA attention kernel (fp16 qkv and output, T % 64 == 0): against fp32 attention on the
    fp16-rounded q, k, v (isolates the kernel from the operand rounding), against the exact result, and against the
    mma.sync kernel on the same operands."""
    g = torch.Generator().manual_seed(T + C)
    qkv = torch.randn(B, 3 * C, T, generator=g)
    ch = C // heads
    r = qkv.reshape(B * heads, 3 * ch, T)
    q, k, v = torch.split(r, ch, dim=1)
    s = 1 / math.sqrt(math.sqrt(ch))
    w = torch.softmax(torch.einsum("bct,bcs->bts", q * s, k * s), -1)
    ref = torch.einsum("bts,bcs->bct", w, v).reshape(B, C, T)
    rh = qkv.half().float().reshape(B * heads, 3 * ch, T)
    qh, kh, vh = torch.split(rh, ch, dim=1)
    wh = torch.softmax(torch.einsum("bct,bcs->bts", qh, kh) / math.sqrt(ch), -1)
    refh = torch.einsum("bts,bcs->bct", wh, vh).reshape(B, C, T)
    q16 = qkv.permute(0, 2, 1).contiguous().to(dev).half()               # [B, T, 3C] head-major channels
    out5 = torch.full((B, T, C), float("nan"), device=dev, dtype=torch.float16)
    outm = torch.full((B, T, C), float("nan"), device=dev, dtype=torch.float16)
    _call("hl_attention", q16.data_ptr(), 1, 3 * C, out5.data_ptr(), 1, C, B, T, C, heads, 0, _stream())
    _call("hl_attention", q16.data_ptr(), 1, 3 * C, outm.data_ptr(), 1, C, B, T, C, heads, 4, _stream())
    torch.cuda.synchronize()
    assert not torch.isnan(out5.float()).any()
    o5, om = out5.float().permute(0, 2, 1), outm.float().permute(0, 2, 1)
    assert rel_l2(o5, refh) < 6e-4, rel_l2(o5, refh)
    assert rel_l2(o5, ref) < 2e-3
    assert rel_l2(o5, om) < 6e-4, rel_l2(o5, om)


def test_ddpm_step_bit_exact(dev):
    from humanliff_b200 import create_gaussian_diffusion
    from oracle.diffusion_oracle import DiffusionOracle
    g = torch.Generator().manual_seed(9)
    B, shape = 4, (4, 27, 16, 16)
    x, eps, z = (torch.randn(shape, generator=g) for _ in range(3))
    for resp in ("250", ""):
        d = create_gaussian_diffusion(steps=1000, timestep_respacing=resp)
        o = DiffusionOracle(1000, resp)
        t = torch.tensor([0, 1, d.num_timesteps // 2, d.num_timesteps - 1])
        sample, x0 = d._fused_step(x.to(dev), eps.to(dev), z.to(dev), t.to(dev), True)
        rs, r0 = o.posterior(x, eps, t, z)
        assert torch.equal(x0.cpu(), r0)
        assert torch.equal(sample.cpu(), rs)


def test_ddim_step_bit_exact(dev):
    """hl_ddim_step vs the oracle's ddim_posterior (same fp32 operation order): bit-exact, eta = 0 and 0.5."""
    from humanliff_b200 import create_gaussian_diffusion
    from humanliff_b200._lib import call
    from oracle.diffusion_oracle import DiffusionOracle
    g = torch.Generator().manual_seed(10)
    shape = (4, 27, 16, 16)
    x, eps, z = (torch.randn(shape, generator=g) for _ in range(3))
    xd, ed, zd = x.to(dev), eps.to(dev), z.to(dev)             # keep the device copies alive across the call
    for resp in ("250", ""):
        d = create_gaussian_diffusion(steps=1000, timestep_respacing=resp)
        o = DiffusionOracle(1000, resp)
        t = torch.tensor([0, 1, d.num_timesteps // 2, d.num_timesteps - 1])
        td = t.to(dev)
        for eta in (0.0, 0.5):
            tb = d._ddim_tables(dev, eta)
            sample, x0 = torch.empty(shape, device=dev), torch.empty(shape, device=dev)
            call("hl_ddim_step", xd.data_ptr(), ed.data_ptr(), zd.data_ptr(), tb["coef"].data_ptr(),
                 tb["sigma"].data_ptr(), td.data_ptr(), sample.data_ptr(), x0.data_ptr(), 4, x[0].numel(), 1,
                 _stream())
            rs, r0 = o.ddim_posterior(x, eps, t, z, eta=eta)
            assert torch.equal(x0.cpu(), r0), (resp, eta)
            assert torch.equal(sample.cpu(), rs), (resp, eta)


# ---------------------------------------------------------------------------------------------------------------
# High-precision operand passes of the fp16 plan: raw residual-stream operands as scaled fp16 hi | lo
# pairs, three tensor-core passes.  Reference = float64 exact products of the fp32 inputs; the bar is the error of a
# ~22-bit operand (1e-5), two orders below a plain fp16 operand (4e-4), and NO overflow at |x| = 1e5.
# ---------------------------------------------------------------------------------------------------------------
def _split_conv_case(dev, B, H, W, Cin, Cout, k, stride, scale, residual=False, out_split=False, force_simt=False, seed=0):
    from humanliff_b200.unet import pack_conv
    from humanliff_b200 import _lib
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, H, W, Cin, generator=g) * scale                      # NHWC raw stream, |x| up to ~5 * scale
    w = torch.randn(Cout, Cin, k, k, generator=g) / math.sqrt(Cin * k * k)
    b = torch.randn(Cout, generator=g) * 0.1
    ref = F.conv2d(x.permute(0, 3, 1, 2).double(), w.double(), b.double(), stride=stride, padding=k // 2)
    res = None
    if residual:
        res = torch.randn(B, ref.shape[2], ref.shape[3], Cout, generator=g) * scale
        ref = ref + res.permute(0, 3, 1, 2).double()
    xd = x.to(dev)
    op = torch.full((B, H, W, 2 * Cin), float("nan"), device=dev, dtype=torch.float16)
    mode = _lib.OP_SPLIT | _lib.OP_SCALED | (Cin << 8)
    _call("hl_cast_operand", xd.data_ptr(), Cin, op.data_ptr(), 1, 2 * Cin, Cin, B * H * W, mode, _stream())
    # the operand pair reconstructs x * 2^-4 to ~2^-22 relative
    rec = (op[..., :Cin].double() + op[..., Cin:].double()).cpu() * 16.0
    assert rel_l2(rec, x) < 2e-6 and torch.isfinite(op).all()
    wpk, bpk = pack_conv(w, b, Cin, "fp16", dev, mode="split")
    assert wpk.shape[0] == 2 * k * k
    Ho, Wo = ref.shape[2:]
    flags = _lib.CONV_SPLIT3 | (_lib.CONV_FORCE_SIMT if force_simt else 0)
    if out_split:
        y = torch.full((B, Ho, Wo, 2 * Cout), float("nan"), device=dev, dtype=torch.float16)
        flags |= _lib.CONV_OUT_F16_SPLIT
        ldy = 2 * Cout
    else:
        y = torch.full((B, Ho, Wo, Cout), float("nan"), device=dev)
        ldy = Cout
    rd = res.to(dev) if residual else None
    _call("hl_conv2d", op.data_ptr(), 1, 2 * Cin, wpk.data_ptr(), bpk.data_ptr(), rd.data_ptr() if residual else None, Cout,
          y.data_ptr(), ldy, None, 0, B, H, W, Cin, Cout, k, stride, flags, _stream())
    torch.cuda.synchronize()
    if out_split:
        out = (y[..., :Cout].double() + y[..., Cout:].double()) * 16.0
    else:
        out = y.double()
    return out.permute(0, 3, 1, 2).cpu(), ref, y


@pytest.mark.parametrize("shape", [(2, 32, 32, 192, 384, 1, 1),      # 1x1 skip / ControlNet projection
                                   (2, 64, 64, 192, 192, 3, 2),      # Downsample
                                   (4, 8, 8, 768, 768, 3, 2),        # Downsample at 8^2 -> 4^2: split-K territory
                                   (1, 128, 128, 192, 192, 1, 1)])
@pytest.mark.parametrize("scale", [1.0, 2e4])
@pytest.mark.parametrize("residual", [False, True])
def test_conv_split3_high_precision_and_range(dev, shape, scale, residual):
    from humanliff_b200 import _lib
    B, H, W, Cin, Cout, k, s = shape
    assert _lib.load().hl_conv2d_uses_tensor_cores(1, B, H, W, Cin, Cout, k, s, 2 * Cin, Cout, _lib.CONV_SPLIT3)
    out, ref, _ = _split_conv_case(dev, B, H, W, Cin, Cout, k, s, scale, residual=residual)
    assert torch.isfinite(out).all()                     # |x| reaches 1e5 at scale 2e4: a plain fp16 operand is inf there
    e = rel_l2(out, ref)
    # K = 3 x 9 x 768 = 20,736 fp32 accumulations (and a split-K second pass) put the 768-channel Downsample at 2e-5
    assert e < (4e-5 if Cin >= 768 else 1e-5), f"rel-L2 {e:.2e} (a plain fp16 operand gives 4e-4)"


def test_conv_split3_simt_fallback_and_split_output(dev):
    """Shapes the tcgen05 kernel does not tile run the same passes on the CUDA cores; the split fp16 output
    (HL_CONV_OUT_F16_SPLIT) of one conv is the split operand of the next (ControlNet block -> projection)."""
    from humanliff_b200.unet import pack_conv
    from humanliff_b200 import _lib
    out, ref, _ = _split_conv_case(dev, 1, 12, 12, 64, 96, 1, 1, 3e4, force_simt=True)
    assert rel_l2(out, ref) < 1e-5
    for simt in (False, True):
        out, ref, y = _split_conv_case(dev, 2, 32, 32, 192, 192, 3, 1, 1e4, residual=True, out_split=True, force_simt=simt)
        assert torch.isfinite(y).all() and rel_l2(out, ref) < 1e-5, simt
        # ... and feeds a projection conv directly
        g = torch.Generator().manual_seed(9)
        w2 = torch.randn(192, 192, 1, 1, generator=g) / math.sqrt(192)
        b2 = torch.zeros(192)
        wpk, bpk = pack_conv(w2, b2, 192, "fp16", dev, mode="split")
        z = torch.empty(2, 32, 32, 192, device=dev)
        _call("hl_conv2d", y.data_ptr(), 1, 384, wpk.data_ptr(), bpk.data_ptr(), None, 0, z.data_ptr(), 192, None, 0,
              2, 32, 32, 192, 192, 1, 1, _lib.CONV_SPLIT3, _stream())
        ref2 = F.conv2d(ref, w2.double())
        assert rel_l2(z.permute(0, 3, 1, 2), ref2) < 2e-5


def test_stem_split_packed(dev):
    """Stem conv (27 -> 192, one 64-channel K chunk): the operand row carries [hi(27) 0.. | lo(27) 0..], two passes."""
    from humanliff_b200.unet import pack_conv
    from humanliff_b200 import _lib
    g = torch.Generator().manual_seed(2)
    B, C, H, W, Cout = 2, 27, 64, 64, 192
    x, x2 = torch.randn(B, C, H, W, generator=g), 0.3 * torch.randn(B, C, H, W, generator=g)
    w = torch.randn(Cout, C, 3, 3, generator=g) / math.sqrt(C * 9)
    b = torch.randn(Cout, generator=g) * 0.1
    op = torch.full((B, H, W, 64), float("nan"), device=dev, dtype=torch.float16)
    xd, x2d = x.to(dev), x2.to(dev)                      # keep the device tensors alive across the launch
    _call("hl_nchw_to_nhwc", xd.data_ptr(), x2d.data_ptr(), op.data_ptr(), 1, B, C, H * W, 64,
          _lib.OP_SPLIT | (32 << 8), _stream())
    torch.cuda.synchronize()
    assert torch.isfinite(op).all() and float(op[..., 27:32].abs().max()) == 0 and float(op[..., 59:].abs().max()) == 0
    wpk, bpk = pack_conv(w, b, 64, "fp16", dev, mode="split_packed")
    y = torch.empty(B, H, W, Cout, device=dev)
    _call("hl_conv2d", op.data_ptr(), 1, 64, wpk.data_ptr(), bpk.data_ptr(), None, 0, y.data_ptr(), Cout, None, 0,
          B, H, W, 64, Cout, 3, 1, _lib.CONV_SPLIT2P, _stream())
    ref = F.conv2d((x + x2).double(), w.double(), b.double(), padding=1)
    assert rel_l2(y.permute(0, 3, 1, 2), ref) < 1e-5


def test_gn_apply_split_outputs(dev):
    """hl_gn_apply op-mode word: a split (unscaled) normalised output for the out conv and a split scaled raw copy
    for the 1x1 skip conv."""
    from humanliff_b200 import _lib
    g = torch.Generator().manual_seed(4)
    B, HW, C = 2, 1024, 192
    x = torch.randn(B, HW, C, generator=g) * 3e4
    gamma, beta = 1 + 0.1 * torch.randn(C, generator=g), 0.1 * torch.randn(C, generator=g)
    xd = x.to(dev)
    stats = torch.zeros(B, C, 2, dtype=torch.float64, device=dev)
    _call("hl_gn_stats", xd.data_ptr(), C, B, HW, C, stats.data_ptr(), C, _stream())
    y = torch.full((B, HW, 2 * C), float("nan"), device=dev, dtype=torch.float16)
    raw = torch.full((B, HW, 2 * C), float("nan"), device=dev, dtype=torch.float16)
    mode = _lib.OP_SPLIT | (C << 8) | ((_lib.OP_SPLIT | _lib.OP_SCALED) << _lib.OP_RAW_SHIFT)
    gd, bd = gamma.to(dev), beta.to(dev)                 # keep the device tensors alive across the launch
    _call("hl_gn_apply", xd.data_ptr(), C, stats.data_ptr(), C, gd.data_ptr(), bd.data_ptr(), None, 0,
          y.data_ptr(), 1, 2 * C, raw.data_ptr(), 2 * C, B, HW, C, 32, 1e-5, 1, mode, _stream())
    torch.cuda.synchronize()
    ref = F.silu(F.group_norm(x.permute(0, 2, 1).double(), 32, gamma.double(), beta.double(), eps=1e-5)).permute(0, 2, 1)
    yy = (y[..., :C].double() + y[..., C:].double()).cpu()
    rr = (raw[..., :C].double() + raw[..., C:].double()).cpu() * 16.0
    assert torch.isfinite(y).all() and torch.isfinite(raw).all()
    assert rel_l2(yy, ref) < 5e-6 and rel_l2(rr, x) < 2e-6
