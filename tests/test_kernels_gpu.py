"""GPU: every C-ABI kernel against a plain torch-CPU statement of the same op (the oracle's arithmetic
library) on seeded inputs.  Tolerances: 1e-5 (fp32 kernels, accumulation-order differences only),
1e-3 rel-L2 for the TF32 tensor-core convolution (north_star bar), bit-exact for the DDPM update."""
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
        _call("hl_nchw_to_nhwc", xd.data_ptr(), x2d.data_ptr(), nhwc.data_ptr(), B, C, H * W, ld, 0, _stream())
        ref = (x + x2).permute(0, 2, 3, 1).reshape(B, H * W, C)
        assert torch.equal(nhwc[:, :, :C].cpu(), ref)
        assert float(nhwc[:, :, C:].abs().max() if ld > C else 0) == 0.0, "channel padding must be zero"
        back = torch.empty(B, C, H, W, device=dev)
        _call("hl_nhwc_to_nchw", nhwc.data_ptr(), ld, back.data_ptr(), B, C, H * W, _stream())
        assert torch.equal(back.cpu(), x + x2)


def test_round_tf32_matches_emulation(dev):
    from oracle.unet_oracle import round_tf32
    x = torch.randn(1000, 8) * torch.logspace(-6, 6, 1000)[:, None]
    xd = x.to(dev)
    out = torch.empty_like(xd)
    _call("hl_round_tf32", xd.data_ptr(), 8, out.data_ptr(), 8, 8, 1000, _stream())
    assert torch.equal(out.cpu(), round_tf32(x))


def test_concat_add_and_upsample(dev):
    g = torch.Generator().manual_seed(1)
    npix, C1, C2 = 50, 8, 12
    a, b, c = torch.randn(npix, C1, generator=g), torch.randn(npix, C2, generator=g), torch.randn(npix, C2, generator=g)
    ad, bd, cd = a.to(dev), b.to(dev), c.to(dev)
    out = torch.empty(npix, C1 + C2, device=dev)
    _call("hl_concat_add", ad.data_ptr(), C1, C1, bd.data_ptr(), C2, cd.data_ptr(), C2, C2, out.data_ptr(),
          C1 + C2, npix, _stream())
    assert torch.equal(out.cpu(), torch.cat([a, b + c], 1))
    _call("hl_concat_add", ad.data_ptr(), C1, C1, bd.data_ptr(), C2, None, C2, C2, out.data_ptr(),
          C1 + C2, npix, _stream())
    assert torch.equal(out.cpu(), torch.cat([a, b], 1))
    x = torch.randn(2, 3, 5, 8, generator=g)                       # NHWC [B,H,W,C]
    up = torch.empty(2, 6, 10, 8, device=dev)
    xd = x.to(dev)
    _call("hl_upsample2x", xd.data_ptr(), 8, up.data_ptr(), 8, 2, 3, 5, 8, 0, _stream())
    ref = F.interpolate(x.permute(0, 3, 1, 2), scale_factor=2, mode="nearest").permute(0, 2, 3, 1)
    assert torch.equal(up.cpu(), ref)


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
    sums = torch.empty(B * 32 * 2, device=dev, dtype=torch.float64)
    y = torch.empty_like(xd)
    _call("hl_gn_stats", xd.data_ptr(), C, B, HW, C, 32, sums.data_ptr(), _stream())
    xn = F.group_norm(x.permute(0, 2, 1), 32, gamma, beta, eps=1e-5)          # [B, C, HW]
    for use_film, silu in [(False, True), (True, True), (False, False)]:
        _call("hl_gn_apply", xd.data_ptr(), C, sums.data_ptr(), gd.data_ptr(), bd.data_ptr(),
              fd.data_ptr() if use_film else None, 2 * C + 10, y.data_ptr(), C, B, HW, C, 32, 1e-5,
              1 if silu else 0, 0, _stream())
        r = xn
        if use_film:
            r = r * (1 + film[:, :C, None]) + film[:, C:2 * C, None]
        if silu:
            r = r * torch.sigmoid(r)
        assert rel_l2(y.permute(0, 2, 1), r) < 2e-6, (use_film, silu)
        assert rel_max(y.permute(0, 2, 1), r) < 2e-5


def _conv_case(dev, B, H, W, Cin, Cout, k, stride, flags=0, residual=True, tf32=False, cin_pad=None, seed=0):
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
    if residual:
        ref = ref + res
    xn = torch.zeros(B, H, W, cin_pad)
    xn[..., :Cin] = x.permute(0, 2, 3, 1)
    xd = xn.to(dev)
    if tf32:
        _call("hl_round_tf32", xd.data_ptr(), cin_pad, xd.data_ptr(), cin_pad, cin_pad, B * H * W, _stream())
    wpk, bpk = pack_conv(w, b, cin_pad, tf32, dev)
    y = torch.full((B, Ho, Wo, Cout), float("nan"), device=dev)
    rd = res.permute(0, 2, 3, 1).contiguous().to(dev) if residual else None
    _call("hl_conv2d", xd.data_ptr(), cin_pad, wpk.data_ptr(), bpk.data_ptr(), rd.data_ptr() if residual else None,
          Cout, y.data_ptr(), Cout, B, H, W, cin_pad, Cout, k, stride, flags, _stream())
    torch.cuda.synchronize()
    return y.permute(0, 3, 1, 2).cpu(), ref, (xd, wpk, bpk, rd, (Ho, Wo))


@pytest.mark.parametrize("shape", [
    (2, 16, 16, 32, 48, 3, 1), (1, 9, 7, 27, 20, 3, 1), (2, 16, 16, 64, 64, 3, 2), (1, 8, 8, 96, 40, 1, 1),
    (2, 5, 5, 33, 7, 3, 2), (1, 4, 4, 768, 768, 3, 1)])
def test_conv_simt_exact(dev, shape):
    from humanliff_b200 import _lib
    B, H, W, Cin, Cout, k, s = shape
    y, ref, _ = _conv_case(dev, B, H, W, Cin, Cout, k, s, flags=_lib.CONV_FORCE_SIMT)
    assert rel_l2(y, ref) < 2e-6 and rel_max(y, ref) < 2e-5


def test_conv_simt_upsample_folded(dev):
    from humanliff_b200 import _lib
    y, ref, _ = _conv_case(dev, 2, 8, 8, 32, 32, 3, 1, flags=_lib.CONV_FORCE_SIMT | _lib.CONV_UPSAMPLE2X)
    assert rel_l2(y, ref) < 2e-6


TC_SHAPES = [
    # B, H, W, Cin, Cout, k   -- tile shapes: (bw,bh,bn)
    (1, 64, 64, 192, 192, 3),     # (64,2,1)  N=192
    (2, 32, 32, 384, 384, 3),     # (32,4,1)  N=192 x2
    (4, 8, 8, 768, 768, 3),       # (8,8,2)   N=256 x3
    (1, 128, 128, 32, 192, 3),    # (128,1,1) stem: Cin padded 27->32
    (1, 256, 256, 64, 64, 3),     # W > 128 : 2 tiles per row
    (2, 16, 16, 384, 192, 1),     # 1x1 skip conv
    (1, 16, 16, 384, 1152, 1),    # qkv GEMM
    (1, 64, 64, 192, 27, 3),      # out conv: Cout 27 (padded N tile, scalar epilogue)
    (3, 8, 8, 64, 32, 3),         # bn=2 with B=3: out-of-range batch rows are zero-filled + masked
    (1, 16, 8, 96, 64, 3),        # non-square
]


@pytest.mark.parametrize("shape", TC_SHAPES)
def test_conv_tensor_core_vs_simt_and_fp32(dev, shape):
    """tcgen05 kernel vs (a) the fp32 CUDA-core kernel on the SAME TF32-rounded operands (only the
    accumulation order differs -> 5e-5) and (b) the fp32 reference conv (TF32 operand error -> 1e-3)."""
    from humanliff_b200 import _lib
    B, H, W, Cin, Cout, k = shape
    lib = _lib.load()
    assert lib.hl_conv2d_uses_tensor_cores(B, H, W, Cin, Cout, k, 1, Cin, 0) == 1, "shape must take the tcgen05 path"
    y, ref, (xd, wpk, bpk, rd, (Ho, Wo)) = _conv_case(dev, B, H, W, Cin, Cout, k, 1, tf32=True, seed=7)
    y2 = torch.full((B, Ho, Wo, Cout), float("nan"), device=dev)
    _call("hl_conv2d", xd.data_ptr(), Cin, wpk.data_ptr(), bpk.data_ptr(), rd.data_ptr(), Cout, y2.data_ptr(),
          Cout, B, H, W, Cin, Cout, k, 1, _lib.CONV_FORCE_SIMT, _stream())
    y2 = y2.permute(0, 3, 1, 2).cpu()
    assert not torch.isnan(y).any()
    assert rel_l2(y, y2) < 5e-5, f"tcgen05 vs fp32-core on identical operands: {rel_l2(y, y2)}"
    assert rel_l2(y, ref) < 1e-3, f"vs fp32 reference: {rel_l2(y, ref)}"


def test_conv_tc_strided_output_and_input(dev):
    """Operands living inside wider (concat) buffers: ldx > Cin, ldy > Cout."""
    from humanliff_b200.unet import pack_conv
    g = torch.Generator().manual_seed(5)
    B, H, W, Cin, Cout = 1, 32, 32, 64, 64
    big = torch.randn(B, H, W, 160, generator=g).to(dev)
    w = torch.randn(Cout, Cin, 3, 3, generator=g) / 24
    b = torch.zeros(Cout)
    wpk, bpk = pack_conv(w, b, Cin, False, dev)
    out = torch.zeros(B, H, W, 96, device=dev)
    x_off, y_off = 32, 16
    _call("hl_conv2d", big.data_ptr() + 4 * x_off, 160, wpk.data_ptr(), bpk.data_ptr(), None, 0,
          out.data_ptr() + 4 * y_off, 96, B, H, W, Cin, Cout, 3, 1, 0, _stream())
    ref = F.conv2d(big[..., x_off:x_off + Cin].permute(0, 3, 1, 2).cpu(), w, b, padding=1)
    assert rel_l2(out[..., y_off:y_off + Cout].permute(0, 3, 1, 2), ref) < 1e-3
    assert float(out[..., :y_off].abs().max()) == 0 and float(out[..., y_off + Cout:].abs().max()) == 0


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
    _call("hl_attention", qd.data_ptr(), 3 * C, out.data_ptr(), C, B, T, C, heads, 0, _stream())
    assert rel_l2(out.permute(0, 2, 1), ref) < 5e-6
    assert rel_max(out.permute(0, 2, 1), ref) < 5e-5


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
