"""CPU interpreter of a ``_StepPlan`` launch list (TEST INFRASTRUCTURE).

``humanliff_b200.unet._StepPlan`` compiles one UNet forward into a flat list of C-ABI calls over a
fixed workspace.  Built on the CPU device, the very same list can be *interpreted* here with
numpy/torch restatements of each entry point's documented semantics (include/humanliff_b200.h), which
checks the whole host-side dataflow -- pointers, pitches, concat slices, statistics rows, FiLM offsets --
against the oracle without a GPU.  The kernels themselves are checked on the GPU (tests/*_gpu.py)."""
import ctypes

import numpy as np
import torch
import torch.nn.functional as F

_CT = {np.float32: ctypes.c_float, np.float16: ctypes.c_uint16, np.float64: ctypes.c_double, np.int64: ctypes.c_int64}


def view(ptr, n, dtype=np.float32):
    """numpy view of n elements of `dtype` at raw address `ptr`."""
    ct = _CT[dtype]
    arr = np.ctypeslib.as_array((ct * int(n)).from_address(int(ptr)))
    return arr.view(dtype) if dtype == np.float16 else arr


def pitched(ptr, rows, cols, ld, dtype=np.float32):
    """[rows, cols] view with row pitch ld (elements)."""
    flat = view(ptr, (rows - 1) * ld + cols, dtype)
    return np.lib.stride_tricks.as_strided(flat, shape=(rows, cols), strides=(ld * flat.itemsize, flat.itemsize))


def _dt(code):
    return np.float16 if code == 1 else np.float32


def _round_tf32(a):
    i = a.astype(np.float32).view(np.int32)
    return ((i + 0x1000) & ~0x1FFF).view(np.float32)


OP_TF32, OP_SCALED, OP_SPLIT = 1, 2, 4          # HL_OP_* of include/humanliff_b200.h


def _store_operand(dst_ptr, dtype_code, rows, cols, ld, values, mode):
    """`mode` = the HL_OP_* word: TF32 rounding (fp32), 2^-4 scaling and hi | lo split with lo at + (mode >> 8) (fp16)."""
    if dtype_code == 1:
        v = values.astype(np.float32) * (np.float32(2.0 ** -4) if mode & OP_SCALED else np.float32(1.0))
        hi = v.astype(np.float16)
        pitched(dst_ptr, rows, cols, ld, np.float16)[...] = hi
        if mode & OP_SPLIT:
            lo_off = mode >> 8
            lo = (v - hi.astype(np.float32)).astype(np.float16)
            pitched(dst_ptr + 2 * lo_off, rows, cols, ld, np.float16)[...] = lo
    else:
        out = pitched(dst_ptr, rows, cols, ld, np.float32)
        out[...] = _round_tf32(values) if mode & OP_TF32 else values


def hl_zero(ptr, nbytes, stream):
    view(ptr, nbytes // 8, np.float64)[...] = 0


def hl_timestep_embedding(t, freqs, B, dim, out, stream):
    half = dim // 2
    tt, ff = view(t, B), view(freqs, half)
    a = (tt[:, None] * ff[None]).astype(np.float32)
    o = view(out, B * dim).reshape(B, dim)
    o[:, :half] = np.cos(a)
    o[:, half:2 * half] = np.sin(a)


def hl_linear_small(x, W, bias, y, B, in_f, out_f, silu_in, add_table, add_idx, stream):
    xv = torch.from_numpy(view(x, B * in_f).reshape(B, in_f).copy())
    if silu_in:
        xv = xv * torch.sigmoid(xv)
    Wv = torch.from_numpy(view(W, out_f * in_f).reshape(out_f, in_f))
    r = xv @ Wv.T
    if bias:
        r = r + torch.from_numpy(view(bias, out_f))
    if add_table:
        idx = view(add_idx, B, np.int64)
        r = r + torch.from_numpy(view(add_table, (int(idx.max()) + 1) * out_f).reshape(-1, out_f))[idx]
    view(y, B * out_f).reshape(B, out_f)[...] = r.numpy()


def hl_nchw_to_nhwc(src, src2, dst, dst_dtype, B, C, HW, ld, round_tf32, stream):
    v = view(src, B * C * HW).reshape(B, C, HW).copy()
    if src2:
        v = v + view(src2, B * C * HW).reshape(B, C, HW)
    full = np.zeros((B * HW, ld), np.float32)
    full[:, :C] = v.transpose(0, 2, 1).reshape(B * HW, C)
    if dst_dtype == 1 and round_tf32 & OP_SPLIT:          # packed pair inside the row: [hi(C) 0.. | lo(C) 0..]
        lo_off = round_tf32 >> 8
        hi = full.astype(np.float16)
        full[:, lo_off:lo_off + C] = (full[:, :C] - hi[:, :C].astype(np.float32))
        full[:, :C] = hi[:, :C].astype(np.float32)
        round_tf32 = 0
    _store_operand(dst, dst_dtype, B * HW, ld, ld, full, round_tf32)


def hl_nhwc_to_nchw(src, ld, dst, B, C, HW, stream):
    v = pitched(src, B * HW, C, ld)
    view(dst, B * C * HW).reshape(B, C, HW)[...] = v.reshape(B, HW, C).transpose(0, 2, 1)


def hl_nhwc_to_nchw_sum2(src, ld, off2, dst, B, C, HW, stream):
    v = pitched(src, B * HW, C, ld) + pitched(src + 4 * off2, B * HW, C, ld)
    view(dst, B * C * HW).reshape(B, C, HW)[...] = v.reshape(B, HW, C).transpose(0, 2, 1)


def hl_cast_operand(src, lds, dst, dst_dtype, ldd, C, npix, round_tf32, stream):
    _store_operand(dst, dst_dtype, npix, C, ldd, pitched(src, npix, C, lds).copy(), round_tf32)


def hl_upsample2x(src, lds, dst, dst_dtype, ldd, B, H, W, C, round_tf32, stream):
    v = pitched(src, B * H * W, C, lds).reshape(B, H, W, C)
    up = v.repeat(2, axis=1).repeat(2, axis=2).reshape(B * 4 * H * W, C)
    _store_operand(dst, dst_dtype, B * 4 * H * W, C, ldd, up, round_tf32)


def hl_gn_stats(x, ldx, B, HW, C, stats, stats_ld, stream, dtype=np.float32):
    v = pitched(x, B * HW, C, ldx, dtype).reshape(B, HW, C).astype(np.float64)
    st = pitched(stats, B, 2 * C, 2 * stats_ld, np.float64).reshape(B, C, 2)
    st[:, :, 0] += v.sum(1)
    st[:, :, 1] += (v * v).sum(1)


def hl_gn_apply(x, ldx, stats, stats_ld, gamma, beta, film, film_ld, y, y_dtype, ldy, raw, ldraw, B, HW, C, groups,
                eps, silu, round_tf32, stream):
    x_f16 = bool(round_tf32 & 8)                      # HL_OP_X_F16: the input is an fp16 tensor
    assert not (x_f16 and raw)
    v = pitched(x, B * HW, C, ldx, np.float16 if x_f16 else np.float32).reshape(B, HW, C).astype(np.float32)
    st = pitched(stats, B, 2 * C, 2 * stats_ld, np.float64).reshape(B, groups, C // groups, 2).sum(2)
    n = HW * (C // groups)
    mean = st[..., 0] / n
    var = np.maximum(st[..., 1] / n - mean * mean, 0.0)
    rstd = (1.0 / np.sqrt(var + eps)).astype(np.float32)
    mean = mean.astype(np.float32)
    cpg = C // groups
    ga = view(gamma, C)[None] * np.repeat(rstd, cpg, axis=1)
    be = view(beta, C)[None] - np.repeat(mean, cpg, axis=1) * ga
    if film:
        f = pitched(film, B, 2 * C, film_ld)
        sc, sh = 1.0 + f[:, :C], f[:, C:]
        ga, be = ga * sc, be * sc + sh
    o = v * ga[:, None, :] + be[:, None, :]
    if silu:
        t = torch.from_numpy(o)
        o = (t * torch.sigmoid(t)).numpy()
    # op-mode word: bits 0-2 = HL_OP_* of y (lo offset in bits 8+), bits 4-6 = HL_OP_* of the raw copy (lo at channel C)
    y_mode = (round_tf32 & 7) | (round_tf32 & ~0xFF)
    raw_mode = ((round_tf32 >> 4) & 7) | (C << 8)
    _store_operand(y, y_dtype, B * HW, C, ldy, o.reshape(B * HW, C).astype(np.float32), y_mode)
    if raw:
        _store_operand(raw, y_dtype, B * HW, C, ldraw, v.reshape(B * HW, C).copy(), raw_mode)


def hl_conv2d(x, x_dtype, ldx, wpk, bias, residual, ldr, y, ldy, stats, stats_ld, B, H, W, Cin, Cout, ksize, stride,
              flags, stream):
    assert not (flags & 2), "the plan upsamples explicitly"
    cout_pad = (Cout + 31) // 32 * 32
    esz = 2 if x_dtype == 1 else 4
    taps = ksize * ksize

    def xop(off):
        return torch.from_numpy(pitched(x + esz * off, B * H * W, Cin, ldx, _dt(x_dtype)).astype(np.float32)) \
            .reshape(B, H, W, Cin).permute(0, 3, 1, 2)

    def wslab(i):
        wv = torch.from_numpy(view(wpk + esz * i * taps * cout_pad * Cin, taps * cout_pad * Cin, _dt(x_dtype)).astype(np.float32))
        return wv.reshape(ksize, ksize, cout_pad, Cin)[:, :, :Cout].permute(2, 3, 0, 1).contiguous()

    bv = torch.from_numpy(view(bias, Cout).copy())
    cv = lambda a, w: F.conv2d(a, w, None, stride=stride, padding=ksize // 2)
    if flags & 16:        # HL_CONV_SPLIT3: x = [hi | lo], w = {W_hi, W_lo}: hi.hi + lo.hi + hi.lo
        assert x_dtype == 1 and ldx >= 2 * Cin
        out = cv(xop(0), wslab(0)) + cv(xop(Cin), wslab(0)) + cv(xop(0), wslab(1))
    elif flags & 128:     # HL_CONV_SPLIT2A: x = [hi | lo], one weight slab: hi.W + lo.W
        assert x_dtype == 1 and ldx >= 2 * Cin
        out = cv(xop(0), wslab(0)) + cv(xop(Cin), wslab(0))
    elif flags & 32:      # HL_CONV_SPLIT2P: hi and lo packed inside the Cin channels, w = {[W_hi | W_hi], [W_lo | 0]}
        out = cv(xop(0), wslab(0)) + cv(xop(0), wslab(1))
    else:
        out = cv(xop(0), wslab(0))
    out = (out + bv[None, :, None, None]).permute(0, 2, 3, 1)
    Ho, Wo = out.shape[1], out.shape[2]
    out = out.reshape(B * Ho * Wo, Cout).numpy()
    if residual:
        out = out + pitched(residual, B * Ho * Wo, Cout, ldr)
    if flags & 64:                                          # HL_CONV_OUT_F16_SPLIT: [hi | lo] of out * 2^-4
        assert not stats and ldy >= 2 * Cout
        _store_operand(y, 1, B * Ho * Wo, Cout, ldy, out, OP_SCALED | OP_SPLIT | (Cout << 8))
        return
    if flags & 8:                                           # HL_CONV_OUT_F16 (statistics: of the rounded values)
        pitched(y, B * Ho * Wo, Cout, ldy, np.float16)[...] = out.astype(np.float16)
        if stats:
            hl_gn_stats(y, ldy, B, Ho * Wo, Cout, stats, stats_ld, stream, dtype=np.float16)
        return
    pitched(y, B * Ho * Wo, Cout, ldy)[...] = out
    if stats:
        hl_gn_stats(y, ldy, B, Ho * Wo, Cout, stats, stats_ld, stream)


def hl_gn_skip(x, ldx, stats, stats_ld, gamma, beta, act, ld_act, wpk, bias, skip, ld_skip, B, HW, Cin, Cout, groups, eps, stream):
    """GroupNorm-1 + SiLU operand pass fused with the 1x1 skip conv = hl_gn_apply with the scaled hi | lo raw copy, then
    hl_conv2d(HL_CONV_SPLIT3) on that copy (H x W = HW x 1: a 1x1 conv does not look at the geometry)."""
    raw = np.zeros((B * HW, 2 * Cin), np.float16)
    raw_ptr = raw.ctypes.data
    hl_gn_apply(x, ldx, stats, stats_ld, gamma, beta, None, 0, act, 1, ld_act, raw_ptr, 2 * Cin, B, HW, Cin, groups, eps, 1,
                (OP_SPLIT | OP_SCALED) << 4, stream)
    hl_conv2d(raw_ptr, 1, 2 * Cin, wpk, bias, None, 0, skip, ld_skip, None, 0, B, HW, 1, Cin, Cout, 1, 1, 16, stream)


def hl_conv2d_dual(x, x_dtype, ldx, wpk, bias, residual, ldr, y, ldy, stats, stats_ld, y2, ldy2, stats2, stats2_ld, B, H, W,
                   Cin, Cout, ksize, stride, flags, stream):
    """y = conv + residual, y2 = conv: the two launches the entry point replaces."""
    hl_conv2d(x, x_dtype, ldx, wpk, bias, None, 0, y2, ldy2, stats2, stats2_ld, B, H, W, Cin, Cout, ksize, stride, flags, stream)
    hl_conv2d(x, x_dtype, ldx, wpk, bias, residual, ldr, y, ldy, stats, stats_ld, B, H, W, Cin, Cout, ksize, stride, flags, stream)


def hl_attention(qkv, qkv_dtype, ldq, out, out_dtype, ldo, B, T, C, heads, round_tf32, stream):
    v = torch.from_numpy(pitched(qkv, B * T, 3 * C, ldq, _dt(qkv_dtype)).astype(np.float32))
    v = v.reshape(B, T, heads, 3, C // heads)
    q, k, vv = v[:, :, :, 0], v[:, :, :, 1], v[:, :, :, 2]            # [B, T, heads, ch]
    w = torch.einsum("bthc,bshc->bhts", q, k) / (C // heads) ** 0.5
    w = torch.softmax(w, dim=-1)
    a = torch.einsum("bhts,bshc->bthc", w, vv).reshape(B * T, C).numpy()
    _store_operand(out, out_dtype, B * T, C, ldo, a, round_tf32)


_OPS = {k: v for k, v in globals().items() if k.startswith("hl_")}


def run_plan(plan, x, timesteps, x_cond, y):
    """Interpret the plan's launch list on the CPU; returns eps [B, C, H, W]."""
    plan.x_in.copy_(x)
    plan.t_in.copy_(timesteps)
    if plan.xc_in is not None:
        plan.xc_in.copy_(x_cond)
    if y is not None:
        plan.y_in.copy_(y)
    for name, args, _branch in plan.calls:
        if name[0] != "#":                      # stream-dependency markers carry no arithmetic
            _OPS[name](*args, None)
    return plan.out.clone()
