// Fused tri-plane volume renderer, tensor-core version (operand mode fp16).
//
// Same per-ray chain as render.cu (recon_NeRF/lib/renderer.py:142-295,504-581, run_nerf_batch.py:29-67):
//   coarse z -> nine-plane bilinear gather -> density MLP -> up_sample / sample_pdf -> merge-sort ->
//   256-sample fine pass (+ view-direction branch) -> alpha compositing,
// but every 128-point x {128|64}-feature layer of the decoder MLP runs on the tensor cores
// (mma.sync m16n8k16, fp16 operands, fp32 accumulate):
//   * one persistent CTA (8 warps) per ray slot; a warp owns 16 of the tile's 128 points for the WHOLE MLP,
//     so the activations never leave registers: the fp32 accumulator fragment of layer L (after bias +
//     softplus) is packed to fp16 and IS the A fragment of layer L+1 (the flash-attention P.V trick);
//   * all weights (fp16, 140 KB) live in shared memory for the lifetime of the CTA, rows padded by 16 B
//     so ldmatrix is bank-conflict free; the per-ray view-direction term of views_linear is folded into
//     its bias once per ray;
//   * alpha (density) and rgb heads are dot products taken from the fp32 fragments (quad shuffles);
//   * gathered features are staged as fp16 [128][32] (27 + zero padding).
// Measured operand-rounding error of this path (CPU emulation on the reference golden): rgb 1.2e-5,
// depth 4e-5 rel-L2 -- two orders inside the 1e-3 bar, because compositing averages 256 samples.
// Bound: MUFU (softplus = ex2 + lg2 on 164 k hidden activations per ray) and shared-memory weight
// traffic, ~2000 cycles each per 128x128x128 layer tile; the tensor pipe is not the limiter.
#include "common.cuh"

#include <cuda_fp16.h>

int hl_num_sms();

namespace {

constexpr int NS = 128;     // samples per pass
constexpr int NT = 256;     // threads per CTA
constexpr int XP = 40;      // Xs row pitch (halves): 32 features + 8 pad -> 80 B, conflict-free ldmatrix

// fp16 weight image (offsets in halves); rows = output features, pitch = K + 8
constexpr int P0 = 40, P1 = 136, P2 = 168, PF = 136, PV = 136;
constexpr int OW0 = 0;                    // pts_linears.0   [128][32+8]   (k 27..31 zero)
constexpr int OW1 = OW0 + 128 * P0;       // pts_linears.1   [128][128+8]
constexpr int OW2 = OW1 + 128 * P1;       // pts_linears.2   [128][160+8]  k = [x(27) pad(5) | h1(128)]
constexpr int OWF = OW2 + 128 * P2;       // feature_linear  [128][128+8]
constexpr int OWV = OWF + 128 * PF;       // views_linear    [64][128+8]   (feature part only)
constexpr int W_HALVES = OWV + 64 * PV;
static_assert(W_HALVES == HL_MLP16_HALVES, "header and kernel disagree on the fp16 MLP image");
// compact fp32 table kept in shared memory (biases, heads, view-direction rows of views_linear)
constexpr int FB_B0 = 0, FB_B1 = 128, FB_B2 = 256, FB_BF = 384, FB_WA = 512, FB_BA = 640, FB_BV = 644;
constexpr int FB_WVPE = 708;                 // [27][64]
constexpr int FB_WR = FB_WVPE + 27 * 64;     // [64][4]
constexpr int FB_BR = FB_WR + 64 * 4;        // [4]
constexpr int FB_FLOATS = FB_BR + 4;

struct RenderArgs {
    const float4 *tex;
    int R;
    const float *mlp;          // fp32 pack (biases, heads, view-direction weights)
    const __half *w16;         // fp16 weight image
    const float *o, *d, *near, *far, *u, *zc_in;
    unsigned long long seed;
    float bmin[3], bmax[3];
    float *rgb, *acc, *depth;
    long long n_rays;
    int clamp_depth;
    unsigned long long *prof;   // optional: cycles of CTA 0 per phase (0 setup, 1 gather, 2 mlp, 3 resample+sort, 4 composite, 5 total)
};

// softplus(x) = max(x, 0) + ln(1 + e^-|x|): 6 instructions (FMUL, MUFU.EX2, FADD, MUFU.LG2, FMNMX, FFMA).  The
// reference's threshold (x > 20 -> x) needs no select: there the log term is < 2.1e-9 and x + it rounds to x.
__device__ __forceinline__ float softplus_fast(float x) {
    const float l = hl_lg2(1.0f + hl_ex2(-1.4426950408889634f * fabsf(x)));
    return fmaf(l, 0.6931471805599453f, fmaxf(x, 0.f));
}
__device__ __forceinline__ float softplus_acc(float x) { return x > 20.f ? x : log1pf(expf(x)); }

__device__ __forceinline__ float uniform_hash(unsigned long long seed, unsigned long long ray, int i) {
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (ray * 128ull + (unsigned long long)i + 1ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (float)(z >> 40) * (1.0f / 16777216.0f);
}

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
                 "{%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
    __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&h);
}
// F.softplus(beta=1, threshold=20) in fp32 (ex2 + lg2 on the MUFU pipe), packed to fp16 afterwards.
// Tried and rejected (measured / emulated): a half2 Horner polynomial for log1p (4e-3 absolute error -- a
// systematic distortion of the activation, not rounding noise) and an fp32 polynomial (one MUFU op less but
// +8 FMA-pipe instructions per activation: the MLP is as much issue-bound as MUFU-bound, it got slower).
__device__ __forceinline__ uint32_t softplus_h2(float a, float b) { return pack_h2(softplus_fast(a), softplus_fast(b)); }
__device__ __forceinline__ float2 unpack_h2(uint32_t v) {
    return __half22float2(*reinterpret_cast<const __half2 *>(&v));
}

// Features of one world-space point: sub-planes c = C0, C0 + CSTEP, ... (< 9) are gathered by this thread and
// written to xrow[c*3 .. c*3+2].  Fully unrolled so that all 4 x NC bilinear taps (16-byte L2 loads) are in
// flight together (the rolled loop exposed one L2 round trip per sub-plane: 4 k cycles per 128-point tile).
template <int C0, int CSTEP>
__device__ __forceinline__ void gather_point(const RenderArgs &a, float px, float py, float pz, __half *xrow) {
    const float cx = 2.f * (px - a.bmin[0]) / (a.bmax[0] - a.bmin[0]) - 1.f;
    const float cy = 2.f * (py - a.bmin[1]) / (a.bmax[1] - a.bmin[1]) - 1.f;
    const float cz = 2.f * (pz - a.bmin[2]) / (a.bmax[2] - a.bmin[2]) - 1.f;
    const int R = a.R;
    const float fR = (float)R, shift = 1.0f / fR;
    constexpr int NC = (9 - C0 + CSTEP - 1) / CSTEP;
    float4 tap[NC][4];
    float wgt[NC][4];
#pragma unroll
    for (int i = 0; i < NC; ++i) {
        const int c = C0 + i * CSTEP;
        const int plane = c / 3, sub = c - plane * 3;
        float u = (plane == 2) ? cz : cx;
        float v = (plane == 1) ? cz : cy;
        if (sub == 1) u += shift;
        if (sub == 2) v += shift;
        const float ix = ((u + 1.f) * fR - 1.f) * 0.5f;
        const float iy = ((v + 1.f) * fR - 1.f) * 0.5f;
        const float fx0 = floorf(ix), fy0 = floorf(iy);
        const float wx1 = ix - fx0, wy1 = iy - fy0;
        const float wx0 = (fx0 + 1.f) - ix, wy0 = (fy0 + 1.f) - iy;
        // clamp before the int cast so far-away points (miss rays) cannot overflow
        const int x0 = (int)fminf(fmaxf(fx0, -2.f), (float)R + 1.f);
        const int y0 = (int)fminf(fmaxf(fy0, -2.f), (float)R + 1.f);
        const float4 *tp = a.tex + (size_t)c * R * R;
        const bool xin0 = x0 >= 0 && x0 < R, xin1 = x0 + 1 >= 0 && x0 + 1 < R;
        const bool yin0 = y0 >= 0 && y0 < R, yin1 = y0 + 1 >= 0 && y0 + 1 < R;
        const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
        tap[i][0] = (yin0 && xin0) ? __ldg(tp + (size_t)y0 * R + x0) : zero;
        tap[i][1] = (yin0 && xin1) ? __ldg(tp + (size_t)y0 * R + x0 + 1) : zero;
        tap[i][2] = (yin1 && xin0) ? __ldg(tp + (size_t)(y0 + 1) * R + x0) : zero;
        tap[i][3] = (yin1 && xin1) ? __ldg(tp + (size_t)(y0 + 1) * R + x0 + 1) : zero;
        wgt[i][0] = wx0 * wy0; wgt[i][1] = wx1 * wy0; wgt[i][2] = wx0 * wy1; wgt[i][3] = wx1 * wy1;
    }
#pragma unroll
    for (int i = 0; i < NC; ++i) {
        const int c = C0 + i * CSTEP;
        float r0 = 0.f, r1 = 0.f, r2 = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {      // same accumulation order as render.cu (taps 00, 01, 10, 11)
            r0 = fmaf(tap[i][k].x, wgt[i][k], r0);
            r1 = fmaf(tap[i][k].y, wgt[i][k], r1);
            r2 = fmaf(tap[i][k].z, wgt[i][k], r2);
        }
        xrow[c * 3 + 0] = __float2half_rn(r0);
        xrow[c * 3 + 1] = __float2half_rn(r1);
        xrow[c * 3 + 2] = __float2half_rn(r2);
    }
}

// 128-point tile: threads p and p + 128 split the nine sub-planes of point p (even / odd)
__device__ __forceinline__ void gather_tile128(const RenderArgs &a, float px, float py, float pz, __half *Xs) {
    __half *xrow = Xs + (threadIdx.x & 127) * XP;
    if (threadIdx.x < 128) {
        gather_point<0, 2>(a, px, py, pz, xrow);
    } else {
        gather_point<1, 2>(a, px, py, pz, xrow);
#pragma unroll
        for (int k = 27; k < 32; ++k) xrow[k] = __float2half_rn(0.f);   // zero padding of features 27..31
    }
}
// 256-point tile: one point per thread
__device__ __forceinline__ void gather_tile256(const RenderArgs &a, float px, float py, float pz, __half *Xs) {
    __half *xrow = Xs + threadIdx.x * XP;
    gather_point<0, 1>(a, px, py, pz, xrow);
#pragma unroll
    for (int k = 27; k < 32; ++k) xrow[k] = __float2half_rn(0.f);
}

// acc[j][.] += A(16 x 16*KT) . W^T for NT8 output tiles of 8; W rows = outputs, pitch PITCH halves,
// starting at input column k_off.  a[kk] = A fragment of k-tile kk.
template <int KT, int NT8, int PITCH>
__device__ __forceinline__ void gemm_frag(float (&acc)[NT8][4], const uint32_t (*a)[4], uint32_t w_addr, int k_off,
                                          int lane) {
    // lane -> (output row within a pair of n-tiles, k half) of the ldmatrix.x4 that yields {b0,b1} of two n-tiles
    const uint32_t lane_off = (uint32_t)((((lane >> 4) * 8 + (lane & 7)) * PITCH + ((lane >> 3) & 1) * 8 + k_off) * 2);
#pragma unroll
    for (int kk = 0; kk < KT; ++kk) {
#pragma unroll
        for (int jp = 0; jp < NT8 / 2; ++jp) {
            uint32_t b0, b1, b2, b3;
            ldsm_x4(w_addr + lane_off + (uint32_t)((jp * 16 * PITCH + kk * 16) * 2), b0, b1, b2, b3);
            mma16816(acc[2 * jp], a[kk], b0, b1);
            mma16816(acc[2 * jp + 1], a[kk], b2, b3);
        }
    }
}

template <int NT8>
__device__ __forceinline__ void init_bias(float (&acc)[NT8][4], const float *bias, int t) {
#pragma unroll
    for (int j = 0; j < NT8; ++j) {
        const float b0 = bias[j * 8 + 2 * t], b1 = bias[j * 8 + 2 * t + 1];
        acc[j][0] = b0; acc[j][1] = b1; acc[j][2] = b0; acc[j][3] = b1;
    }
}

// softplus on the accumulator fragment, then pack it as the A fragments of the next layer
template <int NT8, bool ACT>
__device__ __forceinline__ void act_pack(float (&acc)[NT8][4], uint32_t (*a)[4]) {
#pragma unroll
    for (int j = 0; j < NT8; ++j) {
        a[j >> 1][(j & 1) * 2 + 0] = ACT ? softplus_h2(acc[j][0], acc[j][1]) : pack_h2(acc[j][0], acc[j][1]);
        a[j >> 1][(j & 1) * 2 + 1] = ACT ? softplus_h2(acc[j][2], acc[j][3]) : pack_h2(acc[j][2], acc[j][3]);
    }
}

// The whole decoder MLP for this warp's 16 points.  sig_out[16] / rgb_out[3][...] receive the heads.
// ONE out-of-line copy shared by the coarse pass, both fine tiles and the density-grid kernel: fully unrolled it
// is ~9 k instructions, and three inlined copies thrashed the instruction cache (measured: MLP phase 50 k ->
// 63 k cycles per ray when the fine pass inlined a second copy).
__device__ __noinline__ void mlp_warp(const bool FINE, uint32_t w_addr, const float *fb, const float *peb,
                                      const __half *Xs, float *sig_out, float *rgb_out, int rgb_pitch, int row0,
                                      int lane) {
    const int g = lane >> 2, t = lane & 3;
    // A fragments of the gathered features (2 k-tiles of 16)
    uint32_t ax[2][4];
    {
        const uint32_t xaddr = (uint32_t)__cvta_generic_to_shared(Xs) +
                               (uint32_t)(((row0 + (lane & 7) + ((lane >> 3) & 1) * 8) * XP + (lane >> 4) * 8) * 2);
        ldsm_x4(xaddr, ax[0][0], ax[0][1], ax[0][2], ax[0][3]);
        ldsm_x4(xaddr + 32, ax[1][0], ax[1][1], ax[1][2], ax[1][3]);
    }
    float acc[16][4];
    uint32_t ah[8][4];
    // pts_linears.0 + softplus
    init_bias<16>(acc, fb + FB_B0, t);
    gemm_frag<2, 16, P0>(acc, ax, w_addr + OW0 * 2, 0, lane);
    act_pack<16, true>(acc, ah);
    // pts_linears.1 + softplus
    init_bias<16>(acc, fb + FB_B1, t);
    gemm_frag<8, 16, P1>(acc, ah, w_addr + OW1 * 2, 0, lane);
    act_pack<16, true>(acc, ah);
    // pts_linears.2 on cat([x, h1]) + softplus
    init_bias<16>(acc, fb + FB_B2, t);
    gemm_frag<2, 16, P2>(acc, ax, w_addr + OW2 * 2, 0, lane);
    gemm_frag<8, 16, P2>(acc, ah, w_addr + OW2 * 2, 32, lane);
    act_pack<16, true>(acc, ah);                            // h2 (fp16) = operand of feature_linear
    // alpha_linear (fp32 dot products on the fragment): rows g and g+8
    {
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float w0 = fb[FB_WA + j * 8 + 2 * t], w1 = fb[FB_WA + j * 8 + 2 * t + 1];
            const float2 v0 = unpack_h2(ah[j >> 1][(j & 1) * 2 + 0]), v1 = unpack_h2(ah[j >> 1][(j & 1) * 2 + 1]);
            s0 = fmaf(v0.x, w0, fmaf(v0.y, w1, s0));
            s1 = fmaf(v1.x, w0, fmaf(v1.y, w1, s1));
        }
        s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
        s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
        if (t == 0) {
            sig_out[row0 + g] = s0 + fb[FB_BA];
            sig_out[row0 + g + 8] = s1 + fb[FB_BA];
        }
    }
    if (FINE) {
        // feature_linear (no activation)
        init_bias<16>(acc, fb + FB_BF, t);
        gemm_frag<8, 16, PF>(acc, ah, w_addr + OWF * 2, 0, lane);
        act_pack<16, false>(acc, ah);
        // views_linear on cat([feature, pe(d)]) + softplus: the pe part is the per-ray bias `peb`
        float av[8][4];
        init_bias<8>(av, peb, t);
        gemm_frag<8, 8, PV>(av, ah, w_addr + OWV * 2, 0, lane);
        float r0[3] = {0.f, 0.f, 0.f}, r1[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float2 v0 = unpack_h2(softplus_h2(av[j][0], av[j][1])), v1 = unpack_h2(softplus_h2(av[j][2], av[j][3]));
            const int k0 = j * 8 + 2 * t;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float w0 = fb[FB_WR + k0 * 4 + c], w1 = fb[FB_WR + (k0 + 1) * 4 + c];
                r0[c] = fmaf(v0.x, w0, fmaf(v0.y, w1, r0[c]));
                r1[c] = fmaf(v1.x, w0, fmaf(v1.y, w1, r1[c]));
            }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            r0[c] += __shfl_xor_sync(0xffffffffu, r0[c], 1); r0[c] += __shfl_xor_sync(0xffffffffu, r0[c], 2);
            r1[c] += __shfl_xor_sync(0xffffffffu, r1[c], 1); r1[c] += __shfl_xor_sync(0xffffffffu, r1[c], 2);
            if (t == 0) {
                const float b = fb[FB_BR + c];
                rgb_out[c * rgb_pitch + row0 + g] = 1.0f / (1.0f + expf(-(r0[c] + b)));
                rgb_out[c * rgb_pitch + row0 + g + 8] = 1.0f / (1.0f + expf(-(r1[c] + b)));
            }
        }
    }
}

__device__ __forceinline__ float block_sum(float v, float *red_s) {
    v = hl_warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) red_s[w] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < NT / 32; ++i) t += red_s[i];
    return t;
}

// inclusive scan over 128 / 256 elements held one per thread: T_i = prod_{j<i} f_j  (exclusive product)
// done with warp shuffles + one shared array of per-warp totals.
__device__ __forceinline__ float excl_cumprod(float f, float *warp_tot, int n_threads) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    float v = f;                                  // inclusive product within the warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float up = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v *= up;
    }
    __syncthreads();
    if (lane == 31) warp_tot[w] = v;
    __syncthreads();
    float pre = 1.f;
    for (int i = 0; i < w; ++i) pre *= warp_tot[i];
    float excl = __shfl_up_sync(0xffffffffu, v, 1);
    if (lane == 0) excl = 1.f;
    (void)n_threads;
    return pre * excl;
}

__global__ void __launch_bounds__(NT, 1) k_render_tc(const RenderArgs a) {
    extern __shared__ __align__(16) uint8_t smraw[];
    __half *Wsm = reinterpret_cast<__half *>(smraw);                 // fp16 weight image
    float *fb = reinterpret_cast<float *>(Wsm + W_HALVES);           // fp32 pack (biases, heads)
    __half *Xs = reinterpret_cast<__half *>(fb + FB_FLOATS);
    float *zc = reinterpret_cast<float *>(Xs + 256 * XP);   // [128] coarse z
    float *zn = zc + NS;                  // [128] new z
    float *zf = zn + NS;                  // [256] merged z
    float *sig = zf + 2 * NS;             // [256] raw density
    float *wts = sig + 2 * NS;            // [256] alpha, then weights
    float *cdf = wts + 2 * NS;            // [128]
    float *bins = cdf + NS;               // [128]
    float *rgbs = bins + NS;              // [3][256]
    float *peb = rgbs + 3 * 2 * NS;       // [64] views bias incl. positional-encoding part
    float *part = peb + 64;               // [128] scratch
    float *red = part + NS;               // [8]
    float *pe = red + 8;                  // [28]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // weights + fp32 pack -> shared memory, once per CTA
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(a.w16);
        uint4 *dst = reinterpret_cast<uint4 *>(Wsm);
        for (int i = tid; i < W_HALVES / 8; i += NT) dst[i] = __ldg(src + i);
        for (int i = tid; i < 128; i += NT) {
            fb[FB_B0 + i] = __ldg(a.mlp + HL_MLP_B0 + i);
            fb[FB_B1 + i] = __ldg(a.mlp + HL_MLP_B1 + i);
            fb[FB_B2 + i] = __ldg(a.mlp + HL_MLP_B2 + i);
            fb[FB_BF + i] = __ldg(a.mlp + HL_MLP_BF + i);
            fb[FB_WA + i] = __ldg(a.mlp + HL_MLP_WA + i);
        }
        for (int i = tid; i < 4; i += NT) {
            fb[FB_BA + i] = __ldg(a.mlp + HL_MLP_BA + i);
            fb[FB_BR + i] = __ldg(a.mlp + HL_MLP_BR + i);
        }
        for (int i = tid; i < 64; i += NT) fb[FB_BV + i] = __ldg(a.mlp + HL_MLP_BV + i);
        for (int i = tid; i < 27 * 64; i += NT) fb[FB_WVPE + i] = __ldg(a.mlp + HL_MLP_WV + 128 * 64 + i);
        for (int i = tid; i < 64 * 4; i += NT) fb[FB_WR + i] = __ldg(a.mlp + HL_MLP_WR + i);
    }
    __syncthreads();
    const uint32_t w_addr = (uint32_t)__cvta_generic_to_shared(Wsm);
    const bool prof = a.prof != nullptr && blockIdx.x == 0 && tid == 0;
    long long tp = prof ? clock64() : 0;
    const long long tp0 = tp;
#define RPROF(slot)                                                        \
    if (prof) {                                                            \
        const long long now_ = clock64();                                  \
        atomicAdd(a.prof + (slot), (unsigned long long)(now_ - tp));       \
        tp = now_;                                                         \
    }

    for (long long ray = blockIdx.x; ray < a.n_rays; ray += gridDim.x) {
        const float ox = a.o[ray * 3 + 0], oy = a.o[ray * 3 + 1], oz = a.o[ray * 3 + 2];
        const float dx = a.d[ray * 3 + 0], dy = a.d[ray * 3 + 1], dz = a.d[ray * 3 + 2];
        const float nr = a.near[ray], fr = a.far[ray];
        const float dnorm = sqrtf(dx * dx + dy * dy + dz * dz);

        __syncthreads();   // previous ray fully consumed
        if (tid < NS) {
            const float step = 1.0f / 127.0f;
            const float t = tid < 64 ? step * (float)tid : 1.0f - step * (float)(127 - tid);
            zc[tid] = a.zc_in ? a.zc_in[ray * NS + tid]
                              : __fadd_rn(__fmul_rn(nr, 1.0f - t), __fmul_rn(fr, t));
        } else if (tid < NS + 27) {
            const int k = tid - NS;
            const float dd[3] = {dx / dnorm, dy / dnorm, dz / dnorm};
            float v;
            if (k < 3) {
                v = dd[k];
            } else {
                const int f = (k - 3) / 3, comp = (k - 3) % 3;
                const float freq = (float)(1 << (f >> 1));
                const float phase = (f & 1) ? 1.5707963267948966f : 0.f;
                v = sinf(__fadd_rn(phase, __fmul_rn(dd[comp], freq)));
            }
            pe[k] = v;
        }
        __syncthreads();
        if (tid < 64) {
            float s = fb[FB_BV + tid];
#pragma unroll
            for (int k = 0; k < 27; ++k) s = fmaf(fb[FB_WVPE + k * 64 + tid], pe[k], s);
            peb[tid] = s;
        }

        // ------------------------------- coarse pass (density only) -------------------------------
        RPROF(0)
        {
            const float z = zc[tid & 127];   // pts = o + d*z (separately rounded, as the reference's broadcasting does)
            gather_tile128(a, __fadd_rn(ox, __fmul_rn(dx, z)), __fadd_rn(oy, __fmul_rn(dy, z)),
                           __fadd_rn(oz, __fmul_rn(dz, z)), Xs);
        }
        __syncthreads();
        RPROF(1)
        mlp_warp(false, w_addr, fb, peb, Xs, sig, nullptr, 0, warp * 16, lane);
        __syncthreads();
        RPROF(2)

        // ------------------------------- up_sample + sample_pdf -----------------------------------
        float al = 0.f;
        if (tid < NS) {
            const float dist = (tid < NS - 1 ? zc[tid + 1] - zc[tid] : 1e10f) * dnorm;
            al = 1.0f - expf(-softplus_acc(sig[tid]) * dist);
            if (tid < NS - 1) bins[tid] = 0.5f * (zc[tid + 1] + zc[tid]);
        }
        {   // weights = alpha * cumprod([1, 1-alpha+1e-10])[:-1]
            const float T = excl_cumprod(tid < NS ? (1.0f - al) + 1e-10f : 1.0f, red, NS);
            if (tid < NS) wts[tid] = al * T;
        }
        __syncthreads();
        {
            const float wv = (tid >= 1 && tid <= NS - 2) ? wts[tid] + 1e-5f : 0.f;   // weights[..., 1:-1] + 1e-5
            const float tot = block_sum(wv, red);
            if (tid >= 1 && tid <= NS - 2) part[tid] = wv / tot;                     // pdf, 126 entries
        }
        __syncthreads();
        {   // cdf = [0, cumsum(pdf)] (127 entries): warp-shuffle inclusive scan + per-warp offsets
            float v = (tid >= 1 && tid <= NS - 2) ? part[tid] : 0.f;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const float up = __shfl_up_sync(0xffffffffu, v, o);
                if (lane >= o) v += up;
            }
            if (lane == 31) red[warp] = v;
            __syncthreads();
            float pre = 0.f;
            for (int i = 0; i < warp; ++i) pre += red[i];
            if (tid <= NS - 2) cdf[tid] = pre + v;          // cdf[0] = 0, cdf[i] = pdf[1] + ... + pdf[i]
        }
        __syncthreads();
        if (tid < NS) {
            const float uu = a.u ? a.u[ray * NS + tid] : uniform_hash(a.seed, (unsigned long long)ray, tid);
            int lo = 0, hi = NS - 1;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (cdf[mid] > uu) hi = mid; else lo = mid + 1;
            }
            const int below = max(lo - 1, 0), above = min(NS - 2, lo);
            float den = cdf[above] - cdf[below];
            if (den < 1e-5f) den = 1.0f;
            const float t = (uu - cdf[below]) / den;
            zn[tid] = bins[below] + t * (bins[above] - bins[below]);
        }
        __syncthreads();
        {   // sort(cat(z, z_new)) by ranking: coarse z is already sorted
            const float v = tid < NS ? zc[tid] : zn[tid - NS];
            int rank;
            if (tid < NS) {
                int cnt = 0;
                for (int j = 0; j < NS; ++j) cnt += (zn[j] < v);
                rank = tid + cnt;
            } else {
                int cnt = 0;
                const int me = tid - NS;
                for (int j = 0; j < NS; ++j) {
                    const float w = zn[j];
                    cnt += (w < v) || (w == v && j < me);
                }
                int lo = 0, hi = NS;       // number of coarse z <= v
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (zc[mid] <= v) lo = mid + 1; else hi = mid;
                }
                rank = cnt + lo;
            }
            zf[rank] = v;
        }
        __syncthreads();

        RPROF(3)
        // ------------------------------- fine pass: all 256 sorted samples at once ----------------
        {
            const float z = zf[tid];
            gather_tile256(a, __fadd_rn(ox, __fmul_rn(dx, z)), __fadd_rn(oy, __fmul_rn(dy, z)),
                           __fadd_rn(oz, __fmul_rn(dz, z)), Xs);
        }
        __syncthreads();
        RPROF(1)
        mlp_warp(true, w_addr, fb, peb, Xs, sig, rgbs, 2 * NS, warp * 16, lane);
        mlp_warp(true, w_addr, fb, peb, Xs, sig, rgbs, 2 * NS, NS + warp * 16, lane);
        __syncthreads();
        RPROF(2)

        // ------------------------------- composite (renderer.py:222-239) --------------------------
        {
            const float dist = tid < 2 * NS - 1 ? zf[tid + 1] - zf[tid] : 1e10f;   // NOT scaled by |d|
            const float al2 = 1.0f - expf(-softplus_acc(sig[tid]) * dist);
            const float T = excl_cumprod((1.0f - al2) + 1e-7f, red, 2 * NS);
            const float w = al2 * T;
            // five sums in one pass: warp shuffles, then 8 per-warp partials per quantity in shared memory
            float q5[5] = {w, w * rgbs[0 * 2 * NS + tid], w * rgbs[1 * 2 * NS + tid], w * rgbs[2 * 2 * NS + tid], w * zf[tid]};
#pragma unroll
            for (int k = 0; k < 5; ++k) q5[k] = hl_warp_sum(q5[k]);
            __syncthreads();                      // `part` (pdf scratch) is free; red is in use by excl_cumprod
            if (lane == 0) {
#pragma unroll
                for (int k = 0; k < 5; ++k) part[k * 8 + warp] = q5[k];
            }
            __syncthreads();
            float s_acc = 0.f, s_r = 0.f, s_g = 0.f, s_b = 0.f, s_d = 0.f;
            if (tid == 0) {
                for (int i = 0; i < 8; ++i) {
                    s_acc += part[0 * 8 + i]; s_r += part[1 * 8 + i]; s_g += part[2 * 8 + i];
                    s_b += part[3 * 8 + i]; s_d += part[4 * 8 + i];
                }
            }
            if (tid == 0) {
                a.rgb[ray * 3 + 0] = s_r;
                a.rgb[ray * 3 + 1] = s_g;
                a.rgb[ray * 3 + 2] = s_b;
                a.acc[ray] = s_acc;
                float dep = (s_d - nr) / (fr - nr + 1e-5f);
                if (a.clamp_depth) dep = fminf(fmaxf(dep, 0.f), 1.f);
                a.depth[ray] = dep;
            }
        }
        RPROF(4)
    }
    if (prof) atomicAdd(a.prof + 5, (unsigned long long)(clock64() - tp0));
#undef RPROF
}

// Density on a regular grid (Renderer.extract_geometry, human_diffusion/NeRF/renderer.py:290-318): the
// coarse-pass stage of the renderer (gather + density MLP) over R^3 points x = linspace(bmin, bmax, R),
// out[xi][yi][zi] = -sigma, 128 points per tile.
__global__ void __launch_bounds__(NT, 1) k_density_grid_tc(const RenderArgs a, int res, float *__restrict__ out) {
    extern __shared__ __align__(16) uint8_t smraw[];
    __half *Wsm = reinterpret_cast<__half *>(smraw);
    float *fb = reinterpret_cast<float *>(Wsm + W_HALVES);
    __half *Xs = reinterpret_cast<__half *>(fb + FB_FLOATS);
    float *sig = reinterpret_cast<float *>(Xs + 128 * XP);     // [128]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(a.w16);
        uint4 *dst = reinterpret_cast<uint4 *>(Wsm);
        for (int i = tid; i < W_HALVES / 8; i += NT) dst[i] = __ldg(src + i);
        for (int i = tid; i < 128; i += NT) {
            fb[FB_B0 + i] = __ldg(a.mlp + HL_MLP_B0 + i);
            fb[FB_B1 + i] = __ldg(a.mlp + HL_MLP_B1 + i);
            fb[FB_B2 + i] = __ldg(a.mlp + HL_MLP_B2 + i);
            fb[FB_WA + i] = __ldg(a.mlp + HL_MLP_WA + i);
        }
        if (tid < 4) fb[FB_BA + tid] = __ldg(a.mlp + HL_MLP_BA + tid);
    }
    __syncthreads();
    const uint32_t w_addr = (uint32_t)__cvta_generic_to_shared(Wsm);
    const long long total = (long long)res * res * res;
    const long long tiles = (total + 127) / 128;
    // torch.linspace(lo, hi, R): lo + i*step below the midpoint, hi - (R-1-i)*step above (ATen)
    auto lin = [&](float lo, float hi, int i) {
        const float step = (hi - lo) / (float)(res - 1);
        return i < res / 2 ? __fadd_rn(lo, __fmul_rn(step, (float)i)) : __fsub_rn(hi, __fmul_rn(step, (float)(res - 1 - i)));
    };
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        long long idx = tile * 128 + (tid & 127);
        if (idx >= total) idx = total - 1;
        const int zi = (int)(idx % res), yi = (int)((idx / res) % res), xi = (int)(idx / ((long long)res * res));
        __syncthreads();     // previous tile consumed (Xs, sig)
        gather_tile128(a, lin(a.bmin[0], a.bmax[0], xi), lin(a.bmin[1], a.bmax[1], yi), lin(a.bmin[2], a.bmax[2], zi), Xs);
        __syncthreads();
        mlp_warp(false, w_addr, fb, nullptr, Xs, sig, nullptr, 0, warp * 16, lane);
        __syncthreads();
        if (tid < 128 && tile * 128 + tid < total) out[tile * 128 + tid] = -sig[tid];
    }
}

}  // namespace

extern "C" int hl_density_grid_tc(const float *texels, int R, const float *mlp_packed, const void *mlp_f16,
                                  const float *bounds, int resolution, float *out, void *stream) {
    HL_CHECK_ARG(texels && mlp_packed && mlp_f16 && bounds && out && R > 0 && resolution >= 2);
    HL_CHECK_ARG(((uintptr_t)texels & 15) == 0 && ((uintptr_t)mlp_packed & 15) == 0 && ((uintptr_t)mlp_f16 & 15) == 0);
    RenderArgs a = {};
    a.tex = reinterpret_cast<const float4 *>(texels);
    a.R = R;
    a.mlp = mlp_packed;
    a.w16 = reinterpret_cast<const __half *>(mlp_f16);
    for (int i = 0; i < 3; ++i) { a.bmin[i] = bounds[i]; a.bmax[i] = bounds[3 + i]; }
    const size_t smem = (size_t)W_HALVES * 2 + sizeof(float) * FB_FLOATS + (size_t)128 * XP * 2 + sizeof(float) * 128;
    static HlPerDeviceOnce once;
    if (once.need()) {
        HL_CHECK_CUDA(cudaFuncSetAttribute(k_density_grid_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    const long long tiles = ((long long)resolution * resolution * resolution + 127) / 128;
    long long grid = hl_num_sms();
    if (grid > tiles) grid = tiles;
    k_density_grid_tc<<<(int)grid, NT, smem, (cudaStream_t)stream>>>(a, resolution, out);
    HL_CHECK_LAUNCH();
    return HL_OK;
}

static unsigned long long *g_render_prof = nullptr;
extern "C" int hl_render_set_profile(void *dev_counters) {
    g_render_prof = (unsigned long long *)dev_counters;
    return HL_OK;
}

extern "C" int hl_render_rays_tc(const float *texels, int R, const float *mlp_packed, const void *mlp_f16,
                                 const float *rays_o, const float *rays_d, const float *near, const float *far,
                                 const float *z_coarse, const float *u, uint64_t seed, const float *bounds,
                                 float *rgb, float *acc, float *depth, int64_t n_rays, int clamp_depth,
                                 void *stream) {
    HL_CHECK_ARG(texels && mlp_packed && mlp_f16 && rays_o && rays_d && near && far && bounds && rgb && acc && depth);
    HL_CHECK_ARG(R > 0 && n_rays > 0 && ((uintptr_t)texels & 15) == 0 && ((uintptr_t)mlp_packed & 15) == 0 &&
                 ((uintptr_t)mlp_f16 & 15) == 0);
    RenderArgs a;
    a.tex = reinterpret_cast<const float4 *>(texels);
    a.R = R;
    a.mlp = mlp_packed;
    a.w16 = reinterpret_cast<const __half *>(mlp_f16);
    a.o = rays_o; a.d = rays_d; a.near = near; a.far = far; a.u = u; a.zc_in = z_coarse;
    a.seed = seed;
    for (int i = 0; i < 3; ++i) { a.bmin[i] = bounds[i]; a.bmax[i] = bounds[3 + i]; }
    a.rgb = rgb; a.acc = acc; a.depth = depth;
    a.n_rays = n_rays;
    a.clamp_depth = clamp_depth;
    a.prof = g_render_prof;
    const size_t smem = (size_t)W_HALVES * 2 + sizeof(float) * FB_FLOATS + (size_t)256 * XP * 2 +
                        sizeof(float) * (size_t)(NS * 2 + 2 * NS * 3 + NS * 2 + 3 * 2 * NS + 64 + NS + 8 + 28 + 4);
    static HlPerDeviceOnce once;
    if (once.need()) {
        HL_CHECK_CUDA(cudaFuncSetAttribute(k_render_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    int64_t grid = hl_num_sms();
    if (grid > n_rays) grid = n_rays;
    k_render_tc<<<(int)grid, NT, smem, (cudaStream_t)stream>>>(a);
    HL_CHECK_LAUNCH();
    return HL_OK;
}
