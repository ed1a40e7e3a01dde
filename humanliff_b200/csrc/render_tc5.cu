// Fused tri-plane volume renderer on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// Same per-ray chain as render_tc.cu / render.cu (recon_NeRF/lib/renderer.py:142-295,504-581,
// run_nerf_batch.py:29-67):
//   coarse z -> nine-plane bilinear gather -> density MLP -> up_sample / sample_pdf -> merge-sort ->
//   256-sample fine pass (+ view-direction branch) -> alpha compositing.
// What changed against the mma.sync kernel (220 registers / thread, 8 warps per SM, every phase of a ray alone
// on the SM -- profiles/r1_kernels_full_v10.md): the activations no longer live in registers.
//   * one persistent CTA per SM runs TWO independent ray groups of 128 threads; thread i of a group owns
//     sample i of the group's current ray = row i of every GEMM = TMEM lane i;
//   * every layer is D[128 x N] = A[128 x K] . W^T on tcgen05.mma kind::f16 (M = 128, N = 128 | 64): the fp16
//     weights of all five layers sit in shared memory for the life of the CTA as pre-swizzled K-major
//     SWIZZLE_128B atoms (144 KB, the B operand); the A operand is read FROM TENSOR MEMORY (the ".ts" form):
//     gathered features and the activations of the previous layer are written there with tcgen05.st, so
//     activations never touch shared memory or HBM;
//   * the fp32 accumulator (128 TMEM columns per group) is drained by the same 128 threads: tcgen05.ld of
//     this thread's row, + bias, softplus, cvt to fp16 pairs, tcgen05.st as the next layer's A operand; the
//     alpha / rgb heads are plain per-thread dot products (a thread holds its sample's whole feature row);
//   * while one group waits for its MMAs or for L2 gather latency, the other group's epilogue keeps the MUFU
//     pipe (softplus = ex2 + lg2, the binding unit: 0.33 M transcendentals per ray) busy.
// TMEM columns per group: accumulator 128 | A_h 64 (128 fp16 hidden activations) | A_x 16 (32 fp16 features).
// Operand rounding is identical to render_tc.cu (fp16 features / activations / weights, fp32 accumulate).
#include "common.cuh"
#include "tc5.cuh"

#include <cuda_fp16.h>

int hl_num_sms();

namespace {

constexpr int NS = 128;                 // samples per pass = threads per ray group
constexpr int GROUPS = 2;
constexpr int NT5 = NS * GROUPS;        // threads per CTA

// pre-swizzled fp16 weight image (bytes); every atom = [rows][64 halves] K-major, 128 B rows, SWIZZLE_128B
constexpr int OW0 = 0;                          // pts_linears.0   128 x 64  (k 0..26 used, 27..63 zero)
constexpr int OW1 = OW0 + 16384;                // pts_linears.1   128 x 128 (2 atoms)
constexpr int OW2X = OW1 + 32768;               // pts_linears.2, x part   128 x 64 (k 0..26 used)
constexpr int OW2H = OW2X + 16384;              // pts_linears.2, h1 part  128 x 128 (2 atoms)
constexpr int OWF = OW2H + 32768;               // feature_linear  128 x 128 (2 atoms)
constexpr int OWV = OWF + 32768;                // views_linear (feature part)  64 x 128 (2 atoms of 8 KB)
constexpr int W_BYTES = OWV + 16384;
static_assert(W_BYTES == HL_MLP16S_BYTES, "header and kernel disagree on the swizzled fp16 MLP image");

// fp32 table in shared memory
constexpr int FB_B0 = 0, FB_B1 = 128, FB_B2 = 256, FB_BF = 384, FB_WA = 512, FB_BA = 640, FB_BV = 644;
constexpr int FB_WVPE = 708;                 // [27][64]
constexpr int FB_WR = FB_WVPE + 27 * 64;     // [64][4]
constexpr int FB_BR = FB_WR + 64 * 4;        // [4]
constexpr int FB_FLOATS = FB_BR + 4;

// per-group scratch (floats)
constexpr int SC_ZC = 0, SC_ZN = 128, SC_ZF = 256, SC_CDF = 512, SC_BINS = 640, SC_PEB = 768, SC_PE = 832,
              SC_RED = 864, SC_FLOATS = 928;

constexpr size_t SMEM5 = 1024 + (size_t)W_BYTES + sizeof(float) * (FB_FLOATS + GROUPS * SC_FLOATS);

// TMEM columns (per group: 256-column stride)
constexpr uint32_t TM_ACC = 0, TM_AH = 128, TM_AX = 192, TM_GROUP = 256, TM_COLS = 512;

struct Render5Args {
    const float4 *tex;
    int R;
    const float *mlp;                 // fp32 pack (biases, heads, view-direction weights), HL_MLP_* offsets
    const uint4 *w16s;                // pre-swizzled fp16 weight image (HL_MLP16S_BYTES)
    const float *o, *d, *near, *far, *u, *zc_in;
    unsigned long long seed;
    float bounds[6];
    const float *bounds_dev;          // nullable: device [6] overrides `bounds` (no host sync on a CUDA tensor)
    float *rgb, *acc, *depth;
    long long n_rays;
    int clamp_depth;
    int n_importance;                 // 128, or 0: no coarse pass, composite the 128 coarse samples
    unsigned long long *prof;         // optional cycle counters of (CTA 0, group 0): setup, gather, mlp, resample, composite, total
    // density-grid mode
    int grid_res;
    float *grid_out;
};

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
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
    __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&h);
}

// The 27 features of one world-space point (nine sub-planes x 3 channels), packed as 32 fp16 (27..31 = 0) into
// xa[16].  Sub-planes are fetched in two batches (5 + 4) so that 16-20 independent 16-byte L2 loads are in flight
// per thread without holding all 36 taps in registers.
template <int C0, int NC>
__device__ __forceinline__ void gather_subplanes(const float4 *__restrict__ tex, int R, float cx, float cy, float cz,
                                                 float *f) {
    const float fR = (float)R, shift = 1.0f / fR;
    float4 tap[NC][4];
    float wgt[NC][4];
#pragma unroll
    for (int i = 0; i < NC; ++i) {
        const int c = C0 + i;
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
        const int x0 = (int)fminf(fmaxf(fx0, -2.f), (float)R + 1.f);     // clamp before the cast: miss rays are far away
        const int y0 = (int)fminf(fmaxf(fy0, -2.f), (float)R + 1.f);
        const float4 *tp = tex + (size_t)c * R * R;
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
        float r0 = 0.f, r1 = 0.f, r2 = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {      // same accumulation order as render.cu (taps 00, 01, 10, 11)
            r0 = fmaf(tap[i][k].x, wgt[i][k], r0);
            r1 = fmaf(tap[i][k].y, wgt[i][k], r1);
            r2 = fmaf(tap[i][k].z, wgt[i][k], r2);
        }
        f[(C0 + i) * 3 + 0] = r0;
        f[(C0 + i) * 3 + 1] = r1;
        f[(C0 + i) * 3 + 2] = r2;
    }
}

// ONE out-of-line copy (the coarse pass, both fine tiles and the density grid call it): fully unrolled it is ~1.5 k
// instructions, and the r1 kernel showed that duplicated unrolled bodies thrash the instruction cache.
__device__ __noinline__ void gather_to_tmem(const float4 *__restrict__ tex, int R, const float *bnd, float px,
                                            float py, float pz, uint32_t tm_ax) {
    uint32_t xa[16];
    const float cx = 2.f * (px - bnd[0]) / (bnd[3] - bnd[0]) - 1.f;
    const float cy = 2.f * (py - bnd[1]) / (bnd[4] - bnd[1]) - 1.f;
    const float cz = 2.f * (pz - bnd[2]) / (bnd[5] - bnd[2]) - 1.f;
    float f[28];
    gather_subplanes<0, 5>(tex, R, cx, cy, cz, f);
    gather_subplanes<5, 4>(tex, R, cx, cy, cz, f);
    f[27] = 0.f;
#pragma unroll
    for (int k = 0; k < 14; ++k) xa[k] = pack_h2(f[2 * k], f[2 * k + 1]);
    xa[14] = 0u;
    xa[15] = 0u;
    tmem_st16(tm_ax, xa);
}

// ---- per-group machinery ----------------------------------------------------------------------------------------
struct Group {
    int g, tg, warp, lane;            // group id, thread in group, warp in group, lane
    uint32_t tm;                      // TMEM base of the group, lane quarter of this warp folded in
    uint32_t tm_cols;                 // TMEM base of the group (columns only) -- the MMA's D / A addresses
    uint32_t mbar;                    // the group's MMA-completion barrier
    uint32_t phase;
    uint32_t w_smem;                  // shared-memory address of the weight image
    const float *fb;                  // fp32 table
    float *sc;                        // group scratch
};

__device__ __forceinline__ void gbar(const Group &G) { named_bar(1 + G.g, NS); }

// K-major SWIZZLE_128B descriptor of a weight atom (rows 128 B apart, 8-row groups 1024 B apart)
__device__ __forceinline__ uint64_t wdesc(uint32_t saddr) {
    const uint32_t hi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
    const uint32_t lo = ((saddr & 0x3FFFFu) >> 4) | (1u << 16);
    return ((uint64_t)hi << 32) | lo;
}
// instruction descriptor: D = f32, A = B = f16, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
__device__ __forceinline__ uint32_t idesc_n(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); }

// Issue D[acc] (+)= A[tmem a_col, K = 16 * ksteps] . W[atoms at w_off]^T.  Called by ONE elected thread.
__device__ __forceinline__ void issue_gemm(const Group &G, uint32_t a_col, int ksteps, uint32_t w_off, uint32_t atom_bytes,
                                           int n, bool accumulate) {
    const uint32_t id = idesc_n(n);
    const uint32_t d = G.tm_cols + TM_ACC;
#pragma unroll 1
    for (int k = 0; k < ksteps; ++k) {
        const uint64_t bd = wdesc(G.w_smem + w_off + (uint32_t)(k >> 2) * atom_bytes) + (uint64_t)(2 * (k & 3));
        umma_ts_f16(d, G.tm_cols + a_col + 8u * (uint32_t)k, bd, id, (accumulate || k > 0) ? 1u : 0u);
    }
}

// everyone's tcgen05.st has landed -> one thread issues the layer's MMAs and commits to the group barrier
template <typename F>
__device__ __forceinline__ void run_layer(Group &G, F &&issue) {
    tmem_st_wait();
    tc_fence_before();
    gbar(G);
    if (G.warp == 0) {
        tc_fence_after();
        if (elect_one_sync()) {
            issue();
            umma_commit(G.mbar);
        }
        __syncwarp();
    }
    mbar_wait(G.mbar, G.phase);
    G.phase ^= 1u;
    tc_fence_after();
}

// Drain 32 accumulator columns [c0, c0+32) of this thread's row.
__device__ __forceinline__ void acc_ld(const Group &G, int c0, uint32_t *r) { tmem_ld32_nowait(G.tm + TM_ACC + (uint32_t)c0, r); }

// hidden layer epilogue: h = softplus(acc + bias) -> fp16 -> A_h; optionally the alpha head on the fp32 values
template <bool ALPHA, bool ACT>
__device__ __forceinline__ float epi_hidden(const Group &G, const float *bias, const float *wa) {
    float s = 0.f;
    uint32_t va[32], vb[32];
    acc_ld(G, 0, va);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        uint32_t *cur = (c & 1) ? vb : va;
        uint32_t *nxt = (c & 1) ? va : vb;
        tmem_ld_wait();
        if (c < 3) acc_ld(G, (c + 1) * 32, nxt);
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 b = *reinterpret_cast<const float4 *>(bias + c * 32 + 4 * j);
            float h0 = __uint_as_float(cur[4 * j]) + b.x, h1 = __uint_as_float(cur[4 * j + 1]) + b.y;
            float h2 = __uint_as_float(cur[4 * j + 2]) + b.z, h3 = __uint_as_float(cur[4 * j + 3]) + b.w;
            if (ACT) { h0 = softplus_fast(h0); h1 = softplus_fast(h1); h2 = softplus_fast(h2); h3 = softplus_fast(h3); }
            pk[2 * j] = pack_h2(h0, h1);
            pk[2 * j + 1] = pack_h2(h2, h3);
            if (ALPHA) {
                // same operand as render_tc.cu: the fp16-rounded activation, fp32 weights and accumulation
                const float4 w = *reinterpret_cast<const float4 *>(wa + c * 32 + 4 * j);
                const float2 q0 = __half22float2(*reinterpret_cast<const __half2 *>(&pk[2 * j]));
                const float2 q1 = __half22float2(*reinterpret_cast<const __half2 *>(&pk[2 * j + 1]));
                s = fmaf(q0.x, w.x, s); s = fmaf(q0.y, w.y, s); s = fmaf(q1.x, w.z, s); s = fmaf(q1.y, w.w, s);
            }
        }
        tmem_st16(G.tm + TM_AH + (uint32_t)(c * 16), pk);
    }
    return s;
}

// views_linear epilogue: softplus(acc[0..63] + peb) . rgb_linear -> sigmoid
__device__ __forceinline__ void epi_views(const Group &G, const float *peb, float (&rgb)[3]) {
    float r0 = 0.f, r1 = 0.f, r2 = 0.f;
    uint32_t va[32], vb[32];
    acc_ld(G, 0, va);
    tmem_ld_wait();
    acc_ld(G, 32, vb);
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        uint32_t *cur = c ? vb : va;
        if (c) tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float2 b = *reinterpret_cast<const float2 *>(peb + c * 32 + 2 * j);
            const uint32_t p = pack_h2(softplus_fast(__uint_as_float(cur[2 * j]) + b.x),
                                       softplus_fast(__uint_as_float(cur[2 * j + 1]) + b.y));
            const float2 q = __half22float2(*reinterpret_cast<const __half2 *>(&p));
            const float4 w0 = *reinterpret_cast<const float4 *>(G.fb + FB_WR + (c * 32 + 2 * j) * 4);
            const float4 w1 = *reinterpret_cast<const float4 *>(G.fb + FB_WR + (c * 32 + 2 * j + 1) * 4);
            r0 = fmaf(q.x, w0.x, r0); r1 = fmaf(q.x, w0.y, r1); r2 = fmaf(q.x, w0.z, r2);
            r0 = fmaf(q.y, w1.x, r0); r1 = fmaf(q.y, w1.y, r1); r2 = fmaf(q.y, w1.z, r2);
        }
    }
    rgb[0] = 1.0f / (1.0f + expf(-(r0 + G.fb[FB_BR + 0])));
    rgb[1] = 1.0f / (1.0f + expf(-(r1 + G.fb[FB_BR + 1])));
    rgb[2] = 1.0f / (1.0f + expf(-(r2 + G.fb[FB_BR + 2])));
}

// The decoder MLP for the 128 samples of the group (features already in A_x): returns (sigma, r, g, b).  ONE
// out-of-line copy for every caller; an odd number of layers either way, so the caller flips G.phase once.
__device__ __noinline__ float4 mlp128(Group G, bool fine) {
    // pts_linears.0: K = 32 (features), N = 128
    run_layer(G, [&] { issue_gemm(G, TM_AX, 2, OW0, 16384, 128, false); });
    epi_hidden<false, true>(G, G.fb + FB_B0, nullptr);
    // pts_linears.1: K = 128
    run_layer(G, [&] { issue_gemm(G, TM_AH, 8, OW1, 16384, 128, false); });
    epi_hidden<false, true>(G, G.fb + FB_B1, nullptr);
    // pts_linears.2 on cat([x, h1])
    run_layer(G, [&] {
        issue_gemm(G, TM_AX, 2, OW2X, 16384, 128, false);
        issue_gemm(G, TM_AH, 8, OW2H, 16384, 128, true);
    });
    float4 out;
    out.x = epi_hidden<true, true>(G, G.fb + FB_B2, G.fb + FB_WA) + G.fb[FB_BA];
    out.y = out.z = out.w = 0.f;
    if (fine) {
        // feature_linear (no activation), then views_linear on [feature | pe(d)] (pe part = per-ray bias peb)
        run_layer(G, [&] { issue_gemm(G, TM_AH, 8, OWF, 16384, 128, false); });
        epi_hidden<false, false>(G, G.fb + FB_BF, nullptr);
        run_layer(G, [&] { issue_gemm(G, TM_AH, 8, OWV, 8192, 64, false); });
        float rgb[3];
        epi_views(G, G.sc + SC_PEB, rgb);
        out.y = rgb[0]; out.z = rgb[1]; out.w = rgb[2];
    }
    return out;
}

// exclusive product scan over the group's 128 threads (thread order); *total = product of all 128 factors
__device__ __forceinline__ float excl_cumprod128(const Group &G, float f, float *total) {
    float v = f;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float up = __shfl_up_sync(0xffffffffu, v, o);
        if (G.lane >= o) v *= up;
    }
    float *red = G.sc + SC_RED;
    gbar(G);                                     // previous users of `red` are done
    if (G.lane == 31) red[G.warp] = v;
    gbar(G);
    float pre = 1.f;
    for (int i = 0; i < G.warp; ++i) pre *= red[i];
    float excl = __shfl_up_sync(0xffffffffu, v, 1);
    if (G.lane == 0) excl = 1.f;
    if (total) *total = ((red[0] * red[1]) * red[2]) * red[3];
    return pre * excl;
}

__device__ __forceinline__ float group_sum(const Group &G, float v, int slot) {
    v = hl_warp_sum(v);
    float *red = G.sc + SC_RED + 8 + slot * 4;
    if (G.lane == 0) red[G.warp] = v;
    gbar(G);
    return ((red[0] + red[1]) + red[2]) + red[3];
}

__global__ void __launch_bounds__(NT5, 1) k_render_tc5(const Render5Args a) {
    extern __shared__ __align__(16) uint8_t smraw5[];
    __shared__ __align__(8) uint64_t mbars[GROUPS];
    __shared__ uint32_t tmem_slot;
    const uint32_t base = (smem_u32(smraw5) + 1023u) & ~1023u;
    uint8_t *aligned = smraw5 + (base - smem_u32(smraw5));
    float *fb = reinterpret_cast<float *>(aligned + W_BYTES);
    const int tid = threadIdx.x;

    // one-time: weight image + fp32 table -> shared memory, TMEM allocation, barriers
    {
        uint4 *dst = reinterpret_cast<uint4 *>(aligned);
        for (int i = tid; i < W_BYTES / 16; i += NT5) dst[i] = __ldg(a.w16s + i);
        for (int i = tid; i < 128; i += NT5) {
            fb[FB_B0 + i] = __ldg(a.mlp + HL_MLP_B0 + i);
            fb[FB_B1 + i] = __ldg(a.mlp + HL_MLP_B1 + i);
            fb[FB_B2 + i] = __ldg(a.mlp + HL_MLP_B2 + i);
            fb[FB_BF + i] = __ldg(a.mlp + HL_MLP_BF + i);
            fb[FB_WA + i] = __ldg(a.mlp + HL_MLP_WA + i);
        }
        for (int i = tid; i < 4; i += NT5) {
            fb[FB_BA + i] = __ldg(a.mlp + HL_MLP_BA + i);
            fb[FB_BR + i] = __ldg(a.mlp + HL_MLP_BR + i);
        }
        for (int i = tid; i < 64; i += NT5) fb[FB_BV + i] = __ldg(a.mlp + HL_MLP_BV + i);
        for (int i = tid; i < 27 * 64; i += NT5) fb[FB_WVPE + i] = __ldg(a.mlp + HL_MLP_WV + 128 * 64 + i);
        for (int i = tid; i < 64 * 4; i += NT5) fb[FB_WR + i] = __ldg(a.mlp + HL_MLP_WR + i);
    }
    if (tid < 32) {
        if (tid == 0) {
            for (int g = 0; g < GROUPS; ++g) mbar_init(smem_u32(&mbars[g]), 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                     "r"(TM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_async_smem();                 // generic-proxy writes of the weight image -> visible to the tensor core's reads
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    Group G;
    G.g = tid >> 7;
    G.tg = tid & 127;
    G.warp = G.tg >> 5;
    G.lane = tid & 31;
    G.tm_cols = tmem_slot + (uint32_t)G.g * TM_GROUP;
    G.tm = G.tm_cols + ((uint32_t)(G.warp * 32) << 16);
    G.mbar = smem_u32(&mbars[G.g]);
    G.phase = 0;
    G.w_smem = base;
    G.fb = fb;
    G.sc = fb + FB_FLOATS + G.g * SC_FLOATS;
    float *zc = G.sc + SC_ZC, *zn = G.sc + SC_ZN, *zf = G.sc + SC_ZF, *cdf = G.sc + SC_CDF, *bins = G.sc + SC_BINS;
    float *peb = G.sc + SC_PEB, *pe = G.sc + SC_PE;
    const int tg = G.tg;

    __shared__ float bnd[8];
    if (tid < 6) bnd[tid] = a.bounds_dev ? __ldg(a.bounds_dev + tid) : a.bounds[tid];
    __syncthreads();

    const bool prof = a.prof != nullptr && blockIdx.x == 0 && tid == 0;
    long long tp = prof ? clock64() : 0;
    const long long tp0 = tp;
#define RPROF(slot)                                                        \
    if (prof) {                                                            \
        const long long now_ = clock64();                                  \
        atomicAdd(a.prof + (slot), (unsigned long long)(now_ - tp));       \
        tp = now_;                                                         \
    }

    if (a.grid_out) {
        // ---------------- density-grid mode (Renderer.extract_geometry, human_diffusion/NeRF/renderer.py:290-318) -------
        const int res = a.grid_res;
        const long long total = (long long)res * res * res;
        const long long tiles = (total + 127) / 128;
        auto lin = [&](float lo, float hi, int i) {    // torch.linspace: lo + i*step below the midpoint, hi - (R-1-i)*step above
            const float step = (hi - lo) / (float)(res - 1);
            return i < res / 2 ? __fadd_rn(lo, __fmul_rn(step, (float)i)) : __fsub_rn(hi, __fmul_rn(step, (float)(res - 1 - i)));
        };
        for (long long tile = (long long)blockIdx.x * GROUPS + G.g; tile < tiles; tile += (long long)gridDim.x * GROUPS) {
            long long idx = tile * 128 + tg;
            const bool live = idx < total;
            if (!live) idx = total - 1;
            const int zi = (int)(idx % res), yi = (int)((idx / res) % res), xi = (int)(idx / ((long long)res * res));
            gather_to_tmem(a.tex, a.R, bnd, lin(bnd[0], bnd[3], xi), lin(bnd[1], bnd[4], yi), lin(bnd[2], bnd[5], zi),
                           G.tm + TM_AX);
            const float4 r = mlp128(G, false);
            G.phase ^= 1u;
            if (live) a.grid_out[idx] = -r.x;
        }
    } else {
        const int n_tiles = a.n_importance ? 2 : 1;
        for (long long ray = (long long)blockIdx.x * GROUPS + G.g; ray < a.n_rays; ray += (long long)gridDim.x * GROUPS) {
            const float ox = a.o[ray * 3 + 0], oy = a.o[ray * 3 + 1], oz = a.o[ray * 3 + 2];
            const float dx = a.d[ray * 3 + 0], dy = a.d[ray * 3 + 1], dz = a.d[ray * 3 + 2];
            const float nr = a.near[ray], fr = a.far[ray];
            const float dnorm = sqrtf(dx * dx + dy * dy + dz * dz);
            float zmine;
            {
                const float step = 1.0f / 127.0f;
                const float t = tg < 64 ? step * (float)tg : 1.0f - step * (float)(127 - tg);
                zmine = a.zc_in ? a.zc_in[ray * NS + tg] : __fadd_rn(__fmul_rn(nr, 1.0f - t), __fmul_rn(fr, t));
            }
            gbar(G);                                  // the previous ray of this group is fully consumed
            zc[tg] = zmine;
            if (tg < 27) {                            // positional encoding of the view direction (fields.py:69-85)
                const float dd[3] = {dx / dnorm, dy / dnorm, dz / dnorm};
                float v;
                if (tg < 3) {
                    v = dd[tg];
                } else {
                    const int f = (tg - 3) / 3, comp = (tg - 3) % 3;
                    const float freq = (float)(1 << (f >> 1));
                    const float phase = (f & 1) ? 1.5707963267948966f : 0.f;
                    v = sinf(__fadd_rn(phase, __fmul_rn(dd[comp], freq)));
                }
                pe[tg] = v;
            }
            gbar(G);
            if (tg < 64) {                            // views_linear bias incl. the positional-encoding columns
                float s = fb[FB_BV + tg];
#pragma unroll
                for (int k = 0; k < 27; ++k) s = fmaf(fb[FB_WVPE + k * 64 + tg], pe[k], s);
                peb[tg] = s;
            }
            RPROF(0)
            float rgb0[3] = {0.f, 0.f, 0.f}, rgb1[3] = {0.f, 0.f, 0.f}, sig0 = 0.f, sig1 = 0.f;
            if (a.n_importance) {
                // ------------------------------- coarse pass (density only) -------------------------------
                gather_to_tmem(a.tex, a.R, bnd, __fadd_rn(ox, __fmul_rn(dx, zmine)), __fadd_rn(oy, __fmul_rn(dy, zmine)),
                               __fadd_rn(oz, __fmul_rn(dz, zmine)), G.tm + TM_AX);
                RPROF(1)
                const float sigma = mlp128(G, false).x;
                G.phase ^= 1u;
                RPROF(2)
                // ------------------------------- up_sample + sample_pdf -----------------------------------
                const float dist = (tg < NS - 1 ? zc[tg + 1] - zmine : 1e10f) * dnorm;
                const float al = 1.0f - expf(-softplus_acc(sigma) * dist);
                if (tg < NS - 1) bins[tg] = 0.5f * (zc[tg + 1] + zmine);
                const float T = excl_cumprod128(G, (1.0f - al) + 1e-10f, nullptr);
                const float w = al * T;
                const float wv = (tg >= 1 && tg <= NS - 2) ? w + 1e-5f : 0.f;     // weights[..., 1:-1] + 1e-5
                const float tot = group_sum(G, wv, 0);
                {   // cdf = [0, cumsum(pdf)] (127 entries)
                    float v = wv / tot;
                    if (!(tg >= 1 && tg <= NS - 2)) v = 0.f;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const float up = __shfl_up_sync(0xffffffffu, v, o);
                        if (G.lane >= o) v += up;
                    }
                    float *red = G.sc + SC_RED + 16;
                    if (G.lane == 31) red[G.warp] = v;
                    gbar(G);
                    float pre = 0.f;
                    for (int i = 0; i < G.warp; ++i) pre += red[i];
                    if (tg <= NS - 2) cdf[tg] = pre + v;        // cdf[0] = 0, cdf[i] = pdf[1] + ... + pdf[i]
                }
                gbar(G);
                float znew;
                {
                    const float uu = a.u ? a.u[ray * NS + tg] : uniform_hash(a.seed, (unsigned long long)ray, tg);
                    int lo = 0, hi = NS - 1;
                    while (lo < hi) {
                        const int mid = (lo + hi) >> 1;
                        if (cdf[mid] > uu) hi = mid; else lo = mid + 1;
                    }
                    const int below = max(lo - 1, 0), above = min(NS - 2, lo);
                    float den = cdf[above] - cdf[below];
                    if (den < 1e-5f) den = 1.0f;
                    const float t = (uu - cdf[below]) / den;
                    znew = bins[below] + t * (bins[above] - bins[below]);
                    zn[tg] = znew;
                }
                gbar(G);
                {   // sort(cat(z, z_new)) by ranking: coarse z is already sorted
                    int c_lt = 0, n_lt = 0;
                    const float4 *zn4 = reinterpret_cast<const float4 *>(zn);
#pragma unroll 4
                    for (int j = 0; j < NS / 4; ++j) {
                        const float4 q = zn4[j];
                        c_lt += (q.x < zmine) + (q.y < zmine) + (q.z < zmine) + (q.w < zmine);
                        const int j4 = 4 * j;
                        n_lt += ((q.x < znew) || (q.x == znew && j4 < tg)) + ((q.y < znew) || (q.y == znew && j4 + 1 < tg)) +
                                ((q.z < znew) || (q.z == znew && j4 + 2 < tg)) + ((q.w < znew) || (q.w == znew && j4 + 3 < tg));
                    }
                    int lo = 0, hi = NS;                   // number of coarse z <= znew
                    while (lo < hi) {
                        const int mid = (lo + hi) >> 1;
                        if (zc[mid] <= znew) lo = mid + 1; else hi = mid;
                    }
                    zf[tg + c_lt] = zmine;
                    zf[n_lt + lo] = znew;
                }
                gbar(G);
                RPROF(3)
            } else {
                zf[tg] = zmine;
                gbar(G);
            }
            // ------------------------------- fine pass: sorted samples tg and 128 + tg ----------------
            const float z0 = zf[tg], z1 = a.n_importance ? zf[NS + tg] : 0.f;
#pragma unroll 1
            for (int t = 0; t < n_tiles; ++t) {
                const float z = t ? z1 : z0;
                gather_to_tmem(a.tex, a.R, bnd, __fadd_rn(ox, __fmul_rn(dx, z)), __fadd_rn(oy, __fmul_rn(dy, z)),
                               __fadd_rn(oz, __fmul_rn(dz, z)), G.tm + TM_AX);
                RPROF(1)
                const float4 r = mlp128(G, true);
                G.phase ^= 1u;
                if (t == 0) { sig0 = r.x; rgb0[0] = r.y; rgb0[1] = r.z; rgb0[2] = r.w; }
                else { sig1 = r.x; rgb1[0] = r.y; rgb1[1] = r.z; rgb1[2] = r.w; }
                RPROF(2)
            }
            // ------------------------------- composite (renderer.py:222-239) --------------------------
            {
                const int last = n_tiles * NS - 1;
                const float d0 = tg < last ? zf[tg + 1] - z0 : 1e10f;                  // NOT scaled by |d|
                const float al0 = 1.0f - expf(-softplus_acc(sig0) * d0);
                float P0;
                const float T0 = excl_cumprod128(G, (1.0f - al0) + 1e-7f, &P0);
                const float w0 = al0 * T0;
                float w1 = 0.f;
                if (n_tiles == 2) {
                    const float d1 = NS + tg < last ? zf[NS + tg + 1] - z1 : 1e10f;
                    const float al1 = 1.0f - expf(-softplus_acc(sig1) * d1);
                    const float T1 = P0 * excl_cumprod128(G, (1.0f - al1) + 1e-7f, nullptr);
                    w1 = al1 * T1;
                }
                float q5[5] = {w0 + w1, w0 * rgb0[0] + w1 * rgb1[0], w0 * rgb0[1] + w1 * rgb1[1],
                               w0 * rgb0[2] + w1 * rgb1[2], w0 * z0 + w1 * z1};
#pragma unroll
                for (int k = 0; k < 5; ++k) q5[k] = hl_warp_sum(q5[k]);
                float *red = G.sc + SC_RED + 24;          // [5][4]
                gbar(G);
                if (G.lane == 0) {
#pragma unroll
                    for (int k = 0; k < 5; ++k) red[k * 4 + G.warp] = q5[k];
                }
                gbar(G);
                if (tg == 0) {
                    float s[5];
#pragma unroll
                    for (int k = 0; k < 5; ++k) s[k] = ((red[k * 4] + red[k * 4 + 1]) + red[k * 4 + 2]) + red[k * 4 + 3];
                    a.rgb[ray * 3 + 0] = s[1];
                    a.rgb[ray * 3 + 1] = s[2];
                    a.rgb[ray * 3 + 2] = s[3];
                    a.acc[ray] = s[0];
                    float dep = (s[4] - nr) / (fr - nr + 1e-5f);
                    if (a.clamp_depth) dep = fminf(fmaxf(dep, 0.f), 1.f);
                    a.depth[ray] = dep;
                }
            }
            RPROF(4)
        }
    }
    if (prof) atomicAdd(a.prof + 5, (unsigned long long)(clock64() - tp0));
#undef RPROF

    tc_fence_before();
    __syncthreads();
    if (tid < 32) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_slot), "r"(TM_COLS) : "memory");
    }
}

unsigned long long *g_prof5 = nullptr;

int launch5(Render5Args &a, long long units, cudaStream_t stream) {
    static bool configured[64] = {};
    int dev = 0;
    HL_CHECK_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        HL_CHECK_CUDA(cudaFuncSetAttribute(k_render_tc5, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM5));
        configured[dev] = true;
    }
    long long grid = hl_num_sms();
    const long long need = (units + GROUPS - 1) / GROUPS;
    if (grid > need) grid = need;
    k_render_tc5<<<(int)grid, NT5, SMEM5, stream>>>(a);
    HL_CHECK_LAUNCH();
    return HL_OK;
}

}  // namespace

extern "C" int hl_render5_set_profile(void *dev_counters) {
    g_prof5 = (unsigned long long *)dev_counters;
    return HL_OK;
}

extern "C" int hl_render_rays_tc5(const float *texels, int R, const float *mlp_packed, const void *mlp_f16_swizzled,
                                  const float *rays_o, const float *rays_d, const float *near, const float *far,
                                  const float *z_coarse, const float *u, uint64_t seed, const float *bounds,
                                  int bounds_on_device, float *rgb, float *acc, float *depth, int64_t n_rays,
                                  int n_importance, int clamp_depth, void *stream) {
    HL_CHECK_ARG(texels && mlp_packed && mlp_f16_swizzled && rays_o && rays_d && near && far && bounds && rgb && acc && depth);
    HL_CHECK_ARG(R > 0 && n_rays > 0 && ((uintptr_t)texels & 15) == 0 && ((uintptr_t)mlp_packed & 15) == 0 &&
                 ((uintptr_t)mlp_f16_swizzled & 15) == 0);
    HL_CHECK_ARG(n_importance == 0 || n_importance == NS);
    Render5Args a = {};
    a.tex = reinterpret_cast<const float4 *>(texels);
    a.R = R;
    a.mlp = mlp_packed;
    a.w16s = reinterpret_cast<const uint4 *>(mlp_f16_swizzled);
    a.o = rays_o; a.d = rays_d; a.near = near; a.far = far; a.u = u; a.zc_in = z_coarse;
    a.seed = seed;
    if (bounds_on_device) a.bounds_dev = bounds;
    else for (int i = 0; i < 6; ++i) a.bounds[i] = bounds[i];
    a.rgb = rgb; a.acc = acc; a.depth = depth;
    a.n_rays = n_rays;
    a.clamp_depth = clamp_depth;
    a.n_importance = n_importance;
    a.prof = g_prof5;
    return launch5(a, n_rays, (cudaStream_t)stream);
}

extern "C" int hl_density_grid_tc5(const float *texels, int R, const float *mlp_packed, const void *mlp_f16_swizzled,
                                   const float *bounds, int bounds_on_device, int resolution, float *out, void *stream) {
    HL_CHECK_ARG(texels && mlp_packed && mlp_f16_swizzled && bounds && out && R > 0 && resolution >= 2);
    HL_CHECK_ARG(((uintptr_t)texels & 15) == 0 && ((uintptr_t)mlp_packed & 15) == 0 && ((uintptr_t)mlp_f16_swizzled & 15) == 0);
    Render5Args a = {};
    a.tex = reinterpret_cast<const float4 *>(texels);
    a.R = R;
    a.mlp = mlp_packed;
    a.w16s = reinterpret_cast<const uint4 *>(mlp_f16_swizzled);
    if (bounds_on_device) a.bounds_dev = bounds;
    else for (int i = 0; i < 6; ++i) a.bounds[i] = bounds[i];
    a.grid_res = resolution;
    a.grid_out = out;
    a.n_importance = NS;
    const long long tiles = ((long long)resolution * resolution * resolution + 127) / 128;
    return launch5(a, tiles, (cudaStream_t)stream);
}
