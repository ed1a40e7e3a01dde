// Fused tri-plane volume renderer on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// Same per-ray chain as render_tc.cu / render.cu (recon_NeRF/lib/renderer.py:142-295,504-581,
// run_nerf_batch.py:29-67):
//   coarse z -> nine-plane bilinear gather -> density MLP -> up_sample / sample_pdf -> merge-sort ->
//   256-sample fine pass (+ view-direction branch) -> alpha compositing.
// The coarse samples go through the WHOLE network once: the reference evaluates their density layers for up_sample and
// then again, identically, inside render_core on the sorted union (renderer.py:258-279); here the colour branch runs on
// the coarse activations while they still sit in tensor memory, the fine pass evaluates only the 128 new samples, and the
// (sigma, rgb) of both sets are scattered into sorted order for compositing: 10 layer passes per ray instead of 13.
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
//   * the epilogue is 5.5 instructions per activation: biases ride in the GEMMs (feature slot 27 of x is 1.0 and
//     carries the bias column of pts_linears.0 / .2; a per-ray constant tile A_c = [1 | pe(d) | 0] carries the
//     biases of pts_linears.1 / feature_linear and the view-direction columns + bias of views_linear), and the
//     softplus runs in the log2 domain, h' = max(a', 0) + lg2(1 + 2^-|a'|) with a' = a * log2(e): the factor
//     log2(e) is folded into the weights that produce a', the factor ln 2 into the weights that consume h'.
// TMEM columns per group: accumulator 128 | A_h 64 (128 fp16 hidden activations) | A_x 16 (32 fp16 features) |
// A_c 16 (the constant tile).  Operand rounding: fp16 features / activations / weights (and biases), fp32 accumulate.
#include "common.cuh"
#include "tc5.cuh"
#include "canon.cuh"

#ifndef HL_R5_CANON_TU
#define HL_R5_CANON_TU 0
#endif

#include <cuda_fp16.h>

int hl_num_sms();

namespace {

constexpr int NS = 128;                 // samples per pass = threads per ray group
constexpr int GROUPS = 2;
constexpr int NT5 = NS * GROUPS;        // threads per CTA

// tcgen05 MLP image (bytes): pre-swizzled fp16 atoms [rows][64 halves] K-major, 128 B rows, SWIZZLE_128B, then a
// small fp32 table.  L2E = log2(e), LN2 = ln 2 (see the header comment).
constexpr int OW0 = 0;                          // pts_linears.0 * L2E   128 x 64  (k 0..26 weights, k 27 = bias * L2E)
constexpr int OW1 = OW0 + 16384;                // pts_linears.1          128 x 128 (2 atoms)
constexpr int OW2X = OW1 + 32768;               // pts_linears.2 x part * L2E   128 x 64 (k 27 = bias * L2E)
constexpr int OW2H = OW2X + 16384;              // pts_linears.2 h1 part  128 x 128 (2 atoms)
constexpr int OWF = OW2H + 32768;               // feature_linear * LN2   128 x 128 (2 atoms)
constexpr int OWV = OWF + 32768;                // views_linear (feature part) * L2E   64 x 128 (2 atoms of 8 KB)
constexpr int OWB = OWV + 16384;                // bias atom 128 x 64: k 0 = pts_linears.1 bias * L2E, k 16 = feature_linear bias
constexpr int OWVP = OWB + 16384;               // 64 x 64: k 0 = views bias * L2E, k 1..27 = view-direction columns * L2E
constexpr int W_BYTES = OWVP + 8192;
constexpr int FB_WA = 0, FB_BA = 128, FB_WR = 132, FB_BR = 388, FB_FLOATS = 392;   // alpha_linear * LN2, rgb_linear^T * LN2 [64][4]
static_assert(W_BYTES + 4 * FB_FLOATS == HL_MLP_TC5_BYTES, "header and kernel disagree on the tcgen05 MLP image");

// per-group scratch (floats)
constexpr int SC_ZC = 0, SC_ZN = 128, SC_ZF = 256, SC_CDF = 512, SC_BINS = 640, SC_PE = 768, SC_RED = 800,
              SC_VAL = 864 /* (sigma, r, g, b) of the 256 samples in sorted order */, SC_FLOATS = 864 + 1024;

constexpr size_t SMEM5 = 1024 + (size_t)W_BYTES + sizeof(float) * (FB_FLOATS + GROUPS * SC_FLOATS);

// TMEM columns (per group: 256-column stride)
constexpr uint32_t TM_ACC = 0, TM_AH = 128, TM_AX = 192, TM_AC = 208, TM_GROUP = 256, TM_COLS = 512;

struct Render5Args {
    const uint4 *tex;                 // quad texels (hl_triplane_to_quads): [9][R + 1][R + 1] x 32 B
    int R;
    const uint4 *w16s;                // tcgen05 MLP image (HL_MLP_TC5_BYTES)
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
    // canonical space (canon != 0): every sample is deformed before the gather (canon.cuh); `bounds` = t_world_bounds
    CanonTables ct;
    int canon;
};

// softplus in the log2 domain: a = x * log2(e) in, softplus(x) / ln 2 out (MUFU.EX2, FADD, MUFU.LG2, FMNMX, FADD)
__device__ __forceinline__ float softplus2(float a) { return fmaxf(a, 0.f) + hl_lg2(1.0f + hl_ex2(-fabsf(a))); }
// The same function with the lg2 on the FMA pipe: lg2(1 + t) = t * q(t) on t in (0, 1], q of degree 5 (Lawson / Remez
// fit, max abs error 2.3e-6 -- two orders below the fp16 rounding of the activation).  The epilogue is bound by the
// MUFU pipe (16 lanes / clk / SM) while the FMA pipe idles, so a fraction of the activations (HL_R5_POLY of every 4)
// takes this route: 1 MUFU + 7 FMA-pipe instructions instead of 2 MUFU + 3.
#ifndef HL_R5_POLY
#define HL_R5_POLY 3
#endif
__device__ __forceinline__ float softplus2_poly(float a) {
    const float t = hl_ex2(-fabsf(a));
    float q = fmaf(-0.026457004986457876f, t, 0.12345017983674147f);
    q = fmaf(q, t, -0.27953670731627006f);
    q = fmaf(q, t, 0.4582700935445837f);
    q = fmaf(q, t, -0.7182817634084746f);
    q = fmaf(q, t, 1.4425531338453292f);
    return fmaf(q, t, fmaxf(a, 0.f));
}
// element e (0..3) of a quad: which route
template <int E>
__device__ __forceinline__ float softplus2_mix(float a) {
    constexpr bool poly = (HL_R5_POLY == 1 && E == 1) || (HL_R5_POLY == 2 && (E & 1)) || (HL_R5_POLY == 3 && E != 0) ||
                          (HL_R5_POLY == 4);
    return poly ? softplus2_poly(a) : softplus2(a);
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ float4 lds_f128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
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

// The 27 features of one world-space point (nine sub-planes x 3 channels), packed as 32 fp16 into xa[16].
// Texels come as "quads" (hl_triplane_to_quads): entry (y0, x0) of a sub-plane holds the 2 x 2 bilinear footprint
// {(y0,x0), (y0,x0+1), (y0+1,x0), (y0+1,x0+1)} x 3 channels as fp16 in ONE 32-byte sector (out-of-range taps stored as
// the zeros grid_sample pads with), so a sub-plane costs one 256-bit load / one L2 sector instead of four 16-byte
// loads over 2-4 sectors.  Measured with the four-tap fp32 layout: the gather took the same 13 k cycles per ray with
// 36 loads per thread or 18 (two threads per sample) -- sector throughput, not latency, bound it.
template <int C0, int NC>
__device__ __forceinline__ void gather_subplanes(const uint4 *__restrict__ tex, int R, float cx, float cy, float cz,
                                                 float *f) {
    const float fR = (float)R, shift = 1.0f / fR;
    const int R1 = R + 1;
    uint32_t q[NC][8];
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
        const bool in = x0 >= -1 && x0 < R && y0 >= -1 && y0 < R;         // at least one tap inside the plane
        const uint4 *tp = tex + 2 * ((size_t)c * R1 * R1 + (size_t)(in ? y0 + 1 : 0) * R1 + (in ? x0 + 1 : 0));
        asm volatile("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(q[i][0]), "=r"(q[i][1]), "=r"(q[i][2]), "=r"(q[i][3]), "=r"(q[i][4]), "=r"(q[i][5]),
                       "=r"(q[i][6]), "=r"(q[i][7])
                     : "l"(tp));
        const float m = in ? 1.f : 0.f;
        wgt[i][0] = wx0 * wy0 * m; wgt[i][1] = wx1 * wy0 * m; wgt[i][2] = wx0 * wy1 * m; wgt[i][3] = wx1 * wy1 * m;
    }
#pragma unroll
    for (int i = 0; i < NC; ++i) {
        // halves: t00.{0,1,2} t01.{0,1,2} t10.{0,1,2} t11.{0,1,2}; accumulation order 00, 01, 10, 11 as render.cu
        float h[12];
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const float2 p = __half22float2(*reinterpret_cast<const __half2 *>(&q[i][k]));
            h[2 * k] = p.x;
            h[2 * k + 1] = p.y;
        }
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            float r = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) r = fmaf(h[3 * k + ch], wgt[i][k], r);
            f[(C0 + i) * 3 + ch] = r;
        }
    }
}

// ONE out-of-line copy (the coarse pass, both fine tiles and the density grid call it): fully unrolled it is ~1.5 k
// instructions, and the r1 kernel showed that duplicated unrolled bodies thrash the instruction cache.
__device__ __noinline__ void gather_to_tmem(const uint4 *__restrict__ tex, int R, const float *bnd, float px,
                                            float py, float pz, uint32_t tm_ax) {
    uint32_t xa[16];
    const float cx = 2.f * (px - bnd[0]) / (bnd[3] - bnd[0]) - 1.f;
    const float cy = 2.f * (py - bnd[1]) / (bnd[4] - bnd[1]) - 1.f;
    const float cz = 2.f * (pz - bnd[2]) / (bnd[5] - bnd[2]) - 1.f;
    float f[28];
    gather_subplanes<0, 9>(tex, R, cx, cy, cz, f);
    f[27] = 1.0f;              // the bias column of pts_linears.0 / pts_linears.2 multiplies this slot
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
    uint32_t w_smem;                  // shared-memory address of the MLP image
    uint32_t fb;                      // shared-memory address of the fp32 table
    uint32_t sc;                      // shared-memory address of the group's scratch
    unsigned long long *prof;         // non-null in the one profiled thread: [6] barrier + issue, [7] wait for the MMAs
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

// Issue D[acc + n_off .. + n) (+)= A[tmem a_col, K = 16 * ksteps] . W[rows n_off .. + n of the atoms at w_off]^T.
// Called by ONE elected thread.
__device__ __forceinline__ void issue_gemm(const Group &G, uint32_t a_col, int ksteps, uint32_t w_off, uint32_t atom_bytes,
                                           int n, bool accumulate, int n_off = 0) {
    const uint32_t id = idesc_n(n);
    const uint32_t d = G.tm_cols + TM_ACC + (uint32_t)n_off;
    const uint32_t w0 = G.w_smem + w_off + (uint32_t)n_off * 128u;
#pragma unroll 1
    for (int k = 0; k < ksteps; ++k) {
        const uint64_t bd = wdesc(w0 + (uint32_t)(k >> 2) * atom_bytes) + (uint64_t)(2 * (k & 3));
        umma_ts_f16(d, G.tm_cols + a_col + 8u * (uint32_t)k, bd, id, (accumulate || k > 0) ? 1u : 0u);
    }
}

// Everyone's tcgen05.st has landed -> one thread issues the layer's MMAs in TWO column halves, each committed to its own
// barrier: the epilogue drains half 0 while the tensor core still computes half 1 (`issue(half)` issues one half;
// views_linear, N = 64, is one half and commits both barriers together).
template <typename F>
__device__ __forceinline__ void start_layer(Group &G, bool two_halves, F &&issue) {
    const long long t0 = G.prof ? clock64() : 0;
    tmem_st_wait();
    tc_fence_before();
    gbar(G);
    if (G.warp == 0) {
        tc_fence_after();
        if (elect_one_sync()) {
            issue(0);
            if (two_halves) {
                umma_commit(G.mbar);
                issue(1);
            } else {
                umma_commit(G.mbar);
            }
            umma_commit(G.mbar + 8u);
        }
        __syncwarp();
    }
    if (G.prof) atomicAdd(G.prof + 6, (unsigned long long)(clock64() - t0));
}
__device__ __forceinline__ void wait_half(Group &G, int half) {
    const long long t0 = G.prof ? clock64() : 0;
    mbar_wait(G.mbar + 8u * (uint32_t)half, G.phase);
    tc_fence_after();
    if (G.prof) atomicAdd(G.prof + 7, (unsigned long long)(clock64() - t0));
}

// Drain 32 accumulator columns [c0, c0+32) of this thread's row.
__device__ __forceinline__ void acc_ld(const Group &G, int c0, uint32_t *r) { tmem_ld32_nowait(G.tm + TM_ACC + (uint32_t)c0, r); }

// The hidden-layer epilogue as two passes over a pair of chunks instead of four unrolled chunks: halves its instruction
// footprint and costs the overlap across the pass boundary.  Same-box A/B (profiles/r2_render_epi_roll_ab.log): canonical
// mode 55.3 -> 53.9 ms per frame (instruction-fetch bound), plain mode 27.62 -> 29.35 ms -- so only the canonical kernel
// (render_tc5_canon.cu) uses it.
#ifndef HL_R5_EPI_ROLL
#define HL_R5_EPI_ROLL HL_R5_CANON_TU
#endif
// 32 accumulator columns of this thread's row -> activation -> 16 packed fp16 pairs (+ the alpha head's partial dot product)
template <bool ACT>
__device__ __forceinline__ void epi_chunk(const uint32_t *cur, uint32_t *pk, bool do_alpha, uint32_t wa_addr, float &s) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float h0 = __uint_as_float(cur[4 * j]), h1 = __uint_as_float(cur[4 * j + 1]);
        float h2 = __uint_as_float(cur[4 * j + 2]), h3 = __uint_as_float(cur[4 * j + 3]);
        if (ACT) { h0 = softplus2_mix<0>(h0); h1 = softplus2_mix<1>(h1); h2 = softplus2_mix<2>(h2); h3 = softplus2_mix<3>(h3); }
        pk[2 * j] = pack_h2(h0, h1);
        pk[2 * j + 1] = pack_h2(h2, h3);
        if (do_alpha) {
            const float4 w = lds_f128(wa_addr + 16u * (uint32_t)j);
            s = fmaf(h0, w.x, s); s = fmaf(h1, w.y, s); s = fmaf(h2, w.z, s); s = fmaf(h3, w.w, s);
        }
    }
}

// hidden layer epilogue: h' = softplus2(acc) (the bias is already in the accumulator) -> fp16 -> A_h; optionally the
// alpha head on the fp32 values.  Column half 0 is drained as soon as ITS MMAs have retired; its packed activations wait
// in registers until half 1 has retired too (those MMAs still read the old A_h), then both halves are stored.
// ALPHA: 0 = no alpha head, 1 = alpha head, 2 = decided at run time by `alpha_rt` (the shared out-of-line copy below)
template <int ALPHA, bool ACT>
__device__ __forceinline__ float epi_hidden(Group &G, bool alpha_rt = false) {
    const bool do_alpha = ALPHA == 2 ? alpha_rt : ALPHA == 1;
    float s = 0.f;
#if HL_R5_EPI_ROLL
    // Two passes over a pair of 32-column chunks (one copy of the pair's code instead of four chunk copies: the unrolled
    // epilogues are the kernel's instruction footprint).  Chunk 2 * it is packed into pk0 and waits in registers for its
    // neighbour: in pass 0 because half 1's MMAs still read the old A_h, in pass 1 only to keep the passes identical.
    uint32_t va[32], vb[32], pk0[16], pk[16];
    wait_half(G, 0);
    acc_ld(G, 0, va);
#pragma unroll 1
    for (int it = 0; it < 2; ++it) {
        tmem_ld_wait();
        acc_ld(G, (2 * it + 1) * 32, vb);
        epi_chunk<ACT>(va, pk0, do_alpha, G.fb + 4u * (uint32_t)(FB_WA + 64 * it), s);
        tmem_ld_wait();
        wait_half(G, 1);                                // pass 0: chunk 2 belongs to half 1 (pass 1: returns at once)
        if (it == 0) acc_ld(G, 64, va);
        epi_chunk<ACT>(vb, pk, do_alpha, G.fb + 4u * (uint32_t)(FB_WA + 64 * it + 32), s);
        tmem_st16(G.tm + TM_AH + (uint32_t)(32 * it), pk0);
        tmem_st16(G.tm + TM_AH + (uint32_t)(32 * it) + 16u, pk);
    }
#else
    uint32_t va[32], vb[32], pk0[32];
    wait_half(G, 0);
    acc_ld(G, 0, va);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        uint32_t *cur = (c & 1) ? vb : va;
        uint32_t *nxt = (c & 1) ? va : vb;
        tmem_ld_wait();
        if (c == 1) wait_half(G, 1);                    // chunk 2 belongs to half 1; chunk 1's data is already in registers
        if (c < 3) acc_ld(G, (c + 1) * 32, nxt);
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float h0 = __uint_as_float(cur[4 * j]), h1 = __uint_as_float(cur[4 * j + 1]);
            float h2 = __uint_as_float(cur[4 * j + 2]), h3 = __uint_as_float(cur[4 * j + 3]);
            if (ACT) { h0 = softplus2_mix<0>(h0); h1 = softplus2_mix<1>(h1); h2 = softplus2_mix<2>(h2); h3 = softplus2_mix<3>(h3); }
            pk[2 * j] = pack_h2(h0, h1);
            pk[2 * j + 1] = pack_h2(h2, h3);
            if (do_alpha) {
                const float4 w = lds_f128(G.fb + 4u * (uint32_t)(FB_WA + c * 32 + 4 * j));
                s = fmaf(h0, w.x, s); s = fmaf(h1, w.y, s); s = fmaf(h2, w.z, s); s = fmaf(h3, w.w, s);
            }
        }
        if (c == 0) {
#pragma unroll
            for (int j = 0; j < 16; ++j) pk0[j] = pk[j];
        } else if (c == 1) {
            // half 1 has retired (waited above): nobody reads the old A_h any more
            tmem_st16(G.tm + TM_AH, pk0);
            tmem_st16(G.tm + TM_AH + 16u, pk);
        } else {
            tmem_st16(G.tm + TM_AH + (uint32_t)(c * 16), pk);
        }
    }
#endif
    G.phase ^= 1u;
    return s;
}

// views_linear epilogue: softplus2(acc[0..63]) . rgb_linear -> sigmoid
__device__ __forceinline__ void epi_views(Group &G, float (&rgb)[3]) {
    float r0 = 0.f, r1 = 0.f, r2 = 0.f;
    uint32_t va[32], vb[32];
    wait_half(G, 0);
    wait_half(G, 1);
    G.phase ^= 1u;
    acc_ld(G, 0, va);
    tmem_ld_wait();
    acc_ld(G, 32, vb);
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        uint32_t *cur = c ? vb : va;
        if (c) tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const float h = (j & 1) ? softplus2_mix<1>(__uint_as_float(cur[j])) : softplus2_mix<0>(__uint_as_float(cur[j]));
            const float4 w = lds_f128(G.fb + 4u * (uint32_t)(FB_WR + (c * 32 + j) * 4));
            r0 = fmaf(h, w.x, r0); r1 = fmaf(h, w.y, r1); r2 = fmaf(h, w.z, r2);
        }
    }
    rgb[0] = 1.0f / (1.0f + expf(-(r0 + lds_f32(G.fb + 4u * (FB_BR + 0)))));
    rgb[1] = 1.0f / (1.0f + expf(-(r1 + lds_f32(G.fb + 4u * (FB_BR + 1)))));
    rgb[2] = 1.0f / (1.0f + expf(-(r2 + lds_f32(G.fb + 4u * (FB_BR + 2)))));
}

__device__ float epi_hidden_shared(Group G, bool alpha);      // (defined after the kernel, next to mlp128_compact)

// The decoder MLP for the 128 samples of the group (features already in A_x, the constant tile in A_c): returns
// (sigma, r, g, b).  ONE out-of-line copy for every caller; an odd number of layers either way, so the caller flips
// G.phase once.
template <bool compact>
__device__ __forceinline__ float4 mlp128_body(Group &G, bool fine) {
    // pts_linears.0: K = 32 (27 features + the ones slot), N = 128 in two column halves
    start_layer(G, true, [&](int h) { issue_gemm(G, TM_AX, 2, OW0, 16384, 64, false, 64 * h); });
    if (compact) { epi_hidden_shared(G, false); G.phase ^= 1u; } else epi_hidden<0, true>(G);
    // pts_linears.1: K = 128 (+ bias through the constant tile)
    start_layer(G, true, [&](int h) {
        issue_gemm(G, TM_AH, 8, OW1, 16384, 64, false, 64 * h);
        issue_gemm(G, TM_AC, 1, OWB, 16384, 64, true, 64 * h);
    });
    if (compact) { epi_hidden_shared(G, false); G.phase ^= 1u; } else epi_hidden<0, true>(G);
    // pts_linears.2 on cat([x, h1])
    start_layer(G, true, [&](int h) {
        issue_gemm(G, TM_AX, 2, OW2X, 16384, 64, false, 64 * h);
        issue_gemm(G, TM_AH, 8, OW2H, 16384, 64, true, 64 * h);
    });
    float4 out;
    if (compact) { out.x = epi_hidden_shared(G, true); G.phase ^= 1u; } else out.x = epi_hidden<1, true>(G);
    out.x += lds_f32(G.fb + 4u * FB_BA);
    out.y = out.z = out.w = 0.f;
    if (fine) {
        // feature_linear (no activation), then views_linear on [feature | 1 | pe(d)]
        start_layer(G, true, [&](int h) {
            issue_gemm(G, TM_AH, 8, OWF, 16384, 64, false, 64 * h);
            issue_gemm(G, TM_AC, 1, OWB + 32, 16384, 64, true, 64 * h);
        });
        epi_hidden<0, false>(G);
        start_layer(G, false, [&](int) {
            issue_gemm(G, TM_AH, 8, OWV, 8192, 64, false);
            issue_gemm(G, TM_AC, 2, OWVP, 8192, 64, true);
        });
        float rgb[3];
        epi_views(G, rgb);
        out.y = rgb[0]; out.z = rgb[1]; out.w = rgb[2];
    }
    return out;
}
// two out-of-line copies, one per launch mode; each is called from its own instantiation of the kernel (COMPACT_MLP), so
// that the plain mode's code is what it was without the canonical one (a second call target at the three call sites, or a
// trampoline inside mlp128, cost the plain mode 0.9 - 1.3 %: profiles/r2_render_shared_epi_ab.log)
__device__ __noinline__ float4 mlp128(Group G, bool fine) { return mlp128_body<false>(G, fine); }
__device__ float4 mlp128_compact(Group G, bool fine);

// exclusive product scan over the group's 128 threads (thread order); *total = product of all 128 factors
__device__ __forceinline__ float excl_cumprod128(const Group &G, float f, float *total) {
    float v = f;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float up = __shfl_up_sync(0xffffffffu, v, o);
        if (G.lane >= o) v *= up;
    }
    const uint32_t red = G.sc + 4u * SC_RED;
    gbar(G);                                     // previous users of `red` are done
    if (G.lane == 31) sts_f32(red + 4u * (uint32_t)G.warp, v);
    gbar(G);
    const float4 r = lds_f128(red);
    const float pre = G.warp == 0 ? 1.f : G.warp == 1 ? r.x : G.warp == 2 ? r.x * r.y : (r.x * r.y) * r.z;
    float excl = __shfl_up_sync(0xffffffffu, v, 1);
    if (G.lane == 0) excl = 1.f;
    if (total) *total = ((r.x * r.y) * r.z) * r.w;
    return pre * excl;
}

__device__ __forceinline__ float group_sum(const Group &G, float v, int slot) {
    v = hl_warp_sum(v);
    const uint32_t red = G.sc + 4u * (uint32_t)(SC_RED + 8 + slot * 4);
    if (G.lane == 0) sts_f32(red + 4u * (uint32_t)G.warp, v);
    gbar(G);
    const float4 r = lds_f128(red);
    return ((r.x + r.y) + r.z) + r.w;
}

// Canonical-space mode, one sample per thread: SMPL-frame point q -> nearest vertex -> canonical point; the sample's own
// canonical view direction M sv goes, positionally encoded, into this thread's row of the constant tile
// A_c = [1 | pe(d) (27) | 0 (4)] (per ray in the other modes, per sample here: human_diffusion/NeRF/renderer.py:107-110,
// 155-157).  Out of line: its registers (and the sincos code) stay out of the render loop.
__device__ __noinline__ void canon_sample(const float4 *sph_s, float4 *stage, const float4 *__restrict__ verts,
                                          const float4 *__restrict__ aff, int NC, int CL, float qx, float qy, float qz,
                                          float svx, float svy, float svz, uint32_t tm_ac, float *pc,
                                          unsigned long long *prof) {
    float best;
    int v;
    hl_nearest_vertex_impl(sph_s, verts, stage, NC, CL, qx, qy, qz, 0, NC, best, v, prof);
    const float4 m0 = __ldg(aff + (size_t)v * 3), m1 = __ldg(aff + (size_t)v * 3 + 1), m2 = __ldg(aff + (size_t)v * 3 + 2);
    pc[0] = fmaf(m0.z, qz, fmaf(m0.y, qy, fmaf(m0.x, qx, m0.w)));
    pc[1] = fmaf(m1.z, qz, fmaf(m1.y, qy, fmaf(m1.x, qx, m1.w)));
    pc[2] = fmaf(m2.z, qz, fmaf(m2.y, qy, fmaf(m2.x, qx, m2.w)));
    const float dd[3] = {fmaf(m0.z, svz, fmaf(m0.y, svy, m0.x * svx)), fmaf(m1.z, svz, fmaf(m1.y, svy, m1.x * svx)),
                         fmaf(m2.z, svz, fmaf(m2.y, svy, m2.x * svx))};
    float h[32];
    h[0] = 1.0f;
#pragma unroll
    for (int k = 0; k < 27; ++k) {
        float val;
        if (k < 3) {
            val = dd[k];
        } else {
            const int f = (k - 3) / 3, comp = (k - 3) % 3;
            const float freq = (float)(1 << (f >> 1));
            const float phase = (f & 1) ? 1.5707963267948966f : 0.f;
            val = __sinf(__fadd_rn(phase, __fmul_rn(dd[comp], freq)));     // MUFU: the value is rounded to fp16 next
        }
        h[1 + k] = val;
    }
#pragma unroll
    for (int k = 28; k < 32; ++k) h[k] = 0.f;
    uint32_t ac[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) ac[k] = pack_h2(h[2 * k], h[2 * k + 1]);
    tmem_st16(tm_ac, ac);
}

// This file is compiled TWICE: as itself (the kernel k_render_tc5, the plain MLP copy, the whole C ABI) and through
// render_tc5_canon.cu with HL_R5_CANON_TU = 1 (the kernel k_render_tc5_canon for canonical-mode launches, which calls the
// compact MLP copy, + its launcher).  Two kernels, because any way of reaching both MLP copies from ONE kernel cost the
// plain mode 0.9 - 1.3 % (profiles/r2_render_shared_epi_ab.log); two translation units, because ptxas 12.9 segfaults on
// two entries that share this file's out-of-line device functions (and cicc on a templated kernel).
#if HL_R5_CANON_TU
#define HL_R5_KERNEL k_render_tc5_canon
constexpr bool COMPACT_MLP = true;
#else
#define HL_R5_KERNEL k_render_tc5
constexpr bool COMPACT_MLP = false;
#endif
__global__ void __launch_bounds__(NT5, 1) HL_R5_KERNEL(const Render5Args a) {
    const bool CANON = a.canon != 0;
    extern __shared__ __align__(16) uint8_t smraw5[];
    __shared__ __align__(8) uint64_t mbars[2 * GROUPS];      // per group: column half 0 / half 1 of the layer in flight
    __shared__ uint32_t tmem_slot;
    __shared__ float bnd[8];
    const uint32_t base = (smem_u32(smraw5) + 1023u) & ~1023u;
    uint8_t *aligned = smraw5 + (base - smem_u32(smraw5));
    const int tid = threadIdx.x;

    // one-time: MLP image (fp16 atoms + fp32 table) -> shared memory, TMEM allocation, barriers
    {
        uint4 *dst = reinterpret_cast<uint4 *>(aligned);
        for (int i = tid; i < (W_BYTES + 4 * FB_FLOATS) / 16; i += NT5) dst[i] = __ldg(a.w16s + i);
    }
    // canonical-space search scratch (only allocated by canon launches): spheres | one staging row per warp
    float4 *canon_f4 = reinterpret_cast<float4 *>(aligned + W_BYTES + 4 * (FB_FLOATS + GROUPS * SC_FLOATS));
    float4 *canon_stage = canon_f4 + HL_CANON_NC_MAX + (tid >> 5) * HL_CANON_CL_MAX;
    if (CANON) hl_canon_stage_spheres(a.ct, canon_f4);
    if (tid < 6) bnd[tid] = a.bounds_dev ? __ldg(a.bounds_dev + tid) : a.bounds[tid];
    if (tid < 32) {
        if (tid == 0) {
            for (int g = 0; g < 2 * GROUPS; ++g) mbar_init(smem_u32(&mbars[g]), 1);
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
    G.mbar = smem_u32(&mbars[2 * G.g]);
    G.phase = 0;
    G.w_smem = base;
    G.fb = base + W_BYTES;
    G.sc = G.fb + 4u * (uint32_t)(FB_FLOATS + G.g * SC_FLOATS);
    const uint32_t zc = G.sc + 4u * SC_ZC, zn = G.sc + 4u * SC_ZN, zf = G.sc + 4u * SC_ZF, cdf = G.sc + 4u * SC_CDF,
                   bins = G.sc + 4u * SC_BINS, pe = G.sc + 4u * SC_PE;
    const int tg = G.tg;
    const uint32_t tg4 = 4u * (uint32_t)tg;

    const bool prof = a.prof != nullptr && blockIdx.x == 0 && tid == 0;
    G.prof = prof ? a.prof : nullptr;
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
        {   // the constant tile: [1 | 0 ...] (no view direction here; only the bias of pts_linears.1 uses it)
            uint32_t ac[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) ac[k] = 0u;
            ac[0] = pack_h2(1.0f, 0.f);
            tmem_st16(G.tm + TM_AC, ac);
        }
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
            const float4 r = (COMPACT_MLP ? mlp128_compact(G, false) : mlp128(G, false));
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
            sts_f32(zc + tg4, zmine);
            if (tg < 32) {                            // positional encoding of the view direction (fields.py:69-85)
                const float dd[3] = {dx / dnorm, dy / dnorm, dz / dnorm};
                float v = 0.f;
                if (tg < 3) {
                    v = dd[tg];
                } else if (tg < 27) {
                    const int f = (tg - 3) / 3, comp = (tg - 3) % 3;
                    const float freq = (float)(1 << (f >> 1));
                    const float phase = (f & 1) ? 1.5707963267948966f : 0.f;
                    v = sinf(__fadd_rn(phase, __fmul_rn(dd[comp], freq)));
                }
                sts_f32(pe + tg4, v);
            }
            gbar(G);
            float sv[3] = {0.f, 0.f, 0.f};
            if (CANON) {   // smpl_viewdir = (viewdir - Th) R  (renderer.py:125 subtracts Th from the direction too)
                hl_to_smpl_frame(a.ct, dx / dnorm, dy / dnorm, dz / dnorm, sv[0], sv[1], sv[2]);
            } else {       // the constant tile of this ray: [1 | pe(27) | 0(4)] in every row
                uint32_t ac[16];
                float prev = 1.0f;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float4 q = lds_f128(pe + 16u * (uint32_t)k);
                    ac[2 * k] = pack_h2(prev, q.x);
                    ac[2 * k + 1] = pack_h2(q.y, q.z);
                    prev = q.w;
                }
                tmem_st16(G.tm + TM_AC, ac);          // (pe[27..31] = 0, so halves 28..31 are zero)
            }
            RPROF(0)
            float rgb0[3] = {0.f, 0.f, 0.f}, rgb1[3] = {0.f, 0.f, 0.f}, sig0 = 0.f, sig1 = 0.f;
            // ------------------------------- coarse samples: the WHOLE network, once ---------------------------
            // The reference evaluates NeRF_network twice on the coarse points: density only for up_sample
            // (renderer.py:258-264), then again -- same points, same weights -- inside render_core on the sorted
            // union of coarse and new samples (renderer.py:266-279,199-201).  The second evaluation reproduces the
            // first bit for bit, so the colour branch runs right here on the activations still sitting in tensor
            // memory and the fine pass only evaluates the 128 NEW samples: 10 layer passes per ray instead of 13,
            // two gathers instead of three.
            {
                float pc[3] = {__fadd_rn(ox, __fmul_rn(dx, zmine)), __fadd_rn(oy, __fmul_rn(dy, zmine)),
                               __fadd_rn(oz, __fmul_rn(dz, zmine))};
                if (CANON) {
                    float qx, qy, qz;
                    hl_to_smpl_frame(a.ct, pc[0], pc[1], pc[2], qx, qy, qz);
                    canon_sample(canon_f4, canon_stage, a.ct.verts, a.ct.aff, a.ct.NC, a.ct.CL, qx, qy, qz, sv[0], sv[1], sv[2],
                                 G.tm + TM_AC, pc, prof ? a.prof + 8 : nullptr);
                }
                gather_to_tmem(a.tex, a.R, bnd, pc[0], pc[1], pc[2], G.tm + TM_AX);
            }
            RPROF(1)
            const float4 rc = (COMPACT_MLP ? mlp128_compact(G, true) : mlp128(G, true));
            G.phase ^= 1u;
            RPROF(2)
            if (a.n_importance) {
                const float sigma = rc.x;
                // ------------------------------- up_sample + sample_pdf -----------------------------------
                const float znext = tg < NS - 1 ? lds_f32(zc + tg4 + 4u) : 0.f;
                const float dist = (tg < NS - 1 ? znext - zmine : 1e10f) * dnorm;
                const float al = 1.0f - expf(-softplus_acc(sigma) * dist);
                if (tg < NS - 1) sts_f32(bins + tg4, 0.5f * (znext + zmine));
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
                    const uint32_t red = G.sc + 4u * (SC_RED + 16);
                    if (G.lane == 31) sts_f32(red + 4u * (uint32_t)G.warp, v);
                    gbar(G);
                    const float4 r = lds_f128(red);
                    const float pre = G.warp == 0 ? 0.f : G.warp == 1 ? r.x : G.warp == 2 ? r.x + r.y : (r.x + r.y) + r.z;
                    if (tg <= NS - 2) sts_f32(cdf + tg4, pre + v);        // cdf[0] = 0, cdf[i] = pdf[1] + ... + pdf[i]
                }
                gbar(G);
                float znew;
                int pos_c, pos_n;
                {
                    const float uu = a.u ? a.u[ray * NS + tg] : uniform_hash(a.seed, (unsigned long long)ray, tg);
                    int lo = 0, hi = NS - 1;
#pragma unroll
                    for (int it = 0; it < 7; ++it) {          // 127 entries: 7 halvings always terminate
                        const int mid = (lo + hi) >> 1;
                        const bool gt = lds_f32(cdf + 4u * (uint32_t)mid) > uu;
                        if (lo < hi) { if (gt) hi = mid; else lo = mid + 1; }
                    }
                    const int below = max(lo - 1, 0), above = min(NS - 2, lo);
                    const float cb = lds_f32(cdf + 4u * (uint32_t)below), ca = lds_f32(cdf + 4u * (uint32_t)above);
                    float den = ca - cb;
                    if (den < 1e-5f) den = 1.0f;
                    const float t = (uu - cb) / den;
                    const float bb = lds_f32(bins + 4u * (uint32_t)below), ba = lds_f32(bins + 4u * (uint32_t)above);
                    znew = bb + t * (ba - bb);
                    sts_f32(zn + tg4, znew);
                }
                gbar(G);
                {   // sort(cat(z, z_new)) by ranking: coarse z is already sorted
                    int c_lt = 0, n_lt = 0;
#pragma unroll 4
                    for (int j = 0; j < NS / 4; ++j) {
                        const float4 q = lds_f128(zn + 16u * (uint32_t)j);
                        c_lt += (q.x < zmine) + (q.y < zmine) + (q.z < zmine) + (q.w < zmine);
                        const int j4 = 4 * j;
                        n_lt += ((q.x < znew) || (q.x == znew && j4 < tg)) + ((q.y < znew) || (q.y == znew && j4 + 1 < tg)) +
                                ((q.z < znew) || (q.z == znew && j4 + 2 < tg)) + ((q.w < znew) || (q.w == znew && j4 + 3 < tg));
                    }
                    int lo = 0, hi = NS;                   // number of coarse z <= znew
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        const int mid = (lo + hi) >> 1;
                        const bool le = lds_f32(zc + 4u * (uint32_t)min(mid, NS - 1)) <= znew;
                        if (lo < hi) { if (le) lo = mid + 1; else hi = mid; }
                    }
                    pos_c = tg + c_lt;                     // sorted position of this thread's coarse / new sample
                    pos_n = n_lt + lo;
                    sts_f32(zf + 4u * (uint32_t)pos_c, zmine);
                    sts_f32(zf + 4u * (uint32_t)pos_n, znew);
                    if (CANON) {
                        // Hand the new samples out in depth order (thread tg takes the tg-th smallest): the uniforms are
                        // unsorted, so a warp's 32 samples would otherwise be scattered over the whole ray and the
                        // nearest-vertex search -- a warp scans the union of the clusters its lanes need -- would touch
                        // four times as many clusters.  The set of (position, value) pairs is unchanged.
                        gbar(G);                           // everyone has finished reading zn
                        sts_f32(zn + 4u * (uint32_t)n_lt, znew);
                        gbar(G);
                        znew = lds_f32(zn + tg4);
                        int l2 = 0, h2 = NS;               // number of coarse z <= znew
#pragma unroll
                        for (int it = 0; it < 8; ++it) {
                            const int mid = (l2 + h2) >> 1;
                            const bool le = lds_f32(zc + 4u * (uint32_t)min(mid, NS - 1)) <= znew;
                            if (l2 < h2) { if (le) l2 = mid + 1; else h2 = mid; }
                        }
                        pos_n = tg + l2;
                    }
                }
                RPROF(3)
                // ------------------------------- fine pass: the 128 new samples ----------------------------
                {
                    float pc[3] = {__fadd_rn(ox, __fmul_rn(dx, znew)), __fadd_rn(oy, __fmul_rn(dy, znew)),
                                   __fadd_rn(oz, __fmul_rn(dz, znew))};
                    if (CANON) {
                        float qx, qy, qz;
                        hl_to_smpl_frame(a.ct, pc[0], pc[1], pc[2], qx, qy, qz);
                        canon_sample(canon_f4, canon_stage, a.ct.verts, a.ct.aff, a.ct.NC, a.ct.CL, qx, qy, qz, sv[0], sv[1], sv[2],
                                     G.tm + TM_AC, pc, prof ? a.prof + 8 : nullptr);
                    }
                    gather_to_tmem(a.tex, a.R, bnd, pc[0], pc[1], pc[2], G.tm + TM_AX);
                }
                RPROF(1)
                const float4 rn = (COMPACT_MLP ? mlp128_compact(G, true) : mlp128(G, true));
                G.phase ^= 1u;
                RPROF(2)
                // both sets of (sigma, r, g, b) into sorted order; this thread composites samples tg and 128 + tg
                const uint32_t val = G.sc + 4u * SC_VAL;
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(val + 16u * (uint32_t)pos_c), "f"(rc.x), "f"(rc.y),
                             "f"(rc.z), "f"(rc.w) : "memory");
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(val + 16u * (uint32_t)pos_n), "f"(rn.x), "f"(rn.y),
                             "f"(rn.z), "f"(rn.w) : "memory");
                gbar(G);                                   // also orders the sorted z (zf) before the reads below
                const float4 v0 = lds_f128(val + 4u * tg4), v1 = lds_f128(val + 16u * (uint32_t)(NS + tg));
                sig0 = v0.x; rgb0[0] = v0.y; rgb0[1] = v0.z; rgb0[2] = v0.w;
                sig1 = v1.x; rgb1[0] = v1.y; rgb1[1] = v1.z; rgb1[2] = v1.w;
            } else {
                sts_f32(zf + tg4, zmine);
                gbar(G);
                sig0 = rc.x; rgb0[0] = rc.y; rgb0[1] = rc.z; rgb0[2] = rc.w;
            }
            const float z0 = lds_f32(zf + tg4), z1 = a.n_importance ? lds_f32(zf + 4u * NS + tg4) : 0.f;
            // ------------------------------- composite (renderer.py:222-239) --------------------------
            {
                const int last = n_tiles * NS - 1;
                const float d0 = tg < last ? lds_f32(zf + tg4 + 4u) - z0 : 1e10f;                  // NOT scaled by |d|
                const float al0 = 1.0f - expf(-softplus_acc(sig0) * d0);
                float P0;
                const float T0 = excl_cumprod128(G, (1.0f - al0) + 1e-7f, &P0);
                const float w0 = al0 * T0;
                float w1 = 0.f;
                if (n_tiles == 2) {
                    const float d1 = NS + tg < last ? lds_f32(zf + 4u * NS + tg4 + 4u) - z1 : 1e10f;
                    const float al1 = 1.0f - expf(-softplus_acc(sig1) * d1);
                    const float T1 = P0 * excl_cumprod128(G, (1.0f - al1) + 1e-7f, nullptr);
                    w1 = al1 * T1;
                }
                float q5[5] = {w0 + w1, w0 * rgb0[0] + w1 * rgb1[0], w0 * rgb0[1] + w1 * rgb1[1],
                               w0 * rgb0[2] + w1 * rgb1[2], w0 * z0 + w1 * z1};
#pragma unroll
                for (int k = 0; k < 5; ++k) q5[k] = hl_warp_sum(q5[k]);
                const uint32_t red = G.sc + 4u * (SC_RED + 24);          // [5][4]
                gbar(G);
                if (G.lane == 0) {
#pragma unroll
                    for (int k = 0; k < 5; ++k) sts_f32(red + 4u * (uint32_t)(k * 4 + G.warp), q5[k]);
                }
                gbar(G);
                if (tg == 0) {
                    float s[5];
#pragma unroll
                    for (int k = 0; k < 5; ++k) {
                        const float4 r = lds_f128(red + 16u * (uint32_t)k);
                        s[k] = ((r.x + r.y) + r.z) + r.w;
                    }
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

// ONE out-of-line copy of the activated hidden epilogue for its three call sites (pts_linears.0 / .1 / .2; the alpha head
// of .2 decided at run time), used in canonical mode:
// the fully unrolled epilogues are what the kernel's instruction footprint consists of (ncu: instruction requests at the
// GPC-level cache at 51 % of its peak, 72 % in canonical mode, where the search and the per-sample encoding compete for
// the instruction caches and "no instruction" is the top stall), and a copy less is 1.1 k instructions less to stream per
// MLP evaluation.  Same-box A/B (profiles/r2_render_shared_epi_ab.log): canonical 61.8 -> 58.2 ms per frame, the plain
// mode 27.58 -> 28.07 ms (the call costs more than the fetches it saves) -- hence chosen per launch mode.
__device__ __noinline__ float epi_hidden_shared(Group G, bool alpha) { return epi_hidden<2, true>(G, alpha); }
// (the canonical-mode copies are placed after the kernel so that the plain mode's code layout stays what it was without them)
__device__ __noinline__ float4 mlp128_compact(Group G, bool fine) { return mlp128_body<true>(G, fine); }

// planes [3][9][R][R] fp32 (channel = sub * 3 + ch) -> quad texels [9][R + 1][R + 1] x 16 halves (see gather_subplanes):
// entry (yq, xq) = the footprint whose top-left tap is (yq - 1, xq - 1).
__global__ void k_triplane_to_quads(const float *__restrict__ planes, uint4 *__restrict__ out, int R) {
    const int R1 = R + 1;
    const size_t n = (size_t)9 * R1 * R1;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int xq = (int)(i % R1), yq = (int)((i / R1) % R1), c = (int)(i / ((size_t)R1 * R1));
        const int plane = c / 3, sub = c % 3;
        const float *src = planes + ((size_t)plane * 9 + sub * 3) * R * R;
        __half h[16];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int y = yq - 1 + (k >> 1), x = xq - 1 + (k & 1);
            const bool in = y >= 0 && y < R && x >= 0 && x < R;
#pragma unroll
            for (int ch = 0; ch < 3; ++ch)
                h[3 * k + ch] = __float2half_rn(in ? src[(size_t)ch * R * R + (size_t)y * R + x] : 0.f);
        }
#pragma unroll
        for (int k = 12; k < 16; ++k) h[k] = __float2half_rn(0.f);
        const uint4 *hv = reinterpret_cast<const uint4 *>(h);
        out[2 * i] = hv[0];
        out[2 * i + 1] = hv[1];
    }
}

unsigned long long *g_prof5 = nullptr;

}  // namespace

#if HL_R5_CANON_TU
namespace { constexpr size_t SMEM5_CANON = SMEM5 + 16 * (size_t)HL_CANON_SMEM_F4; }
// the canonical-mode kernel's launcher (called by launch5 of the main translation unit; `args` = its Render5Args)
int hl_r5_launch_canon(const void *args, int grid, void *stream) {
    static HlPerDeviceOnce once;
    if (once.need()) {
        HL_CHECK_CUDA(cudaFuncSetAttribute(k_render_tc5_canon, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM5_CANON));
    }
    k_render_tc5_canon<<<grid, NT5, SMEM5_CANON, (cudaStream_t)stream>>>(*reinterpret_cast<const Render5Args *>(args));
    HL_CHECK_LAUNCH();
    return HL_OK;
}
#else
int hl_r5_launch_canon(const void *args, int grid, void *stream);      // render_tc5_canon.cu

namespace {
int launch5(Render5Args &a, long long units, cudaStream_t stream, bool canon = false) {
    static HlPerDeviceOnce once;
    if (once.need()) {
        HL_CHECK_CUDA(cudaFuncSetAttribute(k_render_tc5, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM5));
    }
    long long grid = hl_num_sms();
    const long long need = (units + GROUPS - 1) / GROUPS;
    if (grid > need) grid = need;
    a.canon = canon ? 1 : 0;
    if (canon) return hl_r5_launch_canon(&a, (int)grid, stream);
    k_render_tc5<<<(int)grid, NT5, SMEM5, stream>>>(a);
    HL_CHECK_LAUNCH();
    return HL_OK;
}

}  // namespace

extern "C" int hl_triplane_to_quads(const float *planes, void *quads, int R, void *stream) {
    HL_CHECK_ARG(planes && quads && R > 0 && ((uintptr_t)quads & 31) == 0);
    const size_t n = (size_t)9 * (R + 1) * (R + 1);
    int grid = (int)((n + 255) / 256);
    const int cap = hl_num_sms() * 8;
    if (grid > cap) grid = cap;
    k_triplane_to_quads<<<grid, 256, 0, (cudaStream_t)stream>>>(planes, reinterpret_cast<uint4 *>(quads), R);
    HL_CHECK_LAUNCH();
    return HL_OK;
}

extern "C" int hl_render5_set_profile(void *dev_counters) {
    g_prof5 = (unsigned long long *)dev_counters;
    return HL_OK;
}

extern "C" int hl_render_rays_tc5(const void *texels, int R, const void *mlp_tc5,
                                  const float *rays_o, const float *rays_d, const float *near, const float *far,
                                  const float *z_coarse, const float *u, uint64_t seed, const float *bounds,
                                  int bounds_on_device, float *rgb, float *acc, float *depth, int64_t n_rays,
                                  int n_importance, int clamp_depth, void *stream) {
    HL_CHECK_ARG(texels && mlp_tc5 && rays_o && rays_d && near && far && bounds && rgb && acc && depth);
    HL_CHECK_ARG(R > 0 && n_rays > 0 && ((uintptr_t)texels & 31) == 0 && ((uintptr_t)mlp_tc5 & 15) == 0);
    HL_CHECK_ARG(n_importance == 0 || n_importance == NS);
    Render5Args a = {};
    a.tex = reinterpret_cast<const uint4 *>(texels);
    a.R = R;
    a.w16s = reinterpret_cast<const uint4 *>(mlp_tc5);
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

int hl_set_canon_tables(CanonTables &t, const float *knn_table, const float *affine_table, int n_clusters, int cluster_slots,
                        const float *rot, const float *trans);      // render.cu

extern "C" int hl_render_rays_tc5_canon(const void *texels, int R, const void *mlp_tc5, const float *rays_o,
                                        const float *rays_d, const float *near, const float *far, const float *z_coarse,
                                        const float *u, uint64_t seed, const float *t_bounds, const float *knn_table,
                                        const float *affine_table, int n_clusters, int cluster_slots, const float *rot,
                                        const float *trans, float *rgb, float *acc, float *depth, int64_t n_rays,
                                        int n_importance, int clamp_depth, void *stream) {
    HL_CHECK_ARG(texels && mlp_tc5 && rays_o && rays_d && near && far && t_bounds && rgb && acc && depth);
    HL_CHECK_ARG(R > 0 && n_rays > 0 && ((uintptr_t)texels & 31) == 0 && ((uintptr_t)mlp_tc5 & 15) == 0);
    HL_CHECK_ARG(n_importance == 0 || n_importance == NS);
    Render5Args a = {};
    a.tex = reinterpret_cast<const uint4 *>(texels);
    a.R = R;
    a.w16s = reinterpret_cast<const uint4 *>(mlp_tc5);
    a.o = rays_o; a.d = rays_d; a.near = near; a.far = far; a.u = u; a.zc_in = z_coarse;
    a.seed = seed;
    for (int i = 0; i < 6; ++i) a.bounds[i] = t_bounds[i];
    a.rgb = rgb; a.acc = acc; a.depth = depth;
    a.n_rays = n_rays;
    a.clamp_depth = clamp_depth;
    a.n_importance = n_importance;
    a.prof = g_prof5;
    if (int rc = hl_set_canon_tables(a.ct, knn_table, affine_table, n_clusters, cluster_slots, rot, trans)) return rc;
    return launch5(a, n_rays, (cudaStream_t)stream, true);
}

extern "C" int hl_density_grid_tc5(const void *texels, int R, const void *mlp_tc5,
                                   const float *bounds, int bounds_on_device, int resolution, float *out, void *stream) {
    HL_CHECK_ARG(texels && mlp_tc5 && bounds && out && R > 0 && resolution >= 2);
    HL_CHECK_ARG(((uintptr_t)texels & 31) == 0 && ((uintptr_t)mlp_tc5 & 15) == 0);
    Render5Args a = {};
    a.tex = reinterpret_cast<const uint4 *>(texels);
    a.R = R;
    a.w16s = reinterpret_cast<const uint4 *>(mlp_tc5);
    if (bounds_on_device) a.bounds_dev = bounds;
    else for (int i = 0; i < 6; ++i) a.bounds[i] = bounds[i];
    a.grid_res = resolution;
    a.grid_out = out;
    a.n_importance = NS;
    const long long tiles = ((long long)resolution * resolution * resolution + 127) / 128;
    return launch5(a, tiles, (cudaStream_t)stream);
}
#endif      // !HL_R5_CANON_TU
