// GroupNorm-1 + SiLU operand pass FUSED with the ResBlock's 1x1 skip convolution (unet.py:198-201,208-219), sm_100a.
//
//   act  = fp16( SiLU( GroupNorm32(x) ) )                       the operand of in_layers' 3x3 conv
//   skip = W_skip . x + b                                       x + h -> skip_connection(x) + h  (unet.py:219)
//
// The two-kernel form reads x twice in effect: k_gn_apply writes, next to `act`, a scaled fp16 hi | lo copy of x (4 B per
// element) that the tcgen05 conv reads back for its three operand passes (hi.W_hi + lo.W_hi + hi.W_lo).  For
// the decoder's concat inputs at 256^2 / 128^2 that copy is 0.8 GB per block written and read again.  Here x is read ONCE
// (TMA, fp32), every pixel row is owned by one thread that
//   * applies the per-(sample, channel) affine + SiLU and writes its 128 B of `act` into a staging tile that leaves by
//     TMA store (per-thread 16-byte global stores were measured L2-request-bound: 2.5 TB/s),
//   * splits x * 2^-4 into fp16 hi + lo and writes both as A operands into TENSOR MEMORY (tcgen05.st; double-buffered, so
//     the split of chunk c + 1 overlaps the MMAs of chunk c -- with the operand tiles in shared memory there was room for
//     one buffer and a two-deep weight ring only, and the kernel sat at 3.2 TB/s),
// and the skip conv's MMAs (tcgen05 kind::f16 in the .ts form: A from tensor memory, B = W_hi / W_lo tiles streamed by TMA
// into a three-deep ring, fp32 accumulators in tensor memory) run on them.  HBM bytes per pixel: 4 Cin in, 2 Cin + 4 Cout
// out -- half of the two-kernel form.  The arithmetic is the two-kernel form's: `act` equals k_gn_apply_f16's (same
// coefficient formulas and reduction order; a few elements per million land one fp16 ulp away), `skip` differs only in
// fp32 accumulation order (chunk-major instead of pass-major).
//
//   tile        128 consecutive pixels of one sample (H*W % 128 == 0), all Cout channels (Cout <= 256, or 2 x 192)
//   warps 0-7   compute: thread = (pixel i of the tile = TMEM lane i, one 32-channel half of the 64-channel chunk): x row
//               from the X ring -> act into the staging tile, hi / lo operand halves -> tensor memory -> arrive on the
//               MMA warp's barrier; one of them issues the act tile's TMA store
//   warp 8      X producer: two TMA boxes {32 ch, 128 px} fp32 per chunk into a ring of 3 slots (the bytes in flight)
//   warp 9      W producer: per chunk the W_hi tiles of every n tile, then the W_lo tiles ({64 ch, nt rows} fp16)
//   warp 10     MMA issuer: per chunk and n tile  acc += x_hi.W_hi + x_lo.W_hi, then acc += x_hi.W_lo
//   warps 11-14 epilogue: accumulator row + bias -> staging tile -> TMA store of `skip`; two accumulator tiles when
//               Cout <= 192, so the epilogue of tile t overlaps the MMAs of tile t + 1
// Measured on B200 (tools/probe_gn_skip.py, 384 -> 192 at 256^2, B = 4; two launches: 295 us):
//   v0  4 compute warps, per-thread 16-byte global stores, MMAs issued by warp 0, operands in shared memory     323 us
//   v1  outputs through staging tiles + TMA stores (the direct stores were L2-request-bound)                      266 us
//   v2  8 compute warps (two per scheduler)                                                                       248 us
//   v3  operands in tensor memory (double-buffered), three-deep weight ring                                       243 us
//   v4  dedicated MMA-issue warp (issue proceeds at the pace of execution: 1.7 k cycles per chunk of a compute
//       warp's time), coefficient refresh on 8 lanes per group                                                    205 us
//   v5  dedicated epilogue warps, two accumulator tiles                                                           170 us = 4.7 TB/s
#include "common.cuh"
#include "tc5.cuh"

#include <cuda.h>
#include <cuda_fp16.h>

int hl_num_sms();

namespace {

constexpr int GS_ROWS = 128;
constexpr int GS_XSLOT = 32 * 1024;      // one 64-channel fp32 chunk of 128 pixels = two 16 KB boxes
constexpr int GS_ATILE = 16 * 1024;      // one fp16 operand tile [128 rows][128 B]
constexpr int GS_MAX_C = 1536;
constexpr int GS_COMPUTE = 256;         // 8 compute warps: two threads per pixel row
constexpr int GS_EPI = 128;             // 4 epilogue warps (one thread per TMEM lane)
constexpr int GS_THREADS = GS_COMPUTE + 96 + GS_EPI;      // + X producer, W producer, MMA issuer, epilogue
constexpr int GS_MAX_NX = 4, GS_MAX_NW = 3;
constexpr uint32_t GS_TM_A = 384;        // tensor memory: accumulators in columns [0, Cout <= 384), two A buffers (hi 32 | lo 32) behind

struct GsParams {
    int Cin, Cout, kchunks, nt, n_tiles, nx, nw, HW, total_tiles, groups, stats_ld, ld_act, ld_skip, tmem_cols;
    int acc_bufs;                 // 2 when two accumulator tiles fit in front of the A buffers (Cout <= 192): the epilogue of
                                  // tile t overlaps the MMAs of tile t + 1
    float eps;
    const double *stats;
    const float *gamma, *beta, *bias;
    __half *act;
    float *skip;
    unsigned long long *prof;     // optional cycle counters of (CTA 0, thread 0): hl_gn_skip_set_profile
};

__device__ __forceinline__ uint32_t gs_h2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t *>(&h);
}
__device__ __forceinline__ uint32_t gs_lo(float a, float b, uint32_t hi) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&hi));
    return gs_h2(a - f.x, b - f.y);
}
__device__ __forceinline__ uint64_t gs_desc(uint32_t saddr) {       // K-major SWIZZLE_128B, 8-row groups 1024 B apart
    const uint32_t hi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
    return ((uint64_t)hi << 32) | (((saddr & 0x3FFFFu) >> 4) | (1u << 16));
}

__global__ void __launch_bounds__(GS_THREADS, 1)
k_gn_skip_tc5(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
              const __grid_constant__ CUtensorMap tmAct, const __grid_constant__ CUtensorMap tmSkip, const GsParams p) {
    extern __shared__ uint8_t gs_smem[];
    __shared__ __align__(8) uint64_t bars[2 * GS_MAX_NX + 2 * GS_MAX_NW + 8];
    __shared__ uint32_t tmem_slot;
    __shared__ float gmean[64], grstd[64];

    const int warp = threadIdx.x >> 5;
    const uint32_t base = (smem_u32(gs_smem) + 1023u) & ~1023u;
    const uint32_t sm_x = base;
    const uint32_t sm_w = sm_x + (uint32_t)p.nx * GS_XSLOT;
    const uint32_t w_bytes = (uint32_t)p.nt * 128u;
    const uint32_t sm_o = sm_w + (uint32_t)p.nw * w_bytes;                  // two act staging tiles [128 rows][128 B] + one for skip
    const uint32_t sm_s = sm_o + 2u * GS_ATILE;
    float *coef = reinterpret_cast<float *>(gs_smem + (sm_s + GS_ATILE - smem_u32(gs_smem)));   // ca[Cin] | cb[Cin]
    const uint32_t bar_x_full = smem_u32(&bars[0]), bar_x_empty = smem_u32(&bars[GS_MAX_NX]);
    const uint32_t bar_w_full = smem_u32(&bars[2 * GS_MAX_NX]), bar_w_empty = smem_u32(&bars[2 * GS_MAX_NX + GS_MAX_NW]);
    const uint32_t bar_a_empty = smem_u32(&bars[2 * GS_MAX_NX + 2 * GS_MAX_NW]);
    const uint32_t bar_a_full = smem_u32(&bars[2 * GS_MAX_NX + 2 * GS_MAX_NW + 2]);
    const uint32_t bar_acc_full = smem_u32(&bars[2 * GS_MAX_NX + 2 * GS_MAX_NW + 4]);
    const uint32_t bar_acc_free = smem_u32(&bars[2 * GS_MAX_NX + 2 * GS_MAX_NW + 6]);
    const uint32_t acc_stride = (uint32_t)(p.nt * p.n_tiles);

    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmAct) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmSkip) : "memory");
        for (int s = 0; s < GS_MAX_NX; ++s) { mbar_init(bar_x_full + 8 * s, 1); mbar_init(bar_x_empty + 8 * s, GS_COMPUTE); }
        for (int s = 0; s < GS_MAX_NW; ++s) { mbar_init(bar_w_full + 8 * s, 1); mbar_init(bar_w_empty + 8 * s, 1); }
        mbar_init(bar_a_empty, 1);
        mbar_init(bar_a_empty + 8, 1);
        mbar_init(bar_a_full, GS_COMPUTE);
        mbar_init(bar_a_full + 8, GS_COMPUTE);
        mbar_init(bar_acc_full, 1);
        mbar_init(bar_acc_full + 8, 1);
        mbar_init(bar_acc_free, GS_EPI);
        mbar_init(bar_acc_free + 8, GS_EPI);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                     "r"((uint32_t)p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    hl_pdl_trigger_early();
    tc_fence_before();
    __syncthreads();
    hl_pdl_wait();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const int n_w = 2 * p.n_tiles;               // weight tiles per chunk: W_hi of every n tile, then W_lo

    if (warp == 8) {
        // ---------------------------------- X producer ----------------------------------
        uint32_t s = 0, ph = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
            for (int c = 0; c < p.kchunks; ++c) {
                mbar_wait(bar_x_empty + 8 * s, ph ^ 1u);
                if (elect_one_sync()) {
                    mbar_expect_tx(bar_x_full + 8 * s, GS_XSLOT);
                    tma_load_2d(sm_x + s * GS_XSLOT, &tmX, bar_x_full + 8 * s, c * 64, tile * GS_ROWS);
                    tma_load_2d(sm_x + s * GS_XSLOT + GS_XSLOT / 2, &tmX, bar_x_full + 8 * s, c * 64 + 32, tile * GS_ROWS);
                }
                __syncwarp();
                if (++s == (uint32_t)p.nx) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 9) {
        // ---------------------------------- W producer ----------------------------------
        uint32_t s = 0, ph = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
            for (int c = 0; c < p.kchunks; ++c) {
                for (int wi = 0; wi < n_w; ++wi) {
                    const int slab = wi / p.n_tiles, nti = wi - slab * p.n_tiles;
                    mbar_wait(bar_w_empty + 8 * s, ph ^ 1u);
                    if (elect_one_sync()) {
                        mbar_expect_tx(bar_w_full + 8 * s, w_bytes);
                        tma_load_3d(sm_w + s * w_bytes, &tmW, bar_w_full + 8 * s, c * 64, nti * p.nt, slab);
                    }
                    __syncwarp();
                    if (++s == (uint32_t)p.nw) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 10) {
        // ---------------------------------- MMA issuer ----------------------------------
        // A dedicated warp: the issue of a chunk's 12 MMAs proceeds at the pace they EXECUTE (measured 1.7 k cycles per chunk
        // when a compute warp issued them, stalling all eight at the next barrier).  Per n tile:
        //   acc += x_hi.W_hi + x_lo.W_hi   (slab 0),   acc += x_hi.W_lo   (slab 1)
        const uint32_t idesc = (1u << 4) | ((uint32_t)(p.nt >> 3) << 17) | ((uint32_t)(GS_ROWS >> 4) << 24);   // f16 x f16 -> f32, K-major
        uint32_t swr = 0, phw = 0, g = 0, it = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
            // accumulator tile of this output tile; its previous contents must have left tensor memory
            const uint32_t buf = p.acc_bufs == 2 ? (it & 1u) : 0u, use = it / (uint32_t)p.acc_bufs;
            if (use) mbar_wait(bar_acc_free + 8 * buf, (use - 1u) & 1u);
            tc_fence_after();
            for (int c = 0; c < p.kchunks; ++c, ++g) {
                const uint32_t ab = g & 1u;
                mbar_wait(bar_a_full + 8 * ab, (g >> 1) & 1u);
                tc_fence_after();
                if (elect_one_sync()) {
                    const uint32_t ah = tmem + GS_TM_A + ab * 64u, al = ah + 32u;      // A_hi / A_lo: 64 halves = 32 columns each
                    for (int wi = 0; wi < n_w; ++wi) {
                        const int slab = wi / p.n_tiles, nti = wi - slab * p.n_tiles;
                        mbar_wait(bar_w_full + 8 * swr, phw);
                        tc_fence_after();
                        const uint64_t bd = gs_desc(sm_w + swr * w_bytes);
                        const uint32_t d = tmem + buf * acc_stride + (uint32_t)(nti * p.nt);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_ts_f16(d, ah + 8u * (uint32_t)k, bd + 2 * k, idesc, (c | slab | k) != 0 ? 1u : 0u);
                        if (!slab) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) umma_ts_f16(d, al + 8u * (uint32_t)k, bd + 2 * k, idesc, 1u);
                        }
                        umma_commit(bar_w_empty + 8 * swr);
                        if (++swr == (uint32_t)p.nw) { swr = 0; phw ^= 1u; }
                    }
                    umma_commit(bar_a_empty + 8 * ab);
                    if (c == p.kchunks - 1) umma_commit(bar_acc_full + 8 * buf);     // the tile's accumulators are complete
                }
                __syncwarp();
            }
        }
    } else if (warp >= 11) {
        // ---------------------------------- epilogue: skip = accumulators + bias ----------------------------------
        // 4 warps, thread = TMEM lane = pixel row; 32 columns at a time through ONE staging tile and a TMA store.  Its own
        // warps, so that the compute warps go straight to the next tile (as part of their loop it was 22 % of a tile).
        const int et = threadIdx.x - (GS_COMPUTE + 96);           // 0..127
        const int q = warp & 3;                                    // the lane quarter this warp may access
        const int row = q * 32 + (threadIdx.x & 31);
        const uint32_t sw = (uint32_t)(row & 7);
        const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16);
        const bool issuer = et == 0;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
            const uint32_t buf = p.acc_bufs == 2 ? (it & 1u) : 0u, use = it / (uint32_t)p.acc_bufs;
            mbar_wait(bar_acc_full + 8 * buf, use & 1u);
            tc_fence_after();
            for (int col = 0; col < p.Cout; col += 32) {
                float v[32];
                tmem_ld32(tlane + buf * acc_stride + (uint32_t)col, v);
                const float4 *bias4 = reinterpret_cast<const float4 *>(p.bias + col);
                if (issuer) bulk_wait_read<0>();           // the previous chunk's store has read the tile out
                named_bar(2, GS_EPI);
                const uint32_t orow = sm_s + (uint32_t)row * 128u;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 bz = __ldg(bias4 + j);
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(orow + (((uint32_t)j ^ sw) << 4)), "f"(v[4 * j] + bz.x),
                                 "f"(v[4 * j + 1] + bz.y), "f"(v[4 * j + 2] + bz.z), "f"(v[4 * j + 3] + bz.w) : "memory");
                }
                fence_async_smem();
                named_bar(2, GS_EPI);
                if (issuer) {
                    tma_store_2d(&tmSkip, sm_s, col, tile * GS_ROWS);
                    bulk_commit();
                }
            }
            tc_fence_before();
            mbar_arrive(bar_acc_free + 8 * buf);            // every row of this accumulator tile has left tensor memory
        }
        if (issuer) bulk_wait_read<0>();
    } else {
        // ------------- compute: 8 warps; thread = (pixel row = TMEM lane, channel half of the 64-channel chunk) -------------
        const int tid = threadIdx.x;                       // 0..255
        const int row = tid & (GS_ROWS - 1), half = tid >> 7;
        const uint32_t sw = (uint32_t)(row & 7);
        const uint32_t tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        float *ca = coef, *cb = coef + p.Cin;
        uint32_t sx = 0, phx = 0;
        uint32_t so = 0;                                    // output staging tile of the next chunk
        const bool issuer = tid == 32;                      // issues the TMA stores (warp 0's elected lane issues the MMAs)
        uint32_t g = 0;                                     // chunks issued so far (a_empty completes once per chunk)
        int cur_b = -1;
        const int cpg = p.Cin / p.groups;
        const bool prof = p.prof != nullptr && blockIdx.x == 0 && tid == 0;
        long long tp = prof ? clock64() : 0;
#define GSPROF(slot)                                                      \
    if (prof) {                                                           \
        const long long now_ = clock64();                                 \
        atomicAdd(p.prof + (slot), (unsigned long long)(now_ - tp));      \
        tp = now_;                                                        \
    }
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
            const int b = (int)(((long long)tile * GS_ROWS) / p.HW);
            if (b != cur_b) {
                // coefficients of this sample: the formulas and the reduction order of elementwise.cu::gn_coefficients
                // (8 strided partial sums per group, xor-tree 4, 2, 1)
                cur_b = b;
                const double n = (double)p.HW * cpg;
                for (int g0 = 0; g0 < p.groups; g0 += GS_COMPUTE / 8) {
                    const int gi = g0 + (tid >> 3), l = tid & 7;
                    double a = 0.0, a2 = 0.0;
                    if (gi < p.groups) {
                        const double2 *st = reinterpret_cast<const double2 *>(p.stats + ((long long)b * p.stats_ld + (long long)gi * cpg) * 2);
                        for (int c = l; c < cpg; c += 8) { const double2 v = st[c]; a += v.x; a2 += v.y; }
                    }
#pragma unroll
                    for (int o = 4; o > 0; o >>= 1) {
                        a += __shfl_xor_sync(0xffffffffu, a, o);
                        a2 += __shfl_xor_sync(0xffffffffu, a2, o);
                    }
                    if (gi < p.groups && l == 0) {
                        const double mean = a / n;
                        double var = a2 / n - mean * mean;
                        if (var < 0.0) var = 0.0;
                        gmean[gi] = (float)mean;
                        grstd[gi] = (float)(1.0 / sqrt(var + (double)p.eps));
                    }
                }
                named_bar(1, GS_COMPUTE);
                for (int c = tid; c < p.Cin; c += GS_COMPUTE) {
                    const int gi = c / cpg;
                    const float ga = p.gamma[c] * grstd[gi];
                    ca[c] = ga;
                    cb[c] = p.beta[c] - gmean[gi] * ga;
                }
                named_bar(1, GS_COMPUTE);
            }
            for (int c = 0; c < p.kchunks; ++c, ++g) {
                // ---- this thread's 32 fp32 channels: its pixel's row of box `half` ----
                GSPROF(0)                                           // (tile / chunk bookkeeping, coefficient refresh)
                mbar_wait(bar_x_full + 8 * sx, phx);
                GSPROF(1)                                           // waiting for x
                float xv[32];
                {
                    const uint32_t r0 = sm_x + sx * GS_XSLOT + (uint32_t)half * (GS_XSLOT / 2) + (uint32_t)row * 128u;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float4 t;
                        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                     : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w)
                                     : "r"(r0 + (((uint32_t)j ^ sw) << 4)));
                        xv[4 * j] = t.x; xv[4 * j + 1] = t.y; xv[4 * j + 2] = t.z; xv[4 * j + 3] = t.w;
                    }
                }
                const uint32_t sx_used = sx;          // released below, once every value has been CONSUMED (see there)
                if (++sx == (uint32_t)p.nx) { sx = 0; phx ^= 1u; }
                // ---- act = fp16(SiLU(x * ca + cb)): this thread's 64 B of the pixel's row in the staging tile ----
                {
                    const float4 *a4 = reinterpret_cast<const float4 *>(ca + c * 64 + half * 32);
                    const float4 *b4 = reinterpret_cast<const float4 *>(cb + c * 64 + half * 32);
                    const uint32_t orow = sm_o + so * GS_ATILE + (uint32_t)row * 128u;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4 a0 = a4[2 * j], a1 = a4[2 * j + 1], b0 = b4[2 * j], b1 = b4[2 * j + 1];
                        uint4 o;
                        o.x = gs_h2(hl_silu_fast(fmaf(xv[8 * j], a0.x, b0.x)), hl_silu_fast(fmaf(xv[8 * j + 1], a0.y, b0.y)));
                        o.y = gs_h2(hl_silu_fast(fmaf(xv[8 * j + 2], a0.z, b0.z)), hl_silu_fast(fmaf(xv[8 * j + 3], a0.w, b0.w)));
                        o.z = gs_h2(hl_silu_fast(fmaf(xv[8 * j + 4], a1.x, b1.x)), hl_silu_fast(fmaf(xv[8 * j + 5], a1.y, b1.y)));
                        o.w = gs_h2(hl_silu_fast(fmaf(xv[8 * j + 6], a1.z, b1.z)), hl_silu_fast(fmaf(xv[8 * j + 7], a1.w, b1.w)));
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(orow + (((uint32_t)(half * 4 + j) ^ sw) << 4)), "r"(o.x),
                                     "r"(o.y), "r"(o.z), "r"(o.w) : "memory");
                    }
                }
                // ---- hi / lo A operands of x * 2^-4 into tensor memory: buffer g & 1 (the MMAs of chunk g - 2 read it last) ----
                GSPROF(2)                                           // x row -> act staging
                const uint32_t ab = g & 1u;
                if (g >= 2) mbar_wait(bar_a_empty + 8 * ab, ((g >> 1) - 1u) & 1u);
                GSPROF(3)                                           // waiting for the A buffer
                tc_fence_after();
                {
                    uint32_t hi[16], lo[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float v0 = xv[2 * j] * HL_OP_SCALE, v1 = xv[2 * j + 1] * HL_OP_SCALE;
                        hi[j] = gs_h2(v0, v1);
                        lo[j] = gs_lo(v0, v1, hi[j]);
                    }
                    const uint32_t acol = tlane + GS_TM_A + ab * 64u + (uint32_t)half * 16u;
                    tmem_st16(acol, hi);
                    tmem_st16(acol + 32u, lo);
                }
                tmem_st_wait();
                // The X slot goes back to the producer only here: every loaded value has been used by now.  Arriving right
                // after the ld.shared instructions let the refill (TMA) overtake the last loads of slow threads -- measured:
                // the final 16-32 bytes of a few rows per launch came from the slot's NEXT chunk once the compute warps no
                // longer waited for the MMA issue.
                mbar_arrive(bar_x_empty + 8 * sx_used);
                tc_fence_before();
                mbar_arrive(bar_a_full + 8 * ab);                   // -> the MMA warp
                fence_async_smem();
                tc_fence_before();
                // every store issued so far has left its staging tile: after the barrier the OTHER tile may be rewritten
                GSPROF(4)                                           // hi / lo -> tensor memory
                if (issuer) bulk_wait_read<0>();
                named_bar(1, GS_COMPUTE);
                GSPROF(5)                                           // chunk barrier
                if (issuer) {
                    tma_store_2d(&tmAct, sm_o + so * GS_ATILE, c * 64, tile * GS_ROWS);
                    bulk_commit();
                }
                so ^= 1u;
                GSPROF(6)                                           // act store issue
            }
            if (prof) atomicAdd(p.prof + 9, 1ull);                  // tiles of CTA 0
        }
#undef GSPROF
    }

    if (threadIdx.x == 32) bulk_wait_read<0>();          // the last stores have read their staging tiles
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)p.tmem_cols) : "memory");
    }
}

bool gs_plan(int B, int HW, int Cin, int Cout, GsParams *p, size_t *smem) {
    if (HW % GS_ROWS || Cin % 64 || Cin > GS_MAX_C || Cout % 32) return false;
    int nt, n_tiles;
    if (Cout <= 256 && Cout % 16 == 0) { nt = Cout; n_tiles = 1; }
    else if (Cout % 192 == 0 && Cout / 192 <= 2) { nt = 192; n_tiles = Cout / 192; }
    else if (Cout % 256 == 0 && Cout / 256 <= 2) { nt = 256; n_tiles = Cout / 256; }
    else return false;
    p->nt = nt;
    p->n_tiles = n_tiles;
    if (nt * n_tiles > (int)GS_TM_A) return false;
    p->tmem_cols = 512;
    p->acc_bufs = 2 * nt * n_tiles <= (int)GS_TM_A ? 2 : 1;
    p->nw = 3;
    const int fixed = 1024 + p->nw * nt * 128 + 3 * GS_ATILE /* act staging x 2, skip staging */ + 2 * Cin * (int)sizeof(float);
    const int budget = 227 * 1024 - 2048;        // static shared memory: barriers, group statistics
    int nx = (budget - fixed) / GS_XSLOT;
    if (nx > GS_MAX_NX) nx = GS_MAX_NX;
    if (nx < 2) return false;
    p->nx = nx;
    p->kchunks = Cin / 64;
    p->total_tiles = (int)((long long)B * HW / GS_ROWS);
    *smem = (size_t)fixed + (size_t)nx * GS_XSLOT;
    return true;
}

}  // namespace

static unsigned long long *g_gs_prof = nullptr;
extern "C" int hl_gn_skip_set_profile(void *dev_counters) {
    g_gs_prof = (unsigned long long *)dev_counters;
    return HL_OK;
}

extern "C" int hl_gn_skip_supported(int B, int HW, int Cin, int Cout) {
    GsParams p = {};
    size_t smem = 0;
    return gs_plan(B, HW, Cin, Cout, &p, &smem) ? 1 : 0;        // host-only shape query (no driver needed: plans built on a CPU match)
}

extern "C" int hl_gn_skip(const float *x, int ldx, const double *stats, int stats_ld, const float *gamma,
                          const float *beta, void *act, int ld_act, const void *wpk, const float *bias, float *skip,
                          int ld_skip, int B, int HW, int Cin, int Cout, int groups, float eps, void *stream) {
    HL_CHECK_ARG(x && stats && gamma && beta && act && wpk && bias && skip && B > 0 && HW > 0);
    HL_CHECK_ARG(groups > 0 && groups <= 64 && Cin % groups == 0 && stats_ld >= Cin && ldx >= Cin && ld_act >= Cin && ld_skip >= Cout);
    HL_CHECK_ARG(ldx % 4 == 0 && ld_act % 8 == 0 && ld_skip % 4 == 0);
    HL_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)act & 15) == 0 && ((uintptr_t)skip & 15) == 0 &&
                 ((uintptr_t)wpk & 15) == 0 && ((uintptr_t)bias & 15) == 0);
    GsParams p = {};
    size_t smem = 0;
    HL_CHECK_ARG(gs_plan(B, HW, Cin, Cout, &p, &smem));
    PFN_hl_encodeTiled encode = hl_get_encode_tiled();
    if (!encode) {
        hl_set_error("cuTensorMapEncodeTiled unavailable");
        return HL_E_CUDA;
    }
    p.Cin = Cin; p.Cout = Cout; p.HW = HW; p.groups = groups; p.stats_ld = stats_ld; p.ld_act = ld_act; p.ld_skip = ld_skip;
    p.eps = eps; p.stats = stats; p.gamma = gamma; p.beta = beta; p.bias = bias; p.act = (__half *)act; p.skip = skip; p.prof = g_gs_prof;
    CUtensorMap tmX, tmW;
    {
        cuuint64_t gdim[2] = {(cuuint64_t)Cin, (cuuint64_t)B * HW};
        cuuint64_t gstr[1] = {(cuuint64_t)ldx * 4};
        cuuint32_t box[2] = {32, GS_ROWS}, estr[2] = {1, 1};
        CUresult r = encode(&tmX, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)x, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            hl_set_error("cuTensorMapEncodeTiled(x) failed: %d (Cin=%d rows=%lld ldx=%d)", (int)r, Cin, (long long)B * HW, ldx);
            return HL_E_CUDA;
        }
    }
    {
        const int cout_pad = (Cout + 31) / 32 * 32;
        cuuint64_t gdim[3] = {(cuuint64_t)Cin, (cuuint64_t)cout_pad, 2};
        cuuint64_t gstr[2] = {(cuuint64_t)Cin * 2, (cuuint64_t)cout_pad * Cin * 2};
        cuuint32_t box[3] = {64, (cuuint32_t)p.nt, 1}, estr[3] = {1, 1, 1};
        CUresult r = encode(&tmW, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, (void *)wpk, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            hl_set_error("cuTensorMapEncodeTiled(w) failed: %d (Cin=%d Cout=%d)", (int)r, Cin, Cout);
            return HL_E_CUDA;
        }
    }
    CUtensorMap tmAct, tmSkip;
    {
        cuuint64_t gdim[2] = {(cuuint64_t)Cin, (cuuint64_t)B * HW};
        cuuint64_t gstr[1] = {(cuuint64_t)ld_act * 2};
        cuuint32_t box[2] = {64, GS_ROWS}, estr[2] = {1, 1};
        CUresult r = encode(&tmAct, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, act, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            hl_set_error("cuTensorMapEncodeTiled(act) failed: %d (Cin=%d ld=%d)", (int)r, Cin, ld_act);
            return HL_E_CUDA;
        }
    }
    {
        cuuint64_t gdim[2] = {(cuuint64_t)Cout, (cuuint64_t)B * HW};
        cuuint64_t gstr[1] = {(cuuint64_t)ld_skip * 4};
        cuuint32_t box[2] = {32, GS_ROWS}, estr[2] = {1, 1};
        CUresult r = encode(&tmSkip, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, skip, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            hl_set_error("cuTensorMapEncodeTiled(skip) failed: %d (Cout=%d ld=%d)", (int)r, Cout, ld_skip);
            return HL_E_CUDA;
        }
    }
    static bool configured[64] = {};
    int dev = 0;
    HL_CHECK_CUDA(cudaGetDevice(&dev));
    if (!configured[dev & 63]) {
        HL_CHECK_CUDA(cudaFuncSetAttribute(k_gn_skip_tc5, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048));
        configured[dev & 63] = true;
    }
    const int sms = hl_num_sms();
    const int grid = p.total_tiles < sms ? p.total_tiles : sms;
    HL_CHECK_CUDA(hl_launch(k_gn_skip_tc5, dim3(grid), dim3(GS_THREADS), smem, (cudaStream_t)stream, tmX, tmW, tmAct, tmSkip, p));
    HL_CHECK_LAUNCH();
    return HL_OK;
}
