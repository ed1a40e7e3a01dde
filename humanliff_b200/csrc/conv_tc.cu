// tcgen05 (5th-gen tensor core) implicit-GEMM convolution for sm_100a.
// NHWC activations (fp16, or fp32 read as TF32), fp32 accumulation in TMEM, fp32 NHWC output.
//
//   GEMM view   M = B*H*W output pixels   tile: mh x 128 pixels (mh = 1 or 2 "halves", one TMEM
//                                         accumulator of 128 lanes x n_tile columns per half)
//               N = Cout                  tile: n_tile = 32..256 output channels
//               K = taps * Cin            chunk: 128 bytes of channels (64 fp16 / 32 tf32) of one tap
//
//   A operand   the activation tensor itself, fetched by TMA; out-of-bounds pixels are zero-filled by
//               the TMA unit, which IS the conv's zero padding -- no im2col buffer ever exists.
//       TAP  mode: one 4-D box {chunk, bw, bh, bn} per (tap, chunk, half), origin shifted by the tap
//                  (stride-2 convs use the tensor map's element strides).
//       HALO mode (3x3, stride 1, W % 128 == 0): a half is one image row segment of 128 pixels; the
//                  mh+2 halo rows {chunk, 130 px} are loaded ONCE per channel chunk into a ring of
//                  row slots and all 9 taps read them through shifted shared-memory descriptors
//                  (start = slot + (dx+1)*128 B; the hardware derives the swizzle phase from the absolute
//                  shared-memory address, so the descriptor's base_offset field stays 0 -- measured):
//                  2.9x (mh=1) .. 4.3x (mh=2) fewer A bytes from L2 than per-tap boxes.
//   B operand   packed weights [tap][Cout_pad][Cin], 3-D TMA box {chunk, n_tile, 1}; with mh = 2 one
//               weight tile feeds both halves (halves the weight traffic per flop).
//   both land in shared memory in the SWIZZLE_128B K-major layout tcgen05.mma consumes directly.
//
//   persistent  grid = min(tiles, SMs); each CTA walks tiles t = blockIdx.x + i*gridDim.x (n-tile
//               fastest so concurrently running CTAs share activations in L2).
//   CTA pairs   wherever the 128-pixel boxes pair up the kernel runs as 2-CTA clusters issuing
//               tcgen05.mma.cta_group::2 (M = 256: each SM supplies its 128 A rows and half of the B rows;
//               the leader CTA issues, barriers are reached through mapa / shared::cluster).
//   warp roles  warp 0: A producer (TMA)   warp 1: B producer (TMA)   warp 2: TMEM alloc + MMA issue
//               warps 3-10: two epilogue warpgroups on alternate 32-column chunks.  Two independent smem
//               rings (A, B) with full/empty mbarriers; tcgen05.commit releases slots; TMEM accumulators
//               are double-buffered when they fit (acc_stages = 2) so the epilogue of tile i overlaps the
//               main loop of tile i+1.
//   split-K     sub-wave 3x3 layers: K cut into S channel-chunk slices over S x the CTAs, fp32 partial tiles
//               into a caller-registered workspace, fixed-order second pass (k_splitk_reduce).
//   planning    N tile and (N tile, S) are chosen by a shared-memory-bandwidth cost model (make_plan,
//               choose_split): measured, the kernel is bound by the ~92 B/clk/SM that MMA
//               operand fetch, TMA operand writes and the epilogue share.
//   epilogue    tcgen05.ld (32 columns) -> + bias (+ residual, TMA-prefetched into the staging
//               buffer, added in place) -> swizzled smem staging -> TMA store (coalesced, clipped
//               at the tensor edge; fp32, or the fp16 operand directly with HL_CONV_OUT_F16); optional
//               per-channel sum / sum-of-squares of the OUTPUT (the following GroupNorm's statistics),
//               combined in a fixed order per chunk, accumulated per CTA in fp64 shared memory and flushed
//               with fp64 atomics once per sample.
//
// Algorithmic work per launch: 2*M*N*K flop; compulsory HBM bytes: e*(M*Cin + taps*Cout*Cin) +
// 4*M*Cout (+ 4*M*Cout residual), e = operand bytes per element.
#include "common.cuh"
#include "tc5.cuh"

#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

int hl_num_sms();

namespace {

constexpr int ROW_BYTES = 128;                   // one K chunk of one pixel / one weight row
constexpr int BLOCK_M = 128;
constexpr int A_BOX_BYTES = BLOCK_M * ROW_BYTES;  // 16 KB
constexpr int HALO_PIX = BLOCK_M + 2;
constexpr int HALO_ROW_BYTES = HALO_PIX * ROW_BYTES;   // 16640
constexpr int HALO_SLOT_BYTES = 17 * 1024;            // rounded up to the 1024 B swizzle period
constexpr int STAGE_BUF_BYTES = BLOCK_M * 128;        // epilogue staging: 128 rows x 32 fp32
constexpr int MAX_SLOTS = 8;
constexpr int MAX_NBUF = 4;
constexpr int EPI_GROUPS = 2;                         // independent epilogue warpgroups (alternate chunks)
constexpr int EPI_THREADS = 128;                      // threads per epilogue group (one per TMEM lane)
constexpr int NUM_THREADS = (3 + 4 * EPI_GROUPS) * 32;
constexpr int STATS_MAX_C = 768;
constexpr int SMEM_LIMIT = 227 * 1024;
constexpr int DYN_SMEM_MAX = SMEM_LIMIT - 15 * 1024;  // static smem: barriers, fp64 statistics accumulators (12 KB), partials (2 KB)

struct TcParams {
    int B, H, W;             // OUTPUT spatial size
    int Cout, stride, ksize, taps, kchunks, kind;   // kind: 0 = tf32, 1 = f16
    int bw, bh, bn, tiles_w, tiles_h;
    int mh, n_tile, n_tiles, nchunks;
    int pair;                // CTAs cooperating on one MMA (cta_group): 1, or 2 = CTA pair, M = 256
    int total_tiles;
    int halo, hp;            // hp = H / mh (halo mode)
    int base_off;            // experiment: halo descriptors carry base_offset = (addr >> 7) & 7
    unsigned long long *prof;   // optional cycle counters of CTA 0 (hl_conv_set_profile)
    int acc_stages, acc_stride, tmem_cols;
    int a_slots, a_slot_bytes, b_slots, b_slot_bytes, nbuf;
    const float *bias;
    int has_res;
    int y_f16;               // 1: output written as fp16 (an operand buffer) instead of fp32; 2: as a scaled fp16 hi | lo pair
    // operand passes (hi + lo operands): the K loop runs npass * kchunks "virtual" chunks; pass q reads
    // the A channels [a_off[q], a_off[q] + Cin) and the weight slab b_slab[q] (taps b_slab[q]*taps ..)
    int npass, a_off[3], b_slab[3], vchunks;
    int ksplit, kc_split, tiles_mn;   // split-K: tile = ks * tiles_mn + (m, n); split ks covers kc_split (virtual) chunks
    int split_b;             // batch-coordinate offset per split in the partial-sum workspace (= B)
    double *stats;
    int stats_ld;
    // dual output (hl_conv2d_dual): epilogue group 0 writes y = conv + bias + residual (tmY, stats), group 1 writes
    // y2 = conv + bias (tmY2, stats2) from the same accumulators -- both groups drain EVERY chunk
    int dual;
    double *stats2;
    int stats2_ld;
    int sacc2_off;           // byte offset (from the aligned dynamic shared-memory base) of group 1's fp64 accumulators
    // split-K with the second pass INSIDE the kernel (red = 1): every slice CTA stores its fp32 partial tile, bumps the
    // tile's counter, and the CTA that arrives last adds the S partial tiles in slice order and runs the real epilogue
    // (bias, residual, rounding, statistics) through the final output map tmF
    int red;
    const float *red_ws;     // partial-sum workspace [S][split_b][H][W][red_ldw]
    long long red_slice;     // elements between slices
    int red_ldw;
    const float *red_bias, *red_res;
    int red_ldr, red_yf16;
    double *red_stats;
    int red_stats_ld, red_rows;   // red_rows: pixels per sample inside a 128-pixel box (128 | 64)
    unsigned *red_ctr;       // one counter per (128-pixel box, n tile); zero between launches (the last CTA resets it)
    int stats_rows;          // 0: one sample per 128-pixel box (per-CTA fp64 accumulators); 32 | 64: a box holds 128 / rows
                             // samples of `rows` pixels each (8^2: 64): the per-quarter sums go straight to global memory
};

// profiling: cycles CTA 0 spends blocked in each wait, accumulated into p.prof[slot]
struct ProfTimer {
    unsigned long long *dst;
    long long t0;
    __device__ __forceinline__ ProfTimer(unsigned long long *prof, int slot, bool on = true)
        : dst(prof && on && blockIdx.x == 0 ? prof + slot : nullptr), t0(0) {
        if (dst) t0 = clock64();
    }
    __device__ __forceinline__ ~ProfTimer() {
        if (dst) atomicAdd(dst, (unsigned long long)(clock64() - t0));
    }
};
#define PROF(slot) ProfTimer prof_timer_##slot(p.prof, slot)
#define PROF_IF(slot, cond) ProfTimer prof_timer_##slot(p.prof, slot, cond)   // warp-convergent: all threads time

// origin (output pixel coordinates) of half `half` of m-tile `mt`
// (a cluster tile holds p.pair * p.mh boxes; CTA `rank` of the pair owns boxes rank*mh .. rank*mh + mh-1)
__device__ __forceinline__ void box_origin(const TcParams &p, int mt, int half, int rank, int &w0, int &h0, int &n0) {
    const int per = p.mh * p.pair, sub = rank * p.mh + half;
    if (p.halo) {
        const int tw = mt % p.tiles_w;
        const int r = mt / p.tiles_w;
        w0 = tw * BLOCK_M;
        h0 = (r % p.hp) * per + sub;
        n0 = r / p.hp;
    } else {
        const int bi = mt * per + sub;
        const int tw = bi % p.tiles_w;
        const int r = bi / p.tiles_w;
        w0 = tw * p.bw;
        h0 = (r % p.tiles_h) * p.bh;
        n0 = (r / p.tiles_h) * p.bn;
    }
}

// RED: split-K launch whose second pass runs inside the kernel (hl_conv_set_split_reduce(1); a separate instantiation, so
// that the one-pass launches keep the lean epilogue: with the two-pass loop compiled into every launch the step lost 1.2 ms).
// EXPERIMENT, off by default -- measured on B200: bit-identical results, but +0.9 ms/step at B = 4 and +1.0 ms at B = 1
// against the separate k_splitk_reduce launch: the last CTA of a tile re-reads the S partial tiles with 256 threads after a
// full store-completion wait and a device-scope fence, where the separate launch spreads the same reads over the whole
// GPU.  The 49 (B = 4) / 72 (B = 1) reduction launches are cheaper than that serialisation.
template <bool CTA2, bool RED>
__global__ void __launch_bounds__(NUM_THREADS, 1)
k_conv_tc(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
          const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmR,
          const __grid_constant__ CUtensorMap tmY2, const __grid_constant__ CUtensorMap tmF, const TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[4 * MAX_SLOTS + 4 + EPI_GROUPS * MAX_NBUF];
    __shared__ uint32_t tmem_slot;
    // per-CTA channel sums in fp64: floating-point atomics commute only up to rounding, and fp64 rounding
    // (1e-16) is far below the fp32 resolution of everything downstream -> results are reproducible run to run
    __shared__ double sacc[2][STATS_MAX_C];
    __shared__ float2 spart[EPI_GROUPS][4][32];     // per chunk: (sum, sum of squares) of each 32-row quarter
    __shared__ int s_last;                          // split-K: this CTA arrived last at its tile's counter

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t smem_a = smem_base;
    const uint32_t smem_b = smem_a + (uint32_t)(p.a_slots * p.a_slot_bytes);
    const uint32_t smem_e = smem_b + (uint32_t)(p.b_slots * p.b_slot_bytes);
    const uint32_t bar_a_full = smem_u32(&bars[0]);
    const uint32_t bar_a_empty = smem_u32(&bars[MAX_SLOTS]);
    const uint32_t bar_b_full = smem_u32(&bars[2 * MAX_SLOTS]);
    const uint32_t bar_b_empty = smem_u32(&bars[3 * MAX_SLOTS]);
    const uint32_t bar_t_full = smem_u32(&bars[4 * MAX_SLOTS]);
    const uint32_t bar_t_empty = smem_u32(&bars[4 * MAX_SLOTS + 2]);
    const uint32_t bar_r_full = smem_u32(&bars[4 * MAX_SLOTS + 4]);
    const int chunk_elems = p.kind ? 64 : 32;
    // CTA pair: rank 0 is the leader (issues the MMAs, owns the barriers the MMA thread waits on)
    const int rank = CTA2 ? (int)cluster_rank() : 0;
    hl_pdl_trigger_early();
    const int cid = CTA2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;    // first tile of this CTA / pair
    const int nct = CTA2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;      // tile stride
    const int nprod = CTA2 ? 2 : 1;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmY) : "memory");
        if (p.has_res) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmR) : "memory");
        if (p.dual) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmY2) : "memory");
        if (RED) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmF) : "memory");
    }
    if (warp == 2) {
        if (lane == 0) {
            for (int s = 0; s < MAX_SLOTS; ++s) {
                mbar_init(bar_a_full + 8 * s, nprod);
                mbar_init(bar_a_empty + 8 * s, 1);
                mbar_init(bar_b_full + 8 * s, (p.halo ? 1 : 2) * nprod);   // TAP: A and B producers fill one stage
                mbar_init(bar_b_empty + 8 * s, 1);
            }
            for (int s = 0; s < 2; ++s) {
                mbar_init(bar_t_full + 8 * s, 1);
                mbar_init(bar_t_empty + 8 * s, nprod * EPI_GROUPS * EPI_THREADS);
            }
            for (int s = 0; s < EPI_GROUPS * MAX_NBUF; ++s) mbar_init(bar_r_full + 8 * s, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        if (CTA2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                             smem_u32(&tmem_slot)),
                         "r"((uint32_t)p.tmem_cols)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                             smem_u32(&tmem_slot)),
                         "r"((uint32_t)p.tmem_cols)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    // group 1's accumulators in dual mode live in dynamic shared memory (only those launches pay for them)
    double *const sacc2 = reinterpret_cast<double *>(smem_raw + (smem_base - smem_u32(smem_raw)) + p.sacc2_off);
    if (warp >= 3 && p.stats) {
        for (int i = threadIdx.x - 96; i < 2 * STATS_MAX_C; i += EPI_GROUPS * EPI_THREADS) {
            (&sacc[0][0])[i] = 0.0;
            if (p.dual) sacc2[i] = 0.0;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (CTA2) cluster_sync_all(); else __syncthreads();     // barriers initialised + TMEM allocated in both CTAs
    hl_pdl_wait();       // everything above overlapped the previous kernel's tail; from here on its results are visible
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_slot;
    const int pad = p.ksize / 2;

    // The three single-thread roles below are latency-critical: one thread issues everything, so
    // their inner loops carry no divisions / tile decoding (ring positions are advanced incrementally,
    // tile coordinates are decoded once per tile, descriptors are 32-bit adds on a precomputed base).
    // Measured on B200: ~1000 cycles of scalar overhead per stage (runtime div/mod) capped the first
    // version of this kernel at 27 % of the tensor peak regardless of tile shape.
    if (warp == 0) {
        {
            // ------------------------------ A producer ------------------------------
            // TAP mode shares the stage ring (and its barriers) with the B producer; HALO mode owns
            // the row ring.
            const uint32_t full_l = p.halo ? bar_a_full : bar_b_full, empty0 = p.halo ? bar_a_empty : bar_b_empty;
            const uint32_t full0 = CTA2 ? mapa_u32(full_l, 0) : full_l;     // the leader's barrier
            const uint32_t nslots = (uint32_t)p.a_slots, slot_bytes = (uint32_t)p.a_slot_bytes;
            uint32_t slot = 0, phase = 0;
            for (int tile = cid; tile < p.total_tiles; tile += nct) {
                const int ks = tile / p.tiles_mn;
                const int mt = (tile - ks * p.tiles_mn) / p.n_tiles;
                int pass = (ks * p.kc_split) / p.kchunks, kcp = ks * p.kc_split - pass * p.kchunks;   // first virtual chunk
                int w0[2], h0[2], n0[2];
                box_origin(p, mt, 0, rank, w0[0], h0[0], n0[0]);
                box_origin(p, mt, p.mh - 1, rank, w0[1], h0[1], n0[1]);
                if (p.halo) {
                    const int rows = p.mh + 2;
                    for (int kc = 0; kc < p.kc_split; ++kc) {
                        const int c0 = p.a_off[pass] + kcp * chunk_elems;
                        if (++kcp == p.kchunks) { kcp = 0; ++pass; }
                        for (int r = 0; r < rows; ++r) {
                            { PROF_IF(7, lane == 0); mbar_wait(empty0 + 8 * slot, phase ^ 1u); }
                            if (elect_one_sync()) {
                                if (CTA2) {
                                    mbar_expect_tx_cluster(full0 + 8 * slot, HALO_ROW_BYTES);
                                    tma_load_4d_2sm(smem_a + slot * slot_bytes, &tmA, full0 + 8 * slot, c0, w0[0] - 1,
                                                    h0[0] - 1 + r, n0[0]);
                                } else {
                                    mbar_expect_tx(full0 + 8 * slot, HALO_ROW_BYTES);
                                    tma_load_4d(smem_a + slot * slot_bytes, &tmA, full0 + 8 * slot, c0, w0[0] - 1,
                                                h0[0] - 1 + r, n0[0]);
                                }
                            }
                            __syncwarp();
                            if (++slot == nslots) { slot = 0; phase ^= 1u; }
                        }
                    }
                } else {
                    const uint32_t bytes = (uint32_t)(p.mh * A_BOX_BYTES);
                    const int xs0 = w0[0] * p.stride - pad, ys0 = h0[0] * p.stride - pad;
                    const int xs1 = w0[1] * p.stride - pad, ys1 = h0[1] * p.stride - pad;
                    for (int kc = 0; kc < p.kc_split; ++kc) {
                        const int c0 = p.a_off[pass] + kcp * chunk_elems;
                        if (++kcp == p.kchunks) { kcp = 0; ++pass; }
                        for (int ty = 0; ty < p.ksize; ++ty) {
                            for (int tx = 0; tx < p.ksize; ++tx) {
                                { PROF_IF(7, lane == 0); mbar_wait(empty0 + 8 * slot, phase ^ 1u); }
                                if (elect_one_sync()) {
                                    const uint32_t dst = smem_a + slot * slot_bytes;
                                    if (CTA2) {
                                        mbar_expect_tx_cluster(full0 + 8 * slot, bytes);
                                        tma_load_4d_2sm(dst, &tmA, full0 + 8 * slot, c0, xs0 + tx, ys0 + ty, n0[0]);
                                    } else {
                                        mbar_expect_tx(full0 + 8 * slot, bytes);
                                        tma_load_4d(dst, &tmA, full0 + 8 * slot, c0, xs0 + tx, ys0 + ty, n0[0]);
                                        if (p.mh == 2)
                                            tma_load_4d(dst + A_BOX_BYTES, &tmA, full0 + 8 * slot, c0, xs1 + tx,
                                                        ys1 + ty, n0[1]);
                                    }
                                }
                                __syncwarp();
                                if (++slot == nslots) { slot = 0; phase ^= 1u; }
                            }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        {
            // ------------------------------ B producer ------------------------------
            // a CTA of a pair loads its half of the weight tile (rows rank*N/2 .. +N/2)
            const int rows_b = p.n_tile / p.pair;
            const uint32_t bytes = (uint32_t)(rows_b * ROW_BYTES);
            const uint32_t nslots = (uint32_t)p.b_slots, slot_bytes = (uint32_t)p.b_slot_bytes;
            const uint32_t bfull = CTA2 ? mapa_u32(bar_b_full, 0) : bar_b_full;
            uint32_t slot = 0, phase = 0;
            for (int tile = cid; tile < p.total_tiles; tile += nct) {
                const int nt0 = (tile % p.n_tiles) * p.n_tile + rank * rows_b;     // tiles_mn is a multiple of n_tiles
                const int v0 = (tile / p.tiles_mn) * p.kc_split;
                int pass = v0 / p.kchunks, kcp = v0 - pass * p.kchunks;
                for (int kc = 0; kc < p.kc_split; ++kc) {
                    const int c0 = kcp * chunk_elems, tap0 = p.b_slab[pass] * p.taps;
                    if (++kcp == p.kchunks) { kcp = 0; ++pass; }
                    for (int tap = tap0; tap < tap0 + p.taps; ++tap) {
                        { PROF_IF(8, lane == 0); mbar_wait(bar_b_empty + 8 * slot, phase ^ 1u); }
                        if (elect_one_sync()) {
                            if (CTA2) {
                                mbar_expect_tx_cluster(bfull + 8 * slot, bytes);
                                tma_load_3d_2sm(smem_b + slot * slot_bytes, &tmB, bfull + 8 * slot, c0, nt0, tap);
                            } else {
                                mbar_expect_tx(bfull + 8 * slot, bytes);
                                tma_load_3d(smem_b + slot * slot_bytes, &tmB, bfull + 8 * slot, c0, nt0, tap);
                            }
                        }
                        __syncwarp();
                        if (++slot == nslots) { slot = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 2) {
        if (!CTA2 || rank == 0) {
            // ------------------------------ MMA issuer (leader CTA of a pair) ---------
            // instruction descriptor: D=f32, A/B format (0 = f16, 2 = tf32), both K-major, N>>3 @17, M>>4 @24
            const uint32_t fmt = p.kind ? 0u : 2u;
            const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(p.n_tile >> 3) << 17) |
                                   ((uint32_t)((BLOCK_M * (CTA2 ? 2 : 1)) >> 4) << 24);
            auto mma = [&](uint32_t d, uint64_t ad, uint64_t bd, uint32_t acc) {
                if (CTA2) umma2(p.kind, d, ad, bd, idesc, acc); else umma(p.kind, d, ad, bd, idesc, acc);
            };
            auto commit = [&](uint32_t bar) {
                if (CTA2) umma_commit2(bar); else umma_commit(bar);
            };
            // shared-memory descriptor: constant high word (SBO = 1024 B, version 1, SWIZZLE_128B), low
            // word = LBO field (unused, 1) | start address >> 4; a K step of 32 B adds 2 to the low word
            const uint32_t desc_hi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
            auto desc = [&](uint32_t saddr) -> uint64_t {
                uint32_t lo = ((saddr & 0x3FFFFu) >> 4) | (1u << 16);
                if (p.base_off) return umma_desc_sw128(saddr, 1u);
                return ((uint64_t)desc_hi << 32) | lo;
            };
            const uint32_t nb = (uint32_t)p.b_slots, b_bytes = (uint32_t)p.b_slot_bytes;
            const uint32_t na = (uint32_t)p.a_slots, a_bytes = (uint32_t)p.a_slot_bytes;
            uint32_t sb = 0, phb = 0;      // B ring (= the stage ring in TAP mode)
            uint32_t sa = 0, pha = 0;      // A row ring (HALO mode)
            int it = 0;
            PROF_IF(0, lane == 0);
            if (p.prof && blockIdx.x == 0 && lane == 0) p.prof[10] = (unsigned long long)((p.total_tiles + gridDim.x - 1) / gridDim.x);
            const int steps = p.kc_split * p.taps;
            for (int tile = cid; tile < p.total_tiles; tile += nct, ++it) {
                const int as = p.acc_stages == 2 ? (it & 1) : 0;
                const uint32_t aph = (uint32_t)(p.acc_stages == 2 ? (it >> 1) : it) & 1u;
                { PROF_IF(2, lane == 0); mbar_wait(bar_t_empty + 8 * as, aph ^ 1u); }
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t acc0 = tmem_base + (uint32_t)(as * p.mh * p.acc_stride);
                const uint32_t acc1 = acc0 + (uint32_t)p.acc_stride;
                if (p.halo) {
                    for (int kc = 0; kc < p.kc_split; ++kc) {
                        // claim the mh+2 row slots of this chunk (addresses + the parity to wait for)
                        uint32_t row_addr[4], row_bar[4], row_par[4];
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            if (r < p.mh + 2) {
                                row_addr[r] = smem_a + sa * a_bytes;
                                row_bar[r] = sa;
                                row_par[r] = pha;
                                if (++sa == na) { sa = 0; pha ^= 1u; }
                            }
                        }
#pragma unroll
                        for (int ty = 0; ty < 3; ++ty) {
                            // tap row ty reads halo rows ty (half 0) and ty+1 (half 1)
                            if (ty == 0) {
                                { PROF_IF(1, lane == 0); mbar_wait(bar_a_full + 8 * row_bar[0], row_par[0]); }
                                if (p.mh == 2) { PROF_IF(1, lane == 0); mbar_wait(bar_a_full + 8 * row_bar[1], row_par[1]); }
                            } else {
                                PROF_IF(1, lane == 0);
                                const uint32_t rb = p.mh == 2 ? row_bar[ty + 1] : row_bar[ty];
                                const uint32_t rp = p.mh == 2 ? row_par[ty + 1] : row_par[ty];
                                mbar_wait(bar_a_full + 8 * rb, rp);
                            }
#pragma unroll
                            for (int tx = 0; tx < 3; ++tx) {
                                { PROF_IF(3, lane == 0); mbar_wait(bar_b_full + 8 * sb, phb); }
                                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                                if (elect_one_sync()) {
                                    const uint64_t bd = desc(smem_b + sb * b_bytes);
                                    const uint64_t ad0 = desc(row_addr[ty] + tx * ROW_BYTES);
                                    const uint32_t accf = (kc | ty | tx) != 0;
#pragma unroll
                                    for (int k = 0; k < 4; ++k)
                                        mma(acc0, ad0 + 2 * k, bd + 2 * k, accf | (k != 0));
                                    if (p.mh == 2) {
                                        const uint64_t ad1 = desc(row_addr[ty + 1] + tx * ROW_BYTES);
#pragma unroll
                                        for (int k = 0; k < 4; ++k)
                                            mma(acc1, ad1 + 2 * k, bd + 2 * k, accf | (k != 0));
                                    }
                                    commit(bar_b_empty + 8 * sb);
                                    // halo row `ty` is not needed by later taps; the last tap row frees the rest
                                    if (tx == 2) {
                                        commit(bar_a_empty + 8 * row_bar[ty]);
                                        if (ty == 2 && p.mh == 2) commit(bar_a_empty + 8 * row_bar[3]);
                                    }
                                }
                                __syncwarp();
                                if (++sb == nb) { sb = 0; phb ^= 1u; }
                            }
                        }
                    }
                } else {
                    for (int st = 0; st < steps; ++st) {
                        { PROF_IF(3, lane == 0); mbar_wait(bar_b_full + 8 * sb, phb); }
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        if (elect_one_sync()) {
                            const uint64_t bd = desc(smem_b + sb * b_bytes);
                            const uint64_t ad0 = desc(smem_a + sb * a_bytes);
                            const uint32_t accf = st != 0;
#pragma unroll
                            for (int k = 0; k < 4; ++k) mma(acc0, ad0 + 2 * k, bd + 2 * k, accf | (k != 0));
                            if (p.mh == 2) {
                                const uint64_t ad1 = ad0 + (A_BOX_BYTES >> 4);
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    mma(acc1, ad1 + 2 * k, bd + 2 * k, accf | (k != 0));
                            }
                            commit(bar_b_empty + 8 * sb);
                        }
                        __syncwarp();
                        if (++sb == nb) { sb = 0; phb ^= 1u; }
                    }
                }
                if (elect_one_sync()) commit(bar_t_full + 8 * as);   // accumulators complete -> epilogue
                __syncwarp();
            }
        }
        hl_pdl_trigger_late();     // only the last epilogue is left: let the next kernel's CTAs move in
    } else {
        // ---------------------------------- epilogue ----------------------------------
        // Two independent warpgroups take alternate 32-column chunks (each with its own staging ring,
        // named barrier, TMA-issuing thread and residual prefetch cursor): per chunk a thread runs a
        // serial chain of ~200 instructions on its one accumulator row, so a single group is latency-
        // bound (measured 1400-2900 cycles per chunk) and two groups double the drain rate.
        const int ew = warp - 3;
        const int eg = ew >> 2;                   // epilogue group
        const int et = (ew & 3) * 32 + lane;      // 0..127 within the group
        const int q = warp & 3;                   // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;            // GEMM row = pixel index inside the box
        const bool e0 = et == 0;
        const bool pt = et == 1 && eg == 0;       // the thread whose waits are profiled
        const uint32_t sw = (uint32_t)(row & 7);
        const uint32_t smem_g = smem_e + (uint32_t)(eg * p.nbuf) * STAGE_BUF_BYTES;
        const uint32_t bar_r = bar_r_full + 8u * (uint32_t)(eg * MAX_NBUF);
        const uint32_t nbuf = (uint32_t)p.nbuf;   // staging ring of this group: 2, 3 or 4 buffers
        // dual output: group 1 is the "plain" group (no residual, second output map, second statistics row)
        const bool plain_g = p.dual && eg == 1;
        const bool g_res = p.has_res && !plain_g;
        const CUtensorMap *const tm_out = plain_g ? &tmY2 : &tmY;
        double *const sacc_s = plain_g ? sacc2 : &sacc[0][0];                  // [2][STATS_MAX_C]
        const uint32_t own_mask = p.dual ? 0u : (uint32_t)(EPI_GROUPS - 1);   // dual: every chunk belongs to both groups
        uint32_t rb = 0, rph = 0;                 // ring slot / phase of the chunk being processed
        uint32_t qn = 0;                          // global chunk counter of this CTA
        uint32_t ql = 0;                          // chunks owned by this group so far (staging ring position)
        int cur_n = -1;                           // sample whose statistics sit in sacc

        // residual prefetch cursor (thread e0 of the group): walks the same (tile, half, chunk) sequence
        // and issues loads for the chunks this group owns
        int l_tile = cid, l_half = 0, l_cc = 0;
        uint32_t l_qn = 0, l_b = 0;               // cursor: global chunk index, ring slot of the next load
        auto issue_res_load = [&]() {
            for (;;) {
                if (l_tile >= p.total_tiles) return;
                const int nt0 = (l_tile % p.n_tiles) * p.n_tile;
                const bool mine = (int)(l_qn & own_mask) == (p.dual ? 0 : eg);
                if (mine) {
                    int w0, h0, n0;
                    box_origin(p, (l_tile % p.tiles_mn) / p.n_tiles, l_half, rank, w0, h0, n0);
                    mbar_expect_tx(bar_r + 8 * l_b, STAGE_BUF_BYTES);
                    tma_load_4d(smem_g + l_b * STAGE_BUF_BYTES, &tmR, bar_r + 8 * l_b, nt0 + l_cc * 32, w0, h0, n0);
                    if (++l_b == nbuf) l_b = 0;
                }
                ++l_qn;
                ++l_cc;
                if (l_cc == p.nchunks || nt0 + l_cc * 32 >= p.Cout) {
                    l_cc = 0;
                    if (++l_half == p.mh) { l_half = 0; l_tile += nct; }
                }
                if (mine) return;
            }
        };
        if (g_res && e0)
            for (int i = 0; i < p.nbuf - 1; ++i) issue_res_load();

        auto flush_stats = [&]() {
            named_bar(3, EPI_GROUPS * EPI_THREADS);
            for (int c = eg * EPI_THREADS + et; c < p.Cout; c += EPI_GROUPS * EPI_THREADS) {
                double *dst = p.stats + ((size_t)cur_n * p.stats_ld + c) * 2;
                atomicAdd(dst, sacc[0][c]);
                atomicAdd(dst + 1, sacc[1][c]);
                sacc[0][c] = 0.0;
                sacc[1][c] = 0.0;
                if (p.dual) {
                    double *dst2 = p.stats2 + ((size_t)cur_n * p.stats2_ld + c) * 2;
                    atomicAdd(dst2, sacc2[c]);
                    atomicAdd(dst2 + 1, sacc2[STATS_MAX_C + c]);
                    sacc2[c] = 0.0;
                    sacc2[STATS_MAX_C + c] = 0.0;
                }
            }
            named_bar(3, EPI_GROUPS * EPI_THREADS);
        };

        int it = 0;
        for (int tile = cid; tile < p.total_tiles; tile += nct, ++it) {
            const int as = p.acc_stages == 2 ? (it & 1) : 0;
            const uint32_t aph = (uint32_t)(p.acc_stages == 2 ? (it >> 1) : it) & 1u;
            const int nt0 = (tile % p.n_tiles) * p.n_tile;
            const int ks = tile / p.tiles_mn;
            const int mt = (tile - ks * p.tiles_mn) / p.n_tiles;
            { PROF_IF(4, pt); mbar_wait(bar_t_full + 8 * as, aph); }
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // pass 0 drains the accumulators; pass 1 (split-K with in-kernel reduction, last CTA of the tile only) re-runs
            // the chunk loop on the sum of the S partial tiles with the real bias / residual / rounding / statistics
            for (int fin_i = 0; fin_i <= (RED ? 1 : 0); ++fin_i) {
            const bool fin = RED && fin_i != 0;                       // compile-time false in the one-pass instantiations
            if (fin) {
                if (e0) bulk_wait_all();                              // this group's partial tile is in global memory
                asm volatile("fence.proxy.async;" ::: "memory");
                named_bar(3, EPI_GROUPS * EPI_THREADS);
                if (threadIdx.x == 96) {
                    __threadfence();
                    unsigned *ctr = p.red_ctr + (size_t)(tile - ks * p.tiles_mn) * p.pair + rank;
                    const unsigned old = atomicAdd(ctr, 1u);
                    s_last = old == (unsigned)(p.ksplit - 1);
                    if (s_last) *ctr = 0u;                            // every slice has arrived: ready for the next launch
                }
                named_bar(3, EPI_GROUPS * EPI_THREADS);
                if (!s_last) break;
                __threadfence();
                asm volatile("fence.proxy.async;" ::: "memory");
            }
            const float *const bias_c = fin ? p.red_bias : p.bias;
            const int yf16_c = fin ? p.red_yf16 : p.y_f16;
            const CUtensorMap *const tm_c = fin ? &tmF : tm_out;
            const bool stats_c = fin ? p.red_stats != nullptr : p.stats != nullptr;
            const int srows_c = fin ? p.red_rows : p.stats_rows;
            double *const gstats_c = fin ? p.red_stats : (plain_g ? p.stats2 : p.stats);
            const int gstats_ld_c = fin ? p.red_stats_ld : (plain_g ? p.stats2_ld : p.stats_ld);
            for (int half = 0; half < p.mh; ++half) {
                int w0, h0, n0;
                box_origin(p, mt, half, rank, w0, h0, n0);
                if (!fin && p.stats && !p.stats_rows && n0 != cur_n) {
                    if (cur_n >= 0) flush_stats();
                    cur_n = n0;
                }
                // pass 1: this thread's pixel inside the box -> its row of the partial tiles (and of the residual)
                size_t m_pix = 0;
                int nn_pix = 0;
                if (fin) {
                    int xr = row, yr = 0, nr = 0;
                    if (!p.halo) { xr = row % p.bw; yr = (row / p.bw) % p.bh; nr = row / (p.bw * p.bh); }
                    nn_pix = n0 + nr;
                    m_pix = ((size_t)nn_pix * p.H + (size_t)(h0 + yr)) * p.W + (size_t)(w0 + xr);
                }
                const uint32_t acc = tmem_base + ((uint32_t)(q * 32) << 16) +
                                     (uint32_t)((as * p.mh + half) * p.acc_stride);
                for (int cc = 0; cc < p.nchunks; ++cc, ++qn) {
                    const int nbase = nt0 + cc * 32;
                    if (nbase >= p.Cout) break;
                    if (!p.dual && (int)(qn & (EPI_GROUPS - 1)) != eg) continue;       // the other group's chunk
                    const uint32_t b = rb;
                    const uint32_t sbuf = smem_g + b * STAGE_BUF_BYTES;
                    const uint32_t srow = sbuf + (uint32_t)row * 128u;
                    float v[32];
                    if (!fin) {
                        tmem_ld32(acc + (uint32_t)(cc * 32), v);
                    } else {
                        // the S partial tiles in slice order (fixed order: bit-reproducible), then the residual
                        const float *src = p.red_ws + m_pix * (size_t)p.red_ldw + nbase;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 t = __ldcg(reinterpret_cast<const float4 *>(src) + j);
                            v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
                        }
                        for (int sl = 1; sl < p.ksplit; ++sl) {
                            src += p.red_slice;
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float4 t = __ldcg(reinterpret_cast<const float4 *>(src) + j);
                                v[4 * j] += t.x; v[4 * j + 1] += t.y; v[4 * j + 2] += t.z; v[4 * j + 3] += t.w;
                            }
                        }
                    }
                    if (g_res) {
                        PROF_IF(5, pt);
                        mbar_wait(bar_r + 8 * b, rph);
                    }
                    const float4 *bias4 = reinterpret_cast<const float4 *>(bias_c + nbase);
                    if (fin && p.red_res && nn_pix < p.B) {
                        // (the bias is added below, after the residual: the order of the one-pass epilogue is bias first --
                        // fp32 addition of three terms, so add in the same order: partial sum + bias, then + residual)
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            if (nbase + 4 * j < p.Cout) {
                                const float4 bz = bias_c ? __ldg(bias4 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                                const float4 r = __ldg(reinterpret_cast<const float4 *>(p.red_res + m_pix * (size_t)p.red_ldr + nbase) + j);
                                v[4 * j] = (v[4 * j] + bz.x) + r.x; v[4 * j + 1] = (v[4 * j + 1] + bz.y) + r.y;
                                v[4 * j + 2] = (v[4 * j + 2] + bz.z) + r.z; v[4 * j + 3] = (v[4 * j + 3] + bz.w) + r.w;
                                // bias consumed for these columns
                            }
                        }
                    }
                    const bool bias_on = bias_c != nullptr && !(fin && p.red_res && nn_pix < p.B);
                    if (!yf16_c) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 bz = bias_on ? __ldg(bias4 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                            float4 o = make_float4(v[4 * j] + bz.x, v[4 * j + 1] + bz.y, v[4 * j + 2] + bz.z,
                                                   v[4 * j + 3] + bz.w);
                            const uint32_t addr = srow + (((uint32_t)j ^ sw) << 4);
                            if (g_res) {
                                float4 r;
                                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                             : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                                             : "r"(addr));
                                o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
                            }
                            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(o.x), "f"(o.y),
                                         "f"(o.z), "f"(o.w)
                                         : "memory");
                        }
                    } else {
                        // fp16 output (the tensor is only ever read as a conv / attention operand): same fp32
                        // arithmetic, then one rounding; staged as [128 rows][64 B] in the SWIZZLE_64B pattern
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 bz = bias_on ? __ldg(bias4 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                            v[4 * j] += bz.x; v[4 * j + 1] += bz.y; v[4 * j + 2] += bz.z; v[4 * j + 3] += bz.w;
                            if (g_res) {
                                float4 r;
                                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                             : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                                             : "r"(srow + (((uint32_t)j ^ sw) << 4)));
                                v[4 * j] += r.x; v[4 * j + 1] += r.y; v[4 * j + 2] += r.z; v[4 * j + 3] += r.w;
                            }
                        }
                        if (g_res) named_bar(6 + eg, EPI_THREADS);   // all residual rows read before the fp16 tile lands
                        const uint32_t hrow = sbuf + (uint32_t)row * 64u, hsw = ((uint32_t)row >> 1) & 3u;
                        if (yf16_c == 2) {
                            // scaled hi | lo pair (a raw residual-stream operand of a high-precision conv): v * 2^-4 =
                            // hi + lo to ~22 bits; the hi tile is staged in the first 8 KB of the buffer, lo in the second
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] *= HL_OP_SCALE;
                        }
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const __half2 h0 = __floats2half2_rn(v[8 * j], v[8 * j + 1]), h1 = __floats2half2_rn(v[8 * j + 2], v[8 * j + 3]);
                            const __half2 h2 = __floats2half2_rn(v[8 * j + 4], v[8 * j + 5]), h3 = __floats2half2_rn(v[8 * j + 6], v[8 * j + 7]);
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(hrow + (((uint32_t)j ^ hsw) << 4)),
                                         "r"(*reinterpret_cast<const uint32_t *>(&h0)), "r"(*reinterpret_cast<const uint32_t *>(&h1)),
                                         "r"(*reinterpret_cast<const uint32_t *>(&h2)), "r"(*reinterpret_cast<const uint32_t *>(&h3))
                                         : "memory");
                            if (yf16_c == 2) {
                                const float2 f0 = __half22float2(h0), f1 = __half22float2(h1), f2 = __half22float2(h2), f3 = __half22float2(h3);
                                const __half2 l0 = __floats2half2_rn(v[8 * j] - f0.x, v[8 * j + 1] - f0.y);
                                const __half2 l1 = __floats2half2_rn(v[8 * j + 2] - f1.x, v[8 * j + 3] - f1.y);
                                const __half2 l2 = __floats2half2_rn(v[8 * j + 4] - f2.x, v[8 * j + 5] - f2.y);
                                const __half2 l3 = __floats2half2_rn(v[8 * j + 6] - f3.x, v[8 * j + 7] - f3.y);
                                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(hrow + 8192u + (((uint32_t)j ^ hsw) << 4)),
                                             "r"(*reinterpret_cast<const uint32_t *>(&l0)), "r"(*reinterpret_cast<const uint32_t *>(&l1)),
                                             "r"(*reinterpret_cast<const uint32_t *>(&l2)), "r"(*reinterpret_cast<const uint32_t *>(&l3))
                                             : "memory");
                            }
                        }
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    if (!g_res && e0) {
                        // the buffer the group's NEXT chunk writes must have been read out by its old store
                        PROF_IF(9, eg == 0);
                        if (nbuf == 2) bulk_wait_read<0>(); else if (nbuf == 3) bulk_wait_read<1>(); else bulk_wait_read<2>();
                    }
                    { PROF_IF(6, pt); named_bar(1 + eg, EPI_THREADS); }
                    if (e0) {
                        tma_store_4d(tm_c, sbuf, nbase, w0, h0, n0 + (fin ? 0 : ks * p.split_b));
                        if (yf16_c == 2) tma_store_4d(tm_c, sbuf + 8192u, p.Cout + nbase, w0, h0, n0);
                        bulk_commit();
                        if (g_res) {
                            { PROF_IF(9, eg == 0); bulk_wait_read<1>(); }   // previous store drained -> refill its buffer
                            issue_res_load();
                        }
                    }
                    if (stats_c) {
                        // column sums of the finished chunk: each thread sums one column over its 32-row quarter
                        // (fixed order), the four quarters are combined in a fixed order by the first warp of the
                        // group, and only then added (one uncontended fp64 atomic per channel) to the CTA totals
                        const int col = et & 31, rq = et >> 5;
                        float s = 0.f, ss = 0.f;
                        if (!yf16_c) {
#pragma unroll 8
                            for (int r = 0; r < 32; ++r) {
                                const int rr = rq * 32 + r;
                                const uint32_t addr = sbuf + (uint32_t)rr * 128u +
                                                      ((((uint32_t)col >> 2) ^ (uint32_t)(rr & 7)) << 4) +
                                                      (((uint32_t)col & 3u) << 2);
                                float x;
                                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(addr));
                                s += x;
                                ss = fmaf(x, x, ss);
                            }
                        } else {
                            // fp16 tile ([128 rows][64 B], SWIZZLE_64B pattern): the statistics of the ROUNDED values,
                            // i.e. of the tensor the following GroupNorm actually reads
#pragma unroll 8
                            for (int r = 0; r < 32; ++r) {
                                const int rr = rq * 32 + r;
                                const uint32_t addr = sbuf + (uint32_t)rr * 64u +
                                                      ((((uint32_t)col >> 3) ^ (((uint32_t)rr >> 1) & 3u)) << 4) +
                                                      (((uint32_t)col & 7u) << 1);
                                unsigned short hb;
                                asm volatile("ld.shared.u16 %0, [%1];" : "=h"(hb) : "r"(addr));
                                const float x = __half2float(__ushort_as_half(hb));
                                s += x;
                                ss = fmaf(x, x, ss);
                            }
                        }
                        spart[eg][rq][col] = make_float2(s, ss);
                        named_bar(4 + eg, EPI_THREADS);
                        if (srows_c) {
                            // several samples per box (8^2: rows 0-63 sample n0, 64-127 sample n0 + 1), or the final pass of
                            // a split-K tile: quarter sums of one sample are combined in a fixed order and added to the
                            // global fp64 row directly (a handful of CTAs per launch at these sizes: no contention worth
                            // a shared-memory stage)
                            const int per = srows_c >> 5;                      // quarters per sample: 1 | 2 | 4
                            if (rq % per == 0 && nbase + col < p.Cout) {
                                const int nn = n0 + rq / per;
                                if (nn < p.B) {
                                    float2 t = spart[eg][rq][col];
                                    for (int k = 1; k < per; ++k) { const float2 u = spart[eg][rq + k][col]; t.x += u.x; t.y += u.y; }
                                    double *dst = gstats_c + ((size_t)nn * gstats_ld_c + nbase + col) * 2;
                                    atomicAdd(dst, (double)t.x);
                                    atomicAdd(dst + 1, (double)t.y);
                                }
                            }
                        } else if (rq == 0 && nbase + col < p.Cout) {
                            const float2 p0 = spart[eg][0][col], p1 = spart[eg][1][col], p2 = spart[eg][2][col],
                                         p3 = spart[eg][3][col];
                            atomicAdd(&sacc_s[nbase + col], (double)((p0.x + p1.x) + (p2.x + p3.x)));
                            atomicAdd(&sacc_s[STATS_MAX_C + nbase + col], (double)((p0.y + p1.y) + (p2.y + p3.y)));
                        }
                    }
                    if (++rb == nbuf) { rb = 0; rph ^= 1u; }
                    ++ql;
                }
            }
            if (!fin) {
                // all tcgen05.ld of this tile have completed (wait::ld) -> hand the accumulators back
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                if (CTA2) mbar_arrive_cluster(mapa_u32(bar_t_empty + 8 * as, 0)); else mbar_arrive(bar_t_empty + 8 * as);
            }
            }   // fin
        }
        if (p.stats && cur_n >= 0) flush_stats();
        if (e0) bulk_wait_read<0>();
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (CTA2) cluster_sync_all(); else __syncthreads();     // no remote arrive / MMA may still target this CTA
    if (warp == 2) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (CTA2)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                         "r"((uint32_t)p.tmem_cols)
                         : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                         "r"((uint32_t)p.tmem_cols)
                         : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// split-K second pass.  The 8^2 / 16^2 layers have M = 256 / 1024 GEMM rows: 24-48 CTAs each streaming
// 100+ operand stages from L2 at the per-SM rate while 100 SMs idle.  With a workspace registered
// (hl_conv_set_workspace) the K loop is cut into S slices run by S x as many CTAs, every slice stores its
// fp32 partial tile into ws[s][B*H*W][cout_pad], and this kernel adds the slices in a FIXED order
// (deterministic), then bias, residual, the per-channel GroupNorm statistics and the output rounding
// exactly as the one-pass epilogue does.  One block = 32 channels of one pixel slab of one sample.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_splitk_reduce(const float *__restrict__ ws, int S, int64_t slice, int ldw,
                                                       const float *__restrict__ bias,
                                                       const float *__restrict__ res, int ldr, void *__restrict__ y,
                                                       int y_f16, int ldy, double *__restrict__ stats, int stats_ld,
                                                       int HW, int Cout, int slab) {
    hl_pdl_enter();
    // block = (32-channel group, sample); warp w, lane = (pixel lane pl, channel quad q): a warp reads 4 pixels x 128 B
    __shared__ double red[8][8][8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = lane & 7, pl = lane >> 3;
    const int c = blockIdx.x * 32 + q * 4, b = blockIdx.y;
    const bool live = c < Cout;
    double sm[4] = {0, 0, 0, 0}, sq[4] = {0, 0, 0, 0};
    if (live) {
        const float4 bz = __ldg(reinterpret_cast<const float4 *>(bias + c));
        // blockIdx.z = pixel slab of the sample (B = 1 at 32^2 would otherwise run 12 blocks of 32 serial rounds)
        const int pend = min(HW, ((int)blockIdx.z + 1) * slab);
#pragma unroll 2
        for (int pp = (int)blockIdx.z * slab + warp * 4 + pl; pp < pend; pp += 32) {
            const int64_t m = (int64_t)b * HW + pp;
            const float *src = ws + m * ldw + c;
            float4 a = __ldcs(reinterpret_cast<const float4 *>(src));
#pragma unroll 4
            for (int k = 1; k < S; ++k) {                      // fixed order: bit-reproducible
                const float4 t = __ldcs(reinterpret_cast<const float4 *>(src + k * slice));
                a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
            }
            a.x += bz.x; a.y += bz.y; a.z += bz.z; a.w += bz.w;
            if (res) {
                const float4 r = *reinterpret_cast<const float4 *>(res + m * ldr + c);
                a.x += r.x; a.y += r.y; a.z += r.z; a.w += r.w;
            }
            if (y_f16) {
                const float sc = y_f16 == 2 ? HL_OP_SCALE : 1.0f;           // 2: scaled hi | lo pair, lo at channel Cout + c
                const float4 as = make_float4(a.x * sc, a.y * sc, a.z * sc, a.w * sc);
                const __half2 h0 = __floats2half2_rn(as.x, as.y), h1 = __floats2half2_rn(as.z, as.w);
                uint2 w;
                w.x = *reinterpret_cast<const uint32_t *>(&h0);
                w.y = *reinterpret_cast<const uint32_t *>(&h1);
                *reinterpret_cast<uint2 *>(reinterpret_cast<__half *>(y) + m * ldy + c) = w;
                if (y_f16 == 2) {
                    const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
                    const __half2 l0 = __floats2half2_rn(as.x - f0.x, as.y - f0.y), l1 = __floats2half2_rn(as.z - f1.x, as.w - f1.y);
                    w.x = *reinterpret_cast<const uint32_t *>(&l0);
                    w.y = *reinterpret_cast<const uint32_t *>(&l1);
                    *reinterpret_cast<uint2 *>(reinterpret_cast<__half *>(y) + m * ldy + Cout + c) = w;
                } else {
                    const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);     // statistics of the rounded tensor
                    a = make_float4(f0.x, f0.y, f1.x, f1.y);
                }
            } else {
                *reinterpret_cast<float4 *>(reinterpret_cast<float *>(y) + m * ldy + c) = a;
            }
            sm[0] += a.x; sm[1] += a.y; sm[2] += a.z; sm[3] += a.w;
            sq[0] += (double)a.x * a.x; sq[1] += (double)a.y * a.y; sq[2] += (double)a.z * a.z; sq[3] += (double)a.w * a.w;
        }
    }
    if (!stats) return;
    // per-channel sums over the sample's pixels: pixel lanes (shuffles), then warps (shared memory), fixed order
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        sm[e] += __shfl_xor_sync(0xffffffffu, sm[e], 8);
        sm[e] += __shfl_xor_sync(0xffffffffu, sm[e], 16);
        sq[e] += __shfl_xor_sync(0xffffffffu, sq[e], 8);
        sq[e] += __shfl_xor_sync(0xffffffffu, sq[e], 16);
    }
    if (pl == 0) {
#pragma unroll
        for (int e = 0; e < 4; ++e) { red[warp][q][e] = sm[e]; red[warp][q][4 + e] = sq[e]; }
    }
    __syncthreads();
    if (threadIdx.x < 64) {                       // (quad, value) pairs: 8 quads x 8 values
        const int qq = threadIdx.x >> 3, v = threadIdx.x & 7;
        const int cc = blockIdx.x * 32 + qq * 4 + (v & 3);
        if (cc < Cout) {
            double t = 0.0;
#pragma unroll
            for (int w = 0; w < 8; ++w) t += red[w][qq][v];
            atomicAdd(stats + ((size_t)b * stats_ld + cc) * 2 + (v >> 2), t);    // fp64: slab order perturbs at 1e-16
        }
    }
}

// partial-sum workspaces, one per stream that issues split-K convolutions (two streams run concurrently)
constexpr int RED_CTR_BYTES = 8192;          // 2048 tile counters behind the partial sums
int g_tune_red = 0;                          // 1: second pass inside the conv kernel (experiment, measured slower); 0: the k_splitk_reduce launch
struct Workspace { cudaStream_t stream; float *ptr; size_t bytes; };
Workspace g_ws[8];
int g_n_ws = 0;

const Workspace *find_ws(cudaStream_t st) {
    for (int i = 0; i < g_n_ws; ++i)
        if (g_ws[i].stream == st && g_ws[i].ptr) return &g_ws[i];
    return nullptr;
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                    const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                    const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) ==
                cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)ptr;
    }
    return fn;
}

bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

struct Tiling {
    int bw, bh, bn;
};

// 128-pixel boxes {bw, bh, bn}: whole rows first, then rows, then samples.  Power-of-two maps only.
bool pick_tiling(int H, int W, Tiling *t) {
    int bw;
    if (W >= 128) {
        if (W % 128) return false;
        bw = 128;
    } else {
        if (!is_pow2(W)) return false;
        bw = W;
    }
    int rows = 128 / bw;
    int bh = rows < H ? rows : H;
    if (!is_pow2(bh) || H % bh) return false;
    t->bw = bw;
    t->bh = bh;
    t->bn = 128 / (bw * bh);
    return true;
}

// tuning overrides (-1 = automatic); set through hl_conv_set_tuning (tests / experiments)
int g_tune_split = -1;      // -1 automatic, 0 off, > 0 forced number of K slices (when it divides the chunk count)
int g_tune_mh = -1, g_tune_ntile = -1, g_tune_halo = -1, g_tune_epi_stats = -1, g_tune_base_off = -1;
unsigned long long *g_prof = nullptr;
int g_tune_stages = -1, g_tune_nbuf = -1, g_tune_cta2 = -1;   // experiment: cap on pipeline slots / staging buffers

struct Plan {
    TcParams p;
    Tiling t;
    int cout_pad;
    size_t smem;
    int grid;
};

int acc_stride_for(int n_tile) {
    int s = 32;
    while (s < n_tile) s <<= 1;
    return s;
}

bool make_plan(int kind, int B, int H, int W, int Cin, int Cout, int ksize, int stride, bool has_res,
               bool want_stats, Plan *pl, int npass = 1, int reserve = 0 /* bytes of dynamic shared memory kept free */) {
    // H, W are OUTPUT dims here
    Tiling t;
    if (!pick_tiling(H, W, &t)) return false;
    const int chunk = kind ? 64 : 32;
    if (Cin % chunk) return false;
    TcParams &p = pl->p;
    p.B = B; p.H = H; p.W = W; p.Cout = Cout; p.stride = stride; p.ksize = ksize;
    p.taps = ksize * ksize;
    p.kchunks = Cin / chunk;
    p.npass = npass;
    p.vchunks = npass * p.kchunks;
    for (int q = 0; q < 3; ++q) { p.a_off[q] = 0; p.b_slab[q] = 0; }
    p.kind = kind;
    p.bw = t.bw; p.bh = t.bh; p.bn = t.bn;
    p.tiles_w = W / t.bw;
    p.tiles_h = H / t.bh;
    const int tiles_n = (B + t.bn - 1) / t.bn;
    const int boxes = p.tiles_w * p.tiles_h * tiles_n;
    const int cout_pad = hl_conv_cout_pad(Cout);
    pl->cout_pad = cout_pad;
    pl->t = t;
    const int sms = hl_num_sms();

    bool halo_ok = ksize == 3 && stride == 1 && W % 128 == 0;
    int halo = halo_ok ? 1 : 0;
    if (g_tune_halo >= 0) halo = halo_ok && g_tune_halo;

    // halves per CTA tile: 2 shares each weight tile between two 128-pixel boxes
    int mh = 1;
    if (halo ? (H % 2 == 0) : (boxes % 2 == 0)) mh = 2;
    if (g_tune_mh == 1) mh = 1;
    if (g_tune_mh == 2 && !(halo ? (H % 2 == 0) : (boxes % 2 == 0))) mh = 1;

    // (pair, mh, n_tile) -- measured on B200 (profiles/r1_conv_tilings.md):
    //  * a tcgen05.mma with M = 128 per SM costs ~(128 + N_sm)/2 cycles of operand fetch from shared memory
    //    (N_sm = B rows held by that SM), so a CTA pair (cta_group::2, M = 256, N_sm = N/2) beats a single
    //    CTA for every 3x3 shape tried and ties on 1x1: use it whenever the boxes pair up;
    //  * a double-buffered TMEM accumulator (epilogue overlapped with the next main loop) beats sharing a
    //    weight tile between two halves of one CTA, so mh = 2 is only reachable through the tuning hook;
    //  * big problems: the widest legal N with at least one tile per SM; small problems (one wave or less):
    //    the N that yields the most CTAs without exceeding one wave (their K loops are latency-bound, more
    //    CTAs = more TMA streams), ties to the wider tile.
    static const int cand[] = {256, 192, 128, 96, 64, 32};
    const bool pair_ok = (halo ? (H % 2 == 0) : (boxes % 2 == 0)) && g_tune_mh != 2;
    int pair = (pair_ok && g_tune_cta2 != 0) ? 2 : 1;
    const int mh_max = mh;
    mh = (g_tune_mh == 2) ? mh_max : 1;
    if (mh == 2) pair = 1;
    auto legal = [&](int c) {
        if (cout_pad % c || c % 32) return false;
        if (mh * acc_stride_for(c) > 512) return false;
        if (pair == 2 && ((c / 2) % 8 || c < 64)) return false;
        return true;
    };
    // N tile: the candidate with the cheapest K loop under the shared-memory model of DESIGN.md 5.1 -- rounds of
    // tiles per CTA (pair) x bytes through shared memory per pipeline stage (MMA operand fetch of 4 instructions
    // + the TMA writes of the weight rows and of the activation box / halo slab); the stage count is the same for
    // every candidate.  Large problems get the widest tile (bytes per output column fall with N), problems of one
    // to two waves the tile that avoids a second round (32^2: N = 128 in one round beats N = 64 in two: measured
    // 26 vs 37 us), sub-wave problems the narrowest (most CTAs; split-K then takes over).  Ties go to the wider.
    auto tiles_of = [&](int c) { return (boxes / (mh * pair)) * (cout_pad / c); };
    auto stage_bytes = [&](int c) {
        const double rows_b = (double)c / pair;
        const double a_bytes = halo ? (double)HALO_ROW_BYTES * (mh + 2) / 9.0 : (double)mh * A_BOX_BYTES;
        return 4.0 * (128.0 * mh + rows_b) * 32.0 + rows_b * ROW_BYTES + a_bytes;
    };
    int n_tile = 0;
    {
        const int units = pair == 2 ? sms / 2 : sms;
        double best = 0.0;
        for (int c : cand) {
            if (!legal(c) || c < 64 || c == 96) continue;         // swept: 64 / 128 / 192 / 256
            const int rounds = (tiles_of(c) + units - 1) / units;
            const double cost = rounds * stage_bytes(c);
            if (n_tile == 0 || cost < best) { best = cost; n_tile = c; }
        }
        if (n_tile == 0)
            for (int c : cand)
                if (legal(c)) { n_tile = c; break; }
    }
    if (n_tile == 0 && pair == 2) {                      // e.g. Cout_pad = 32: no legal pair tile
        pair = 1;
        for (int c : cand)
            if (legal(c)) { n_tile = c; break; }
    }
    if (g_tune_ntile > 0 && legal(g_tune_ntile)) n_tile = g_tune_ntile;
    if (n_tile == 0) return false;

    p.halo = halo;
    p.prof = g_prof;
    p.base_off = g_tune_base_off == 1 ? 1 : 0;   // measured on B200: the swizzle XOR uses absolute smem address bits
    p.mh = mh;
    p.pair = pair;
    p.hp = halo ? H / (mh * pair) : 1;
    p.n_tile = n_tile;
    p.n_tiles = cout_pad / n_tile;
    p.nchunks = n_tile / 32;
    p.total_tiles = (boxes / (mh * pair)) * p.n_tiles;
    p.acc_stride = acc_stride_for(n_tile);
    p.acc_stages = 512 / (mh * p.acc_stride) >= 2 ? 2 : 1;
    int cols = 32;
    while (cols < p.acc_stages * mh * p.acc_stride) cols <<= 1;
    p.tmem_cols = cols;

    // shared-memory budget: staging ring, then A slots, the rest to B slots
    const int budget = DYN_SMEM_MAX - 1024 /*alignment slack*/ - reserve;
    p.b_slot_bytes = (n_tile / pair) * ROW_BYTES;
    p.has_res = has_res ? 1 : 0;
    // staging buffers PER epilogue group: 2; the HBM-bound 1x1 convs with a residual want a deeper residual
    // prefetch (measured 0.124 -> 0.102 ms on 192->192 @ 256^2) and have the shared memory to spare
    int nbuf = (has_res && ksize == 1) ? 4 : 2;    // 3x3: measured 2 > 3 (the pipeline slots matter more)
    if (g_tune_nbuf >= 2 && g_tune_nbuf <= 4) nbuf = g_tune_nbuf;
    for (;;) {
        int rest = budget - EPI_GROUPS * nbuf * STAGE_BUF_BYTES;
        if (halo) {
            p.a_slot_bytes = HALO_SLOT_BYTES;
            p.a_slots = mh + 3;
            int b_slots = (rest - p.a_slots * p.a_slot_bytes) / p.b_slot_bytes;
            if (b_slots < 3 && p.a_slots > mh + 2) {
                p.a_slots = mh + 2;
                b_slots = (rest - p.a_slots * p.a_slot_bytes) / p.b_slot_bytes;
            }
            p.b_slots = b_slots > MAX_SLOTS ? MAX_SLOTS : b_slots;
            if (g_tune_stages >= 2 && p.b_slots > g_tune_stages) p.b_slots = g_tune_stages;
        } else {
            p.a_slot_bytes = mh * A_BOX_BYTES;
            int stages = rest / (p.a_slot_bytes + p.b_slot_bytes);
            if (stages > MAX_SLOTS) stages = MAX_SLOTS;
            if (g_tune_stages >= 2 && stages > g_tune_stages) stages = g_tune_stages;
            p.a_slots = p.b_slots = stages;
        }
        if (p.b_slots >= 3 || nbuf == 2) break;
        --nbuf;                        // trade residual prefetch depth for pipeline depth
    }
    if (p.b_slots < 2 || p.a_slots < (halo ? mh + 2 : 2)) return false;
    p.nbuf = nbuf;
    pl->smem = (size_t)p.a_slots * p.a_slot_bytes + (size_t)p.b_slots * p.b_slot_bytes +
               (size_t)EPI_GROUPS * nbuf * STAGE_BUF_BYTES + 1024;

    p.stats = nullptr;     // filled by the caller when the epilogue computes the statistics
    p.stats_ld = 0;
    if (pair == 2) {
        const int clusters = p.total_tiles < sms / 2 ? p.total_tiles : sms / 2;
        pl->grid = 2 * clusters;
    } else {
        pl->grid = p.total_tiles < sms ? p.total_tiles : sms;
    }
    (void)want_stats;
    p.ksplit = 1;
    p.kc_split = p.vchunks;
    p.tiles_mn = p.total_tiles;
    p.split_b = 0;
    return true;
}

// split-K planning: 3x3 layers whose K loop is long and whose grid is under one wave.  Every (N tile, S) with at
// most one wave of CTAs is costed with the shared-memory model of DESIGN.md 5.1 -- stages x bytes through shared
// memory per stage (MMA operand fetch + TMA operand writes) -- and the cheapest is taken if it at least halves the
// K-loop cost of the one-pass plan (the second pass is not free).  May replace `pl` by a plan with a wider N tile.
// H, W are OUTPUT dims.  Returns the number of K slices (1 = no split).
inline int split_batch(const Plan &pl, int B) { return (B + pl.t.bn - 1) / pl.t.bn * pl.t.bn; }

int choose_split(Plan &pl, int kind, int B, int H, int W, int Cin, int Cout, int ksize, int stride, size_t ws_bytes) {
    TcParams &p = pl.p;
    int S = 1;
    if (!ws_bytes || g_tune_split == 0 || Cout % 4 || p.vchunks < 2) return 1;
    const int sms = hl_num_sms();
    // a slice holds whole batch boxes: B rounded up to the box's sample count (the padding rows receive the zero
    // rows TMA fills in for samples beyond B and are never read back)
    const size_t slice_bytes = (size_t)split_batch(pl, B) * H * W * pl.cout_pad * sizeof(float);
    auto fits = [&](const Plan &q, int s) {
        return q.p.vchunks % s == 0 && (size_t)s * slice_bytes <= ws_bytes && q.grid * s <= sms &&
               (q.p.vchunks / s) * q.p.taps >= 12;
    };
    auto cost = [&](const Plan &q, int s) {        // bytes through shared memory on one CTA's K loop
        const double rows_b = (double)q.p.n_tile / q.p.pair;
        const double stage = 4.0 * (128.0 * q.p.mh + rows_b) * 32.0 + rows_b * ROW_BYTES + (double)q.p.mh * A_BOX_BYTES;
        const int units = q.p.pair == 2 ? q.grid / 2 : q.grid;
        const double tiles = (double)((q.p.total_tiles + units - 1) / units);
        return tiles * (q.p.vchunks / s) * q.p.taps * stage;
    };
    if (g_tune_split > 1) {
        if (p.vchunks % g_tune_split == 0 && (size_t)g_tune_split * slice_bytes <= ws_bytes) S = g_tune_split;
    } else if (ksize == 3 && !p.halo && pl.grid < sms && p.vchunks * p.taps >= 54) {
#ifndef HL_SPLIT_GAIN
#define HL_SPLIT_GAIN 0.5
#endif
        double best = cost(pl, 1) * HL_SPLIT_GAIN;
        Plan best_pl = pl;
        const int keep_ntile = g_tune_ntile;
        static const int widths[] = {0, 128, 192, 256};          // 0 = the one-pass plan's own N tile
        for (int wdt : widths) {
            Plan q = pl;
            if (wdt) {
                if (keep_ntile > 0) continue;                       // a forced tile is not second-guessed
                q = Plan{};
                g_tune_ntile = wdt;
                const bool ok = make_plan(kind, B, H, W, Cin, Cout, ksize, stride, false, false, &q, p.npass);
                g_tune_ntile = keep_ntile;
                if (ok) for (int e = 0; e < 3; ++e) { q.p.a_off[e] = p.a_off[e]; q.p.b_slab[e] = p.b_slab[e]; }
                if (!ok || q.p.n_tile != wdt || q.p.halo || q.t.bn != pl.t.bn) continue;
            }
            for (int s = 8; s >= 2; --s) {
                if (!fits(q, s)) continue;
                const double c = cost(q, s);
                if (c < best) { best = c; best_pl = q; S = s; }
            }
        }
        if (S > 1) pl = best_pl;
    }
    return S;
}

// in-epilogue GroupNorm statistics need one sample per 128-pixel box (every box row valid)
bool plan_epi_stats(const Plan &pl, bool want_stats) {
    if (!want_stats || g_tune_epi_stats == 0) return false;
    if (pl.cout_pad > STATS_MAX_C) return false;
    // one sample per box, or whole 32-row quarters per sample (8^2: two samples of 64 rows; 4^2 and below: stats kernel)
    return pl.t.bn == 1 || (128 / pl.t.bn) % 32 == 0;
}
int plan_stats_rows(const Plan &pl) { return pl.t.bn == 1 ? 0 : 128 / pl.t.bn; }

}  // namespace

extern "C" int hl_conv_cout_pad(int Cout) { return (Cout + 31) / 32 * 32; }

extern "C" int hl_conv_set_tuning(int mh, int n_tile, int halo, int epi_stats, int base_off) {
    g_tune_base_off = base_off;
    g_tune_mh = mh;
    g_tune_ntile = n_tile;
    g_tune_halo = halo;
    g_tune_epi_stats = epi_stats;
    return HL_OK;
}

extern "C" int hl_conv_set_tuning2(int max_stages, int nbuf, int cta2) {
    g_tune_stages = max_stages;
    g_tune_nbuf = nbuf;
    g_tune_cta2 = cta2;
    return HL_OK;
}

extern "C" int hl_conv_set_workspace(void *ws, int64_t bytes, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    HL_CHECK_ARG(((uintptr_t)ws & 15) == 0 && bytes >= 0);
    // the last RED_CTR_BYTES hold the tile counters of the in-kernel reduction: zero now, and every launch leaves them zero
    if (ws) {
        HL_CHECK_ARG(bytes > RED_CTR_BYTES && bytes % 16 == 0);
        HL_CHECK_CUDA(cudaMemsetAsync((char *)ws + bytes - RED_CTR_BYTES, 0, RED_CTR_BYTES, st));
        bytes -= RED_CTR_BYTES;
    }
    for (int i = 0; i < g_n_ws; ++i)
        if (g_ws[i].stream == st) {
            g_ws[i].ptr = (float *)ws;
            g_ws[i].bytes = ws ? (size_t)bytes : 0;
            return HL_OK;
        }
    if (!ws) return HL_OK;
    for (int i = 0; i < g_n_ws; ++i)
        if (!g_ws[i].ptr) {                      // reuse an unregistered slot
            g_ws[i] = {st, (float *)ws, (size_t)bytes};
            return HL_OK;
        }
    if (g_n_ws == 8) {
        hl_set_error("hl_conv_set_workspace: more than 8 streams registered");
        return HL_E_INVALID;
    }
    g_ws[g_n_ws++] = {st, (float *)ws, (size_t)bytes};
    return HL_OK;
}

extern "C" int hl_conv_set_split(int ksplit) {
    g_tune_split = ksplit;
    return HL_OK;
}

extern "C" int hl_conv_set_split_reduce(int in_kernel) {
    g_tune_red = in_kernel;
    return HL_OK;
}

extern "C" int hl_conv2d_plan_info(int x_dtype, int B, int H, int W, int Cin, int Cout, int ksize, int stride,
                                   int has_res, int want_stats, int64_t ws_bytes, int *out) {
    HL_CHECK_ARG(out && B > 0 && H > 0 && W > 0 && stride > 0);
    for (int i = 0; i < 16; ++i) out[i] = 0;
    const int kind = x_dtype == HL_DT_F16 ? 1 : 0;
    Plan pl = {};
    if ((stride != 1 && stride != 2) || (ksize != 1 && ksize != 3) || (stride == 2 && (ksize != 3 || H % 2 || W % 2)) ||
        !make_plan(kind, B, H / stride, W / stride, Cin, Cout, ksize, stride, has_res != 0, want_stats != 0, &pl, 1,
                   has_res == 2 ? 2 * STATS_MAX_C * (int)sizeof(double) : 0))      // has_res == 2: the dual-output launch
        return HL_OK;                                   // out[0] = 0: the CUDA-core kernel serves this shape
    const int S = has_res == 2 ? 1 : choose_split(pl, kind, B, H / stride, W / stride, Cin, Cout, ksize, stride,
                                                  ws_bytes > 0 ? (size_t)ws_bytes : 0);
    const TcParams &p = pl.p;
    const int units = p.pair == 2 ? pl.grid / 2 : pl.grid;
    int grid = pl.grid;
    if (S > 1) {
        const int total = p.total_tiles * S, sms = hl_num_sms();
        grid = p.pair == 2 ? 2 * (total < sms / 2 ? total : sms / 2) : (total < sms ? total : sms);
    }
    (void)units;
    out[0] = 1; out[1] = p.pair; out[2] = p.mh; out[3] = p.n_tile; out[4] = p.halo; out[5] = p.a_slots;
    out[6] = p.b_slots; out[7] = p.nbuf; out[8] = p.acc_stages; out[9] = p.tmem_cols; out[10] = (int)pl.smem;
    out[11] = grid; out[12] = p.total_tiles * S; out[13] = S; out[14] = p.vchunks / S;
    out[15] = plan_epi_stats(pl, want_stats != 0 && S == 1) ? 1 : 0;
    return HL_OK;
}

extern "C" int hl_conv_set_profile(void *dev_counters) {
    g_prof = (unsigned long long *)dev_counters;
    return HL_OK;
}

bool hl_conv_tc_applicable(int x_dtype, int B, int H, int W, int Cin, int Cout, int ksize, int stride,
                           int ldx, int ldy, int flags) {
    if (flags & (HL_CONV_FORCE_SIMT | HL_CONV_UPSAMPLE2X)) return false;
    if ((stride != 1 && stride != 2) || (ksize != 1 && ksize != 3)) return false;
    if (stride == 2 && (ksize != 3 || (H % 2) || (W % 2))) return false;
    if (x_dtype == HL_DT_F32 && !(flags & HL_CONV_TF32)) return false;
    const int esz = x_dtype == HL_DT_F16 ? 2 : 4;
    if ((ldx * esz) % 16 || Cin > ldx || ldy % ((flags & (HL_CONV_OUT_F16 | HL_CONV_OUT_F16_SPLIT)) ? 8 : 4)) return false;
    if ((flags & (HL_CONV_SPLIT3 | HL_CONV_SPLIT2P | HL_CONV_SPLIT2A)) && x_dtype != HL_DT_F16) return false;
    if ((flags & (HL_CONV_SPLIT3 | HL_CONV_SPLIT2A)) && ldx < 2 * Cin) return false;
    Plan pl = {};
    if (!make_plan(x_dtype == HL_DT_F16 ? 1 : 0, B, H / stride, W / stride, Cin, Cout, ksize, stride, false, false,
                   &pl))
        return false;
    return get_encode() != nullptr;
}

int hl_gn_stats_launch(const void *x, int x_f16, int ldx, int B, int HW, int C, double *stats, int stats_ld,
                       cudaStream_t stream);

int hl_conv2d_tc(const void *x, int x_dtype, int ldx, const void *wpk, const float *bias,
                 const float *residual, int ldr, void *y, int y_f16, int ldy, double *stats, int stats_ld, int B,
                 int Hin, int Win, int Cin, int Cout, int ksize, int stride, int flags, cudaStream_t stream,
                 float *y2 = nullptr, int ldy2 = 0, double *stats2 = nullptr, int stats2_ld = 0) {
    PFN_encodeTiled encode = get_encode();
    if (!encode) {
        hl_set_error("cuTensorMapEncodeTiled unavailable");
        return HL_E_CUDA;
    }
    const int kind = x_dtype == HL_DT_F16 ? 1 : 0;
    const int esz = kind ? 2 : 4;
    const int chunk = kind ? 64 : 32;
    const int H = Hin / stride, W = Win / stride;
    // hi + lo operand passes (HL_CONV_SPLIT3: x = [hi | lo] with lo at channel Cin, weights {W_hi, W_lo}: hi.hi +
    // lo.hi + hi.lo; HL_CONV_SPLIT2P: hi and lo packed inside the Cin channels, weights {[W_hi | W_hi], [W_lo | 0]})
    // HL_CONV_SPLIT2A: x = [hi | lo] as SPLIT3, one weight slab: hi.W + lo.W (the activation pair only)
    const int npass = (flags & HL_CONV_SPLIT3) ? 3 : (flags & (HL_CONV_SPLIT2P | HL_CONV_SPLIT2A)) ? 2 : 1;
    const int nslab = (flags & (HL_CONV_SPLIT3 | HL_CONV_SPLIT2P)) ? 2 : 1;
    Plan pl = {};
    constexpr int DUAL_RESERVE = 2 * STATS_MAX_C * (int)sizeof(double);
    HL_CHECK_ARG(make_plan(kind, B, H, W, Cin, Cout, ksize, stride, residual != nullptr, stats != nullptr, &pl, npass,
                           y2 ? DUAL_RESERVE : 0));
    if (npass == 3) { pl.p.a_off[1] = Cin; pl.p.b_slab[2] = 1; }
    if (flags & HL_CONV_SPLIT2P) pl.p.b_slab[1] = 1;
    if (flags & HL_CONV_SPLIT2A) pl.p.a_off[1] = Cin;
    HL_CHECK_ARG(npass == 1 || kind == 1);
    HL_CHECK_ARG(y_f16 != 2 || (Cout % 32 == 0 && ldy >= 2 * Cout));
    HL_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)wpk & 15) == 0 && ((uintptr_t)y & 15) == 0 &&
                 bias != nullptr && ((uintptr_t)bias & 15) == 0);
    HL_CHECK_ARG(!residual || (((uintptr_t)residual & 15) == 0 && ldr % 4 == 0));
    TcParams &p = pl.p;
    p.bias = bias;
    p.y_f16 = y_f16;
    HL_CHECK_ARG(!(y_f16 == 2 && stats));
    bool epi_stats = plan_epi_stats(pl, stats != nullptr);
    // dual output: y = conv + residual, y2 = conv (one pass over the operands; see TcParams::dual)
    const bool dual = y2 != nullptr;
    HL_CHECK_ARG(!dual || (residual && !y_f16 && ((uintptr_t)y2 & 15) == 0 && ldy2 % 4 == 0 && ldy2 >= Cout &&
                           (stats == nullptr) == (stats2 == nullptr)));
    p.dual = dual ? 1 : 0;
    p.stats2 = nullptr;
    p.stats2_ld = 0;
    p.sacc2_off = 0;
    if (dual) {
        p.sacc2_off = (int)pl.smem - 1024;            // pl.smem = rings + staging + 1024 B of alignment slack
        pl.smem += DUAL_RESERVE;
        HL_CHECK_ARG(pl.smem <= (size_t)DYN_SMEM_MAX);
    }

    // split-K (see choose_split / k_splitk_reduce)
    const Workspace *wsp = find_ws(stream);
    const float *keep_bias = p.bias;
    const int S = dual ? 1 : choose_split(pl, kind, B, H, W, Cin, Cout, ksize, stride, wsp ? wsp->bytes : 0);
    p.bias = keep_bias;
    p.y_f16 = y_f16;
    const void *y_final = y;
    const int ldy_final = ldy, yf16_final = y_f16;
    p.red = 0;
    p.red_ctr = nullptr;
    if (S > 1) {
        p.ksplit = S;
        p.kc_split = p.vchunks / S;
        p.split_b = split_batch(pl, B);
        p.total_tiles = p.tiles_mn * S;
        const int sms = hl_num_sms();
        if (p.pair == 2) {
            const int clusters = p.total_tiles < sms / 2 ? p.total_tiles : sms / 2;
            pl.grid = 2 * clusters;
        } else {
            pl.grid = p.total_tiles < sms ? p.total_tiles : sms;
        }
        p.bias = nullptr;          // bias, residual, statistics and rounding move to the second pass
        p.has_res = 0;
        p.y_f16 = 0;
        epi_stats = false;
        y = wsp->ptr;
        ldy = pl.cout_pad;
        y_f16 = 0;
        // second pass inside the kernel (the last CTA of a tile reduces): whole 32-row quarters per sample, counters fit
        const int rows = 128 / pl.t.bn;
        if (g_tune_red == 1 && rows % 32 == 0 && rows >= 64 && p.tiles_mn * p.pair <= RED_CTR_BYTES / 4 &&
            (!residual || ldr % 4 == 0)) {
            p.red = 1;
            p.red_ws = wsp->ptr;
            p.red_slice = (long long)split_batch(pl, B) * H * W * pl.cout_pad;
            p.red_ldw = pl.cout_pad;
            p.red_bias = keep_bias;
            p.red_res = residual;
            p.red_ldr = ldr;
            p.red_yf16 = yf16_final;
            p.red_stats = stats;
            p.red_stats_ld = stats_ld;
            p.red_rows = rows;
            p.red_ctr = reinterpret_cast<unsigned *>(reinterpret_cast<char *>(wsp->ptr) + wsp->bytes);
        }
    }
    p.stats_rows = 0;
    if (epi_stats) {
        p.stats = stats;
        p.stats_ld = stats_ld;
        p.stats2 = stats2;
        p.stats2_ld = stats2_ld;
        p.stats_rows = plan_stats_rows(pl);
    }
    const CUtensorMapDataType dt = kind ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;

    CUtensorMap tmA, tmB, tmY, tmR, tmY2, tmF;
    {
        cuuint64_t gdim[4] = {(cuuint64_t)((flags & (HL_CONV_SPLIT3 | HL_CONV_SPLIT2A)) ? 2 * Cin : Cin), (cuuint64_t)Win,
                              (cuuint64_t)Hin, (cuuint64_t)B};
        cuuint64_t gstr[3] = {(cuuint64_t)ldx * esz, (cuuint64_t)Win * ldx * esz,
                              (cuuint64_t)Hin * Win * ldx * esz};
        cuuint32_t box[4], estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
        if (p.halo) {
            box[0] = chunk; box[1] = HALO_PIX; box[2] = 1; box[3] = 1;
        } else {
            box[0] = chunk; box[1] = pl.t.bw * stride; box[2] = pl.t.bh * stride; box[3] = pl.t.bn;
        }
        CUresult r = encode(&tmA, dt, 4, (void *)x, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            hl_set_error("cuTensorMapEncodeTiled(A) failed: %d (B=%d H=%d W=%d Cin=%d ldx=%d stride=%d halo=%d)",
                         (int)r, B, Hin, Win, Cin, ldx, stride, p.halo);
            return HL_E_CUDA;
        }
    }
    {
        cuuint64_t gdim[3] = {(cuuint64_t)Cin, (cuuint64_t)pl.cout_pad, (cuuint64_t)(p.taps * nslab)};
        cuuint64_t gstr[2] = {(cuuint64_t)Cin * esz, (cuuint64_t)pl.cout_pad * Cin * esz};
        cuuint32_t box[3] = {(cuuint32_t)chunk, (cuuint32_t)(p.n_tile / p.pair), 1};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = encode(&tmB, dt, 3, (void *)wpk, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            hl_set_error("cuTensorMapEncodeTiled(B) failed: %d (Cin=%d Cout_pad=%d taps=%d)", (int)r, Cin,
                         pl.cout_pad, p.taps);
            return HL_E_CUDA;
        }
    }
    for (int which = 0; which < 4; ++which) {
        // 0: output (the partial-sum workspace when K is split), 1: residual, 2: second output (dual), 3: the real
        // output of a split-K launch that reduces inside the kernel
        const void *ptr = which == 1 ? (const void *)residual : which == 2 ? (const void *)y2
                        : which == 3 ? (p.red ? y_final : nullptr) : (const void *)y;
        const int ld = which == 1 ? ldr : which == 2 ? ldy2 : which == 3 ? ldy_final : ldy;
        const int yf = which == 0 ? y_f16 : which == 3 ? yf16_final : 0;
        const bool f16 = yf != 0;
        const int esz_o = f16 ? 2 : 4;
        CUtensorMap *tm = which == 1 ? &tmR : which == 2 ? &tmY2 : which == 3 ? &tmF : &tmY;
        if (!ptr || (which == 1 && S > 1)) { *tm = tmY; continue; }
        const bool wsmap = which == 0 && S > 1;
        cuuint64_t gdim[4] = {(cuuint64_t)(wsmap ? pl.cout_pad : yf == 2 ? 2 * Cout : Cout), (cuuint64_t)W,
                              (cuuint64_t)H, (cuuint64_t)(wsmap ? split_batch(pl, B) * S : B)};
        cuuint64_t gstr[3] = {(cuuint64_t)ld * esz_o, (cuuint64_t)W * ld * esz_o, (cuuint64_t)H * W * ld * esz_o};
        cuuint32_t box[4] = {32, (cuuint32_t)pl.t.bw, (cuuint32_t)pl.t.bh, (cuuint32_t)pl.t.bn};
        if (p.halo) { box[1] = BLOCK_M; box[2] = 1; box[3] = 1; }
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = encode(tm, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void *)ptr,
                            gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            f16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            hl_set_error("cuTensorMapEncodeTiled(%s) failed: %d (B=%d H=%d W=%d Cout=%d ld=%d)",
                         which == 1 ? "residual" : which == 2 ? "second output" : which == 3 ? "final output" : "output", (int)r, B, H, W, Cout, ld);
            return HL_E_CUDA;
        }
    }
    // cudaFuncSetAttribute is per device: remember which ordinals have been configured
    static bool configured[64] = {};
    int dev_ord = 0;
    HL_CHECK_CUDA(cudaGetDevice(&dev_ord));
    bool &smem_configured = configured[dev_ord & 63];
    if (!smem_configured) {
        HL_CHECK_CUDA(cudaFuncSetAttribute(k_conv_tc<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, DYN_SMEM_MAX));
        HL_CHECK_CUDA(cudaFuncSetAttribute(k_conv_tc<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, DYN_SMEM_MAX));
        HL_CHECK_CUDA(cudaFuncSetAttribute(k_conv_tc<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, DYN_SMEM_MAX));
        HL_CHECK_CUDA(cudaFuncSetAttribute(k_conv_tc<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, DYN_SMEM_MAX));
        smem_configured = true;
    }
    HL_CHECK_ARG(pl.smem <= (size_t)DYN_SMEM_MAX);
    if (p.pair == 2) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(pl.grid);
        cfg.blockDim = dim3(NUM_THREADS);
        cfg.dynamicSmemBytes = pl.smem;
        cfg.stream = stream;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = hl_pdl_attr(attr, 1);
        if (p.red) HL_CHECK_CUDA(cudaLaunchKernelEx(&cfg, k_conv_tc<true, true>, tmA, tmB, tmY, tmR, tmY2, tmF, p));
        else HL_CHECK_CUDA(cudaLaunchKernelEx(&cfg, k_conv_tc<true, false>, tmA, tmB, tmY, tmR, tmY2, tmF, p));
    } else {
        if (p.red) HL_CHECK_CUDA(hl_launch(k_conv_tc<false, true>, dim3(pl.grid), dim3(NUM_THREADS), pl.smem, stream, tmA, tmB, tmY, tmR, tmY2, tmF, p));
        else HL_CHECK_CUDA(hl_launch(k_conv_tc<false, false>, dim3(pl.grid), dim3(NUM_THREADS), pl.smem, stream, tmA, tmB, tmY, tmR, tmY2, tmF, p));
    }
    HL_CHECK_LAUNCH();
    if (S > 1 && p.red) return HL_OK;          // the last CTA of every tile has run the second pass
    if (S > 1) {
        const int HW = H * W;
        // pixel slabs (multiples of 32 pixels) until the grid covers the SMs about twice
        int slabs = 1;
        while (hl_cdiv(Cout, 32) * B * slabs < 2 * hl_num_sms() && HW / (2 * slabs) >= 64) slabs *= 2;
        const int slab = hl_cdiv(hl_cdiv(HW, slabs), 32) * 32;
        dim3 grid(hl_cdiv(Cout, 32), B, hl_cdiv(HW, slab));
        HL_CHECK_CUDA(hl_launch(k_splitk_reduce, grid, dim3(256), 0, stream, (const float *)wsp->ptr, S,
                                (int64_t)split_batch(pl, B) * HW * pl.cout_pad, pl.cout_pad, bias, residual, ldr, (void *)y_final,
                                yf16_final, ldy_final, stats, stats_ld, HW, Cout, slab));
        return HL_OK;
    }
    if (stats && !epi_stats) {
        int rc = hl_gn_stats_launch(y, y_f16, ldy, B, H * W, Cout, stats, stats_ld, stream);
        if (rc != HL_OK || !dual) return rc;
        return hl_gn_stats_launch(y2, 0, ldy2, B, H * W, Cout, stats2, stats2_ld, stream);
    }
    return HL_OK;
}

// ---------------------------------------------------------------------------------------------
// public entry: dispatch
// ---------------------------------------------------------------------------------------------
int hl_conv2d_simt(const void *x, int x_dtype, int ldx, const void *wpk, const float *bias,
                   const float *residual, int ldr, void *y, int ldy, int B, int H, int W, int Cin, int Cout,
                   int ksize, int stride, int flags, cudaStream_t stream);

extern "C" int hl_conv2d_uses_tensor_cores(int x_dtype, int B, int H, int W, int Cin, int Cout, int ksize,
                                           int stride, int ldx, int ldy, int flags) {
    return hl_conv_tc_applicable(x_dtype, B, H, W, Cin, Cout, ksize, stride, ldx, ldy, flags) ? 1 : 0;
}

extern "C" int hl_conv2d(const void *x, int x_dtype, int ldx, const void *wpk, const float *bias,
                         const float *residual, int ldr, void *y, int ldy, double *stats, int stats_ld,
                         int B, int H, int W, int Cin, int Cout, int ksize, int stride, int flags,
                         void *stream) {
    HL_CHECK_ARG(x && wpk && y && B > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0);
    HL_CHECK_ARG(x_dtype == HL_DT_F32 || x_dtype == HL_DT_F16);
    HL_CHECK_ARG(ksize == 1 || ksize == 3);
    HL_CHECK_ARG(stride == 1 || stride == 2);
    HL_CHECK_ARG(ldx >= Cin && ldy >= Cout && (!residual || ldr >= Cout));
    HL_CHECK_ARG(!((flags & HL_CONV_UPSAMPLE2X) && stride != 1));
    HL_CHECK_ARG(!stats || stats_ld >= Cout);
    HL_CHECK_ARG(!((flags & HL_CONV_OUT_F16_SPLIT) && stats));    // no statistics of a scaled hi | lo pair
    if (hl_conv_tc_applicable(x_dtype, B, H, W, Cin, Cout, ksize, stride, ldx, ldy, flags))
        return hl_conv2d_tc(x, x_dtype, ldx, wpk, bias, residual, ldr, y,
                            (flags & HL_CONV_OUT_F16_SPLIT) ? 2 : (flags & HL_CONV_OUT_F16) ? 1 : 0, ldy, stats,
                            stats_ld, B, H, W, Cin, Cout, ksize, stride, flags, (cudaStream_t)stream);
    HL_CHECK_ARG(!(flags & (HL_CONV_SPLIT3 | HL_CONV_SPLIT2P | HL_CONV_SPLIT2A)) || x_dtype == HL_DT_F16);
    HL_CHECK_ARG(!(flags & (HL_CONV_SPLIT3 | HL_CONV_SPLIT2A)) || ldx >= 2 * Cin);
    HL_CHECK_ARG(!(flags & HL_CONV_OUT_F16_SPLIT) || ldy >= 2 * Cout);
    int rc = hl_conv2d_simt(x, x_dtype, ldx, wpk, bias, residual, ldr, y, ldy, B, H, W, Cin, Cout, ksize, stride,
                            flags, (cudaStream_t)stream);
    if (rc != HL_OK || !stats) return rc;
    const int ups = (flags & HL_CONV_UPSAMPLE2X) ? 2 : 1;
    const int pad = ksize / 2;
    const int Ho = (H * ups + 2 * pad - ksize) / stride + 1, Wo = (W * ups + 2 * pad - ksize) / stride + 1;
    return hl_gn_stats_launch(y, (flags & HL_CONV_OUT_F16) ? 1 : 0, ldy, B, Ho * Wo, Cout, stats, stats_ld, (cudaStream_t)stream);
}

// y = conv(x) + bias + residual AND y2 = conv(x) + bias from ONE pass over the operands (see the header): the ControlNet
// projection, whose result both feeds the next ControlNet block (y2 = h_cond, unet.py:600) and, added to the main
// encoder's skip tensor, fills the decoder's concat slice (y = hs + hs_cond, unet.py:606).  Shapes the tcgen05 kernel
// does not serve (and the fp32 plan) run as the two launches this call replaces.
extern "C" int hl_conv2d_dual(const void *x, int x_dtype, int ldx, const void *wpk, const float *bias,
                              const float *residual, int ldr, float *y, int ldy, double *stats, int stats_ld,
                              float *y2, int ldy2, double *stats2, int stats2_ld, int B, int H, int W, int Cin,
                              int Cout, int ksize, int stride, int flags, void *stream) {
    HL_CHECK_ARG(x && wpk && y && y2 && residual && B > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0);
    HL_CHECK_ARG(x_dtype == HL_DT_F32 || x_dtype == HL_DT_F16);
    HL_CHECK_ARG(ksize == 1 || ksize == 3);
    HL_CHECK_ARG(stride == 1 || stride == 2);
    HL_CHECK_ARG(ldx >= Cin && ldy >= Cout && ldy2 >= Cout && ldr >= Cout);
    HL_CHECK_ARG(!(flags & (HL_CONV_OUT_F16 | HL_CONV_OUT_F16_SPLIT | HL_CONV_UPSAMPLE2X)));
    HL_CHECK_ARG((stats == nullptr) == (stats2 == nullptr));
    HL_CHECK_ARG(!stats || (stats_ld >= Cout && stats2_ld >= Cout));
    if (hl_conv_tc_applicable(x_dtype, B, H, W, Cin, Cout, ksize, stride, ldx, ldy, flags) && ldy2 % 4 == 0 &&
        ((uintptr_t)y2 & 15) == 0) {
        // the second staging ring shares the budget of the first: check that the plan still fits with both
        Plan pl = {};
        const int npass = (flags & HL_CONV_SPLIT3) ? 3 : (flags & (HL_CONV_SPLIT2P | HL_CONV_SPLIT2A)) ? 2 : 1;
        if (make_plan(x_dtype == HL_DT_F16 ? 1 : 0, B, H / stride, W / stride, Cin, Cout, ksize, stride, true, stats != nullptr,
                      &pl, npass, 2 * STATS_MAX_C * (int)sizeof(double)))
            return hl_conv2d_tc(x, x_dtype, ldx, wpk, bias, residual, ldr, y, 0, ldy, stats, stats_ld, B, H, W, Cin, Cout,
                                ksize, stride, flags, (cudaStream_t)stream, y2, ldy2, stats2, stats2_ld);
    }
    int rc = hl_conv2d(x, x_dtype, ldx, wpk, bias, nullptr, 0, y2, ldy2, stats2, stats2_ld, B, H, W, Cin, Cout, ksize,
                       stride, flags, stream);
    if (rc != HL_OK) return rc;
    return hl_conv2d(x, x_dtype, ldx, wpk, bias, residual, ldr, y, ldy, stats, stats_ld, B, H, W, Cin, Cout, ksize, stride,
                     flags, stream);
}
