// tcgen05 (5th-gen tensor core) implicit-GEMM convolution for sm_100a, NHWC fp32 in / fp32 out,
// TF32 operands with fp32 accumulation in TMEM.
//
//   GEMM view   M = B*H*W output pixels (tile: 128 pixels = a [bn x bh x bw] box of the image)
//               N = Cout               (tile: UMMA_N = 32..256 output channels)
//               K = taps * Cin         (chunk: 32 channels of one filter tap = one 128-byte row)
//
//   A operand   the activation tensor itself, fetched by TMA as a 4-D box {32ch, bw, bh, bn} whose
//               (x, y) origin is shifted by the filter tap; out-of-bounds pixels are zero-filled by
//               the TMA unit, which IS the conv's zero padding -- no im2col buffer ever exists.
//   B operand   packed weights [tap][Cout_pad][Cin], 3-D TMA box {32ch, UMMA_N, 1}.
//   both land in shared memory in the SWIZZLE_128B K-major layout tcgen05.mma consumes directly.
//
//   warp roles  warp 0: TMA producer (one lane)      warp 1: TMEM alloc + MMA issue (one lane)
//               warps 2-5: epilogue (tcgen05.ld -> +bias (+residual) -> global)
//   pipeline    `stages`-deep full/empty mbarrier ring between TMA and MMA; tcgen05.commit releases
//               a stage when the MMAs that read it retire; a final commit hands TMEM to the epilogue.
//
// Algorithmic work per launch: 2*M*N*K flop; compulsory HBM bytes: 4*(M*Cin + taps*Cout*Cin + M*Cout
// (+ M*Cout residual)).
#include "common.cuh"

#include <cuda.h>

int hl_num_sms();

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 32;   // fp32/tf32 elements per K chunk (128 bytes)
constexpr int UMMA_K = 8;     // tf32: 32 bytes per MMA K step
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 4;
constexpr int MAX_STAGES = 8;
constexpr int NUM_THREADS = 192;

struct TcParams {
    int taps, ksize, kchunks_per_tap, n_tile, stages, tmem_cols;
    int bw, bh, bn, tiles_w, tiles_h;
    int B, H, W, Cout;
    const float *bias;
    const float *res;
    int ldr;
    float *y;
    int ldy;
    int vec_ok;   // 1: 16-byte epilogue accesses are legal
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}
// Bounded spin: a protocol bug becomes a trap (an error the host sees) instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t it = 0; it < (1u << 26); ++it) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) return;
    }
    __trap();
}

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0,
                                            int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0,
                                            int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 format, version 1):
// rows of 128 bytes, 8-row swizzle atoms 1024 bytes apart (SBO), LBO unused for swizzled K-major.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
                 : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float *v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
k_conv_tc(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
          const TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2 * MAX_STAGES + 1];
    __shared__ uint32_t tmem_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int b_stage_bytes = p.n_tile * BLOCK_K * 4;
    const uint32_t smem_a = smem_base;
    const uint32_t smem_b = smem_base + (uint32_t)p.stages * A_STAGE_BYTES;
    const uint32_t bar_full = smem_u32(&bars[0]);
    const uint32_t bar_empty = smem_u32(&bars[MAX_STAGES]);
    const uint32_t bar_tmem = smem_u32(&bars[2 * MAX_STAGES]);

    // tile coordinates
    const int mt = blockIdx.x;
    const int tw = mt % p.tiles_w;
    const int th = (mt / p.tiles_w) % p.tiles_h;
    const int tn = mt / (p.tiles_w * p.tiles_h);
    const int w0 = tw * p.bw, h0 = th * p.bh, n0 = tn * p.bn;
    const int nt0 = blockIdx.y * p.n_tile;
    const int total_chunks = p.taps * p.kchunks_per_tap;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < p.stages; ++s) {
                mbar_init(bar_full + 8 * s, 1);
                mbar_init(bar_empty + 8 * s, 1);
            }
            mbar_init(bar_tmem, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32(&tmem_slot)),
                     "r"((uint32_t)p.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ------------------------------ TMA producer ------------------------------
            const int pad = p.ksize / 2;
            const uint32_t tx_bytes = (uint32_t)(A_STAGE_BYTES + b_stage_bytes);
            int s = 0;
            uint32_t ph = 0;
            for (int kc = 0; kc < total_chunks; ++kc) {
                mbar_wait(bar_empty + 8 * s, ph ^ 1u);
                mbar_expect_tx(bar_full + 8 * s, tx_bytes);
                const int tap = kc / p.kchunks_per_tap;
                const int c0 = (kc - tap * p.kchunks_per_tap) * BLOCK_K;
                const int dy = tap / p.ksize - pad, dx = tap % p.ksize - pad;
                tma_load_4d(smem_a + s * A_STAGE_BYTES, &tmA, bar_full + 8 * s, c0, w0 + dx, h0 + dy, n0);
                tma_load_3d(smem_b + s * b_stage_bytes, &tmB, bar_full + 8 * s, c0, nt0, tap);
                if (++s == p.stages) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ------------------------------ MMA issuer --------------------------------
            // instruction descriptor: D=f32, A=B=tf32, both K-major, N>>3 @17, M>>4 @24
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) |
                                   ((uint32_t)(p.n_tile >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
            int s = 0;
            uint32_t ph = 0;
            for (int kc = 0; kc < total_chunks; ++kc) {
                mbar_wait(bar_full + 8 * s, ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a0 = smem_a + s * A_STAGE_BYTES;
                const uint32_t b0 = smem_b + s * b_stage_bytes;
#pragma unroll
                for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                    umma_tf32(tmem_base, umma_desc_sw128(a0 + k * UMMA_K * 4),
                              umma_desc_sw128(b0 + k * UMMA_K * 4), idesc, (kc | k) != 0);
                }
                umma_commit(bar_empty + 8 * s);   // stage reusable once these MMAs retire
                if (++s == p.stages) { s = 0; ph ^= 1u; }
            }
            umma_commit(bar_tmem);                // accumulator complete -> epilogue
        }
    } else {
        // ---------------------------------- epilogue ----------------------------------
        mbar_wait(bar_tmem, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int q = warp & 3;                   // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;            // GEMM row = pixel index inside the box
        const int pw = row % p.bw;
        const int phh = (row / p.bw) % p.bh;
        const int pn = row / (p.bw * p.bh);
        const bool valid = (n0 + pn) < p.B && (h0 + phh) < p.H && (w0 + pw) < p.W;
        const int64_t pix = ((int64_t)(n0 + pn) * p.H + (h0 + phh)) * p.W + (w0 + pw);
        float *yrow = p.y + pix * p.ldy;
        const float *rrow = p.res ? p.res + pix * p.ldr : nullptr;
        for (int c = 0; c < p.n_tile; c += 32) {
            float v[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
            const int nbase = nt0 + c;
            if (!valid || nbase >= p.Cout) {
                // nothing to store for this lane (padding row / padded channels)
            } else if (p.vec_ok && nbase + 32 <= p.Cout) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    float4 bz = *reinterpret_cast<const float4 *>(p.bias + nbase + j);
                    float4 o = make_float4(v[j] + bz.x, v[j + 1] + bz.y, v[j + 2] + bz.z, v[j + 3] + bz.w);
                    if (rrow) {
                        float4 r = *reinterpret_cast<const float4 *>(rrow + nbase + j);
                        o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
                    }
                    *reinterpret_cast<float4 *>(yrow + nbase + j) = o;
                }
            } else {
                for (int j = 0; j < 32 && nbase + j < p.Cout; ++j) {
                    float o = v[j] + p.bias[nbase + j];
                    if (rrow) o += rrow[nbase + j];
                    yrow[nbase + j] = o;
                }
            }
            __syncwarp();   // tcgen05.ld is warp-collective: reconverge before the next one
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "r"((uint32_t)p.tmem_cols)
                     : "memory");
    }
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

int pick_n_tile(int cout_pad) {
    if (cout_pad % 256 == 0) return 256;
    if (cout_pad % 192 == 0) return 192;
    if (cout_pad % 128 == 0) return 128;
    if (cout_pad % 64 == 0) return 64;
    return 32;
}

struct Tiling {
    int bw, bh, bn;
};

bool pick_tiling(int B, int H, int W, Tiling *t) {
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
    (void)B;
    return true;
}

}  // namespace

extern "C" int hl_conv_cout_pad(int Cout) { return (Cout + 31) / 32 * 32; }

bool hl_conv_tc_applicable(int B, int H, int W, int Cin, int Cout, int ksize, int stride, int ldx,
                           int flags) {
    if (flags & (HL_CONV_FORCE_SIMT | HL_CONV_UPSAMPLE2X)) return false;
    if (stride != 1 || (ksize != 1 && ksize != 3)) return false;
    if (Cin % 32 || ldx % 4 || Cin > ldx) return false;
    if ((int64_t)B * H * W < 128) return false;     // tiny maps: the fp32 kernel is as fast
    Tiling t;
    if (!pick_tiling(B, H, W, &t)) return false;
    (void)Cout;
    return get_encode() != nullptr;
}

int hl_conv2d_tc(const float *x, int ldx, const float *wpk, const float *bias, const float *residual,
                 int ldr, float *y, int ldy, int B, int H, int W, int Cin, int Cout, int ksize,
                 cudaStream_t stream) {
    PFN_encodeTiled encode = get_encode();
    if (!encode) {
        hl_set_error("cuTensorMapEncodeTiled unavailable");
        return HL_E_CUDA;
    }
    Tiling t;
    HL_CHECK_ARG(pick_tiling(B, H, W, &t));
    HL_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)wpk & 15) == 0 && bias != nullptr);
    const int cout_pad = hl_conv_cout_pad(Cout);
    TcParams p;
    p.taps = ksize * ksize;
    p.ksize = ksize;
    p.kchunks_per_tap = Cin / BLOCK_K;
    p.n_tile = pick_n_tile(cout_pad);
    const int b_stage = p.n_tile * BLOCK_K * 4;
    int stages = (227 * 1024 - 1024 - 1024) / (A_STAGE_BYTES + b_stage);
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    const int total_chunks = p.taps * p.kchunks_per_tap;
    if (stages > total_chunks) stages = total_chunks;
    if (stages < 1) stages = 1;
    p.stages = stages;
    int cols = 32;
    while (cols < p.n_tile) cols <<= 1;
    p.tmem_cols = cols;
    p.bw = t.bw; p.bh = t.bh; p.bn = t.bn;
    p.tiles_w = W / t.bw;
    p.tiles_h = H / t.bh;
    const int tiles_n = (B + t.bn - 1) / t.bn;
    p.B = B; p.H = H; p.W = W; p.Cout = Cout;
    p.bias = bias; p.res = residual; p.ldr = ldr; p.y = y; p.ldy = ldy;
    p.vec_ok = (ldy % 4 == 0) && (((uintptr_t)y & 15) == 0) && (((uintptr_t)bias & 15) == 0) &&
               (!residual || (ldr % 4 == 0 && ((uintptr_t)residual & 15) == 0));

    CUtensorMap tmA, tmB;
    {
        cuuint64_t gdim[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
        cuuint64_t gstr[3] = {(cuuint64_t)ldx * 4, (cuuint64_t)W * ldx * 4, (cuuint64_t)H * W * ldx * 4};
        cuuint32_t box[4] = {(cuuint32_t)BLOCK_K, (cuuint32_t)t.bw, (cuuint32_t)t.bh, (cuuint32_t)t.bn};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = encode(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void *)x, gdim, gstr, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            hl_set_error("cuTensorMapEncodeTiled(A) failed: %d (B=%d H=%d W=%d Cin=%d ldx=%d)", (int)r, B,
                         H, W, Cin, ldx);
            return HL_E_CUDA;
        }
    }
    {
        cuuint64_t gdim[3] = {(cuuint64_t)Cin, (cuuint64_t)cout_pad, (cuuint64_t)p.taps};
        cuuint64_t gstr[2] = {(cuuint64_t)Cin * 4, (cuuint64_t)cout_pad * Cin * 4};
        cuuint32_t box[3] = {(cuuint32_t)BLOCK_K, (cuuint32_t)p.n_tile, 1};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = encode(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void *)wpk, gdim, gstr, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            hl_set_error("cuTensorMapEncodeTiled(B) failed: %d (Cin=%d Cout_pad=%d taps=%d)", (int)r, Cin,
                         cout_pad, p.taps);
            return HL_E_CUDA;
        }
    }
    const size_t smem = (size_t)p.stages * (A_STAGE_BYTES + b_stage) + 1024;
    static bool smem_configured = false;
    if (!smem_configured) {
        // 227 KB per CTA minus the kernel's static shared memory (barriers + TMEM slot)
        HL_CHECK_CUDA(cudaFuncSetAttribute(k_conv_tc, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           227 * 1024 - 1024));
        smem_configured = true;
    }
    dim3 grid(p.tiles_w * p.tiles_h * tiles_n, cout_pad / p.n_tile);
    k_conv_tc<<<grid, NUM_THREADS, smem, stream>>>(tmA, tmB, p);
    HL_CHECK_LAUNCH();
    return HL_OK;
}

// ---------------------------------------------------------------------------------------------
// public entry: dispatch
// ---------------------------------------------------------------------------------------------
int hl_conv2d_simt(const float *x, int ldx, const float *wpk, const float *bias, const float *residual,
                   int ldr, float *y, int ldy, int B, int H, int W, int Cin, int Cout, int ksize,
                   int stride, int flags, cudaStream_t stream);

extern "C" int hl_conv2d_uses_tensor_cores(int B, int H, int W, int Cin, int Cout, int ksize,
                                           int stride, int ldx, int flags) {
    return hl_conv_tc_applicable(B, H, W, Cin, Cout, ksize, stride, ldx, flags) ? 1 : 0;
}

extern "C" int hl_conv2d(const float *x, int ldx, const float *wpk, const float *bias,
                         const float *residual, int ldr, float *y, int ldy, int B, int H, int W, int Cin,
                         int Cout, int ksize, int stride, int flags, void *stream) {
    HL_CHECK_ARG(x && wpk && y && B > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0);
    HL_CHECK_ARG(ksize == 1 || ksize == 3);
    HL_CHECK_ARG(stride == 1 || stride == 2);
    HL_CHECK_ARG(ldx >= Cin && ldy >= Cout && (!residual || ldr >= Cout));
    HL_CHECK_ARG(!((flags & HL_CONV_UPSAMPLE2X) && stride != 1));
    if (hl_conv_tc_applicable(B, H, W, Cin, Cout, ksize, stride, ldx, flags))
        return hl_conv2d_tc(x, ldx, wpk, bias, residual, ldr, y, ldy, B, H, W, Cin, Cout, ksize,
                            (cudaStream_t)stream);
    return hl_conv2d_simt(x, ldx, wpk, bias, residual, ldr, y, ldy, B, H, W, Cin, Cout, ksize, stride,
                          flags, (cudaStream_t)stream);
}
