// QKVAttention (human_diffusion/improved_diffusion/unet.py:255-274) on the 5th-generation tensor cores.
//
//   W = softmax_fp32((q s) . (k s)^T),  s = ch^(-1/4);   a = W . v          per (sample, head), T = H*W tokens
//
// One CTA = 128 query rows of one (sample, head); thread i owns query row i = TMEM lane i (flash-style online
// softmax, no T x T matrix ever exists):
//   * Q tile and the K / V tiles of 128 (or 64) keys arrive by TMA (cp.async.bulk.tensor, SWIZZLE_128B boxes of
//     64 channels x rows) straight from the head-major fp16 qkv tensor the qkv conv wrote ([B*T][3C]);
//   * S = Q . K^T : tcgen05.mma kind::f16, both operands from shared memory (K tile = K-major B operand), fp32
//     accumulator in tensor memory (128 columns);
//   * the thread reads its row of S (tcgen05.ld), takes the running max / sum in the exp2 domain (the 1/sqrt(ch)
//     scale is applied to the fp32 logits), rounds the weights to fp16 and writes them back to tensor memory
//     (tcgen05.st) where they ARE the A operand of the second product;
//   * O += P . V : tcgen05.mma with A from tensor memory and the V tile as an MN-major B operand (the same TMA
//     box layout as K: rows = keys, 64 channels contiguous), fp32 accumulator in tensor memory (ch columns);
//     when the running max moves, the thread rescales its O row in place (tcgen05.ld / st);
//   * K of tile j+1 is fetched as soon as S_j has been multiplied, V of tile j+1 as soon as O has consumed V_j.
// Served shapes: fp16 qkv and output, T % 64 == 0, head width in {64, 96, 128, 192}; everything else stays on the
// mma.sync kernel of attention.cu (same operand rounding: fp16 q, k, v and weights, fp32 everything else).
#include "common.cuh"
#include "tc5.cuh"

#include <cuda_fp16.h>

int hl_num_sms();

namespace {

constexpr int QT = 128;                       // query rows per CTA
constexpr uint32_t TM_S = 0, TM_P = 0 /* P overwrites the S columns this thread has already read */, TM_O = 128;

struct Attn5Params {
    int T, ch, heads, ldo, kt, nat;           // kt = keys per tile (64 | 128), nat = 64-channel atoms per operand row
    float scale_log2;                         // log2(e) / sqrt(ch)
    __half *out;
};

// K-major / MN-major SWIZZLE_128B shared-memory descriptors over a TMA box [rows][64 halves] (128 B rows, 8-row groups
// 1024 B apart).  K-major (Q, K): LBO unused.  MN-major (V: N = channels contiguous, K = key rows): LBO = distance
// between 64-channel blocks, SBO = distance between 8-key groups.
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t saddr) {
    const uint32_t hi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
    return ((uint64_t)hi << 32) | (((saddr & 0x3FFFFu) >> 4) | (1u << 16));
}
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t saddr, uint32_t lbo_bytes) {
    const uint32_t hi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
    return ((uint64_t)hi << 32) | (((saddr & 0x3FFFFu) >> 4) | (((lbo_bytes >> 4) & 0x3FFFu) << 16));
}
__device__ __forceinline__ void umma_ss_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}

__global__ void __launch_bounds__(QT, 1)
k_attention_tc5(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, const Attn5Params p) {
    extern __shared__ uint8_t smraw[];
    __shared__ __align__(8) uint64_t bars[4];          // q/k full, v full, S done, O done
    __shared__ uint32_t tmem_slot;
    const uint32_t base = (smem_u32(smraw) + 1023u) & ~1023u;
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t atom_q = QT * 128u, atom_kv = (uint32_t)p.kt * 128u;
    const uint32_t sQ = base, sK = sQ + (uint32_t)p.nat * atom_q, sV = sK + (uint32_t)p.nat * atom_kv;
    const uint32_t bar_k = smem_u32(&bars[0]), bar_v = smem_u32(&bars[1]), bar_s = smem_u32(&bars[2]), bar_o = smem_u32(&bars[3]);
    hl_pdl_trigger_early();
    if (warp == 0) {
        if (tid == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQ) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmKV) : "memory");
            for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&bars[i]), 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    hl_pdl_wait();
    tc_fence_after();
    const uint32_t tmem = tmem_slot, tlane = tmem + ((uint32_t)(warp * 32) << 16);

    const int bh = blockIdx.y, b = bh / p.heads, h = bh - b * p.heads;
    const int q0 = blockIdx.x * QT;                                  // first query token of this CTA
    const int row0 = b * p.T;                                        // first token row of the sample
    const int cq = h * 3 * p.ch, ck = cq + p.ch, cv = cq + 2 * p.ch;  // head-major channel order [q | k | v] per head
    const int n_tiles = p.T / p.kt;
    const uint32_t kv_bytes = (uint32_t)p.nat * atom_kv;

    if (tid == 0) {
        mbar_expect_tx(bar_k, (uint32_t)p.nat * atom_q + kv_bytes);
        for (int a = 0; a < p.nat; ++a) tma_load_2d(sQ + (uint32_t)a * atom_q, &tmQ, bar_k, cq + 64 * a, row0 + q0);
        for (int a = 0; a < p.nat; ++a) tma_load_2d(sK + (uint32_t)a * atom_kv, &tmKV, bar_k, ck + 64 * a, row0);
        mbar_expect_tx(bar_v, kv_bytes);
        for (int a = 0; a < p.nat; ++a) tma_load_2d(sV + (uint32_t)a * atom_kv, &tmKV, bar_v, cv + 64 * a, row0);
    }
    // instruction descriptors: D = f32, A = B = f16; S: N = kt, both K-major; O: N = ch, B MN-major (bit 16)
    const uint32_t id_s = (1u << 4) | ((uint32_t)(p.kt >> 3) << 17) | ((uint32_t)(QT >> 4) << 24);
    const uint32_t id_o = (1u << 4) | (1u << 16) | ((uint32_t)(p.ch >> 3) << 17) | ((uint32_t)(QT >> 4) << 24);
    const int ks_qk = p.ch / 16, ks_pv = p.kt / 16;

    float m_run = -INFINITY, l_run = 0.f;
    for (int j = 0; j < n_tiles; ++j) {
        const uint32_t ph = (uint32_t)j & 1u;
        // ---- S = Q . K_j^T ----
        if (warp == 0) {
            mbar_wait(bar_k, ph);
            tc_fence_after();
            if (elect_one_sync()) {
                for (int k = 0; k < ks_qk; ++k) {
                    const uint64_t ad = desc_kmajor(sQ + (uint32_t)(k >> 2) * atom_q) + (uint64_t)(2 * (k & 3));
                    const uint64_t bd = desc_kmajor(sK + (uint32_t)(k >> 2) * atom_kv) + (uint64_t)(2 * (k & 3));
                    umma_ss_f16(tmem + TM_S, ad, bd, id_s, k > 0 ? 1u : 0u);
                }
                umma_commit(bar_s);
            }
            __syncwarp();
        }
        mbar_wait(bar_s, ph);
        tc_fence_after();
        if (tid == 0 && j + 1 < n_tiles) {                           // K_j consumed: fetch K_{j+1}
            mbar_expect_tx(bar_k, kv_bytes);
            for (int a = 0; a < p.nat; ++a)
                tma_load_2d(sK + (uint32_t)a * atom_kv, &tmKV, bar_k, ck + 64 * a, row0 + (j + 1) * p.kt);
        }
        // ---- online softmax on this thread's row (exp2 domain) ----
        float mx = -INFINITY;
        for (int c = 0; c < p.kt; c += 32) {
            float v[32];
            tmem_ld32(tlane + TM_S + (uint32_t)c, v);
#pragma unroll
            for (int i = 0; i < 32; ++i) mx = fmaxf(mx, v[i]);
        }
        const float m_new = fmaxf(m_run, mx * p.scale_log2);
        const float alpha = exp2f(m_run - m_new);                     // 0 on the first tile (m_run = -inf)
        float rs = 0.f;
        for (int c = 0; c < p.kt; c += 32) {
            float v[32];
            tmem_ld32(tlane + TM_S + (uint32_t)c, v);
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const __half2 hp = __floats2half2_rn(hl_ex2(fmaf(v[2 * i], p.scale_log2, -m_new)),
                                                     hl_ex2(fmaf(v[2 * i + 1], p.scale_log2, -m_new)));
                const float2 f = __half22float2(hp);                  // the sum runs over the ROUNDED weights that get multiplied
                rs += f.x + f.y;
                pk[i] = *reinterpret_cast<const uint32_t *>(&hp);
            }
            tmem_st16(tlane + TM_P + (uint32_t)(c / 2), pk);          // columns [c/2, c/2 + 16): all already read by this thread
        }
        l_run = l_run * alpha + rs;
        m_run = m_new;
        if (j > 0) {                                                  // rescale the running output row
            for (int c = 0; c < p.ch; c += 32) {
                float v[32];
                tmem_ld32(tlane + TM_O + (uint32_t)c, v);
                uint32_t r[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(v[i] * alpha);
                tmem_st16(tlane + TM_O + (uint32_t)c, r);
                tmem_st16(tlane + TM_O + (uint32_t)c + 16u, r + 16);
            }
        }
        tmem_st_wait();
        tc_fence_before();
        __syncthreads();
        // ---- O += P . V_j ----
        if (warp == 0) {
            mbar_wait(bar_v, ph);
            tc_fence_after();
            if (elect_one_sync()) {
                for (int k = 0; k < ks_pv; ++k) {
                    const uint64_t bd = desc_mnmajor(sV + (uint32_t)k * 2048u, atom_kv);     // 16 keys = two 8-row groups
                    umma_ts_f16(tmem + TM_O, tmem + TM_P + 8u * (uint32_t)k, bd, id_o, (j > 0 || k > 0) ? 1u : 0u);
                }
                umma_commit(bar_o);
            }
            __syncwarp();
        }
        mbar_wait(bar_o, ph);
        tc_fence_after();
        if (tid == 0 && j + 1 < n_tiles) {                           // V_j consumed: fetch V_{j+1}
            mbar_expect_tx(bar_v, kv_bytes);
            for (int a = 0; a < p.nat; ++a)
                tma_load_2d(sV + (uint32_t)a * atom_kv, &tmKV, bar_v, cv + 64 * a, row0 + (j + 1) * p.kt);
        }
    }
    // ---- a = O / l -> fp16 -> out[token][h * ch ..] ----
    const int tok = q0 + tid;
    const float inv = 1.0f / l_run;
    if (tok < p.T) {
        __half *dst = p.out + (size_t)(row0 + tok) * p.ldo + h * p.ch;
        for (int c = 0; c < p.ch; c += 32) {
            float v[32];
            tmem_ld32(tlane + TM_O + (uint32_t)c, v);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                uint4 w;
                __half2 h0 = __floats2half2_rn(v[8 * i] * inv, v[8 * i + 1] * inv), h1 = __floats2half2_rn(v[8 * i + 2] * inv, v[8 * i + 3] * inv);
                __half2 h2 = __floats2half2_rn(v[8 * i + 4] * inv, v[8 * i + 5] * inv), h3 = __floats2half2_rn(v[8 * i + 6] * inv, v[8 * i + 7] * inv);
                w.x = *reinterpret_cast<uint32_t *>(&h0); w.y = *reinterpret_cast<uint32_t *>(&h1);
                w.z = *reinterpret_cast<uint32_t *>(&h2); w.w = *reinterpret_cast<uint32_t *>(&h3);
                *reinterpret_cast<uint4 *>(dst + c + 8 * i) = w;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

}  // namespace

// true if the tcgen05 kernel serves this shape (fp16 qkv / output)
bool hl_attention_tc5_applicable(const void *qkv, int ldq, const void *out, int ldo, int T, int ch) {
    if (T % 64 || (ch != 64 && ch != 96 && ch != 128 && ch != 192)) return false;
    if (((uintptr_t)qkv & 15) || ((uintptr_t)out & 15) || ldq % 8 || ldo % 8) return false;
    return hl_get_encode_tiled() != nullptr;
}

int hl_attention_tc5(const void *qkv, int ldq, void *out, int ldo, int B, int T, int C, int heads, cudaStream_t stream) {
    PFN_hl_encodeTiled encode = hl_get_encode_tiled();
    if (!encode) {
        hl_set_error("cuTensorMapEncodeTiled unavailable");
        return HL_E_CUDA;
    }
    const int ch = C / heads;
    Attn5Params p;
    p.T = T; p.ch = ch; p.heads = heads; p.ldo = ldo;
    p.kt = T % 128 == 0 ? 128 : 64;
    p.nat = (ch + 63) / 64;
    p.scale_log2 = 1.4426950408889634f / sqrtf((float)ch);
    p.out = (__half *)out;
    CUtensorMap tmQ, tmKV;
    for (int which = 0; which < 2; ++which) {
        // qkv as a 2-D tensor [B*T rows][3C channels]; a box = 64 channels x (128 query | kt key) rows.  A 96-wide head
        // reads a second 64-channel box whose upper half belongs to the next operand -- loaded, never multiplied.
        cuuint64_t gdim[2] = {(cuuint64_t)(3 * C), (cuuint64_t)B * T};
        cuuint64_t gstr[1] = {(cuuint64_t)ldq * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)(which ? p.kt : QT)};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(which ? &tmKV : &tmQ, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void *)qkv, gdim, gstr, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            hl_set_error("cuTensorMapEncodeTiled(qkv) failed: %d (B=%d T=%d C=%d ldq=%d)", (int)r, B, T, C, ldq);
            return HL_E_CUDA;
        }
    }
    const size_t smem = 1024 + (size_t)p.nat * (QT * 128 + 2 * p.kt * 128);
    static bool configured[64] = {};
    int dev = 0;
    HL_CHECK_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        HL_CHECK_CUDA(cudaFuncSetAttribute(k_attention_tc5, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        configured[dev] = true;
    }
    dim3 grid((T + QT - 1) / QT, B * heads);
    HL_CHECK_CUDA(hl_launch(k_attention_tc5, grid, dim3(QT), smem, stream, tmQ, tmKV, p));
    HL_CHECK_LAUNCH();
    return HL_OK;
}
