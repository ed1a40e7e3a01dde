// QKVAttention (unet.py:255-274) as one fused flash-style kernel: S = (q.k)/sqrt(ch) per 64x64 tile,
// online softmax in fp32, O += P.V, never materialising the T x T weight matrix the reference builds.
// fp32 CUDA-core version (exact-precision companion; attention is 0.9 % of the step's FLOPs).
//
// qkv rows are [B, T, 3C] (NHWC pixels as tokens) with the reference's head-major channel order
// [head][q(ch) | k(ch) | v(ch)]  (qkv.reshape(b*heads, 3*ch, T), unet.py:248,267-268).
#include "common.cuh"

#include <cuda_fp16.h>

namespace {

constexpr int TQ = 64, TK = 64;

template <int CH>
__global__ void __launch_bounds__(256) k_attention(const float *__restrict__ qkv, int ldq,
                                                   void *__restrict__ out, int out_dtype, int ldo, int T,
                                                   int heads, float scale, int round_tf32) {
    hl_pdl_enter();
    constexpr int LD = CH + 4;           // padded row pitch (floats): conflict-free float4 rows
    constexpr int NJ = CH / 16;          // output columns per thread
    extern __shared__ float sm[];
    float *Qs = sm;                      // [TQ][LD]
    float *Ks = Qs + TQ * LD;            // [TK][LD]
    float *Vs = Ks + TK * LD;            // [TK][LD]
    float *Ss = Vs + TK * LD;            // [TQ][TK+1]
    float *row_m = Ss + TQ * (TK + 1);   // [TQ] running max
    float *row_l = row_m + TQ;           // [TQ] running sum
    float *row_a = row_l + TQ;           // [TQ] rescale factor of this tile

    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * TQ;
    const float *base = qkv + (int64_t)b * T * ldq + h * 3 * CH;

    for (int i = tid; i < TQ * (CH / 4); i += 256) {
        int r = i / (CH / 4), c4 = i % (CH / 4);
        float4 v = make_float4(0, 0, 0, 0);
        if (q0 + r < T) v = *reinterpret_cast<const float4 *>(base + (int64_t)(q0 + r) * ldq + 4 * c4);
        *reinterpret_cast<float4 *>(&Qs[r * LD + 4 * c4]) = v;
    }
    if (tid < TQ) { row_m[tid] = -INFINITY; row_l[tid] = 0.f; }

    float o[4][NJ];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) o[i][j] = 0.f;

    for (int k0 = 0; k0 < T; k0 += TK) {
        __syncthreads();
        for (int i = tid; i < TK * (CH / 4); i += 256) {
            int r = i / (CH / 4), c4 = i % (CH / 4);
            float4 kv = make_float4(0, 0, 0, 0), vv = make_float4(0, 0, 0, 0);
            if (k0 + r < T) {
                const float *rowp = base + (int64_t)(k0 + r) * ldq + 4 * c4;
                kv = *reinterpret_cast<const float4 *>(rowp + CH);
                vv = *reinterpret_cast<const float4 *>(rowp + 2 * CH);
            }
            *reinterpret_cast<float4 *>(&Ks[r * LD + 4 * c4]) = kv;
            *reinterpret_cast<float4 *>(&Vs[r * LD + 4 * c4]) = vv;
        }
        __syncthreads();

        // S tile: rows ty+16i, cols tx+16j
        float s[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
        for (int c = 0; c < CH; c += 4) {
            float4 qa[4], kb[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) qa[i] = *reinterpret_cast<const float4 *>(&Qs[(ty + 16 * i) * LD + c]);
#pragma unroll
            for (int j = 0; j < 4; ++j) kb[j] = *reinterpret_cast<const float4 *>(&Ks[(tx + 16 * j) * LD + c]);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    s[i][j] = fmaf(qa[i].x, kb[j].x, s[i][j]);
                    s[i][j] = fmaf(qa[i].y, kb[j].y, s[i][j]);
                    s[i][j] = fmaf(qa[i].z, kb[j].z, s[i][j]);
                    s[i][j] = fmaf(qa[i].w, kb[j].w, s[i][j]);
                }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int col = tx + 16 * j;
                Ss[(ty + 16 * i) * (TK + 1) + col] = (k0 + col < T) ? s[i][j] * scale : -INFINITY;
            }
        __syncthreads();

        // online softmax: 4 threads per row
        {
            int r = tid >> 2, part = tid & 3;
            float *srow = &Ss[r * (TK + 1)];
            float mx = -INFINITY;
            for (int c = part; c < TK; c += 4) mx = fmaxf(mx, srow[c]);
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
            float m_old = row_m[r];
            float m_new = fmaxf(m_old, mx);
            float sum = 0.f;
            for (int c = part; c < TK; c += 4) {
                float pv = expf(srow[c] - m_new);   // exp(-inf) = 0 for masked keys
                srow[c] = pv;
                sum += pv;
            }
            sum += __shfl_xor_sync(0xffffffffu, sum, 1);
            sum += __shfl_xor_sync(0xffffffffu, sum, 2);
            __syncwarp();
            if (part == 0) {
                float a = (m_old == -INFINITY) ? 0.f : expf(m_old - m_new);
                row_a[r] = a;
                row_l[r] = row_l[r] * a + sum;
                row_m[r] = m_new;
            }
        }
        __syncthreads();

        // O = O*alpha + P.V
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float a = row_a[ty + 16 * i];
#pragma unroll
            for (int j = 0; j < NJ; ++j) o[i][j] *= a;
        }
        for (int kk = 0; kk < TK; ++kk) {
            float pv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) pv[i] = Ss[(ty + 16 * i) * (TK + 1) + kk];
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                float vv = Vs[kk * LD + tx + 16 * j];
#pragma unroll
                for (int i = 0; i < 4; ++i) o[i][j] = fmaf(pv[i], vv, o[i][j]);
            }
        }
    }

#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int r = ty + 16 * i;
        if (q0 + r >= T) continue;
        float inv = 1.0f / row_l[r];
        const int64_t obase = ((int64_t)b * T + q0 + r) * ldo + h * CH;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            float v = o[i][j] * inv;
            if (out_dtype == HL_DT_F16) reinterpret_cast<__half *>(out)[obase + tx + 16 * j] = __float2half_rn(v);
            else reinterpret_cast<float *>(out)[obase + tx + 16 * j] = round_tf32 ? hl_rna_tf32(v) : v;
        }
    }
}

template <int CH>
int launch(const float *qkv, int ldq, void *out, int out_dtype, int ldo, int B, int T, int heads,
           int round_tf32, cudaStream_t stream) {
    constexpr int LD = CH + 4;
    size_t smem = sizeof(float) * (size_t)(3 * 64 * LD + 64 * 65 + 3 * 64);
    static HlPerDeviceOnce once;
    if (once.need()) {
        HL_CHECK_CUDA(cudaFuncSetAttribute(k_attention<CH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)smem));
    }
    dim3 grid(hl_cdiv(T, TQ), heads, B);
    float scale = 1.0f / sqrtf((float)CH);
    HL_CHECK_CUDA(hl_launch(k_attention<CH>, grid, dim3(256), smem, stream, qkv, ldq, out, out_dtype, ldo, T, heads, scale, round_tf32));
    HL_CHECK_LAUNCH();
    return HL_OK;
}

// ---------------------------------------------------------------------------------------------
// Tensor-core version (operand mode fp16): the same flash-style algorithm with both contractions on
// mma.sync m16n8k16 (fp16 operands, fp32 accumulate) -- Q, K, V and the softmax weights P are rounded
// to fp16 exactly like every other GEMM operand of the denoise path; scores, running max / sum and the
// output accumulator stay fp32.  One CTA = 64 query rows of one (sample, head), 4 warps x 16 rows;
// K / V tiles of 64 keys are converted to fp16 on their way into shared memory (row pitch padded by
// 16 B so ldmatrix is conflict-free).  T x T scores never leave registers.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
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

// F16IN: qkv was written as fp16 by the producing conv (HL_CONV_OUT_F16); K / V tiles then stream into a
// double-buffered ring with cp.async (16 B, zero-filled beyond T) while the previous tile is being multiplied.
template <int CH, bool F16IN>
__global__ void __launch_bounds__(128) k_attention_mma(const void *__restrict__ qkv_, int ldq,
                                                       __half *__restrict__ out, int ldo, int T, float scale_log2e) {
    hl_pdl_enter();
    constexpr int PITCH = CH * 2 + 16;          // bytes per smem row
    constexpr int KS = CH / 16;                 // k steps of Q.K^T
    constexpr int NO = CH / 8;                  // n tiles of the output
    constexpr int KV_BYTES = 2 * 64 * PITCH;    // one K tile + one V tile
    extern __shared__ __align__(16) uint8_t smraw[];
    uint8_t *Qs = smraw, *Ks = Qs + 64 * PITCH, *Vs = Ks + 64 * PITCH;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * 64;
    const float *base = reinterpret_cast<const float *>(qkv_) + (int64_t)b * T * ldq + h * 3 * CH;
    const __half *base_h = reinterpret_cast<const __half *>(qkv_) + (int64_t)b * T * ldq + h * 3 * CH;

    auto cp_tile = [&](uint8_t *dst, int row0, int col0) {     // 64 rows x CH fp16, asynchronous
        constexpr int N8 = 64 * (CH / 8);
        for (int i = tid; i < N8; i += 128) {
            const int r = i / (CH / 8), c8 = i % (CH / 8);
            const bool ok = row0 + r < T;
            const __half *src = base_h + (int64_t)(ok ? row0 + r : T - 1) * ldq + col0 + 8 * c8;
            const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst + r * PITCH + c8 * 16);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(ok ? 16 : 0) : "memory");
        }
    };

    auto load_tile = [&](uint8_t *dst, int row0, int col0) {   // 64 rows x CH fp32 -> fp16 smem
        constexpr int N4 = 64 * (CH / 4), U = 8;               // U independent 128-bit loads in flight per thread
        for (int i0 = tid; i0 < N4; i0 += 128 * U) {
            float4 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = i0 + u * 128;
                const int r = i / (CH / 4), c4 = i % (CH / 4);
                v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (i < N4 && row0 + r < T)
                    v[u] = *reinterpret_cast<const float4 *>(base + (int64_t)(row0 + r) * ldq + col0 + 4 * c4);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = i0 + u * 128;
                if (i >= N4) break;
                const int r = i / (CH / 4), c4 = i % (CH / 4);
                uint2 w;
                w.x = pack_h2(v[u].x, v[u].y);
                w.y = pack_h2(v[u].z, v[u].w);
                *reinterpret_cast<uint2 *>(dst + r * PITCH + c4 * 8) = w;
            }
        }
    };
    if constexpr (F16IN) {
        cp_tile(Qs, q0, 0);
        cp_tile(Ks, 0, CH);
        cp_tile(Vs, 0, 2 * CH);
        asm volatile("cp.async.commit_group;" ::: "memory");
    } else {
        load_tile(Qs, q0, 0);
    }

    float o[NO][4];
#pragma unroll
    for (int j = 0; j < NO; ++j) { o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f; }
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;    // rows g and g+8 of this warp's 16
    const uint32_t q_addr = (uint32_t)__cvta_generic_to_shared(Qs) +
                            (uint32_t)((warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * PITCH + (lane >> 4) * 16);
    const uint32_t k_addr0 = (uint32_t)__cvta_generic_to_shared(Ks) +
                             (uint32_t)(((lane >> 4) * 8 + (lane & 7)) * PITCH + ((lane >> 3) & 1) * 16);
    const uint32_t v_addr0 = (uint32_t)__cvta_generic_to_shared(Vs) +
                             (uint32_t)((((lane >> 3) & 1) * 8 + (lane & 7)) * PITCH + (lane >> 4) * 16);

    for (int k0 = 0, it = 0; k0 < T; k0 += 64, ++it) {
        uint32_t k_addr = k_addr0, v_addr = v_addr0;
        if constexpr (F16IN) {
            __syncthreads();                     // the other buffer (tile it-1) fully consumed
            if (k0 + 64 < T) {
                uint8_t *nk = Ks + ((it + 1) & 1) * KV_BYTES;
                cp_tile(nk, k0 + 64, CH);
                cp_tile(nk + 64 * PITCH, k0 + 64, 2 * CH);
                asm volatile("cp.async.commit_group;" ::: "memory");
                asm volatile("cp.async.wait_group 1;" ::: "memory");
            } else {
                asm volatile("cp.async.wait_group 0;" ::: "memory");
            }
            __syncthreads();
            k_addr += (it & 1) * KV_BYTES;
            v_addr += (it & 1) * KV_BYTES;
        } else {
            __syncthreads();                     // previous tile fully consumed (and Q visible)
            load_tile(Ks, k0, CH);
            load_tile(Vs, k0, 2 * CH);
            __syncthreads();
        }

        float s[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) { s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f; }
#pragma unroll
        for (int kk = 0; kk < KS; ++kk) {
            uint32_t a[4];
            ldsm_x4(q_addr + kk * 32, a[0], a[1], a[2], a[3]);
#pragma unroll
            for (int jp = 0; jp < 4; ++jp) {
                uint32_t b0, b1, b2, b3;
                ldsm_x4(k_addr + jp * 16 * PITCH + kk * 32, b0, b1, b2, b3);
                mma16816(s[2 * jp], a, b0, b1);
                mma16816(s[2 * jp + 1], a, b2, b3);
            }
        }
        // scale (log2 domain), mask keys beyond T, online softmax
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int key = k0 + j * 8 + 2 * t;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const bool ok = key + (e & 1) < T;
                s[j][e] = ok ? s[j][e] * scale_log2e : -INFINITY;
            }
            mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
            mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);   // finite: every tile holds >= 1 valid key
        const float al0 = hl_ex2(m0 - mn0), al1 = hl_ex2(m1 - mn1);
        float sum0 = 0.f, sum1 = 0.f;
        uint32_t pa[4][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float p0 = hl_ex2(s[j][0] - mn0), p1 = hl_ex2(s[j][1] - mn0);
            const float p2 = hl_ex2(s[j][2] - mn1), p3 = hl_ex2(s[j][3] - mn1);
            sum0 += p0 + p1;
            sum1 += p2 + p3;
            pa[j >> 1][(j & 1) * 2 + 0] = pack_h2(p0, p1);
            pa[j >> 1][(j & 1) * 2 + 1] = pack_h2(p2, p3);
        }
        sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
        sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
        sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
        sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
        l0 = l0 * al0 + sum0;
        l1 = l1 * al1 + sum1;
        m0 = mn0;
        m1 = mn1;
#pragma unroll
        for (int j = 0; j < NO; ++j) { o[j][0] *= al0; o[j][1] *= al0; o[j][2] *= al1; o[j][3] *= al1; }
        // O += P . V
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
            for (int np = 0; np < NO / 2; ++np) {
                uint32_t b0, b1, b2, b3;
                ldsm_x4_t(v_addr + kk * 16 * PITCH + np * 32, b0, b1, b2, b3);
                mma16816(o[2 * np], pa[kk], b0, b1);
                mma16816(o[2 * np + 1], pa[kk], b2, b3);
            }
        }
    }
    const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
    const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
#pragma unroll
    for (int j = 0; j < NO; ++j) {
        const int col = h * CH + j * 8 + 2 * t;
        if (r0 < T) *reinterpret_cast<uint32_t *>(out + ((int64_t)b * T + r0) * ldo + col) = pack_h2(o[j][0] * inv0, o[j][1] * inv0);
        if (r1 < T) *reinterpret_cast<uint32_t *>(out + ((int64_t)b * T + r1) * ldo + col) = pack_h2(o[j][2] * inv1, o[j][3] * inv1);
    }
}

template <int CH, bool F16IN>
int launch_mma_t(const void *qkv, int ldq, __half *out, int ldo, int B, int T, int heads, cudaStream_t stream) {
    constexpr int PITCH = CH * 2 + 16;
    size_t smem = (size_t)(F16IN ? 5 : 3) * 64 * PITCH;
    static HlPerDeviceOnce once;
    if (once.need()) {
        HL_CHECK_CUDA(cudaFuncSetAttribute(k_attention_mma<CH, F16IN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    dim3 grid(hl_cdiv(T, 64), heads, B);
    const float scale_log2e = 1.4426950408889634f / sqrtf((float)CH);
    HL_CHECK_CUDA(hl_launch(k_attention_mma<CH, F16IN>, grid, dim3(128), smem, stream, qkv, ldq, out, ldo, T, scale_log2e));
    HL_CHECK_LAUNCH();
    return HL_OK;
}
template <int CH>
int launch_mma(const void *qkv, int qkv_f16, int ldq, __half *out, int ldo, int B, int T, int heads, cudaStream_t stream) {
    return qkv_f16 ? launch_mma_t<CH, true>(qkv, ldq, out, ldo, B, T, heads, stream)
                   : launch_mma_t<CH, false>(qkv, ldq, out, ldo, B, T, heads, stream);
}

}  // namespace

bool hl_attention_tc5_applicable(const void *qkv, int ldq, const void *out, int ldo, int T, int ch);
int hl_attention_tc5(const void *qkv, int ldq, void *out, int ldo, int B, int T, int C, int heads, cudaStream_t stream);

extern "C" int hl_attention(const void *qkv, int qkv_dtype, int ldq, void *out, int out_dtype, int ldo, int B, int T,
                            int C, int heads, int round_tf32, void *stream) {
    HL_CHECK_ARG(qkv && out && B > 0 && T > 0 && C > 0 && heads > 0 && C % heads == 0);
    const int qf16 = qkv_dtype == HL_DT_F16;
    HL_CHECK_ARG(ldq >= 3 * C && ldo >= C && ldq % (qf16 ? 8 : 4) == 0);
    int ch = C / heads;
    cudaStream_t st = (cudaStream_t)stream;
    if (out_dtype == HL_DT_F16 && !(round_tf32 & 2) && ldo % 2 == 0 && ((uintptr_t)out & 3) == 0 &&
        (!qf16 || ((uintptr_t)qkv & 15) == 0)) {
        __half *oh = (__half *)out;
        // fp16 in / out: the tcgen05 kernel (attention_tc5.cu) where it tiles the shape, else the mma.sync kernel
        if (qf16 && !(round_tf32 & 4) && hl_attention_tc5_applicable(qkv, ldq, out, ldo, T, ch))
            return hl_attention_tc5(qkv, ldq, out, ldo, B, T, C, heads, st);
        switch (ch) {
            case 32: return launch_mma<32>(qkv, qf16, ldq, oh, ldo, B, T, heads, st);
            case 64: return launch_mma<64>(qkv, qf16, ldq, oh, ldo, B, T, heads, st);
            case 96: return launch_mma<96>(qkv, qf16, ldq, oh, ldo, B, T, heads, st);
            case 128: return launch_mma<128>(qkv, qf16, ldq, oh, ldo, B, T, heads, st);
            case 192: return launch_mma<192>(qkv, qf16, ldq, oh, ldo, B, T, heads, st);
            default: break;   // other head widths: CUDA-core kernel below
        }
    }
    if (qf16) {
        hl_set_error("hl_attention: fp16 qkv needs the tensor-core kernel (fp16 output, head width 32/64/96/128/192; got %d)", ch);
        return HL_E_UNSUPPORTED;
    }
    const float *qf = (const float *)qkv;
    switch (ch) {
        case 32: return launch<32>(qf, ldq, out, out_dtype, ldo, B, T, heads, round_tf32 & 1, st);
        case 64: return launch<64>(qf, ldq, out, out_dtype, ldo, B, T, heads, round_tf32 & 1, st);
        case 96: return launch<96>(qf, ldq, out, out_dtype, ldo, B, T, heads, round_tf32 & 1, st);
        case 128: return launch<128>(qf, ldq, out, out_dtype, ldo, B, T, heads, round_tf32 & 1, st);
        case 192: return launch<192>(qf, ldq, out, out_dtype, ldo, B, T, heads, round_tf32 & 1, st);
        default:
            hl_set_error("hl_attention: unsupported head width %d (supported: 32,64,96,128,192)", ch);
            return HL_E_UNSUPPORTED;
    }
}
