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
    static bool configured = false;
    if (!configured) {
        HL_CHECK_CUDA(cudaFuncSetAttribute(k_attention<CH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)smem));
        configured = true;
    }
    dim3 grid(hl_cdiv(T, TQ), heads, B);
    float scale = 1.0f / sqrtf((float)CH);
    k_attention<CH><<<grid, 256, smem, stream>>>(qkv, ldq, out, out_dtype, ldo, T, heads, scale, round_tf32);
    HL_CHECK_LAUNCH();
    return HL_OK;
}

}  // namespace

extern "C" int hl_attention(const float *qkv, int ldq, void *out, int out_dtype, int ldo, int B, int T, int C,
                            int heads, int round_tf32, void *stream) {
    HL_CHECK_ARG(qkv && out && B > 0 && T > 0 && C > 0 && heads > 0 && C % heads == 0);
    HL_CHECK_ARG(ldq >= 3 * C && ldo >= C && ldq % 4 == 0);
    int ch = C / heads;
    cudaStream_t st = (cudaStream_t)stream;
    switch (ch) {
        case 32: return launch<32>(qkv, ldq, out, out_dtype, ldo, B, T, heads, round_tf32, st);
        case 64: return launch<64>(qkv, ldq, out, out_dtype, ldo, B, T, heads, round_tf32, st);
        case 96: return launch<96>(qkv, ldq, out, out_dtype, ldo, B, T, heads, round_tf32, st);
        case 128: return launch<128>(qkv, ldq, out, out_dtype, ldo, B, T, heads, round_tf32, st);
        case 192: return launch<192>(qkv, ldq, out, out_dtype, ldo, B, T, heads, round_tf32, st);
        default:
            hl_set_error("hl_attention: unsupported head width %d (supported: 32,64,96,128,192)", ch);
            return HL_E_UNSUPPORTED;
    }
}
