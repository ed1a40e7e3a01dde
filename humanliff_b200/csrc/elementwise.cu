// Memory-bound kernels of the denoise path: layout changes, GroupNorm32 + SiLU + FiLM,
// embeddings, DDPM posterior update.  All are HBM-bound streaming kernels: 128-bit accesses,
// grids sized in multiples of the SM count.
#include "common.cuh"

#include <cuda_fp16.h>

// ------------------------------------------------------------------------------------------
// diagnostics
// ------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

void hl_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" int hl_version(void) { return 100; }
extern "C" const char *hl_last_error(void) { return g_err; }

static int g_num_sms = 0;
int hl_num_sms() {
    if (g_num_sms == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            g_num_sms = n;
        else
            g_num_sms = 148;
    }
    return g_num_sms;
}

// Operand mode word (HL_OP_* in the header): bit 0 = TF32-round an fp32 operand, bit 1 = store value * 2^-4 (the
// packed weights carry 2^4: a raw residual-stream operand keeps fp16 range up to 1.0e6), bit 2 = store an fp16
// hi | lo pair (lo = fp16(v - hi), ~22 significant bits) with lo at + (mode >> 8) elements.
__device__ __forceinline__ uint32_t h2_bits(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&h);
}
__device__ __forceinline__ uint32_t lo_bits(float a, float b, uint32_t hi) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&hi));
    return h2_bits(a - f.x, b - f.y);
}
// store 4 consecutive channels of an operand buffer: fp32 (optionally TF32-rounded) or fp16
__device__ __forceinline__ void store_quad(void *dst, int dtype, int64_t idx, float4 v, int mode) {
    if (dtype == HL_DT_F16) {
        if (mode & HL_OP_SCALED) { v.x *= HL_OP_SCALE; v.y *= HL_OP_SCALE; v.z *= HL_OP_SCALE; v.w *= HL_OP_SCALE; }
        uint2 u;
        u.x = h2_bits(v.x, v.y);
        u.y = h2_bits(v.z, v.w);
        *reinterpret_cast<uint2 *>(reinterpret_cast<__half *>(dst) + idx) = u;
        if (mode & HL_OP_SPLIT) {
            uint2 l;
            l.x = lo_bits(v.x, v.y, u.x);
            l.y = lo_bits(v.z, v.w, u.y);
            *reinterpret_cast<uint2 *>(reinterpret_cast<__half *>(dst) + idx + (mode >> 8)) = l;
        }
    } else {
        if (mode & HL_OP_TF32) {
            v.x = hl_rna_tf32(v.x); v.y = hl_rna_tf32(v.y);
            v.z = hl_rna_tf32(v.z); v.w = hl_rna_tf32(v.w);
        }
        *reinterpret_cast<float4 *>(reinterpret_cast<float *>(dst) + idx) = v;
    }
}

// 8 consecutive channels: one 128-bit store for fp16 operands, two for fp32
__device__ __forceinline__ void store_oct(void *dst, int dtype, int64_t idx, float4 a, float4 b, int mode) {
    if (dtype == HL_DT_F16) {
        if (mode & HL_OP_SCALED) {
            a.x *= HL_OP_SCALE; a.y *= HL_OP_SCALE; a.z *= HL_OP_SCALE; a.w *= HL_OP_SCALE;
            b.x *= HL_OP_SCALE; b.y *= HL_OP_SCALE; b.z *= HL_OP_SCALE; b.w *= HL_OP_SCALE;
        }
        uint4 u;
        u.x = h2_bits(a.x, a.y);
        u.y = h2_bits(a.z, a.w);
        u.z = h2_bits(b.x, b.y);
        u.w = h2_bits(b.z, b.w);
        *reinterpret_cast<uint4 *>(reinterpret_cast<__half *>(dst) + idx) = u;
        if (mode & HL_OP_SPLIT) {
            uint4 l;
            l.x = lo_bits(a.x, a.y, u.x);
            l.y = lo_bits(a.z, a.w, u.y);
            l.z = lo_bits(b.x, b.y, u.z);
            l.w = lo_bits(b.z, b.w, u.w);
            *reinterpret_cast<uint4 *>(reinterpret_cast<__half *>(dst) + idx + (mode >> 8)) = l;
        }
    } else {
        store_quad(dst, dtype, idx, a, mode);
        store_quad(dst, dtype, idx + 4, b, mode);
    }
}

static inline int grid_for(int64_t work_items, int per_block, int max_waves = 8) {
    int64_t g = (work_items + per_block - 1) / per_block;
    int64_t cap = (int64_t)hl_num_sms() * max_waves;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

// ------------------------------------------------------------------------------------------
// NCHW <-> NHWC.  One block transposes a [32 pixels x Cpad] tile through shared memory so that
// both the NCHW side (contiguous pixels) and the NHWC side (contiguous channels) are coalesced.
// ------------------------------------------------------------------------------------------
__global__ void k_nchw_to_nhwc(const float *__restrict__ src, const float *__restrict__ src2,
                               void *__restrict__ dst, int dst_dtype, int C, int HW, int ld, int round_tf32,
                               int64_t n_tiles, int tiles_per_img) {
    hl_pdl_enter();
    __shared__ float tile[32][33];
    for (int64_t tidx = blockIdx.x; tidx < n_tiles; tidx += gridDim.x) {
        int b = (int)(tidx / tiles_per_img);
        int p0 = (int)(tidx % tiles_per_img) * 32;
        for (int c0 = 0; c0 < ld; c0 += 32) {
            // load: threadIdx.x -> pixel, threadIdx.y -> channel
            for (int cy = threadIdx.y; cy < 32; cy += blockDim.y) {
                int c = c0 + cy, p = p0 + threadIdx.x;
                float v = 0.f;
                if (c < C && p < HW) {
                    int64_t o = ((int64_t)b * C + c) * HW + p;
                    v = src[o];
                    if (src2) v += src2[o];
                }
                tile[cy][threadIdx.x] = v;
            }
            __syncthreads();
            for (int py = threadIdx.y; py < 32; py += blockDim.y) {
                int c = c0 + threadIdx.x, p = p0 + py;
                if (c < ld && p < HW) {
                    float v = tile[threadIdx.x][py];
                    const int64_t o = ((int64_t)b * HW + p) * ld + c;
                    if (dst_dtype == HL_DT_F16) {
                        const __half h = __float2half_rn(v);
                        if (!(round_tf32 & HL_OP_SPLIT)) {
                            reinterpret_cast<__half *>(dst)[o] = h;
                        } else if (c < (round_tf32 >> 8)) {      // packed pair inside the row: [hi | lo at + lo_off]
                            reinterpret_cast<__half *>(dst)[o] = h;
                            reinterpret_cast<__half *>(dst)[o + (round_tf32 >> 8)] = __float2half_rn(v - __half2float(h));
                        }
                    } else {
                        reinterpret_cast<float *>(dst)[o] = (round_tf32 & HL_OP_TF32) ? hl_rna_tf32(v) : v;
                    }
                }
            }
            __syncthreads();
        }
    }
}

extern "C" int hl_nchw_to_nhwc(const float *src, const float *src2, void *dst, int dst_dtype, int B, int C,
                               int HW, int ld, int round_tf32, void *stream) {
    HL_CHECK_ARG(src && dst && B > 0 && C > 0 && HW > 0 && ld >= C);
    int tiles_per_img = hl_cdiv(HW, 32);
    int64_t n_tiles = (int64_t)B * tiles_per_img;
    dim3 blk(32, 8);
    HL_CHECK_CUDA(hl_launch(k_nchw_to_nhwc, dim3(grid_for(n_tiles, 1, 16)), dim3(blk), 0, (cudaStream_t)stream, 
        src, src2, dst, dst_dtype, C, HW, ld, round_tf32, n_tiles, tiles_per_img));
    HL_CHECK_LAUNCH();
    return HL_OK;
}

__global__ void k_nhwc_to_nchw(const float *__restrict__ src, int ld, int off2, float *__restrict__ dst,
                               int C, int HW, int64_t n_tiles, int tiles_per_img) {
    hl_pdl_enter();
    __shared__ float tile[32][33];
    for (int64_t tidx = blockIdx.x; tidx < n_tiles; tidx += gridDim.x) {
        int b = (int)(tidx / tiles_per_img);
        int p0 = (int)(tidx % tiles_per_img) * 32;
        for (int c0 = 0; c0 < C; c0 += 32) {
            for (int py = threadIdx.y; py < 32; py += blockDim.y) {
                int c = c0 + threadIdx.x, p = p0 + py;
                float v = 0.f;
                if (c < C && p < HW) {
                    const float *row = src + ((int64_t)b * HW + p) * ld + c;
                    v = off2 ? row[0] + row[off2] : row[0];       // off2: the two halves of a split-weight conv result
                }
                tile[py][threadIdx.x] = v;
            }
            __syncthreads();
            for (int cy = threadIdx.y; cy < 32; cy += blockDim.y) {
                int c = c0 + cy, p = p0 + threadIdx.x;
                if (c < C && p < HW) dst[((int64_t)b * C + c) * HW + p] = tile[threadIdx.x][cy];
            }
            __syncthreads();
        }
    }
}

static int nhwc_to_nchw_launch(const float *src, int ld, int off2, float *dst, int B, int C, int HW, void *stream) {
    HL_CHECK_ARG(src && dst && B > 0 && C > 0 && HW > 0 && ld >= C && off2 >= 0 && (off2 == 0 || ld >= off2 + C));
    int tiles_per_img = hl_cdiv(HW, 32);
    int64_t n_tiles = (int64_t)B * tiles_per_img;
    dim3 blk(32, 8);
    HL_CHECK_CUDA(hl_launch(k_nhwc_to_nchw, dim3(grid_for(n_tiles, 1, 16)), dim3(blk), 0, (cudaStream_t)stream,
        src, ld, off2, dst, C, HW, n_tiles, tiles_per_img));
    HL_CHECK_LAUNCH();
    return HL_OK;
}

extern "C" int hl_nhwc_to_nchw(const float *src, int ld, float *dst, int B, int C, int HW,
                               void *stream) {
    return nhwc_to_nchw_launch(src, ld, 0, dst, B, C, HW, stream);
}

extern "C" int hl_nhwc_to_nchw_sum2(const float *src, int ld, int off2, float *dst, int B, int C, int HW,
                                    void *stream) {
    HL_CHECK_ARG(off2 > 0);
    return nhwc_to_nchw_launch(src, ld, off2, dst, B, C, HW, stream);
}

// ------------------------------------------------------------------------------------------
// upsample, operand cast -- float4 over channels (all C are multiples of 4)
// ------------------------------------------------------------------------------------------
__global__ void k_upsample2x(const float *__restrict__ src, int lds, void *__restrict__ dst, int dst_dtype,
                             int ldd, int H, int W, int C, int round_tf32, int64_t total) {
    hl_pdl_enter();
    int q = C >> 2, W2 = 2 * W, H2 = 2 * H;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        int64_t p = i / q;
        int j = (int)(i - p * q);
        int ox = (int)(p % W2);
        int64_t r = p / W2;
        int oy = (int)(r % H2);
        int64_t b = r / H2;
        int64_t sp = (b * H + (oy >> 1)) * W + (ox >> 1);
        float4 v = *reinterpret_cast<const float4 *>(src + sp * lds + 4 * j);
        store_quad(dst, dst_dtype, p * ldd + 4 * j, v, round_tf32);
    }
}

extern "C" int hl_upsample2x(const float *src, int lds, void *dst, int dst_dtype, int ldd, int B, int H, int W,
                             int C, int round_tf32, void *stream) {
    HL_CHECK_ARG(src && dst && B > 0 && H > 0 && W > 0 && C % 4 == 0 && lds % 4 == 0 && ldd % 4 == 0);
    int64_t total = (int64_t)B * 4 * H * W * (C / 4);
    HL_CHECK_CUDA(hl_launch(k_upsample2x, dim3(grid_for(total, 256 * 4)), dim3(256), 0, (cudaStream_t)stream, src, lds, dst, dst_dtype, ldd, H,
                                                                             W, C, round_tf32, total));
    HL_CHECK_LAUNCH();
    return HL_OK;
}

__global__ void k_cast_operand(const float *__restrict__ src, int lds, void *__restrict__ dst, int dst_dtype,
                               int ldd, int C, int64_t npix, int round_tf32) {
    hl_pdl_enter();
    int q = C >> 2;
    int64_t total = npix * q;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        int64_t p = i / q;
        int j = (int)(i - p * q);
        float4 v = *reinterpret_cast<const float4 *>(src + p * lds + 4 * j);
        store_quad(dst, dst_dtype, p * ldd + 4 * j, v, round_tf32);
    }
}

extern "C" int hl_cast_operand(const float *src, int lds, void *dst, int dst_dtype, int ldd, int C,
                               int64_t npix, int round_tf32, void *stream) {
    HL_CHECK_ARG(src && dst && C % 4 == 0 && lds % 4 == 0 && ldd % 4 == 0 && npix > 0);
    HL_CHECK_ARG(dst_dtype == HL_DT_F32 || dst_dtype == HL_DT_F16);
    HL_CHECK_CUDA(hl_launch(k_cast_operand, dim3(grid_for(npix * (C / 4), 256 * 4)), dim3(256), 0, (cudaStream_t)stream, src, lds, dst, dst_dtype,
                                                                                        ldd, C, npix, round_tf32));
    HL_CHECK_LAUNCH();
    return HL_OK;
}

extern "C" int hl_zero(void *ptr, int64_t bytes, void *stream) {
    HL_CHECK_ARG(ptr && bytes >= 0);
    HL_CHECK_CUDA(cudaMemsetAsync(ptr, 0, (size_t)bytes, (cudaStream_t)stream));
    g_hl_pdl_skip = 1;          // a memset node is not a programmatic-launch primary
    return HL_OK;
}

int g_hl_pdl = 0, g_hl_pdl_skip = 0;
long long g_hl_launches = 0;
extern "C" int64_t hl_launch_count(void) { return (int64_t)g_hl_launches; }
extern "C" int hl_set_pdl(int on) {
    const int prev = g_hl_pdl;
    if (on >= 0) g_hl_pdl = on ? 1 : 0;
    return prev;
}
extern "C" void hl_pdl_barrier(void) { g_hl_pdl_skip = 1; }

// ------------------------------------------------------------------------------------------
// embeddings
// ------------------------------------------------------------------------------------------
__global__ void k_timestep_embedding(const float *__restrict__ t, const float *__restrict__ freqs, int B,
                                     int dim, float *__restrict__ out) {
    hl_pdl_enter();
    int half = dim / 2;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * half) return;
    int b = i / half, k = i % half;
    float a = __fmul_rn(t[b], freqs[k]);
    out[b * dim + k] = cosf(a);
    out[b * dim + half + k] = sinf(a);
    if ((dim & 1) && k == 0) out[b * dim + dim - 1] = 0.f;
}

extern "C" int hl_timestep_embedding(const float *t, const float *freqs, int B, int dim, float *out,
                                     void *stream) {
    HL_CHECK_ARG(t && freqs && out && B > 0 && dim >= 2);
    int n = B * (dim / 2);
    HL_CHECK_CUDA(hl_launch(k_timestep_embedding, dim3(hl_cdiv(n, 128)), dim3(128), 0, (cudaStream_t)stream, t, freqs, B, dim, out));
    HL_CHECK_LAUNCH();
    return HL_OK;
}

// One warp per output feature; x (activated) is staged in shared memory in chunks of 8 rows.
// The weight matrix is streamed exactly once per 8 batch rows (HBM-bound GEMV).
#define LIN_ROWS 8
__global__ void k_linear_small(const float *__restrict__ x, const float *__restrict__ W,
                               const float *__restrict__ bias, float *__restrict__ y, int B, int in_f,
                               int out_f, int silu_in, const float *__restrict__ add_table,
                               const int64_t *__restrict__ add_idx) {
    hl_pdl_enter();
    extern __shared__ float xs[];  // [LIN_ROWS][in_f]
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    for (int b0 = 0; b0 < B; b0 += LIN_ROWS) {
        int nb = min(LIN_ROWS, B - b0);
        __syncthreads();
        for (int i = threadIdx.x; i < nb * in_f; i += blockDim.x) {
            float v = x[(int64_t)b0 * in_f + i];
            xs[i] = silu_in ? hl_silu(v) : v;
        }
        __syncthreads();
        for (int o = blockIdx.x * nwarp + warp; o < out_f; o += gridDim.x * nwarp) {
            float acc[LIN_ROWS];
#pragma unroll
            for (int r = 0; r < LIN_ROWS; ++r) acc[r] = 0.f;
            const float *wr = W + (int64_t)o * in_f;
            for (int i = lane; i < in_f; i += 32) {
                float w = wr[i];
#pragma unroll
                for (int r = 0; r < LIN_ROWS; ++r)
                    if (r < nb) acc[r] = fmaf(w, xs[r * in_f + i], acc[r]);
            }
#pragma unroll
            for (int r = 0; r < LIN_ROWS; ++r) {
                if (r < nb) {
                    float s = hl_warp_sum(acc[r]);
                    if (lane == 0) {
                        s += bias ? bias[o] : 0.f;
                        if (add_table) s += add_table[add_idx[b0 + r] * out_f + o];
                        y[(int64_t)(b0 + r) * out_f + o] = s;
                    }
                }
            }
        }
    }
}

extern "C" int hl_linear_small(const float *x, const float *W, const float *bias, float *y, int B,
                               int in_f, int out_f, int silu_in, const float *add_table,
                               const int64_t *add_idx, void *stream) {
    HL_CHECK_ARG(x && W && y && B > 0 && in_f > 0 && out_f > 0);
    HL_CHECK_ARG((add_table == nullptr) == (add_idx == nullptr));
    size_t smem = (size_t)LIN_ROWS * in_f * sizeof(float);
    HL_CHECK_ARG(smem <= 48 * 1024);
    int warps = 8;
    int grid = hl_cdiv(out_f, warps);
    int cap = hl_num_sms() * 4;
    if (grid > cap) grid = cap;
    HL_CHECK_CUDA(hl_launch(k_linear_small, dim3(grid), dim3(warps * 32), smem, (cudaStream_t)stream, x, W, bias, y, B, in_f, out_f,
                                                                    silu_in, add_table, add_idx));
    HL_CHECK_LAUNCH();
    return HL_OK;
}

// ------------------------------------------------------------------------------------------
// GroupNorm32.  Statistics are kept PER CHANNEL (sum, sum of squares, fp64) so that they can be
// produced by whoever writes the tensor (the tcgen05 conv epilogue, or k_gn_stats below) and folded
// into any group layout by the consumer -- including groups straddling the two halves of the
// decoder's never-materialised concat.
// stats kernel: each block owns a slab of pixels of one sample, each thread a fixed channel quad
// (128-bit coalesced loads); per-channel partials are combined in shared memory and added to the
// fp64 accumulators (fp64 atomics: arrival order changes the result far below fp32 resolution).
// ------------------------------------------------------------------------------------------
#define GN_MAX_C 2048
// four consecutive channels of an fp32 tensor or of an fp16 tensor (HL_CONV_OUT_F16 results) as floats
__device__ __forceinline__ float4 load_quad(const void *x, int64_t idx, bool f16) {
    if (!f16) return *reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(x) + idx);
    const uint2 u = *reinterpret_cast<const uint2 *>(reinterpret_cast<const __half *>(x) + idx);
    const float2 a = __half22float2(*reinterpret_cast<const __half2 *>(&u.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2 *>(&u.y));
    return make_float4(a.x, a.y, b.x, b.y);
}

__global__ void k_gn_stats(const void *__restrict__ x, int x_f16, int ldx, int HW, int C, int pix_per_block,
                           double *__restrict__ stats, int stats_ld) {
    hl_pdl_enter();
    __shared__ double s_sum[GN_MAX_C];   // fp64: atomics then commute to ~1e-16 -> reproducible statistics
    __shared__ double s_sq[GN_MAX_C];
    int b = blockIdx.y;
    int q = C >> 2;
    int lanes_p = blockDim.x / q;  // pixels processed concurrently (blockDim.x is a multiple of q)
    int tq = threadIdx.x % q, tp = threadIdx.x / q;
    for (int i = threadIdx.x; i < C; i += blockDim.x) { s_sum[i] = 0.0; s_sq[i] = 0.0; }
    __syncthreads();
    int p0 = blockIdx.x * pix_per_block;
    int p1 = min(HW, p0 + pix_per_block);
    float4 s = make_float4(0, 0, 0, 0), ss = make_float4(0, 0, 0, 0);
    if (tp < lanes_p) {
        const int64_t base = (int64_t)b * HW * ldx + 4 * tq;
        for (int p = p0 + tp; p < p1; p += lanes_p) {
            float4 v = load_quad(x, base + (int64_t)p * ldx, x_f16 != 0);
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
            ss.x = fmaf(v.x, v.x, ss.x); ss.y = fmaf(v.y, v.y, ss.y);
            ss.z = fmaf(v.z, v.z, ss.z); ss.w = fmaf(v.w, v.w, ss.w);
        }
        atomicAdd(&s_sum[4 * tq + 0], (double)s.x); atomicAdd(&s_sq[4 * tq + 0], (double)ss.x);
        atomicAdd(&s_sum[4 * tq + 1], (double)s.y); atomicAdd(&s_sq[4 * tq + 1], (double)ss.y);
        atomicAdd(&s_sum[4 * tq + 2], (double)s.z); atomicAdd(&s_sq[4 * tq + 2], (double)ss.z);
        atomicAdd(&s_sum[4 * tq + 3], (double)s.w); atomicAdd(&s_sq[4 * tq + 3], (double)ss.w);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        double *dst = stats + ((int64_t)b * stats_ld + c) * 2;
        atomicAdd(dst, s_sum[c]);
        atomicAdd(dst + 1, s_sq[c]);
    }
}

int hl_gn_stats_launch(const void *x, int x_f16, int ldx, int B, int HW, int C, double *stats, int stats_ld,
                       cudaStream_t stream) {
    HL_CHECK_ARG(x && stats && B > 0 && HW > 0 && C > 0 && C <= GN_MAX_C);
    HL_CHECK_ARG(C % 4 == 0 && ldx % 4 == 0 && ldx >= C && stats_ld >= C);
    int q = C / 4;
    HL_CHECK_ARG(q <= 512);
    int threads = (512 / q) * q;   // largest multiple of q not above 512
    int lanes_p = threads / q;
    // ~4 waves over the SMs, but at least 8 pixel rounds per block so the smem fold amortises
    int64_t want_blocks = (int64_t)hl_num_sms() * 4 / B;
    if (want_blocks < 1) want_blocks = 1;
    int pix_per_block = hl_cdiv(HW, want_blocks);
    int min_ppb = lanes_p * 8;
    if (pix_per_block < min_ppb) pix_per_block = min_ppb;
    dim3 grid(hl_cdiv(HW, pix_per_block), B);
    HL_CHECK_CUDA(hl_launch(k_gn_stats, dim3(grid), dim3(threads), 0, stream, x, x_f16, ldx, HW, C, pix_per_block, stats, stats_ld));
    HL_CHECK_LAUNCH();
    return HL_OK;
}

extern "C" int hl_gn_stats(const float *x, int ldx, int B, int HW, int C, double *stats, int stats_ld,
                           void *stream) {
    return hl_gn_stats_launch(x, 0, ldx, B, HW, C, stats, stats_ld, (cudaStream_t)stream);
}

// apply: y = act(x * A[b,c] + Bc[b,c]) where A, Bc fold mean/rstd/gamma/beta and the FiLM
// scale/shift; A and Bc are built per block in shared memory from the per-channel fp64 sums.
__device__ __forceinline__ void gn_coefficients(const double *__restrict__ stats, int stats_ld,
                                                const float *__restrict__ gamma, const float *__restrict__ beta,
                                                const float *__restrict__ film, int film_ld, int b, int HW, int C,
                                                int groups, float eps, float *sA, float *sB, float *gmean,
                                                float *grstd) {
    const int cpg = C / groups;
    const double n = (double)HW * cpg;
    // group statistics: 8 lanes per group, independent loads, shuffle reduction (a serial loop over the
    // group's channels cost ~15 us of dependent L2 latency per launch on the small feature maps)
    for (int g0 = 0; g0 < groups; g0 += blockDim.x / 8) {
        const int g = g0 + (threadIdx.x >> 3), l = threadIdx.x & 7;
        double a = 0.0, a2 = 0.0;
        if (g < groups) {
            const double2 *st = reinterpret_cast<const double2 *>(stats + ((int64_t)b * stats_ld + g * cpg) * 2);
            for (int c = l; c < cpg; c += 8) {
                const double2 v = st[c];
                a += v.x;
                a2 += v.y;
            }
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            a2 += __shfl_xor_sync(0xffffffffu, a2, o);
        }
        if (g < groups && l == 0) {
            double m = a / n;
            double var = a2 / n - m * m;
            if (var < 0.0) var = 0.0;
            gmean[g] = (float)m;
            grstd[g] = (float)(1.0 / sqrt(var + (double)eps));
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        int g = c / cpg;
        float ga = gamma[c] * grstd[g];
        float be = beta[c] - gmean[g] * ga;
        if (film) {
            float sc = 1.0f + film[(int64_t)b * film_ld + c];
            float sh = film[(int64_t)b * film_ld + C + c];
            ga *= sc;
            be = be * sc + sh;
        }
        sA[c] = ga;
        sB[c] = be;
    }
    __syncthreads();
}

// Generic kernel (any dtype / operand mode; the fp32 and TF32 plans, odd channel counts): each thread walks
// (pixel, channel-quad) pairs with an incrementally updated index (no divisions in the loop).
__global__ void k_gn_apply(const float *__restrict__ x, int ldx, const double *__restrict__ stats,
                           int stats_ld, const float *__restrict__ gamma, const float *__restrict__ beta,
                           const float *__restrict__ film, int film_ld, void *__restrict__ y, int y_dtype,
                           int ldy, void *__restrict__ raw, int ldraw, int HW, int C, int groups, float eps,
                           int silu, int op_mode, int pix_per_block) {
    hl_pdl_enter();
    // op_mode: bits 0-2 = HL_OP_* of y (lo offset of a split y in bits 8+), bit 3 = HL_OP_X_F16, bits 4-6 = HL_OP_* of
    // the raw copy (a split raw copy has its lo half at channel C)
    const int y_mode = (op_mode & 7) | (op_mode & ~0xFF);
    const int raw_mode = ((op_mode >> 4) & 7) | (C << 8);
    const bool x_f16 = (op_mode & HL_OP_X_F16) != 0;       // x is an fp16 tensor (a conv's HL_CONV_OUT_F16 result)
    __shared__ float sA[GN_MAX_C];
    __shared__ float sB[GN_MAX_C];
    __shared__ float gmean[64], grstd[64];
    const int b = blockIdx.y;
    gn_coefficients(stats, stats_ld, gamma, beta, film, film_ld, b, HW, C, groups, eps, sA, sB, gmean, grstd);
    const int p0 = blockIdx.x * pix_per_block;
    const int np = min(HW, p0 + pix_per_block) - p0;
    const int64_t pix0 = (int64_t)b * HW + p0;
    const bool fast_silu = (y_dtype == HL_DT_F16) || (y_mode & HL_OP_TF32);
    const int q = C >> 2;
    const int dp = blockDim.x / q, dj = blockDim.x % q;
    int p = threadIdx.x / q, j = threadIdx.x % q;
    while (p < np) {
        const int64_t pix = pix0 + p;
        const float4 v = load_quad(x, pix * ldx + 4 * j, x_f16);
        const float4 a = *reinterpret_cast<const float4 *>(&sA[4 * j]);
        const float4 c = *reinterpret_cast<const float4 *>(&sB[4 * j]);
        float4 o;
        o.x = fmaf(v.x, a.x, c.x); o.y = fmaf(v.y, a.y, c.y);
        o.z = fmaf(v.z, a.z, c.z); o.w = fmaf(v.w, a.w, c.w);
        if (silu) {
            if (fast_silu) { o.x = hl_silu_fast(o.x); o.y = hl_silu_fast(o.y); o.z = hl_silu_fast(o.z); o.w = hl_silu_fast(o.w); }
            else { o.x = hl_silu(o.x); o.y = hl_silu(o.y); o.z = hl_silu(o.z); o.w = hl_silu(o.w); }
        }
        store_quad(y, y_dtype, pix * ldy + 4 * j, o, y_mode);
        if (raw) store_quad(raw, y_dtype, pix * ldraw + 4 * j, v, raw_mode);
        p += dp;
        j += dj;
        if (j >= q) { j -= q; ++p; }
    }
}

// The fp16 plan's kernel, specialised at compile time (the one-kernel-for-every-mode version was 7,200 SASS
// instructions with every flag re-tested inside the unrolled loops: ncu showed instruction-fetch stalls of 2-3
// issue slots per instruction).  y is a plain fp16 operand; XF16: the input is fp16; RAW: additionally the
// 2^-4-scaled hi | lo copy of the input for a 1x1 skip conv (lo at channel C); YSPLIT: y itself is an unscaled
// hi | lo pair with lo at channel C (the output conv's operand).  A thread owns 8 fixed channels (coefficients in
// registers) and walks pixels with 8 x 16 B of loads in flight per round (fp32 input: 4 rows x two 128-bit loads;
// fp16 input: 8 rows x one); the first round is issued BEFORE the statistics prologue, whose two dependent L2
// round trips and barriers then hide under the first DRAM round trip.
template <bool XF16, bool SILU, bool RAW, bool YSPLIT>
__global__ void __launch_bounds__(256, 3)
k_gn_apply_f16(const void *__restrict__ x, int ldx, const double *__restrict__ stats, int stats_ld,
               const float *__restrict__ gamma, const float *__restrict__ beta, const float *__restrict__ film,
               int film_ld, __half *__restrict__ y, int ldy, __half *__restrict__ raw, int ldraw, int HW, int C,
               int groups, float eps, int pix_per_block) {
    hl_pdl_enter();
    __shared__ float sA[GN_MAX_C];
    __shared__ float sB[GN_MAX_C];
    __shared__ float gmean[64], grstd[64];
    const int b = blockIdx.y;
    const int p0 = blockIdx.x * pix_per_block;
    const int np = min(HW, p0 + pix_per_block) - p0;
    const int64_t pix0 = (int64_t)b * HW + p0;
    const int q8 = C >> 3;
    const int lanes_p = blockDim.x / q8;
    const int j8 = threadIdx.x % q8, tp = threadIdx.x / q8;
    constexpr int U = XF16 ? 8 : 4;
    uint4 buf[8];
    auto load_round = [&](int pb) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int pp = pb + u * lanes_p;
            if (pp < np) {
                if (XF16) {
                    buf[u] = __ldcs(reinterpret_cast<const uint4 *>(reinterpret_cast<const __half *>(x) + (pix0 + pp) * ldx + 8 * j8));
                } else {
                    const uint4 *src = reinterpret_cast<const uint4 *>(reinterpret_cast<const float *>(x) + (pix0 + pp) * ldx + 8 * j8);
                    buf[2 * u] = __ldcs(src);
                    buf[2 * u + 1] = __ldcs(src + 1);
                }
            }
        }
    };
    if (tp < np) load_round(tp);
    gn_coefficients(stats, stats_ld, gamma, beta, film, film_ld, b, HW, C, groups, eps, sA, sB, gmean, grstd);
    float ca[8], cb[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { ca[e] = sA[8 * j8 + e]; cb[e] = sB[8 * j8 + e]; }
    for (int pb = tp; pb < np;) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int pp = pb + u * lanes_p;
            if (pp >= np) break;
            float in[8];
            if (XF16) {
                const float2 f0 = __half22float2(*reinterpret_cast<const __half2 *>(&buf[u].x));
                const float2 f1 = __half22float2(*reinterpret_cast<const __half2 *>(&buf[u].y));
                const float2 f2 = __half22float2(*reinterpret_cast<const __half2 *>(&buf[u].z));
                const float2 f3 = __half22float2(*reinterpret_cast<const __half2 *>(&buf[u].w));
                in[0] = f0.x; in[1] = f0.y; in[2] = f1.x; in[3] = f1.y; in[4] = f2.x; in[5] = f2.y; in[6] = f3.x; in[7] = f3.y;
            } else {
                const float4 v0 = *reinterpret_cast<const float4 *>(&buf[2 * u]);
                const float4 v1 = *reinterpret_cast<const float4 *>(&buf[2 * u + 1]);
                in[0] = v0.x; in[1] = v0.y; in[2] = v0.z; in[3] = v0.w; in[4] = v1.x; in[5] = v1.y; in[6] = v1.z; in[7] = v1.w;
            }
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                o[e] = fmaf(in[e], ca[e], cb[e]);
                if (SILU) o[e] = hl_silu_fast(o[e]);
            }
            store_oct(y, HL_DT_F16, (pix0 + pp) * ldy + 8 * j8, make_float4(o[0], o[1], o[2], o[3]),
                      make_float4(o[4], o[5], o[6], o[7]), YSPLIT ? (HL_OP_SPLIT | (C << 8)) : 0);
            if (RAW && !XF16)
                store_oct(raw, HL_DT_F16, (pix0 + pp) * ldraw + 8 * j8, make_float4(in[0], in[1], in[2], in[3]),
                          make_float4(in[4], in[5], in[6], in[7]), HL_OP_SCALED | HL_OP_SPLIT | (C << 8));
        }
        pb += lanes_p * U;
        if (pb < np) load_round(pb);
    }
}

static int g_gn_blocks_per_sm = 0;      // blocks of k_gn_apply_f16 per SM and launch; 0 = one wave by occupancy (hl_gn_set_tuning)
extern "C" int hl_gn_set_tuning(int blocks_per_sm) {
    g_gn_blocks_per_sm = blocks_per_sm > 0 ? blocks_per_sm : 0;
    return HL_OK;
}

typedef void (*GnF16Kernel)(const void *, int, const double *, int, const float *, const float *, const float *, int,
                            __half *, int, __half *, int, int, int, int, float, int);

extern "C" int hl_gn_apply(const float *x, int ldx, const double *stats, int stats_ld, const float *gamma,
                           const float *beta, const float *film, int film_ld, void *y, int y_dtype, int ldy,
                           void *raw, int ldraw, int B, int HW, int C, int groups, float eps, int silu,
                           int round_tf32, void *stream) {
    HL_CHECK_ARG(x && stats && gamma && beta && y && B > 0 && HW > 0 && C > 0 && C <= GN_MAX_C);
    HL_CHECK_ARG(groups > 0 && groups <= 64 && C % groups == 0 && C % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0 &&
                 ldx >= C && ldy >= C && stats_ld >= C);
    HL_CHECK_ARG(y_dtype == HL_DT_F32 || y_dtype == HL_DT_F16);
    HL_CHECK_ARG(!raw || (ldraw % 4 == 0 && ldraw >= C));
    const int op_mode = round_tf32;
    const bool x_f16 = (op_mode & HL_OP_X_F16) != 0;
    HL_CHECK_ARG(!x_f16 || (((uintptr_t)x & 7) == 0 && !raw));
    const int y_mode = op_mode & 7, y_lo = op_mode >> 8, raw_mode = (op_mode >> 4) & 7;
    // the specialised fp16 kernel: whole channel octets, y a plain fp16 operand or an unscaled hi | lo pair with lo
    // at channel C, raw (if any) the scaled hi | lo pair
    const bool fast = y_dtype == HL_DT_F16 && C % 8 == 0 && C / 8 <= 256 && ldy % 8 == 0 && ((uintptr_t)y & 15) == 0 &&
                      ldx % 8 == 0 && ((uintptr_t)x & 15) == 0 &&
                      (y_mode == 0 || (y_mode == HL_OP_SPLIT && y_lo == C && ldy >= 2 * C)) &&
                      (!raw || (raw_mode == (HL_OP_SPLIT | HL_OP_SCALED) && ldraw % 8 == 0 && ldraw >= 2 * C &&
                                ((uintptr_t)raw & 15) == 0));
    int threads = 256;
    if (fast) threads = (256 / (C / 8)) * (C / 8);
    int per_sm = 8;
    GnF16Kernel kern = nullptr;
    if (fast) {
        const bool ys = y_mode == HL_OP_SPLIT, rw = raw != nullptr;
#define HL_GN_PICK(XF, SI, RW, YS) if (x_f16 == XF && (silu != 0) == SI && rw == RW && ys == YS) kern = k_gn_apply_f16<XF, SI, RW, YS>;
        HL_GN_PICK(false, true, false, false) HL_GN_PICK(false, false, false, false) HL_GN_PICK(false, true, true, false)
        HL_GN_PICK(false, false, true, false) HL_GN_PICK(true, true, false, false) HL_GN_PICK(true, false, false, false)
        HL_GN_PICK(false, true, false, true) HL_GN_PICK(false, false, false, true)
#undef HL_GN_PICK
    }
    if (kern) {
        // grid: whole waves of resident blocks (measured at 256^2, B = 4, 3 resident blocks / SM: 3 / SM = one wave
        // 57.7 us, 4 / SM = a wave and a third 68.2 us, 6 / SM = two waves 55.8 us -- profiles/r2_gn_sweep*.log);
        // the store-heavy variant with the raw hi | lo copy (10 B / element, 6 of them written) prefers many short
        // blocks (16 / SM: 167.6 us vs 185.5 us for one wave).  Small tensors are bounded below by min_ppb.
        // hl_gn_set_tuning overrides the blocks-per-SM figure.
        per_sm = g_gn_blocks_per_sm;
        if (per_sm <= 0) {
            int occ = 0;
            HL_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, 0));
            per_sm = raw ? 16 : 2 * (occ > 0 ? occ : 1);
        }
    }
    int64_t want_blocks = (int64_t)hl_num_sms() * per_sm / B;
    if (want_blocks < 1) want_blocks = 1;
    int pix_per_block = hl_cdiv(HW, want_blocks);
    int min_ppb = hl_cdiv(256 * 4 * 4, C / 4);  // >= 4 float4 per thread
    if (pix_per_block < min_ppb) pix_per_block = min_ppb;
    dim3 grid(hl_cdiv(HW, pix_per_block), B);
    if (kern) {
        HL_CHECK_CUDA(hl_launch(kern, dim3(grid), dim3(threads), 0, (cudaStream_t)stream, (const void *)x, ldx, stats, stats_ld,
                                gamma, beta, film, film_ld, (__half *)y, ldy, (__half *)raw, ldraw, HW, C, groups, eps,
                                pix_per_block));
    } else {
        HL_CHECK_CUDA(hl_launch(k_gn_apply, dim3(grid), dim3(threads), 0, (cudaStream_t)stream, x, ldx, stats, stats_ld, gamma,
                                beta, film, film_ld, y, y_dtype, ldy, raw, ldraw, HW, C, groups, eps, silu, op_mode,
                                pix_per_block));
    }
    HL_CHECK_LAUNCH();
    return HL_OK;
}

// ------------------------------------------------------------------------------------------
// DDPM posterior update -- 3 streams in, 2 out, 20 B / element (gaussian_diffusion.py:293-387)
// ------------------------------------------------------------------------------------------
__global__ void k_ddpm_step(const float *__restrict__ x, const float *__restrict__ eps,
                            const float *__restrict__ noise, const float *__restrict__ coef,
                            const float *__restrict__ sigma, const int64_t *__restrict__ t,
                            float *__restrict__ sample, float *__restrict__ x0out, int64_t n4,
                            int clip) {
    hl_pdl_enter();
    int b = blockIdx.y;
    int64_t ti = t[b];
    float c0 = coef[ti * 4 + 0], c1 = coef[ti * 4 + 1], c2 = coef[ti * 4 + 2], c3 = coef[ti * 4 + 3];
    float sg = sigma[ti];
    const float4 *x4 = reinterpret_cast<const float4 *>(x) + (int64_t)b * n4;
    const float4 *e4 = reinterpret_cast<const float4 *>(eps) + (int64_t)b * n4;
    const float4 *z4 = reinterpret_cast<const float4 *>(noise) + (int64_t)b * n4;
    float4 *s4 = reinterpret_cast<float4 *>(sample) + (int64_t)b * n4;
    float4 *p4 = x0out ? reinterpret_cast<float4 *>(x0out) + (int64_t)b * n4 : nullptr;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
         i += (int64_t)gridDim.x * blockDim.x) {
        float4 xv = x4[i], ev = e4[i], zv = z4[i], x0, s;
        // reference order: c0*x - c1*eps ; clamp ; c2*x0 + c3*x ; + sigma*noise (separate roundings)
        x0.x = __fsub_rn(__fmul_rn(c0, xv.x), __fmul_rn(c1, ev.x));
        x0.y = __fsub_rn(__fmul_rn(c0, xv.y), __fmul_rn(c1, ev.y));
        x0.z = __fsub_rn(__fmul_rn(c0, xv.z), __fmul_rn(c1, ev.z));
        x0.w = __fsub_rn(__fmul_rn(c0, xv.w), __fmul_rn(c1, ev.w));
        if (clip) {
            x0.x = fminf(fmaxf(x0.x, -1.f), 1.f); x0.y = fminf(fmaxf(x0.y, -1.f), 1.f);
            x0.z = fminf(fmaxf(x0.z, -1.f), 1.f); x0.w = fminf(fmaxf(x0.w, -1.f), 1.f);
        }
        s.x = __fadd_rn(__fadd_rn(__fmul_rn(c2, x0.x), __fmul_rn(c3, xv.x)), __fmul_rn(sg, zv.x));
        s.y = __fadd_rn(__fadd_rn(__fmul_rn(c2, x0.y), __fmul_rn(c3, xv.y)), __fmul_rn(sg, zv.y));
        s.z = __fadd_rn(__fadd_rn(__fmul_rn(c2, x0.z), __fmul_rn(c3, xv.z)), __fmul_rn(sg, zv.z));
        s.w = __fadd_rn(__fadd_rn(__fmul_rn(c2, x0.w), __fmul_rn(c3, xv.w)), __fmul_rn(sg, zv.w));
        s4[i] = s;
        if (p4) p4[i] = x0;
    }
}

extern "C" int hl_ddpm_step(const float *x, const float *eps, const float *noise, const float *coef,
                            const float *sigma, const int64_t *t, float *sample, float *pred_xstart,
                            int B, int64_t n, int clip, void *stream) {
    HL_CHECK_ARG(x && eps && noise && coef && sigma && t && sample && B > 0 && n > 0 && n % 4 == 0);
    int64_t n4 = n / 4;
    int gx = (int)((n4 + 255) / 256);
    int cap = hl_num_sms() * 8 / B;
    if (cap < 1) cap = 1;
    if (gx > cap) gx = cap;
    dim3 grid(gx, B);
    HL_CHECK_CUDA(hl_launch(k_ddpm_step, dim3(grid), dim3(256), 0, (cudaStream_t)stream, x, eps, noise, coef, sigma, t, sample,
                                                        pred_xstart, n4, clip));
    HL_CHECK_LAUNCH();
    return HL_OK;
}

// ------------------------------------------------------------------------------------------
// DDIM update (gaussian_diffusion.py:484-529): same streams as the DDPM step, epsilon re-derived from the
// clipped x0 exactly as the reference does, every operation separately rounded in the reference's order.
// ------------------------------------------------------------------------------------------
__global__ void k_ddim_step(const float *__restrict__ x, const float *__restrict__ eps,
                            const float *__restrict__ noise, const float *__restrict__ coef,
                            const float *__restrict__ sigma, const int64_t *__restrict__ t,
                            float *__restrict__ sample, float *__restrict__ x0out, int64_t n4, int clip) {
    hl_pdl_enter();
    int b = blockIdx.y;
    int64_t ti = t[b];
    const float c0 = coef[ti * 4 + 0], c1 = coef[ti * 4 + 1], ca = coef[ti * 4 + 2], cb = coef[ti * 4 + 3];
    const float sg = sigma[ti];
    const float4 *x4 = reinterpret_cast<const float4 *>(x) + (int64_t)b * n4;
    const float4 *e4 = reinterpret_cast<const float4 *>(eps) + (int64_t)b * n4;
    const float4 *z4 = noise ? reinterpret_cast<const float4 *>(noise) + (int64_t)b * n4 : nullptr;
    float4 *s4 = reinterpret_cast<float4 *>(sample) + (int64_t)b * n4;
    float4 *p4 = x0out ? reinterpret_cast<float4 *>(x0out) + (int64_t)b * n4 : nullptr;
    auto one = [&](float xv, float ev, float zv, float &x0, float &s) {
        const float cx = __fmul_rn(c0, xv);
        x0 = __fsub_rn(cx, __fmul_rn(c1, ev));
        if (clip) x0 = fminf(fmaxf(x0, -1.f), 1.f);
        const float e2 = __fdiv_rn(__fsub_rn(cx, x0), c1);            // _predict_eps_from_xstart
        const float mean = __fadd_rn(__fmul_rn(x0, ca), __fmul_rn(cb, e2));
        s = __fadd_rn(mean, __fmul_rn(sg, zv));
    };
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
         i += (int64_t)gridDim.x * blockDim.x) {
        const float4 xv = x4[i], ev = e4[i];
        const float4 zv = z4 ? z4[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 x0, s;
        one(xv.x, ev.x, zv.x, x0.x, s.x);
        one(xv.y, ev.y, zv.y, x0.y, s.y);
        one(xv.z, ev.z, zv.z, x0.z, s.z);
        one(xv.w, ev.w, zv.w, x0.w, s.w);
        s4[i] = s;
        if (p4) p4[i] = x0;
    }
}

extern "C" int hl_ddim_step(const float *x, const float *eps, const float *noise, const float *coef,
                            const float *sigma, const int64_t *t, float *sample, float *pred_xstart, int B,
                            int64_t n, int clip, void *stream) {
    HL_CHECK_ARG(x && eps && coef && sigma && t && sample && B > 0 && n > 0 && n % 4 == 0);
    int64_t n4 = n / 4;
    int gx = (int)((n4 + 255) / 256);
    int cap = hl_num_sms() * 8 / B;
    if (cap < 1) cap = 1;
    if (gx > cap) gx = cap;
    dim3 grid(gx, B);
    HL_CHECK_CUDA(hl_launch(k_ddim_step, dim3(grid), dim3(256), 0, (cudaStream_t)stream, x, eps, noise, coef, sigma, t, sample, pred_xstart, n4,
                                                        clip));
    HL_CHECK_LAUNCH();
    return HL_OK;
}
