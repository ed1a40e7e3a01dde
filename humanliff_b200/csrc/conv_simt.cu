// CUDA-core implicit-GEMM convolution (NHWC), fp32 accumulate; operands fp32 or fp16.  General shapes: ksize 1/3, stride 1/2, any
// Cin/Cout, optional nearest-x2 upsampled input, bias + residual epilogue.  This is the exact-fp32
// companion of the tcgen05 kernel in conv_tc.cu: it serves the shapes the tensor-core path does
// not tile (stride 2, tiny feature maps, W not a power of two) and the `precision="fp32"` mode.
//
// GEMM view: M = B*Ho*Wo output pixels, N = Cout, K = taps*Cin.  64x64 output tile per CTA,
// K chunk 16, 256 threads each owning a 4x4 micro-tile; A/B chunks are register-prefetched while
// the previous chunk is multiplied out of shared memory.
#include "common.cuh"

#include <cuda_fp16.h>

namespace {

constexpr int BM = 64, BN = 64, BK = 16;

__device__ __forceinline__ float ldf(const float *p, int64_t i) { return p[i]; }
__device__ __forceinline__ float ldf(const __half *p, int64_t i) { return __half2float(p[i]); }

template <typename T>
struct ConvArgs {
    const T *x; int ldx;
    const T *w; const float *bias;
    const float *res; int ldr;
    void *y; int ldy; int y_f16;
    int B, H, W, Cin, Cout, Cout_pad, ksize, stride, ups;
    int Ho, Wo;      // output size
    int Hi, Wi;      // logical input size (after optional upsample)
    int64_t M;
    int K;           // taps * Cin per operand pass
    int npass, a_off[3], b_slab[3];   // hi + lo operand passes (HL_CONV_SPLIT3 / SPLIT2P), as in conv_tc.cu
};

template <typename T>
__global__ void __launch_bounds__(256) k_conv_simt(ConvArgs<T> a) {
    hl_pdl_enter();
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    const int tid = threadIdx.x;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int pad = a.ksize / 2;

    // loader mapping: thread -> (row within tile, 4 consecutive k)
    const int lrow = tid >> 2;        // 0..63
    const int lk = (tid & 3) * 4;     // 0,4,8,12
    // decode the A row (output pixel) once
    int64_t m = m0 + lrow;
    bool mvalid = m < a.M;
    int ox = 0, oy = 0, ob = 0;
    if (mvalid) {
        ox = (int)(m % a.Wo);
        int64_t r = m / a.Wo;
        oy = (int)(r % a.Ho);
        ob = (int)(r / a.Ho);
    }
    const int nrow = n0 + lrow;       // B row (output channel)
    const bool nvalid = nrow < a.Cout;

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const int ty = tid >> 4, tx = tid & 15;   // 16x16 thread grid, 4x4 each
    float ra[4], rb[4];

    auto load_chunk = [&](int k0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int k = k0 + lk + j;
            float va = 0.f, vb = 0.f;
            if (k < a.K * a.npass) {
                const int pass = k / a.K, kk = k - pass * a.K;
                int tap = kk / a.Cin;
                int ci = kk - tap * a.Cin;
                if (mvalid) {
                    int ky = tap / a.ksize, kx = tap - ky * a.ksize;
                    int iy = oy * a.stride + ky - pad;
                    int ix = ox * a.stride + kx - pad;
                    if (iy >= 0 && iy < a.Hi && ix >= 0 && ix < a.Wi) {
                        if (a.ups) { iy >>= 1; ix >>= 1; }
                        va = ldf(a.x, (((int64_t)ob * a.H + iy) * a.W + ix) * a.ldx + a.a_off[pass] + ci);
                    }
                }
                if (nvalid) vb = ldf(a.w, ((int64_t)(a.b_slab[pass] * a.ksize * a.ksize + tap) * a.Cout_pad + nrow) * a.Cin + ci);
            }
            ra[j] = va;
            rb[j] = vb;
        }
    };

    load_chunk(0);
    const int Ktot = a.K * a.npass;
    for (int k0 = 0; k0 < Ktot; k0 += BK) {
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            As[lk + j][lrow] = ra[j];
            Bs[lk + j][lrow] = rb[j];
        }
        __syncthreads();
        if (k0 + BK < Ktot) load_chunk(k0 + BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float4 av = *reinterpret_cast<const float4 *>(&As[k][ty * 4]);
            float4 bv = *reinterpret_cast<const float4 *>(&Bs[k][tx * 4]);
            float aa[4] = {av.x, av.y, av.z, av.w};
            float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
        }
    }

#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int64_t mm = m0 + ty * 4 + i;
        if (mm >= a.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int n = n0 + tx * 4 + j;
            if (n >= a.Cout) continue;
            float v = acc[i][j] + (a.bias ? a.bias[n] : 0.f);
            if (a.res) v += a.res[mm * a.ldr + n];
            if (a.y_f16 == 2) {                          // scaled hi | lo pair (HL_CONV_OUT_F16_SPLIT)
                const float vs = v * HL_OP_SCALE;
                const __half h = __float2half_rn(vs);
                reinterpret_cast<__half *>(a.y)[mm * a.ldy + n] = h;
                reinterpret_cast<__half *>(a.y)[mm * a.ldy + a.Cout + n] = __float2half_rn(vs - __half2float(h));
            } else if (a.y_f16) {
                reinterpret_cast<__half *>(a.y)[mm * a.ldy + n] = __float2half_rn(v);
            } else {
                reinterpret_cast<float *>(a.y)[mm * a.ldy + n] = v;
            }
        }
    }
}

template <typename T>
static int launch_simt(const T *x, int ldx, const T *wpk, const float *bias, const float *residual, int ldr,
                       void *y, int ldy, int B, int H, int W, int Cin, int Cout, int ksize, int stride,
                       int flags, cudaStream_t stream) {
    ConvArgs<T> a;
    a.x = x; a.ldx = ldx; a.w = wpk; a.bias = bias; a.res = residual; a.ldr = ldr; a.y = y; a.ldy = ldy; a.y_f16 = (flags & HL_CONV_OUT_F16_SPLIT) ? 2 : (flags & HL_CONV_OUT_F16) ? 1 : 0;
    a.npass = (flags & HL_CONV_SPLIT3) ? 3 : (flags & (HL_CONV_SPLIT2P | HL_CONV_SPLIT2A)) ? 2 : 1;
    for (int q = 0; q < 3; ++q) { a.a_off[q] = 0; a.b_slab[q] = 0; }
    if (a.npass == 3) { a.a_off[1] = Cin; a.b_slab[2] = 1; }
    if (flags & HL_CONV_SPLIT2P) a.b_slab[1] = 1;
    if (flags & HL_CONV_SPLIT2A) a.a_off[1] = Cin;
    a.B = B; a.H = H; a.W = W; a.Cin = Cin; a.Cout = Cout; a.Cout_pad = hl_conv_cout_pad(Cout);
    a.ksize = ksize; a.stride = stride; a.ups = (flags & HL_CONV_UPSAMPLE2X) ? 1 : 0;
    a.Hi = a.ups ? 2 * H : H;
    a.Wi = a.ups ? 2 * W : W;
    int pad = ksize / 2;
    a.Ho = (a.Hi + 2 * pad - ksize) / stride + 1;
    a.Wo = (a.Wi + 2 * pad - ksize) / stride + 1;
    a.M = (int64_t)B * a.Ho * a.Wo;
    a.K = ksize * ksize * Cin;
    dim3 grid(hl_cdiv(a.M, BM), hl_cdiv(Cout, BN));
    HL_CHECK_CUDA(hl_launch(k_conv_simt<T>, grid, dim3(256), 0, stream, a));
    HL_CHECK_LAUNCH();
    return HL_OK;
}

}  // namespace

int hl_conv2d_simt(const void *x, int x_dtype, int ldx, const void *wpk, const float *bias,
                   const float *residual, int ldr, void *y, int ldy, int B, int H, int W, int Cin, int Cout,
                   int ksize, int stride, int flags, cudaStream_t stream) {
    if (x_dtype == HL_DT_F16)
        return launch_simt<__half>((const __half *)x, ldx, (const __half *)wpk, bias, residual, ldr, y, ldy, B, H,
                                   W, Cin, Cout, ksize, stride, flags, stream);
    return launch_simt<float>((const float *)x, ldx, (const float *)wpk, bias, residual, ldr, y, ldy, B, H, W,
                              Cin, Cout, ksize, stride, flags, stream);
}
