// Fused tri-plane volume renderer: one persistent CTA walks rays; per ray it runs the whole
// reference chain without touching HBM in between --
//   coarse z (128) -> nine-plane bilinear gather -> density MLP -> importance resampling (PDF/CDF,
//   inverse-CDF with the caller's uniforms) -> merge-sort to 256 samples -> gather -> full MLP
//   (+ view-direction branch) -> alpha compositing.
// Follows recon_NeRF/lib/renderer.py:142-178 (NeRF_network, up_sample), :180-241 (render_core),
// :244-295 (render), :504-581 (project_onto_planes, sample_from_planes, sample_pdf),
// lib/fields.py:69-85 (PositionalEncoding) and run_nerf_batch.py:29-67 (render()).
//
// fp32 CUDA-core version.  Points of a ray form the GEMM M dimension (128-point tiles), activations
// live in shared memory as [feature][point] so each layer is a 128x128xK register-tiled product
// (8x8 micro-tile per thread) whose weights stream from L1/L2 (k-major packed, 267 KB total).
// Compulsory HBM traffic: 32 B in + 20 B out per ray (+512 B of uniforms) + the 9.4 MB texel array once.
#include "common.cuh"

int hl_num_sms();

namespace {

constexpr int NS = 128;     // samples per pass (n_samples == n_importance == 128, renderer.py:266)
constexpr int NT = 256;     // threads per CTA
constexpr int LDP = 128;    // points per tile (row pitch of the [feature][point] matrices)

struct RenderArgs {
    const float4 *tex;
    int R;
    const float *mlp;
    const float *o, *d, *near, *far, *u, *zc_in;
    unsigned long long seed;
    float bmin[3], bmax[3];
    float *rgb, *acc, *depth;
    long long n_rays;
    int clamp_depth;
};

// F.softplus(beta=1, threshold=20).  log1p(exp(x)) = max(x,0) + log(1 + exp(-|x|)).
__device__ __forceinline__ float softplus_fast(float x) {
    float r = fmaxf(x, 0.f) + __logf(1.0f + __expf(-fabsf(x)));
    return x > 20.f ? x : r;
}
__device__ __forceinline__ float softplus_acc(float x) { return x > 20.f ? x : log1pf(expf(x)); }

__device__ __forceinline__ float uniform_hash(unsigned long long seed, unsigned long long ray, int i) {
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (ray * 128ull + (unsigned long long)i + 1ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (float)(z >> 40) * (1.0f / 16777216.0f);
}

// acc[i][j] += sum_k in_s[k][pt_i] * Wt[k][og*8 + j];  thread's points: pg*4..+3 and 64+pg*4..+3
__device__ __forceinline__ void mac8x8(float (&acc)[8][8], const float *__restrict__ Wt,
                                       const float *in_s, int K, int pg, int og) {
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
        const float4 a0 = *reinterpret_cast<const float4 *>(in_s + k * LDP + pg * 4);
        const float4 a1 = *reinterpret_cast<const float4 *>(in_s + k * LDP + 64 + pg * 4);
        const float4 w0 = __ldg(reinterpret_cast<const float4 *>(Wt + k * 128 + og * 8));
        const float4 w1 = __ldg(reinterpret_cast<const float4 *>(Wt + k * 128 + og * 8 + 4));
        const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
}

template <bool ACT>
__device__ __forceinline__ void store8x8(const float (&acc)[8][8], const float *__restrict__ bias,
                                         float *out_s, int pg, int og) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float b = __ldg(bias + og * 8 + j);
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            v[i] = acc[i][j] + b;
            if (ACT) v[i] = softplus_fast(v[i]);
        }
        float *row = out_s + (og * 8 + j) * LDP;
        *reinterpret_cast<float4 *>(row + pg * 4) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4 *>(row + 64 + pg * 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
}

__device__ __forceinline__ void zero8x8(float (&acc)[8][8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
}

// Nine-plane gather of one 128-point tile into Xs[27][128]  (renderer.py:504-549; A.5 of SURVEY)
__device__ __forceinline__ void gather_tile(const RenderArgs &a, const float *z_s, float ox, float oy,
                                            float oz, float dx, float dy, float dz, float *Xs) {
    const int p = threadIdx.x & 127, half = threadIdx.x >> 7;
    const float z = z_s[p];
    // pts = o + d*z (separately rounded, as the reference's broadcasting arithmetic does)
    const float px = __fadd_rn(ox, __fmul_rn(dx, z));
    const float py = __fadd_rn(oy, __fmul_rn(dy, z));
    const float pz = __fadd_rn(oz, __fmul_rn(dz, z));
    const float cx = 2.f * (px - a.bmin[0]) / (a.bmax[0] - a.bmin[0]) - 1.f;
    const float cy = 2.f * (py - a.bmin[1]) / (a.bmax[1] - a.bmin[1]) - 1.f;
    const float cz = 2.f * (pz - a.bmin[2]) / (a.bmax[2] - a.bmin[2]) - 1.f;
    const int R = a.R;
    const float fR = (float)R, shift = 1.0f / fR;
    for (int c = half; c < 9; c += 2) {
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
        // clamp before the int cast so far-away points (miss rays) cannot overflow
        const int x0 = (int)fminf(fmaxf(fx0, -2.f), (float)R + 1.f);
        const int y0 = (int)fminf(fmaxf(fy0, -2.f), (float)R + 1.f);
        const float4 *tp = a.tex + (size_t)c * R * R;
        float r0 = 0.f, r1 = 0.f, r2 = 0.f;
        const bool xin0 = x0 >= 0 && x0 < R, xin1 = x0 + 1 >= 0 && x0 + 1 < R;
        const bool yin0 = y0 >= 0 && y0 < R, yin1 = y0 + 1 >= 0 && y0 + 1 < R;
        if (yin0 && xin0) { float4 t = __ldg(tp + (size_t)y0 * R + x0);           float w = wx0 * wy0; r0 = fmaf(t.x, w, r0); r1 = fmaf(t.y, w, r1); r2 = fmaf(t.z, w, r2); }
        if (yin0 && xin1) { float4 t = __ldg(tp + (size_t)y0 * R + x0 + 1);       float w = wx1 * wy0; r0 = fmaf(t.x, w, r0); r1 = fmaf(t.y, w, r1); r2 = fmaf(t.z, w, r2); }
        if (yin1 && xin0) { float4 t = __ldg(tp + (size_t)(y0 + 1) * R + x0);     float w = wx0 * wy1; r0 = fmaf(t.x, w, r0); r1 = fmaf(t.y, w, r1); r2 = fmaf(t.z, w, r2); }
        if (yin1 && xin1) { float4 t = __ldg(tp + (size_t)(y0 + 1) * R + x0 + 1); float w = wx1 * wy1; r0 = fmaf(t.x, w, r0); r1 = fmaf(t.y, w, r1); r2 = fmaf(t.z, w, r2); }
        Xs[(c * 3 + 0) * LDP + p] = r0;
        Xs[(c * 3 + 1) * LDP + p] = r1;
        Xs[(c * 3 + 2) * LDP + p] = r2;
    }
}

// pts_linears 0..2 (renderer.py:144-151): Xs -> Ha (h2);  uses Hb as scratch.
__device__ __forceinline__ void trunk(const float *__restrict__ mlp, const float *Xs, float *Ha,
                                      float *Hb, int pg, int og) {
    float acc[8][8];
    zero8x8(acc);
    mac8x8(acc, mlp + HL_MLP_W0, Xs, 27, pg, og);
    store8x8<true>(acc, mlp + HL_MLP_B0, Ha, pg, og);
    __syncthreads();
    zero8x8(acc);
    mac8x8(acc, mlp + HL_MLP_W1, Ha, 128, pg, og);
    store8x8<true>(acc, mlp + HL_MLP_B1, Hb, pg, og);
    __syncthreads();
    zero8x8(acc);
    mac8x8(acc, mlp + HL_MLP_W2, Xs, 27, pg, og);                 // skip: h = cat([x, h1])
    mac8x8(acc, mlp + HL_MLP_W2 + 27 * 128, Hb, 128, pg, og);
    store8x8<true>(acc, mlp + HL_MLP_B2, Ha, pg, og);
    __syncthreads();
}

// alpha_linear: sig_s[p] = wa . h2[:, p] + ba     (256 threads: 2 partial sums per point)
__device__ __forceinline__ void alpha_head(const float *__restrict__ mlp, const float *Ha, float *part_s,
                                           float *sig_out) {
    const int p = threadIdx.x & 127, half = threadIdx.x >> 7;
    float s = 0.f;
#pragma unroll 8
    for (int k = half * 64; k < half * 64 + 64; ++k) s = fmaf(__ldg(mlp + HL_MLP_WA + k), Ha[k * LDP + p], s);
    if (half) part_s[p] = s;
    __syncthreads();
    if (!half) sig_out[p] = s + part_s[p] + __ldg(mlp + HL_MLP_BA);
    __syncthreads();
}

__device__ __forceinline__ float block_sum(float v, float *red_s) {
    v = hl_warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) red_s[w] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < NT / 32; ++i) t += red_s[i];
    return t;
}

__global__ void __launch_bounds__(NT, 1) k_render(const RenderArgs a) {
    extern __shared__ __align__(16) float sm[];
    float *Xs = sm;                       // [28][128]
    float *Ha = Xs + 28 * LDP;            // [128][128]
    float *Hb = Ha + 128 * LDP;           // [128][128]
    float *zc = Hb + 128 * LDP;           // [128] coarse z
    float *zn = zc + NS;                  // [128] new z
    float *zf = zn + NS;                  // [256] merged z
    float *sig = zf + 2 * NS;             // [256] raw density
    float *wts = sig + 2 * NS;            // [256] alpha, then weights
    float *cdf = wts + 2 * NS;            // [128]
    float *bins = cdf + NS;               // [128]
    float *rgbs = bins + NS;              // [3][256]
    float *peb = rgbs + 3 * 2 * NS;       // [64] views bias incl. positional-encoding part
    float *part = peb + 64;               // [128] scratch
    float *red = part + NS;               // [8]
    float *pe = red + 8;                  // [28]

    const int tid = threadIdx.x;
    const int pg = tid & 15, og = tid >> 4;
    const float *mlp = a.mlp;

    for (long long ray = blockIdx.x; ray < a.n_rays; ray += gridDim.x) {
        const float ox = a.o[ray * 3 + 0], oy = a.o[ray * 3 + 1], oz = a.o[ray * 3 + 2];
        const float dx = a.d[ray * 3 + 0], dy = a.d[ray * 3 + 1], dz = a.d[ray * 3 + 2];
        const float nr = a.near[ray], fr = a.far[ray];
        const float dnorm = sqrtf(dx * dx + dy * dy + dz * dz);

        __syncthreads();   // previous ray fully consumed
        if (tid < NS) {
            // t = linspace(0,1,128) (ATen: start + i*step below the midpoint, end - (n-1-i)*step above)
            const float step = 1.0f / 127.0f;
            const float t = tid < 64 ? step * (float)tid : 1.0f - step * (float)(127 - tid);
            zc[tid] = a.zc_in ? a.zc_in[ray * NS + tid]
                              : __fadd_rn(__fmul_rn(nr, 1.0f - t), __fmul_rn(fr, t));
        } else if (tid < NS + 27) {
            // positional encoding of the unit view direction (fields.py:69-85)
            const int k = tid - NS;
            const float dd[3] = {dx / dnorm, dy / dnorm, dz / dnorm};
            float v;
            if (k < 3) {
                v = dd[k];
            } else {
                const int f = (k - 3) / 3, comp = (k - 3) % 3;
                const float freq = (float)(1 << (f >> 1));
                const float phase = (f & 1) ? 1.5707963267948966f : 0.f;
                v = sinf(__fadd_rn(phase, __fmul_rn(dd[comp], freq)));
            }
            pe[k] = v;
        }
        __syncthreads();
        if (tid < 64) {
            float s = __ldg(mlp + HL_MLP_BV + tid);
#pragma unroll
            for (int k = 0; k < 27; ++k) s = fmaf(__ldg(mlp + HL_MLP_WV + (128 + k) * 64 + tid), pe[k], s);
            peb[tid] = s;
        }

        // ------------------------------- coarse pass (density only) -------------------------------
        gather_tile(a, zc, ox, oy, oz, dx, dy, dz, Xs);
        __syncthreads();
        trunk(mlp, Xs, Ha, Hb, pg, og);
        alpha_head(mlp, Ha, part, sig);

        // ------------------------------- up_sample + sample_pdf -----------------------------------
        if (tid < NS) {
            const float dist = (tid < NS - 1 ? zc[tid + 1] - zc[tid] : 1e10f) * dnorm;
            wts[tid] = 1.0f - expf(-softplus_acc(sig[tid]) * dist);
            if (tid < NS - 1) bins[tid] = 0.5f * (zc[tid + 1] + zc[tid]);
        }
        __syncthreads();
        if (tid == 0) {
            float T = 1.0f;
            for (int i = 0; i < NS; ++i) {   // weights = alpha * cumprod([1, 1-alpha+1e-10])[:-1]
                const float al = wts[i];
                wts[i] = al * T;
                T *= (1.0f - al) + 1e-10f;
            }
        }
        __syncthreads();
        {
            const float wv = (tid >= 1 && tid <= NS - 2) ? wts[tid] + 1e-5f : 0.f;   // weights[..., 1:-1] + 1e-5
            const float tot = block_sum(wv, red);
            if (tid >= 1 && tid <= NS - 2) part[tid] = wv / tot;                     // pdf, 126 entries
        }
        __syncthreads();
        if (tid == 0) {
            float c = 0.f;
            cdf[0] = 0.f;
            for (int i = 1; i <= NS - 2; ++i) { c += part[i]; cdf[i] = c; }         // 127 entries
        }
        __syncthreads();
        if (tid < NS) {
            const float uu = a.u ? a.u[ray * NS + tid] : uniform_hash(a.seed, (unsigned long long)ray, tid);
            // searchsorted(cdf[0..126], u, right=True): first index with cdf[idx] > u
            int lo = 0, hi = NS - 1;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (cdf[mid] > uu) hi = mid; else lo = mid + 1;
            }
            const int below = max(lo - 1, 0), above = min(NS - 2, lo);
            float den = cdf[above] - cdf[below];
            if (den < 1e-5f) den = 1.0f;
            const float t = (uu - cdf[below]) / den;
            zn[tid] = bins[below] + t * (bins[above] - bins[below]);
        }
        __syncthreads();
        // sort(cat(z, z_new)) by ranking: coarse z is already sorted
        {
            const float v = tid < NS ? zc[tid] : zn[tid - NS];
            int rank;
            if (tid < NS) {
                int cnt = 0;
                for (int j = 0; j < NS; ++j) cnt += (zn[j] < v);
                rank = tid + cnt;
            } else {
                int cnt = 0;
                const int me = tid - NS;
                for (int j = 0; j < NS; ++j) {
                    const float w = zn[j];
                    cnt += (w < v) || (w == v && j < me);
                }
                int lo = 0, hi = NS;       // number of coarse z <= v
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (zc[mid] <= v) lo = mid + 1; else hi = mid;
                }
                rank = cnt + lo;
            }
            zf[rank] = v;
        }
        __syncthreads();

        // ------------------------------- fine pass: 2 tiles of 128 --------------------------------
        for (int t = 0; t < 2; ++t) {
            gather_tile(a, zf + t * NS, ox, oy, oz, dx, dy, dz, Xs);
            __syncthreads();
            trunk(mlp, Xs, Ha, Hb, pg, og);
            alpha_head(mlp, Ha, part, sig + t * NS);
            {   // feature_linear (no activation): Hb = Wf . h2 + bf
                float acc[8][8];
                zero8x8(acc);
                mac8x8(acc, mlp + HL_MLP_WF, Ha, 128, pg, og);
                store8x8<false>(acc, mlp + HL_MLP_BF, Hb, pg, og);
            }
            __syncthreads();
            {   // views_linear + softplus: Ha[0..63] = sp(Wv[:, :128] . feature + peb)
                const int pg4 = tid & 31, og8 = tid >> 5;    // 4 points x 8 outputs per thread
                float acc[4][8];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
#pragma unroll 4
                for (int k = 0; k < 128; ++k) {
                    const float4 a0 = *reinterpret_cast<const float4 *>(Hb + k * LDP + pg4 * 4);
                    const float4 w0 = __ldg(reinterpret_cast<const float4 *>(mlp + HL_MLP_WV + k * 64 + og8 * 8));
                    const float4 w1 = __ldg(reinterpret_cast<const float4 *>(mlp + HL_MLP_WV + k * 64 + og8 * 8 + 4));
                    const float av[4] = {a0.x, a0.y, a0.z, a0.w};
                    const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], w[j], acc[i][j]);
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float b = peb[og8 * 8 + j];
                    float4 v;
                    v.x = softplus_fast(acc[0][j] + b); v.y = softplus_fast(acc[1][j] + b);
                    v.z = softplus_fast(acc[2][j] + b); v.w = softplus_fast(acc[3][j] + b);
                    *reinterpret_cast<float4 *>(Ha + (og8 * 8 + j) * LDP + pg4 * 4) = v;
                }
            }
            __syncthreads();
            // rgb_linear + sigmoid: 3 x 128 outputs, K = 64
            for (int idx = tid; idx < 3 * NS; idx += NT) {
                const int c = idx >> 7, p = idx & 127;
                float s = __ldg(mlp + HL_MLP_BR + c);
#pragma unroll 8
                for (int k = 0; k < 64; ++k) s = fmaf(__ldg(mlp + HL_MLP_WR + k * 4 + c), Ha[k * LDP + p], s);
                rgbs[c * 2 * NS + t * NS + p] = 1.0f / (1.0f + expf(-s));
            }
            __syncthreads();
        }

        // ------------------------------- composite (renderer.py:222-239) --------------------------
        {
            const float dist = tid < 2 * NS - 1 ? zf[tid + 1] - zf[tid] : 1e10f;   // NOT scaled by |d|
            wts[tid] = 1.0f - expf(-softplus_acc(sig[tid]) * dist);
        }
        __syncthreads();
        if (tid == 0) {
            float T = 1.0f;
            for (int i = 0; i < 2 * NS; ++i) {
                const float al = wts[i];
                wts[i] = al * T;
                T *= (1.0f - al) + 1e-7f;
            }
        }
        __syncthreads();
        {
            const float w = wts[tid];
            const float s_acc = block_sum(w, red);
            const float s_r = block_sum(w * rgbs[0 * 2 * NS + tid], red);
            const float s_g = block_sum(w * rgbs[1 * 2 * NS + tid], red);
            const float s_b = block_sum(w * rgbs[2 * 2 * NS + tid], red);
            const float s_d = block_sum(w * zf[tid], red);
            if (tid == 0) {
                a.rgb[ray * 3 + 0] = s_r;
                a.rgb[ray * 3 + 1] = s_g;
                a.rgb[ray * 3 + 2] = s_b;
                a.acc[ray] = s_acc;
                float dep = (s_d - nr) / (fr - nr + 1e-5f);
                if (a.clamp_depth) dep = fminf(fmaxf(dep, 0.f), 1.f);
                a.depth[ray] = dep;
            }
        }
    }
}

__global__ void k_triplane_to_texels(const float *__restrict__ planes, float4 *__restrict__ tex, int R) {
    // planes: [3][9][R][R], channel = sub*3 + c ; tex: [3*3][R][R] float4
    const size_t n = (size_t)9 * R * R;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t pix = i % ((size_t)R * R);
        const int c = (int)(i / ((size_t)R * R));   // plane*3 + sub
        const int plane = c / 3, sub = c % 3;
        const float *src = planes + ((size_t)plane * 9 + sub * 3) * R * R + pix;
        tex[i] = make_float4(src[0], src[(size_t)R * R], src[2 * (size_t)R * R], 0.f);
    }
}

}  // namespace

extern "C" int hl_triplane_to_texels(const float *planes, float *texels, int R, void *stream) {
    HL_CHECK_ARG(planes && texels && R > 0 && ((uintptr_t)texels & 15) == 0);
    size_t n = (size_t)9 * R * R;
    int grid = (int)((n + 255) / 256);
    int cap = hl_num_sms() * 8;
    if (grid > cap) grid = cap;
    k_triplane_to_texels<<<grid, 256, 0, (cudaStream_t)stream>>>(planes, reinterpret_cast<float4 *>(texels), R);
    HL_CHECK_LAUNCH();
    return HL_OK;
}

extern "C" int hl_render_rays(const float *texels, int R, const float *mlp_packed, const float *rays_o,
                              const float *rays_d, const float *near, const float *far,
                              const float *z_coarse, const float *u, uint64_t seed, const float *bounds,
                              float *rgb, float *acc, float *depth,
                              int64_t n_rays, int clamp_depth, void *stream) {
    HL_CHECK_ARG(texels && mlp_packed && rays_o && rays_d && near && far && bounds && rgb && acc && depth);
    HL_CHECK_ARG(R > 0 && n_rays > 0 && ((uintptr_t)texels & 15) == 0 && ((uintptr_t)mlp_packed & 15) == 0);
    RenderArgs a;
    a.tex = reinterpret_cast<const float4 *>(texels);
    a.R = R;
    a.mlp = mlp_packed;
    a.o = rays_o; a.d = rays_d; a.near = near; a.far = far; a.u = u; a.zc_in = z_coarse;
    a.seed = seed;
    for (int i = 0; i < 3; ++i) { a.bmin[i] = bounds[i]; a.bmax[i] = bounds[3 + i]; }
    a.rgb = rgb; a.acc = acc; a.depth = depth;
    a.n_rays = n_rays;
    a.clamp_depth = clamp_depth;
    const size_t smem = sizeof(float) * (size_t)(28 * LDP + 2 * 128 * LDP + NS * 2 + 2 * NS * 3 + NS * 2 +
                                                 3 * 2 * NS + 64 + NS + 8 + 28 + 4);
    static bool configured = false;
    if (!configured) {
        HL_CHECK_CUDA(cudaFuncSetAttribute(k_render, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    int64_t grid = hl_num_sms();
    if (grid > n_rays) grid = n_rays;
    k_render<<<(int)grid, NT, smem, (cudaStream_t)stream>>>(a);
    HL_CHECK_LAUNCH();
    return HL_OK;
}
