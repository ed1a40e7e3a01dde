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
//
// Canonical-space variant (use_canonical_space=True, the TightCap branch of triplane_sample_layered.py:73-76;
// human_diffusion/NeRF/renderer.py:52-133 deform_target2c / deform_target2c_op): every sample point is taken to the
// posed body's SMPL frame, snapped to its nearest body vertex (pytorch3d knn_points, K = 1) and carried to the canonical
// "big pose" by that vertex's skinning transforms; the view direction follows the same rotations, so the positional
// encoding of views_linear becomes per sample instead of per ray.  Everything the reference does per *point* after the
// nearest-vertex lookup depends on the vertex only (blend weights, blend-shape offsets and both joint-transform blends
// are indexed by vert_ids), so k_smpl_vertex_tables folds the whole chain -- inverse skinning, minus pose and shape
// offsets, plus big-pose offsets, forward skinning -- into one 3x4 affine per vertex (fp64 arithmetic, stored fp32),
// once per frame.  Per point the render kernel then does: nearest vertex (exact; cluster bounding spheres prune the scan,
// canon.cuh; the two threads of a point split the clusters) and one affine.
#include "common.cuh"
#include "canon.cuh"

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
    CanonTables ct;         // canonical space (k_render<true> / k_canon_points / k_density_grid_canon only)
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

// deform_target2c for the tile's 128 points (two threads per point, each searching half of the clusters): world point ->
// canonical point; optionally the canonical view direction of the sample into vdc[3][128].  `nn_s` = [2][256] (distance,
// index) exchange.  Ends with every thread holding its point's canonical position.
__device__ __forceinline__ void canon_tile(const RenderArgs &a, float *nn_s, float &px, float &py, float &pz,
                                           const float *sv, float *vdc) {
    float4 *sm_f4 = reinterpret_cast<float4 *>(nn_s + 512);      // search scratch follows the exchange area
    const int p = threadIdx.x & 127, half = threadIdx.x >> 7;
    float qx, qy, qz;
    hl_to_smpl_frame(a.ct, px, py, pz, qx, qy, qz);
    const int mid = a.ct.NC >> 1;
    float best;
    int bi;
    hl_nearest_vertex(a.ct, sm_f4, qx, qy, qz, half ? mid : 0, half ? a.ct.NC : mid, best, bi);
    nn_s[half * 256 + p] = best;
    nn_s[half * 256 + 128 + p] = __int_as_float(bi);
    __syncthreads();
    const float d0 = nn_s[p], d1 = nn_s[256 + p];
    const int i0 = __float_as_int(nn_s[128 + p]), i1 = __float_as_int(nn_s[256 + 128 + p]);
    const int v = (d1 < d0 || (d1 == d0 && i1 < i0)) ? i1 : i0;
    float dc[3];
    hl_canon_affine(a.ct, v, qx, qy, qz, px, py, pz, sv, vdc ? dc : nullptr);
    if (vdc && half == 0) {
        vdc[p] = dc[0];
        vdc[LDP + p] = dc[1];
        vdc[2 * LDP + p] = dc[2];
    }
    __syncthreads();          // nn_s may be overwritten from here on
}

// Nine-plane gather of one 128-point tile into Xs[27][128]  (renderer.py:504-549; A.5 of SURVEY); thread (p, half) holds
// point p = (px, py, pz)
__device__ __forceinline__ void gather_point(const RenderArgs &a, float px, float py, float pz, float *Xs) {
    const int p = threadIdx.x & 127, half = threadIdx.x >> 7;
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

// one 128-sample tile of a ray: pts = o + d*z (separately rounded, as the reference's broadcasting arithmetic does)
template <bool CANON>
__device__ __forceinline__ void gather_tile(const RenderArgs &a, const float *z_s, float ox, float oy,
                                            float oz, float dx, float dy, float dz, float *Xs,
                                            float *nn_s = nullptr, const float *sv = nullptr, float *vdc = nullptr) {
    const float z = z_s[threadIdx.x & 127];
    float px = __fadd_rn(ox, __fmul_rn(dx, z));
    float py = __fadd_rn(oy, __fmul_rn(dy, z));
    float pz = __fadd_rn(oz, __fmul_rn(dz, z));
    if (CANON) canon_tile(a, nn_s, px, py, pz, sv, vdc);
    gather_point(a, px, py, pz, Xs);
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

template <bool CANON>
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
    float *nn = pe + 28 + 4;              // CANON: [2][256] nearest-vertex exchange | search scratch (canon.cuh)
    float *vdc = nn + 512 + 4 * HL_CANON_SMEM_F4;   // CANON: [3][128] canonical view direction of the tile's samples
    if (CANON) {
        hl_canon_stage_spheres(a.ct, reinterpret_cast<float4 *>(nn + 512));
        __syncthreads();
    }

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
            if (!CANON) {          // per-ray view direction: its encoding folds into the bias of views_linear
#pragma unroll
                for (int k = 0; k < 27; ++k) s = fmaf(__ldg(mlp + HL_MLP_WV + (128 + k) * 64 + tid), pe[k], s);
            }
            peb[tid] = s;
        }
        // CANON: smpl_viewdir = (viewdir - Th) R  (renderer.py:125 subtracts Th from the direction too)
        float sv[3] = {0.f, 0.f, 0.f};
        if (CANON) {
            hl_to_smpl_frame(a.ct, dx / dnorm, dy / dnorm, dz / dnorm, sv[0], sv[1], sv[2]);
        }

        // ------------------------------- coarse pass (density only) -------------------------------
        gather_tile<CANON>(a, zc, ox, oy, oz, dx, dy, dz, Xs, nn);
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
            gather_tile<CANON>(a, zf + t * NS, ox, oy, oz, dx, dy, dz, Xs, nn, sv, vdc);
            __syncthreads();
            trunk(mlp, Xs, Ha, Hb, pg, og);
            alpha_head(mlp, Ha, part, sig + t * NS);
            if (CANON) {   // Xs is free after the trunk: rows 0..26 <- positional encoding of the canonical direction
                const int p = tid & 127;
                for (int k = tid >> 7; k < 27; k += 2) {
                    float v;
                    if (k < 3) {
                        v = vdc[k * LDP + p];
                    } else {
                        const int f = (k - 3) / 3, comp = (k - 3) % 3;
                        const float freq = (float)(1 << (f >> 1));
                        const float phase = (f & 1) ? 1.5707963267948966f : 0.f;
                        v = sinf(__fadd_rn(phase, __fmul_rn(vdc[comp * LDP + p], freq)));
                    }
                    Xs[k * LDP + p] = v;
                }
            }
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
                if (CANON) {       // views_linear columns 128..154: the per-sample view encoding
                    for (int k = 0; k < 27; ++k) {
                        const float4 a0 = *reinterpret_cast<const float4 *>(Xs + k * LDP + pg4 * 4);
                        const float4 w0 = __ldg(reinterpret_cast<const float4 *>(mlp + HL_MLP_WV + (128 + k) * 64 + og8 * 8));
                        const float4 w1 = __ldg(reinterpret_cast<const float4 *>(mlp + HL_MLP_WV + (128 + k) * 64 + og8 * 8 + 4));
                        const float av[4] = {a0.x, a0.y, a0.z, a0.w};
                        const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                        for (int i = 0; i < 4; ++i)
#pragma unroll
                            for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], w[j], acc[i][j]);
                    }
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

// ------------------------------------------------------------------------------------------------------------------
// Per-frame vertex tables of the canonical-space deformation.  One warp per table slot (cluster order).
//   consts (device, fp64; J = joints, 24 for SMPL): A_pose [J][12] | A_big [J][12] (rows of the 3x4 joint transforms of
//   get_transform_params_torch for params and for t_params with zero shape) | pose_feature [9(J-1)] | pose_feature_big
//   [9(J-1)] (rot_mats[1:] - I, renderer.py:80-83,98-100) | betas [16] | R [9] | Th [3]
// Per vertex v (everything deform_target2c_op indexes by vert_ids):
//   A  = sum_j w[v][j] A_pose[j],  Ab = sum_j w[v][j] A_big[j]
//   po = posedirs[v] . pose_feature,  pob = posedirs[v] . pose_feature_big,  so = shapedirs[v] . betas
//   canonical(q) = Ab_R (A_R^-1 (q - A_t) - po - so + pob) + Ab_t  =  M q + c,   M = Ab_R A_R^-1
// and the vertex position in the SMPL frame, (vertices[v] - Th) R, for the nearest-vertex search.
__global__ void k_smpl_vertex_tables(const float *__restrict__ weights, const float *__restrict__ posedirs,
                                     const float *__restrict__ shapedirs, int S_asset, int S,
                                     const float *__restrict__ vertices, const double *__restrict__ cst, int V, int J,
                                     const int *__restrict__ slot_vertex, int n_slots, float4 *__restrict__ verts,
                                     float4 *__restrict__ aff) {
    const int PF = (J - 1) * 9;           // pose-feature length (207 for SMPL)
    const int SC_APOSE = 0, SC_ABIG = J * 12, SC_PF = 2 * J * 12, SC_PFB = SC_PF + PF, SC_BETAS = SC_PFB + PF,
              SC_R = SC_BETAS + 16, SC_TH = SC_R + 9;
    const int slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (slot >= n_slots) return;
    const int v = slot_vertex[slot];
    if (v < 0 || v >= V) {      // unused slot of its cluster: never the nearest
        if (lane == 0) hl_pair_put(verts, slot, 1e18f, 1e18f, 1e18f, __int_as_float(0x7fffffff));
        return;
    }
    double po[3] = {0, 0, 0}, pob[3] = {0, 0, 0};
    for (int k = lane; k < 3 * PF; k += 32) {
        const int c = k / PF, f = k - c * PF;
        const double pd = (double)posedirs[(size_t)v * 3 * PF + k];
        po[c] += pd * cst[SC_PF + f];
        pob[c] += pd * cst[SC_PFB + f];
    }
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            po[c] += __shfl_xor_sync(0xffffffffu, po[c], o);
            pob[c] += __shfl_xor_sync(0xffffffffu, pob[c], o);
        }
    if (lane) return;
    double A[12], Ab[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) A[i] = Ab[i] = 0.0;
    for (int j = 0; j < J; ++j) {
        const double w = (double)weights[(size_t)v * J + j];
        if (w == 0.0) continue;
#pragma unroll
        for (int i = 0; i < 12; ++i) {
            A[i] += w * cst[SC_APOSE + j * 12 + i];
            Ab[i] += w * cst[SC_ABIG + j * 12 + i];
        }
    }
    double so[3] = {0, 0, 0};
    for (int c = 0; c < 3; ++c)
        for (int b = 0; b < S; ++b) so[c] += (double)shapedirs[((size_t)v * 3 + c) * S_asset + b] * cst[SC_BETAS + b];
    // inverse of the 3x3 block of A (row-major, row r = A[r*4 .. r*4+2], translation A[r*4+3])
    const double a00 = A[0], a01 = A[1], a02 = A[2], a10 = A[4], a11 = A[5], a12 = A[6], a20 = A[8], a21 = A[9], a22 = A[10];
    const double c00 = a11 * a22 - a12 * a21, c01 = a12 * a20 - a10 * a22, c02 = a10 * a21 - a11 * a20;
    const double det = a00 * c00 + a01 * c01 + a02 * c02;
    const double id = 1.0 / det;
    const double inv[9] = {c00 * id, (a02 * a21 - a01 * a22) * id, (a01 * a12 - a02 * a11) * id,
                           c01 * id, (a00 * a22 - a02 * a20) * id, (a02 * a10 - a00 * a12) * id,
                           c02 * id, (a01 * a20 - a00 * a21) * id, (a00 * a11 - a01 * a10) * id};
    double off[3];           // -inv t - po - so + pob
    for (int r = 0; r < 3; ++r)
        off[r] = -(inv[r * 3] * A[3] + inv[r * 3 + 1] * A[7] + inv[r * 3 + 2] * A[11]) - po[r] - so[r] + pob[r];
    for (int r = 0; r < 3; ++r) {
        double m[3];
        for (int c = 0; c < 3; ++c)
            m[c] = Ab[r * 4] * inv[c] + Ab[r * 4 + 1] * inv[3 + c] + Ab[r * 4 + 2] * inv[6 + c];
        const double cc = Ab[r * 4] * off[0] + Ab[r * 4 + 1] * off[1] + Ab[r * 4 + 2] * off[2] + Ab[r * 4 + 3];
        aff[(size_t)v * 3 + r] = make_float4((float)m[0], (float)m[1], (float)m[2], (float)cc);
    }
    // smpl_pts = (vertices - Th) R in fp32, as the reference computes the search set (renderer.py:61)
    const float ex = vertices[(size_t)v * 3] - (float)cst[SC_TH], ey = vertices[(size_t)v * 3 + 1] - (float)cst[SC_TH + 1],
                ez = vertices[(size_t)v * 3 + 2] - (float)cst[SC_TH + 2];
    float q[3];
    for (int c = 0; c < 3; ++c)
        q[c] = fmaf(ez, (float)cst[SC_R + 6 + c], fmaf(ey, (float)cst[SC_R + 3 + c], ex * (float)cst[SC_R + c]));
    hl_pair_put(verts, slot, q[0], q[1], q[2], __int_as_float(v));      // pair layout (canon.cuh); CL is even, so a pair never straddles two clusters
}

// Bounding sphere of each cluster's posed vertices (one warp per cluster): centre = centroid, radius = the largest
// distance, rounded up so that the bounds of hl_nearest_vertex stay conservative.
__global__ void k_smpl_cluster_bounds(const float4 *__restrict__ verts, int NC, int CL, float4 *__restrict__ spheres) {
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (c >= NC) return;
    float sx = 0.f, sy = 0.f, sz = 0.f, n = 0.f;
    for (int k = lane; k < CL; k += 32) {
        const float4 p = hl_pair_get(verts + (size_t)c * CL, k);
        if (p.x < 1e17f) { sx += p.x; sy += p.y; sz += p.z; n += 1.f; }
    }
    sx = hl_warp_sum(sx); sy = hl_warp_sum(sy); sz = hl_warp_sum(sz); n = hl_warp_sum(n);
    const float cx = n > 0.f ? sx / n : 0.f, cy = n > 0.f ? sy / n : 0.f, cz = n > 0.f ? sz / n : 0.f;
    float r = 0.f;
    for (int k = lane; k < CL; k += 32) {
        const float4 p = hl_pair_get(verts + (size_t)c * CL, k);
        if (p.x < 1e17f) r = fmaxf(r, sqrtf(hl_dist2(cx, cy, cz, p.x, p.y, p.z)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r = fmaxf(r, __shfl_xor_sync(0xffffffffu, r, o));
    // an empty cluster gets a far-away centre: its lower bound never passes, its upper bound never wins
    if (lane == 0) {
        if (n > 0.f) hl_pair_put(spheres, c, cx, cy, cz, r * (1.0f + 4e-6f) + 1e-7f);
        else hl_pair_put(spheres, c, 1e18f, 1e18f, 1e18f, 0.f);
    }
}

// deform_target2c on arbitrary points (tests / Renderer.deform_target2c): 128 points per block, two threads per point
__global__ void __launch_bounds__(NT) k_canon_points(const RenderArgs a, const float *__restrict__ pts,
                                                     const float *__restrict__ dirs, long long n,
                                                     float *__restrict__ out_pts, float *__restrict__ out_dirs) {
    __shared__ __align__(16) float nn[512 + 4 * HL_CANON_SMEM_F4];
    __shared__ float vdc[3 * LDP];
    const int p = threadIdx.x & 127, half = threadIdx.x >> 7;
    hl_canon_stage_spheres(a.ct, reinterpret_cast<float4 *>(nn + 512));
    __syncthreads();
    for (long long base = (long long)blockIdx.x * 128; base < n; base += (long long)gridDim.x * 128) {
        const long long i = base + p < n ? base + p : n - 1;
        float px = pts[i * 3], py = pts[i * 3 + 1], pz = pts[i * 3 + 2];
        float sv[3] = {0.f, 0.f, 0.f};
        if (dirs) hl_to_smpl_frame(a.ct, dirs[i * 3], dirs[i * 3 + 1], dirs[i * 3 + 2], sv[0], sv[1], sv[2]);
        canon_tile(a, nn, px, py, pz, sv, dirs ? vdc : nullptr);
        if (half == 0 && base + p < n) {
            out_pts[i * 3] = px; out_pts[i * 3 + 1] = py; out_pts[i * 3 + 2] = pz;
            if (dirs) { out_dirs[i * 3] = vdc[p]; out_dirs[i * 3 + 1] = vdc[LDP + p]; out_dirs[i * 3 + 2] = vdc[2 * LDP + p]; }
        }
        __syncthreads();
    }
}

// Density on a regular grid of the posed-space box with every grid point deformed to the canonical space
// (extract_geometry with use_canonical_space=True, human_diffusion/NeRF/renderer.py:290-318: linspace^3 of
// tp_input['world_bounds'], deform_target2c, features inside t_world_bounds, -sigma); 128 points per tile.
__global__ void __launch_bounds__(NT, 1) k_density_grid_canon(const RenderArgs a, float3 wmin, float3 wmax, int res,
                                                              float *__restrict__ out) {
    extern __shared__ __align__(16) float sm[];
    float *Xs = sm;
    float *Ha = Xs + 28 * LDP;
    float *Hb = Ha + 128 * LDP;
    float *sig = Hb + 128 * LDP;          // [128]
    float *part = sig + NS;               // [128]
    float *nn = part + NS;                // [512] | search scratch
    const int tid = threadIdx.x, pg = tid & 15, og = tid >> 4;
    hl_canon_stage_spheres(a.ct, reinterpret_cast<float4 *>(nn + 512));
    __syncthreads();
    const long long total = (long long)res * res * res;
    const long long tiles = (total + 127) / 128;
    // torch.linspace(lo, hi, R): lo + i*step below the midpoint, hi - (R-1-i)*step above (ATen)
    auto lin = [&](float lo, float hi, int i) {
        const float step = (hi - lo) / (float)(res - 1);
        return i < res / 2 ? __fadd_rn(lo, __fmul_rn(step, (float)i)) : __fsub_rn(hi, __fmul_rn(step, (float)(res - 1 - i)));
    };
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        long long idx = tile * 128 + (tid & 127);
        if (idx >= total) idx = total - 1;
        const int zi = (int)(idx % res), yi = (int)((idx / res) % res), xi = (int)(idx / ((long long)res * res));
        float px = lin(wmin.x, wmax.x, xi), py = lin(wmin.y, wmax.y, yi), pz = lin(wmin.z, wmax.z, zi);
        __syncthreads();     // previous tile consumed
        canon_tile(a, nn, px, py, pz, nullptr, nullptr);
        gather_point(a, px, py, pz, Xs);
        __syncthreads();
        trunk(a.mlp, Xs, Ha, Hb, pg, og);
        alpha_head(a.mlp, Ha, part, sig);
        if (tid < 128 && tile * 128 + tid < total) out[tile * 128 + tid] = -sig[tid];
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


static size_t render_smem_bytes() {
    return sizeof(float) * (size_t)(28 * LDP + 2 * 128 * LDP + NS * 2 + 2 * NS * 3 + NS * 2 + 3 * 2 * NS + 64 + NS + 8 +
                                    28 + 4 + 512 + 4 * HL_CANON_SMEM_F4 + 3 * LDP);
}

static void fill_render_args(RenderArgs &a, const float *texels, int R, const float *mlp_packed, const float *rays_o,
                             const float *rays_d, const float *near, const float *far, const float *z_coarse,
                             const float *u, uint64_t seed, const float *bounds, float *rgb, float *acc, float *depth,
                             int64_t n_rays, int clamp_depth) {
    a = RenderArgs{};
    a.tex = reinterpret_cast<const float4 *>(texels);
    a.R = R;
    a.mlp = mlp_packed;
    a.o = rays_o; a.d = rays_d; a.near = near; a.far = far; a.u = u; a.zc_in = z_coarse;
    a.seed = seed;
    for (int i = 0; i < 3; ++i) { a.bmin[i] = bounds[i]; a.bmax[i] = bounds[3 + i]; }
    a.rgb = rgb; a.acc = acc; a.depth = depth;
    a.n_rays = n_rays;
    a.clamp_depth = clamp_depth;
}

extern "C" int hl_render_rays(const float *texels, int R, const float *mlp_packed, const float *rays_o,
                              const float *rays_d, const float *near, const float *far,
                              const float *z_coarse, const float *u, uint64_t seed, const float *bounds,
                              float *rgb, float *acc, float *depth,
                              int64_t n_rays, int clamp_depth, void *stream) {
    HL_CHECK_ARG(texels && mlp_packed && rays_o && rays_d && near && far && bounds && rgb && acc && depth);
    HL_CHECK_ARG(R > 0 && n_rays > 0 && ((uintptr_t)texels & 15) == 0 && ((uintptr_t)mlp_packed & 15) == 0);
    RenderArgs a;
    fill_render_args(a, texels, R, mlp_packed, rays_o, rays_d, near, far, z_coarse, u, seed, bounds, rgb, acc, depth,
                     n_rays, clamp_depth);
    const size_t smem = render_smem_bytes();
    static HlPerDeviceOnce once;
    if (once.need()) {
        HL_CHECK_CUDA(cudaFuncSetAttribute(k_render<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    int64_t grid = hl_num_sms();
    if (grid > n_rays) grid = n_rays;
    k_render<false><<<(int)grid, NT, smem, (cudaStream_t)stream>>>(a);
    HL_CHECK_LAUNCH();
    return HL_OK;
}

extern "C" int hl_smpl_vertex_tables(const float *weights, const float *posedirs, const float *shapedirs, int n_betas_asset,
                                     int n_betas, const float *vertices, const double *consts, int n_verts, int n_joints,
                                     const int *slot_vertex, int n_clusters, int cluster_slots, float *knn_table,
                                     float *affine_table, void *stream) {
    HL_CHECK_ARG(weights && posedirs && shapedirs && vertices && consts && slot_vertex && knn_table && affine_table);
    HL_CHECK_ARG(n_verts > 0 && n_betas >= 0 && n_betas <= 16 && n_betas <= n_betas_asset);
    HL_CHECK_ARG(n_joints >= 2 && n_joints <= 64 && n_clusters >= 2 && n_clusters <= 128 && cluster_slots >= 1);
    HL_CHECK_ARG((int64_t)n_clusters * cluster_slots >= n_verts && n_clusters % 2 == 0 && cluster_slots % 4 == 0);
    HL_CHECK_ARG(((uintptr_t)knn_table & 15) == 0 && ((uintptr_t)affine_table & 15) == 0);
    const int n_slots = n_clusters * cluster_slots;
    float4 *spheres = reinterpret_cast<float4 *>(knn_table);
    float4 *verts = spheres + n_clusters;
    k_smpl_vertex_tables<<<hl_cdiv((int64_t)n_slots * 32, 256), 256, 0, (cudaStream_t)stream>>>(
        weights, posedirs, shapedirs, n_betas_asset, n_betas, vertices, consts, n_verts, n_joints, slot_vertex, n_slots,
        verts, reinterpret_cast<float4 *>(affine_table));
    HL_CHECK_LAUNCH();
    k_smpl_cluster_bounds<<<hl_cdiv((int64_t)n_clusters * 32, 256), 256, 0, (cudaStream_t)stream>>>(verts, n_clusters,
                                                                                                 cluster_slots, spheres);
    HL_CHECK_LAUNCH();
    return HL_OK;
}

// knn_table = [spheres n_clusters | verts n_clusters * cluster_slots] float4 (hl_smpl_vertex_tables)
int hl_set_canon_tables(CanonTables &t, const float *knn_table, const float *affine_table, int n_clusters, int cluster_slots,
                        const float *rot, const float *trans) {
    HL_CHECK_ARG(knn_table && affine_table && rot && trans && n_clusters >= 2 && n_clusters <= HL_CANON_NC_MAX && cluster_slots >= 1 &&
                 cluster_slots <= HL_CANON_CL_MAX);
    HL_CHECK_ARG(((uintptr_t)knn_table & 15) == 0 && ((uintptr_t)affine_table & 15) == 0);
    t.spheres = reinterpret_cast<const float4 *>(knn_table);
    t.verts = t.spheres + n_clusters;
    t.aff = reinterpret_cast<const float4 *>(affine_table);
    t.NC = n_clusters;
    t.CL = cluster_slots;
    for (int i = 0; i < 9; ++i) t.Rm[i] = rot[i];
    for (int i = 0; i < 3; ++i) t.Th[i] = trans[i];
    return HL_OK;
}

extern "C" int hl_render_rays_canon(const float *texels, int R, const float *mlp_packed, const float *rays_o,
                                    const float *rays_d, const float *near, const float *far, const float *z_coarse,
                                    const float *u, uint64_t seed, const float *t_bounds, const float *knn_table,
                                    const float *affine_table, int n_clusters, int cluster_slots, const float *rot,
                                    const float *trans, float *rgb, float *acc, float *depth, int64_t n_rays,
                                    int clamp_depth, void *stream) {
    HL_CHECK_ARG(texels && mlp_packed && rays_o && rays_d && near && far && t_bounds && rgb && acc && depth);
    HL_CHECK_ARG(R > 0 && n_rays > 0 && ((uintptr_t)texels & 15) == 0 && ((uintptr_t)mlp_packed & 15) == 0);
    RenderArgs a;
    fill_render_args(a, texels, R, mlp_packed, rays_o, rays_d, near, far, z_coarse, u, seed, t_bounds, rgb, acc, depth,
                     n_rays, clamp_depth);
    if (int rc = hl_set_canon_tables(a.ct, knn_table, affine_table, n_clusters, cluster_slots, rot, trans)) return rc;
    const size_t smem = render_smem_bytes();
    static HlPerDeviceOnce once;
    if (once.need()) {
        HL_CHECK_CUDA(cudaFuncSetAttribute(k_render<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    int64_t grid = hl_num_sms();
    if (grid > n_rays) grid = n_rays;
    k_render<true><<<(int)grid, NT, smem, (cudaStream_t)stream>>>(a);
    HL_CHECK_LAUNCH();
    return HL_OK;
}

extern "C" int hl_canonical_points(const float *pts, const float *dirs, int64_t n, const float *knn_table,
                                   const float *affine_table, int n_clusters, int cluster_slots, const float *rot,
                                   const float *trans, float *out_pts, float *out_dirs, void *stream) {
    HL_CHECK_ARG(pts && out_pts && n > 0 && (dirs == nullptr) == (out_dirs == nullptr));
    RenderArgs a = {};
    if (int rc = hl_set_canon_tables(a.ct, knn_table, affine_table, n_clusters, cluster_slots, rot, trans)) return rc;
    int64_t grid = (n + 127) / 128;
    if (grid > 8 * hl_num_sms()) grid = 8 * hl_num_sms();
    k_canon_points<<<(int)grid, NT, 0, (cudaStream_t)stream>>>(a, pts, dirs, (long long)n, out_pts, out_dirs);
    HL_CHECK_LAUNCH();
    return HL_OK;
}

extern "C" int hl_density_grid_canon(const float *texels, int R, const float *mlp_packed, const float *world_bounds,
                                     const float *t_bounds, const float *knn_table, const float *affine_table,
                                     int n_clusters, int cluster_slots, const float *rot, const float *trans,
                                     int resolution, float *out, void *stream) {
    HL_CHECK_ARG(texels && mlp_packed && world_bounds && t_bounds && out && R > 0 && resolution >= 2);
    HL_CHECK_ARG(((uintptr_t)texels & 15) == 0 && ((uintptr_t)mlp_packed & 15) == 0);
    RenderArgs a = {};
    a.tex = reinterpret_cast<const float4 *>(texels);
    a.R = R;
    a.mlp = mlp_packed;
    for (int i = 0; i < 3; ++i) { a.bmin[i] = t_bounds[i]; a.bmax[i] = t_bounds[3 + i]; }
    if (int rc = hl_set_canon_tables(a.ct, knn_table, affine_table, n_clusters, cluster_slots, rot, trans)) return rc;
    const size_t smem = sizeof(float) * (size_t)(28 * LDP + 2 * 128 * LDP + 2 * NS + 512 + 4 * HL_CANON_SMEM_F4);
    static HlPerDeviceOnce once;
    if (once.need()) {
        HL_CHECK_CUDA(cudaFuncSetAttribute(k_density_grid_canon, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    const long long tiles = ((long long)resolution * resolution * resolution + 127) / 128;
    long long grid = hl_num_sms();
    if (grid > tiles) grid = tiles;
    k_density_grid_canon<<<(int)grid, NT, smem, (cudaStream_t)stream>>>(
        a, make_float3(world_bounds[0], world_bounds[1], world_bounds[2]),
        make_float3(world_bounds[3], world_bounds[4], world_bounds[5]), resolution, out);
    HL_CHECK_LAUNCH();
    return HL_OK;
}
