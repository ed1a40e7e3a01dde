// Sampling-loop kernels that keep p_sample_loop free of library (ATen / cuRAND) launches:
//   hl_randn            x_T ~ N(0, I)                        (gaussian_diffusion.py:460  th.randn(*shape))
//   hl_ddpm_step_rng    the posterior update of elementwise.cu::k_ddpm_step with the per-step Gaussian drawn
//                       inside the kernel                    (gaussian_diffusion.py:383-387  th.randn_like(x))
//   hl_ddpm_posterior   mean / sample from a caller-supplied x0 (the denoised_fn route, :294-295,312-314)
//   hl_loop_advance     t <- t - 1, model timestep <- map[t] (respace.py:117-122), RNG counter += 1 -- on the
//                       device, so that one CUDA graph (UNet + posterior + advance) is replayed per step with no
//                       host-side tensor work in between.
// Generator: Philox4x32-10 keyed by a 64-bit seed, counter = (element index / 4, draw counter); Box-Muller on
// pairs.  The stream differs from torch.randn's (the reference draws on whatever device / generator PyTorch picks;
// parity tests inject the noise instead).  HBM traffic of the fused step: 16 B / element (x, eps in; sample,
// optional x0 out) instead of 20 + the 8 B / element a separate randn_like costs.
#include "common.cuh"

int hl_num_sms();

namespace {

__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    c[0] = hi1 ^ c[1] ^ k0;
    c[1] = lo1;
    c[2] = hi0 ^ c[3] ^ k1;
    c[3] = lo0;
}

// four N(0, 1) variates for elements 4*idx .. 4*idx+3 of draw number `draw`
__device__ __forceinline__ float4 randn4(uint64_t seed, uint64_t draw, uint64_t idx) {
    uint32_t c[4] = {(uint32_t)idx, (uint32_t)(idx >> 32), (uint32_t)draw, (uint32_t)(draw >> 32)};
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        philox_round(c, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    // (0, 1] uniforms from the top 24 bits; Box-Muller
    const float s = 1.0f / 16777216.0f;
    const float u0 = ((float)(c[0] >> 8) + 1.0f) * s, u1 = (float)(c[1] >> 8) * s;
    const float u2 = ((float)(c[2] >> 8) + 1.0f) * s, u3 = (float)(c[3] >> 8) * s;
    const float r0 = sqrtf(-2.0f * logf(u0)), r1 = sqrtf(-2.0f * logf(u2));
    float s0, c0, s1, c1;
    sincospif(2.0f * u1, &s0, &c0);
    sincospif(2.0f * u3, &s1, &c1);
    return make_float4(r0 * c0, r0 * s0, r1 * c1, r1 * s1);
}

__global__ void k_randn(float *__restrict__ out, int64_t n4, const uint64_t *__restrict__ rng, uint64_t seed,
                        uint64_t draw, int64_t off4) {
    hl_pdl_enter();
    if (rng) { seed = rng[0]; draw = rng[1]; }
    float4 *o4 = reinterpret_cast<float4 *>(out);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x)
        o4[i] = randn4(seed, draw, (uint64_t)(off4 + i));
}

// MODE 0: x0 from eps (k_ddpm_step's arithmetic, noise drawn here)   MODE 1: x0 supplied (posterior only)
template <int MODE>
__global__ void k_ddpm_step_rng(const float *__restrict__ x, const float *__restrict__ eps_or_x0,
                                const float *__restrict__ noise, const float *__restrict__ coef,
                                const float *__restrict__ sigma, const int64_t *__restrict__ t, int T,
                                float *__restrict__ sample, float *__restrict__ x0out, int64_t n4, int clip,
                                const uint64_t *__restrict__ rng, uint64_t seed, uint64_t draw, int64_t sample_offset) {
    hl_pdl_enter();
    const int b = blockIdx.y;
    const int64_t ti = t[b];
    const bool bad = ti < 0 || ti >= T;          // an out-of-range timestep poisons the row instead of reading out of bounds
    const int64_t tc = bad ? 0 : ti;
    const float c0 = coef[tc * 4 + 0], c1 = coef[tc * 4 + 1], c2 = coef[tc * 4 + 2], c3 = coef[tc * 4 + 3];
    const float sg = bad ? __int_as_float(0x7fc00000) : sigma[tc];
    if (rng) { seed = rng[0]; draw = rng[1]; }
    const float4 *x4 = reinterpret_cast<const float4 *>(x) + (int64_t)b * n4;
    const float4 *e4 = reinterpret_cast<const float4 *>(eps_or_x0) + (int64_t)b * n4;
    const float4 *z4 = noise ? reinterpret_cast<const float4 *>(noise) + (int64_t)b * n4 : nullptr;
    float4 *s4 = reinterpret_cast<float4 *>(sample) + (int64_t)b * n4;
    float4 *p4 = x0out ? reinterpret_cast<float4 *>(x0out) + (int64_t)b * n4 : nullptr;
    auto one = [&](float xv, float ev, float zv, float &x0, float &s) {
        // reference order: c0*x - c1*eps ; clamp ; c2*x0 + c3*x ; + sigma*noise (separate roundings)
        x0 = MODE == 0 ? __fsub_rn(__fmul_rn(c0, xv), __fmul_rn(c1, ev)) : ev;
        if (clip) x0 = fminf(fmaxf(x0, -1.f), 1.f);
        s = __fadd_rn(__fadd_rn(__fmul_rn(c2, x0), __fmul_rn(c3, xv)), __fmul_rn(sg, zv));
    };
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 xv = x4[i], ev = e4[i];
        const float4 zv = z4 ? z4[i] : randn4(seed, draw, (uint64_t)((sample_offset + b) * n4 + i));
        float4 x0, s;
        one(xv.x, ev.x, zv.x, x0.x, s.x);
        one(xv.y, ev.y, zv.y, x0.y, s.y);
        one(xv.z, ev.z, zv.z, x0.z, s.z);
        one(xv.w, ev.w, zv.w, x0.w, s.w);
        s4[i] = s;
        if (p4) p4[i] = x0;
    }
}

__global__ void k_loop_advance(int64_t *__restrict__ t, float *__restrict__ t_model, const int64_t *__restrict__ map,
                               float scale, int B, uint64_t *__restrict__ rng) {
    hl_pdl_enter();
    const int b = threadIdx.x;
    if (b < B) {
        const int64_t ti = t[b] - 1;
        t[b] = ti;
        if (ti >= 0) t_model[b] = map ? (float)map[ti] * scale : (float)ti * scale;
    }
    if (b == 0 && rng) rng[1] += 1;
}

int grid_for(int64_t n4, int B) {
    int gx = (int)((n4 + 255) / 256);
    int cap = hl_num_sms() * 8 / (B > 0 ? B : 1);
    if (cap < 1) cap = 1;
    return gx > cap ? cap : gx;
}

}  // namespace

extern "C" int hl_randn(float *out, int64_t n, const uint64_t *rng_state, uint64_t seed, uint64_t draw,
                        int64_t element_offset, void *stream) {
    HL_CHECK_ARG(out && n > 0 && n % 4 == 0 && element_offset % 4 == 0 && ((uintptr_t)out & 15) == 0);
    HL_CHECK_CUDA(hl_launch(k_randn, dim3(grid_for(n / 4, 1)), dim3(256), 0, (cudaStream_t)stream, out, n / 4, rng_state,
                            seed, draw, element_offset / 4));
    HL_CHECK_LAUNCH();
    return HL_OK;
}

extern "C" int hl_ddpm_step_rng(const float *x, const float *eps, const float *noise, const float *coef,
                                const float *sigma, const int64_t *t, int T, float *sample, float *pred_xstart, int B,
                                int64_t n, int clip, const uint64_t *rng_state, uint64_t seed, uint64_t draw,
                                int64_t sample_offset, void *stream) {
    HL_CHECK_ARG(x && eps && coef && sigma && t && sample && B > 0 && T > 0 && n > 0 && n % 4 == 0);
    HL_CHECK_CUDA(hl_launch(k_ddpm_step_rng<0>, dim3(grid_for(n / 4, B), B), dim3(256), 0, (cudaStream_t)stream, x, eps, noise,
                            coef, sigma, t, T, sample, pred_xstart, n / 4, clip, rng_state, seed, draw, sample_offset));
    HL_CHECK_LAUNCH();
    return HL_OK;
}

extern "C" int hl_ddpm_posterior(const float *x, const float *x0, const float *noise, const float *coef,
                                 const float *sigma, const int64_t *t, int T, float *sample, float *x0_clipped, int B,
                                 int64_t n, int clip, const uint64_t *rng_state, uint64_t seed, uint64_t draw,
                                 int64_t sample_offset, void *stream) {
    HL_CHECK_ARG(x && x0 && coef && sigma && t && sample && B > 0 && T > 0 && n > 0 && n % 4 == 0);
    HL_CHECK_CUDA(hl_launch(k_ddpm_step_rng<1>, dim3(grid_for(n / 4, B), B), dim3(256), 0, (cudaStream_t)stream, x, x0, noise,
                            coef, sigma, t, T, sample, x0_clipped, n / 4, clip, rng_state, seed, draw, sample_offset));
    HL_CHECK_LAUNCH();
    return HL_OK;
}

extern "C" int hl_loop_advance(int64_t *t, float *t_model, const int64_t *timestep_map, float scale, int B,
                               uint64_t *rng_state, void *stream) {
    HL_CHECK_ARG(t && t_model && B > 0 && B <= 1024);
    HL_CHECK_CUDA(hl_launch(k_loop_advance, dim3(1), dim3((B + 31) / 32 * 32), 0, (cudaStream_t)stream, t, t_model,
                            timestep_map, scale, B, rng_state));
    HL_CHECK_LAUNCH();
    return HL_OK;
}
