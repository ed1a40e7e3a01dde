// Shared helpers for libhumanliff_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/humanliff_b200.h"

void hl_set_error(const char *fmt, ...);

#define HL_CHECK_ARG(cond)                                                        \
    do {                                                                          \
        if (!(cond)) {                                                            \
            hl_set_error("%s:%d: invalid argument: %s", __FILE__, __LINE__, #cond); \
            return HL_E_INVALID;                                                  \
        }                                                                         \
    } while (0)

#define HL_CHECK_LAUNCH()                                                          \
    do {                                                                           \
        cudaError_t e__ = cudaGetLastError();                                      \
        if (e__ != cudaSuccess) {                                                  \
            hl_set_error("%s:%d: CUDA error: %s", __FILE__, __LINE__,              \
                         cudaGetErrorString(e__));                                 \
            return HL_E_CUDA;                                                      \
        }                                                                          \
    } while (0)

#define HL_CHECK_CUDA(call)                                                        \
    do {                                                                           \
        cudaError_t e__ = (call);                                                  \
        if (e__ != cudaSuccess) {                                                  \
            hl_set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call,        \
                         cudaGetErrorString(e__));                                 \
            return HL_E_CUDA;                                                      \
        }                                                                          \
    } while (0)

static inline int hl_cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// cvt.rna.tf32.f32 : round-to-nearest, ties away -- the operand format of tcgen05 kind::tf32.
__device__ __forceinline__ float hl_rna_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

__device__ __forceinline__ float hl_silu(float x) { return x / (1.0f + expf(-x)); }
// SiLU for values that are about to be rounded to an 11-bit significand (fp16 / TF32 operands):
// ex2.approx + rcp.approx, relative error ~2e-7, 3x fewer instructions than the exact form
__device__ __forceinline__ float hl_silu_fast(float x) { return __fdividef(x, 1.0f + __expf(-x)); }

__device__ __forceinline__ float hl_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
