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

// cudaFuncSetAttribute (opt-in dynamic shared memory) is per device: `static HlPerDeviceOnce once; if (once.need()) ...`
// runs the configuration once for every device ordinal a process launches on.
struct HlPerDeviceOnce {
    bool done[64] = {};
    bool need() {
        int d = 0;
        cudaGetDevice(&d);
        d &= 63;
        if (done[d]) return false;
        done[d] = true;
        return true;
    }
};

static inline int hl_cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// cvt.rna.tf32.f32 : round-to-nearest, ties away -- the operand format of tcgen05 kind::tf32.
__device__ __forceinline__ float hl_rna_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// MUFU ex2 / lg2 without the denormal-range fix-up sequence nvcc wraps around __expf / __logf / exp2f
// (FSETP + 2 FMUL per call: measured 12 -> 6 instructions per softplus in the render MLP).  Flush-to-zero is exact
// enough wherever the result feeds 1 + e^x or a value that is rounded to an 11-bit significand.
__device__ __forceinline__ float hl_ex2(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float hl_lg2(float x) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ float hl_silu(float x) { return x / (1.0f + expf(-x)); }
// SiLU for values that are about to be rounded to an 11-bit significand (fp16 / TF32 operands):
// ex2.approx + rcp.approx, relative error ~2e-7, 3x fewer instructions than the exact form
__device__ __forceinline__ float hl_silu_fast(float x) { return __fdividef(x, 1.0f + hl_ex2(-1.4426950408889634f * x)); }

__device__ __forceinline__ float hl_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL).  A denoise step is ~540 short launches; with the attribute set
// the next kernel's CTAs are scheduled as soon as every CTA of the previous one has started (or exited),
// run their prologue (barrier init, TMEM alloc, descriptor prefetch) and block in griddepcontrol.wait
// until the previous grid has completed and its memory is visible -- the launch gap and the prologue
// disappear from the critical path.  Rule: every kernel that can be launched with the attribute calls
// hl_pdl_wait() before its first global access to anything another kernel writes or reads (so
// completion is transitive along the stream); both instructions are no-ops in a normal launch.
// ---------------------------------------------------------------------------------------------
#ifndef HL_PDL_TRIGGER
#define HL_PDL_TRIGGER 1     // 0: never (dependents start at completion), 1: at kernel entry, 2: conv after its last MMA
#endif
__device__ __forceinline__ void hl_pdl_trigger() {
    if (HL_PDL_TRIGGER != 0) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void hl_pdl_trigger_early() { if (HL_PDL_TRIGGER == 1) hl_pdl_trigger(); }
__device__ __forceinline__ void hl_pdl_trigger_late() { if (HL_PDL_TRIGGER == 2) hl_pdl_trigger(); }
__device__ __forceinline__ void hl_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void hl_pdl_enter() { hl_pdl_trigger(); hl_pdl_wait(); }

extern int g_hl_pdl;            // hl_set_pdl(): 1 = launch with programmatic stream serialization
extern int g_hl_pdl_skip;       // hl_pdl_barrier(): the next launch is a normal (fully serialized) one
extern long long g_hl_launches; // kernels launched (or captured) by this library: hl_launch_count()

// fills cfg.attrs[n_attrs..] ; returns the new attribute count
static inline unsigned hl_pdl_attr(cudaLaunchAttribute *attrs, unsigned n) {
    ++g_hl_launches;
    if (g_hl_pdl_skip) { g_hl_pdl_skip = 0; return n; }
    if (!g_hl_pdl) return n;
    attrs[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[n].val.programmaticStreamSerializationAllowed = 1;
    return n + 1;
}

template <typename... KArgs, typename... Args>
static inline cudaError_t hl_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                    Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attrs[1];
    cfg.attrs = attrs;
    cfg.numAttrs = hl_pdl_attr(attrs, 0);
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
