// Canonical-space deformation shared by the render kernels (render.cu: exact fp32 MLP; render_tc5.cu: tcgen05 MLP).
// human_diffusion/NeRF/renderer.py:52-133 (deform_target2c, deform_target2c_op) == recon_NeRF/lib/renderer.py:60-140.
//
// Per frame hl_smpl_vertex_tables leaves three tables in HBM (L2-resident: 6,890 vertices -> 110 KB + 2 KB + 330 KB):
//   verts   [NC * CL] float4, two vertices per PAIR of float4 {x0, x1, y0, y1} {z0, z1, index0, index1} (the layout the packed
//           fp32 instructions of sm_100 -- FADD2 / FMUL2 / FFMA2: two lanes per instruction -- want): body vertices in the SMPL frame, grouped into NC spatial clusters of CL slots (clusters are fixed per asset: a k-d split of the template, humanliff_b200/smpl.py; skinning is smooth,
//           so they stay compact under any pose); unused slots hold x = 1e18
//   spheres [NC] float4 in the same pair layout {cx0, cx1, cy0, cy1} {cz0, cz1, r0, r1}: centre and radius of each cluster's
//           posed vertices
//   aff     [V][3] float4: rows of M | c -- everything deform_target2c_op gathers per point depends on the nearest vertex
//           only, so the whole chain is one affine per vertex: canonical = M q + c, canonical direction = M d
//
// Nearest vertex (pytorch3d knn_points, K = 1), EXACT: the cluster with the nearest centre is scanned first; after it only
// clusters whose lower bound |q - centre| - radius does not exceed the best distance so far can hold the answer.  The distance of a vertex is evaluated with the same fp32 operations in the same order as the
// brute-force scan, and ties go to the lowest ORIGINAL vertex index, so the result equals the full scan's -- the bounds
// only skip vertices that cannot win (slack factors make them conservative under rounding).  A warp scans a cluster when
// any of its lanes needs it (adjacent samples of a ray need the same few clusters): loads stay warp-uniform broadcasts
// served by L1.  Typical cost: 2 x NC sphere tests + 5-10 clusters x CL vertices instead of 6,890 vertices.
#include <stdint.h>
#pragma once
#include <cuda_runtime.h>

struct CanonTables {
    const float4 *verts;
    const float4 *spheres;
    const float4 *aff;
    int NC, CL;
    float Rm[9], Th[3];      // params['R'] (row-major), params['Th']:  q = (p - Th) R
};

// (v - Th) R  (row vector times matrix, renderer.py:124-125)
__device__ __forceinline__ void hl_to_smpl_frame(const CanonTables &t, float x, float y, float z, float &qx, float &qy,
                                                 float &qz) {
    const float ex = x - t.Th[0], ey = y - t.Th[1], ez = z - t.Th[2];
    qx = fmaf(ez, t.Rm[6], fmaf(ey, t.Rm[3], ex * t.Rm[0]));
    qy = fmaf(ez, t.Rm[7], fmaf(ey, t.Rm[4], ex * t.Rm[1]));
    qz = fmaf(ez, t.Rm[8], fmaf(ey, t.Rm[5], ex * t.Rm[2]));
}

__device__ __forceinline__ float hl_dist2(float qx, float qy, float qz, float x, float y, float z) {
    const float ax = qx - x, ay = qy - y, az = qz - z;
    return __fadd_rn(__fadd_rn(__fmul_rn(ax, ax), __fmul_rn(ay, ay)), __fmul_rn(az, az));
}

// ---- packed fp32 (sm_100: FADD2 / FMUL2 / FFMA2, two IEEE single lanes per instruction) ----
// ptxas 12.9 CONTRACTS mul.rn.f32x2 + add.rn.f32x2 into FFMA2 (unlike the scalar mul.rn / add.rn, which it leaves alone:
// 40 of 172 packed PTX instructions of this file came out fused), so a packed distance cannot reproduce the separately
// rounded hl_dist2.  The packed distance is therefore DEFINED with explicit fused multiply-adds and used only where a
// bound with slack is all that is needed: the sphere passes, and the FILTER of the vertex scans.  Every nearest-vertex
// decision is taken on the scalar, separately rounded hl_dist2 (the oracle's arithmetic).
typedef unsigned long long hl_f32x2;
__device__ __forceinline__ hl_f32x2 hl_pk(float lo, float hi) {
    hl_f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void hl_unpk(hl_f32x2 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ hl_f32x2 hl_sub2(hl_f32x2 a, hl_f32x2 b) { hl_f32x2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ hl_f32x2 hl_add2(hl_f32x2 a, hl_f32x2 b) { hl_f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ hl_f32x2 hl_mul2(hl_f32x2 a, hl_f32x2 b) { hl_f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ hl_f32x2 hl_fma2(hl_f32x2 a, hl_f32x2 b, hl_f32x2 c) { hl_f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
// squared distances of q to the two points of a pair {x0, x1, y0, y1} {z0, z1, ., .}: fma(dz, dz, fma(dy, dy, dx * dx)) per
// lane, 6 packed instructions.  With d = hl_dist2 of the same point and D the real-number sum of the three squares (the
// differences dx, dy, dz are the same fp32 values on both paths): |d - D| <= 3 * 2^-24 D and |d_fma - D| <= 3 * 2^-24 D, so
// d_fma <= d (1 + 4e-7): a group of vertices can only hold a winner (d <= best) if min(d_fma) (1 - 1e-6) <= best.
constexpr float HL_CANON_FILTER_SLACK = 1.0f - 1e-6f;
__device__ __forceinline__ hl_f32x2 hl_dist2x2_fma(hl_f32x2 qx, hl_f32x2 qy, hl_f32x2 qz, const ulonglong2 &a, const ulonglong2 &b) {
    const hl_f32x2 dx = hl_sub2(qx, a.x), dy = hl_sub2(qy, a.y), dz = hl_sub2(qz, b.x);
    return hl_fma2(dz, dz, hl_fma2(dy, dy, hl_mul2(dx, dx)));
}
// element e of pair-layout entry i (spheres: x, y, z = centre, w = radius; vertices: w = index bits)
__device__ __forceinline__ float4 hl_pair_get(const float4 *t, int i) {
    const float *a = reinterpret_cast<const float *>(t + 2 * (i >> 1)), *b = a + 4;
    const int e = i & 1;
    return make_float4(a[e], a[2 + e], b[e], b[2 + e]);
}
__device__ __forceinline__ void hl_pair_put(float4 *t, int i, float x, float y, float z, float w) {
    float *a = reinterpret_cast<float *>(t + 2 * (i >> 1)), *b = a + 4;
    const int e = i & 1;
    a[e] = x; a[2 + e] = y; b[e] = z; b[2 + e] = w;
}

// sqrt.approx (2 ulp): every use below carries a relative slack of 4e-6 or more
__device__ __forceinline__ float hl_sqrt_approx(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// Shared-memory scratch of the search (float4 units): the NC bounding spheres, staged once per CTA, then one staging row
// of HL_CANON_CL_MAX vertices per warp (8 warps).  With ~200 KB of a CTA's shared memory holding MLP weights the L1 is
// too small to keep the 110 KB vertex table, and a warp-uniform load that misses it costs an L2 round trip per vertex:
// a cluster is therefore fetched with ONE coalesced load per lane and scanned out of shared memory.
#ifndef HL_CANON_UNROLL
#define HL_CANON_UNROLL 2      // blocks of four per loop iteration of the sphere / vertex scans (measured: profiles/r2_canon_search_ab.log)
#endif
constexpr int kCanonUnroll = HL_CANON_UNROLL;      // (#pragma unroll takes a constant expression, not a macro)
constexpr int HL_CANON_NC_MAX = 128, HL_CANON_CL_MAX = 96;
constexpr int HL_CANON_SMEM_F4 = HL_CANON_NC_MAX + 8 * HL_CANON_CL_MAX;

// all HL_CANON_NC_MAX entries are written (unused ones far away, radius 0: never the arg-min, never a candidate), so the
// loops over the spheres run a fixed trip count in blocks of four without bounds checks
__device__ __forceinline__ void hl_canon_stage_spheres(const CanonTables &t, float4 *sm_f4) {
    for (int i = threadIdx.x; i < HL_CANON_NC_MAX; i += blockDim.x)      // pair layout: even float4 = x | y, odd = z | radius
        sm_f4[i] = i < t.NC ? __ldg(t.spheres + i) : ((i & 1) ? make_float4(1e18f, 1e18f, 0.f, 0.f) : make_float4(1e18f, 1e18f, 1e18f, 1e18f));
}

// A cluster in flight: lane l holds slots l, l + 32, l + 64 of the cluster (CL <= 96) in registers.
struct HlClusterRegs { float4 a, b, c; };
__device__ __forceinline__ void hl_cluster_load(HlClusterRegs &r, const float4 *__restrict__ v, int CL, int lane) {
    const float4 pad = make_float4(1e18f, 1e18f, 1e18f, 1e18f);      // (never used: CL float4 are always present)
    r.a = lane < CL ? __ldg(v + lane) : pad;
    r.b = lane + 32 < CL ? __ldg(v + lane + 32) : pad;
    r.c = lane + 64 < CL ? __ldg(v + lane + 64) : pad;
}
__device__ __forceinline__ void hl_cluster_store(const HlClusterRegs &r, float4 *stage, int CL, int lane) {
    if (lane < CL) stage[lane] = r.a;
    if (lane + 32 < CL) stage[lane + 32] = r.b;
    if (lane + 64 < CL) stage[lane + 64] = r.c;
}
// scan the staged cluster: every lane updates its own (best, bi); branch-free, the loads are warp-uniform broadcasts
__device__ __forceinline__ void hl_scan_staged(const float4 *stage, int CL, float qx, float qy, float qz, float &best,
                                               int &bi) {
    // four vertices = two pairs at a time (CL is a multiple of 4): 12 packed instructions for the four FILTER distances, then
    // ONE comparison of their minimum against `best` -- after the first cluster almost no group can win, so the exact
    // distances and the serial compare-select chain of the update are off the common path
    const hl_f32x2 q2x = hl_pk(qx, qx), q2y = hl_pk(qy, qy), q2z = hl_pk(qz, qz);
    const ulonglong2 *st2 = reinterpret_cast<const ulonglong2 *>(stage);
#pragma unroll kCanonUnroll
    for (int k = 0; k < CL; k += 4) {
        const ulonglong2 a0 = st2[k], b0 = st2[k + 1], a1 = st2[k + 2], b1 = st2[k + 3];
        float d0, d1, d2, d3;
        hl_unpk(hl_dist2x2_fma(q2x, q2y, q2z, a0, b0), d0, d1);
        hl_unpk(hl_dist2x2_fma(q2x, q2y, q2z, a1, b1), d2, d3);
        if (fminf(fminf(d0, d1), fminf(d2, d3)) * HL_CANON_FILTER_SLACK <= best) {
            float x0, x1, y0, y1, z0, z1;                               // the separately rounded distances decide
            hl_unpk(a0.x, x0, x1); hl_unpk(a0.y, y0, y1); hl_unpk(b0.x, z0, z1);
            d0 = hl_dist2(qx, qy, qz, x0, y0, z0); d1 = hl_dist2(qx, qy, qz, x1, y1, z1);
            hl_unpk(a1.x, x0, x1); hl_unpk(a1.y, y0, y1); hl_unpk(b1.x, z0, z1);
            d2 = hl_dist2(qx, qy, qz, x0, y0, z0); d3 = hl_dist2(qx, qy, qz, x1, y1, z1);
            const float ds[4] = {d0, d1, d2, d3};
            const int is[4] = {(int)(unsigned)b0.y, (int)(unsigned)(b0.y >> 32), (int)(unsigned)b1.y, (int)(unsigned)(b1.y >> 32)};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const bool take = (ds[j] < best) | ((ds[j] == best) & (is[j] < bi));
                best = take ? ds[j] : best;
                bi = take ? is[j] : bi;
            }
        }
    }
}

// Nearest vertex of q.  Must be called by all 32 lanes of a warp; NC <= 128, CL <= HL_CANON_CL_MAX.
//   pass 1: c* = the cluster with the nearest centre; U = |q - centre| + radius of c* bounds the answer from above;
//   phase A: the warp scans the c* of each of its lanes (usually one or two distinct clusters: adjacent samples of a ray) --
//            after it `best` is within a cluster diameter of the true minimum, which is what makes phase B selective even
//            for points metres away from the body (miss rays), where all upper bounds look alike;
//   phase B: candidate mask = clusters of [c0, c1) not scanned yet whose lower bound |q - centre| - radius does not exceed
//            `best` for ANY lane (one pipelined pass over the spheres, OR-reduced over the warp); the candidates are then
//            visited in turn -- re-tested against the current `best`, fetched one ahead into registers while the
//            previous one is scanned out of shared memory.
// `sph_s` = the spheres in shared memory, `stage` = this warp's staging row.  Callers that split [c0, c1) between threads
// combine the parts by (best, bi), smallest first.
static __device__ __noinline__ void hl_nearest_vertex_impl(const float4 *sph_s, const float4 *__restrict__ verts,
                                                           float4 *stage, int NC, int CL, float qx, float qy, float qz,
                                                           int c0, int c1, float &best_out, int &bi_out,
                                                           unsigned long long *prof = nullptr) {
    // prof (one thread of a probe run): [0] pass 1, [1] phase A, [2] candidate mask, [3] phase B cycles; [4] clusters
    // scanned in phase A, [5] candidates, [6] candidates scanned, [7] calls
    long long tq = prof ? clock64() : 0;
#define HL_CPROF(slot)                                                         \
    if (prof) {                                                                \
        const long long now_ = clock64();                                      \
        atomicAdd(prof + (slot), (unsigned long long)(now_ - tq));             \
        tq = now_;                                                             \
    }
    const int lane = threadIdx.x & 31;
    float best;            // registers, not the caller's stack slots: the references are written once at the end
    int bi;
    // pass 1: c* = the cluster with the nearest centre (any choice is correct; this one needs no square root per cluster)
    float dmin = 3.0e38f;
    int cstar = 0;
    const hl_f32x2 q2x = hl_pk(qx, qx), q2y = hl_pk(qy, qy), q2z = hl_pk(qz, qz);
    const ulonglong2 *sp2 = reinterpret_cast<const ulonglong2 *>(sph_s);
#pragma unroll kCanonUnroll
    for (int c = 0; c < HL_CANON_NC_MAX; c += 4) {
        const ulonglong2 a0 = sp2[c], b0 = sp2[c + 1], a1 = sp2[c + 2], b1 = sp2[c + 3];
        float u0, u1, u2, u3;
        hl_unpk(hl_dist2x2_fma(q2x, q2y, q2z, a0, b0), u0, u1);
        hl_unpk(hl_dist2x2_fma(q2x, q2y, q2z, a1, b1), u2, u3);
        const bool b01 = u1 < u0, b23 = u3 < u2;
        const float m01 = b01 ? u1 : u0, m23 = b23 ? u3 : u2;
        const int i01 = b01 ? c + 1 : c, i23 = b23 ? c + 3 : c + 2;
        const bool b = m23 < m01;
        const float mm = b ? m23 : m01;
        const int im = b ? i23 : i01;
        if (mm < dmin) { dmin = mm; cstar = im; }
    }
    const float U = hl_sqrt_approx(dmin) + hl_pair_get(sph_s, cstar).w;
    best = U * U * (1.0f + 8e-6f) + 1e-20f;       // >= the squared distance of the farthest vertex of cluster c*
    bi = 0x7fffffff;
    HL_CPROF(0)
    uint32_t done0 = 0u, done1 = 0u, done2 = 0u, done3 = 0u;     // warp-uniform: clusters scanned in phase A
    uint32_t todo = 0xffffffffu;
    HlClusterRegs regs;
    while (todo) {
        const int c = __shfl_sync(0xffffffffu, cstar, __ffs(todo) - 1);
        hl_cluster_load(regs, verts + (size_t)c * CL, CL, lane);
        hl_cluster_store(regs, stage, CL, lane);
        __syncwarp();
        hl_scan_staged(stage, CL, qx, qy, qz, best, bi);
        __syncwarp();
        const uint32_t bit = 1u << (c & 31);
        if ((c >> 5) == 0) done0 |= bit; else if ((c >> 5) == 1) done1 |= bit; else if ((c >> 5) == 2) done2 |= bit; else done3 |= bit;
        todo &= ~__ballot_sync(0xffffffffu, cstar == c);
        if (prof) atomicAdd(prof + 4, 1ull);
    }
    HL_CPROF(1)
    // candidate mask: |q - centre| - radius <= sqrt(best)  <=>  |q - centre|^2 <= (sqrt(best) + radius)^2
    float sb = hl_sqrt_approx(best) * (1.0f + 4e-6f) + 1e-18f;
    // (compact loops throughout: the render kernels are instruction-cache bound -- ncu: 2.3 warps stalled on instruction
    // fetch per issued instruction before this function was shrunk -- so nothing here is unrolled beyond a block of four)
    uint32_t m[4] = {0u, 0u, 0u, 0u};
    const hl_f32x2 sb2 = hl_pk(sb, sb), slack2 = hl_pk(1.0f - 8e-6f, 1.0f - 8e-6f);
#pragma unroll 1
    for (int w = 0; w < 4; ++w) {
        uint32_t bits = 0u;
#pragma unroll kCanonUnroll
        for (int b = 0; b < 32; b += 4) {
            const ulonglong2 a0 = sp2[w * 32 + b], b0 = sp2[w * 32 + b + 1], a1 = sp2[w * 32 + b + 2], b1 = sp2[w * 32 + b + 3];
            const hl_f32x2 r01 = hl_add2(sb2, b0.y), r23 = hl_add2(sb2, b1.y);            // sb + radius
            float e0, e1, e2, e3, f0, f1, f2, f3;
            hl_unpk(hl_mul2(hl_dist2x2_fma(q2x, q2y, q2z, a0, b0), slack2), e0, e1);
            hl_unpk(hl_mul2(hl_dist2x2_fma(q2x, q2y, q2z, a1, b1), slack2), e2, e3);
            hl_unpk(hl_mul2(r01, r01), f0, f1);
            hl_unpk(hl_mul2(r23, r23), f2, f3);
            const uint32_t n0 = e0 <= f0, n1 = e1 <= f1, n2 = e2 <= f2, n3 = e3 <= f3;
            bits |= (n0 | (n1 << 1) | (n2 << 2) | (n3 << 3)) << b;
        }
        // clusters [c0, c1) only
        const int lo = c0 - w * 32, hi = c1 - w * 32;
        const uint32_t range = (hi <= 0 || lo >= 32) ? 0u
                               : ((hi >= 32 ? 0xffffffffu : ((1u << hi) - 1u)) & (lo <= 0 ? 0xffffffffu : ~((1u << lo) - 1u)));
        bits = __reduce_or_sync(0xffffffffu, bits & range);
        m[0] = w == 0 ? bits : m[0];
        m[1] = w == 1 ? bits : m[1];
        m[2] = w == 2 ? bits : m[2];
        m[3] = w == 3 ? bits : m[3];
    }
    m[0] &= ~done0; m[1] &= ~done1; m[2] &= ~done2; m[3] &= ~done3;
    HL_CPROF(2)
    if (prof) atomicAdd(prof + 5, (unsigned long long)(__popc(m[0]) + __popc(m[1]) + __popc(m[2]) + __popc(m[3])));
    // visit the candidates; the next one is in flight while the current one is scanned
    auto pop = [&]() -> int {          // lowest candidate, or -1
        int c = -1;
        if (m[0]) { c = __ffs(m[0]) - 1; m[0] &= m[0] - 1u; }
        else if (m[1]) { c = 32 + __ffs(m[1]) - 1; m[1] &= m[1] - 1u; }
        else if (m[2]) { c = 64 + __ffs(m[2]) - 1; m[2] &= m[2] - 1u; }
        else if (m[3]) { c = 96 + __ffs(m[3]) - 1; m[3] &= m[3] - 1u; }
        return c;
    };
    int cur = pop();
    if (cur >= 0) hl_cluster_load(regs, verts + (size_t)cur * CL, CL, lane);
#pragma unroll 1
    while (cur >= 0) {
        hl_cluster_store(regs, stage, CL, lane);
        __syncwarp();
        const int now = cur;
        cur = pop();
        if (cur >= 0) hl_cluster_load(regs, verts + (size_t)cur * CL, CL, lane);
        const float4 s = hl_pair_get(sph_s, now);
        const float reach = sb + s.w;
        const bool need = hl_dist2(qx, qy, qz, s.x, s.y, s.z) * (1.0f - 8e-6f) <= reach * reach;
        if (__any_sync(0xffffffffu, need)) {
            hl_scan_staged(stage, CL, qx, qy, qz, best, bi);
            sb = hl_sqrt_approx(best) * (1.0f + 4e-6f) + 1e-18f;
            if (prof) atomicAdd(prof + 6, 1ull);
        }
        __syncwarp();
    }
    HL_CPROF(3)
    if (prof) atomicAdd(prof + 7, 1ull);
#undef HL_CPROF
    best_out = best;
    bi_out = bi;
}

// `sm_f4`: the CTA's search scratch (HL_CANON_SMEM_F4 float4, spheres staged by hl_canon_stage_spheres)
__device__ __forceinline__ void hl_nearest_vertex(const CanonTables &t, float4 *sm_f4, float qx, float qy, float qz, int c0,
                                                  int c1, float &best, int &bi) {
    hl_nearest_vertex_impl(sm_f4, t.verts, sm_f4 + HL_CANON_NC_MAX + ((threadIdx.x >> 5) & 7) * HL_CANON_CL_MAX, t.NC, t.CL,
                           qx, qy, qz, c0, c1, best, bi);
}

// canonical position of q (SMPL frame) through vertex v's affine; optionally the canonical direction of sv
__device__ __forceinline__ void hl_canon_affine(const CanonTables &t, int v, float qx, float qy, float qz, float &px,
                                                float &py, float &pz, const float *sv, float *dc) {
    const float4 m0 = __ldg(t.aff + (size_t)v * 3), m1 = __ldg(t.aff + (size_t)v * 3 + 1), m2 = __ldg(t.aff + (size_t)v * 3 + 2);
    px = fmaf(m0.z, qz, fmaf(m0.y, qy, fmaf(m0.x, qx, m0.w)));
    py = fmaf(m1.z, qz, fmaf(m1.y, qy, fmaf(m1.x, qx, m1.w)));
    pz = fmaf(m2.z, qz, fmaf(m2.y, qy, fmaf(m2.x, qx, m2.w)));
    if (dc) {
        dc[0] = fmaf(m0.z, sv[2], fmaf(m0.y, sv[1], m0.x * sv[0]));
        dc[1] = fmaf(m1.z, sv[2], fmaf(m1.y, sv[1], m1.x * sv[0]));
        dc[2] = fmaf(m2.z, sv[2], fmaf(m2.y, sv[1], m2.x * sv[0]));
    }
}
