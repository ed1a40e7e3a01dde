// Canonical-space deformation shared by the render kernels (render.cu: exact fp32 MLP; render_tc5.cu: tcgen05 MLP).
// human_diffusion/NeRF/renderer.py:52-133 (deform_target2c, deform_target2c_op) == recon_NeRF/lib/renderer.py:60-140.
//
// Per frame hl_smpl_vertex_tables leaves three tables in HBM (L2-resident: 6,890 vertices -> 110 KB + 2 KB + 330 KB):
//   verts   [NC * CL] float4 {x, y, z, vertex index}: body vertices in the SMPL frame, grouped into NC spatial clusters of
//           CL slots (clusters are fixed per asset: a k-d split of the template, humanliff_b200/smpl.py; skinning is smooth,
//           so they stay compact under any pose); unused slots hold x = 1e18
//   spheres [NC] float4 {centre, radius} of each cluster's posed vertices
//   aff     [V][3] float4: rows of M | c -- everything deform_target2c_op gathers per point depends on the nearest vertex
//           only, so the whole chain is one affine per vertex: canonical = M q + c, canonical direction = M d
//
// Nearest vertex (pytorch3d knn_points, K = 1), EXACT: pass 1 bounds the answer by U = min over clusters of
// (|q - centre| + radius); pass 2 scans only clusters whose lower bound |q - centre| - radius does not exceed the best
// distance so far.  The distance of a vertex is evaluated with the same fp32 operations in the same order as the
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

// Shared-memory scratch of the search (float4 units): the NC bounding spheres, staged once per CTA, then one staging row
// of HL_CANON_CL_MAX vertices per warp (8 warps).  With ~200 KB of a CTA's shared memory holding MLP weights the L1 is
// too small to keep the 110 KB vertex table, and a warp-uniform load that misses it costs an L2 round trip per vertex:
// a cluster is therefore fetched with ONE coalesced load per lane and scanned out of shared memory.
constexpr int HL_CANON_NC_MAX = 128, HL_CANON_CL_MAX = 96;
constexpr int HL_CANON_SMEM_F4 = HL_CANON_NC_MAX + 8 * HL_CANON_CL_MAX;

__device__ __forceinline__ void hl_canon_stage_spheres(const CanonTables &t, float4 *sm_f4) {
    for (int i = threadIdx.x; i < t.NC; i += blockDim.x) sm_f4[i] = __ldg(t.spheres + i);
}

// scan one cluster: the warp copies it into its staging row, then every lane updates its own (best, bi)
__device__ __forceinline__ void hl_scan_cluster(const float4 *__restrict__ v, float4 *stage, int CL, int lane, float qx,
                                                float qy, float qz, float &best, int &bi) {
    for (int k = lane; k < CL; k += 32) stage[k] = __ldg(v + k);
    __syncwarp();
#pragma unroll 4
    for (int k = 0; k < CL; ++k) {
        const float4 p = stage[k];
        const float d = hl_dist2(qx, qy, qz, p.x, p.y, p.z);
        const int idx = __float_as_int(p.w);
        if (d < best || (d == best && idx < bi)) { best = d; bi = idx; }
    }
    __syncwarp();
}

// Nearest vertex of q.  Must be called by all 32 lanes of a warp; NC <= 128, CL <= HL_CANON_CL_MAX.
//   pass 1: U = min_c (|q - centre_c| + radius_c) bounds the answer from above, c* = its arg-min;
//   phase A: the warp scans the c* of each of its lanes (usually one or two distinct clusters: adjacent samples of a ray) --
//            after it `best` is within a cluster diameter of the true minimum, which is what makes phase B selective even
//            for points metres away from the body (miss rays), where all upper bounds look alike;
//   phase B: clusters [c0, c1) not scanned yet whose lower bound |q - centre| - radius does not exceed `best`.
// `sph_s` = the spheres in shared memory, `stage` = this warp's staging row.  Callers that split [c0, c1) between threads
// combine the parts by (best, bi), smallest first.
static __device__ __noinline__ void hl_nearest_vertex_impl(const float4 *sph_s, const float4 *__restrict__ verts,
                                                           float4 *stage, int NC, int CL, float qx, float qy, float qz,
                                                           int c0, int c1, float &best, int &bi) {
    const int lane = threadIdx.x & 31;
    float U = 3.0e38f;
    int cstar = 0;
#pragma unroll 4
    for (int c = 0; c < NC; ++c) {
        const float4 s = sph_s[c];
        const float ub = sqrtf(hl_dist2(qx, qy, qz, s.x, s.y, s.z)) + s.w;
        if (ub < U) { U = ub; cstar = c; }
    }
    best = U * U * (1.0f + 8e-6f) + 1e-20f;       // >= the squared distance of the farthest vertex of cluster c*
    bi = 0x7fffffff;
    uint32_t done[4] = {0u, 0u, 0u, 0u};            // warp-uniform: clusters scanned in phase A
    uint32_t todo = 0xffffffffu;
    while (todo) {
        const int c = __shfl_sync(0xffffffffu, cstar, __ffs(todo) - 1);
        hl_scan_cluster(verts + (size_t)c * CL, stage, CL, lane, qx, qy, qz, best, bi);
        done[(c >> 5) & 3] |= 1u << (c & 31);
        todo &= ~__ballot_sync(0xffffffffu, cstar == c);
    }
    for (int c = c0; c < c1; ++c) {
        if ((done[(c >> 5) & 3] >> (c & 31)) & 1u) continue;
        const float4 s = sph_s[c];
        const float lb = sqrtf(hl_dist2(qx, qy, qz, s.x, s.y, s.z)) * (1.0f - 4e-6f) - s.w;
        const bool need = lb <= 0.f || lb * lb <= best;
        if (__any_sync(0xffffffffu, need)) hl_scan_cluster(verts + (size_t)c * CL, stage, CL, lane, qx, qy, qz, best, bi);
    }
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
