// Device helpers shared by the fused cell-run movers (cellrun.cu: orders 1 and 2; cellrun3.cu: order 3): streaming /
// keep-in-L1 loads, the packed node-centred field gather and the sort-key classification.
#pragma once
#include "tgpu_internal.h"
#ifndef CR_LDG_PLAIN
#define CR_LDG_PLAIN 1       // 0: ld.global.nc.L1::evict_last for the field nodes -- measured: no change in time, L1 hit rate or DRAM bytes
#endif

// sum over an NW^3 block of node-centred fields, in the reference's order: x innermost (sum()), then *Sy*Sz
// (particles_movedeposit.F90:801-815); packed fp32 (FFMA2 / FMUL2): the six components sit in three aligned register
// pairs straight out of the two 128-bit loads; same operations and roundings as the scalar form
// node-centred field loads.  (An L1::evict_last priority for them -- the particle records stream through L1 exactly once,
// the field lines are re-read by the next steps -- was measured: no change in time, L1 hit rate or DRAM bytes.)
__device__ __forceinline__ float4 ldg_keep(const float4 *p)
{
#if CR_LDG_PLAIN
    return __ldg(p);
#else
    float4 v;
    asm("ld.global.nc.L1::evict_last.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
#endif
}

// particle records are read exactly once: do not let them displace the field lines in L1
__device__ __forceinline__ float ldg_stream(const float *p)
{
    float v; asm volatile("ld.global.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p)); return v;
}
__device__ __forceinline__ int ldg_stream(const int32_t *p)
{
    int v; asm volatile("ld.global.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p)); return v;
}
__device__ __forceinline__ void prefetch_l1_keep(const void *p)
{
    asm volatile("prefetch.global.L1 [%0];" :: "l"(p));
}

template <int NW>
__device__ __forceinline__ void gather_nodes(const float4 *__restrict__ prim8, int nbase, int mx, int my,
                                             const float *wxs, const float *wys, const float *wzs,
                                             float &e0, float &e1, float &e2, float &b0, float &b1, float &b2)
{
    float2 e01 = make_float2(e0, e1), e2b0 = make_float2(e2, b0), b12 = make_float2(b1, b2);
#pragma unroll
    for (int c3 = 0; c3 < NW; c3++) {
#pragma unroll
        for (int c2 = 0; c2 < NW; c2++) {
            float2 s01 = make_float2(0.f, 0.f), s23 = s01, s45 = s01;
            // 32-bit node index: the fast path is only taken for grids below 2^30 nodes (cellrun_supported)
            const float4 *row = prim8 + (unsigned)(2 * (nbase + mx * (c2 + my * c3)));
#pragma unroll
            for (int c1 = 0; c1 < NW; c1++) {
                const float4 lo = ldg_keep(row + 2 * c1), hi = ldg_keep(row + 2 * c1 + 1);
                const float2 w2 = make_float2(wxs[c1], wxs[c1]);
                s01 = __ffma2_rn(make_float2(lo.x, lo.y), w2, s01);
                s23 = __ffma2_rn(make_float2(lo.z, lo.w), w2, s23);
                s45 = __ffma2_rn(make_float2(hi.x, hi.y), w2, s45);
            }
            const float2 wy2 = make_float2(wys[c2], wys[c2]), wz2 = make_float2(wzs[c3], wzs[c3]);
            e01 = __ffma2_rn(__fmul2_rn(s01, wy2), wz2, e01);
            e2b0 = __ffma2_rn(__fmul2_rn(s23, wy2), wz2, e2b0);
            b12 = __ffma2_rn(__fmul2_rn(s45, wy2), wz2, b12);
        }
    }
    e0 = e01.x; e1 = e01.y; e2 = e2b0.x; b0 = e2b0.y; b1 = b12.x; b2 = b12.y;
}

// periodic wrap / shift into the destination's frame (deposit_particles loop B, particles_movedeposit.F90:1553-1633),
// branch-free; lo / hi tell which side was crossed
__device__ __forceinline__ float wrap1(float x, float lo_edge, float hi_edge, float shift_lo, float shift_hi, bool &lo, bool &hi)
{
    lo = x < lo_edge; hi = x > hi_edge;
    return x + (lo ? shift_lo : hi ? -shift_hi : 0.f);
}

// same classification as k_classify_key (particles.cu), on a copy of the position
__device__ __forceinline__ uint32_t sort_key(const DevGeom &G, unsigned keyoff, int general, float x, float y, float z)
{
    bool lx, hx, ly, hy, lz, hz;
    const float xs = wrap1(x, G.minx, G.maxx, G.shiftx_lo, G.shiftx_hi, lx, hx);
    const float ys = wrap1(y, G.miny, G.maxy, G.shifty_lo, G.shifty_hi, ly, hy);
    const float zs = wrap1(z, G.minz, G.maxz, G.shiftz_lo, G.shiftz_hi, lz, hz);
    // (a NaN position converts to 0 and the unsigned min below keeps the key inside the table)
    uint32_t key = G.rowblk ? cell_key(G, (int)xs, (int)ys, (int)zs)
                            : (uint32_t)((int)xs + G.mx * ((int)ys + G.my * (int)zs)) - keyoff;
    key = min(key, (uint32_t)G.nkeys - 1u);
    if (general) {                                           // kernel-uniform: open or split axes
        bool in = true;
        if (!G.perx) in = (x + G.mxcum > G.x1in) && (x + G.mxcum < G.x2in);
        if (!G.pery && in) in = (y + G.mycum > G.y1in) && (y + G.mycum < G.y2in);
        if (!G.perz && in) in = (z + G.mzcum > G.z1in) && (z + G.mzcum < G.z2in);
        const int dy = (int)hy - (int)ly, dz = (int)hz - (int)lz;
        const int code = ((G.sendy ? dy : 0) + 1) + 3 * ((G.sendz ? dz : 0) + 1);
        if (code != 4) key = (uint32_t)G.nkeys + (uint32_t)code;
        if (!in) key = (uint32_t)G.nkeys + 9u;
    }
    return key;
}

