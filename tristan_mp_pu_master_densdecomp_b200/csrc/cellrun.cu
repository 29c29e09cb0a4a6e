// Cell-run kernels: the fast path for 3D, shape orders 1 and 2 (-Ddd1 / -Ddd2).
//
// What they replace:  mover_1ord / mover_2ord        code/particles_movedeposit.F90:356-610, 619-933
//                     densdecomp_1ord / _2ord        code/particles.F90:678-854, 864-1102
//                     loop A of deposit_particles    code/particles_movedeposit.F90:1381-1401, 1717-1737
//
// Design.  Shared-memory fp32 atomics are CAS loops on sm_100 (ATOMS.CAST.SPIN) and a 2nd-order particle touches up
// to 4x4x4 cells x 3 components, so a scatter-per-particle deposit is atomic-bound far below the HBM roofline.
// Particles are kept sorted by cell (x fastest) by the counting sort that follows every deposit.  All particles of a
// cell share the same 4x4x4 output footprint (slots 2..5 of the reference's 6-slot stencil; |dx| < c < 1/2 cell per
// step keeps both the old and the new shape inside it).  So the deposit is made OUTPUT-STATIONARY:
//   record pipeline: each lane fetches the record of its NEXT step (through the lazy sort's permutation, whose entry was
//           loaded one step earlier) with 4-byte cp.async into per-lane landing slots while the current step is deposited;
//   phase 1 (lane = particle): read the landed record, [gather node-centred fields (packed FFMA2) + Boris push + store +
//           sort key of the new cell + rank in its bin], build the 1-D factors of the Esirkepov sum from the old and new
//           shapes and park them in shared memory (44 floats per particle).  Deposit-only launches recompute the old
//           position as the reference does (x - u/gamma*c); fused launches use the gather's shape at the true pre-push
//           position (equal to round-off);
//   phase 2 (half-warp = one footprint): lane (j,k) of a half-warp owns the 4 x-cells of row (j,k) of the footprint
//           for all three components = 12 register accumulators (6 FFMA2 pairs).  It walks its 16 particles, reading the
//           factors with broadcast 128-bit shared loads; the two half-warps run in lockstep.  The x-planes form a ring
//           (plane of cell x lives in register x mod 4, phase 1 stores the x factors pre-rotated): when the cell advances
//           along x the completed plane is flushed with one predicated fp32 RED per (cell, component) into the tiled
//           shadow arrays (tgpu_internal.h row_index) and cleared; nothing moves between registers.
// With ~8 particles per cell and species, that is ~6 global REDs per particle instead of up to 192 atomics, and the
// arithmetic is the factorised form of Appendix A.3 (Jx = q*Wx(j,k)*prefix_i(dSx), ...).
// Nothing here depends on the particles being sorted for correctness -- an unsorted tail (fresh arrivals) only makes
// the window jump and flush more often (which is why freshly uploaded records are sorted once, cellrun_move_deposit).
// Measured character (profiles/README.md): 56 warp-instructions per particle, issue slots 70 % and the L1/shared data
// pipe 78 % busy, DRAM 24 % of peak: co-limited by instruction issue and LSU wavefronts, not by HBM.
#include "tgpu_internal.h"
#include "shapes.cuh"

#ifndef CR_WARPS
#define CR_WARPS 8
#endif
#define CR_STRIDE 44
#ifndef CR_MINB
#define CR_MINB 3
#endif
#ifndef CR_CHUNK
#define CR_CHUNK 128          // particles per half-warp
#endif
#ifndef CR_UNROLL
#define CR_UNROLL 4           // phase-2 unroll (particles per half-warp per loop trip)
#endif
#define CR_STR(x) #x
#define CR_DO_PRAGMA(x) _Pragma(CR_STR(x))
#ifndef CR_FFMA2
#define CR_FFMA2 1            // packed fp32 (FFMA2/FMUL2) in the gather and the deposit accumulation: measured 2-3 % faster
#endif
// factor staging of one warp: two halves of 16 particles x CR_STRIDE floats; the second half starts 16 floats (half of
// the 32 banks) further on, so the lockstep phase-2 loads of the two half-warps never fall into the same banks
#define CR_HALF_FLOATS (16 * CR_STRIDE + 16)
#define CR_WARP_FLOATS (2 * CR_HALF_FLOATS)
#define CR_SMEM_BYTES (sizeof(float) * CR_WARPS * CR_WARP_FLOATS + sizeof(uint32_t) * CR_WARPS * 2 * 9 * 32)

struct CRArgs {
    Species s;                // source records
    Species d;                // FUSED: destination records (logical order); may alias s when perm == nullptr
    const int32_t *perm;      // FUSED: logical position t reads physical record perm[t] (lazy sort), or nullptr
    long long n;
    const float4 *prim8;
    float *cx, *cy, *cz;     // FUSED: the tiled shadow arrays (see row_index), else curx, cury, curz in Fortran order
    int nty;                  // FUSED: number of 4-row tiles along y
    DevGeom G;
    float qm, qs;
    uint32_t *key;            // FUSED: sort key of the pushed particle (prt_sort skips its classify pass)
    int32_t *slot, *bincount;
};

__device__ __forceinline__ void red3(float *cx, float *cy, float *cz, size_t idx, float vx, float vy, float vz)
{
    // (a 16-byte red.global.add.v4.f32 into an interleaved array was measured 4 % slower than three scalar REDs)
    red_nz(cx + idx, vx); red_nz(cy + idx, vy); red_nz(cz + idx, vz);
}

// sum over an NW^3 block of node-centred fields, in the reference's order: x innermost (sum()), then *Sy*Sz
// (particles_movedeposit.F90:801-815)
template <int NW>
__device__ __forceinline__ void gather_nodes(const float4 *__restrict__ prim8, int nbase, int mx, int my,
                                             const float *wxs, const float *wys, const float *wzs,
                                             float &e0, float &e1, float &e2, float &b0, float &b1, float &b2)
{
#if CR_FFMA2
    // packed fp32 (FFMA2 / FMUL2): the six components sit in three aligned register pairs straight out of the two
    // 128-bit loads; same operations and roundings as the scalar form below
    float2 e01 = make_float2(e0, e1), e2b0 = make_float2(e2, b0), b12 = make_float2(b1, b2);
#pragma unroll
    for (int c3 = 0; c3 < NW; c3++) {
#pragma unroll
        for (int c2 = 0; c2 < NW; c2++) {
            float2 s01 = make_float2(0.f, 0.f), s23 = s01, s45 = s01;
            const float4 *row = prim8 + (unsigned)(2 * (nbase + mx * (c2 + my * c3)));
#pragma unroll
            for (int c1 = 0; c1 < NW; c1++) {
                const float4 lo = __ldg(row + 2 * c1), hi = __ldg(row + 2 * c1 + 1);
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
#else
#pragma unroll
    for (int c3 = 0; c3 < NW; c3++) {
#pragma unroll
        for (int c2 = 0; c2 < NW; c2++) {
            float s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0, s5 = 0;
            // 32-bit node index: the fast path is only taken for grids below 2^30 nodes (cellrun_supported)
            const float4 *row = prim8 + (unsigned)(2 * (nbase + mx * (c2 + my * c3)));
#pragma unroll
            for (int c1 = 0; c1 < NW; c1++) {
                float4 lo = __ldg(row + 2 * c1), hi = __ldg(row + 2 * c1 + 1);
                s0 = s0 + lo.x * wxs[c1]; s1 = s1 + lo.y * wxs[c1]; s2 = s2 + lo.z * wxs[c1];
                s3 = s3 + lo.w * wxs[c1]; s4 = s4 + hi.x * wxs[c1]; s5 = s5 + hi.y * wxs[c1];
            }
            const float wy_ = wys[c2], wz_ = wzs[c3];
            e0 = e0 + s0 * wy_ * wz_; e1 = e1 + s1 * wy_ * wz_; e2 = e2 + s2 * wy_ * wz_;
            b0 = b0 + s3 * wy_ * wz_; b1 = b1 + s4 * wy_ * wz_; b2 = b2 + s5 * wy_ * wz_;
        }
    }
#endif
}

// 3x3x3 path (particle exactly on a node under Q1, or quirks = fixed): kept out of line so that its register
// demand does not set the occupancy of the common 2x2x2 path
__device__ __noinline__ void gather27(const float4 *__restrict__ prim8, long long nbase, int mx, int my, const float *w9, float *eb)
{
    float e0 = 0, e1 = 0, e2 = 0, b0 = 0, b1 = 0, b2 = 0;
#pragma unroll 1
    for (int c3 = 0; c3 < 3; c3++) {
#pragma unroll 1
        for (int c2 = 0; c2 < 3; c2++) {
            float s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0, s5 = 0;
            const float4 *row = prim8 + 2 * (nbase + (long long)mx * (c2 + (long long)my * c3));
#pragma unroll 1
            for (int c1 = 0; c1 < 3; c1++) {
                const float4 lo = __ldg(row + 2 * c1), hi = __ldg(row + 2 * c1 + 1);
                const float wx = w9[c1];
                s0 = s0 + lo.x * wx; s1 = s1 + lo.y * wx; s2 = s2 + lo.z * wx;
                s3 = s3 + lo.w * wx; s4 = s4 + hi.x * wx; s5 = s5 + hi.y * wx;
            }
            const float wy_ = w9[3 + c2], wz_ = w9[6 + c3];
            e0 = e0 + s0 * wy_ * wz_; e1 = e1 + s1 * wy_ * wz_; e2 = e2 + s2 * wy_ * wz_;
            b0 = b0 + s3 * wy_ * wz_; b1 = b1 + s4 * wy_ * wz_; b2 = b2 + s5 * wy_ * wz_;
        }
    }
    eb[0] = e0; eb[1] = e1; eb[2] = e2; eb[3] = b0; eb[4] = b1; eb[5] = b2;
}

// 1-D factors of one axis -> shared staging.  MODE 0: x (q*prefix, XA, XB); MODE 1: y/z rows (S1, dS, q*prefix, tag)
// rotate a 4-vector left... component (s + r) & 3 of the result holds v[s]
__device__ __forceinline__ float4 rot4(float v0, float v1, float v2, float v3, int r)
{
    float a0 = (r & 2) ? v2 : v0, a1 = (r & 2) ? v3 : v1, a2 = (r & 2) ? v0 : v2, a3 = (r & 2) ? v1 : v3;   // by 2
    return (r & 1) ? make_float4(a3, a0, a1, a2) : make_float4(a0, a1, a2, a3);                          // by 1
}

// MODE 0: x factors (q*prefix, XA, XB), stored ROTATED by r = (i1 - 1) & 3: component m belongs to the footprint cell
// whose x index is congruent to m mod 4, so the phase-2 accumulators never have to move when the window slides.
// MODE 1: y/z rows (S1, dS, q*prefix, tag).
template <int MODE>
__device__ __forceinline__ void stage_axis(float *st, const float S1[4], const float S2[4], float q, float tag, int r)
{
    const float third = 1.f / 3.f;
    const float d0 = S2[0] - S1[0], d1 = S2[1] - S1[1], d2 = S2[2] - S1[2], d3 = S2[3] - S1[3];
    const float p0 = d0, p1 = p0 + d1, p2 = p1 + d2, p3 = p2 + d3;
    if (MODE == 0) {
        *(float4 *)(st + 0) = rot4(q * p0, q * p1, q * p2, q * p3, r);
        *(float4 *)(st + 4) = rot4(S1[0] + 0.5f * d0, S1[1] + 0.5f * d1, S1[2] + 0.5f * d2, S1[3] + 0.5f * d3, r);
        *(float4 *)(st + 8) = rot4(0.5f * S1[0] + third * d0, 0.5f * S1[1] + third * d1,
                                   0.5f * S1[2] + third * d2, 0.5f * S1[3] + third * d3, r);
    } else {
        *(float4 *)(st + 0) = make_float4(S1[0], d0, q * p0, tag);
        *(float4 *)(st + 4) = make_float4(S1[1], d1, q * p1, tag);
        *(float4 *)(st + 8) = make_float4(S1[2], d2, q * p2, tag);
        *(float4 *)(st + 12) = make_float4(S1[3], d3, q * p3, tag);
    }
}

// flush the accumulators of the first `nplanes` x-planes of the window whose first cell is wi-1 (plane t lives in
// physical register (wi - 1 + t) & 3) and clear them
template <int PS>
__device__ __forceinline__ void flush_planes(float *cx, float *cy, float *cz, size_t idx0, int wi, int nplanes,
                                             float (&ax)[4], float (&ay)[4], float (&az)[4])
{
#pragma unroll
    for (int m = 0; m < 4; m++) {
        const int t = (m - (wi - 1)) & 3;                    // plane held by register m
        if (t < nplanes) { red3(cx, cy, cz, idx0 + (size_t)(t * PS), ax[m], ay[m], az[m]); ax[m] = 0.f; ay[m] = 0.f; az[m] = 0.f; }
    }
}

// the common case in sorted order: the window slides one cell along x, plane 0 (register (wi-1)&3) is complete.
// The ring register is uniform across the half-warp, so a 4-way switch costs one short divergent body instead of
// select chains over all twelve accumulators.
__device__ __forceinline__ void flush_one(float *cx, float *cy, float *cz, size_t idx0, int wi, float (&ax)[4], float (&ay)[4], float (&az)[4])
{
    float *px = cx + idx0, *py = cy + idx0, *pz = cz + idx0;
    switch ((wi - 1) & 3) {
    case 0: red_nz(px, ax[0]); red_nz(py, ay[0]); red_nz(pz, az[0]); ax[0] = 0.f; ay[0] = 0.f; az[0] = 0.f; break;
    case 1: red_nz(px, ax[1]); red_nz(py, ay[1]); red_nz(pz, az[1]); ax[1] = 0.f; ay[1] = 0.f; az[1] = 0.f; break;
    case 2: red_nz(px, ax[2]); red_nz(py, ay[2]); red_nz(pz, az[2]); ax[2] = 0.f; ay[2] = 0.f; az[2] = 0.f; break;
    default: red_nz(px, ax[3]); red_nz(py, ay[3]); red_nz(pz, az[3]); ax[3] = 0.f; ay[3] = 0.f; az[3] = 0.f; break;
    }
}

// same classification as k_classify_key (particles.cu), on a copy of the position
__device__ __forceinline__ uint32_t sort_key(const DevGeom &G, float x, float y, float z)
{
    int dx = 0, dy = 0, dz = 0;
    if (x < G.minx) dx = -1; else if (x > G.maxx) dx = 1;
    if (y < G.miny) dy = -1; else if (y > G.maxy) dy = 1;
    if (z < G.minz) dz = -1; else if (z > G.maxz) dz = 1;
    bool in = true;
    if (!G.perx) in = (x + G.mxcum > G.x1in) && (x + G.mxcum < G.x2in);
    if (!G.pery && in) in = (y + G.mycum > G.y1in) && (y + G.mycum < G.y2in);
    if (!G.perz && in) in = (z + G.mzcum > G.z1in) && (z + G.mzcum < G.z2in);
    if (!in) return (uint32_t)G.lot + 9u;
    if (dx < 0) x = x + G.shiftx_lo; else if (dx > 0) x = x - G.shiftx_hi;
    if (dy < 0) y = y + G.shifty_lo; else if (dy > 0) y = y - G.shifty_hi;
    if (dz < 0) z = z + G.shiftz_lo; else if (dz > 0) z = z - G.shiftz_hi;
    const int da = G.sendy ? dy : 0, db = G.sendz ? dz : 0;
    const int code = (da + 1) + 3 * (db + 1);
    if (code != 4) return (uint32_t)G.lot + (uint32_t)code;
    int i = min(max((int)x, 1), G.mx), j = min(max((int)y, 1), G.my), k = min(max((int)z, 1), G.mz);
    return (uint32_t)((i - 1) + G.mx * ((j - 1) + G.my * (k - 1)));
}

template <int ORDER, bool FUSED>
__global__ void __launch_bounds__(CR_WARPS * 32, CR_MINB) k_cellrun(CRArgs A)
{
    // dynamic shared memory (CR_SMEM_BYTES > 48 KB): the factor staging of phase 1 -> phase 2, then the record pipeline
    extern __shared__ __align__(16) unsigned char cr_smem[];
    float *stage = reinterpret_cast<float *>(cr_smem);
    uint32_t (*rec)[2][9][32] = reinterpret_cast<uint32_t (*)[2][9][32]>(cr_smem + sizeof(float) * CR_WARPS * CR_WARP_FLOATS);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int half = lane >> 4, hl = lane & 15;
    const long long gw = (long long)blockIdx.x * CR_WARPS + warp;
    const long long base = gw * (2 * CR_CHUNK) + (long long)half * CR_CHUNK;
    if (gw * (2 * CR_CHUNK) >= A.n) return;                  // whole warp idle (no block-level barriers are used)
    const DevGeom &G = A.G;
    const int mx = G.mx, my = G.my;
    const int j = lane & 3, k = (lane >> 2) & 3;
    constexpr bool TILED = FUSED && TGPU_SHADOW_TILED;
    constexpr int PS = TILED ? 16 : 1;                       // distance between consecutive x-planes of a row

    int wi = 0, wrow = 0;                                    // window cell (1-based i, row id)
    bool have = false;
    float ax[4] = {0.f, 0.f, 0.f, 0.f}, ay[4] = {0.f, 0.f, 0.f, 0.f}, az[4] = {0.f, 0.f, 0.f, 0.f};

    // ---- record pipeline: rec[buf][field][lane], fields x y z u v w ch ind tag; each lane only ever touches its own slots
    const bool lazy = FUSED && A.perm != nullptr;
    constexpr int NIT = CR_CHUNK / 16;
    auto fetch = [&](int buf, long long tt, int pp32) {
        if (tt < A.n) {
            const long long pp = lazy ? (long long)pp32 : tt;
            uint32_t *r = &rec[warp][buf][0][lane];
            cp_async4(r + 0 * 32, A.s.x + pp); cp_async4(r + 1 * 32, A.s.y + pp); cp_async4(r + 2 * 32, A.s.z + pp);
            cp_async4(r + 3 * 32, A.s.u + pp); cp_async4(r + 4 * 32, A.s.v + pp); cp_async4(r + 5 * 32, A.s.w + pp);
            cp_async4(r + 6 * 32, A.s.ch + pp);
            if (lazy) { cp_async4(r + 7 * 32, A.s.ind + pp); cp_async4(r + 8 * 32, A.s.tag + pp); }
        }
        cp_async_commit();
    };
    {
        const long long t0 = base + hl;
        fetch(0, t0, (lazy && t0 < A.n) ? A.perm[t0] : 0);
    }
    // permutation entry of this lane's particle of the next step: loaded one step ahead and kept as the raw 32-bit
    // value (no instruction touches it until the following step, so the load never stalls the warp)
    int pnext = 0;
    if (lazy && base + 16 + hl < A.n) pnext = A.perm[base + 16 + hl];

    for (int it = 0; it < NIT; ++it) {
        const long long t = base + it * 16 + hl;
        float *st = stage + warp * CR_WARP_FLOATS + half * CR_HALF_FLOATS + hl * CR_STRIDE;
        int ci = -1, crow = -1;                              // deposit base cell of this lane's particle
        cp_async_wait_all();                                 // this step's record has landed in rec[it & 1]
        if (it + 1 < NIT) {
            fetch((it + 1) & 1, t + 16, pnext);              // next step's record, in flight during this step
            if (lazy && it + 2 < NIT && t + 32 < A.n) pnext = A.perm[t + 32];
        }
        // ------------------------------------------------------------------ phase 1: lane = particle
        if (t < A.n) {
            const uint32_t *r = &rec[warp][it & 1][0][lane];
            float x = __uint_as_float(r[0 * 32]), y = __uint_as_float(r[1 * 32]), z = __uint_as_float(r[2 * 32]);
            float u = __uint_as_float(r[3 * 32]), v = __uint_as_float(r[4 * 32]), w = __uint_as_float(r[5 * 32]);
            const float ch = __uint_as_float(r[6 * 32]);
            if (lazy) {
                // lazily sorted input: the record still carries last lap's unwrapped position; apply the periodic wrap /
                // frame shift its sort key was computed with (deposit_particles loop B), and carry the passive fields along
                if (x < A.G.minx) x += A.G.shiftx_lo; else if (x > A.G.maxx) x -= A.G.shiftx_hi;
                if (y < A.G.miny) y += A.G.shifty_lo; else if (y > A.G.maxy) y -= A.G.shifty_hi;
                if (z < A.G.minz) z += A.G.shiftz_lo; else if (z > A.G.maxz) z -= A.G.shiftz_hi;
                A.d.ch[t] = ch; A.d.ind[t] = (int32_t)r[7 * 32]; A.d.tag[t] = (int32_t)r[8 * 32];
            }
            const float q = ch * A.qs;
            float S1[4], S2[4];
            if (FUSED) {
                const float half_ = 0.5f;
                const int ip = (int)x, jp = (int)y, kp = (int)z;
                const float dxp = x - ip, dyp = y - jp, dzp = z - kp;
                float Wx[4], Wy[4], Wz[4];
                shape_window<ORDER>(dxp, 0, Wx); shape_window<ORDER>(dyp, 0, Wy); shape_window<ORDER>(dzp, 0, Wz);
                float e0 = 0, e1 = 0, e2 = 0, b0 = 0, b1 = 0, b2 = 0;
                // Order 1, and order 2 under quirk Q1 (loop bounds taken from the dual branch,
                // particles_movedeposit.F90:752-790): only slots 3,4 carry weight inside the loop range unless the
                // particle sits exactly on a node, so the sum runs over 2x2x2 nodes (the dropped terms are exact zeros).
                const bool q1 = ORDER == 2 && (G.quirks & TGPU_Q1_MOVER2_RANGE);
                const bool fast = ORDER == 1 || (q1 && dxp != 0.f && dyp != 0.f && dzp != 0.f);
                if (fast) {
                    const float wxs[2] = {Wx[1], Wx[2]}, wys[2] = {Wy[1], Wy[2]}, wzs[2] = {Wz[1], Wz[2]};
                    const int nbase = (ip - 1) + mx * ((jp - 1) + my * (kp - 1));
                    gather_nodes<2>(A.prim8, nbase, mx, my, wxs, wys, wzs, e0, e1, e2, b0, b1, b2);
                } else if (ORDER == 2) {
                    int lox, loy, loz;
                    if (q1) {
                        const float dxd = x - half_ - (int)(x - half_), dyd = y - half_ - (int)(y - half_), dzd = (z - half_) - (int)(z - half_);
                        lox = dxd <= half_ ? 0 : 1; loy = dyd <= half_ ? 0 : 1; loz = dzd <= half_ ? 0 : 1;
                    } else { lox = dxp <= half_ ? 0 : 1; loy = dyp <= half_ ? 0 : 1; loz = dzp <= half_ ? 0 : 1; }
                    float w9[9], eb[6];
#pragma unroll
                    for (int a = 0; a < 3; a++) {
                        w9[a] = lox ? Wx[a + 1] : Wx[a]; w9[3 + a] = loy ? Wy[a + 1] : Wy[a]; w9[6 + a] = loz ? Wz[a + 1] : Wz[a];
                    }
                    const long long nbase = (ip - 2 + lox) + (long long)mx * ((jp - 2 + loy) + (long long)my * (kp - 2 + loz));
                    gather27(A.prim8, nbase, mx, my, w9, eb);
                    e0 = eb[0]; e1 = eb[1]; e2 = eb[2]; b0 = eb[3]; b1 = eb[4]; b2 = eb[5];
                }
                const float cinv = G.cinv, qm = A.qm;
                e0 = 0.5f * e0 * qm; e1 = 0.5f * e1 * qm; e2 = 0.5f * e2 * qm;
                b0 = 0.5f * b0 * qm * cinv; b1 = 0.5f * b1 * qm * cinv; b2 = 0.5f * b2 * qm * cinv;
                if (G.external_fields) {
                    b0 = b0 + G.ext[3] * 0.5f * qm * cinv; b1 = b1 + G.ext[4] * 0.5f * qm * cinv; b2 = b2 + G.ext[5] * 0.5f * qm * cinv;
                    e0 = e0 + G.ext[0] * 0.5f * qm; e1 = e1 + G.ext[1] * 0.5f * qm; e2 = e2 + G.ext[2] * 0.5f * qm;
                }
                push_particle<true>(G.c, G.pusher, e0, e1, e2, b0, b1, b2, x, y, z, u, v, w, cinv);
                A.d.x[t] = x; A.d.y[t] = y; A.d.z[t] = z; A.d.u[t] = u; A.d.v[t] = v; A.d.w[t] = w;
                // The deposit's "old" shape is the gather's shape at the true pre-push position (the reference
                // recomputes it as x - u/gamma*c, particles_movedeposit.F90:1384-1388, equal to round-off).
                // sort key of the pushed particle + its rank inside the destination bin; the rank comes back from L2 while
                // the deposit factors are being staged and is stored at the end of the phase
                const uint32_t ky = sort_key(G, x, y, z);
                A.key[t] = ky;
                const int rank_in_bin = atomicAdd(&A.bincount[ky], 1);
                crow = (jp - 1) | ((kp - 1) << 16); ci = ip;
                shape_window<ORDER>(x - (int)x, (int)x - ip, S2);
                stage_axis<0>(st, Wx, S2, q, 0.f, (ip - 1) & 3);
                shape_window<ORDER>(y - (int)y, (int)y - jp, S2);
                stage_axis<1>(st + 12, Wy, S2, q, __int_as_float(ci), 0);
                shape_window<ORDER>(z - (int)z, (int)z - kp, S2);
                stage_axis<1>(st + 28, Wz, S2, q, __int_as_float(crow), 0);
                A.slot[t] = rank_in_bin;
            } else {
                // deposit_particles loop A: old position recomputed from the new one (particles_movedeposit.F90:1384-1390)
                const float invgam = 1.f / sqrtf(1 + u * u + v * v + w * w);
                const float x1 = x - u * invgam * G.c, y1 = y - v * invgam * G.c, z1 = z - w * invgam * G.c;
                const int i1 = (int)x1, j1 = (int)y1, k1 = (int)z1;
                crow = (j1 - 1) | ((k1 - 1) << 16); ci = i1;
                shape_window<ORDER>(x1 - i1, 0, S1); shape_window<ORDER>(x - (int)x, (int)x - i1, S2);
                stage_axis<0>(st, S1, S2, q, 0.f, (i1 - 1) & 3);
                shape_window<ORDER>(y1 - j1, 0, S1); shape_window<ORDER>(y - (int)y, (int)y - j1, S2);
                stage_axis<1>(st + 12, S1, S2, q, __int_as_float(ci), 0);
                shape_window<ORDER>(z1 - k1, 0, S1); shape_window<ORDER>(z - (int)z, (int)z - k1, S2);
                stage_axis<1>(st + 28, S1, S2, q, __int_as_float(crow), 0);
            }
        }
        else {
            // past the end (last warp only): stage zeros so that phase 2 can always walk all 16 slots of the half
#pragma unroll
            for (int q4 = 0; q4 < CR_STRIDE / 4; q4++) *(float4 *)(st + 4 * q4) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        // run starts inside each half: a particle whose cell differs from its predecessor's
        const int pci = __shfl_up_sync(0xffffffffu, ci, 1), pcrow = __shfl_up_sync(0xffffffffu, crow, 1);
        const bool start = t < A.n && (hl == 0 || ci != pci || crow != pcrow);
        unsigned starts = (__ballot_sync(0xffffffffu, start) >> (half << 4)) & 0xFFFFu;
        __syncwarp();                                        // staging written by all lanes is visible to the warp
        // ------------------------------------------------------------------ phase 2: half-warp = footprint
        // Both halves walk their 16 particles in lockstep; only the (rare) window moves diverge.
        const float4 *sp = (const float4 *)(stage + warp * CR_WARP_FLOATS + half * CR_HALF_FLOATS);
CR_DO_PRAGMA(unroll CR_UNROLL)
        for (int tt = 0; tt < 16; ++tt, sp += CR_STRIDE / 4) {
            if ((starts >> tt) & 1u) {
                // particle tt opens a run of particles that share one footprint
                const int ni = __float_as_int(sp[3].w), nrow = __float_as_int(sp[7].w);
                if (!have || ni != wi || nrow != wrow) {
                    if (have) {
                        const size_t idx0 = row_index<TILED>(mx, my, A.nty, (wrow & 0xFFFF) + j - 1, (wrow >> 16) + k - 1, wi - 2);
                        const int di = ni - wi;
                        // sliding 1..3 cells along x completes that many planes; anything else flushes the window
                        // sliding 1..3 cells along x completes that many planes; anything else flushes the window
                        if (nrow == wrow && di == 1) flush_one(A.cx, A.cy, A.cz, idx0, wi, ax, ay, az);
                        else flush_planes<PS>(A.cx, A.cy, A.cz, idx0, wi, (nrow == wrow && di > 1 && di < 4) ? di : 4, ax, ay, az);
                    }
                    wi = ni; wrow = nrow; have = true;
                }
            }
            const float4 yv = sp[3 + j], zv = sp[7 + k];
            const float4 qpsx = sp[0], xa = sp[1], xb = sp[2];
            const float sy1 = yv.x, dsy = yv.y, qpsy = yv.z;
            const float sz1 = zv.x, dsz = zv.y, qpsz = zv.z;
            const float ya = fmaf(0.5f, dsy, sy1), yb = fmaf(1.f / 3.f, dsy, 0.5f * sy1);
            const float wx = fmaf(yb, dsz, ya * sz1);      // Wx(j,k)
            const float a = qpsy * sz1, b = qpsy * dsz;    // Jy = XA*a + XB*b
            const float c = qpsz * sy1, d = qpsz * dsy;    // Jz = XA*c + XB*d
#if CR_FFMA2
            // sm_100 packed fp32 (FFMA2, fma.rn.f32x2): the x-ring registers pair up as (0,1) and (2,3), the per-lane
            // factors are broadcast into both halves of a 64-bit register; 10 issue slots instead of 20
            {
                const float2 wx2 = make_float2(wx, wx), a2 = make_float2(a, a), b2 = make_float2(b, b),
                             c2 = make_float2(c, c), d2 = make_float2(d, d);
                const float2 px01 = make_float2(qpsx.x, qpsx.y), px23 = make_float2(qpsx.z, qpsx.w);
                const float2 xa01 = make_float2(xa.x, xa.y), xa23 = make_float2(xa.z, xa.w);
                const float2 xb01 = make_float2(xb.x, xb.y), xb23 = make_float2(xb.z, xb.w);
                float2 t;
                t = __ffma2_rn(px01, wx2, make_float2(ax[0], ax[1])); ax[0] = t.x; ax[1] = t.y;
                t = __ffma2_rn(px23, wx2, make_float2(ax[2], ax[3])); ax[2] = t.x; ax[3] = t.y;
                t = __ffma2_rn(xa01, a2, __ffma2_rn(xb01, b2, make_float2(ay[0], ay[1]))); ay[0] = t.x; ay[1] = t.y;
                t = __ffma2_rn(xa23, a2, __ffma2_rn(xb23, b2, make_float2(ay[2], ay[3]))); ay[2] = t.x; ay[3] = t.y;
                t = __ffma2_rn(xa01, c2, __ffma2_rn(xb01, d2, make_float2(az[0], az[1]))); az[0] = t.x; az[1] = t.y;
                t = __ffma2_rn(xa23, c2, __ffma2_rn(xb23, d2, make_float2(az[2], az[3]))); az[2] = t.x; az[3] = t.y;
            }
#else
            ax[0] = fmaf(qpsx.x, wx, ax[0]); ax[1] = fmaf(qpsx.y, wx, ax[1]);
            ax[2] = fmaf(qpsx.z, wx, ax[2]); ax[3] = fmaf(qpsx.w, wx, ax[3]);
            ay[0] = fmaf(xa.x, a, fmaf(xb.x, b, ay[0])); ay[1] = fmaf(xa.y, a, fmaf(xb.y, b, ay[1]));
            ay[2] = fmaf(xa.z, a, fmaf(xb.z, b, ay[2])); ay[3] = fmaf(xa.w, a, fmaf(xb.w, b, ay[3]));
            az[0] = fmaf(xa.x, c, fmaf(xb.x, d, az[0])); az[1] = fmaf(xa.y, c, fmaf(xb.y, d, az[1]));
            az[2] = fmaf(xa.z, c, fmaf(xb.z, d, az[2])); az[3] = fmaf(xa.w, c, fmaf(xb.w, d, az[3]));
#endif
        }
        __syncwarp();
    }
    if (have) {
        const size_t idx0 = row_index<TILED>(mx, my, A.nty, (wrow & 0xFFFF) + j - 1, (wrow >> 16) + k - 1, wi - 2);
        flush_planes<PS>(A.cx, A.cy, A.cz, idx0, wi, 4, ax, ay, az);
    }
}

int cellrun_supported(const tgpu_ctx *h)
{
    return h->P.dim == 3 && (h->P.order == 1 || h->P.order == 2) && h->G.lot < (1ll << 30) && h->P.my < 65536 && h->P.mz < 32768;
}

template <bool FUSED>
static int launch(tgpu_ctx *h, float *cx, float *cy, float *cz)
{
    for (int s = 0; s < 2; s++) {
        Species &S = h->sp[s];
        if (S.n == 0) continue;
        CRArgs A;
        A.s = S; A.d = S; A.perm = nullptr;
        A.n = S.n; A.prim8 = h->prim8; A.cx = cx; A.cy = cy; A.cz = cz; A.nty = h->nty; A.G = h->G;
        A.qm = s ? h->P.qme : h->P.qmi; A.qs = s ? h->P.qe : h->P.qi;
        const size_t nb = (size_t)h->G.lot + TGPU_NBIN_EXTRA;
        A.key = h->key[s]; A.slot = h->slot + (size_t)s * h->maxhlf; A.bincount = h->bincount + (size_t)s * nb;
        if (FUSED) CK(cudaMemsetAsync(A.bincount, 0, nb * sizeof(int32_t), h->stream));
        if (FUSED && h->lazy[s]) { A.perm = h->perm[s]; A.d = h->alt[s]; }     // gather through the pending permutation
        long long warps = (S.n + 2 * CR_CHUNK - 1) / (2 * CR_CHUNK);
        int blocks = (int)((warps + CR_WARPS - 1) / CR_WARPS);
        if (h->P.order == 2) {
            CK(cudaFuncSetAttribute(k_cellrun<2, FUSED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CR_SMEM_BYTES));
            k_cellrun<2, FUSED><<<blocks, CR_WARPS * 32, CR_SMEM_BYTES, h->stream>>>(A);
        } else {
            CK(cudaFuncSetAttribute(k_cellrun<1, FUSED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CR_SMEM_BYTES));
            k_cellrun<1, FUSED><<<blocks, CR_WARPS * 32, CR_SMEM_BYTES, h->stream>>>(A);
        }
        CKK(h);
        if (FUSED && h->lazy[s]) {
            // the pushed records now sit, in sorted order and wrapped, in the other buffer
            int n = S.n;
            Species tmp = h->sp[s]; h->sp[s] = h->alt[s]; h->alt[s] = tmp;
            h->sp[s].n = n; h->lazy[s] = 0; h->nphys[s] = n;
        }
    }
    return 0;
}

// tgpu_move_particles fast path: gather + push + deposit in one pass over the particles.  The currents go to shadow[]
// because mainloop resets cur between move_particles and deposit_particles (tristanmainloop.F90:134-183).
int cellrun_move_deposit(tgpu_ctx *h)
{
    int rc = fld_primal(h); if (rc) return rc;
    if (h->presort) {
        // Freshly uploaded records are in the host's order.  The cell-run deposit is correct for any order but a particle
        // that does not share its cell with its predecessor costs a full window flush (measured: 85 ms instead of 8 ms per
        // launch on a randomly ordered load), so the records are counting-sorted once (~8 ms) before the first fused mover.
        // At this point of the lap every particle is inside the rank, so the sort only orders them.
        h->presort = 0;
        rc = prt_sort(h, false); if (rc) return rc;
    }
    rc = launch<true>(h, h->shadow[0], h->shadow[1], h->shadow[2]); if (rc) return rc;
    h->keys_valid = 1;                       // prt_sort may skip its classify pass
    return 0;
}

// Streamed mirror lap (api.cu tgpu_step_mirror): fused gather + push + deposit of the records [off, off + cnt) of species s,
// in place.  The caller has refreshed the node-centred fields (fld_primal) and zeroed nothing: key / slot / bincount are
// written as in a resident lap but not used.
int cellrun_move_deposit_range(tgpu_ctx *h, int s, int off, int cnt)
{
    if (cnt <= 0) return 0;
    const Species &S = h->sp[s];
    CRArgs A;
    Species R = S;
    R.x += off; R.y += off; R.z += off; R.u += off; R.v += off; R.w += off; R.ch += off; R.ind += off; R.tag += off; R.n = cnt;
    A.s = R; A.d = R; A.perm = nullptr;
    A.n = cnt; A.prim8 = h->prim8; A.cx = h->shadow[0]; A.cy = h->shadow[1]; A.cz = h->shadow[2]; A.nty = h->nty; A.G = h->G;
    A.qm = s ? h->P.qme : h->P.qmi; A.qs = s ? h->P.qe : h->P.qi;
    const size_t nb = (size_t)h->G.lot + TGPU_NBIN_EXTRA;
    A.key = h->key[s] + off; A.slot = h->slot + (size_t)s * h->maxhlf + off; A.bincount = h->bincount + (size_t)s * nb;
    long long warps = ((long long)cnt + 2 * CR_CHUNK - 1) / (2 * CR_CHUNK);
    int blocks = (int)((warps + CR_WARPS - 1) / CR_WARPS);
    if (h->P.order == 2) {
        CK(cudaFuncSetAttribute(k_cellrun<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CR_SMEM_BYTES));
        k_cellrun<2, true><<<blocks, CR_WARPS * 32, CR_SMEM_BYTES, h->stream>>>(A);
    } else {
        CK(cudaFuncSetAttribute(k_cellrun<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CR_SMEM_BYTES));
        k_cellrun<1, true><<<blocks, CR_WARPS * 32, CR_SMEM_BYTES, h->stream>>>(A);
    }
    CKK(h);
    return 0;
}

// tgpu_deposit_particles fast path when the particles were moved elsewhere (mirror mode, tests)
int cellrun_deposit(tgpu_ctx *h)
{
    { int rc = prt_materialize(h); if (rc) return rc; }
    return launch<false>(h, h->f[6], h->f[7], h->f[8]);
}
