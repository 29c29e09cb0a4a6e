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
//   phase 1 (lane = particle, 32 particles per warp step): read the landed record, [gather node-centred fields (packed
//           FFMA2) + Boris push + store + sort key of the new cell + rank in its bin], build the 1-D factors of the
//           Esirkepov sum from the old and new shapes and park them in shared memory (40 floats per particle).
//           The deposit's old position is recomputed from the new one as the reference does (x - u/gamma*c);
//   phase 2 (QUARTER-warp = one footprint): lane (j, kp) of a quarter owns the 4 x-cells of the two footprint rows
//           (j, kp) and (j, kp + 2) for all three components = 24 register accumulators (12 FFMA2 pairs).  It walks the 8
//           particles its quarter staged; the four quarters run in lockstep.  The x-planes form a ring (plane of cell x
//           lives in register x mod 4, phase 1 stores the x factors pre-rotated): when the cell advances along x the
//           completed plane is flushed with predicated fp32 REDs into the tiled shadow arrays (tgpu_internal.h
//           row_index) and cleared; nothing moves between registers.
// Why a quarter-warp: the factors reach phase 2 through shared memory, and profiles/micro/lds_rate.cu measures what that
// costs on B200 -- an LDS.128 takes 2 LSU cycles per warp when consecutive lanes share addresses and 4 otherwise (an
// LDS.32 takes 1): the data path delivers 8 bytes per lane per cycle whatever the broadcast.  With 16 lanes per
// footprint every particle cost 6 LSU cycles (of the ~14 the whole kernel spent per particle, 78 % of the pipe); with 8
// lanes owning two rows each the x factors are loaded once for twice the outputs: 13 cycles per 4 particles.  The same
// move cuts the per-lane preparation (9 scalar FP per 12 outputs -> 3 scalar + 6 packed per 24 outputs).
// With ~8 particles per cell and species, that is ~6 global REDs per particle instead of up to 192 atomics, and the
// arithmetic is the factorised form of Appendix A.3 (Jx = q*Wx(j,k)*prefix_i(dSx), ...).
// Nothing here depends on the particles being sorted for correctness -- an unsorted tail (fresh arrivals) only makes
// the window jump and flush more often (which is why freshly uploaded records are sorted once, cellrun_move_deposit).
#include "tgpu_internal.h"
#include "shapes.cuh"
#include "cellrun_common.cuh"

#ifndef CR_WARPS
#define CR_WARPS 4           // 4-warp blocks x 6 per SM (measured: 15.64 ms per lap against 15.94 for 8 x 3)
#endif
#ifndef CR_MINB
#define CR_MINB 6
#endif
#ifndef CR_CHUNK
#define CR_CHUNK 128          // particles per quarter-warp
#endif
#ifndef CR_P2UNROLL
#define CR_P2UNROLL 4         // phase-2 unroll: the window-move code is inlined once per copy and the kernel has to stay inside
                              // the instruction cache (fully unrolled: 4096 instructions, 29 % of the stall samples "no instruction")
#endif
#define CR_STR(x) #x
#define CR_DO_PRAGMA(x) _Pragma(CR_STR(x))
#ifndef CR_OLDPOS_REF
#define CR_OLDPOS_REF 1       // fused launches recompute the deposit's old position as the reference does (0: reuse the gather's shape)
#endif
#ifndef CR_PREFETCH
#define CR_PREFETCH 0
#endif
#define CR_STRIDE 44          // floats between the staged factors of consecutive particles (40 used; 44 keeps the
                              // phase-1 STS.128 of a quarter-warp conflict-free: 12 banks apart)
#define CR_QFLOATS (8 * CR_STRIDE + 8)     // one quarter's staging; the four quarters start 8 banks apart so that their
                                           // lockstep phase-2 loads never fall into the same banks
#define CR_WARP_FLOATS (4 * CR_QFLOATS)
#define CR_REC_WORDS (2 * 9 * 32)          // record landing slots of one warp: [buf][field][lane]
#ifndef CR_REGPF
#define CR_REGPF 0            // 1: the next step's record is prefetched into registers (streaming loads that do not allocate
                              // in L1); 0: 4-byte cp.async into per-lane shared-memory landing slots
#endif
#define CR_SMEM_BYTES (sizeof(float) * CR_WARPS * CR_WARP_FLOATS + (CR_REGPF ? 0 : sizeof(uint32_t) * CR_WARPS * CR_REC_WORDS))
#define CR_NOWIN (-0x40000000)             // "no window yet"

// staging layout of one particle (floats):
//    0.. 3  q * prefix(dSx)   \  rotated by (i1 - 1) & 3: component m belongs to the footprint cell whose x index is
//    4.. 7  XA = S1 + dS/2     > congruent to m mod 4, so the phase-2 accumulators never move when the window slides
//    8..11  XB = S1/2 + dS/3  /
//   12..15  Sy1[0..3]    16..19  dSy[0..3]    20..23  q * prefix(dSy)[0..3]          (read as scalars: lane's row j)
//   24..27  (Sz1[0], Sz1[2], dSz[0], dSz[2])   28..31  (Sz1[1], Sz1[3], dSz[1], dSz[3])     row pair kp = 0 / 1
//   32..35  (qPz[0], qPz[2], cell i, row id)   36..39  (qPz[1], qPz[3], cell i, row id)

struct CRArgs {
    Species s;                // source records
    Species d;                // FUSED: destination records (logical order); may alias s when perm == nullptr
    const int32_t *perm;      // LAZY: logical position t reads physical record perm[t] (lazy sort)
    unsigned n;
    const float4 *prim8;
    float *cx, *cy, *cz;     // FUSED: the tiled shadow arrays (see row_index), else curx, cury, curz in Fortran order
    int nty;                  // FUSED: number of 4-row tiles along y
    DevGeom G;
    float qm, qs;
    uint32_t *key;            // FUSED: sort key of the pushed particle (prt_sort skips its classify pass)
    int32_t *slot, *bincount;
    unsigned keyoff;          // 1 + mx + mx*my: key = i + mx*(j + my*k) - keyoff for 1-based cell indices
    int general;              // any open axis or any split axis: the key needs the slow classification
};

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

// rotate a 4-vector left: component (s + r) & 3 of the result holds v[s]
__device__ __forceinline__ float4 rot4(float v0, float v1, float v2, float v3, int r)
{
    const bool r2 = (r & 2) != 0, r1 = (r & 1) != 0;
    const float a0 = r2 ? v2 : v0, a1 = r2 ? v3 : v1, a2 = r2 ? v0 : v2, a3 = r2 ? v1 : v3;           // by 2
    return r1 ? make_float4(a3, a0, a1, a2) : make_float4(a0, a1, a2, a3);                            // by 1
}

// 1-D factors of the three axes -> shared staging (layout above)
__device__ __forceinline__ void stage_x(float *st, const float S1[4], const float S2[4], float q, int r)
{
    const float third = 1.f / 3.f;
    const float d0 = S2[0] - S1[0], d1 = S2[1] - S1[1], d2 = S2[2] - S1[2], d3 = S2[3] - S1[3];
    const float p0 = d0, p1 = p0 + d1, p2 = p1 + d2, p3 = p2 + d3;
    *(float4 *)(st + 0) = rot4(q * p0, q * p1, q * p2, q * p3, r);
    *(float4 *)(st + 4) = rot4(fmaf(0.5f, d0, S1[0]), fmaf(0.5f, d1, S1[1]), fmaf(0.5f, d2, S1[2]), fmaf(0.5f, d3, S1[3]), r);
    *(float4 *)(st + 8) = rot4(fmaf(third, d0, 0.5f * S1[0]), fmaf(third, d1, 0.5f * S1[1]),
                               fmaf(third, d2, 0.5f * S1[2]), fmaf(third, d3, 0.5f * S1[3]), r);
}
__device__ __forceinline__ void stage_y(float *st, const float S1[4], const float S2[4], float q)
{
    const float d0 = S2[0] - S1[0], d1 = S2[1] - S1[1], d2 = S2[2] - S1[2], d3 = S2[3] - S1[3];
    const float p0 = d0, p1 = p0 + d1, p2 = p1 + d2, p3 = p2 + d3;
    *(float4 *)(st + 12) = make_float4(S1[0], S1[1], S1[2], S1[3]);
    *(float4 *)(st + 16) = make_float4(d0, d1, d2, d3);
    *(float4 *)(st + 20) = make_float4(q * p0, q * p1, q * p2, q * p3);
}
__device__ __forceinline__ void stage_z(float *st, const float S1[4], const float S2[4], float q, int ci, int crow)
{
    const float d0 = S2[0] - S1[0], d1 = S2[1] - S1[1], d2 = S2[2] - S1[2], d3 = S2[3] - S1[3];
    const float p0 = d0, p1 = p0 + d1, p2 = p1 + d2, p3 = p2 + d3;
    const float fi = __int_as_float(ci), fr = __int_as_float(crow);
    *(float4 *)(st + 24) = make_float4(S1[0], S1[2], d0, d2);
    *(float4 *)(st + 28) = make_float4(S1[1], S1[3], d1, d3);
    *(float4 *)(st + 32) = make_float4(q * p0, q * p2, fi, fr);
    *(float4 *)(st + 36) = make_float4(q * p1, q * p3, fi, fr);
}

template <int ORDER, bool FUSED, bool LAZY, bool FASTP>
__global__ void __launch_bounds__(CR_WARPS * 32, CR_MINB) k_cellrun(const CRArgs A)
{
    // dynamic shared memory (CR_SMEM_BYTES > 48 KB): the factor staging of phase 1 -> phase 2, then the record pipeline
    extern __shared__ __align__(16) unsigned char cr_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int qtr = lane >> 3, ql = lane & 7;
    const unsigned n = A.n;
    const unsigned wbase = (blockIdx.x * CR_WARPS + warp) * (4u * CR_CHUNK);
    if (wbase >= n) return;                                  // whole warp idle (no block-level barriers are used)
    const unsigned base = wbase + qtr * CR_CHUNK;
    const DevGeom &G = A.G;
    const int mx = G.mx, my = G.my;
    const int j = ql & 3, kp = ql >> 2;                      // this lane's footprint rows: (j, kp) and (j, kp + 2)
    constexpr bool TILED = FUSED && TGPU_SHADOW_TILED;
    constexpr unsigned PS = TILED ? 16 : 1;                  // distance between consecutive x-planes of a row

    float *const sq = reinterpret_cast<float *>(cr_smem) + warp * CR_WARP_FLOATS + qtr * CR_QFLOATS;   // quarter's staging
    float *const st = sq + ql * CR_STRIDE;                   // phase 1: this lane's particle
    const float *const syl = sq + 12 + j;                    // phase 2: row j of the y block
    const float4 *const szl = reinterpret_cast<const float4 *>(sq + 24 + 4 * kp);   // phase 2: row pair kp of the z block
#if !CR_REGPF
    uint32_t *const rec = reinterpret_cast<uint32_t *>(cr_smem + sizeof(float) * CR_WARPS * CR_WARP_FLOATS) + warp * CR_REC_WORDS + lane;
#endif

    // window: cell wi (1-based), row id wrow = (j0 - 1) | (k0 - 1) << 16, element offsets of the lane's two rows at plane wi - 1
    int wi = CR_NOWIN, wrow = -1;
    unsigned woffA = 0, woffB = 0;
    float2 axA01 = make_float2(0.f, 0.f), axA23 = axA01, ayA01 = axA01, ayA23 = axA01, azA01 = axA01, azA23 = axA01;
    float2 axB01 = axA01, axB23 = axA01, ayB01 = axA01, ayB23 = axA01, azB01 = axA01, azB23 = axA01;

    // flush the lowest plane of the window (cell wi - 1, ring register (wi - 1) & 3), clear it and slide the window by one
    // cell.  The ring register is uniform across the quarter, so the 4-way switch costs one short divergent body.
    // (order 1 leaves half of the 4-wide window exactly zero: skip those; order 2 fills it, the test would only cost issue slots)
    auto red = [&](float *p, float v) { if (ORDER == 1) red_nz(p, v); else red_add(p, v); };
    auto flush_slide = [&]() {
        float *pxA = A.cx + woffA, *pyA = A.cy + woffA, *pzA = A.cz + woffA;
        float *pxB = A.cx + woffB, *pyB = A.cy + woffB, *pzB = A.cz + woffB;
        switch ((wi - 1) & 3) {
        case 0: red(pxA, axA01.x); red(pyA, ayA01.x); red(pzA, azA01.x); red(pxB, axB01.x); red(pyB, ayB01.x); red(pzB, azB01.x);
                axA01.x = ayA01.x = azA01.x = axB01.x = ayB01.x = azB01.x = 0.f; break;
        case 1: red(pxA, axA01.y); red(pyA, ayA01.y); red(pzA, azA01.y); red(pxB, axB01.y); red(pyB, ayB01.y); red(pzB, azB01.y);
                axA01.y = ayA01.y = azA01.y = axB01.y = ayB01.y = azB01.y = 0.f; break;
        case 2: red(pxA, axA23.x); red(pyA, ayA23.x); red(pzA, azA23.x); red(pxB, axB23.x); red(pyB, ayB23.x); red(pzB, azB23.x);
                axA23.x = ayA23.x = azA23.x = axB23.x = ayB23.x = azB23.x = 0.f; break;
        default: red(pxA, axA23.y); red(pyA, ayA23.y); red(pzA, azA23.y); red(pxB, axB23.y); red(pyB, ayB23.y); red(pzB, azB23.y);
                axA23.y = ayA23.y = azA23.y = axB23.y = ayB23.y = azB23.y = 0.f; break;
        }
        wi += 1; woffA += PS; woffB += PS;
    };
    // the particle that opens a run sits in cell ni of row nrow: bring the window there
    auto move_window = [&](int ni, int nrow) {
        const int di = ni - wi;
        const bool slide = nrow == wrow && (unsigned)di < 4u;     // 0: a run continued from the last step; 1: the common case
        // sliding 1..3 cells along x completes that many planes; anything else flushes the whole window (one inlined copy of
        // the flush: the code of this kernel has to stay inside the instruction cache)
        const int ns = slide ? di : (wi != CR_NOWIN ? 4 : 0);
#pragma unroll 1
        for (int s = 0; s < ns; s++) flush_slide();
        if (!slide) {
            wi = ni; wrow = nrow;
            const int J = (nrow & 0xFFFF) + j - 1, K = (nrow >> 16) + kp - 1;
            woffA = (unsigned)row_index<TILED>(mx, my, A.nty, J, K, ni - 2);
            woffB = (unsigned)row_index<TILED>(mx, my, A.nty, J, K + 2, ni - 2);
        }
    };

    // ---- record pipeline.  CR_REGPF: seven streaming loads per lane one step ahead, straight into registers (nothing touches
    // them until the next step starts, so they are in flight during this step's phase 1 and phase 2; ind / tag are only
    // copied and are loaded at the start of their own step).  Otherwise: rec[buf][field][lane] landing slots filled by
    // 4-byte cp.async, fields x y z u v w ch ind tag; each lane only ever touches its own slots.
    constexpr int NIT = CR_CHUNK / 8;
#if CR_REGPF
    float nx_ = 0.f, ny_ = 0.f, nz_ = 0.f, nu_ = 0.f, nv_ = 0.f, nw_ = 0.f, nch_ = 0.f;
    unsigned pcur = 0;
    auto fetch = [&](int, unsigned tt, int pp32) {
        if (tt < n) {
            const unsigned pp = LAZY ? (unsigned)pp32 : tt;
            nx_ = ldg_stream(A.s.x + pp); ny_ = ldg_stream(A.s.y + pp); nz_ = ldg_stream(A.s.z + pp);
            nu_ = ldg_stream(A.s.u + pp); nv_ = ldg_stream(A.s.v + pp); nw_ = ldg_stream(A.s.w + pp);
            nch_ = ldg_stream(A.s.ch + pp);
        }
    };
#else
    auto fetch = [&](int buf, unsigned tt, int pp32) {
        if (tt < n) {
            const unsigned pp = LAZY ? (unsigned)pp32 : tt;
            uint32_t *r = rec + buf * (9 * 32);
            cp_async4(r + 0 * 32, A.s.x + pp); cp_async4(r + 1 * 32, A.s.y + pp); cp_async4(r + 2 * 32, A.s.z + pp);
            cp_async4(r + 3 * 32, A.s.u + pp); cp_async4(r + 4 * 32, A.s.v + pp); cp_async4(r + 5 * 32, A.s.w + pp);
            cp_async4(r + 6 * 32, A.s.ch + pp);
            if (LAZY) { cp_async4(r + 7 * 32, A.s.ind + pp); cp_async4(r + 8 * 32, A.s.tag + pp); }
        }
        cp_async_commit();
    };
#endif
    // permutation entry of this lane's particle of the next step: loaded one step ahead and kept as the raw 32-bit
    // value (no instruction touches it until the following step, so the load never stalls the warp)
    int pnext = 0;
    {
        const unsigned t0 = base + ql;
        const int p0 = (LAZY && t0 < n) ? __ldcs(A.perm + t0) : 0;
        fetch(0, t0, p0);
        if (LAZY && base + 8 + ql < n) pnext = __ldcs(A.perm + base + 8 + ql);
#if CR_REGPF
        pcur = LAZY ? (unsigned)p0 : t0;
#endif
    }

#pragma unroll 1
    for (int it = 0; it < NIT; ++it) {
        const unsigned t = base + it * 8 + ql;
        int ci = -1, crow = -1;                              // deposit base cell of this lane's particle
#if CR_REGPF
        float x = nx_, y = ny_, z = nz_, u = nu_, v = nv_, w = nw_;
        const float ch = nch_;
        int pind = 0, ptag = 0;
        if (LAZY && t < n) { pind = ldg_stream(A.s.ind + pcur); ptag = ldg_stream(A.s.tag + pcur); }
        if (it + 1 < NIT) {
            pcur = LAZY ? (unsigned)pnext : t + 8;
            fetch(0, t + 8, pnext);                          // next step's record, in flight during this step
            if (LAZY && it + 2 < NIT && t + 16 < n) pnext = __ldcs(A.perm + t + 16);
        }
#else
        cp_async_wait_all();                                 // this step's record has landed in rec[it & 1]
        if (it + 1 < NIT) {
            fetch((it + 1) & 1, t + 8, pnext);               // next step's record, in flight during this step
            if (LAZY && it + 2 < NIT && t + 16 < n) pnext = __ldcs(A.perm + t + 16);
        }
#endif
        // ------------------------------------------------------------------ phase 1: lane = particle
        if (t < n) {
#if !CR_REGPF
            const uint32_t *r = rec + (it & 1) * (9 * 32);
            float x = __uint_as_float(r[0 * 32]), y = __uint_as_float(r[1 * 32]), z = __uint_as_float(r[2 * 32]);
            float u = __uint_as_float(r[3 * 32]), v = __uint_as_float(r[4 * 32]), w = __uint_as_float(r[5 * 32]);
            const float ch = __uint_as_float(r[6 * 32]);
            const int pind = LAZY ? (int)r[7 * 32] : 0, ptag = LAZY ? (int)r[8 * 32] : 0;
#endif
            if (LAZY) {
                // lazily sorted input: the record still carries last lap's unwrapped position; apply the periodic wrap /
                // frame shift its sort key was computed with (deposit_particles loop B), and carry the passive fields along
                bool lo, hi;
                x = wrap1(x, G.minx, G.maxx, G.shiftx_lo, G.shiftx_hi, lo, hi);
                y = wrap1(y, G.miny, G.maxy, G.shifty_lo, G.shifty_hi, lo, hi);
                z = wrap1(z, G.minz, G.maxz, G.shiftz_lo, G.shiftz_hi, lo, hi);
                A.d.ch[t] = ch; A.d.ind[t] = pind; A.d.tag[t] = ptag;
            }
            const float q = ch * A.qs;
            float S1[4], S2[4];
            if (FUSED) {
                const float half_ = 0.5f;
                const int ip = (int)x, jp = (int)y, kq = (int)z;
                const float dxp = x - ip, dyp = y - jp, dzp = z - kq;
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
                    const int nbase = (ip - 1) + mx * ((jp - 1) + my * (kq - 1));
#if CR_PREFETCH
                    // In sorted order the next step of this quarter works one or two cells further along x: its new x-node
                    // column (2 x 2 rows) is not in L1 yet.  One prefetch per lane -- rows by (ql & 1, ql >> 1 & 1), nodes
                    // ip + 2 (lanes 0-3) and ip + 3 (lanes 4-7) -- brings it in while phase 2 runs.
                    {
                        const unsigned pn = (unsigned)(nbase + 2 + (ql >> 2) + mx * ((ql & 1) + my * ((ql >> 1) & 1)));
                        prefetch_l1_keep(A.prim8 + 2u * min(pn, (unsigned)G.lot - 1u));
                    }
#endif
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
                    const long long nbase = (ip - 2 + lox) + (long long)mx * ((jp - 2 + loy) + (long long)my * (kq - 2 + loz));
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
                push_particle<FASTP>(G.c, G.pusher, e0, e1, e2, b0, b1, b2, x, y, z, u, v, w, cinv);
                A.d.x[t] = x; A.d.y[t] = y; A.d.z[t] = z; A.d.u[t] = u; A.d.v[t] = v; A.d.w[t] = w;
                // sort key of the pushed particle + its rank inside the destination bin; the rank comes back from L2 while
                // the deposit factors are being staged and is stored at the end of the phase
                const uint32_t ky = sort_key(G, A.keyoff, A.general, x, y, z);
                A.key[t] = ky;
                const int rank_in_bin = atomicAdd(&A.bincount[ky], 1);
#if CR_OLDPOS_REF
                // deposit_particles loop A: the old position is RECOMPUTED from the new one, x - u/gamma*c
                // (particles_movedeposit.F90:1384-1390), not remembered -- at x ~ 130 one ulp of x is 7e-5 of a step, so the
                // gather's shape at the true pre-push position would give currents that differ from the reference's by
                // ~1e-5 of the gross current.  
                {
                    const float invgam = FASTP ? rsqrtf(1.f + u * u + v * v + w * w) : 1.f / sqrtf(1.f + u * u + v * v + w * w);
                    const float x1 = x - u * invgam * G.c, y1 = y - v * invgam * G.c, z1 = z - w * invgam * G.c;
                    const int i1 = (int)x1, j1 = (int)y1, k1 = (int)z1;
                    crow = (j1 - 1) | ((k1 - 1) << 16); ci = i1;
                    shape_window<ORDER>(x1 - i1, 0, S1); shape_window<ORDER>(x - (int)x, (int)x - i1, S2);
                    stage_x(st, S1, S2, q, (i1 - 1) & 3);
                    shape_window<ORDER>(y1 - j1, 0, S1); shape_window<ORDER>(y - (int)y, (int)y - j1, S2);
                    stage_y(st, S1, S2, q);
                    shape_window<ORDER>(z1 - k1, 0, S1); shape_window<ORDER>(z - (int)z, (int)z - k1, S2);
                    stage_z(st, S1, S2, q, ci, crow);
                }
#else
                // The deposit's "old" shape is the gather's shape at the true pre-push position (the reference
                // recomputes it as x - u/gamma*c, particles_movedeposit.F90:1384-1388, equal to round-off).
                crow = (jp - 1) | ((kq - 1) << 16); ci = ip;
                const int in_ = (int)x, jn = (int)y, kn = (int)z;
                shape_window<ORDER>(x - in_, in_ - ip, S2);
                stage_x(st, Wx, S2, q, (ip - 1) & 3);
                shape_window<ORDER>(y - jn, jn - jp, S2);
                stage_y(st, Wy, S2, q);
                shape_window<ORDER>(z - kn, kn - kq, S2);
                stage_z(st, Wz, S2, q, ci, crow);
#endif
                A.slot[t] = rank_in_bin;
            } else {
                // deposit_particles loop A: old position recomputed from the new one (particles_movedeposit.F90:1384-1390)
                const float invgam = 1.f / sqrtf(1 + u * u + v * v + w * w);
                const float x1 = x - u * invgam * G.c, y1 = y - v * invgam * G.c, z1 = z - w * invgam * G.c;
                const int i1 = (int)x1, j1 = (int)y1, k1 = (int)z1;
                crow = (j1 - 1) | ((k1 - 1) << 16); ci = i1;
                shape_window<ORDER>(x1 - i1, 0, S1); shape_window<ORDER>(x - (int)x, (int)x - i1, S2);
                stage_x(st, S1, S2, q, (i1 - 1) & 3);
                shape_window<ORDER>(y1 - j1, 0, S1); shape_window<ORDER>(y - (int)y, (int)y - j1, S2);
                stage_y(st, S1, S2, q);
                shape_window<ORDER>(z1 - k1, 0, S1); shape_window<ORDER>(z - (int)z, (int)z - k1, S2);
                stage_z(st, S1, S2, q, ci, crow);
            }
        }
        else {
            // past the end (last warp only): stage zeros so that phase 2 can always walk all 8 slots of the quarter
#pragma unroll
            for (int q4 = 0; q4 < 10; q4++) *(float4 *)(st + 4 * q4) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        // run starts inside each quarter: a particle whose cell differs from its predecessor's
        const int pci = __shfl_up_sync(0xffffffffu, ci, 1), pcrow = __shfl_up_sync(0xffffffffu, crow, 1);
        const bool start = t < n && (ql == 0 || ci != pci || crow != pcrow);
        const unsigned sb = __ballot_sync(0xffffffffu, start);         // staging written by all lanes is visible after this
        const unsigned any = (sb | (sb >> 8) | (sb >> 16) | (sb >> 24)) & 0xFFu;    // warp-uniform: some quarter starts a run at tt
        const unsigned mine = (sb >> (qtr << 3)) & 0xFFu;
        __syncwarp();
        // ------------------------------------------------------------------ phase 2: quarter-warp = footprint
        // The four quarters walk their 8 particles in lockstep; only the window moves diverge.
CR_DO_PRAGMA(unroll CR_P2UNROLL)
        for (int tt = 0; tt < 8; ++tt) {
            const float4 v1 = szl[tt * (CR_STRIDE / 4) + 2];             // (qPz[kp], qPz[kp + 2], cell i, row id)
            if (any & (1u << tt)) {
                if (mine & (1u << tt)) move_window(__float_as_int(v1.z), __float_as_int(v1.w));
            }
            const float4 *sx = reinterpret_cast<const float4 *>(sq + tt * CR_STRIDE);
            const float4 qpx = sx[0], xa = sx[1], xb = sx[2];
            const float sy1 = syl[tt * CR_STRIDE], dsy = syl[tt * CR_STRIDE + 4], qpy = syl[tt * CR_STRIDE + 8];
            const float4 v0 = szl[tt * (CR_STRIDE / 4)];                 // (Sz1[kp], Sz1[kp + 2], dSz[kp], dSz[kp + 2])
            const float ya = fmaf(0.5f, dsy, sy1), yb = fmaf(1.f / 3.f, dsy, 0.5f * sy1);
            // packed over the lane's two rows (.x = row kp, .y = row kp + 2); FFMA2 / FMUL2 take the per-lane scalar as a
            // broadcast operand
            const float2 sz1 = make_float2(v0.x, v0.y), dsz = make_float2(v0.z, v0.w), qpz = make_float2(v1.x, v1.y);
            const float2 wx = __ffma2_rn(dsz, make_float2(yb, yb), __fmul2_rn(sz1, make_float2(ya, ya)));   // Wx(j,k)
            const float2 ja = __fmul2_rn(sz1, make_float2(qpy, qpy)), jb = __fmul2_rn(dsz, make_float2(qpy, qpy));   // Jy = XA*ja + XB*jb
            const float2 jc = __fmul2_rn(qpz, make_float2(sy1, sy1)), jd = __fmul2_rn(qpz, make_float2(dsy, dsy));   // Jz = XA*jc + XB*jd
            const float2 px01 = make_float2(qpx.x, qpx.y), px23 = make_float2(qpx.z, qpx.w);
            const float2 xa01 = make_float2(xa.x, xa.y), xa23 = make_float2(xa.z, xa.w);
            const float2 xb01 = make_float2(xb.x, xb.y), xb23 = make_float2(xb.z, xb.w);
            {
                const float2 s = make_float2(wx.x, wx.x), a2 = make_float2(ja.x, ja.x), b2 = make_float2(jb.x, jb.x),
                             c2 = make_float2(jc.x, jc.x), d2 = make_float2(jd.x, jd.x);
                axA01 = __ffma2_rn(px01, s, axA01); axA23 = __ffma2_rn(px23, s, axA23);
                ayA01 = __ffma2_rn(xa01, a2, __ffma2_rn(xb01, b2, ayA01)); ayA23 = __ffma2_rn(xa23, a2, __ffma2_rn(xb23, b2, ayA23));
                azA01 = __ffma2_rn(xa01, c2, __ffma2_rn(xb01, d2, azA01)); azA23 = __ffma2_rn(xa23, c2, __ffma2_rn(xb23, d2, azA23));
            }
            {
                const float2 s = make_float2(wx.y, wx.y), a2 = make_float2(ja.y, ja.y), b2 = make_float2(jb.y, jb.y),
                             c2 = make_float2(jc.y, jc.y), d2 = make_float2(jd.y, jd.y);
                axB01 = __ffma2_rn(px01, s, axB01); axB23 = __ffma2_rn(px23, s, axB23);
                ayB01 = __ffma2_rn(xa01, a2, __ffma2_rn(xb01, b2, ayB01)); ayB23 = __ffma2_rn(xa23, a2, __ffma2_rn(xb23, b2, ayB23));
                azB01 = __ffma2_rn(xa01, c2, __ffma2_rn(xb01, d2, azB01)); azB23 = __ffma2_rn(xa23, c2, __ffma2_rn(xb23, d2, azB23));
            }
        }
        __syncwarp();
    }
    if (wi != CR_NOWIN) {
#pragma unroll 1
        for (int s = 0; s < 4; s++) flush_slide();
    }
}

int cellrun_supported(const tgpu_ctx *h)
{
    return h->P.dim == 3 && (h->P.order == 1 || h->P.order == 2) && h->G.lot < (1ll << 30) && h->P.my < 65536 && h->P.mz < 32768;
}

// FASTP (tgpu_set_option "fast_push", default 1): SFU reciprocal / reciprocal square root in the Boris push (<= 2 ulp)
// instead of the IEEE division and square root; measured 3 % of the mover (15.33 against 15.79 ms per lap)
template <int ORDER, bool FUSED, bool LAZY, bool FASTP>
static int launch_one(tgpu_ctx *h, const CRArgs &A)
{
    static bool attr_set = false;
    if (!attr_set) {
        CK(cudaFuncSetAttribute(k_cellrun<ORDER, FUSED, LAZY, FASTP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CR_SMEM_BYTES));
        attr_set = true;
    }
    const unsigned per_block = CR_WARPS * 4u * CR_CHUNK;
    const unsigned blocks = (A.n + per_block - 1) / per_block;
    k_cellrun<ORDER, FUSED, LAZY, FASTP><<<blocks, CR_WARPS * 32, CR_SMEM_BYTES, h->stream>>>(A);
    CKK(h);
    return 0;
}
template <bool FUSED>
static int launch_any(tgpu_ctx *h, const CRArgs &A)
{
    const bool lazy = FUSED && A.perm != nullptr;
    if (FUSED && !h->opt_fast_push) {
        if (h->P.order == 2) return lazy ? launch_one<2, FUSED, FUSED, false>(h, A) : launch_one<2, FUSED, false, false>(h, A);
        return lazy ? launch_one<1, FUSED, FUSED, false>(h, A) : launch_one<1, FUSED, false, false>(h, A);
    }
    if (h->P.order == 2) return lazy ? launch_one<2, FUSED, FUSED, FUSED>(h, A) : launch_one<2, FUSED, false, FUSED>(h, A);
    return lazy ? launch_one<1, FUSED, FUSED, FUSED>(h, A) : launch_one<1, FUSED, false, FUSED>(h, A);
}

static void fill_args(tgpu_ctx *h, int s, CRArgs &A, float *cx, float *cy, float *cz)
{
    A.prim8 = h->prim8; A.cx = cx; A.cy = cy; A.cz = cz; A.nty = h->nty; A.G = h->G;
    A.qm = s ? h->P.qme : h->P.qmi; A.qs = s ? h->P.qe : h->P.qi;
    A.keyoff = 1u + (unsigned)h->G.mx + (unsigned)h->G.mx * (unsigned)h->G.my;
    A.general = !(h->G.perx && h->G.pery && h->G.perz) || h->G.sendy || h->G.sendz;
}

template <bool FUSED>
static int launch(tgpu_ctx *h, float *cx, float *cy, float *cz)
{
    for (int s = 0; s < 2; s++) {
        Species &S = h->sp[s];
        if (S.n == 0) continue;
        CRArgs A;
        fill_args(h, s, A, cx, cy, cz);
        A.s = S; A.d = S; A.perm = nullptr; A.n = (unsigned)S.n;
        const size_t nb = (size_t)h->G.nkeys + TGPU_NBIN_EXTRA;
        A.key = h->key[s]; A.slot = h->slot + (size_t)s * h->maxhlf; A.bincount = h->bincount + (size_t)s * nb;
        if (FUSED) CK(cudaMemsetAsync(A.bincount, 0, nb * sizeof(int32_t), h->stream));
        if (FUSED && h->lazy[s]) { A.perm = h->perm[s]; A.d = h->alt[s]; }     // gather through the pending permutation
        int rc = launch_any<FUSED>(h, A); if (rc) return rc;
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
    fill_args(h, s, A, h->shadow[0], h->shadow[1], h->shadow[2]);
    A.s = R; A.d = R; A.perm = nullptr; A.n = (unsigned)cnt;
    const size_t nb = (size_t)h->G.nkeys + TGPU_NBIN_EXTRA;
    A.key = h->key[s] + off; A.slot = h->slot + (size_t)s * h->maxhlf + off; A.bincount = h->bincount + (size_t)s * nb;
    return launch_any<true>(h, A);
}

// tgpu_deposit_particles fast path when the particles were moved elsewhere (mirror mode, tests)
int cellrun_deposit(tgpu_ctx *h)
{
    { int rc = prt_materialize(h); if (rc) return rc; }
    return launch<false>(h, h->f[6], h->f[7], h->f[8]);
}
