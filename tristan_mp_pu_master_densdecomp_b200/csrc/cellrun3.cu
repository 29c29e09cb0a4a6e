// Cell-run kernel for 3D, 3rd-order shapes (-Ddd3): densdecomp_3ord, code/particles.F90:1112-1358, loop A of
// deposit_particles, code/particles_movedeposit.F90:1381-1401, and -- in FUSED launches -- mover_3ord,
// code/particles_movedeposit.F90:943-1271 (4x4x4-node gather of the node-centred fields + Boris push), the sort key of the
// pushed particle and, through the lazy sort's permutation, the data-moving pass of the counting sort: one pass over the
// particles per lap instead of three (mover, physical scatter, deposit).
//
// Same output-stationary idea as cellrun.cu, widened: the old shape sits on slots 2..5 and the new one on slots 2..5
// shifted by -1/0/+1, so a cell's particles write a 6x6x6 footprint (slots 1..6).  One WARP owns one footprint: lane =
// one of the 32 non-corner (j,k) rows, 6 x-planes x 3 components = 18 register accumulators.  The four corner rows
// (slot 1 or 6 in both y and z) only receive charge from a particle that changes cell in y AND z in the same step;
// that particle's own lane deposits its single corner row with plain atomics in phase 1.
// The six x-planes form a ring: plane t of the window that starts at cell wi-2 lives in register (wi-2+t) mod 6, so
// sliding along x flushes and clears registers without moving any.  Phase 1 stores each particle's x factors already
// rotated into ring order (component m of the staged vectors belongs to the cell congruent to m mod 6), so phase 2 is one
// fixed sequence of 15 packed FMAs (FFMA2, register pairs (0,1) (2,3) (4,5)) with no rotation switch.
#include "tgpu_internal.h"
#include "shapes.cuh"
#include "cellrun_common.cuh"

#define C3_WARPS 8
#define C3_STRIDE 68
#define C3_CHUNK 256           // particles per warp

struct C3Args {
    Species s;
    Species d;               // FUSED: destination records (logical order); aliases s unless a permutation is pending
    const int32_t *perm;     // LAZY: logical position t reads physical record perm[t]
    long long n;
    const float4 *prim8;     // FUSED: node-centred fields
    float *cx, *cy, *cz;     // the tiled shadow arrays (tgpu_internal.h row_index) when TGPU_SHADOW_TILED, else curx..curz
    int nty;
    DevGeom G;
    float qs, qm;
    uint32_t *key; int32_t *slot, *bincount;     // FUSED: sort key / rank in bin of the pushed particle (prt_sort skips its classify pass)
    unsigned keyoff; int general;
};
#define C3_TILED (TGPU_SHADOW_TILED != 0)
#define C3_PS (C3_TILED ? 16 : 1)          // distance between consecutive x-planes of a row

// cubic B-spline weights on slots 2..5 (particles.F90:1175-1188); W6[0..5] <-> slots 1..6 of the cell `base`,
// for a particle whose own cell is base + shift
__device__ __forceinline__ void shape6(float d, int shift, float W6[6])
{
    const float half = 0.5f, one = 1.f, two = 2.f, twoth = 2.f / 3.f, sixth = 1.f / 6.f, negsixth = -1.f / 6.f, negone = -1.f;
    float s2, s3, s4, s5;
    if (d <= half) {
        s2 = negsixth * (d - one) * (d - one) * (d - one);
        s3 = twoth + half * (d - two) * d * d;
        s5 = sixth * d * d * d;
        s4 = one - s5 - s3 - s2;
    } else {
        s5 = sixth * d * d * d;
        s4 = twoth + half * (negone - d) * (one - d) * (one - d);
        s2 = sixth * (one - d) * (one - d) * (one - d);
        s3 = one - s5 - s4 - s2;
    }
    W6[0] = shift < 0 ? s2 : 0.f;
    W6[1] = shift < 0 ? s3 : shift == 0 ? s2 : 0.f;
    W6[2] = shift < 0 ? s4 : shift == 0 ? s3 : s2;
    W6[3] = shift < 0 ? s5 : shift == 0 ? s4 : s3;
    W6[4] = shift < 0 ? 0.f : shift == 0 ? s5 : s4;
    W6[5] = shift > 0 ? s5 : 0.f;
}

__device__ __forceinline__ void red3c(float *cx, float *cy, float *cz, size_t idx, float vx, float vy, float vz)
{
    red_nz(cx + idx, vx); red_nz(cy + idx, vy); red_nz(cz + idx, vz);
}

template <bool FUSED, bool LAZY>
__global__ void __launch_bounds__(C3_WARPS * 32, 2) k_cellrun3(const C3Args A)
{
    extern __shared__ __align__(16) float stage3[];        // [C3_WARPS][32][C3_STRIDE] factor staging, then the record pipeline
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long gw = (long long)blockIdx.x * C3_WARPS + warp;
    const long long base = gw * C3_CHUNK;
    if (base >= A.n) return;
    const DevGeom &G = A.G;
    const int mx = G.mx, my = G.my;
    // lane -> one of the 32 non-corner rows of the 6x6 (j,k) footprint
    const int nrow = lane < 4 ? lane + 1 : lane < 28 ? lane + 2 : lane + 3;
    const int j = nrow % 6, k = nrow / 6;
    float *wst = stage3 + (size_t)warp * 32 * C3_STRIDE;

    int wi = 0, wrow = 0, rot = 0;
    bool have = false;
    float ax[6] = {0, 0, 0, 0, 0, 0}, ay[6] = {0, 0, 0, 0, 0, 0}, az[6] = {0, 0, 0, 0, 0, 0};

    // record pipeline (same idea as cellrun.cu): the next step's seven floats per lane are fetched with 4-byte cp.async
    // into per-lane landing slots while the current step is being deposited
    uint32_t *rec = reinterpret_cast<uint32_t *>(stage3 + (size_t)C3_WARPS * 32 * C3_STRIDE) + (size_t)warp * 2 * 9 * 32;
    auto fetch = [&](int buf, long long tt, int pp32) {
        if (tt < A.n) {
            const long long pp = LAZY ? (long long)pp32 : tt;
            uint32_t *r = rec + (size_t)buf * 9 * 32 + lane;
            cp_async4(r + 0 * 32, A.s.x + pp); cp_async4(r + 1 * 32, A.s.y + pp); cp_async4(r + 2 * 32, A.s.z + pp);
            cp_async4(r + 3 * 32, A.s.u + pp); cp_async4(r + 4 * 32, A.s.v + pp); cp_async4(r + 5 * 32, A.s.w + pp);
            cp_async4(r + 6 * 32, A.s.ch + pp);
            if (LAZY) { cp_async4(r + 7 * 32, A.s.ind + pp); cp_async4(r + 8 * 32, A.s.tag + pp); }
        }
        cp_async_commit();
    };
    {
        const long long t0 = base + lane;
        fetch(0, t0, (LAZY && t0 < A.n) ? __ldcs(A.perm + t0) : 0);
    }
    int pnext = 0;                     // permutation entry of the next step's particle, loaded one step ahead
    if (LAZY && base + 32 + lane < A.n) pnext = __ldcs(A.perm + base + 32 + lane);
    for (int it = 0; it < C3_CHUNK / 32; ++it) {
        const long long t = base + it * 32 + lane;
        float *st = wst + lane * C3_STRIDE;
        int ci = -1, crow = -1;
        cp_async_wait_all();
        if (it + 1 < C3_CHUNK / 32) {
            fetch((it + 1) & 1, t + 32, pnext);
            if (LAZY && it + 2 < C3_CHUNK / 32 && t + 64 < A.n) pnext = __ldcs(A.perm + t + 64);
        }
        if (t < A.n) {
            const uint32_t *r = rec + (size_t)(it & 1) * 9 * 32 + lane;
            float x = __uint_as_float(r[0 * 32]), y = __uint_as_float(r[1 * 32]), z = __uint_as_float(r[2 * 32]);
            float u = __uint_as_float(r[3 * 32]), v = __uint_as_float(r[4 * 32]), w = __uint_as_float(r[5 * 32]);
            const float ch = __uint_as_float(r[6 * 32]);
            const float q = ch * A.qs;
            int rank_in_bin = 0;
            if (FUSED) {
                if (LAZY) {
                    // the record still carries last lap's unwrapped position: apply the wrap its sort key was computed with
                    bool lo, hi;
                    x = wrap1(x, G.minx, G.maxx, G.shiftx_lo, G.shiftx_hi, lo, hi);
                    y = wrap1(y, G.miny, G.maxy, G.shifty_lo, G.shifty_hi, lo, hi);
                    z = wrap1(z, G.minz, G.maxz, G.shiftz_lo, G.shiftz_hi, lo, hi);
                    A.d.ch[t] = ch; A.d.ind[t] = (int32_t)r[7 * 32]; A.d.tag[t] = (int32_t)r[8 * 32];
                }
                // mover_3ord (particles_movedeposit.F90:1035-1128, 1130-1160): cubic B-spline weights on slots 2..5 of each
                // axis = nodes ip-1..ip+2; the six node-centred fields summed x innermost, then *Sy*Sz
                const int ip = (int)x, jp = (int)y, kq = (int)z;
                float W6[6], wxs[4], wys[4], wzs[4];
                shape6(x - ip, 0, W6); wxs[0] = W6[1]; wxs[1] = W6[2]; wxs[2] = W6[3]; wxs[3] = W6[4];
                shape6(y - jp, 0, W6); wys[0] = W6[1]; wys[1] = W6[2]; wys[2] = W6[3]; wys[3] = W6[4];
                shape6(z - kq, 0, W6); wzs[0] = W6[1]; wzs[1] = W6[2]; wzs[2] = W6[3]; wzs[3] = W6[4];
                float e0 = 0, e1 = 0, e2 = 0, b0 = 0, b1 = 0, b2 = 0;
                const int nbase = (ip - 2) + mx * ((jp - 2) + my * (kq - 2));
                gather_nodes<4>(A.prim8, nbase, mx, my, wxs, wys, wzs, e0, e1, e2, b0, b1, b2);
                const float cinv = G.cinv, qm = A.qm;
                e0 = 0.5f * e0 * qm; e1 = 0.5f * e1 * qm; e2 = 0.5f * e2 * qm;
                b0 = 0.5f * b0 * qm * cinv; b1 = 0.5f * b1 * qm * cinv; b2 = 0.5f * b2 * qm * cinv;
                if (G.external_fields) {
                    b0 = b0 + G.ext[3] * 0.5f * qm * cinv; b1 = b1 + G.ext[4] * 0.5f * qm * cinv; b2 = b2 + G.ext[5] * 0.5f * qm * cinv;
                    e0 = e0 + G.ext[0] * 0.5f * qm; e1 = e1 + G.ext[1] * 0.5f * qm; e2 = e2 + G.ext[2] * 0.5f * qm;
                }
                push_particle<false>(G.c, G.pusher, e0, e1, e2, b0, b1, b2, x, y, z, u, v, w);
                A.d.x[t] = x; A.d.y[t] = y; A.d.z[t] = z; A.d.u[t] = u; A.d.v[t] = v; A.d.w[t] = w;
                const uint32_t ky = sort_key(G, A.keyoff, A.general, x, y, z);
                A.key[t] = ky;
                rank_in_bin = atomicAdd(&A.bincount[ky], 1);
            }
            // old position recomputed from the new one (particles_movedeposit.F90:1384-1390)
            const float invgam = 1.f / sqrtf(1 + u * u + v * v + w * w);
            const float x1 = x - u * invgam * G.c, y1 = y - v * invgam * G.c, z1 = z - w * invgam * G.c;
            const int i1 = (int)x1, j1 = (int)y1, k1 = (int)z1;
            const int shi = (int)x - i1, shj = (int)y - j1, shk = (int)z - k1;
            ci = i1; crow = (j1 - 1) | ((k1 - 1) << 16);
            const float third = 1.f / 3.f;
            float S1[6], S2[6], dSy[6], dSz[6], XB6[6], qPSx[6], qPSy[6], qPSz[6];
            shape6(x1 - i1, 0, S1); shape6(x - (int)x, shi, S2);
            {
                // ring order: the factor of logical plane s (cell i1-2+s) goes to component (i1-2+s) mod 6
                int m = (i1 - 2) % 6; m = m < 0 ? m + 6 : m;
                float ps = 0.f;
#pragma unroll
                for (int s = 0; s < 6; s++) {
                    const float dS = S2[s] - S1[s];
                    ps = ps + dS; qPSx[s] = q * ps;
                    XB6[s] = 0.5f * S1[s] + third * dS;
                    st[m] = qPSx[s]; st[6 + m] = S1[s] + 0.5f * dS; st[12 + m] = XB6[s];
                    m = m == 5 ? 0 : m + 1;
                }
                st[18] = 0.f; st[19] = 0.f;
            }
            shape6(y1 - j1, 0, S1); shape6(y - (int)y, shj, S2);
            {
                float ps = 0.f;
#pragma unroll
                for (int s = 0; s < 6; s++) {
                    dSy[s] = S2[s] - S1[s]; ps = ps + dSy[s]; qPSy[s] = q * ps;
                    *(float4 *)(st + 20 + 4 * s) = make_float4(S1[s], dSy[s], qPSy[s], __int_as_float(ci));
                }
            }
            shape6(z1 - k1, 0, S1); shape6(z - (int)z, shk, S2);
            {
                float ps = 0.f;
#pragma unroll
                for (int s = 0; s < 6; s++) {
                    dSz[s] = S2[s] - S1[s]; ps = ps + dSz[s]; qPSz[s] = q * ps;
                    *(float4 *)(st + 44 + 4 * s) = make_float4(S1[s], dSz[s], qPSz[s], __int_as_float(crow));
                }
            }
            if (FUSED) A.slot[t] = rank_in_bin;
            if (shj != 0 && shk != 0) {
                // the one corner row this particle reaches: the old shape is zero there, only the dS*dS terms survive
                const int jc = shj < 0 ? 0 : 5, kc = shk < 0 ? 0 : 5;
                const float dy = jc ? dSy[5] : dSy[0], dz = kc ? dSz[5] : dSz[0];
                const float py = jc ? qPSy[5] : qPSy[0], pz = kc ? qPSz[5] : qPSz[0];
                const size_t idx0 = row_index<C3_TILED>(mx, my, A.nty, (j1 - 1) + jc - 2, (k1 - 1) + kc - 2, i1 - 3);
                const float wxc = third * dy * dz;
#pragma unroll
                for (int s = 0; s < 6; s++) red3c(A.cx, A.cy, A.cz, idx0 + (size_t)(s * C3_PS), qPSx[s] * wxc, py * (XB6[s] * dz), pz * (XB6[s] * dy));
            }
        }
        const int pci = __shfl_up_sync(0xffffffffu, ci, 1), pcrow = __shfl_up_sync(0xffffffffu, crow, 1);
        const unsigned starts = __ballot_sync(0xffffffffu, lane == 0 || ci != pci || crow != pcrow);
        __syncwarp();
        const long long rem = A.n - (base + it * 32);
        const int cnt = rem >= 32 ? 32 : (int)rem;
        const float *sp = wst;
        for (int tt = 0; tt < cnt; ++tt, sp += C3_STRIDE) {
            if ((starts >> tt) & 1u) {
                const int ni = __float_as_int(sp[23]), nrw = __float_as_int(sp[47]);
                if (!have || ni != wi || nrw != wrow) {
                    if (have) {
                        const size_t idx0 = row_index<C3_TILED>(mx, my, A.nty, (wrow & 0xFFFF) + j - 2, (wrow >> 16) + k - 2, wi - 3);
                        const int di = ni - wi;
                        if (nrw == wrow && di == 1) {
                            // the common case in sorted order: the window slides one cell, plane 0 (register rot) is complete;
                            // rot is uniform across the warp, so this is one short non-divergent case
                            float *px = A.cx + idx0, *py = A.cy + idx0, *pz = A.cz + idx0;
                            switch (rot) {
#define C3_FL(q) case q: red_nz(px, ax[q]); red_nz(py, ay[q]); red_nz(pz, az[q]); ax[q] = 0.f; ay[q] = 0.f; az[q] = 0.f; break;
                            C3_FL(0) C3_FL(1) C3_FL(2) C3_FL(3) C3_FL(4)
                            default: red_nz(px, ax[5]); red_nz(py, ay[5]); red_nz(pz, az[5]); ax[5] = 0.f; ay[5] = 0.f; az[5] = 0.f; break;
#undef C3_FL
                            }
                        } else {
                            const int nplanes = (nrw == wrow && di > 0 && di < 6) ? di : 6;
#pragma unroll
                            for (int m = 0; m < 6; m++) {
                                int tp = m - rot; tp = tp < 0 ? tp + 6 : tp;          // plane held by register m
                                if (tp < nplanes) { red3c(A.cx, A.cy, A.cz, idx0 + (size_t)(tp * C3_PS), ax[m], ay[m], az[m]); ax[m] = 0.f; ay[m] = 0.f; az[m] = 0.f; }
                            }
                        }
                    }
                    wi = ni; wrow = nrw; have = true;
                    rot = (wi - 2) % 6; rot = rot < 0 ? rot + 6 : rot;
                }
            }
            const float4 x0 = *(const float4 *)(sp + 0), x1v = *(const float4 *)(sp + 4), x2v = *(const float4 *)(sp + 8),
                         x3 = *(const float4 *)(sp + 12), x4 = *(const float4 *)(sp + 16);
            const float4 yv = *(const float4 *)(sp + 20 + 4 * j), zv = *(const float4 *)(sp + 44 + 4 * k);
            const float sy1 = yv.x, dsy = yv.y, qpsy = yv.z, sz1 = zv.x, dsz = zv.y, qpsz = zv.z;
            const float ya = fmaf(0.5f, dsy, sy1), yb = fmaf(1.f / 3.f, dsy, 0.5f * sy1);
            const float wx = fmaf(yb, dsz, ya * sz1);
            const float a = qpsy * sz1, b = qpsy * dsz, c = qpsz * sy1, d = qpsz * dsy;
            // staged vectors are in ring order: XQ = (x0, x1v.xy), XA = (x1v.zw, x2v), XB = (x3, x4.xy)
            const float2 w2 = make_float2(wx, wx), a2 = make_float2(a, a), b2 = make_float2(b, b), c2 = make_float2(c, c),
                         d2 = make_float2(d, d);
            const float2 q01 = make_float2(x0.x, x0.y), q23 = make_float2(x0.z, x0.w), q45 = make_float2(x1v.x, x1v.y);
            const float2 A01 = make_float2(x1v.z, x1v.w), A23 = make_float2(x2v.x, x2v.y), A45 = make_float2(x2v.z, x2v.w);
            const float2 B01 = make_float2(x3.x, x3.y), B23 = make_float2(x3.z, x3.w), B45 = make_float2(x4.x, x4.y);
            float2 t2;
#define C3_P(acc, m, expr) t2 = expr; acc[m] = t2.x; acc[m + 1] = t2.y;
            C3_P(ax, 0, __ffma2_rn(q01, w2, make_float2(ax[0], ax[1])))
            C3_P(ax, 2, __ffma2_rn(q23, w2, make_float2(ax[2], ax[3])))
            C3_P(ax, 4, __ffma2_rn(q45, w2, make_float2(ax[4], ax[5])))
            C3_P(ay, 0, __ffma2_rn(A01, a2, __ffma2_rn(B01, b2, make_float2(ay[0], ay[1]))))
            C3_P(ay, 2, __ffma2_rn(A23, a2, __ffma2_rn(B23, b2, make_float2(ay[2], ay[3]))))
            C3_P(ay, 4, __ffma2_rn(A45, a2, __ffma2_rn(B45, b2, make_float2(ay[4], ay[5]))))
            C3_P(az, 0, __ffma2_rn(A01, c2, __ffma2_rn(B01, d2, make_float2(az[0], az[1]))))
            C3_P(az, 2, __ffma2_rn(A23, c2, __ffma2_rn(B23, d2, make_float2(az[2], az[3]))))
            C3_P(az, 4, __ffma2_rn(A45, c2, __ffma2_rn(B45, d2, make_float2(az[4], az[5]))))
#undef C3_P
        }
        __syncwarp();
    }
    if (have) {
        const size_t idx0 = row_index<C3_TILED>(mx, my, A.nty, (wrow & 0xFFFF) + j - 2, (wrow >> 16) + k - 2, wi - 3);
#pragma unroll
        for (int m = 0; m < 6; m++) {
            int tp = m - rot; tp = tp < 0 ? tp + 6 : tp;
            red3c(A.cx, A.cy, A.cz, idx0 + (size_t)(tp * C3_PS), ax[m], ay[m], az[m]);
        }
    }
}

int cellrun3_supported(const tgpu_ctx *h) { return h->P.dim == 3 && h->P.order == 3 && h->P.my < 65536 && h->P.mz < 32768; }

#define C3_SMEM ((size_t)C3_WARPS * 32 * C3_STRIDE * sizeof(float) + (size_t)C3_WARPS * 2 * 9 * 32 * sizeof(uint32_t))

template <bool FUSED, bool LAZY>
static int launch3(tgpu_ctx *h, const C3Args &A)
{
    static bool attr_set = false;
    if (!attr_set) {
        CK(cudaFuncSetAttribute(k_cellrun3<FUSED, LAZY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C3_SMEM));
        attr_set = true;
    }
    const long long warps = (A.n + C3_CHUNK - 1) / C3_CHUNK;
    const int blocks = (int)((warps + C3_WARPS - 1) / C3_WARPS);
    k_cellrun3<FUSED, LAZY><<<blocks, C3_WARPS * 32, C3_SMEM, h->stream>>>(A);
    CKK(h);
    return 0;
}

template <bool FUSED>
static int run3(tgpu_ctx *h)
{
    for (int s = 0; s < 2; s++) {
        Species &S = h->sp[s];
        if (S.n == 0) continue;
        C3Args A;
        A.s = S; A.d = S; A.perm = nullptr; A.n = S.n; A.G = h->G; A.nty = h->nty; A.prim8 = h->prim8;
        A.qs = s ? h->P.qe : h->P.qi; A.qm = s ? h->P.qme : h->P.qmi;
        // deposits go to the shadow arrays (tiled: a flush touches a few 64 B tile rows instead of 32 cache lines) and
        // are folded into curx..curz by fld_add_shadow
        A.cx = h->shadow[0]; A.cy = h->shadow[1]; A.cz = h->shadow[2];
        const size_t nb = (size_t)h->G.nkeys + TGPU_NBIN_EXTRA;
        A.key = h->key[s]; A.slot = h->slot + (size_t)s * h->maxhlf; A.bincount = h->bincount + (size_t)s * nb;
        A.keyoff = 1u + (unsigned)h->G.mx + (unsigned)h->G.mx * (unsigned)h->G.my;
        A.general = !(h->G.perx && h->G.pery && h->G.perz) || h->G.sendy || h->G.sendz;
        int rc;
        if (FUSED) {
            CK(cudaMemsetAsync(A.bincount, 0, nb * sizeof(int32_t), h->stream));
            if (h->lazy[s]) { A.perm = h->perm[s]; A.d = h->alt[s]; rc = launch3<true, true>(h, A); }
            else rc = launch3<true, false>(h, A);
            if (rc) return rc;
            if (h->lazy[s]) {
                // the pushed records now sit, in sorted order and wrapped, in the other buffer
                const int n = S.n;
                Species tmp = h->sp[s]; h->sp[s] = h->alt[s]; h->alt[s] = tmp;
                h->sp[s].n = n; h->lazy[s] = 0; h->nphys[s] = n;
            }
        } else {
            rc = launch3<false, false>(h, A); if (rc) return rc;
        }
    }
    return 0;
}

// tgpu_move_particles fast path for -Ddd3: gather + push + deposit (into shadow[], folded into cur at deposit_particles time
// because mainloop resets cur in between) + sort keys, one pass; the same life cycle as cellrun_move_deposit (cellrun.cu)
int cellrun3_move_deposit(tgpu_ctx *h)
{
    int rc = fld_primal(h); if (rc) return rc;
    if (h->G.lot >= (1ll << 30)) { tgpu_set_error("cellrun3: grid too large for 32-bit node indices"); return TGPU_EINVAL; }
    if (h->presort) {
        h->presort = 0;                       // host order -> cell order once (see cellrun_move_deposit)
        rc = prt_sort(h, false); if (rc) return rc;
    }
    rc = run3<true>(h); if (rc) return rc;
    h->keys_valid = 1;
    return 0;
}

// tgpu_deposit_particles fast path for -Ddd3 when the particles were moved elsewhere (currents only; wrap / compaction /
// sort follow in prt_sort)
int cellrun3_deposit(tgpu_ctx *h)
{
    int rc = prt_materialize(h); if (rc) return rc;
    rc = run3<false>(h); if (rc) return rc;
    return fld_add_shadow(h);
}
