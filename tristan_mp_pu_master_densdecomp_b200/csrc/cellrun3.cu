// Cell-run deposit for 3D, 3rd-order shapes (-Ddd3): densdecomp_3ord, code/particles.F90:1112-1358, and loop A of
// deposit_particles, code/particles_movedeposit.F90:1381-1401.  (The 3rd-order mover stays on the generic kernel: its
// 4x4x4-node gather is load-bound and already fast; the deposit is what the per-particle atomics cripple.)
//
// Same output-stationary idea as cellrun.cu, widened: the old shape sits on slots 2..5 and the new one on slots 2..5
// shifted by -1/0/+1, so a cell's particles write a 6x6x6 footprint (slots 1..6).  One WARP owns one footprint: lane =
// one of the 32 non-corner (j,k) rows, 6 x-planes x 3 components = 18 register accumulators.  The four corner rows
// (slot 1 or 6 in both y and z) only receive charge from a particle that changes cell in y AND z in the same step;
// that particle's own lane deposits its single corner row with plain atomics in phase 1.
// The six x-planes form a ring: plane t of the window that starts at cell wi-2 lives in register (wi-2+t) mod 6, so
// sliding along x flushes and clears registers without moving any; the accumulate code is instantiated for the six
// rotations and selected with one uniform switch per particle.
#include "tgpu_internal.h"
#include "shapes.cuh"

#define C3_WARPS 8
#define C3_STRIDE 68
#define C3_CHUNK 256           // particles per warp

struct C3Args {
    Species s;
    long long n;
    float *cx, *cy, *cz;
    DevGeom G;
    float qs;
};

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
    if (vx != 0.f) atomicAdd(cx + idx, vx);
    if (vy != 0.f) atomicAdd(cy + idx, vy);
    if (vz != 0.f) atomicAdd(cz + idx, vz);
}

// accumulate one particle into the ring, logical plane s -> register (R + s) % 6
#define C3_ACC1(m, s)                                                                 \
    ax[m] = fmaf(XQ[s], wx, ax[m]);                                                   \
    ay[m] = fmaf(XAv[s], a, fmaf(XBv[s], b, ay[m]));                                  \
    az[m] = fmaf(XAv[s], c, fmaf(XBv[s], d, az[m]));
#define C3_ACC(R)                                                                     \
    C3_ACC1((R + 0) % 6, 0) C3_ACC1((R + 1) % 6, 1) C3_ACC1((R + 2) % 6, 2)           \
    C3_ACC1((R + 3) % 6, 3) C3_ACC1((R + 4) % 6, 4) C3_ACC1((R + 5) % 6, 5)

__global__ void __launch_bounds__(C3_WARPS * 32, 2) k_cellrun3(C3Args A)
{
    extern __shared__ __align__(16) float stage3[];        // [C3_WARPS][32][C3_STRIDE]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long gw = (long long)blockIdx.x * C3_WARPS + warp;
    const long long base = gw * C3_CHUNK;
    if (base >= A.n) return;
    const DevGeom &G = A.G;
    const int mx = G.mx, my = G.my;
    // lane -> one of the 32 non-corner rows of the 6x6 (j,k) footprint
    const int nrow = lane < 4 ? lane + 1 : lane < 28 ? lane + 2 : lane + 3;
    const int j = nrow % 6, k = nrow / 6;
    const int loff = (j - 2) + my * (k - 2);
    float *wst = stage3 + (size_t)warp * 32 * C3_STRIDE;

    int wi = 0, wrow = 0, rot = 0;
    bool have = false;
    float ax[6] = {0, 0, 0, 0, 0, 0}, ay[6] = {0, 0, 0, 0, 0, 0}, az[6] = {0, 0, 0, 0, 0, 0};

    for (int it = 0; it < C3_CHUNK / 32; ++it) {
        const long long t = base + it * 32 + lane;
        float *st = wst + lane * C3_STRIDE;
        int ci = -1, crow = -1;
        if (t < A.n) {
            const float x = A.s.x[t], y = A.s.y[t], z = A.s.z[t], u = A.s.u[t], v = A.s.v[t], w = A.s.w[t];
            const float q = A.s.ch[t] * A.qs;
            // old position recomputed from the new one (particles_movedeposit.F90:1384-1390)
            const float invgam = 1.f / sqrtf(1 + u * u + v * v + w * w);
            const float x1 = x - u * invgam * G.c, y1 = y - v * invgam * G.c, z1 = z - w * invgam * G.c;
            const int i1 = (int)x1, j1 = (int)y1, k1 = (int)z1;
            const int shi = (int)x - i1, shj = (int)y - j1, shk = (int)z - k1;
            ci = i1; crow = (j1 - 1) + my * (k1 - 1);
            const float third = 1.f / 3.f;
            float S1[6], S2[6], dSy[6], dSz[6], XB6[6], qPSx[6], qPSy[6], qPSz[6];
            shape6(x1 - i1, 0, S1); shape6(x - (int)x, shi, S2);
            {
                float ps = 0.f;
#pragma unroll
                for (int s = 0; s < 6; s++) {
                    const float dS = S2[s] - S1[s];
                    ps = ps + dS; qPSx[s] = q * ps;
                    st[s] = qPSx[s]; st[6 + s] = S1[s] + 0.5f * dS; XB6[s] = 0.5f * S1[s] + third * dS; st[12 + s] = XB6[s];
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
            if (shj != 0 && shk != 0) {
                // the one corner row this particle reaches: the old shape is zero there, only the dS*dS terms survive
                const int jc = shj < 0 ? 0 : 5, kc = shk < 0 ? 0 : 5;
                const float dy = jc ? dSy[5] : dSy[0], dz = kc ? dSz[5] : dSz[0];
                const float py = jc ? qPSy[5] : qPSy[0], pz = kc ? qPSz[5] : qPSz[0];
                const size_t idx0 = (size_t)((long long)mx * (crow + (jc - 2) + my * (kc - 2)) + (i1 - 3));
                const float wxc = third * dy * dz;
#pragma unroll
                for (int s = 0; s < 6; s++) red3c(A.cx, A.cy, A.cz, idx0 + s, qPSx[s] * wxc, py * (XB6[s] * dz), pz * (XB6[s] * dy));
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
                        const size_t idx0 = (size_t)((long long)mx * (wrow + loff) + (wi - 3));
                        const int di = ni - wi;
                        const int nplanes = (nrw == wrow && di > 0 && di < 6) ? di : 6;
#pragma unroll
                        for (int m = 0; m < 6; m++) {
                            int tp = m - rot; tp = tp < 0 ? tp + 6 : tp;          // plane held by register m
                            if (tp < nplanes) { red3c(A.cx, A.cy, A.cz, idx0 + tp, ax[m], ay[m], az[m]); ax[m] = 0.f; ay[m] = 0.f; az[m] = 0.f; }
                        }
                    }
                    wi = ni; wrow = nrw; have = true;
                    rot = (wi - 2) % 6; rot = rot < 0 ? rot + 6 : rot;
                }
            }
            const float4 x0 = *(const float4 *)(sp + 0), x1v = *(const float4 *)(sp + 4), x2v = *(const float4 *)(sp + 8),
                         x3 = *(const float4 *)(sp + 12), x4 = *(const float4 *)(sp + 16);
            const float4 yv = *(const float4 *)(sp + 20 + 4 * j), zv = *(const float4 *)(sp + 44 + 4 * k);
            const float XQ[6] = {x0.x, x0.y, x0.z, x0.w, x1v.x, x1v.y};
            const float XAv[6] = {x1v.z, x1v.w, x2v.x, x2v.y, x2v.z, x2v.w};
            const float XBv[6] = {x3.x, x3.y, x3.z, x3.w, x4.x, x4.y};
            const float sy1 = yv.x, dsy = yv.y, qpsy = yv.z, sz1 = zv.x, dsz = zv.y, qpsz = zv.z;
            const float ya = fmaf(0.5f, dsy, sy1), yb = fmaf(1.f / 3.f, dsy, 0.5f * sy1);
            const float wx = fmaf(yb, dsz, ya * sz1);
            const float a = qpsy * sz1, b = qpsy * dsz, c = qpsz * sy1, d = qpsz * dsy;
            switch (rot) {
            case 0: C3_ACC(0) break;
            case 1: C3_ACC(1) break;
            case 2: C3_ACC(2) break;
            case 3: C3_ACC(3) break;
            case 4: C3_ACC(4) break;
            default: C3_ACC(5) break;
            }
        }
        __syncwarp();
    }
    if (have) {
        const size_t idx0 = (size_t)((long long)mx * (wrow + loff) + (wi - 3));
#pragma unroll
        for (int m = 0; m < 6; m++) {
            int tp = m - rot; tp = tp < 0 ? tp + 6 : tp;
            red3c(A.cx, A.cy, A.cz, idx0 + tp, ax[m], ay[m], az[m]);
        }
    }
}

int cellrun3_supported(const tgpu_ctx *h) { return h->P.dim == 3 && h->P.order == 3; }

// tgpu_deposit_particles fast path for -Ddd3 (currents only; wrap / compaction / sort follow in prt_sort)
int cellrun3_deposit(tgpu_ctx *h)
{
    int rc = prt_materialize(h); if (rc) return rc;
    const size_t smem = (size_t)C3_WARPS * 32 * C3_STRIDE * sizeof(float);
    CK(cudaFuncSetAttribute(k_cellrun3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int s = 0; s < 2; s++) {
        Species &S = h->sp[s];
        if (S.n == 0) continue;
        C3Args A;
        A.s = S; A.n = S.n; A.cx = h->f[6]; A.cy = h->f[7]; A.cz = h->f[8]; A.G = h->G; A.qs = s ? h->P.qe : h->P.qi;
        long long warps = (S.n + C3_CHUNK - 1) / C3_CHUNK;
        int blocks = (int)((warps + C3_WARPS - 1) / C3_WARPS);
        k_cellrun3<<<blocks, C3_WARPS * 32, smem, h->stream>>>(A);
        CKK(h);
    }
    return 0;
}
