// Particle-side kernels, generic path: one thread per particle, any order (0..3), 2D and 3D.
//   mover / mover_{1,2,3}ord       code/particles_movedeposit.F90:98-347, 356-610, 619-933, 943-1271
//   zigzag / densdecomp_{1,2,3}ord code/particles.F90:550-669, 678-854, 864-1102, 1112-1358
//   deposit_particles loops B, C   code/particles_movedeposit.F90:1546-1705 (wrap / leave / discard)
//   reorder_particles_             code/particles.F90:418-497 (counting sort by cell)
//   exchange_particles, inject_others code/particles.F90:1865-2116, 1368-1852
// The cell-run (output-stationary) fast path for 3D orders 1 and 2 lives in cellrun.cu.
#include <cub/device/device_scan.cuh>
#include "tgpu_internal.h"
#include "shapes.cuh"

// ---------------------------------------------------------------------------------------------
// AoS (reference `type particle`, 40 B) <-> device SoA
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_aos2soa(const tgpu_particle *__restrict__ a, Species s, int off, int n)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    tgpu_particle p = a[t];
    int d = off + t;
    s.x[d] = p.x; s.y[d] = p.y; s.z[d] = p.z; s.u[d] = p.u; s.v[d] = p.v; s.w[d] = p.w; s.ch[d] = p.ch;
    s.ind[d] = p.ind; s.tag[d] = (p.proc & 0xFFFFFF) | (p.splitlev << 24);
}
__global__ void __launch_bounds__(256) k_soa2aos(tgpu_particle *__restrict__ a, Species s, int off, int n)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    int d = off + t;
    tgpu_particle p;
    p.x = s.x[d]; p.y = s.y[d]; p.z = s.z[d]; p.u = s.u[d]; p.v = s.v[d]; p.w = s.w[d]; p.ch = s.ch[d];
    p.ind = s.ind[d]; int tg = s.tag[d]; p.proc = tg & 0xFFFFFF; p.splitlev = (tg >> 24) & 0xFF;
    a[t] = p;
}

// SoA -> AoS of freshly pushed records, applying the periodic wrap of deposit_particles loop B
// (particles_movedeposit.F90:1546-1633) on the way; the wrapped position is also written back to the SoA
__global__ void __launch_bounds__(256) k_soa2aos_wrap(tgpu_particle *__restrict__ a, Species s, int off, int n, DevGeom G)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    int d = off + t;
    tgpu_particle p;
    float x = s.x[d], y = s.y[d], z = s.z[d];
    if (x < G.minx) x += G.shiftx_lo; else if (x > G.maxx) x -= G.shiftx_hi;
    if (y < G.miny) y += G.shifty_lo; else if (y > G.maxy) y -= G.shifty_hi;
    if (G.dim == 3) { if (z < G.minz) z += G.shiftz_lo; else if (z > G.maxz) z -= G.shiftz_hi; }
    s.x[d] = x; s.y[d] = y; s.z[d] = z;
    p.x = x; p.y = y; p.z = z; p.u = s.u[d]; p.v = s.v[d]; p.w = s.w[d]; p.ch = s.ch[d];
    p.ind = s.ind[d]; int tg = s.tag[d]; p.proc = tg & 0xFFFFFF; p.splitlev = (tg >> 24) & 0xFF;
    a[t] = p;
}

__global__ void __launch_bounds__(256) k_iota(int32_t *__restrict__ a, int first, int n)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) a[t] = first + t;
}

int prt_append(tgpu_ctx *h, int s, const tgpu_particle *p, int n, bool host)
{
    if (n <= 0) return 0;
    h->keys_valid = 0;
    Species &S = h->sp[s];
    if (h->lazy[s] && h->nphys[s] + n > h->maxhlf) { int rc = prt_materialize(h); if (rc) return rc; }
    const int phys0 = h->lazy[s] ? h->nphys[s] : S.n;          // where the new records physically go
    if (phys0 + n > h->maxhlf) { tgpu_set_error("particle capacity (maxhlf) exceeded"); return TGPU_EOVERFLOW; }
    if (!host) {
        k_aos2soa<<<cdiv(n, 256), 256, 0, h->stream>>>(p, S, phys0, n); CKK(h);
    } else {
        // double-buffered: the PCIe copy of chunk i+1 (copy stream) overlaps the AoS->SoA transpose of chunk i
        const size_t half = h->stage_particles / 2;
        cudaStream_t ck = h->stream, cc = h->stream == h->stream_main ? h->stream_prt : h->stream_main;
        CK(cudaEventRecord(h->ev_stage_free[0], ck)); CK(cudaEventRecord(h->ev_stage_free[1], ck));
        int done = 0, i = 0;
        while (done < n) {
            const int b = i & 1;
            int chunk = n - done; if ((size_t)chunk > half) chunk = (int)half;
            tgpu_particle *stg = h->stage + (size_t)b * half;
            CK(cudaStreamWaitEvent(cc, h->ev_stage_free[b], 0));
            CK(cudaMemcpyAsync(stg, p + done, (size_t)chunk * sizeof(tgpu_particle), cudaMemcpyHostToDevice, cc));
            CK(cudaEventRecord(h->ev_stage_full[b], cc));
            CK(cudaStreamWaitEvent(ck, h->ev_stage_full[b], 0));
            k_aos2soa<<<cdiv(chunk, 256), 256, 0, ck>>>(stg, S, phys0 + done, chunk); CKK(h);
            CK(cudaEventRecord(h->ev_stage_free[b], ck));
            done += chunk; i++;
        }
    }
    if (h->lazy[s]) {
        k_iota<<<cdiv(n, 256), 256, 0, h->stream>>>(h->perm[s] + S.n, phys0, n); CKK(h);
        h->nphys[s] += n;
    }
    S.n += n;
    return 0;
}
int prt_h2d(tgpu_ctx *h, const tgpu_particle *p, int ions, int lecs)
{
    if (ions < 0 || lecs < 0 || ions > h->maxhlf || lecs > h->maxhlf) { tgpu_set_error("bad particle counts"); return TGPU_EINVAL; }
    h->sp[0].n = 0; h->sp[1].n = 0; h->keys_valid = 0;
    h->lazy[0] = h->lazy[1] = 0; h->nphys[0] = h->nphys[1] = 0;
    int rc = prt_append(h, 0, p, ions, true); if (rc) return rc;
    rc = prt_append(h, 1, p + h->maxhlf, lecs, true); if (rc) return rc;
    CK(cudaStreamSynchronize(h->stream));
    h->presort = 1;          // host order (the reference sorts only every 10 laps): see cellrun_move_deposit
    return 0;
}
int prt_d2h(tgpu_ctx *h, tgpu_particle *p, int *ions, int *lecs)
{
    { int rc = prt_materialize(h); if (rc) return rc; }
    // double-buffered: the PCIe copy of chunk i (copy stream) overlaps the SoA->AoS transpose of chunk i+1
    const size_t half = h->stage_particles / 2;
    cudaStream_t ck = h->stream, cc = h->stream == h->stream_main ? h->stream_prt : h->stream_main;
    CK(cudaEventRecord(h->ev_stage_free[0], cc)); CK(cudaEventRecord(h->ev_stage_free[1], cc));
    int i = 0;
    for (int s = 0; s < 2; s++) {
        Species &S = h->sp[s];
        tgpu_particle *dst = p + (s ? h->maxhlf : 0);
        int done = 0;
        while (done < S.n) {
            const int b = i & 1;
            int chunk = S.n - done; if ((size_t)chunk > half) chunk = (int)half;
            tgpu_particle *stg = h->stage + (size_t)b * half;
            CK(cudaStreamWaitEvent(ck, h->ev_stage_free[b], 0));
            k_soa2aos<<<cdiv(chunk, 256), 256, 0, ck>>>(stg, S, done, chunk); CKK(h);
            CK(cudaEventRecord(h->ev_stage_full[b], ck));
            CK(cudaStreamWaitEvent(cc, h->ev_stage_full[b], 0));
            CK(cudaMemcpyAsync(dst + done, stg, (size_t)chunk * sizeof(tgpu_particle), cudaMemcpyDeviceToHost, cc));
            CK(cudaEventRecord(h->ev_stage_free[b], cc));
            done += chunk; i++;
        }
    }
    CK(cudaStreamSynchronize(cc));
    CK(cudaStreamSynchronize(h->stream));
    *ions = h->sp[0].n; *lecs = h->sp[1].n;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// generic mover
// ---------------------------------------------------------------------------------------------
struct Fields6 { const float *f[6]; const float4 *prim8; };

template <int ORDER, int DIM>
__global__ void __launch_bounds__(256) k_move(Species s, int n, Fields6 F, DevGeom G, float qm, float qme_abs)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    float x = s.x[t], y = s.y[t], z = s.z[t], u = s.u[t], v = s.v[t], w = s.w[t];
    const float cinv = 1.f / G.c;
    const long mx = G.mx, my = G.my;
    float e0 = 0, e1 = 0, e2 = 0, b0 = 0, b1 = 0, b2 = 0;
    float qm1 = qm;
    if (ORDER == 0) {
        // particles_movedeposit.F90:125-248 : trilinear staggered gather of the zigzag build
        if (s.ind[t] < 0 && qm > 0) qm1 = qme_abs;
        const float *ex = F.f[0], *ey = F.f[1], *ez = F.f[2], *bx = F.f[3], *by = F.f[4], *bz = F.f[5];
        int i = (int)x; float dx = x - i;
        int j = (int)y; float dy = y - j;
        int k = (int)z; float dz = z - k;
        const long ix = 1, iy = mx, iz = DIM == 3 ? mx * my : 0;
        if (DIM == 2) { k = 1; dz = 0; }
        long l = (i - 1) + iy * (j - 1) + iz * (k - 1);
        float f, g;
        f = ex[l] + ex[l - ix] + dx * (ex[l + ix] - ex[l - ix]);
        f = f + dy * (ex[l + iy] + ex[l - ix + iy] + dx * (ex[l + ix + iy] - ex[l - ix + iy]) - f);
        g = ex[l + iz] + ex[l - ix + iz] + dx * (ex[l + ix + iz] - ex[l - ix + iz]);
        g = g + dy * (ex[l + iy + iz] + ex[l - ix + iy + iz] + dx * (ex[l + ix + iy + iz] - ex[l - ix + iy + iz]) - g);
        e0 = (f + dz * (g - f)) * (.25f * qm1);
        f = ey[l] + ey[l - iy] + dy * (ey[l + iy] - ey[l - iy]);
        f = f + dz * (ey[l + iz] + ey[l - iy + iz] + dy * (ey[l + iy + iz] - ey[l - iy + iz]) - f);
        g = ey[l + ix] + ey[l - iy + ix] + dy * (ey[l + iy + ix] - ey[l - iy + ix]);
        g = g + dz * (ey[l + iz + ix] + ey[l - iy + iz + ix] + dy * (ey[l + iy + iz + ix] - ey[l - iy + iz + ix]) - g);
        e1 = (f + dx * (g - f)) * (.25f * qm1);
        f = ez[l] + ez[l - iz] + dz * (ez[l + iz] - ez[l - iz]);
        f = f + dx * (ez[l + ix] + ez[l - iz + ix] + dz * (ez[l + iz + ix] - ez[l - iz + ix]) - f);
        g = ez[l + iy] + ez[l - iz + iy] + dz * (ez[l + iz + iy] - ez[l - iz + iy]);
        g = g + dx * (ez[l + ix + iy] + ez[l - iz + ix + iy] + dz * (ez[l + iz + ix + iy] - ez[l - iz + ix + iy]) - g);
        e2 = (f + dy * (g - f)) * (.25f * qm1);
        f = bx[l - iy] + bx[l - iy - iz] + dz * (bx[l - iy + iz] - bx[l - iy - iz]);
        f = bx[l] + bx[l - iz] + dz * (bx[l + iz] - bx[l - iz]) + f +
            dy * (bx[l + iy] + bx[l + iy - iz] + dz * (bx[l + iy + iz] - bx[l + iy - iz]) - f);
        g = bx[l + ix - iy] + bx[l + ix - iy - iz] + dz * (bx[l + ix - iy + iz] - bx[l + ix - iy - iz]);
        g = bx[l + ix] + bx[l + ix - iz] + dz * (bx[l + ix + iz] - bx[l + ix - iz]) + g +
            dy * (bx[l + ix + iy] + bx[l + ix + iy - iz] + dz * (bx[l + ix + iy + iz] - bx[l + ix + iy - iz]) - g);
        b0 = (f + dx * (g - f)) * (.125f * qm1 * cinv);
        f = by[l - iz] + by[l - iz - ix] + dx * (by[l - iz + ix] - by[l - iz - ix]);
        f = by[l] + by[l - ix] + dx * (by[l + ix] - by[l - ix]) + f +
            dz * (by[l + iz] + by[l + iz - ix] + dx * (by[l + iz + ix] - by[l + iz - ix]) - f);
        g = by[l + iy - iz] + by[l + iy - iz - ix] + dx * (by[l + iy - iz + ix] - by[l + iy - iz - ix]);
        g = by[l + iy] + by[l + iy - ix] + dx * (by[l + iy + ix] - by[l + iy - ix]) + g +
            dz * (by[l + iy + iz] + by[l + iy + iz - ix] + dx * (by[l + iy + iz + ix] - by[l + iy + iz - ix]) - g);
        b1 = (f + dy * (g - f)) * (.125f * qm1 * cinv);
        f = bz[l - ix] + bz[l - ix - iy] + dy * (bz[l - ix + iy] - bz[l - ix - iy]);
        f = bz[l] + bz[l - iy] + dy * (bz[l + iy] - bz[l - iy]) + f +
            dx * (bz[l + ix] + bz[l + ix - iy] + dy * (bz[l + ix + iy] - bz[l + ix - iy]) - f);
        g = bz[l + iz - ix] + bz[l + iz - ix - iy] + dy * (bz[l + iz - ix + iy] - bz[l + iz - ix - iy]);
        g = bz[l + iz] + bz[l + iz - iy] + dy * (bz[l + iz + iy] - bz[l + iz - iy]) + g +
            dx * (bz[l + iz + ix] + bz[l + iz + ix - iy] + dy * (bz[l + iz + ix + iy] - bz[l + iz + ix - iy]) - g);
        b2 = (f + dz * (g - f)) * (.125f * qm1 * cinv);
    } else {
        // particles_movedeposit.F90:686-839 (order 2; orders 1 and 3 alike)
        const float half = 0.5f;
        float Sxp[8], Syp[8], Szp[8], Sxd[8], Syd[8];
        int pmin[3], pmax[3], dmin[3], dmax[3];
        int ip = (int)x, jp = (int)y, kp = (int)z;
        float dxp = x - ip, dyp = y - jp, dzp = z - kp;
        int id = (int)(x - half), jd = (int)(y - half), kd = (int)(z - half);
        float dxd = x - half - id, dyd = y - half - jd, dzd = (z - half) - kd;
        shape_slots<ORDER>(dxp, 0, Sxp, pmin[0], pmax[0]);
        shape_slots<ORDER>(dyp, 0, Syp, pmin[1], pmax[1]);
        shape_slots<ORDER>(dxd, 0, Sxd, dmin[0], dmax[0]);
        shape_slots<ORDER>(dyd, 0, Syd, dmin[1], dmax[1]);
        pmin[2] = pmax[2] = dmin[2] = dmax[2] = 3;
        if (DIM == 3) {
            float Szd[8];
            shape_slots<ORDER>(dzp, 0, Szp, pmin[2], pmax[2]);
            shape_slots<ORDER>(dzd, 0, Szd, dmin[2], dmax[2]);
        }
        const bool q1 = ORDER == 2 && (G.quirks & TGPU_Q1_MOVER2_RANGE);
        int imin[3], imax[3];
#pragma unroll
        for (int a = 0; a < 3; a++) {
            if (q1) { imin[a] = dmin[a]; imax[a] = dmax[a]; }
            else if (ORDER == 2 && DIM == 2) { imin[a] = 2; imax[a] = 5; }
            else { imin[a] = pmin[a]; imax[a] = pmax[a]; }
        }
        if (DIM == 3) {
            for (int i3 = imin[2]; i3 <= imax[2]; i3++)
                for (int i2 = imin[1]; i2 <= imax[1]; i2++) {
                    float sacc[6] = {0, 0, 0, 0, 0, 0};
                    long base = (ip - 3 - 1) + mx * ((jp - 3 + i2 - 1) + my * (long)(kp - 3 + i3 - 1));
                    for (int i1 = imin[0]; i1 <= imax[0]; i1++) {
                        float wx = Sxp[i1];
                        float4 lo = __ldg(&F.prim8[2 * (base + i1)]), hi = __ldg(&F.prim8[2 * (base + i1) + 1]);
                        sacc[0] = sacc[0] + lo.x * wx; sacc[1] = sacc[1] + lo.y * wx; sacc[2] = sacc[2] + lo.z * wx;
                        sacc[3] = sacc[3] + lo.w * wx; sacc[4] = sacc[4] + hi.x * wx; sacc[5] = sacc[5] + hi.y * wx;
                    }
                    float wyz_y = Syp[i2], wyz_z = Szp[i3];
                    e0 = e0 + sacc[0] * wyz_y * wyz_z; e1 = e1 + sacc[1] * wyz_y * wyz_z; e2 = e2 + sacc[2] * wyz_y * wyz_z;
                    b0 = b0 + sacc[3] * wyz_y * wyz_z; b1 = b1 + sacc[4] * wyz_y * wyz_z; b2 = b2 + sacc[5] * wyz_y * wyz_z;
                }
        } else {
            const float *ex = F.f[0], *ey = F.f[1], *ez = F.f[2], *bx = F.f[3], *by = F.f[4], *bz = F.f[5];
            for (int i2 = imin[1]; i2 <= imax[1]; i2++)
                for (int i1 = imin[0]; i1 <= imax[0]; i1++) {
                    long lpp = (ip - 3 + i1) + mx * (jp - 3 + i2 - 1) - 1;
                    long lpd = (ip - 3 + i1) + mx * (jd - 3 + i2 - 1) - 1;
                    long ldp = (id - 3 + i1) + mx * (jp - 3 + i2 - 1) - 1;
                    long ldd = (id - 3 + i1) + mx * (jd - 3 + i2 - 1) - 1;
                    e0 = e0 + ex[ldp] * Sxd[i1] * Syp[i2];
                    e1 = e1 + ey[lpd] * Sxp[i1] * Syd[i2];
                    e2 = e2 + ez[lpp] * Sxp[i1] * Syp[i2];
                    b0 = b0 + bx[lpd] * Sxp[i1] * Syd[i2];
                    b1 = b1 + by[ldp] * Sxd[i1] * Syp[i2];
                    b2 = b2 + bz[ldd] * Sxd[i1] * Syd[i2];
                }
        }
        e0 = 0.5f * e0 * qm; e1 = 0.5f * e1 * qm; e2 = 0.5f * e2 * qm;
        b0 = 0.5f * b0 * qm * cinv; b1 = 0.5f * b1 * qm * cinv; b2 = 0.5f * b2 * qm * cinv;
    }
    if (G.external_fields) {
        b0 = b0 + G.ext[3] * 0.5f * qm1 * cinv; b1 = b1 + G.ext[4] * 0.5f * qm1 * cinv; b2 = b2 + G.ext[5] * 0.5f * qm1 * cinv;
        e0 = e0 + G.ext[0] * 0.5f * qm1; e1 = e1 + G.ext[1] * 0.5f * qm1; e2 = e2 + G.ext[2] * 0.5f * qm1;
    }
    push_particle(G.c, G.pusher, e0, e1, e2, b0, b1, b2, x, y, z, u, v, w);
    s.x[t] = x; s.y[t] = y; s.z[t] = z; s.u[t] = u; s.v[t] = v; s.w[t] = w;
}

template <int ORDER>
static int launch_move(tgpu_ctx *h, int s, float qm)
{
    Species &S = h->sp[s];
    if (S.n == 0) return 0;
    Fields6 F;
    for (int a = 0; a < 6; a++) F.f[a] = h->f[a];
    F.prim8 = h->prim8;
    float qa = fabsf(h->P.qme);
    if (h->P.dim == 3) k_move<ORDER, 3><<<cdiv(S.n, 256), 256, 0, h->stream>>>(S, S.n, F, h->G, qm, qa);
    else k_move<ORDER, 2><<<cdiv(S.n, 256), 256, 0, h->stream>>>(S, S.n, F, h->G, qm, qa);
    CKK(h);
    return 0;
}

int prt_move_generic(tgpu_ctx *h)
{
    if (h->P.order > 0 && h->P.dim == 3) { int rc = fld_primal(h); if (rc) return rc; }
    for (int s = 0; s < 2; s++) {
        float qm = s ? h->P.qme : h->P.qmi;
        int rc;
        switch (h->P.order) {
        case 0: rc = launch_move<0>(h, s, qm); break;
        case 1: rc = launch_move<1>(h, s, qm); break;
        case 2: rc = launch_move<2>(h, s, qm); break;
        default: rc = launch_move<3>(h, s, qm); break;
        }
        if (rc) return rc;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// generic deposit (global fp32 atomics = REDG.ADD.F32)
// ---------------------------------------------------------------------------------------------
template <int ORDER, int DIM>
__global__ void __launch_bounds__(256) k_deposit(Species s, int n, float *__restrict__ curx, float *__restrict__ cury,
                                                 float *__restrict__ curz, DevGeom G, float qs)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    float x2 = s.x[t], y2 = s.y[t], z2 = s.z[t], u = s.u[t], v = s.v[t], w = s.w[t];
    // particles_movedeposit.F90:1384-1390
    float invgam = 1.f / sqrtf(1 + u * u + v * v + w * w);
    float x1 = x2 - u * invgam * G.c, y1 = y2 - v * invgam * G.c, z1 = z2 - w * invgam * G.c;
    float q = s.ch[t] * qs;
    const long mx = G.mx, my = G.my;
#define LI(i, j, k) ((size_t)((i)-1) + (size_t)mx * ((size_t)((j)-1) + (size_t)my * (size_t)((k)-1)))
    if (ORDER == 0) {
        // zigzag, particles.F90:578-664
        int i1 = (int)x1, i2 = (int)x2, j1 = (int)y1, j2 = (int)y2, k1 = (int)z1, k2 = (int)z2;
        float xr = fminf((float)(min(i1, i2) + 1), fmaxf((float)max(i1, i2), .5f * (x1 + x2)));
        float yr = fminf((float)(min(j1, j2) + 1), fmaxf((float)max(j1, j2), .5f * (y1 + y2)));
        float zr = fminf((float)(min(k1, k2) + 1), fmaxf((float)max(k1, k2), .5f * (z1 + z2)));
        if (DIM == 2) { k1 = 1; k2 = 1; }
        float Fx1 = -q * (xr - x1), Fy1 = -q * (yr - y1), Fz1 = -q * (zr - z1);
        float Wx1 = .5f * (x1 + xr) - i1, Wy1 = .5f * (y1 + yr) - j1, Wz1 = DIM == 3 ? .5f * (z1 + zr) - k1 : 0.f;
        float Wx2 = .5f * (x2 + xr) - i2, Wy2 = .5f * (y2 + yr) - j2, Wz2 = DIM == 3 ? .5f * (z2 + zr) - k2 : 0.f;
        float Fx2 = -q * (x2 - xr), Fy2 = -q * (y2 - yr), Fz2 = -q * (z2 - zr);
        atomicAdd(&curx[LI(i1, j1, k1)], Fx1 * (1.f - Wy1) * (1.f - Wz1));
        atomicAdd(&curx[LI(i1, j1 + 1, k1)], Fx1 * Wy1 * (1.f - Wz1));
        atomicAdd(&curx[LI(i2, j2, k2)], Fx2 * (1.f - Wy2) * (1.f - Wz2));
        atomicAdd(&curx[LI(i2, j2 + 1, k2)], Fx2 * Wy2 * (1.f - Wz2));
        atomicAdd(&cury[LI(i1, j1, k1)], Fy1 * (1.f - Wx1) * (1.f - Wz1));
        atomicAdd(&cury[LI(i1 + 1, j1, k1)], Fy1 * Wx1 * (1.f - Wz1));
        atomicAdd(&cury[LI(i2, j2, k2)], Fy2 * (1.f - Wx2) * (1.f - Wz2));
        atomicAdd(&cury[LI(i2 + 1, j2, k2)], Fy2 * Wx2 * (1.f - Wz2));
        if (DIM == 3) {
            atomicAdd(&curx[LI(i1, j1, k1 + 1)], Fx1 * (1 - Wy1) * Wz1);
            atomicAdd(&curx[LI(i1, j1 + 1, k1 + 1)], Fx1 * Wy1 * Wz1);
            atomicAdd(&curx[LI(i2, j2, k2 + 1)], Fx2 * (1.f - Wy2) * Wz2);
            atomicAdd(&curx[LI(i2, j2 + 1, k2 + 1)], Fx2 * Wy2 * Wz2);
            atomicAdd(&cury[LI(i1, j1, k1 + 1)], Fy1 * (1.f - Wx1) * Wz1);
            atomicAdd(&cury[LI(i1 + 1, j1, k1 + 1)], Fy1 * Wx1 * Wz1);
            atomicAdd(&cury[LI(i2, j2, k2 + 1)], Fy2 * (1.f - Wx2) * Wz2);
            atomicAdd(&cury[LI(i2 + 1, j2, k2 + 1)], Fy2 * Wx2 * Wz2);
        }
        atomicAdd(&curz[LI(i1, j1, k1)], Fz1 * (1.f - Wx1) * (1.f - Wy1));
        atomicAdd(&curz[LI(i1 + 1, j1, k1)], Fz1 * Wx1 * (1.f - Wy1));
        atomicAdd(&curz[LI(i1, j1 + 1, k1)], Fz1 * (1.f - Wx1) * Wy1);
        atomicAdd(&curz[LI(i1 + 1, j1 + 1, k1)], Fz1 * Wx1 * Wy1);
        atomicAdd(&curz[LI(i2, j2, k2)], Fz2 * (1.f - Wx2) * (1.f - Wy2));
        atomicAdd(&curz[LI(i2 + 1, j2, k2)], Fz2 * Wx2 * (1.f - Wy2));
        atomicAdd(&curz[LI(i2, j2 + 1, k2)], Fz2 * (1.f - Wx2) * Wy2);
        atomicAdd(&curz[LI(i2 + 1, j2 + 1, k2)], Fz2 * Wx2 * Wy2);
    } else {
        // densdecomp_*: particles.F90:886-1097.  The prefix sums along x (curx), y (cury) and z (curz) are
        // carried in registers; quirk Q4's stale carries are exactly zero in exact arithmetic and are dropped.
        const float half = 0.5f, third = 1.f / 3.f;
        float Sx1[8], Sy1[8], Sz1[8], Sx2[8], Sy2[8], Sz2[8];
        int i1 = (int)x1, j1 = (int)y1, k1 = (int)z1;
        int shifti = (int)x2 - i1, shiftj = (int)y2 - j1, shiftk = (int)z2 - k1;
        float dx1 = x1 - (int)x1, dy1 = y1 - (int)y1, dz1 = z1 - (int)z1;
        float dx2 = x2 - (int)x2, dy2 = y2 - (int)y2, dz2 = z2 - (int)z2;
        float deltaz = z2 - z1;
        int a1, b1, a2, b2, xmin, xmax, ymin, ymax, zmin = 3, zmax = 3;
        shape_slots<ORDER>(dx1, 0, Sx1, a1, b1); shape_slots<ORDER>(dx2, shifti, Sx2, a2, b2);
        xmin = min(a1, a2); xmax = max(b1, b2);
        shape_slots<ORDER>(dy1, 0, Sy1, a1, b1); shape_slots<ORDER>(dy2, shiftj, Sy2, a2, b2);
        ymin = min(a1, a2); ymax = max(b1, b2);
        if (DIM == 3) {
            shape_slots<ORDER>(dz1, 0, Sz1, a1, b1); shape_slots<ORDER>(dz2, shiftk, Sz2, a2, b2);
            zmin = min(a1, a2); zmax = max(b1, b2);
        } else k1 = 1;
        if (DIM == 3) {
            float cury_prev[8], curz_prev[8][8];
#pragma unroll
            for (int a = 0; a < 8; a++) { cury_prev[a] = 0.f;
#pragma unroll
                for (int b = 0; b < 8; b++) curz_prev[a][b] = 0.f; }
            for (int it2 = zmin; it2 <= zmax; it2++) {
                for (int a = 0; a < 8; a++) cury_prev[a] = 0.f;
                for (int it1 = ymin; it1 <= ymax; it1++) {
                    float cxp = 0.f;
                    for (int it = xmin; it <= xmax; it++) {
                        size_t l2 = LI(i1 - 3 + it, j1 - 3 + it1, k1 - 3 + it2);
                        float cx = q * ((Sx2[it] - Sx1[it]) *
                                        (Sy1[it1] * Sz1[it2] + half * (Sy2[it1] - Sy1[it1]) * Sz1[it2] +
                                         half * Sy1[it1] * (Sz2[it2] - Sz1[it2]) +
                                         third * (Sy2[it1] - Sy1[it1]) * (Sz2[it2] - Sz1[it2]))) + cxp;
                        float cy = q * ((Sy2[it1] - Sy1[it1]) *
                                        (Sx1[it] * Sz1[it2] + half * (Sx2[it] - Sx1[it]) * Sz1[it2] +
                                         half * Sx1[it] * (Sz2[it2] - Sz1[it2]) +
                                         third * (Sx2[it] - Sx1[it]) * (Sz2[it2] - Sz1[it2]))) + cury_prev[it];
                        float cz = q * ((Sz2[it2] - Sz1[it2]) *
                                        (Sx1[it] * Sy1[it1] + half * (Sx2[it] - Sx1[it]) * Sy1[it1] +
                                         half * Sx1[it] * (Sy2[it1] - Sy1[it1]) +
                                         third * (Sx2[it] - Sx1[it]) * (Sy2[it1] - Sy1[it1]))) + curz_prev[it][it1];
                        atomicAdd(&curx[l2], cx); atomicAdd(&cury[l2], cy); atomicAdd(&curz[l2], cz);
                        cxp = cx; cury_prev[it] = cy; curz_prev[it][it1] = cz;
                    }
                }
            }
        } else {
            float cury_prev[8];
#pragma unroll
            for (int a = 0; a < 8; a++) cury_prev[a] = 0.f;
            for (int it1 = ymin; it1 <= ymax; it1++) {
                float cxp = 0.f;
                for (int it = xmin; it <= xmax; it++) {
                    size_t l2 = LI(i1 - 3 + it, j1 - 3 + it1, k1);
                    float cx = q * ((Sx2[it] - Sx1[it]) * (Sy1[it1] + half * (Sy2[it1] - Sy1[it1]))) + cxp;
                    float cy = q * ((Sy2[it1] - Sy1[it1]) * (Sx1[it] + half * (Sx2[it] - Sx1[it]))) + cury_prev[it];
                    float cz = -1 * q * deltaz *
                               (Sx1[it] * Sy1[it1] + half * (Sx2[it] - Sx1[it]) * Sy1[it1] +
                                half * Sx1[it] * (Sy2[it1] - Sy1[it1]) + third * (Sx2[it] - Sx1[it]) * (Sy2[it1] - Sy1[it1]));
                    atomicAdd(&curx[l2], cx); atomicAdd(&cury[l2], cy); atomicAdd(&curz[l2], cz);
                    cxp = cx; cury_prev[it] = cy;
                }
            }
        }
    }
#undef LI
}

template <int ORDER>
static int launch_deposit(tgpu_ctx *h, int s, float qs)
{
    Species &S = h->sp[s];
    if (S.n == 0) return 0;
    if (h->P.dim == 3) k_deposit<ORDER, 3><<<cdiv(S.n, 256), 256, 0, h->stream>>>(S, S.n, h->f[6], h->f[7], h->f[8], h->G, qs);
    else k_deposit<ORDER, 2><<<cdiv(S.n, 256), 256, 0, h->stream>>>(S, S.n, h->f[6], h->f[7], h->f[8], h->G, qs);
    CKK(h);
    return 0;
}

int prt_deposit_generic(tgpu_ctx *h)
{
    for (int s = 0; s < 2; s++) {
        float qs = s ? h->P.qe : h->P.qi;
        int rc;
        switch (h->P.order) {
        case 0: rc = launch_deposit<0>(h, s, qs); break;
        case 1: rc = launch_deposit<1>(h, s, qs); break;
        case 2: rc = launch_deposit<2>(h, s, qs); break;
        default: rc = launch_deposit<3>(h, s, qs); break;
        }
        if (rc) return rc;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// classification + counting sort.  Stayers are keyed by cell (the reference's reorder key,
// particles.F90:441), leavers by neighbour code, discarded particles by the last bin; one scan and one
// scatter then (a) sort by cell, (b) compact, (c) partition the outboxes -- loops B and C of
// deposit_particles and reorder_particles_ in one pass.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void classify(const DevGeom &G, float &x, float &y, float &z, int &code, bool &discard)
{
    // particles_movedeposit.F90:1553-1633
    int dx = 0, dy = 0, dz = 0;
    if (x < G.minx) dx = -1; else if (x > G.maxx) dx = 1;
    if (y < G.miny) dy = -1; else if (y > G.maxy) dy = 1;
    if (z < G.minz) dz = -1; else if (z > G.maxz) dz = 1;
    bool in = true;
    if (!G.perx) in = (x + G.mxcum > G.x1in) && (x + G.mxcum < G.x2in);
    if (!G.pery && in) in = (y + G.mycum > G.y1in) && (y + G.mycum < G.y2in);
    if (G.dim == 3 && !G.perz && in) in = (z + G.mzcum > G.z1in) && (z + G.mzcum < G.z2in);
    discard = !in;
    code = 4;
    if (!in) return;
    // The reference tests x, then y, then z; once a particle is marked for sending along a split axis the remaining
    // shifts are zeroed (:1583-1586, :1598-1602).  In 3D the receiver re-tests z (inject_others, particles.F90:1431),
    // so every shift ends up applied.  In 2D a particle sent in x keeps an out-of-range y when y is not split, and any
    // sent particle keeps an out-of-range z: those wraps happen one lap later, at the next deposit_particles.
    const bool sent_x = G.dim == 2 && G.sendx && dx != 0;
    const bool sent_y = G.sendy && dy != 0;
    if (dx < 0) x = x + G.shiftx_lo; else if (dx > 0) x = x - G.shiftx_hi;
    if (!sent_x || G.sendy) { if (dy < 0) y = y + G.shifty_lo; else if (dy > 0) y = y - G.shifty_hi; }
    if (G.dim == 3 || !(sent_x || sent_y)) { if (dz < 0) z = z + G.shiftz_lo; else if (dz > 0) z = z - G.shiftz_hi; }
    int da, db;
    if (G.dim == 3) { da = G.sendy ? dy : 0; db = G.sendz ? dz : 0; }
    else { da = G.sendx ? dx : 0; db = G.sendy ? dy : 0; }
    code = (da + 1) + 3 * (db + 1);
}

__global__ void __launch_bounds__(256) k_classify_key(Species s, int n, DevGeom G, uint32_t *__restrict__ key,
                                                      int32_t *__restrict__ slot, int32_t *__restrict__ bincount)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    float x = s.x[t], y = s.y[t], z = s.z[t];
    int code; bool discard;
    classify(G, x, y, z, code, discard);
    s.x[t] = x; s.y[t] = y; s.z[t] = z;
    uint32_t k;
    if (discard) k = (uint32_t)G.nkeys + 9u;
    else if (code != 4) k = (uint32_t)G.nkeys + (uint32_t)code;
    else {
        int i = (int)x, j = (int)y, kk = G.dim == 3 ? (int)z : 1;
        i = min(max(i, 1), G.mx); j = min(max(j, 1), G.my); kk = min(max(kk, 1), G.mz);
        k = cell_key(G, i, j, kk);
    }
    key[t] = k;
    slot[t] = atomicAdd(&bincount[k], 1);
}

// WRAP: the keys were computed by the fused mover on a copy of the position; apply the periodic wrap / frame shift here
template <bool WRAP>
__global__ void __launch_bounds__(256) k_scatter(Species a, Species b, int n, const uint32_t *__restrict__ key,
                                                 const int32_t *__restrict__ slot, const int32_t *__restrict__ binoff, DevGeom G)
{
    // one thread per particle (measured faster than grid-stride loops over a bounded grid: 6.3 vs 8.2 ms per lap)
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    int d = binoff[key[t]] + slot[t];
    float x = a.x[t], y = a.y[t], z = a.z[t];
    if (WRAP) { int code; bool discard; classify(G, x, y, z, code, discard); }
    b.x[d] = x; b.y[d] = y; b.z[d] = z; b.u[d] = a.u[t]; b.v[d] = a.v[t]; b.w[d] = a.w[t];
    b.ch[d] = a.ch[t]; b.ind[d] = a.ind[t]; b.tag[d] = a.tag[t];
}

// lazy sort: instead of moving the records, remember where each sorted position's record is
__global__ void __launch_bounds__(256) k_build_perm(int32_t *__restrict__ perm, int n, const uint32_t *__restrict__ key,
                                                    const int32_t *__restrict__ slot, const int32_t *__restrict__ binoff)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) perm[binoff[key[t]] + slot[t]] = t;
}
// apply the permutation (and the wrap that the key was computed with)
__global__ void __launch_bounds__(256) k_gather_perm(Species a, Species b, int n, const int32_t *__restrict__ perm, DevGeom G)
{
    int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= n) return;
    int t = perm[d];
    float x = a.x[t], y = a.y[t], z = a.z[t];
    int code; bool discard; classify(G, x, y, z, code, discard);
    b.x[d] = x; b.y[d] = y; b.z[d] = z; b.u[d] = a.u[t]; b.v[d] = a.v[t]; b.w[d] = a.w[t];
    b.ch[d] = a.ch[t]; b.ind[d] = a.ind[t]; b.tag[d] = a.tag[t];
}
__global__ void __launch_bounds__(256) k_soa2aos_perm(tgpu_particle *__restrict__ a, Species s, const int32_t *__restrict__ perm, int off, int n, DevGeom G)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    int d = perm[off + t];
    tgpu_particle p;
    p.x = s.x[d]; p.y = s.y[d]; p.z = s.z[d];
    int code; bool discard; classify(G, p.x, p.y, p.z, code, discard);     // into the destination's frame
    p.u = s.u[d]; p.v = s.v[d]; p.w = s.w[d]; p.ch = s.ch[d];
    p.ind = s.ind[d]; int tg = s.tag[d]; p.proc = tg & 0xFFFFFF; p.splitlev = (tg >> 24) & 0xFF;
    a[t] = p;
}

int prt_materialize(tgpu_ctx *h)
{
    for (int s = 0; s < 2; s++) {
        if (!h->lazy[s]) continue;
        Species &S = h->sp[s];
        if (S.n) { k_gather_perm<<<cdiv(S.n, 256), 256, 0, h->stream>>>(S, h->alt[s], S.n, h->perm[s], h->G); CKK(h); }
        int n = S.n;
        Species tmp = h->sp[s]; h->sp[s] = h->alt[s]; h->alt[s] = tmp;
        h->sp[s].n = n; h->lazy[s] = 0; h->nphys[s] = n;
    }
    return 0;
}

// After prt_sort: sp[s].n = stayers; h_small[s*16 + c] / [s*16 + 8.. ] hold leaver ranges (offset table of the 11 tail bins).
int prt_sort(tgpu_ctx *h, bool)
{
    const int nb = (int)h->G.nkeys + TGPU_NBIN_EXTRA;
    const bool have_keys = h->keys_valid != 0;      // written by the fused mover for the records as they now sit in sp[]
    h->keys_valid = 0;
    if (!have_keys) { int rc = prt_materialize(h); if (rc) return rc; }
    const bool lazy = have_keys && h->opt_lazy;
    for (int s = 0; s < 2 && !have_keys; s++) {
        Species &S = h->sp[s];
        int32_t *cnt = h->bincount + (size_t)s * nb;
        CK(cudaMemsetAsync(cnt, 0, (size_t)nb * sizeof(int32_t), h->stream));
        if (S.n == 0) continue;
        k_classify_key<<<cdiv(S.n, 256), 256, 0, h->stream>>>(S, S.n, h->G, h->key[s], h->slot + (size_t)s * h->maxhlf, cnt);
        CKK(h);
    }
    for (int s = 0; s < 2; s++) {
        size_t bytes = h->cub_bytes;
        CK(cub::DeviceScan::ExclusiveSum(h->cub_tmp, bytes, h->bincount + (size_t)s * nb, h->binoff + (size_t)s * (nb + 1), nb + 1, h->stream));
        h->launches++;
    }
    for (int s = 0; s < 2; s++) {
        Species &S = h->sp[s];
        const int32_t *slot = h->slot + (size_t)s * h->maxhlf, *off = h->binoff + (size_t)s * (nb + 1);
        if (S.n) {
            if (lazy) k_build_perm<<<cdiv(S.n, 256), 256, 0, h->stream>>>(h->perm[s], S.n, h->key[s], slot, off);
            else if (have_keys) k_scatter<true><<<cdiv(S.n, 256), 256, 0, h->stream>>>(S, h->alt[s], S.n, h->key[s], slot, off, h->G);
            else k_scatter<false><<<cdiv(S.n, 256), 256, 0, h->stream>>>(S, h->alt[s], S.n, h->key[s], slot, off, h->G);
            CKK(h);
        }
        CK(cudaMemcpyAsync(h->h_small + s * 16, off + (size_t)h->G.nkeys, 11 * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    }
    CK(cudaStreamSynchronize(h->stream));
    for (int s = 0; s < 2; s++) {
        int n_old = h->sp[s].n;
        if (lazy) { h->lazy[s] = n_old > 0; h->nphys[s] = n_old; }
        else { Species tmp = h->sp[s]; h->sp[s] = h->alt[s]; h->alt[s] = tmp; h->lazy[s] = 0; h->nphys[s] = 0; }
        // h_small[s*16 + c] = offset of tail bin c (c = 0..9), [10] = total
        h->sp[s].n = n_old ? h->h_small[s * 16 + 0] : 0;
        if (!n_old) for (int c = 0; c < 11; c++) h->h_small[s * 16 + c] = 0;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// migration: one round over the (up to) 8 neighbours in the two decomposed axes replaces the reference's
// z/y/x pairwise SendRecv + second corner round (particles.F90:1904-2112, tristanmainloop.F90:264-275).
// Leavers sit, already shifted into the destination's frame, in the tail bins written by prt_sort.
// ---------------------------------------------------------------------------------------------
int prt_exchange(tgpu_ctx *h)
{
    if (h->size0 == 1) return 0;
    if (!h->nccl_comm) { tgpu_set_error("tgpu_exchange_particles: communicator not initialised (tgpu_comm_init)"); return TGPU_ENCCL; }
    const int B = h->P.buffsize;
    int nout[2][9], off[2][9], nin[2][9];
    // Error protocol.  The reference `stop`s the whole job on a full buffer (particles.F90:1933-1938).  Here nothing may
    // return between ncclGroupStart and ncclGroupEnd, and no rank may skip a group its neighbours enter: a rank whose
    // outbox overflows still takes part in both groups -- sending zero particles and an overflow flag with its counts --
    // and reports TGPU_EOVERFLOW afterwards, together with every neighbour that saw the flag.
    int overflow = 0;
    for (int s = 0; s < 2; s++) for (int c = 0; c < 9; c++) {
        off[s][c] = h->h_small[s * 16 + c]; nout[s][c] = h->h_small[s * 16 + c + 1] - h->h_small[s * 16 + c];
        if (c == 4) nout[s][c] = 0;
    }
    for (int c = 0; c < 9; c++) if (nout[0][c] + nout[1][c] > B) overflow = 1;
    if (overflow) for (int s = 0; s < 2; s++) for (int c = 0; c < 9; c++) nout[s][c] = 0;
    // pack: sendbuf[c] = ions then electrons
    int32_t *hc = h->h_small + 32;          // [c][3] out: ions, electrons, overflow flag; then [c][3] in
    for (int c = 0; c < 9; c++) {
        hc[3 * c + 2] = overflow;
        if (c == 4) { hc[3 * c] = hc[3 * c + 1] = 0; continue; }
        int o = 0;
        for (int s = 0; s < 2; s++) {
            if (nout[s][c]) {
                if (h->lazy[s]) k_soa2aos_perm<<<cdiv(nout[s][c], 256), 256, 0, h->stream>>>(h->sendbuf + (size_t)c * B + o, h->sp[s], h->perm[s], off[s][c], nout[s][c], h->G);
                else k_soa2aos<<<cdiv(nout[s][c], 256), 256, 0, h->stream>>>(h->sendbuf + (size_t)c * B + o, h->sp[s], off[s][c], nout[s][c]);
                CKK(h);
            }
            o += nout[s][c];
            hc[3 * c + s] = nout[s][c];
        }
    }
    CK(cudaMemcpyAsync(h->d_small, hc, 27 * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
    int rc = comm_group_begin(h); if (rc) return rc;
    int rc_in = 0;
    for (int c = 0; c < 9; c++) {
        if (c == 4) continue;
        int da = c % 3 - 1, db = c / 3 - 1;
        int to = topo_neighbour2(h, da, db), from = topo_neighbour2(h, -da, -db);
        int r1 = comm_send(h, h->d_small + 3 * c, 3 * sizeof(int32_t), to);
        int r2 = comm_recv(h, h->d_small + 27 + 3 * c, 3 * sizeof(int32_t), from);
        if (!rc_in) rc_in = r1 ? r1 : r2;
    }
    rc = comm_group_end(h); if (rc_in) return rc_in; if (rc) return rc;
    CK(cudaMemcpyAsync(hc + 27, h->d_small + 27, 27 * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    int remote_overflow = 0;
    for (int c = 0; c < 9; c++) {
        for (int s = 0; s < 2; s++) nin[s][c] = c == 4 ? 0 : hc[27 + 3 * c + s];
        if (c != 4 && hc[27 + 3 * c + 2]) remote_overflow = 1;
        // cannot happen with a job-wide buffsize (the sender clips at the same B); checked BEFORE the payload group opens
        if (nin[0][c] + nin[1][c] > B) { tgpu_set_error("migration inbox overflow: buffsize differs between ranks"); return TGPU_EOVERFLOW; }
    }
    rc = comm_group_begin(h); if (rc) return rc;
    for (int c = 0; c < 9; c++) {
        if (c == 4) continue;
        int da = c % 3 - 1, db = c / 3 - 1;
        int to = topo_neighbour2(h, da, db), from = topo_neighbour2(h, -da, -db);
        int ns = nout[0][c] + nout[1][c], nr = nin[0][c] + nin[1][c];
        int r1 = ns ? comm_send(h, h->sendbuf + (size_t)c * B, (size_t)ns * sizeof(tgpu_particle), to) : 0;
        int r2 = nr ? comm_recv(h, h->recvbuf + (size_t)c * B, (size_t)nr * sizeof(tgpu_particle), from) : 0;
        if (!rc_in) rc_in = r1 ? r1 : r2;
    }
    rc = comm_group_end(h); if (rc_in) return rc_in; if (rc) return rc;
    for (int c = 0; c < 9; c++) {
        if (c == 4) continue;
        rc = prt_append(h, 0, h->recvbuf + (size_t)c * B, nin[0][c], false); if (rc) return rc;
        rc = prt_append(h, 1, h->recvbuf + (size_t)c * B + nin[0][c], nin[1][c], false); if (rc) return rc;
    }
    for (int s = 0; s < 2; s++) for (int c = 0; c < 11; c++) h->h_small[s * 16 + c] = h->sp[s].n;   // outboxes consumed
    if (overflow) { tgpu_set_error("migration outbox overflow (buffsize)"); return TGPU_EOVERFLOW; }
    if (remote_overflow) { tgpu_set_error("migration outbox overflow (buffsize) on a neighbouring rank"); return TGPU_EOVERFLOW; }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// particle_bc_user of the shock problem: user/user_shock.F90:377-457 (gammawall = 1, betawall = 0).
// Runs between the mover and deposit_particles; deposits two zigzag segments (particles.F90:550-669) per reflected
// particle: the path up to the wall, and minus the piece behind the wall that deposit_particles will add.
// ---------------------------------------------------------------------------------------------
template <int DIM>
__device__ __forceinline__ void zigzag_dev(float *curx, float *cury, float *curz, const DevGeom &G, float x2, float y2,
                                           float z2, float x1, float y1, float z1, float q)
{
    const long mx = G.mx, my = G.my;
#define LI(i, j, k) ((size_t)((i)-1) + (size_t)mx * ((size_t)((j)-1) + (size_t)my * (size_t)((k)-1)))
    int i1 = (int)x1, i2 = (int)x2, j1 = (int)y1, j2 = (int)y2, k1 = (int)z1, k2 = (int)z2;
    float xr = fminf((float)(min(i1, i2) + 1), fmaxf((float)max(i1, i2), .5f * (x1 + x2)));
    float yr = fminf((float)(min(j1, j2) + 1), fmaxf((float)max(j1, j2), .5f * (y1 + y2)));
    float zr = fminf((float)(min(k1, k2) + 1), fmaxf((float)max(k1, k2), .5f * (z1 + z2)));
    if (DIM == 2) { k1 = 1; k2 = 1; }
    float Fx1 = -q * (xr - x1), Fy1 = -q * (yr - y1), Fz1 = -q * (zr - z1);
    float Wx1 = .5f * (x1 + xr) - i1, Wy1 = .5f * (y1 + yr) - j1, Wz1 = DIM == 3 ? .5f * (z1 + zr) - k1 : 0.f;
    float Wx2 = .5f * (x2 + xr) - i2, Wy2 = .5f * (y2 + yr) - j2, Wz2 = DIM == 3 ? .5f * (z2 + zr) - k2 : 0.f;
    float Fx2 = -q * (x2 - xr), Fy2 = -q * (y2 - yr), Fz2 = -q * (z2 - zr);
    atomicAdd(&curx[LI(i1, j1, k1)], Fx1 * (1.f - Wy1) * (1.f - Wz1));
    atomicAdd(&curx[LI(i1, j1 + 1, k1)], Fx1 * Wy1 * (1.f - Wz1));
    atomicAdd(&curx[LI(i2, j2, k2)], Fx2 * (1.f - Wy2) * (1.f - Wz2));
    atomicAdd(&curx[LI(i2, j2 + 1, k2)], Fx2 * Wy2 * (1.f - Wz2));
    atomicAdd(&cury[LI(i1, j1, k1)], Fy1 * (1.f - Wx1) * (1.f - Wz1));
    atomicAdd(&cury[LI(i1 + 1, j1, k1)], Fy1 * Wx1 * (1.f - Wz1));
    atomicAdd(&cury[LI(i2, j2, k2)], Fy2 * (1.f - Wx2) * (1.f - Wz2));
    atomicAdd(&cury[LI(i2 + 1, j2, k2)], Fy2 * Wx2 * (1.f - Wz2));
    if (DIM == 3) {
        atomicAdd(&curx[LI(i1, j1, k1 + 1)], Fx1 * (1 - Wy1) * Wz1);
        atomicAdd(&curx[LI(i1, j1 + 1, k1 + 1)], Fx1 * Wy1 * Wz1);
        atomicAdd(&curx[LI(i2, j2, k2 + 1)], Fx2 * (1.f - Wy2) * Wz2);
        atomicAdd(&curx[LI(i2, j2 + 1, k2 + 1)], Fx2 * Wy2 * Wz2);
        atomicAdd(&cury[LI(i1, j1, k1 + 1)], Fy1 * (1.f - Wx1) * Wz1);
        atomicAdd(&cury[LI(i1 + 1, j1, k1 + 1)], Fy1 * Wx1 * Wz1);
        atomicAdd(&cury[LI(i2, j2, k2 + 1)], Fy2 * (1.f - Wx2) * Wz2);
        atomicAdd(&cury[LI(i2 + 1, j2, k2 + 1)], Fy2 * Wx2 * Wz2);
    }
    atomicAdd(&curz[LI(i1, j1, k1)], Fz1 * (1.f - Wx1) * (1.f - Wy1));
    atomicAdd(&curz[LI(i1 + 1, j1, k1)], Fz1 * Wx1 * (1.f - Wy1));
    atomicAdd(&curz[LI(i1, j1 + 1, k1)], Fz1 * (1.f - Wx1) * Wy1);
    atomicAdd(&curz[LI(i1 + 1, j1 + 1, k1)], Fz1 * Wx1 * Wy1);
    atomicAdd(&curz[LI(i2, j2, k2)], Fz2 * (1.f - Wx2) * (1.f - Wy2));
    atomicAdd(&curz[LI(i2 + 1, j2, k2)], Fz2 * Wx2 * (1.f - Wy2));
    atomicAdd(&curz[LI(i2, j2 + 1, k2)], Fz2 * (1.f - Wx2) * Wy2);
    atomicAdd(&curz[LI(i2 + 1, j2 + 1, k2)], Fz2 * Wx2 * Wy2);
#undef LI
}

template <int DIM>
__global__ void __launch_bounds__(256) k_wall(Species s, int n, DevGeom G, float walloc, float q0, float *curx, float *cury, float *curz)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    float x = s.x[t];
    if (!(x + G.mxcum < walloc)) return;
    const float c = G.c, gammawall = 1.f, betawall = 0.f;
    float y = s.y[t], z = s.z[t], u = s.u[t], v = s.v[t], w = s.w[t];
    float gamma = sqrtf(1 + (u * u + v * v + w * w));
    float x0 = x - u / gamma * c, y0 = y, z0 = z;
    float walloc0 = walloc - betawall * c - G.mxcum;
    float tfrac = fabsf((x0 - walloc0) / (betawall * c - u / gamma * c));
    float xcolis = x0 + u / gamma * c * tfrac, ycolis = y0, zcolis = z0;
    float q = s.ch[t] * q0;
    zigzag_dev<DIM>(curx, cury, curz, G, xcolis, ycolis, zcolis, x0, y0, z0, q);
    u = gammawall * gammawall * gamma * (2 * betawall - u / gamma * (1 + betawall * betawall));
    gamma = sqrtf(1 + (u * u + v * v + w * w));
    tfrac = fminf(fabsf((x - xcolis) / fmaxf(fabsf(x - x0), 1e-9f)), 1.f);
    x = xcolis + u / gamma * c * tfrac; y = ycolis; z = zcolis;
    q = -q;
    zigzag_dev<DIM>(curx, cury, curz, G, xcolis, ycolis, zcolis, x - u / gamma * c, y - v / gamma * c, z - w / gamma * c, q);
    s.x[t] = x; s.y[t] = y; s.z[t] = z; s.u[t] = u;
}

int prt_wall(tgpu_ctx *h, float leftwall)
{
    int rc = prt_materialize(h); if (rc) return rc;
    h->keys_valid = 0;
    for (int s = 0; s < 2; s++) {
        Species &S = h->sp[s];
        if (!S.n) continue;
        float q0 = s ? h->P.qe : h->P.qi;
        if (h->P.dim == 3) k_wall<3><<<cdiv(S.n, 256), 256, 0, h->stream>>>(S, S.n, h->G, leftwall, q0, h->f[6], h->f[7], h->f[8]);
        else k_wall<2><<<cdiv(S.n, 256), 256, 0, h->stream>>>(S, S.n, h->G, leftwall, q0, h->f[6], h->f[7], h->f[8]);
        CKK(h);
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Output-side moments on the device: meanq_fld_cur(totname), output.F90:5229-5486.  The reference adds every particle's
// term to the (2*2+1)^3 box around its cell (idx = idy = idz = 2, idz = 0 in 2D; :189-195), folds the ghosts with
// exchange_current, divides by the (clipped) box volume and, for the averaged quantities, by the weight.  Here each
// particle adds once to its own cell (cell-sorted particles -> contiguous atomics), the box sum is a stencil over
// those cell sums, and only curx needs to travel to the host instead of the whole particle array.
// curx / cury are the scratch arrays, exactly as in the reference.
// ---------------------------------------------------------------------------------------------
struct MeanQ { int kind, comp, species, ind_sign, ratio, weight; };   // kind 0 ch, 1 one, 2 beta, 3 momentum, 4 energy, 5 beta^2
static bool meanq_parse(const char *name, MeanQ *q)
{
    auto is = [&](const char *s) { return strncmp(name, s, 5) == 0; };
    auto comp_of = [](char c) { return c == 'x' ? 0 : c == 'y' ? 1 : 2; };
    q->kind = 0; q->comp = 0; q->species = 3; q->ind_sign = 0; q->ratio = 1; q->weight = 1;
    if (is("tdens")) { q->ratio = 0; return true; }
    if (is("idens")) { q->species = 1; q->ratio = 0; return true; }
    if (is("hdens")) { q->species = 2; q->ind_sign = 1; q->ratio = 0; return true; }
    if (is("ldens")) { q->kind = 1; q->species = 2; q->ind_sign = -1; q->ratio = 0; return true; }
    // the beam densities set no weight (addprty stays 0, output.F90:5289-5297) but are not in the density list of :5470-5471,
    // so the reference's final where(cury /= 0) zeroes them; reproduced as is
    if (is("btden")) { q->ind_sign = -1; q->weight = 0; return true; }
    if (is("biden")) { q->species = 1; q->ind_sign = -1; q->weight = 0; return true; }
    const char f = name[0];
    if ((f == 't' || f == 'e' || f == 'i') && strncmp(name + 1, "bet", 3) == 0 && strchr("xyz", name[4]) && name[4]) {
        q->kind = 2; q->comp = comp_of(name[4]); q->species = f == 't' ? 3 : f == 'e' ? 2 : 1; return true;
    }
    if ((f == 't' || f == 'i') && strncmp(name + 1, "mom", 3) == 0 && strchr("xyz", name[4]) && name[4]) {
        q->kind = 3; q->comp = comp_of(name[4]); q->species = f == 't' ? 3 : 1; return true;
    }
    if (is("eener") || is("iener")) { q->kind = 4; q->species = f == 'e' ? 2 : 1; return true; }
    if ((f == 'e' || f == 'i') && name[1] == 'e' && name[2] == 't' && name[3] && strchr("xyz", name[3]) && name[4] == '2') {
        q->kind = 5; q->comp = comp_of(name[3]); q->species = f == 'e' ? 2 : 1; return true;
    }
    return false;
}
__global__ void __launch_bounds__(256) k_meanq_cells(Species s, int n, MeanQ q, DevGeom G, float *__restrict__ sx, float *__restrict__ sy)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int ind = s.ind[t];
    if ((q.ind_sign > 0 && !(ind > 0)) || (q.ind_sign < 0 && !(ind < 0))) return;
    const float u = s.u[t], v = s.v[t], w = s.w[t], ch = s.ch[t];
    const float gam = 1.f / sqrtf(1.f + u * u + v * v + w * w);
    const float uu = q.comp == 0 ? u : q.comp == 1 ? v : w;
    float ax, ay = ch;
    switch (q.kind) {
    case 0: ax = ch; break;
    case 1: ax = 1.f; break;
    case 2: ax = uu * gam * ch; break;
    case 3: ax = uu * ch; break;
    case 4: ax = (1.f / gam - 1.f) * ch; break;
    default: ax = (uu * gam) * (uu * gam) * ch; break;
    }
    const int i = (int)s.x[t], j = (int)s.y[t], k = G.dim == 3 ? (int)s.z[t] : 1;
    const size_t l = (size_t)(i - 1) + (size_t)G.mx * ((size_t)(j - 1) + (size_t)G.my * (size_t)(k - 1));
    atomicAdd(sx + l, ax);
    if (q.ratio && q.weight) atomicAdd(sy + l, ay);
}
// box sum of the cell sums, then nothing else: the fold and the normalisation follow
__global__ void __launch_bounds__(256) k_meanq_box(const float *__restrict__ sx, const float *__restrict__ sy,
                                                   float *__restrict__ cx, float *__restrict__ cy, int mx, int my, int mz, int idz)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1, k = blockIdx.z + 1;
    if (i > mx) return;
    float a = 0.f, b = 0.f;
    for (int kk = max(k - idz, 1); kk <= min(k + idz, mz); kk++)
        for (int jj = max(j - 2, 1); jj <= min(j + 2, my); jj++)
            for (int ii = max(i - 2, 1); ii <= min(i + 2, mx); ii++) {
                const size_t l = (size_t)(ii - 1) + (size_t)mx * ((size_t)(jj - 1) + (size_t)my * (size_t)(kk - 1));
                a += sx[l]; b += sy[l];
            }
    const size_t l = (size_t)(i - 1) + (size_t)mx * ((size_t)(j - 1) + (size_t)my * (size_t)(k - 1));
    cx[l] = a; cy[l] = b;
}
__global__ void __launch_bounds__(256) k_meanq_norm(float *__restrict__ cx, float *__restrict__ cy, int mx, int my, int mz, int idz, int ratio)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y + 1, k = blockIdx.z + 1;
    if (i > mx) return;
    const int nx = min(i + 2, mx) - max(i - 2, 1) + 1, ny = min(j + 2, my) - max(j - 2, 1) + 1, nz = min(k + idz, mz) - max(k - idz, 1) + 1;
    const float vol = (float)(nx * ny * nz);
    const size_t l = (size_t)(i - 1) + (size_t)mx * ((size_t)(j - 1) + (size_t)my * (size_t)(k - 1));
    float a = cx[l] / vol, b = cy[l] / vol;                  // output.F90:5458-5459
    if (ratio) a = b != 0.f ? a / b : 0.f;                   // :5470-5477
    cx[l] = a; cy[l] = b;
}
int prt_meanq(tgpu_ctx *h, const char *totname)
{
    MeanQ q;
    if (!totname || strlen(totname) < 5) { tgpu_set_error("meanq_fld_cur: totname must have 5 characters"); return TGPU_EINVAL; }
    const bool known = meanq_parse(totname, &q);
    int rc = prt_materialize(h); if (rc) return rc;
    const size_t lot = (size_t)h->G.lot;
    for (int c = 0; c < 3; c++) CK(cudaMemsetAsync(h->f[6 + c], 0, lot * sizeof(float), h->stream));     // :5257-5259
    h->fused_pending = 0;
    if (!known) return 0;                                    // unknown names add nothing (select case falls through)
    float *sx = h->ftmp[0], *sy = h->ftmp[1];
    CK(cudaMemsetAsync(sx, 0, lot * sizeof(float), h->stream)); CK(cudaMemsetAsync(sy, 0, lot * sizeof(float), h->stream));
    for (int s = 0; s < 2; s++) {
        Species &S = h->sp[s];
        if (!S.n || !(q.species & (s ? 2 : 1))) continue;
        k_meanq_cells<<<cdiv(S.n, 256), 256, 0, h->stream>>>(S, S.n, q, h->G, sx, sy); CKK(h);
    }
    const int mx = h->P.mx, my = h->P.my, mz = h->P.dim == 3 ? h->P.mz : 1, idz = h->P.dim == 3 ? 2 : 0;
    dim3 grid(cdiv(mx, 256), my, mz);
    k_meanq_box<<<grid, 256, 0, h->stream>>>(sx, sy, h->f[6], h->f[7], mx, my, mz, idz); CKK(h);
    rc = fld_fold(h); if (rc) return rc;                     // exchange_current(), :5436
    k_meanq_norm<<<grid, 256, 0, h->stream>>>(h->f[6], h->f[7], mx, my, mz, idz, q.ratio); CKK(h);
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Streamed mirror lap, particle side (tgpu_step_mirror): the host array crosses PCIe in both directions at once.
// Four staging quarters (two inbound, two outbound); per chunk:  H2D copy (stream_prt)  ->  AoS->SoA, fused mover +
// deposit, wrap + SoA->AoS (stream_main)  ->  D2H copy (stream_d2h).  Chunk k+1 is inbound while chunk k is pushed and
// chunk k-1 is outbound, so a lap costs max(H2D, D2H) instead of H2D + lap + D2H.
// Only for one rank with all axes periodic (nobody leaves or is discarded): the caller checks.
// ---------------------------------------------------------------------------------------------
int cellrun_move_deposit_range(tgpu_ctx *h, int s, int off, int cnt);
int prt_mirror_stream(tgpu_ctx *h, tgpu_particle *p, int ions, int lecs)
{
    if (ions < 0 || lecs < 0 || ions > h->maxhlf || lecs > h->maxhlf) { tgpu_set_error("bad particle counts"); return TGPU_EINVAL; }
    const size_t quarter = h->stage_particles / 4;
    if (quarter < 1) { tgpu_set_error("staging buffer too small for the streamed mirror lap"); return TGPU_EINVAL; }
    cudaStream_t sm = h->stream_main, sh = h->stream_prt, sd = h->stream_d2h;
    h->sp[0].n = ions; h->sp[1].n = lecs; h->keys_valid = 0;
    h->lazy[0] = h->lazy[1] = 0; h->nphys[0] = ions; h->nphys[1] = lecs;
    const size_t nb = (size_t)h->G.nkeys + TGPU_NBIN_EXTRA;
    CK(cudaMemsetAsync(h->bincount, 0, 2 * nb * sizeof(int32_t), sm));
    for (int b = 0; b < 2; b++) { CK(cudaEventRecord(h->ev_stage_free[b], sm)); CK(cudaEventRecord(h->ev_out_free[b], sd)); }
    int i = 0;
    for (int s = 0; s < 2; s++) {
        Species &S = h->sp[s];
        tgpu_particle *hp = p + (s ? h->maxhlf : 0);
        int done = 0;
        while (done < S.n) {
            const int b = i & 1;
            int chunk = S.n - done; if ((size_t)chunk > quarter) chunk = (int)quarter;
            tgpu_particle *sin = h->stage + (size_t)b * quarter, *sout = h->stage + (size_t)(2 + b) * quarter;
            CK(cudaStreamWaitEvent(sh, h->ev_stage_free[b], 0));
            CK(cudaMemcpyAsync(sin, hp + done, (size_t)chunk * sizeof(tgpu_particle), cudaMemcpyHostToDevice, sh));
            CK(cudaEventRecord(h->ev_stage_full[b], sh));
            CK(cudaStreamWaitEvent(sm, h->ev_stage_full[b], 0));
            k_aos2soa<<<cdiv(chunk, 256), 256, 0, sm>>>(sin, S, done, chunk); CKK(h);
            CK(cudaEventRecord(h->ev_stage_free[b], sm));
            { int rc = cellrun_move_deposit_range(h, s, done, chunk); if (rc) return rc; }
            CK(cudaStreamWaitEvent(sm, h->ev_out_free[b], 0));
            k_soa2aos_wrap<<<cdiv(chunk, 256), 256, 0, sm>>>(sout, S, done, chunk, h->G); CKK(h);
            CK(cudaEventRecord(h->ev_out_full[b], sm));
            CK(cudaStreamWaitEvent(sd, h->ev_out_full[b], 0));
            CK(cudaMemcpyAsync(hp + done, sout, (size_t)chunk * sizeof(tgpu_particle), cudaMemcpyDeviceToHost, sd));
            CK(cudaEventRecord(h->ev_out_free[b], sd));
            done += chunk; i++;
        }
    }
    return 0;          // the caller joins stream_d2h after the field phase
}

// ---------------------------------------------------------------------------------------------
// Output-side particle sub-sampling: the prtl.tot selection of output.F90:3526-3551, modulo(ind/2, stride) == 0.
// Stream compaction on the device so that an output lap moves 1/stride of the particles across PCIe, not all of them.
// Selected ions land at out[0 .. n_ion), electrons at out[capacity .. capacity + n_lec) (order within a species is not the
// array order of the reference; identify particles by (proc, ind)).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_select(Species s, int n, int stride, tgpu_particle *__restrict__ out, int cap, int32_t *__restrict__ counter)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int ind = s.ind[t];
    if ((ind / 2) % stride != 0) return;                       // Fortran integer division truncates, like C
    const int pos = atomicAdd(counter, 1);
    if (pos >= cap) return;
    tgpu_particle p;
    p.x = s.x[t]; p.y = s.y[t]; p.z = s.z[t]; p.u = s.u[t]; p.v = s.v[t]; p.w = s.w[t]; p.ch = s.ch[t];
    p.ind = ind; const int tg = s.tag[t]; p.proc = tg & 0xFFFFFF; p.splitlev = (tg >> 24) & 0xFF;
    out[pos] = p;
}
int prt_select(tgpu_ctx *h, int stride, tgpu_particle *out_host, int capacity, int *n_ion, int *n_lec)
{
    if (stride < 1 || capacity < 0 || !out_host || !n_ion || !n_lec) { tgpu_set_error("select_particles: bad arguments"); return TGPU_EINVAL; }
    int rc = prt_materialize(h); if (rc) return rc;
    const size_t half = h->stage_particles / 2;
    if ((size_t)capacity > half) { tgpu_set_error("select_particles: capacity exceeds the staging buffer (stage_particles / 2)"); return TGPU_EINVAL; }
    CK(cudaMemsetAsync(h->d_small + 64, 0, 2 * sizeof(int32_t), h->stream));
    for (int s = 0; s < 2; s++) {
        Species &S = h->sp[s];
        if (!S.n || !capacity) continue;
        k_select<<<cdiv(S.n, 256), 256, 0, h->stream>>>(S, S.n, stride, h->stage + (size_t)s * half, capacity, h->d_small + 64 + s); CKK(h);
    }
    int32_t cnt[2];
    CK(cudaMemcpyAsync(cnt, h->d_small + 64, sizeof cnt, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (cnt[0] > capacity || cnt[1] > capacity) { tgpu_set_error("select_particles: more selected particles than capacity"); return TGPU_EOVERFLOW; }
    for (int s = 0; s < 2; s++)
        if (cnt[s]) CK(cudaMemcpyAsync(out_host + (size_t)s * capacity, h->stage + (size_t)s * half, (size_t)cnt[s] * sizeof(tgpu_particle), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    *n_ion = cnt[0]; *n_lec = cnt[1];
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Output-side spectra: the per-rank part of save_spectrum (output.F90:380-633) on the device.
//   tgpu_spectrum_gamma_range : local min / max of gamma (:440-455); the host allreduces them (:458-463)
//   tgpu_spectrum             : per species, slice-mean flow velocity (:477-497, 561-578), lab-frame spectrum (:503-511)
//                               and flow-rest-frame spectrum (:513-537), nbins x-slices x gambins logarithmic bins,
//                               returned as this rank's sums BEFORE mpi_allreduce and the division by xgamma (:539-552).
// Histograms are accumulated per block in shared memory (when they fit) and merged with one atomic per bin per block.
// The reference's electron loops start one slot early (a dead record, :562, 583); that slot is not read here.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_gamma_range(Species s, int n, unsigned *__restrict__ mm)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    float g = 1.f;
    if (t < n) { const float u = s.u[t], v = s.v[t], w = s.w[t]; g = sqrtf(1.f + (u * u + v * v + w * w)); }
    float lo = g, hi = g;
    for (int o = 16; o; o >>= 1) { lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o)); }
    if ((threadIdx.x & 31) == 0) { atomicMin(mm, __float_as_uint(lo)); atomicMax(mm + 1, __float_as_uint(hi)); }   // gamma >= 1 > 0
}
struct SpecArgs { int mxcum, mxmin, nbins, gambins; float dxslice, splitratio, lg0, dgam; };
__device__ __forceinline__ bool spec_particle(const Species &s, int t, const SpecArgs &A, int &xbin, float &wgt, float &gam,
                                              float &u, float &v, float &w)
{
    xbin = (int)((s.x[t] + A.mxcum - A.mxmin) / A.dxslice + 1);
    if (xbin < 1 || xbin > A.nbins) return false;
    u = s.u[t]; v = s.v[t]; w = s.w[t];
    const int splitlev = (s.tag[t] >> 24) & 0xFF;
    wgt = powf(A.splitratio, 1.f - (float)splitlev) * s.ch[t];
    gam = sqrtf(1.f + (u * u + v * v + w * w));
    return true;
}
// acc[0..nbins) = numdens, then umean, vmean, wmean sums
__global__ void __launch_bounds__(256) k_spec_means(Species s, int n, SpecArgs A, float *__restrict__ acc)
{
    extern __shared__ float sh[];
    for (int i = threadIdx.x; i < 4 * A.nbins; i += blockDim.x) sh[i] = 0.f;
    __syncthreads();
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    int xbin; float wgt, gam, u, v, w;
    if (t < n && spec_particle(s, t, A, xbin, wgt, gam, u, v, w)) {
        atomicAdd(sh + xbin - 1, wgt); atomicAdd(sh + A.nbins + xbin - 1, u / gam * wgt);
        atomicAdd(sh + 2 * A.nbins + xbin - 1, v / gam * wgt); atomicAdd(sh + 3 * A.nbins + xbin - 1, w / gam * wgt);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 4 * A.nbins; i += blockDim.x) if (sh[i] != 0.f) atomicAdd(acc + i, sh[i]);
}
__global__ void k_spec_norm(float *acc, int nbins)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < nbins) { const float nd = acc[b]; acc[nbins + b] /= nd; acc[2 * nbins + b] /= nd; acc[3 * nbins + b] /= nd; }
}
// spec[0 .. nb*gb) = lab frame, spec[nb*gb .. 2*nb*gb) = flow rest frame; SMEM: per-block shared histogram
template <bool SMEM>
__global__ void __launch_bounds__(256) k_spec_hist(Species s, int n, SpecArgs A, const float *__restrict__ mean, float *__restrict__ spec)
{
    extern __shared__ float sh[];
    const int nh = 2 * A.nbins * A.gambins;
    if (SMEM) { for (int i = threadIdx.x; i < nh; i += blockDim.x) sh[i] = 0.f; __syncthreads(); }
    float *dst = SMEM ? sh : spec;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    int xbin; float wgt, gam, u, v, w;
    if (t < n && spec_particle(s, t, A, xbin, wgt, gam, u, v, w)) {
        int gbin = (int)((log10f(gam - 1.f) - A.lg0) / A.dgam + 1);
        if (gbin >= 1 && gbin <= A.gambins) atomicAdd(dst + (xbin - 1) + A.nbins * (gbin - 1), wgt);
        const float vx = mean[A.nbins + xbin - 1], vy = mean[2 * A.nbins + xbin - 1], vz = mean[3 * A.nbins + xbin - 1];
        const float vr = sqrtf(vx * vx + vy * vy + vz * vz), gvr = 1.f / sqrtf(1.f - vr * vr);
        const float up = -vx * gvr * gam + (1 + (gvr - 1) * vx * vx / (vr * vr)) * u + (gvr - 1) * vx * vy / (vr * vr) * v
                         + (gvr - 1) * vx * vz / (vr * vr) * w;
        const float vp = -vy * gvr * gam + (gvr - 1) * vx * vy / (vr * vr) * u + (1 + (gvr - 1) * vy * vy / (vr * vr)) * v
                         + (gvr - 1) * vy * vz / (vr * vr) * w;
        const float wp = -vz * gvr * gam + (gvr - 1) * vx * vz / (vr * vr) * u + (gvr - 1) * vy * vz / (vr * vr) * v
                         + (1 + (gvr - 1) * vz * vz / (vr * vr)) * w;
        const float gp = sqrtf(1.f + (up * up + vp * vp + wp * wp));
        gbin = (int)((log10f(gp - 1.f) - A.lg0) / A.dgam + 1);
        if (gbin >= 1 && gbin <= A.gambins) atomicAdd(dst + A.nbins * A.gambins + (xbin - 1) + A.nbins * (gbin - 1), wgt);
    }
    if (SMEM) { __syncthreads(); for (int i = threadIdx.x; i < nh; i += blockDim.x) if (sh[i] != 0.f) atomicAdd(spec + i, sh[i]); }
}
int prt_gamma_range(tgpu_ctx *h, float *gammin, float *gammax)
{
    int rc = prt_materialize(h); if (rc) return rc;
    unsigned init[2] = {0x3f800000u, 0x3f800000u}, out[2];             // gammin = gammax = 1 (:424-425)
    unsigned *mm = (unsigned *)(h->d_small + 66);
    CK(cudaMemcpyAsync(mm, init, sizeof init, cudaMemcpyHostToDevice, h->stream));
    for (int s = 0; s < 2; s++) if (h->sp[s].n) { k_gamma_range<<<cdiv(h->sp[s].n, 256), 256, 0, h->stream>>>(h->sp[s], h->sp[s].n, mm); CKK(h); }
    CK(cudaMemcpyAsync(out, mm, sizeof out, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    memcpy(gammin, &out[0], 4); memcpy(gammax, &out[1], 4);
    return 0;
}
int prt_spectrum(tgpu_ctx *h, float gammin, float gammax, int mx0, float splitratio, int nbins, int gambins,
                 float *specp, float *spece, float *specprest, float *specerest)
{
    if (nbins < 1 || gambins < 1 || !specp || !spece || !specprest || !specerest || !(gammax > 1.f)) {
        tgpu_set_error("spectrum: bad arguments (need nbins, gambins >= 1 and gammax > 1)"); return TGPU_EINVAL;
    }
    int rc = prt_materialize(h); if (rc) return rc;
    SpecArgs A;
    A.mxcum = h->P.mxcum; A.mxmin = 3; A.nbins = nbins; A.gambins = gambins; A.splitratio = splitratio;
    A.dxslice = 1.f * ((mx0 - 2) - A.mxmin) / nbins;                                    // :432-438
    gammin = gammin > 1.f + 1e-6f ? gammin : 1.f + 1e-6f;                               // :465
    A.lg0 = log10f(gammin - 1.f); A.dgam = (log10f(gammax - 1.f) - A.lg0) / gambins;   // :466
    const size_t nh = 2 * (size_t)nbins * gambins, nacc = 4 * (size_t)nbins;
    struct Scratch { float *p = nullptr; ~Scratch() { if (p) cudaFree(p); } } scratch;     // freed on every return path
    CK(cudaMalloc(&scratch.p, (nh + nacc) * sizeof(float)));
    float *buf = scratch.p;
    float *hout[2][2] = {{specp, specprest}, {spece, specerest}};
    for (int s = 0; s < 2; s++) {
        Species &S = h->sp[s];
        CK(cudaMemsetAsync(buf, 0, (nh + nacc) * sizeof(float), h->stream));
        float *acc = buf + nh;
        if (S.n) {
            k_spec_means<<<cdiv(S.n, 256), 256, nacc * sizeof(float), h->stream>>>(S, S.n, A, acc); CKK(h);
            k_spec_norm<<<cdiv(nbins, 64), 64, 0, h->stream>>>(acc, nbins); CKK(h);
            if (nh * sizeof(float) <= 40 * 1024) k_spec_hist<true><<<cdiv(S.n, 256), 256, nh * sizeof(float), h->stream>>>(S, S.n, A, acc, buf);
            else k_spec_hist<false><<<cdiv(S.n, 256), 256, 0, h->stream>>>(S, S.n, A, acc, buf);
            CKK(h);
        }
        CK(cudaMemcpyAsync(hout[s][0], buf, nh / 2 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(hout[s][1], buf + nh / 2, nh / 2 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    return 0;
}
