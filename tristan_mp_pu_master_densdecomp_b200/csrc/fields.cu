// Field-side kernels: Yee half/full steps, ghost refresh, current fold, digital filters.
// Compiled with -fmad=false: every fp32 operation is rounded separately, in the order the
// reference writes it, so these kernels are bit-exact against the CPU restatement.
//   advance_b_halfstep  code/fields.F90:586-728      advance_e_fullstep  code/fields.F90:739-870
//   bc_b1 / bc_e1       code/fieldboundaries.F90:181-263, 306-392
//   exchange_current    code/fieldboundaries.F90:1768-2189
//   apply_filter1_opt   code/filter.F90:8-221        apply_filter2_opt   code/optimized_filters.F90:9-227
#include "tgpu_internal.h"

#define LIDX(i, j, k) ((size_t)((i)-1) + (size_t)mx * ((size_t)((j)-1) + (size_t)my * (size_t)((k)-1)))

// ---------------------------------------------------------------------------------------------
// stencil index ranges (fields.F90:599-669 for B, :752-819 for E)
// ---------------------------------------------------------------------------------------------
static void axis_info(const tgpu_ctx *h, int axis, int *m, int *g, int *per, int *size, int *pos)
{
    const tgpu_params &P = h->P;
    *m = axis == 0 ? P.mx : axis == 1 ? P.my : P.mz;
    *g = (axis == 2 ? P.nghostz : P.nghost) / 2;
    *per = axis == 0 ? P.periodicx : axis == 1 ? P.periodicy : P.periodicz;
    *size = axis == 0 ? P.sizex : axis == 1 ? P.sizey : P.sizez;
    *pos = axis == 0 ? P.rank % P.sizex : axis == 1 ? (P.rank % (P.sizex * P.sizey)) / P.sizex : P.rank / (P.sizex * P.sizey);
}
static void range_b(const tgpu_ctx *h, int axis, int *a1, int *a2)
{
    int m, g, per, sz, pos; axis_info(h, axis, &m, &g, &per, &sz, &pos);
    *a1 = g + 1; *a2 = m - (g + 1);
    if (!per) {
        if (pos == 0) { *a1 = 1; *a2 = m - (g + 1); }
        if (pos == sz - 1) { *a1 = g + 1; *a2 = m - 1; }
        if (axis == 2 ? (h->size0 == 1) : (h->size0 == 1 || sz == 1)) { *a1 = 1; *a2 = m - 1; }
    }
}
static void range_e(const tgpu_ctx *h, int axis, int *a1, int *a2)
{
    int m, g, per, sz, pos; axis_info(h, axis, &m, &g, &per, &sz, &pos);
    *a1 = g + 1; *a2 = m - (g + 1);
    if (!per) {
        if (pos == 0) { *a1 = g; *a2 = m - (g + 1); }
        if (pos == sz - 1) { *a1 = g + 1; *a2 = m; }
        if (axis == 2 ? (h->size0 == 1) : (h->size0 == 1 || sz == 1)) { *a1 = g; *a2 = m; }
    }
    if (axis == 2) { *a1 = g; *a2 = m; }      // fields.F90:818-819
}

struct Range3 { int i1, i2, j1, j2, k1, k2; };

// One thread per cell, x fastest => fully coalesced; each array is read/written once per launch.
template <int DIM>
__global__ void __launch_bounds__(256) k_bhalf(float *__restrict__ bx, float *__restrict__ by, float *__restrict__ bz,
                                               const float *__restrict__ ex, const float *__restrict__ ey,
                                               const float *__restrict__ ez, int mx, int my, Range3 r, float cnst)
{
    int i = r.i1 + blockIdx.x * blockDim.x + threadIdx.x;
    int j = r.j1 + blockIdx.y;
    int k = r.k1 + blockIdx.z;
    if (i > r.i2) return;
    size_t l = LIDX(i, j, k), lip = l + 1, ljp = l + mx;
    if (DIM == 3) {
        size_t lkp = l + (size_t)mx * my;
        bx[l] = bx[l] + cnst * (ey[lkp] - ey[l] - ez[ljp] + ez[l]);
        by[l] = by[l] + cnst * (ez[lip] - ez[l] - ex[lkp] + ex[l]);
        bz[l] = bz[l] + cnst * (ex[ljp] - ex[l] - ey[lip] + ey[l]);
    } else {
        bx[l] = bx[l] + cnst * (-ez[ljp] + ez[l]);
        by[l] = by[l] + cnst * (ez[lip] - ez[l]);
        bz[l] = bz[l] + cnst * (ex[ljp] - ex[l] - ey[lip] + ey[l]);
    }
}

template <int DIM>
__global__ void __launch_bounds__(256) k_efull(float *__restrict__ ex, float *__restrict__ ey, float *__restrict__ ez,
                                               const float *__restrict__ bx, const float *__restrict__ by,
                                               const float *__restrict__ bz, int mx, int my, Range3 r, float cnst)
{
    int i = r.i1 + blockIdx.x * blockDim.x + threadIdx.x;
    int j = r.j1 + blockIdx.y;
    int k = r.k1 + blockIdx.z;
    if (i > r.i2) return;
    size_t l = LIDX(i, j, k), lim = l - 1, ljm = l - mx;
    if (DIM == 3) {
        size_t lkm = l - (size_t)mx * my;
        ex[l] = ex[l] + cnst * (by[lkm] - by[l] - bz[ljm] + bz[l]);
        ey[l] = ey[l] + cnst * (bz[lim] - bz[l] - bx[lkm] + bx[l]);
        ez[l] = ez[l] + cnst * (bx[ljm] - bx[l] - by[lim] + by[l]);
    } else {
        ex[l] = ex[l] + cnst * (-bz[ljm] + bz[l]);
        ey[l] = ey[l] + cnst * (bz[lim] - bz[l]);
        ez[l] = ez[l] + cnst * (bx[ljm] - bx[l] - by[lim] + by[l]);
    }
}

// ---------------------------------------------------------------------------------------------
// 4th-order `_42` solver (highorder = 1): fields.F90:1039-1212 (B half step), 1223-1361 (E full step).
// coef1 = 9/8, coef2 = -1/24 of the 2nd-order constant; on an open x axis the two outermost planes get the 2nd-order
// update (:1158-1190, 1327-1357).  EDGE launches run that update on the plane pair (i = ia, ib).
// ---------------------------------------------------------------------------------------------
static void range42_b(const tgpu_ctx *h, int axis, int *a1, int *a2)
{
    int m, g, per, sz, pos; axis_info(h, axis, &m, &g, &per, &sz, &pos);
    if (per) { *a1 = g + 1; *a2 = m - (g + 1); } else { *a1 = g; *a2 = m - g; }
    if (axis == 0 && h->P.wall_i2 > 0 && h->P.wall_i2 < *a2) *a2 = h->P.wall_i2;
}
static void range42_e(const tgpu_ctx *h, int axis, int *a1, int *a2)
{
    int m, g, per, sz, pos; axis_info(h, axis, &m, &g, &per, &sz, &pos);
    *a1 = g + 1; *a2 = per ? m - (g + 1) : m - 1;
    if (axis == 0 && h->P.wall_i2 > 0 && h->P.wall_i2 < *a2) *a2 = h->P.wall_i2;
}
template <int DIM, bool EDGE>
__global__ void __launch_bounds__(256) k_bhalf42(float *__restrict__ bx, float *__restrict__ by, float *__restrict__ bz,
                                                 const float *__restrict__ ex, const float *__restrict__ ey,
                                                 const float *__restrict__ ez, int mx, int my, Range3 r, float coef1, float coef2)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = EDGE ? (t == 0 ? r.i1 : r.i2) : r.i1 + t;
    const int j = r.j1 + blockIdx.y, k = r.k1 + blockIdx.z;
    if (EDGE ? t > 1 : i > r.i2) return;
    const size_t sx = 1, sy = mx, sz = (size_t)mx * my;
    const size_t l = LIDX(i, j, k);
    if (EDGE) {                       // coef1 carries the plain 2nd-order constant
        if (DIM == 3) {
            bx[l] = bx[l] + coef1 * (ey[l + sz] - ey[l] - ez[l + sy] + ez[l]);
            by[l] = by[l] + coef1 * (ez[l + sx] - ez[l] - ex[l + sz] + ex[l]);
        } else {
            bx[l] = bx[l] + coef1 * (-ez[l + sy] + ez[l]);
            by[l] = by[l] + coef1 * (ez[l + sx] - ez[l]);
        }
        bz[l] = bz[l] + coef1 * (ex[l + sy] - ex[l] - ey[l + sx] + ey[l]);
        return;
    }
    if (DIM == 3) {
        bx[l] = bx[l] + coef1 * (ey[l + sz] - ey[l] - ez[l + sy] + ez[l])
                      + coef2 * (ey[l + 2 * sz] - ey[l - sz] - ez[l + 2 * sy] + ez[l - sy]);
        by[l] = by[l] + coef1 * (ez[l + sx] - ez[l] - ex[l + sz] + ex[l])
                      + coef2 * (ez[l + 2 * sx] - ez[l - sx] - ex[l + 2 * sz] + ex[l - sz]);
    } else {
        bx[l] = bx[l] + coef1 * (-ez[l + sy] + ez[l]) + coef2 * (-ez[l + 2 * sy] + ez[l - sy]);
        by[l] = by[l] + coef1 * (ez[l + sx] - ez[l]) + coef2 * (ez[l + 2 * sx] - ez[l - sx]);
    }
    bz[l] = bz[l] + coef1 * (ex[l + sy] - ex[l] - ey[l + sx] + ey[l])
                  + coef2 * (ex[l + 2 * sy] - ex[l - sy] - ey[l + 2 * sx] + ey[l - sx]);
}
template <int DIM, bool EDGE>
__global__ void __launch_bounds__(256) k_efull42(float *__restrict__ ex, float *__restrict__ ey, float *__restrict__ ez,
                                                 const float *__restrict__ bx, const float *__restrict__ by,
                                                 const float *__restrict__ bz, int mx, int my, Range3 r, float coef1, float coef2)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = EDGE ? (t == 0 ? r.i1 : r.i2) : r.i1 + t;
    const int j = r.j1 + blockIdx.y, k = r.k1 + blockIdx.z;
    if (EDGE ? t > 1 : i > r.i2) return;
    const size_t sx = 1, sy = mx, sz = (size_t)mx * my;
    const size_t l = LIDX(i, j, k);
    if (EDGE) {
        if (DIM == 3) {
            ex[l] = ex[l] + coef1 * (by[l - sz] - by[l] - bz[l - sy] + bz[l]);
            ey[l] = ey[l] + coef1 * (bz[l - sx] - bz[l] - bx[l - sz] + bx[l]);
        } else {
            ex[l] = ex[l] + coef1 * (-bz[l - sy] + bz[l]);
            ey[l] = ey[l] + coef1 * (bz[l - sx] - bz[l]);
        }
        ez[l] = ez[l] + coef1 * (bx[l - sy] - bx[l] - by[l - sx] + by[l]);
        return;
    }
    if (DIM == 3) {
        ex[l] = ex[l] + coef1 * (by[l - sz] - by[l] - bz[l - sy] + bz[l])
                      + coef2 * (by[l - 2 * sz] - by[l + sz] - bz[l - 2 * sy] + bz[l + sy]);
        ey[l] = ey[l] + coef1 * (bz[l - sx] - bz[l] - bx[l - sz] + bx[l])
                      + coef2 * (bz[l - 2 * sx] - bz[l + sx] - bx[l - 2 * sz] + bx[l + sz]);
    } else {
        ex[l] = ex[l] + coef1 * (-bz[l - sy] + bz[l]) + coef2 * (-bz[l - 2 * sy] + bz[l + sy]);
        ey[l] = ey[l] + coef1 * (bz[l - sx] - bz[l]) + coef2 * (bz[l - 2 * sx] - bz[l + sx]);
    }
    ez[l] = ez[l] + coef1 * (bx[l - sy] - bx[l] - by[l - sx] + by[l])
                  + coef2 * (bx[l - 2 * sy] - bx[l + sy] - by[l - 2 * sx] + by[l + sx]);
}
static int fld_step42(tgpu_ctx *h, bool is_e)
{
    Range3 r; r.k1 = r.k2 = 1;
    if (is_e) { range42_e(h, 0, &r.i1, &r.i2); range42_e(h, 1, &r.j1, &r.j2); if (h->P.dim == 3) range42_e(h, 2, &r.k1, &r.k2); }
    else { range42_b(h, 0, &r.i1, &r.i2); range42_b(h, 1, &r.j1, &r.j2); if (h->P.dim == 3) range42_b(h, 2, &r.k1, &r.k2); }
    const float base = is_e ? h->P.corr * h->P.c : h->P.corr * (.5f * h->P.c);
    const float coef1 = is_e ? 9.f / 8.f * h->P.corr * h->P.c : 9.f / 8.f * h->P.corr * (.5f * h->P.c);
    const float coef2 = is_e ? -1.f / 24.f * h->P.corr * h->P.c : -1.f / 24.f * h->P.corr * (.5f * h->P.c);
    float *e0 = h->f[0], *e1 = h->f[1], *e2 = h->f[2], *b0 = h->f[3], *b1 = h->f[4], *b2 = h->f[5];
    const int mx = h->P.mx, my = h->P.my;
    dim3 grid(cdiv(r.i2 - r.i1 + 1, 256), r.j2 - r.j1 + 1, r.k2 - r.k1 + 1);
    if (is_e) {
        if (h->P.dim == 3) k_efull42<3, false><<<grid, 256, 0, h->stream>>>(e0, e1, e2, b0, b1, b2, mx, my, r, coef1, coef2);
        else k_efull42<2, false><<<grid, 256, 0, h->stream>>>(e0, e1, e2, b0, b1, b2, mx, my, r, coef1, coef2);
    } else {
        if (h->P.dim == 3) k_bhalf42<3, false><<<grid, 256, 0, h->stream>>>(b0, b1, b2, e0, e1, e2, mx, my, r, coef1, coef2);
        else k_bhalf42<2, false><<<grid, 256, 0, h->stream>>>(b0, b1, b2, e0, e1, e2, mx, my, r, coef1, coef2);
    }
    CKK(h);
    if (!h->P.periodicx) {
        Range3 e = r; dim3 ge(1, grid.y, grid.z);
        if (is_e) { e.i1 = 2; e.i2 = mx; } else { e.i1 = 1; e.i2 = mx - 1; }
        if (is_e) {
            if (h->P.dim == 3) k_efull42<3, true><<<ge, 32, 0, h->stream>>>(e0, e1, e2, b0, b1, b2, mx, my, e, base, 0.f);
            else k_efull42<2, true><<<ge, 32, 0, h->stream>>>(e0, e1, e2, b0, b1, b2, mx, my, e, base, 0.f);
        } else {
            if (h->P.dim == 3) k_bhalf42<3, true><<<ge, 32, 0, h->stream>>>(b0, b1, b2, e0, e1, e2, mx, my, e, base, 0.f);
            else k_bhalf42<2, true><<<ge, 32, 0, h->stream>>>(b0, b1, b2, e0, e1, e2, mx, my, e, base, 0.f);
        }
        CKK(h);
    }
    h->need_prim = 1;
    return 0;
}

int fld_bhalf(tgpu_ctx *h)
{
    if (h->P.highorder) return fld_step42(h, false);          // dispatcher, fields.F90:1407-1417
    Range3 r; r.k1 = r.k2 = 1;
    range_b(h, 0, &r.i1, &r.i2); range_b(h, 1, &r.j1, &r.j2);
    if (h->P.dim == 3) range_b(h, 2, &r.k1, &r.k2);
    const float cnst = h->P.corr * (.5f * h->P.c);
    dim3 grid(cdiv(r.i2 - r.i1 + 1, 256), r.j2 - r.j1 + 1, r.k2 - r.k1 + 1);
    if (h->P.dim == 3)
        k_bhalf<3><<<grid, 256, 0, h->stream>>>(h->f[3], h->f[4], h->f[5], h->f[0], h->f[1], h->f[2], h->P.mx, h->P.my, r, cnst);
    else
        k_bhalf<2><<<grid, 256, 0, h->stream>>>(h->f[3], h->f[4], h->f[5], h->f[0], h->f[1], h->f[2], h->P.mx, h->P.my, r, cnst);
    CKK(h);
    h->need_prim = 1;
    return 0;
}

int fld_efull(tgpu_ctx *h)
{
    if (h->P.highorder) return fld_step42(h, true);           // dispatcher, fields.F90:1429-1440
    Range3 r; r.k1 = r.k2 = 1;
    range_e(h, 0, &r.i1, &r.i2); range_e(h, 1, &r.j1, &r.j2);
    if (h->P.dim == 3) range_e(h, 2, &r.k1, &r.k2);
    // guard the reference's out-of-bounds reads at the array edge (i-1 with i = g >= 2 is fine)
    const float cnst = h->P.corr * h->P.c;
    dim3 grid(cdiv(r.i2 - r.i1 + 1, 256), r.j2 - r.j1 + 1, r.k2 - r.k1 + 1);
    if (h->P.dim == 3)
        k_efull<3><<<grid, 256, 0, h->stream>>>(h->f[0], h->f[1], h->f[2], h->f[3], h->f[4], h->f[5], h->P.mx, h->P.my, r, cnst);
    else
        k_efull<2><<<grid, 256, 0, h->stream>>>(h->f[0], h->f[1], h->f[2], h->f[3], h->f[4], h->f[5], h->P.mx, h->P.my, r, cnst);
    CKK(h);
    h->need_prim = 1;
    return 0;
}

int fld_reset(tgpu_ctx *h)
{
    for (int c = 0; c < 3; c++) CK(cudaMemsetAsync(h->f[6 + c], 0, (size_t)h->G.lot * sizeof(float), h->stream));
    return 0;
}

// e += cur over whole arrays (fields.F90:1391-1393); with SHADOW: cur += shadow; shadow = 0
__global__ void __launch_bounds__(256) k_add3(float *__restrict__ a0, float *__restrict__ a1, float *__restrict__ a2,
                                              float *__restrict__ b0, float *__restrict__ b1, float *__restrict__ b2,
                                              size_t n, int zero_src)
{
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t l = (size_t)blockIdx.x * blockDim.x + threadIdx.x; l < n; l += stride) {
        a0[l] = a0[l] + b0[l]; a1[l] = a1[l] + b1[l]; a2[l] = a2[l] + b2[l];
        if (zero_src) { b0[l] = 0.f; b1[l] = 0.f; b2[l] = 0.f; }
    }
}
int fld_add(tgpu_ctx *h)
{
    size_t n = (size_t)h->G.lot;
    int blocks = (int)((n + 255) / 256); if (blocks > 148 * 16) blocks = 148 * 16;
    k_add3<<<blocks, 256, 0, h->stream>>>(h->f[0], h->f[1], h->f[2], h->f[6], h->f[7], h->f[8], n, 0);
    CKK(h);
    h->need_prim = 1;
    return 0;
}
// cur += shadow; shadow = 0, where shadow is the fused mover's tiled deposit target (cellrun.cu row_index): each 4x4 (y,z)
// tile of an x-plane is 16 contiguous floats, y fastest.  One thread = (i, four consecutive y, one z): one 128-bit load per
// component from the tile (lanes = consecutive i), twelve independent read-modify-writes of the Fortran-order arrays,
// each coalesced along x.  All loads are issued before the first store.
__global__ void __launch_bounds__(256) k_add_shadow_tiled(float *__restrict__ c0, float *__restrict__ c1, float *__restrict__ c2,
                                                          float4 *__restrict__ s0, float4 *__restrict__ s1, float4 *__restrict__ s2,
                                                          int mx, int my, int mz, int nty)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x, ty = blockIdx.y, K = blockIdx.z;
    if (i >= mx) return;
    const size_t a = (((size_t)((K >> 2) * nty + ty) * mx + i) << 2) + (K & 3);      // in float4 units
    const float4 v0 = s0[a], v1 = s1[a], v2 = s2[a];
    const float sv[3][4] = {{v0.x, v0.y, v0.z, v0.w}, {v1.x, v1.y, v1.z, v1.w}, {v2.x, v2.y, v2.z, v2.w}};
    float *const cc[3] = {c0, c1, c2};
    const size_t l0 = (size_t)i + (size_t)mx * (4 * ty + (size_t)my * K);
    float cv[3][4];
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
        for (int q = 0; q < 4; q++) cv[c][q] = (4 * ty + q < my) ? cc[c][l0 + (size_t)q * mx] : 0.f;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    s0[a] = zero; s1[a] = zero; s2[a] = zero;
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
        for (int q = 0; q < 4; q++) if (4 * ty + q < my) cc[c][l0 + (size_t)q * mx] = cv[c][q] + sv[c][q];
}
int fld_add_shadow(tgpu_ctx *h)
{
#if !TGPU_SHADOW_TILED
    {
        size_t n = (size_t)h->G.lot;
        int blocks = (int)((n + 255) / 256); if (blocks > 148 * 16) blocks = 148 * 16;
        k_add3<<<blocks, 256, 0, h->stream>>>(h->f[6], h->f[7], h->f[8], h->shadow[0], h->shadow[1], h->shadow[2], n, 1);
        CKK(h);
        return 0;
    }
#endif
    dim3 grid((h->G.mx + 127) / 128, h->nty, h->G.mz);
    k_add_shadow_tiled<<<grid, 128, 0, h->stream>>>(h->f[6], h->f[7], h->f[8], (float4 *)h->shadow[0], (float4 *)h->shadow[1],
                                                    (float4 *)h->shadow[2], h->G.mx, h->G.my, h->G.mz, h->nty);
    CKK(h);
    return 0;
}

// node-centred fields for the 3D shaped movers: particles_movedeposit.F90:395-404 / 658-667 / 982-991
// (cshift is circular).  Quirk Q2: bx_p, by_p are not averaged in k.  The six components of a node are
// interleaved (ex,ey,ez,bx | by,bz,0,0) = 32 B so that the gather issues two 128-bit loads per node.
__global__ void __launch_bounds__(256) k_primal(const float *__restrict__ ex, const float *__restrict__ ey,
                                                const float *__restrict__ ez, const float *__restrict__ bx,
                                                const float *__restrict__ by, const float *__restrict__ bz,
                                                float4 *__restrict__ prim8, int mx, int my, int mz, int q2)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
    int j = blockIdx.y + 1, k = blockIdx.z + 1;
    if (i > mx) return;
    int im = i == 1 ? mx : i - 1, jm = j == 1 ? my : j - 1, km = k == 1 ? mz : k - 1;
    size_t l = LIDX(i, j, k);
    float4 lo, hi;
    lo.x = 0.5f * (ex[l] + ex[LIDX(im, j, k)]);
    lo.y = 0.5f * (ey[l] + ey[LIDX(i, jm, k)]);
    lo.z = 0.5f * (ez[l] + ez[LIDX(i, j, km)]);
    float bxp = 0.5f * (bx[l] + bx[LIDX(i, jm, k)]);
    float byp = 0.5f * (by[l] + by[LIDX(im, j, k)]);
    if (!q2) {
        float bxk = 0.5f * (bx[LIDX(i, j, km)] + bx[LIDX(i, jm, km)]);
        float byk = 0.5f * (by[LIDX(i, j, km)] + by[LIDX(im, j, km)]);
        bxp = 0.5f * (bxp + bxk); byp = 0.5f * (byp + byk);
    }
    lo.w = bxp; hi.x = byp;
    hi.y = 0.5f * (0.5f * (bz[l] + bz[LIDX(im, j, k)]) + 0.5f * (bz[LIDX(i, jm, k)] + bz[LIDX(im, jm, k)]));
    hi.z = 0.f; hi.w = 0.f;
    prim8[2 * l] = lo; prim8[2 * l + 1] = hi;
}
int fld_primal(tgpu_ctx *h)
{
    if (!h->need_prim) return 0;
    dim3 grid(cdiv(h->P.mx, 256), h->P.my, h->P.mz);
    k_primal<<<grid, 256, 0, h->stream>>>(h->f[0], h->f[1], h->f[2], h->f[3], h->f[4], h->f[5], h->prim8, h->P.mx, h->P.my,
                                         h->P.mz, (h->P.quirks & TGPU_Q2_BXBY_NO_KAVG) != 0);
    CKK(h);
    h->need_prim = 0;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// box copy / pack / unpack over three component arrays at once
// ---------------------------------------------------------------------------------------------
struct Box { int lo[3]; int n[3]; };
struct Arr3 { float *a[3]; };

__device__ __forceinline__ void box_decode(const Box &b, size_t idx, int &c, int &i, int &j, int &k)
{
    size_t vol = (size_t)b.n[0] * b.n[1] * b.n[2];
    c = (int)(idx / vol); size_t r = idx - (size_t)c * vol;
    i = (int)(r % b.n[0]); r /= b.n[0];
    j = (int)(r % b.n[1]); k = (int)(r / b.n[1]);
}
// mode 0: dst = src ; mode 1: dst += src
__global__ void __launch_bounds__(256) k_box_copy(Arr3 A, Box src, Box dst, int mx, int my, int mode, size_t total)
{
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    int c, i, j, k; box_decode(src, idx, c, i, j, k);
    size_t ls = LIDX(src.lo[0] + i, src.lo[1] + j, src.lo[2] + k);
    size_t ld = LIDX(dst.lo[0] + i, dst.lo[1] + j, dst.lo[2] + k);
    float v = A.a[c][ls];
    if (mode) A.a[c][ld] = A.a[c][ld] + v; else A.a[c][ld] = v;
}
__global__ void __launch_bounds__(256) k_box_get(Arr3 A, Box src, float *__restrict__ buf, int mx, int my, size_t total)
{
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    int c, i, j, k; box_decode(src, idx, c, i, j, k);
    buf[idx] = A.a[c][LIDX(src.lo[0] + i, src.lo[1] + j, src.lo[2] + k)];
}
__global__ void __launch_bounds__(256) k_box_put(Arr3 A, Box dst, const float *__restrict__ buf, int mx, int my, int mode, size_t total)
{
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    int c, i, j, k; box_decode(dst, idx, c, i, j, k);
    size_t ld = LIDX(dst.lo[0] + i, dst.lo[1] + j, dst.lo[2] + k);
    if (mode) A.a[c][ld] = A.a[c][ld] + buf[idx]; else A.a[c][ld] = buf[idx];
}

static Box full_box(const tgpu_ctx *h)
{
    Box b; b.lo[0] = b.lo[1] = b.lo[2] = 1; b.n[0] = h->P.mx; b.n[1] = h->P.my; b.n[2] = h->P.mz; return b;
}
static size_t box_total(const Box &b) { return (size_t)3 * b.n[0] * b.n[1] * b.n[2]; }

// Move a box of three arrays from this rank's `src` to the `dst` box of the neighbour that lies in
// direction `dir_to` (and receive the matching box from the opposite neighbour).  Local when the
// axis has one rank.  recv_ok = 0 skips the unpack (open boundary, edge rank).
static int box_shift(tgpu_ctx *h, Arr3 A, Box src, Box dst, int axis, int dir_to, int mode, int recv_ok, bool buffered = false)
{
    int m, g, per, sz, pos; axis_info(h, axis, &m, &g, &per, &sz, &pos);
    size_t total = box_total(src);
    if (total == 0) return 0;
    if (sz == 1) {
        if (!recv_ok) return 0;
        if (buffered) {
            // source and destination overlap (an axis narrower than its ghost zone): go through a buffer as the reference
            // does (bufferin1y / bufferin2y, fieldboundaries.F90:1931-1948)
            if (total > h->halo_floats) { tgpu_set_error("halo scratch too small"); return TGPU_EINVAL; }
            k_box_get<<<cdiv(total, 256), 256, 0, h->stream>>>(A, src, h->halo, h->P.mx, h->P.my, total); CKK(h);
            k_box_put<<<cdiv(total, 256), 256, 0, h->stream>>>(A, dst, h->halo, h->P.mx, h->P.my, mode, total); CKK(h);
            return 0;
        }
        k_box_copy<<<cdiv(total, 256), 256, 0, h->stream>>>(A, src, dst, h->P.mx, h->P.my, mode, total);
        CKK(h);
        return 0;
    }
    if (2 * total > h->halo_floats) { tgpu_set_error("halo scratch too small"); return TGPU_EINVAL; }
    float *sb = h->halo, *rb = h->halo + total;
    k_box_get<<<cdiv(total, 256), 256, 0, h->stream>>>(A, src, sb, h->P.mx, h->P.my, total);
    CKK(h);
    int to = topo_neighbour(h->P.rank, h->P.sizex, h->P.sizey, h->P.sizez, 2 * axis + (dir_to > 0 ? 1 : 0));
    int from = topo_neighbour(h->P.rank, h->P.sizex, h->P.sizey, h->P.sizez, 2 * axis + (dir_to > 0 ? 0 : 1));
    int rc = comm_sendrecv(h, sb, total * sizeof(float), to, rb, total * sizeof(float), from);
    if (rc) return rc;
    if (recv_ok) {
        k_box_put<<<cdiv(total, 256), 256, 0, h->stream>>>(A, dst, rb, h->P.mx, h->P.my, mode, total);
        CKK(h);
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Halo step over peer memory (comm.cu comm_peer_setup).  Exchange number s of the run, on a split axis:
//   1. k_sig_set:  READY = s in my own signal words, stream-ordered after everything I produced for this step;
//   2. k_box_pull: for each of my two neighbours, wait until ITS READY reaches s (ld.acquire.sys over NVLink), then read the
//      layers it would have sent me straight out of its arrays and store (ghost refresh) or add (current fold) them;
//   3. k_sig_sync: PULLED = s in my words, then wait until both neighbours' PULLED reach s -- they read my arrays in
//      their step 2, and what follows on my stream may overwrite those layers.
// Every rank runs the same sequence of steps (SPMD), so one counter per rank is enough.  No pack / unpack buffers and no
// NCCL call: three small launches per axis.  A wait gives up after 5 s and raises TGPU_SIG_TIMEOUT instead of hanging.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p)
{
    uint32_t v; asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t;
}
__device__ __forceinline__ void sig_wait(const uint32_t *flag, uint32_t seq, uint32_t *timeout)
{
    const unsigned long long t0 = global_ns();
    while ((int)(ld_acquire_sys(flag) - seq) < 0) {
        if (global_ns() - t0 > 5000000000ull) { *timeout = 1u; return; }
        __nanosleep(64);
    }
}
__global__ void k_sig_set(uint32_t *flag, uint32_t seq)
{
    __threadfence_system();
    st_release_sys(flag, seq);
}
__global__ void k_sig_sync(uint32_t *sig, uint32_t seq, const uint32_t *nb_a, const uint32_t *nb_b)
{
    __threadfence_system();
    st_release_sys(sig + TGPU_SIG_PULLED, seq);
    if (nb_a) sig_wait(nb_a + TGPU_SIG_PULLED, seq, sig + TGPU_SIG_TIMEOUT);
    if (nb_b) sig_wait(nb_b + TGPU_SIG_PULLED, seq, sig + TGPU_SIG_TIMEOUT);
}
struct PullBox {
    Box src, dst;              // src in the neighbour's index space, dst in mine; same extents
    const float *rem[3];       // the neighbour's three arrays
    int rmx, rmy;              // its x and y extents (strides)
    const uint32_t *nb_sig;    // its signal words
    unsigned long long total;  // 3 * volume; 0 = nothing to pull from this side
};
// blockIdx.y = 0 / 1: the box that comes from the lower / upper neighbour.  mode 0: dst = src ; 1: dst += src
__global__ void __launch_bounds__(256) k_box_pull(Arr3 A, PullBox lo, PullBox hi, int mx, int my, int mode, uint32_t seq, uint32_t *sig)
{
    const PullBox &b = blockIdx.y ? hi : lo;
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if ((size_t)blockIdx.x * blockDim.x >= b.total) return;
    if (threadIdx.x == 0) sig_wait(b.nb_sig + TGPU_SIG_READY, seq, sig + TGPU_SIG_TIMEOUT);
    __syncthreads();
    if (idx >= b.total) return;
    int c, i, j, k; box_decode(b.src, idx, c, i, j, k);
    const size_t ls = (size_t)(b.src.lo[0] + i - 1) + (size_t)b.rmx * ((size_t)(b.src.lo[1] + j - 1) + (size_t)b.rmy * (size_t)(b.src.lo[2] + k - 1));
    const size_t ld = LIDX(b.dst.lo[0] + i, b.dst.lo[1] + j, b.dst.lo[2] + k);
    const float v = __ldcv(b.rem[c] + ls);              // never a stale line of an earlier step
    if (mode) A.a[c][ld] = A.a[c][ld] + v; else A.a[c][ld] = v;
}

static const std::vector<int> &axis_extents(const tgpu_ctx *h, int axis) { return axis == 0 ? h->mxl : axis == 1 ? h->myl : h->mzl; }

// One halo step on `axis` for the three arrays first..first+2.  The data that travels UP is the box of n_up layers that
// starts at layer (up_from_m ? m_sender + up_off : up_off) of the sender and lands at layer dst_up of the receiver; same
// for DOWN.  recv_*: whether this rank takes what arrives from below / above (open boundaries, edge ranks).
static int halo_step(tgpu_ctx *h, int first, int axis, int n_up, bool up_from_m, int up_off, int dst_up, int recv_from_below,
                     int n_dn, bool dn_from_m, int dn_off, int dst_dn, int recv_from_above, int mode)
{
    const tgpu_params &P = h->P;
    const int below = topo_neighbour(P.rank, P.sizex, P.sizey, P.sizez, 2 * axis);
    const int above = topo_neighbour(P.rank, P.sizex, P.sizey, P.sizez, 2 * axis + 1);
    Arr3 A; for (int c = 0; c < 3; c++) A.a[c] = h->f[first + c];
    const uint32_t seq = ++h->xseq;
    k_sig_set<<<1, 1, 0, h->stream>>>(h->sig + TGPU_SIG_READY, seq); CKK(h);
    PullBox pb[2];
    for (int side = 0; side < 2; side++) {
        PullBox &b = pb[side];
        const int nb = side ? above : below;
        const PeerRank &R = h->peer[nb];
        b.src = full_box(h); b.dst = full_box(h);
        const int m_nb = axis_extents(h, axis)[nb];
        // from below comes what the lower neighbour sends UP; from above what the upper neighbour sends DOWN
        const int n = side ? n_dn : n_up;
        b.src.lo[axis] = side ? (dn_from_m ? m_nb + dn_off : dn_off) : (up_from_m ? m_nb + up_off : up_off);
        b.dst.lo[axis] = side ? dst_dn : dst_up;
        b.src.n[axis] = b.dst.n[axis] = n;
        for (int c = 0; c < 3; c++) b.rem[c] = R.f[first + c];
        b.rmx = h->mxl[nb]; b.rmy = h->myl[nb];
        b.nb_sig = R.sig;
        b.total = (side ? recv_from_above : recv_from_below) ? (unsigned long long)box_total(b.src) : 0ull;
    }
    const unsigned long long tmax = pb[0].total > pb[1].total ? pb[0].total : pb[1].total;
    if (tmax) {
        dim3 grid((unsigned)cdiv((long long)tmax, 256), 2);
        k_box_pull<<<grid, 256, 0, h->stream>>>(A, pb[0], pb[1], P.mx, P.my, mode, seq, h->sig); CKK(h);
    }
    k_sig_sync<<<1, 1, 0, h->stream>>>(h->sig, seq, h->peer[below].sig, above != below ? h->peer[above].sig : nullptr); CKK(h);
    return 0;
}

// bc_b1 / bc_e1: for iter = 1..g: low ghost g+1-iter <- (-nbr) m-(g+iter); high ghost m-g-1+iter <- (+nbr) g+iter.
// All g layers of one side move as one box: [1..g] <- [m-2g..m-g-1], [m-g..m-1] <- [g+1..2g].  Axis order x,y,z with
// full extents in the other axes so that corners propagate (fieldboundaries.F90:186-262).
int fld_bc(tgpu_ctx *h, int first)
{
    Arr3 A; for (int c = 0; c < 3; c++) A.a[c] = h->f[first + c];
    int naxes = h->P.dim == 3 ? 3 : 2;
    for (int axis = 0; axis < naxes; axis++) {
        int m, g, per, sz, pos; axis_info(h, axis, &m, &g, &per, &sz, &pos);
        if (!per && sz == 1 && axis != 2) continue;           // bc_b1: no copy at all
        if (!per && sz == 1 && axis == 2) continue;           // copy_layrz2 on a single rank: both receives skipped
        Box src = full_box(h), dst = full_box(h);
        if (sz > 1 && h->peer) {
            // over peer memory: the low ghosts [1, g] are the lower neighbour's layers [m' - 2g, m' - g - 1] (m' = its extent),
            // the high ghosts [m - g, m - 1] the upper neighbour's [g + 1, 2g]; layer by layer on a narrow axis
            const int r_lo = per || pos != 0, r_hi = per || pos != sz - 1;
            if (m - 2 * g - 1 < g) {
                for (int iter = 1; iter <= g; iter++) {
                    int rc = halo_step(h, first, axis, 1, true, -(g + iter), g + 1 - iter, r_lo, 1, false, g + iter, m - g - 1 + iter, r_hi, 0);
                    if (rc) return rc;
                }
            } else {
                int rc = halo_step(h, first, axis, g, true, -2 * g, 1, r_lo, g, false, g + 1, m - g, r_hi, 0);
                if (rc) return rc;
            }
            continue;
        }
        if (m - 2 * g - 1 < g) {
            // fewer interior cells than ghost layers (user/input.twostream: my0 = 2): the outer ghost layers are images of
            // layers that are ghosts themselves, so the reference's layer-by-layer order matters (do iter = 1, nghost/2 ...)
            src.n[axis] = dst.n[axis] = 1;
            for (int iter = 1; iter <= g; iter++) {
                src.lo[axis] = m - (g + iter); dst.lo[axis] = g + 1 - iter;
                int rc = box_shift(h, A, src, dst, axis, +1, 0, per || pos != 0); if (rc) return rc;
                src.lo[axis] = g + iter; dst.lo[axis] = m - g - 1 + iter;
                rc = box_shift(h, A, src, dst, axis, -1, 0, per || pos != sz - 1); if (rc) return rc;
            }
            continue;
        }
        // send up: my layers [m-2g, m-g-1] become the + neighbour's low ghosts [1, g]
        src.lo[axis] = m - 2 * g; src.n[axis] = g; dst.lo[axis] = 1; dst.n[axis] = g;
        int rc = box_shift(h, A, src, dst, axis, +1, 0, per || pos != 0);
        if (rc) return rc;
        // send down: my layers [g+1, 2g] become the - neighbour's high ghosts [m-g, m-1]
        src.lo[axis] = g + 1; dst.lo[axis] = m - g;
        rc = box_shift(h, A, src, dst, axis, -1, 0, per || pos != sz - 1);
        if (rc) return rc;
    }
    if (first < 6) h->need_prim = 1;
    return 0;
}
// NOTE on uneven splits: the destination index m-g is evaluated with the receiving rank's m; ranks on one axis
// line share m except the last one, and a box is always unpacked with the receiver's own geometry above.

// exchange_current: high ghosts [m-g..m] are added to the + neighbour's [g+1..nghost]; low ghosts [1..g] to the
// - neighbour's [m-nghost+1..m-g-1]; x, then y, then z (fieldboundaries.F90:1796-1813, 1990-2079, 2108-2185).
int fld_fold(tgpu_ctx *h)
{
    Arr3 A; for (int c = 0; c < 3; c++) A.a[c] = h->f[6 + c];
    int naxes = h->P.dim == 3 ? 3 : 2;
    for (int axis = 0; axis < naxes; axis++) {
        int m, g, per, sz, pos; axis_info(h, axis, &m, &g, &per, &sz, &pos);
        int ng = axis == 2 ? h->P.nghostz : h->P.nghost;
        if (sz == 1 && !per) continue;
        Box src = full_box(h), dst = full_box(h);
        if (sz > 1 && h->peer && m - ng >= g + 1) {
            // the lower neighbour's high ghosts [m' - g, m'] are added to my [g + 1, nghost]; the upper neighbour's low ghosts
            // [1, g] to my [m - nghost + 1, m - g - 1].  Sources (ghosts) and targets (interior) are disjoint: one step.
            int rc = halo_step(h, 6, axis, g + 1, true, -g, g + 1, per || pos != 0, g, false, 1, m - (ng - 1), per || pos != sz - 1, 1);
            if (rc) return rc;
            continue;
        }
        const bool overlap = m - ng < g + 1;                  // narrow axis: the target layers reach into the source ghosts
        src.lo[axis] = m - g; src.n[axis] = g + 1; dst.lo[axis] = g + 1; dst.n[axis] = g + 1;
        int rc = box_shift(h, A, src, dst, axis, +1, 1, per || pos != 0, overlap);
        if (rc) return rc;
        src.lo[axis] = 1; src.n[axis] = g; dst.lo[axis] = m - (ng - 1); dst.n[axis] = g;
        rc = box_shift(h, A, src, dst, axis, -1, 1, per || pos != sz - 1, overlap);
        if (rc) return rc;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// filter1: one 9/27-point pass for all three components (filter.F90:102-216), ghosts refreshed one
// layer per pass (filter.F90:71-99).
// ---------------------------------------------------------------------------------------------
template <int DIM>
__global__ void __launch_bounds__(128) k_filter1(Arr3 S, Arr3 D, int mx, int my, int mz, int g, int gz, int q5)
{
    const float winv = DIM == 3 ? 1.f / 64.f : 1.f / 16.f;
    const float w1 = (DIM == 3 ? 4.f : 2.f) * winv, w0 = (DIM == 3 ? 8.f : 4.f) * winv, w2 = (DIM == 3 ? 2.f : 1.f) * winv;
    const float wz1 = 2.f * winv, wz0 = 4.f * winv, wz2 = 1.f * winv;
    int i = g + 1 + blockIdx.x * blockDim.x + threadIdx.x;
    int j = g + 1 + blockIdx.y;
    int nk = DIM == 3 ? mz - 2 * gz - 1 : 1;
    int c = blockIdx.z / nk;
    int k = DIM == 3 ? gz + 1 + blockIdx.z % nk : 1;
    if (i > mx - (g + 1)) return;
    int jmax = (c == 2 && q5) ? my - g + 1 : my - (g + 1);
    if (j > jmax) return;
    const float *cu = S.a[c];
#define C(di, dj, dk) cu[LIDX(i + (di), j + (dj), k + (dk))]
    float t = w1 * C(-1, 0, 0) + w0 * C(0, 0, 0) + w1 * C(1, 0, 0) + w1 * C(0, -1, 0) + w1 * C(0, 1, 0) +
              w2 * C(-1, 1, 0) + w2 * C(1, 1, 0) + w2 * C(-1, -1, 0) + w2 * C(1, -1, 0);
    if (DIM == 3) {
        t = t + wz1 * C(-1, 0, -1) + wz0 * C(0, 0, -1) + wz1 * C(1, 0, -1) + wz1 * C(0, -1, -1) + wz1 * C(0, 1, -1) +
            wz2 * C(-1, 1, -1) + wz2 * C(1, 1, -1) + wz2 * C(-1, -1, -1) + wz2 * C(1, -1, -1) +
            wz1 * C(-1, 0, 1) + wz0 * C(0, 0, 1) + wz1 * C(1, 0, 1) + wz1 * C(0, -1, 1) + wz1 * C(0, 1, 1) +
            wz2 * C(-1, 1, 1) + wz2 * C(1, 1, 1) + wz2 * C(-1, -1, 1) + wz2 * C(1, -1, 1);
    }
#undef C
    D.a[c][LIDX(i, j, k)] = t;
}

// copy the filtered interior back (filter.F90:121-131)
__global__ void __launch_bounds__(128) k_filter1_back(Arr3 S, Arr3 D, int mx, int my, int mz, int g, int gz, int dim, int q5)
{
    int i = g + 1 + blockIdx.x * blockDim.x + threadIdx.x;
    int j = g + 1 + blockIdx.y;
    int nk = dim == 3 ? mz - 2 * gz - 1 : 1;
    int c = blockIdx.z / nk;
    int k = dim == 3 ? gz + 1 + blockIdx.z % nk : 1;
    if (i > mx - (g + 1)) return;
    int jmax = (c == 2 && q5) ? my - g + 1 : my - (g + 1);
    if (j > jmax) return;
    D.a[c][LIDX(i, j, k)] = S.a[c][LIDX(i, j, k)];
}

static int filter1_passes(tgpu_ctx *h);

// On one rank the ntimes passes are a fixed sequence of small launches with fixed arguments (6 per pass: 192 for the
// shipped ntimes = 32, which is what a 128 x 128 problem's lap consists of): captured once into a CUDA graph and replayed.
int fld_filter1(tgpu_ctx *h)
{
    if (h->size0 > 1 || h->P.ntimes <= 0 || !h->opt_graph) return filter1_passes(h);
    if (!h->f1_graph) {
        const int64_t l0 = h->launches;
        CK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
        const int rc = filter1_passes(h);
        cudaGraph_t gr = nullptr;
        const cudaError_t e = cudaStreamEndCapture(h->stream, &gr);
        if (rc || e != cudaSuccess || !gr) { if (gr) cudaGraphDestroy(gr); cudaGetLastError(); h->opt_graph = 0; h->launches = l0; return filter1_passes(h); }
        cudaGraphExec_t ex = nullptr;
        if (cudaGraphInstantiate(&ex, gr, 0) != cudaSuccess) { cudaGraphDestroy(gr); cudaGetLastError(); h->opt_graph = 0; h->launches = l0; return filter1_passes(h); }
        cudaGraphDestroy(gr);
        h->f1_graph = ex; h->f1_graph_launches = (int)(h->launches - l0); h->launches = l0;
    }
    CK(cudaGraphLaunch((cudaGraphExec_t)h->f1_graph, h->stream));
    h->launches += h->f1_graph_launches;          // kernels executed (one graph launch)
    return 0;
}

static int filter1_passes(tgpu_ctx *h)
{
    const tgpu_params &P = h->P;
    Arr3 A, T; for (int c = 0; c < 3; c++) { A.a[c] = h->f[6 + c]; T.a[c] = h->ftmp[c]; }
    int g = P.nghost / 2, gz = P.nghostz / 2;
    int q5 = (P.quirks & TGPU_Q5_FILTER_CURZ_J) != 0;
    int naxes = P.dim == 3 ? 3 : 2;
    int nk = P.dim == 3 ? P.mz - 2 * gz - 1 : 1;
    int nj = P.my - 2 * g - 1 + (q5 ? 2 : 0);
    dim3 grid(cdiv(P.mx - 2 * g - 1, 128), nj, 3 * nk);
    for (int n = 1; n <= P.ntimes; n++) {
        for (int axis = 0; axis < naxes; axis++) {
            int m, ga, per, sz, pos; axis_info(h, axis, &m, &ga, &per, &sz, &pos);
            Box src = full_box(h), dst = full_box(h);
            src.n[axis] = dst.n[axis] = 1;
            // (lt,ls,nt,ns) = (g, m-g-1, m-g, g+1); copy_layr*2_opt on open axes skips the outer receive
            if (sz > 1 && h->peer) {
                int rc = halo_step(h, 6, axis, 1, true, -ga - 1, ga, per || pos != 0, 1, false, ga + 1, m - ga, per || pos != sz - 1, 0);
                if (rc) return rc;
                continue;
            }
            src.lo[axis] = m - ga - 1; dst.lo[axis] = ga;
            int rc = box_shift(h, A, src, dst, axis, +1, 0, per || pos != 0);
            if (rc) return rc;
            src.lo[axis] = ga + 1; dst.lo[axis] = m - ga;
            rc = box_shift(h, A, src, dst, axis, -1, 0, per || pos != sz - 1);
            if (rc) return rc;
        }
        if (P.dim == 3) k_filter1<3><<<grid, 128, 0, h->stream>>>(A, T, P.mx, P.my, P.mz, g, gz, q5);
        else k_filter1<2><<<grid, 128, 0, h->stream>>>(A, T, P.mx, P.my, P.mz, g, gz, q5);
        CKK(h);
        k_filter1_back<<<grid, 128, 0, h->stream>>>(T, A, P.mx, P.my, P.mz, g, gz, P.dim, q5);
        CKK(h);
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// filter2: ntimes fixed-end 1-2-1 passes per grid line, along x, then y, then z, for curx, cury, curz
// (optimized_filters.F90:459-558 and the y/z twins), with ntimes-deep halos fetched once per axis
// (deep_copy_layr*, :1387-1963).  A CTA stages NL whole extended lines in shared memory and runs
// all passes there (ping-pong); each global element is read and written exactly once per axis.
//   line element p in [0, nt)            : low halo  (neighbour cells fin-nt+1+p, or replicated edge)
//                 p in [nt, nt+ncell)     : cur(str + p - nt)
//                 p in [nt+ncell, L)      : high halo
// ---------------------------------------------------------------------------------------------
struct F2 {
    int mx, my, mz;
    int axis, str, ncell, nt;
    int q_lo[2], q_n[2];       // ranges of the two other axes (a1 < a2)
    int lowmode, highmode;     // 0 = wrap inside this array, 1 = from halo buffer, 2 = replicate edge
};

__device__ __forceinline__ size_t f2_addr(const F2 &f, int p_cell, int q1, int q2)
{
    int ijk[3];
    int a1 = f.axis == 0 ? 1 : 0, a2 = f.axis == 2 ? 1 : 2;
    ijk[f.axis] = p_cell; ijk[a1] = f.q_lo[0] + q1; ijk[a2] = f.q_lo[1] + q2;
    return (size_t)(ijk[0] - 1) + (size_t)f.mx * ((size_t)(ijk[1] - 1) + (size_t)f.my * (size_t)(ijk[2] - 1));
}

// Shared memory holds the tile only for the coalesced load/store.  The passes run in registers with no block-level
// synchronisation: one warp owns one whole extended line, lane l keeps R consecutive elements (R odd, so the strided
// shared-memory reads are conflict-free) and trades one edge value per side per pass with its lane neighbours through
// warp shuffles.  Every global element is read and written once per axis; the arithmetic per pass is exactly
// new(p) = .25*old(p-1) + .5*old(p) + .25*old(p+1) with both end points of the extended line held fixed.
// tile layout: [line][Lp], Lp = 32*R + 1.
// EXACT: the extended line fills the warp exactly (L == 32*R), so the two fixed end points sit at compile-time positions
// (lane 0, r = 0) and (lane 31, r = R-1) and the per-element "is this an end point" predicate of the general form disappears
// from the pass loop: 4 instead of 5 instructions per element-pass, same arithmetic.  (R may then be even: its 2-way bank
// conflicts only touch the two register load/store phases.)
template <int R, bool EXACT = false>
__global__ void __launch_bounds__(512) k_filter2(float *__restrict__ cur, const float *__restrict__ halo_lo,
                                                 const float *__restrict__ halo_hi, F2 f, int NL)
{
    extern __shared__ float tile[];
    const int L = f.ncell + 2 * f.nt;
    constexpr int Lp = 32 * R + 1;
    const int nlines = f.q_n[0] * f.q_n[1];
    const int line0 = blockIdx.x * NL;
    const int nthr = blockDim.x;
    // Load/store index arithmetic without per-element divisions: a line's (q1, q2) and base address are computed once.
    //   axis 0: the line is contiguous in memory -> one warp per line, lanes along the line;
    //   axis 1, 2: consecutive lines are consecutive in x -> a thread keeps one line (NL is a power of two) and walks p.
    const int lane_ = threadIdx.x & 31, warp_ = threadIdx.x >> 5, nwarps_ = blockDim.x >> 5;
    int ln_first, ln_step, p_first, p_step;
    if (f.axis == 0) { ln_first = warp_; ln_step = nwarps_; p_first = lane_; p_step = 32; }
    else { ln_first = threadIdx.x & (NL - 1); ln_step = NL; p_first = threadIdx.x / NL; p_step = nthr / NL; }
    const size_t pstride = f.axis == 0 ? 1 : f.axis == 1 ? (size_t)f.mx : (size_t)f.mx * f.my;
    for (int ln = ln_first; ln < NL; ln += ln_step) {
        const int line = line0 + ln;
        float *trow = tile + ln * Lp;
        if (line >= nlines) { for (int p = p_first; p < L; p += p_step) trow[p] = 0.f; continue; }
        const int q1 = line % f.q_n[0], q2 = line / f.q_n[0];
        const float *cbase = cur + f2_addr(f, f.str, q1, q2);            // cell `str` of this line
        const float *hl_ = halo_lo + (size_t)f.nt * line, *hh_ = halo_hi + (size_t)f.nt * line;
        for (int p = p_first; p < L; p += p_step) {
            const int pc = p - f.nt;                                     // cell offset relative to str
            float v;
            if (pc < 0) v = f.lowmode == 0 ? cbase[(size_t)(f.ncell + pc) * pstride] : f.lowmode == 1 ? hl_[p] : cbase[0];
            else if (pc >= f.ncell)
                v = f.highmode == 0 ? cbase[(size_t)(pc - f.ncell) * pstride] : f.highmode == 1 ? hh_[pc - f.ncell]
                                                                                             : cbase[(size_t)(f.ncell - 1) * pstride];
            else v = cbase[(size_t)pc * pstride];
            trow[p] = v;
        }
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int p0 = lane * R;
    for (int ln = warp; ln < NL; ln += nwarps) {
        float *row = tile + ln * Lp + p0;
        float v[R];
#pragma unroll
        for (int r = 0; r < R; r++) v[r] = (p0 + r < L) ? row[r] : 0.f;
        if (EXACT) {
            const float f0 = v[0], fl = v[R - 1];
            const bool first = lane == 0, last = lane == 31;
            for (int n = 0; n < f.nt; n++) {
                const float left = __shfl_up_sync(0xffffffffu, v[R - 1], 1);
                const float right = __shfl_down_sync(0xffffffffu, v[0], 1);
                // (.25 * prev + .5 * c) + .25 * nx, the reference's order (optimized_filters.F90:487-516).  Scaling by 1/4 and
                // 1/2 is exact, so the two explicit FMAs round exactly where the separate multiplies and adds would, and
                // .25 * c is computed once and reused as the next element's first term: 3 instructions per element-pass
                // instead of 4 (the kernel is issue-bound; this file is built with -fmad=false, fmaf() stays an FMA)
                float tprev = .25f * left;
#pragma unroll
                for (int r = 0; r < R; r++) {
                    const float c = v[r];
                    const float nx = r < R - 1 ? v[r + 1] : right;
                    v[r] = fmaf(.25f, nx, fmaf(.5f, c, tprev));
                    tprev = .25f * c;
                }
                v[0] = first ? f0 : v[0];
                v[R - 1] = last ? fl : v[R - 1];
            }
        } else
        for (int n = 0; n < f.nt; n++) {
            const float left = __shfl_up_sync(0xffffffffu, v[R - 1], 1);
            const float right = __shfl_down_sync(0xffffffffu, v[0], 1);
            float tprev = .25f * left;
#pragma unroll
            for (int r = 0; r < R; r++) {
                const int p = p0 + r;
                const float c = v[r];
                const float nx = r < R - 1 ? v[r + 1] : right;
                const float nv = fmaf(.25f, nx, fmaf(.5f, c, tprev));
                v[r] = (p == 0 || p >= L - 1) ? c : nv;
                tprev = .25f * c;
            }
        }
#pragma unroll
        for (int r = 0; r < R; r++) if (p0 + r < L) row[r] = v[r];
    }
    __syncthreads();
    for (int ln = ln_first; ln < NL; ln += ln_step) {
        const int line = line0 + ln;
        if (line >= nlines) continue;
        const int q1 = line % f.q_n[0], q2 = line / f.q_n[0];
        float *cbase = cur + f2_addr(f, f.str, q1, q2);
        const float *trow = tile + ln * Lp + f.nt;
        for (int p = p_first; p < f.ncell; p += p_step) cbase[(size_t)p * pstride] = trow[p];
    }
}

// pack the nt interior layers next to a face, restricted to the interior of the other axes, in the
// [line][p] order k_filter2 reads its halo in
__global__ void __launch_bounds__(256) k_f2_pack(const float *__restrict__ cur, float *__restrict__ buf, F2 f, int first_cell)
{
    size_t total = (size_t)f.nt * f.q_n[0] * f.q_n[1];
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    int p = (int)(idx % f.nt); int line = (int)(idx / f.nt);
    int q1 = line % f.q_n[0], q2 = line / f.q_n[0];
    buf[idx] = cur[f2_addr(f, first_cell + p, q1, q2)];
}

// filter2 deep halo over peer memory: my low halo is the last nt interior layers of the lower neighbour, my high halo the
// first nt of the upper one (optimized_filters.F90:1665-1674, 1838-1847), read straight out of their cur arrays in the
// [line][p] order k_filter2 wants.  blockIdx.y = 2 * component + side.
struct F2Pull {
    const float *rem[2][3]; float *buf[2][3];
    int first[2], rmx[2], rmy[2], on[2];
    const uint32_t *nb_sig[2];
};
__global__ void __launch_bounds__(256) k_f2_pull(F2 f, F2Pull a, uint32_t seq, uint32_t *sig)
{
    const int side = blockIdx.y & 1, c = blockIdx.y >> 1;
    if (!a.on[side]) return;
    if (threadIdx.x == 0) sig_wait(a.nb_sig[side] + TGPU_SIG_READY, seq, sig + TGPU_SIG_TIMEOUT);
    __syncthreads();
    const size_t total = (size_t)f.nt * f.q_n[0] * f.q_n[1];
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    // thread order follows the REMOTE array (x fastest) so that the NVLink reads are whole lines; the scattered side is the
    // local store into the [line][p] order k_filter2 reads (a strided remote read was measured 2.5x slower than NCCL)
    int p, q1, q2;
    if (f.axis == 0) { p = (int)(idx % f.nt); const int line = (int)(idx / f.nt); q1 = line % f.q_n[0]; q2 = line / f.q_n[0]; }
    else { q1 = (int)(idx % f.q_n[0]); const size_t r = idx / f.q_n[0]; p = (int)(r % f.nt); q2 = (int)(r / f.nt); }
    F2 fr = f; fr.mx = a.rmx[side]; fr.my = a.rmy[side];
    a.buf[side][c][(size_t)p + (size_t)f.nt * (q1 + (size_t)f.q_n[0] * q2)] = __ldcv(a.rem[side][c] + f2_addr(fr, a.first[side] + p, q1, q2));
}

int fld_filter2(tgpu_ctx *h)
{
    const tgpu_params &P = h->P;
    if (P.ntimes <= 0) return 0;
    int naxes = P.dim == 3 ? 3 : 2;
    int lo[3], n[3];
    for (int a = 0; a < 3; a++) {
        int m, g, per, sz, pos; axis_info(h, a, &m, &g, &per, &sz, &pos);
        lo[a] = g + 1; n[a] = m - 2 * g - 1;
    }
    if (P.dim == 2) { lo[2] = 1; n[2] = 1; }
    // The three components are independent, so the loops run axis-major: the ntimes-deep halos of all three components of an
    // axis travel in ONE NCCL group (2 groups per lap on a y/z-decomposed box instead of 6), then the three kernels run.
    for (int axis = 0; axis < naxes; axis++) {
        int m, g, per, sz, pos; axis_info(h, axis, &m, &g, &per, &sz, &pos);
        F2 f; f.mx = P.mx; f.my = P.my; f.mz = P.mz; f.axis = axis; f.str = lo[axis]; f.ncell = n[axis]; f.nt = P.ntimes;
        int a1 = axis == 0 ? 1 : 0, a2 = axis == 2 ? 1 : 2;
        f.q_lo[0] = lo[a1]; f.q_n[0] = n[a1]; f.q_lo[1] = lo[a2]; f.q_n[1] = n[a2];
        if (f.nt > f.ncell) { tgpu_set_error("filter2: ntimes exceeds the local extent"); return TGPU_EINVAL; }
        f.lowmode = per ? 0 : 2; f.highmode = per ? 0 : 2;
        float *hlo_c[3] = {nullptr, nullptr, nullptr}, *hhi_c[3] = {nullptr, nullptr, nullptr};
        if (sz > 1) {
            size_t cnt = (size_t)f.nt * f.q_n[0] * f.q_n[1];
            if (12 * cnt > h->halo_floats) { tgpu_set_error("halo scratch too small for filter2"); return TGPU_EINVAL; }
            int up = topo_neighbour(P.rank, P.sizex, P.sizey, P.sizez, 2 * axis + 1);
            int dn = topo_neighbour(P.rank, P.sizex, P.sizey, P.sizez, 2 * axis);
            if (h->peer) {
                F2Pull a;
                for (int side = 0; side < 2; side++) {
                    const int nb = side ? up : dn;
                    const int m_nb = axis_extents(h, axis)[nb];
                    for (int c = 0; c < 3; c++) {
                        a.rem[side][c] = h->peer[nb].f[6 + c];
                        a.buf[side][c] = h->halo + (4 * c + 2 + side) * cnt;
                    }
                    // lower neighbour: its last nt interior layers; upper neighbour: its first nt
                    a.first[side] = side ? f.str : f.str + (m_nb - 2 * g - 1) - f.nt;
                    a.rmx[side] = h->mxl[nb]; a.rmy[side] = h->myl[nb];
                    a.on[side] = side ? (per || pos != sz - 1) : (per || pos != 0);
                    a.nb_sig[side] = h->peer[nb].sig;
                }
                for (int c = 0; c < 3; c++) { hlo_c[c] = a.buf[0][c]; hhi_c[c] = a.buf[1][c]; }
                const uint32_t seq = ++h->xseq;
                k_sig_set<<<1, 1, 0, h->stream>>>(h->sig + TGPU_SIG_READY, seq); CKK(h);
                k_f2_pull<<<dim3(cdiv(cnt, 256), 6), 256, 0, h->stream>>>(f, a, seq, h->sig); CKK(h);
                // the neighbours read my cur before I filter it in place
                k_sig_sync<<<1, 1, 0, h->stream>>>(h->sig, seq, h->peer[dn].sig, up != dn ? h->peer[up].sig : nullptr); CKK(h);
            } else {
            for (int c = 0; c < 3; c++) {
                float *s_up = h->halo + (4 * c) * cnt, *s_dn = h->halo + (4 * c + 1) * cnt;
                hlo_c[c] = h->halo + (4 * c + 2) * cnt; hhi_c[c] = h->halo + (4 * c + 3) * cnt;
                // my last nt cells go up and become the + neighbour's low halo; my first nt cells go down
                k_f2_pack<<<cdiv(cnt, 256), 256, 0, h->stream>>>(h->f[6 + c], s_up, f, f.str + f.ncell - f.nt); CKK(h);
                k_f2_pack<<<cdiv(cnt, 256), 256, 0, h->stream>>>(h->f[6 + c], s_dn, f, f.str); CKK(h);
            }
            int rc = comm_group_begin(h); if (rc) return rc;
            for (int c = 0; c < 3; c++) {
                float *s_up = h->halo + (4 * c) * cnt, *s_dn = h->halo + (4 * c + 1) * cnt;
                comm_send(h, s_up, cnt * 4, up); comm_recv(h, hlo_c[c], cnt * 4, dn);
                comm_send(h, s_dn, cnt * 4, dn); comm_recv(h, hhi_c[c], cnt * 4, up);
            }
            rc = comm_group_end(h); if (rc) return rc;
            }
            f.lowmode = (per || pos != 0) ? 1 : 2;
            f.highmode = (per || pos != sz - 1) ? 1 : 2;
        }
        for (int c = 0; c < 3; c++) {
            float *hlo = hlo_c[c], *hhi = hhi_c[c];
            const int L = f.ncell + 2 * f.nt;
            int nlines = f.q_n[0] * f.q_n[1];
            // strip length R (odd): smallest instantiated value with 32*R >= L; or the exact-fit kernel when L == 32*R
            int R = L <= 224 ? 7 : L <= 352 ? 11 : L <= 608 ? 19 : L <= 1120 ? 35 : 0;
            if (!R) { tgpu_set_error("filter2: extended line longer than 1120 elements is not instantiated"); return TGPU_EINVAL; }
            const bool exact = L % 32 == 0 && L / 32 >= 4 && L / 32 <= 20;
            if (exact) R = L / 32;
            const int Lp = 32 * R + 1;
            int NL = 32;
            while (NL > 1 && (size_t)Lp * NL * 4 > 100 * 1024) NL >>= 1;
            size_t smem = (size_t)Lp * NL * 4;
            int threads = NL >= 16 ? 512 : NL * 32;
#define LAUNCH_F2(RV)                                                                                              \
    {                                                                                                              \
        CK(cudaFuncSetAttribute(k_filter2<RV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));           \
        k_filter2<RV><<<cdiv(nlines, NL), threads, smem, h->stream>>>(h->f[6 + c], hlo, hhi, f, NL);               \
    }
#define LAUNCH_F2X(RV)                                                                                             \
    case RV: {                                                                                                     \
        CK(cudaFuncSetAttribute(k_filter2<RV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
        k_filter2<RV, true><<<cdiv(nlines, NL), threads, smem, h->stream>>>(h->f[6 + c], hlo, hhi, f, NL);         \
    } break;
            if (exact) {
                switch (R) {
                    LAUNCH_F2X(4) LAUNCH_F2X(5) LAUNCH_F2X(6) LAUNCH_F2X(7) LAUNCH_F2X(8) LAUNCH_F2X(9) LAUNCH_F2X(10)
                    LAUNCH_F2X(11) LAUNCH_F2X(12) LAUNCH_F2X(13) LAUNCH_F2X(14) LAUNCH_F2X(15) LAUNCH_F2X(16) LAUNCH_F2X(17)
                    LAUNCH_F2X(18) LAUNCH_F2X(19) LAUNCH_F2X(20)
                }
            } else
            switch (R) {
            case 7: LAUNCH_F2(7) break;
            case 11: LAUNCH_F2(11) break;
            case 19: LAUNCH_F2(19) break;
            default: LAUNCH_F2(35) break;
            }
#undef LAUNCH_F2X
#undef LAUNCH_F2
            CKK(h);
        }
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Radiation boundary `surface` (fieldboundaries.F90:493-606): the Lindman-type absorbing face that bc_b2 (:274-295) applies
// to the high face of every radiating axis and bc_e2 (:403-426) to the low face, with mirrored strides and E <-> B.
// Arguments keep the reference's meaning (rotated components b1,b2,b3 / e1,e2,e3, strides s1,s2,s3 with s3 normal to the
// face, 1-based first element m00).  The reference sweeps rows then columns; its data flow is
//   (1) b3 += h over the face, (2) b1 and b2 from that b3, (3) b3 += h again,   h = c/2 * curl_3(e),
// so two launches suffice: k_surface_b12 recomputes b3 + h locally (for the point and its two lower neighbours, same
// expression -> same bits) and writes only b1, b2; k_surface_b3 then applies the two half updates.  The twoD variants of the
// reference (:540-583, one stride zero) are the same formulas with a one-point range on the degenerate axis.
// ---------------------------------------------------------------------------------------------
struct SurfArgs {
    float *b1, *b2, *b3; const float *e1, *e2, *e3;
    long long s1, s2, s3, mf;       // strides; mf = 1-based index of the first face element
    int n1, n2;                     // points along s1, s2 that receive the b3 update
    int b1_first, b2_first;         // first index along s1 (s2) that receives the b1 (b2) update
    float c, rs, s, os;
};
#define SF(a, n) (a)[(n) - 1]
__device__ __forceinline__ float surf_h(const SurfArgs &A, long long n)
{
    return .5f * A.c * (SF(A.e1, n + A.s2) - SF(A.e1, n) - SF(A.e2, n + A.s1) + SF(A.e2, n));
}
__global__ void __launch_bounds__(256) k_surface_b12(SurfArgs A)
{
    const int ii = blockIdx.x * blockDim.x + threadIdx.x, jj = blockIdx.y;
    if (ii >= A.n1) return;
    const long long n = A.mf + A.s1 * ii + A.s2 * jj;
    const float bz_n = SF(A.b3, n) + surf_h(A, n);
    if (ii >= A.b1_first) {
        const float bz_m = SF(A.b3, n - A.s1) + surf_h(A, n - A.s1);
        SF(A.b1, n) = SF(A.b1, n) + A.rs * (SF(A.b1, n - A.s3) - SF(A.b1, n) + A.s * (bz_n - bz_m))
                      - A.os * (SF(A.e3, n + A.s2) - SF(A.e3, n)) - (A.os - A.c) * (SF(A.e3, n + A.s2 - A.s3) - SF(A.e3, n - A.s3))
                      - A.c * (SF(A.e2, n) - SF(A.e2, n - A.s3));
    }
    if (jj >= A.b2_first) {
        const float bz_m = SF(A.b3, n - A.s2) + surf_h(A, n - A.s2);
        SF(A.b2, n) = SF(A.b2, n) + A.rs * (SF(A.b2, n - A.s3) - SF(A.b2, n) + A.s * (bz_n - bz_m))
                      + A.os * (SF(A.e3, n + A.s1) - SF(A.e3, n)) + (A.os - A.c) * (SF(A.e3, n + A.s1 - A.s3) - SF(A.e3, n - A.s3))
                      + A.c * (SF(A.e1, n) - SF(A.e1, n - A.s3));
    }
}
__global__ void __launch_bounds__(256) k_surface_b3(SurfArgs A)
{
    const int ii = blockIdx.x * blockDim.x + threadIdx.x, jj = blockIdx.y;
    if (ii >= A.n1) return;
    const long long n = A.mf + A.s1 * ii + A.s2 * jj;
    const float hh = surf_h(A, n);
    SF(A.b3, n) = (SF(A.b3, n) + hh) + hh;
}
#undef SF
static int surface_face(tgpu_ctx *h, float *b1, float *b2, float *b3, const float *e1, const float *e2, const float *e3,
                        long long s1, long long s2, long long s3, int m1, int m2, int m3, long long m00)
{
    SurfArgs A;
    A.b1 = b1; A.b2 = b2; A.b3 = b3; A.e1 = e1; A.e2 = e2; A.e3 = e3; A.s1 = s1; A.s2 = s2; A.s3 = s3;
    A.mf = m00 + s3 * (m3 - 1);
    A.c = h->P.c; A.rs = 2.f * A.c / (1.f + A.c); A.s = .4142136f; A.os = .5f * (1.f - A.s) * A.rs;
    A.n1 = m1 - 1; A.n2 = m2 - 1; A.b1_first = 1; A.b2_first = 1;
    if (h->P.dim == 2) {
        if (s1 == 0) { A.n1 = 1; A.b1_first = 0; }            // fieldboundaries.F90:540-560
        else if (s2 == 0) { A.n2 = 1; A.b2_first = 0; }       // :565-583
        else return 0;
    }
    if (A.n1 <= 0 || A.n2 <= 0) return 0;
    dim3 grid((A.n1 + 255) / 256, A.n2);
    k_surface_b12<<<grid, 256, 0, h->stream>>>(A); CKK(h);
    k_surface_b3<<<grid, 256, 0, h->stream>>>(A); CKK(h);
    return 0;
}
// is_e = 0: the `surface` calls of bc_b2; 1: those of bc_e2
int fld_surface(tgpu_ctx *h, int is_e)
{
    const tgpu_params &P = h->P;
    // fieldboundaries.F90:88-94: "open y also opens z" exists only under #ifdef twoD, where the z call itself is compiled
    // out (:287-291, 418-422) -- so in 3D an open y with periodic z radiates on y only
    int rad[3] = {1 - P.periodicx, 1 - P.periodicy, 1 - P.periodicz};
    if (P.dim == 2) rad[2] = 0;
    if (!rad[0] && !rad[1] && !rad[2]) return 0;
    float *ex = h->f[0], *ey = h->f[1], *ez = h->f[2], *bx = h->f[3], *by = h->f[4], *bz = h->f[5];
    const long long ix = 1, iy = P.mx, iz = P.dim == 3 ? (long long)P.mx * P.my : 0, lot = h->G.lot;
    const int mx = P.mx, my = P.my, mz = P.dim == 3 ? P.mz : 1;
    int rc = 0;
    if (!is_e) {
        if (rad[0]) rc |= surface_face(h, by, bz, bx, ey, ez, ex, iy, iz, ix, my, mz, mx, 1);
        if (rad[1]) rc |= surface_face(h, bz, bx, by, ez, ex, ey, iz, ix, iy, mz, mx, my, 1);
        if (rad[2]) rc |= surface_face(h, bx, by, bz, ex, ey, ez, ix, iy, iz, mx, my, mz, 1);
    } else {
        if (rad[0]) rc |= surface_face(h, ey, ez, ex, by, bz, bx, -iy, -iz, -ix, my, mz, mx, lot);
        if (rad[1]) rc |= surface_face(h, ez, ex, ey, bz, bx, by, -iz, -ix, -iy, mz, mx, my, lot);
        if (rad[2]) rc |= surface_face(h, ex, ey, ez, bx, by, bz, -ix, -iy, -iz, mx, my, mz, lot);
    }
    h->need_prim = 1;
    return rc;
}

// ---------------------------------------------------------------------------------------------
// Edge fixes of an all-open 3D box: preledge / postedge (fieldboundaries.F90:2200-2244, 2371-2407), the #ifndef twoD
// bodies.  Each routine is four sweeps along one edge line of the rotated box; no sweep reads an element that the same
// sweep writes at another n, so a sweep is one launch with one thread per n, and the launches follow the reference's order.
// E1(a, n): the reference's 1-based flat index.  fields.cu is built with -fmad=false: same roundings as the oracle.
// ---------------------------------------------------------------------------------------------
struct EdgeArgs {
    float *bx, *by, *bz; const float *ex, *ey, *ez;
    long long ix, iy, iz, m; int mx, my, mz; float c;
};
#define E1(a, n) (a)[(n) - 1]
template <int POST, int SWEEP>
__global__ void __launch_bounds__(128) k_edge(EdgeArgs A)
{
    const long long ix = A.ix, iy = A.iy, iz = A.iz;
    const float c = A.c, s = .4142136f;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    float *bx = A.bx, *by = A.by, *bz = A.bz; const float *ex = A.ex, *ey = A.ey, *ez = A.ez;
    const float t = c / (2.f * c + 1.f + s), r = 4.f / (2.f * c + 2.f + s);
    const float p = (1.f + c) * 2.f / (1.f + 2.f * c * (1.f + c * s)), q_ = c * s * 2.f / (1.f + 2.f * c * (1.f + c * s));
    const long long corner = A.m + iy * (A.my - 1) + iz * (A.mz - 1);
    if (SWEEP == 0) {                  // bx along x on the (y, z) = (my, mz) edge, n = corner + ix .. corner + ix*(mx-2)
        if (q >= A.mx - 2) return;
        const long long n = corner + ix * (q + 1);
        const float cross = s * t * (E1(by, n) - E1(by, n - ix) + E1(by, n - iz) - E1(by, n - ix - iz)
                                     + E1(bz, n) - E1(bz, n - ix) + E1(bz, n - iy) - E1(bz, n - ix - iy));
        if (!POST) E1(bx, n) = E1(bx, n - iy - iz) + (1.f - 4.f * t) * E1(bx, n) + (1.f - 2.f * t) * (E1(bx, n - iy) + E1(bx, n - iz)) + cross;
        else E1(bx, n) = E1(bx, n) - (1.f - 4.f * t) * E1(bx, n - iy - iz) - (1.f - 2.f * t) * (E1(bx, n - iy) + E1(bx, n - iz)) + cross;
    } else if (SWEEP == 1) {           // bx along y on the z = mz face's first column
        if (q >= A.my - 1) return;
        const long long n = A.m + iz * (A.mz - 1) + iy * q;
        if (!POST) E1(bx, n) = (1.f - c * r) * (E1(bx, n) + E1(bx, n + ix)) + E1(bx, n - iz) + E1(bx, n + ix - iz) - r * (E1(bz, n)
                + c * ((1.f - s) * (E1(ex, n + iy) - E1(ex, n)) + (1.f + s) * .25f * (E1(ez, n + iy)
                - E1(ez, n) + E1(ez, n + ix + iy) - E1(ez, n + ix) + E1(ez, n + iy - iz) - E1(ez, n - iz)
                + E1(ez, n + ix + iy - iz) - E1(ez, n + ix - iz))));
        else E1(bx, n) = E1(bx, n) - E1(bx, n + ix) - (1.f - c * r) * (E1(bx, n - iz) + E1(bx, n + ix - iz)) + r * E1(bz, n);
    } else if (SWEEP == 2) {           // bx along z on the y = my face's first column
        if (q >= A.mz - 1) return;
        const long long n = A.m + iy * (A.my - 1) + iz * q;
        if (!POST) E1(bx, n) = (1.f - c * r) * (E1(bx, n) + E1(bx, n + ix)) + E1(bx, n - iy) + E1(bx, n + ix - iy)
                - r * (E1(by, n) - c * ((1.f - s) * (E1(ex, n + iz) - E1(ex, n))
                + (1.f + s) * .25f * (E1(ey, n + iz) - E1(ey, n) + E1(ey, n + ix + iz)
                - E1(ey, n + ix) + E1(ey, n + iz - iy) - E1(ey, n - iy) + E1(ey, n + ix + iz - iy)
                - E1(ey, n + ix - iy))));
        else E1(bx, n) = E1(bx, n) - E1(bx, n + ix) - (1.f - c * r) * (E1(bx, n - iy) + E1(bx, n + ix - iy)) + r * E1(by, n);
    } else {                           // by, bz along x on the (my, mz) edge, n = corner .. corner + ix*(mx-2)
        if (q >= A.mx - 1) return;
        const long long n = corner + ix * q;
        if (!POST) {
            const float temp = E1(bz, n) - .5f * c * (1.f - s) * (E1(ey, n + ix) - E1(ey, n) + E1(ey, n + ix - iy) - E1(ey, n - iy));
            const float byn = E1(by, n);
            E1(bz, n) = E1(bz, n - iy) - E1(bz, n) + p * temp + q_ * byn;
            E1(by, n) = E1(by, n - iz) - byn + p * byn + q_ * temp;
        } else {
            const float temp = E1(by, n - iz) - .5f * c * (1.f - s) * (E1(ez, n + ix) - E1(ez, n) + E1(ez, n + ix - iz) - E1(ez, n - iz));
            const float bzl = E1(bz, n - iy);
            E1(bz, n) = bzl + E1(bz, n) - q_ * temp - p * bzl;
            E1(by, n) = E1(by, n - iz) + E1(by, n) - q_ * bzl - p * temp;
        }
    }
}
#undef E1
template <int POST>
static int edge_call(tgpu_ctx *h, float *bx, float *by, float *bz, const float *ex, const float *ey, const float *ez,
                     long long ix, long long iy, long long iz, int mx, int my, int mz, long long m)
{
    EdgeArgs A; A.bx = bx; A.by = by; A.bz = bz; A.ex = ex; A.ey = ey; A.ez = ez; A.ix = ix; A.iy = iy; A.iz = iz; A.m = m;
    A.mx = mx; A.my = my; A.mz = mz; A.c = h->P.c;
    // preledge: sweeps 0, 1, 2, 3 (:2218-2243); postedge: the by/bz sweep first, then 0, 1, 2 (:2393-2413)
    if (POST) { k_edge<POST, 3><<<cdiv(mx - 1, 128), 128, 0, h->stream>>>(A); CKK(h); }
    if (mx > 2) { k_edge<POST, 0><<<cdiv(mx - 2, 128), 128, 0, h->stream>>>(A); CKK(h); }
    k_edge<POST, 1><<<cdiv(my - 1, 128), 128, 0, h->stream>>>(A); CKK(h);
    k_edge<POST, 2><<<cdiv(mz - 1, 128), 128, 0, h->stream>>>(A); CKK(h);
    if (!POST) { k_edge<POST, 3><<<cdiv(mx - 1, 128), 128, 0, h->stream>>>(A); CKK(h); }
    return 0;
}
// which = 0 pre_bc_b, 1 post_bc_b, 2 pre_bc_e, 3 post_bc_e (fieldboundaries.F90:114-163, 437-482): the three rotated edge
// calls; the caller follows with bc_b1 / bc_e1.  Acts only in a 3D box whose three axes all radiate.
int fld_edges(tgpu_ctx *h, int which)
{
    const tgpu_params &P = h->P;
    if (P.dim != 3 || P.periodicx || P.periodicy || P.periodicz) return 0;
    float *ex = h->f[0], *ey = h->f[1], *ez = h->f[2], *bx = h->f[3], *by = h->f[4], *bz = h->f[5];
    const long long ix = 1, iy = P.mx, iz = (long long)P.mx * P.my, lot = h->G.lot;
    const int mx = P.mx, my = P.my, mz = P.mz;
    int rc = 0;
    if (which == 0) {
        rc |= edge_call<0>(h, by, bz, bx, ey, ez, ex, iy, iz, ix, my, mz, mx, 1);
        rc |= edge_call<0>(h, bz, bx, by, ez, ex, ey, iz, ix, iy, mz, mx, my, 1);
        rc |= edge_call<0>(h, bx, by, bz, ex, ey, ez, ix, iy, iz, mx, my, mz, 1);
    } else if (which == 1) {
        rc |= edge_call<1>(h, by, bz, bx, ey, ez, ex, iy, iz, ix, my, mz, mx, 1);
        rc |= edge_call<1>(h, bz, bx, by, ez, ex, ey, iz, ix, iy, mz, mx, my, 1);
        rc |= edge_call<1>(h, bx, by, bz, ex, ey, ez, ix, iy, iz, mx, my, mz, 1);
    } else if (which == 2) {
        rc |= edge_call<0>(h, ey, ez, ex, by, bz, bx, -iy, -iz, -ix, my, mz, mx, lot);
        rc |= edge_call<0>(h, ez, ex, ey, bz, bx, by, -iz, -ix, -iy, mz, mx, my, lot);
        rc |= edge_call<0>(h, ex, ey, ez, bx, by, bz, -ix, -iy, -iz, mx, my, mz, lot);
    } else {
        rc |= edge_call<1>(h, ey, ez, ex, by, bz, bx, -iy, -iz, -ix, my, mz, mx, lot);
        rc |= edge_call<1>(h, ez, ex, ey, bz, bx, by, -iz, -ix, -iy, mz, mx, my, lot);
        rc |= edge_call<1>(h, ex, ey, ez, bx, by, bz, -ix, -iy, -iz, mx, my, mz, lot);
    }
    h->need_prim = 1;
    return rc;
}

// ---------------------------------------------------------------------------------------------
// field_bc_user of the shock problem: user/user_shock.F90:342-373
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_fill_x(float *__restrict__ a, int mx, int my, int mz, int i1, int i2, float v)
{
    size_t n = (size_t)(i2 - i1 + 1) * my * mz;
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    int w = i2 - i1 + 1;
    int i = i1 + (int)(idx % w); size_t r = idx / w;
    int j = 1 + (int)(r % my), k = 1 + (int)(r / my);
    a[LIDX(i, j, k)] = v;
}
static int iloc_x(const tgpu_ctx *h, int iglob)
{
    int i = iglob - h->P.mxcum;                        // fields.F90:384-390
    if (i < h->P.nghost / 2 + 1) i = 1;
    if (i > h->P.mx - h->P.nghost / 2) i = h->P.mx;
    return i;
}
static int fill_x(tgpu_ctx *h, int which, int i1, int i2, float v)
{
    size_t n = (size_t)(i2 - i1 + 1) * h->P.my * h->P.mz;
    k_fill_x<<<cdiv(n, 256), 256, 0, h->stream>>>(h->f[which], h->P.mx, h->P.my, h->P.mz, i1, i2, v);
    CKK(h);
    return 0;
}
int fld_bc_shock(tgpu_ctx *h, float leftwall, float binit, float btheta, float bphi, float beta)
{
    // global mx0 (ghosts included) = x2in + nghost/2 (particles.F90:339-344)
    const int mx0g = (int)h->P.x2in + h->P.nghost / 2;
    float xmin = 1.f, xmax = leftwall - 10.f;
    int i1 = iloc_x(h, (int)xmin), i2 = iloc_x(h, (int)xmax), rc = 0;
    if (i1 != i2) { rc |= fill_x(h, 1, i1, i2, 0.f); rc |= fill_x(h, 2, i1, i2, 0.f); }
    xmin = mx0g - 2.f; xmax = (float)mx0g;
    i1 = iloc_x(h, (int)xmin); i2 = iloc_x(h, (int)xmax);
    if (i1 != i2) {
        const float bxv = binit * cosf(btheta), byv = binit * sinf(btheta) * sinf(bphi), bzv = binit * sinf(btheta) * cosf(bphi);
        rc |= fill_x(h, 3, i1, i2, bxv); rc |= fill_x(h, 4, i1, i2, byv); rc |= fill_x(h, 5, i1, i2, bzv);
        rc |= fill_x(h, 0, i1, i2, 0.f); rc |= fill_x(h, 1, i1, i2, (-beta) * bzv); rc |= fill_x(h, 2, i1, i2, -(-beta) * byv);
    }
    h->need_prim = 1;
    return rc ? TGPU_ECUDA : 0;
}
