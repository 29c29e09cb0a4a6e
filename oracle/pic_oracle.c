/*
 * pic_oracle.c -- CPU restatement of the TRISTAN-MP (PU fork, Esirkepov branch)
 * per-timestep PIC hot path.  TEST INFRASTRUCTURE ONLY; see pic_oracle.h.
 *
 * PARITY: pinned bit-exact against the reference's own source text executed by
 * tests/golden/f90run.py (tests/test_ref_golden.py); the reference cannot be
 * compiled here.  Every routine cites the reference file:line it restates.
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -fPIC -shared (oracle/Makefile).
 * -ffp-contract=off keeps every fp32 operation separately rounded, in the
 * order the Fortran source writes it.
 */
#include "pic_oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define IDX(r, i, j, k) \
    ((size_t)((i)-1) + (size_t)(r)->mx * ((size_t)((j)-1) + (size_t)(r)->my * (size_t)((k)-1)))

static int imodulo(int a, int b) { int m = a % b; return m < 0 ? m + b : m; }
static float fsign(float a, float b) { return b >= 0.f ? fabsf(a) : -fabsf(a); } /* Fortran sign() */

size_t orc_sizeof_particle(void) { return sizeof(orc_particle); }

/* ------------------------------------------------------------------------- */
/* geometry: fields.F90:154-192 (ghost widths), :255-338 (decomposition)      */
/* ------------------------------------------------------------------------- */
static void ghost_widths(int dim, int order, int *nghost, int *nghostz)
{
    /* fields.F90:166-184 */
    if (order <= 1) { *nghost = 5; *nghostz = 5; }
    else            { *nghost = 7; *nghostz = 7; }
    if (dim == 2) *nghostz = 5;
}

int orc_range42_ok(const orc_params *P);
orc_world *orc_world_create(const orc_params *Pin)
{
    if (!orc_range42_ok(Pin)) return NULL;               /* _42 ranges that read outside the arrays in the reference */
    orc_world *w = (orc_world *)calloc(1, sizeof(orc_world));
    w->P = *Pin;
    orc_params *P = &w->P;
    if (P->dim == 2) P->sizez = 1;                       /* communications.F90:163-166 */
    int size0 = P->sizex * P->sizey * P->sizez;
    w->size0 = size0;
    int nghost, nghostz;
    ghost_widths(P->dim, P->order, &nghost, &nghostz);
    /* read_input_grid, fields.F90:186-189 */
    w->mx0g = P->mx0 + nghost;
    w->my0g = P->my0 + nghost;
    w->mz0g = P->mz0 + nghostz;
    if (P->dim == 2) w->mz0g = 1;                        /* fields.F90:228-232 */
    w->r = (orc_rank **)calloc(size0, sizeof(orc_rank *));
    w->mxl = (int *)calloc(size0, sizeof(int));
    w->myl = (int *)calloc(size0, sizeof(int));
    w->mzl = (int *)calloc(size0, sizeof(int));
    int sx = P->sizex, sy = P->sizey, sz = P->sizez;
    for (int rk = 0; rk < size0; rk++) {
        /* fields.F90:259-280 */
        int mx = (w->mx0g - nghost) / sx + nghost;
        int my = (w->my0g - nghost) / sy + nghost;
        int mz = (P->dim == 2) ? 1 : (w->mz0g - nghostz) / sz + nghostz;
        if (rk % sx == sx - 1 && w->mx0g != (mx - nghost) * sx + nghost)
            mx = w->mx0g - (mx - nghost) * (sx - 1);
        if ((rk % (sx * sy)) / sx == sy - 1 && w->my0g != (my - nghost) * sy + nghost)
            my = w->my0g - (my - nghost) * (sy - 1);
        if (P->dim == 3 && rk / (sx * sy) == sz - 1 && w->mz0g != (mz - nghostz) * sz + nghostz)
            mz = w->mz0g - (mz - nghostz) * (sz - 1);
        w->mxl[rk] = mx; w->myl[rk] = my; w->mzl[rk] = mz;
    }
    for (int rk = 0; rk < size0; rk++) {
        orc_rank *r = (orc_rank *)calloc(1, sizeof(orc_rank));
        w->r[rk] = r;
        r->P = *P; r->rank = rk; r->size0 = size0;
        r->nghost = nghost; r->nghostz = nghostz;
        r->mx = w->mxl[rk]; r->my = w->myl[rk]; r->mz = w->mzl[rk];
        r->mxl = w->mxl; r->myl = w->myl; r->mzl = w->mzl;
        /* fields.F90:316-328 */
        r->mxcum = 0; for (int i = 0; i < rk % sx; i++) r->mxcum += w->mxl[(rk / sx) * sx + i] - nghost;
        r->mycum = 0; for (int j = 0; j < (rk % (sx * sy)) / sx; j++) r->mycum += w->myl[j * sx] - nghost;
        r->mzcum = 0;
        if (P->dim == 3) for (int k = 0; k < rk / (sx * sy); k++) r->mzcum += w->mzl[k * sx * sy] - nghostz;
        r->iy = r->mx; r->iz = (P->dim == 2) ? 0 : r->mx * r->my;
        r->lot = (size_t)r->mx * r->my * r->mz;
        for (int a = 0; a < ORC_NARR; a++) r->f[a] = (float *)calloc(r->lot, sizeof(float));
        r->temp = (float *)calloc(r->lot, sizeof(float));
        r->maxhlf = P->maxptl / 2;
        r->p = (orc_particle *)calloc((size_t)P->maxptl + 1, sizeof(orc_particle));
        r->pind = (int32_t *)calloc((size_t)P->maxptl + 1, sizeof(int32_t));
        for (int d = 0; d < 6; d++) {
            r->out[d].p = (orc_particle *)calloc((size_t)P->buffsize + 1, sizeof(orc_particle));
            r->in[d].p = (orc_particle *)calloc((size_t)P->buffsize + 1, sizeof(orc_particle));
        }
        /* particles.F90:339-344 : x1in = nghost/2+1, x2in = mx0-nghost/2 (global) */
        r->x1in = (float)(nghost / 2 + 1);  r->x2in = (float)(w->mx0g - nghost / 2);
        r->y1in = (float)(nghost / 2 + 1);  r->y2in = (float)(w->my0g - nghost / 2);
        r->z1in = (float)(nghostz / 2 + 1); r->z2in = (float)(w->mz0g - nghostz / 2);
        r->mx0g = w->mx0g;
        r->dseed = 123457.0 + rk;                           /* communications.F90:228-229 */
        r->totalpartnum = 0;
    }
    return w;
}

void orc_world_destroy(orc_world *w)
{
    if (!w) return;
    for (int rk = 0; rk < w->size0; rk++) {
        orc_rank *r = w->r[rk];
        for (int a = 0; a < ORC_NARR; a++) free(r->f[a]);
        free(r->temp); free(r->p); free(r->pind);
        for (int d = 0; d < 6; d++) { free(r->out[d].p); free(r->in[d].p); }
        free(r);
    }
    free(w->r); free(w->mxl); free(w->myl); free(w->mzl); free(w);
}

orc_rank *orc_world_rank(orc_world *w, int r) { return w->r[r]; }
float *orc_rank_array(orc_rank *r, int which) { return r->f[which]; }
orc_particle *orc_rank_particles(orc_rank *r) { return r->p; }
void orc_rank_dims(const orc_rank *r, int *o)
{
    o[0] = r->mx; o[1] = r->my; o[2] = r->mz; o[3] = r->nghost; o[4] = r->nghostz;
    o[5] = r->mxcum; o[6] = r->mycum; o[7] = r->mzcum; o[8] = r->maxhlf;
}
void orc_rank_counts(const orc_rank *r, int *ions, int *lecs) { *ions = r->ions; *lecs = r->lecs; }
void orc_rank_set_counts(orc_rank *r, int ions, int lecs) { r->ions = ions; r->lecs = lecs; }
orc_particle *orc_rank_box(orc_rank *r, int which, int dir, int *nion, int *nlec)
{
    orc_box *b = which ? &r->in[dir] : &r->out[dir];
    if (nion) *nion = b->nion; if (nlec) *nlec = b->nlec;
    return b->p;
}
void orc_rank_box_set_counts(orc_rank *r, int which, int dir, int nion, int nlec)
{
    orc_box *b = which ? &r->in[dir] : &r->out[dir];
    b->nion = nion; b->nlec = nlec;
}

/* neighbour formulas: fieldboundaries.F90:1170-1171 (x), :1311-1314 (y); particles.F90:1892-1895 (z) */
int orc_neighbour(const orc_rank *r, int dir)
{
    int rank = r->rank, sx = r->P.sizex, sy = r->P.sizey, sz = r->P.sizez;
    switch (dir) {
    case 0: return (rank / sx) * sx + imodulo(rank - 1, sx);
    case 1: return (rank / sx) * sx + imodulo(rank + 1, sx);
    case 2: return imodulo(rank / sx - 1, sy) * sx + rank / (sx * sy) * (sx * sy) + imodulo(rank, sx);
    case 3: return imodulo(rank / sx + 1, sy) * sx + rank / (sx * sy) * (sx * sy) + imodulo(rank, sx);
    case 4: return imodulo(rank / (sx * sy) - 1, sz) * (sx * sy) + imodulo(rank, sx * sy);
    default:return imodulo(rank / (sx * sy) + 1, sz) * (sx * sy) + imodulo(rank, sx * sy);
    }
}
static int rank_ix(const orc_rank *r) { return r->rank % r->P.sizex; }
static int rank_iy(const orc_rank *r) { return (r->rank % (r->P.sizex * r->P.sizey)) / r->P.sizex; }
static int rank_iz(const orc_rank *r) { return r->rank / (r->P.sizex * r->P.sizey); }

/* ------------------------------------------------------------------------- */
/* Yee solver: fields.F90:586-728 (B half step), :739-870 (E full step)        */
/* ------------------------------------------------------------------------- */
static void stencil_range_b(const orc_rank *r, int axis, int *a1, int *a2)
{
    /* fields.F90:599-669 */
    int g = (axis == 2 ? r->nghostz : r->nghost) / 2;
    int m = axis == 0 ? r->mx : axis == 1 ? r->my : r->mz;
    int per = axis == 0 ? r->P.periodicx : axis == 1 ? r->P.periodicy : r->P.periodicz;
    int sz = axis == 0 ? r->P.sizex : axis == 1 ? r->P.sizey : r->P.sizez;
    int pos = axis == 0 ? rank_ix(r) : axis == 1 ? rank_iy(r) : rank_iz(r);
    *a1 = g + 1; *a2 = m - (g + 1);
    if (!per) {
        if (pos == 0) { *a1 = 1; *a2 = m - (g + 1); }
        if (pos == sz - 1) { *a1 = g + 1; *a2 = m - 1; }
        if (axis == 2 ? (r->size0 == 1) : (r->size0 == 1 || sz == 1)) { *a1 = 1; *a2 = m - 1; }
    }
}
static void stencil_range_e(const orc_rank *r, int axis, int *a1, int *a2)
{
    /* fields.F90:752-819 */
    int g = (axis == 2 ? r->nghostz : r->nghost) / 2;
    int m = axis == 0 ? r->mx : axis == 1 ? r->my : r->mz;
    int per = axis == 0 ? r->P.periodicx : axis == 1 ? r->P.periodicy : r->P.periodicz;
    int sz = axis == 0 ? r->P.sizex : axis == 1 ? r->P.sizey : r->P.sizez;
    int pos = axis == 0 ? rank_ix(r) : axis == 1 ? rank_iy(r) : rank_iz(r);
    *a1 = g + 1; *a2 = m - (g + 1);
    if (!per) {
        if (pos == 0) { *a1 = g; *a2 = m - (g + 1); }
        if (pos == sz - 1) { *a1 = g + 1; *a2 = m; }
        if (axis == 2 ? (r->size0 == 1) : (r->size0 == 1 || sz == 1)) { *a1 = g; *a2 = m; }
    }
    if (axis == 2) { *a1 = g; *a2 = m; }                    /* unconditional override, fields.F90:818-819 */
}

/* ------------------------------------------------------------------------- */
/* 4th-order `_42` solver (highorder = 1): fields.F90:1039-1212 (B), 1223-1361 (E) */
/* Index ranges are the reference's.  Where those ranges make the reference read  */
/* outside its arrays (open y or z on a split axis; open z in the E step, :1262-   */
/* 1277 "FIX range") the configuration is rejected by range42_ok().               */
/* ------------------------------------------------------------------------- */
int orc_range42_ok(const orc_params *P)
{
    int size0 = P->sizex * P->sizey * P->sizez;
    if (!P->highorder) return 1;
    if (P->dim == 3 && !P->periodicz) return 0;
    if (!P->periodicy && size0 != 1) return 0;
    return 1;
}
static void range42_b(const orc_rank *r, int axis, int *a1, int *a2)
{
    int g = (axis == 2 ? r->nghostz : r->nghost) / 2;
    int m = axis == 0 ? r->mx : axis == 1 ? r->my : r->mz;
    int per = axis == 0 ? r->P.periodicx : axis == 1 ? r->P.periodicy : r->P.periodicz;
    if (per) { *a1 = g + 1; *a2 = m - (g + 1); }            /* :1054-1056, 1067-1069, 1088-1090 */
    else { *a1 = g; *a2 = m - g; }                          /* :1057-1059; y, z: the size0 == 1 branch (:1080-1083, 1102-1105) */
    if (axis == 0 && r->P.wall_i2 > 0 && r->P.wall_i2 < *a2) *a2 = r->P.wall_i2;   /* :1062-1065 */
}
static void range42_e(const orc_rank *r, int axis, int *a1, int *a2)
{
    int g = (axis == 2 ? r->nghostz : r->nghost) / 2;
    int m = axis == 0 ? r->mx : axis == 1 ? r->my : r->mz;
    int per = axis == 0 ? r->P.periodicx : axis == 1 ? r->P.periodicy : r->P.periodicz;
    *a1 = g + 1; *a2 = per ? m - (g + 1) : m - 1;           /* :1238-1244, 1251-1257, 1260-1262 */
    if (axis == 0 && r->P.wall_i2 > 0 && r->P.wall_i2 < *a2) *a2 = r->P.wall_i2;   /* :1246-1249 */
}
static void advance_b_halfstep_42(orc_rank *r)
{
    float *ex = r->f[ORC_EX], *ey = r->f[ORC_EY], *ez = r->f[ORC_EZ];
    float *bx = r->f[ORC_BX], *by = r->f[ORC_BY], *bz = r->f[ORC_BZ];
    int i1, i2, j1, j2, k1 = 1, k2 = 1;
    range42_b(r, 0, &i1, &i2); range42_b(r, 1, &j1, &j2);
    if (r->P.dim == 3) range42_b(r, 2, &k1, &k2);
    const float cnst = r->P.corr * (.5f * r->P.c);                       /* :1050-1052 */
    const float coef1 = 9.f / 8.f * r->P.corr * (.5f * r->P.c);
    const float coef2 = -1.f / 24.f * r->P.corr * (.5f * r->P.c);
    const int three = r->P.dim == 3;
#define L(i, j, k) IDX(r, i, j, k)
    for (int k = k1; k <= k2; k++) for (int j = j1; j <= j2; j++) for (int i = i1; i <= i2; i++) {
        const size_t l = L(i, j, k);
        if (three) {                                                     /* :1124-1132 */
            bx[l] = bx[l] + coef1 * (ey[L(i, j, k + 1)] - ey[l] - ez[L(i, j + 1, k)] + ez[l])
                          + coef2 * (ey[L(i, j, k + 2)] - ey[L(i, j, k - 1)] - ez[L(i, j + 2, k)] + ez[L(i, j - 1, k)]);
            by[l] = by[l] + coef1 * (ez[L(i + 1, j, k)] - ez[l] - ex[L(i, j, k + 1)] + ex[l])
                          + coef2 * (ez[L(i + 2, j, k)] - ez[L(i - 1, j, k)] - ex[L(i, j, k + 2)] + ex[L(i, j, k - 1)]);
        } else {                                                         /* :1147-1150 */
            bx[l] = bx[l] + coef1 * (-ez[L(i, j + 1, k)] + ez[l]) + coef2 * (-ez[L(i, j + 2, k)] + ez[L(i, j - 1, k)]);
            by[l] = by[l] + coef1 * (ez[L(i + 1, j, k)] - ez[l]) + coef2 * (ez[L(i + 2, j, k)] - ez[L(i - 1, j, k)]);
        }
        bz[l] = bz[l] + coef1 * (ex[L(i, j + 1, k)] - ex[l] - ey[L(i + 1, j, k)] + ey[l])
                      + coef2 * (ex[L(i, j + 2, k)] - ex[L(i, j - 1, k)] - ey[L(i + 2, j, k)] + ey[L(i - 1, j, k)]);
    }
    if (!r->P.periodicx) {                                               /* :1158-1190: 2nd-order update of i = 1 and mx-1 */
        for (int k = k1; k <= k2; k++) for (int j = j1; j <= j2; j++) for (int i = 1; i <= r->mx - 1; i += r->mx - 2) {
            const size_t l = L(i, j, k);
            if (three) {
                bx[l] = bx[l] + cnst * (ey[L(i, j, k + 1)] - ey[l] - ez[L(i, j + 1, k)] + ez[l]);
                by[l] = by[l] + cnst * (ez[L(i + 1, j, k)] - ez[l] - ex[L(i, j, k + 1)] + ex[l]);
            } else {
                bx[l] = bx[l] + cnst * (-ez[L(i, j + 1, k)] + ez[l]);
                by[l] = by[l] + cnst * (ez[L(i + 1, j, k)] - ez[l]);
            }
            bz[l] = bz[l] + cnst * (ex[L(i, j + 1, k)] - ex[l] - ey[L(i + 1, j, k)] + ey[l]);
        }
    }
}
static void advance_e_fullstep_42(orc_rank *r)
{
    float *ex = r->f[ORC_EX], *ey = r->f[ORC_EY], *ez = r->f[ORC_EZ];
    float *bx = r->f[ORC_BX], *by = r->f[ORC_BY], *bz = r->f[ORC_BZ];
    int i1, i2, j1, j2, k1 = 1, k2 = 1;
    range42_e(r, 0, &i1, &i2); range42_e(r, 1, &j1, &j2);
    if (r->P.dim == 3) range42_e(r, 2, &k1, &k2);
    const float cnst = r->P.corr * r->P.c;                               /* :1234-1236 */
    const float coef1 = 9.f / 8.f * r->P.corr * r->P.c;
    const float coef2 = -1.f / 24.f * r->P.corr * r->P.c;
    const int three = r->P.dim == 3;
    for (int k = k1; k <= k2; k++) for (int j = j1; j <= j2; j++) for (int i = i1; i <= i2; i++) {
        const size_t l = L(i, j, k);
        if (three) {                                                     /* :1293-1301 */
            ex[l] = ex[l] + coef1 * (by[L(i, j, k - 1)] - by[l] - bz[L(i, j - 1, k)] + bz[l])
                          + coef2 * (by[L(i, j, k - 2)] - by[L(i, j, k + 1)] - bz[L(i, j - 2, k)] + bz[L(i, j + 1, k)]);
            ey[l] = ey[l] + coef1 * (bz[L(i - 1, j, k)] - bz[l] - bx[L(i, j, k - 1)] + bx[l])
                          + coef2 * (bz[L(i - 2, j, k)] - bz[L(i + 1, j, k)] - bx[L(i, j, k - 2)] + bx[L(i, j, k + 1)]);
        } else {                                                         /* :1316-1319 */
            ex[l] = ex[l] + coef1 * (-bz[L(i, j - 1, k)] + bz[l]) + coef2 * (-bz[L(i, j - 2, k)] + bz[L(i, j + 1, k)]);
            ey[l] = ey[l] + coef1 * (bz[L(i - 1, j, k)] - bz[l]) + coef2 * (bz[L(i - 2, j, k)] - bz[L(i + 1, j, k)]);
        }
        ez[l] = ez[l] + coef1 * (bx[L(i, j - 1, k)] - bx[l] - by[L(i - 1, j, k)] + by[l])
                      + coef2 * (bx[L(i, j - 2, k)] - bx[L(i, j + 1, k)] - by[L(i - 2, j, k)] + by[L(i + 1, j, k)]);
    }
    if (!r->P.periodicx) {                                               /* :1327-1357: 2nd-order update of i = 2 and mx */
        for (int k = k1; k <= k2; k++) for (int j = j1; j <= j2; j++) for (int i = 2; i <= r->mx; i += r->mx - 2) {
            const size_t l = L(i, j, k);
            if (three) {
                ex[l] = ex[l] + cnst * (by[L(i, j, k - 1)] - by[l] - bz[L(i, j - 1, k)] + bz[l]);
                ey[l] = ey[l] + cnst * (bz[L(i - 1, j, k)] - bz[l] - bx[L(i, j, k - 1)] + bx[l]);
            } else {
                ex[l] = ex[l] + cnst * (-bz[L(i, j - 1, k)] + bz[l]);
                ey[l] = ey[l] + cnst * (bz[L(i - 1, j, k)] - bz[l]);
            }
            ez[l] = ez[l] + cnst * (bx[L(i, j - 1, k)] - bx[l] - by[L(i - 1, j, k)] + by[l]);
        }
    }
#undef L
}

void orc_advance_b_halfstep(orc_rank *r)
{
    if (r->P.highorder) { advance_b_halfstep_42(r); return; }          /* dispatcher, fields.F90:1407-1417 */
    float *ex = r->f[ORC_EX], *ey = r->f[ORC_EY], *ez = r->f[ORC_EZ];
    float *bx = r->f[ORC_BX], *by = r->f[ORC_BY], *bz = r->f[ORC_BZ];
    int i1, i2, j1, j2, k1 = 1, k2 = 1;
    stencil_range_b(r, 0, &i1, &i2);
    stencil_range_b(r, 1, &j1, &j2);
    if (r->P.dim == 3) stencil_range_b(r, 2, &k1, &k2);
    const float cnst = r->P.corr * (.5f * r->P.c);           /* fields.F90:672 */
    if (r->P.dim == 3) {
        for (int k = k1; k <= k2; k++) for (int j = j1; j <= j2; j++) for (int i = i1; i <= i2; i++) {
            size_t l = IDX(r, i, j, k), lip = IDX(r, i + 1, j, k), ljp = IDX(r, i, j + 1, k), lkp = IDX(r, i, j, k + 1);
            /* fields.F90:680-685 */
            bx[l] = bx[l] + cnst * (ey[lkp] - ey[l] - ez[ljp] + ez[l]);
            by[l] = by[l] + cnst * (ez[lip] - ez[l] - ex[lkp] + ex[l]);
            bz[l] = bz[l] + cnst * (ex[ljp] - ex[l] - ey[lip] + ey[l]);
        }
    } else {
        for (int j = j1; j <= j2; j++) for (int i = i1; i <= i2; i++) {
            size_t l = IDX(r, i, j, 1), lip = IDX(r, i + 1, j, 1), ljp = IDX(r, i, j + 1, 1);
            /* fields.F90:716-719 */
            bx[l] = bx[l] + cnst * (-ez[ljp] + ez[l]);
            by[l] = by[l] + cnst * (ez[lip] - ez[l]);
            bz[l] = bz[l] + cnst * (ex[ljp] - ex[l] - ey[lip] + ey[l]);
        }
    }
}

void orc_advance_e_fullstep(orc_rank *r)
{
    if (r->P.highorder) { advance_e_fullstep_42(r); return; }          /* dispatcher, fields.F90:1429-1440 */
    float *ex = r->f[ORC_EX], *ey = r->f[ORC_EY], *ez = r->f[ORC_EZ];
    float *bx = r->f[ORC_BX], *by = r->f[ORC_BY], *bz = r->f[ORC_BZ];
    int i1, i2, j1, j2, k1 = 1, k2 = 1;
    stencil_range_e(r, 0, &i1, &i2);
    stencil_range_e(r, 1, &j1, &j2);
    if (r->P.dim == 3) stencil_range_e(r, 2, &k1, &k2);
    const float cnst = r->P.corr * r->P.c;                   /* fields.F90:826 */
    if (r->P.dim == 3) {
        for (int k = k1; k <= k2; k++) for (int j = j1; j <= j2; j++) for (int i = i1; i <= i2; i++) {
            size_t l = IDX(r, i, j, k), lim = IDX(r, i - 1, j, k), ljm = IDX(r, i, j - 1, k), lkm = IDX(r, i, j, k - 1);
            /* fields.F90:836-838 */
            ex[l] = ex[l] + cnst * (by[lkm] - by[l] - bz[ljm] + bz[l]);
            ey[l] = ey[l] + cnst * (bz[lim] - bz[l] - bx[lkm] + bx[l]);
            ez[l] = ez[l] + cnst * (bx[ljm] - bx[l] - by[lim] + by[l]);
        }
    } else {
        for (int j = j1; j <= j2; j++) for (int i = i1; i <= i2; i++) {
            size_t l = IDX(r, i, j, 1), lim = IDX(r, i - 1, j, 1), ljm = IDX(r, i, j - 1, 1);
            /* fields.F90:860-863 */
            ex[l] = ex[l] + cnst * (-bz[ljm] + bz[l]);
            ey[l] = ey[l] + cnst * (bz[lim] - bz[l]);
            ez[l] = ez[l] + cnst * (bx[ljm] - bx[l] - by[lim] + by[l]);
        }
    }
}

/* fields.F90:499-506 */
void orc_reset_currents(orc_rank *r)
{
    for (int a = ORC_CURX; a <= ORC_CURZ; a++) memset(r->f[a], 0, r->lot * sizeof(float));
}
/* fields.F90:1391-1393 : whole arrays, ghosts included */
void orc_add_current(orc_rank *r)
{
    for (int a = 0; a < 3; a++) {
        float *e = r->f[ORC_EX + a]; const float *cu = r->f[ORC_CURX + a];
        for (size_t l = 0; l < r->lot; l++) e[l] = e[l] + cu[l];
    }
}

/* ------------------------------------------------------------------------- */
/* box primitives                                                              */
/* ------------------------------------------------------------------------- */
void orc_box_get(const orc_rank *r, int which, const int lo[3], const int hi[3], float *buf)
{
    const float *a = r->f[which]; size_t n = 0;
    for (int k = lo[2]; k <= hi[2]; k++) for (int j = lo[1]; j <= hi[1]; j++) for (int i = lo[0]; i <= hi[0]; i++)
        buf[n++] = a[IDX(r, i, j, k)];
}
void orc_box_put(orc_rank *r, int which, const int lo[3], const int hi[3], const float *buf)
{
    float *a = r->f[which]; size_t n = 0;
    for (int k = lo[2]; k <= hi[2]; k++) for (int j = lo[1]; j <= hi[1]; j++) for (int i = lo[0]; i <= hi[0]; i++)
        a[IDX(r, i, j, k)] = buf[n++];
}
void orc_box_add(orc_rank *r, int which, const int lo[3], const int hi[3], const float *buf)
{
    float *a = r->f[which]; size_t n = 0;
    for (int k = lo[2]; k <= hi[2]; k++) for (int j = lo[1]; j <= hi[1]; j++) for (int i = lo[0]; i <= hi[0]; i++) {
        size_t l = IDX(r, i, j, k); a[l] = a[l] + buf[n++];
    }
}
static void full_box(const orc_rank *r, int lo[3], int hi[3])
{
    lo[0] = lo[1] = lo[2] = 1; hi[0] = r->mx; hi[1] = r->my; hi[2] = r->mz;
}
static size_t box_count(const int lo[3], const int hi[3])
{
    return (size_t)(hi[0] - lo[0] + 1) * (hi[1] - lo[1] + 1) * (hi[2] - lo[2] + 1);
}
static int axis_m(const orc_rank *r, int axis) { return axis == 0 ? r->mx : axis == 1 ? r->my : r->mz; }
static int axis_g(const orc_rank *r, int axis) { return (axis == 2 ? r->nghostz : r->nghost) / 2; }
static int axis_ng(const orc_rank *r, int axis) { return axis == 2 ? r->nghostz : r->nghost; }
static int axis_per(const orc_rank *r, int axis) { return axis == 0 ? r->P.periodicx : axis == 1 ? r->P.periodicy : r->P.periodicz; }
static int axis_size(const orc_rank *r, int axis) { return axis == 0 ? r->P.sizex : axis == 1 ? r->P.sizey : r->P.sizez; }
static int axis_pos(const orc_rank *r, int axis) { return axis == 0 ? rank_ix(r) : axis == 1 ? rank_iy(r) : rank_iz(r); }

/* ------------------------------------------------------------------------- */
/* one-layer copy between neighbours:                                           */
/* copy_layr{x,y,z}1_opt (periodic) / 2_opt (open), copylayrx/y (local)         */
/* fieldboundaries.F90:714-785, 1149-1428, 1616-1756                            */
/* low target lt <- (-neighbour) layer ls ; high target nt <- (+neighbour) ns  */
/* Open axis: edge ranks skip the receive on their outer face (:1404-1426).     */
/* On open axes with one rank on the axis nothing is copied (bc_b1 :204-212).   */
/* ------------------------------------------------------------------------- */
/* bc_b1 / bc_e1: fieldboundaries.F90:181-263, 306-392 */
void orc_bc_fields(orc_world *w, int first)
{
    orc_rank *r0 = w->r[0];
    int naxes = w->P.dim == 3 ? 3 : 2;
    for (int axis = 0; axis < naxes; axis++) {
        int g = axis_g(r0, axis);
        for (int iter = 1; iter <= g; iter++) {
            /* (lt,ls,nt,ns) = (g+1-iter, m-(g+iter), m-(g+1)+iter, g+iter), m rank-local.
               copy_layr*1_opt (periodic) / *2_opt (open: edge ranks skip the outer receive,
               :1404-1426) / copylayrx,y (one rank on the axis, :714-785) */
            int n = w->size0;
            float **sbuf = (float **)calloc(n, sizeof(float *));
            for (int pass = 0; pass < 2; pass++) {
                for (int rk = 0; rk < n; rk++) {
                    orc_rank *r = w->r[rk]; int lo[3], hi[3]; full_box(r, lo, hi);
                    int m = axis_m(r, axis);
                    lo[axis] = hi[axis] = pass == 0 ? m - (g + iter) : g + iter;
                    size_t cnt = box_count(lo, hi);
                    sbuf[rk] = (float *)malloc(3 * cnt * sizeof(float));
                    for (int c = 0; c < 3; c++) orc_box_get(r, first + c, lo, hi, sbuf[rk] + c * cnt);
                }
                for (int rk = 0; rk < n; rk++) {
                    orc_rank *r = w->r[rk]; int lo[3], hi[3]; full_box(r, lo, hi);
                    int m = axis_m(r, axis);
                    int src = orc_neighbour(r, pass == 0 ? 2 * axis : 2 * axis + 1);
                    int per = axis_per(r, axis), pos = axis_pos(r, axis), sz = axis_size(r, axis);
                    int skip = 0;
                    if (!per) {
                        if (sz == 1 && axis != 2) skip = 1;
                        if (pass == 0 && pos == 0) skip = 1;
                        if (pass == 1 && pos == sz - 1) skip = 1;
                    }
                    if (!skip) {
                        lo[axis] = hi[axis] = pass == 0 ? g + 1 - iter : m - (g + 1) + iter;
                        size_t cnt = box_count(lo, hi);
                        for (int c = 0; c < 3; c++) orc_box_put(r, first + c, lo, hi, sbuf[src] + c * cnt);
                    }
                }
                for (int rk = 0; rk < n; rk++) free(sbuf[rk]);
            }
            free(sbuf);
        }
    }
}

/* ------------------------------------------------------------------------- */
/* radiation boundary `surface` (Lindman-type absorbing face):                  */
/* fieldboundaries.F90:493-606, called by bc_b2 (:274-295) on the high face of   */
/* every radiating axis and by bc_e2 (:403-426) on the low face with mirrored    */
/* strides and the roles of E and B exchanged.  The arguments keep the           */
/* reference's meaning: 1-based flat indices, strides (ix,iy,iz) of the three    */
/* rotated axes (iz = the axis normal to the face), first element m00.           */
/* The #ifdef twoD variants (one stride is zero) are the two branches below.     */
/* ------------------------------------------------------------------------- */
static void surface(float *bx, float *by, float *bz, const float *ex, const float *ey, const float *ez,
                    long ix, long iy, long iz, int mx, int my, int mz, long m00, float c, int twod)
{
#define A1(a, n) (a)[(n) - 1]
    const float rs = 2.f * c / (1.f + c);
    const float s = .4142136f;
    const float os = .5f * (1.f - s) * rs;
    const long mf = m00 + iz * (mz - 1);                      /* first element of the face */
#define BZ_HALF(n) A1(bz, n) = A1(bz, n) + .5f * c * (A1(ex, (n) + iy) - A1(ex, n) - A1(ey, (n) + ix) + A1(ey, n))
#define BX_UPD(n) A1(bx, n) = A1(bx, n) + rs * (A1(bx, (n) - iz) - A1(bx, n) + s * (A1(bz, n) - A1(bz, (n) - ix)))      \
        - os * (A1(ez, (n) + iy) - A1(ez, n)) - (os - c) * (A1(ez, (n) + iy - iz) - A1(ez, (n) - iz))                  \
        - c * (A1(ey, n) - A1(ey, (n) - iz))
#define BY_UPD(n) A1(by, n) = A1(by, n) + rs * (A1(by, (n) - iz) - A1(by, n) + s * (A1(bz, n) - A1(bz, (n) - iy)))      \
        + os * (A1(ez, (n) + ix) - A1(ez, n)) + (os - c) * (A1(ez, (n) + ix - iz) - A1(ez, (n) - iz))                  \
        + c * (A1(ex, n) - A1(ex, (n) - iz))
    if (!twod) {                                              /* :515-537 */
        for (int jj = 0; jj <= my - 2; jj++) {
            const long m = mf + iy * jj;
            for (int ii = 0; ii <= mx - 2; ii++) { const long n = m + ix * ii; BZ_HALF(n); }
            for (int ii = 1; ii <= mx - 2; ii++) { const long n = m + ix * ii; BX_UPD(n); }
        }
        for (int ii = 0; ii <= mx - 2; ii++) {
            const long m = mf + ix * ii;
            for (int jj = 1; jj <= my - 2; jj++) { const long n = m + iy * jj; BY_UPD(n); }
            for (int jj = 0; jj <= my - 2; jj++) { const long n = m + iy * jj; BZ_HALF(n); }
        }
    } else if (ix == 0) {                                     /* :540-560 */
        for (int jj = 0; jj <= my - 2; jj++) { const long n = mf + iy * jj; BZ_HALF(n); BX_UPD(n); }
        for (int jj = 1; jj <= my - 2; jj++) { const long n = mf + iy * jj; BY_UPD(n); }
        for (int jj = 0; jj <= my - 2; jj++) { const long n = mf + iy * jj; BZ_HALF(n); }
    } else if (iy == 0) {                                     /* :565-583 */
        for (int ii = 0; ii <= mx - 2; ii++) { const long n = mf + ix * ii; BZ_HALF(n); }
        for (int ii = 1; ii <= mx - 2; ii++) { const long n = mf + ix * ii; BX_UPD(n); }
        for (int ii = 0; ii <= mx - 2; ii++) { const long n = mf + ix * ii; BY_UPD(n); BZ_HALF(n); }
    }
#undef BZ_HALF
#undef BX_UPD
#undef BY_UPD
#undef A1
}
/* radiation flags: fieldboundaries.F90:90-94 */
static void radiation_flags(const orc_rank *r, int rad[3])
{
    rad[0] = 1 - r->P.periodicx; rad[1] = 1 - r->P.periodicy; rad[2] = 1 - r->P.periodicz;
    /* the "open y opens z" promotion (:91-93) is #ifdef twoD only, and there the z call is compiled out (:287-291, 418-422) */
    if (r->P.dim == 2) rad[2] = 0;
}
/* the `surface` part of bc_b2 (high faces) / bc_e2 (low faces); the ghost refresh that follows is orc_bc_fields */
void orc_surface_b(orc_rank *r)
{
    float *ex = r->f[ORC_EX], *ey = r->f[ORC_EY], *ez = r->f[ORC_EZ];
    float *bx = r->f[ORC_BX], *by = r->f[ORC_BY], *bz = r->f[ORC_BZ];
    const long ix = 1, iy = r->mx, iz = r->P.dim == 3 ? (long)r->mx * r->my : 0;
    const int mx = r->mx, my = r->my, mz = r->P.dim == 3 ? r->mz : 1, twod = r->P.dim == 2;
    int rad[3]; radiation_flags(r, rad);
    const float c = r->P.c;
    if (rad[0]) surface(by, bz, bx, ey, ez, ex, iy, iz, ix, my, mz, mx, 1, c, twod);     /* :279-281 */
    if (rad[1]) surface(bz, bx, by, ez, ex, ey, iz, ix, iy, mz, mx, my, 1, c, twod);     /* :283-285 */
    if (rad[2]) surface(bx, by, bz, ex, ey, ez, ix, iy, iz, mx, my, mz, 1, c, twod);     /* :288-290 */
}
void orc_surface_e(orc_rank *r)
{
    float *ex = r->f[ORC_EX], *ey = r->f[ORC_EY], *ez = r->f[ORC_EZ];
    float *bx = r->f[ORC_BX], *by = r->f[ORC_BY], *bz = r->f[ORC_BZ];
    const long ix = 1, iy = r->mx, iz = r->P.dim == 3 ? (long)r->mx * r->my : 0;
    const int mx = r->mx, my = r->my, mz = r->P.dim == 3 ? r->mz : 1, twod = r->P.dim == 2;
    int rad[3]; radiation_flags(r, rad);
    const float c = r->P.c; const long lot = (long)r->lot;
    if (rad[0]) surface(ey, ez, ex, by, bz, bx, -iy, -iz, -ix, my, mz, mx, lot, c, twod);   /* :410-412 */
    if (rad[1]) surface(ez, ex, ey, bz, bx, by, -iz, -ix, -iy, mz, mx, my, lot, c, twod);   /* :414-416 */
    if (rad[2]) surface(ex, ey, ez, bx, by, bz, -ix, -iy, -iz, mx, my, mz, lot, c, twod);   /* :419-421 */
}

/* ------------------------------------------------------------------------- */
/* edge fixes of an all-open 3D box: preledge (fieldboundaries.F90:2200-2244) and */
/* postedge (:2371-2407), called three times each with rotated arrays by          */
/* pre_bc_b / post_bc_b / pre_bc_e / post_bc_e (:114-163, 437-482) when all three  */
/* axes radiate; each group is followed by bc_b1 / bc_e1.  Arguments keep the      */
/* reference's meaning (1-based flat index n, strides ix,iy,iz, first element m).  */
/* Only the #ifndef twoD bodies exist: the callers compile the calls out in 2D.    */
/* ------------------------------------------------------------------------- */
static void preledge(float *bx, float *by, float *bz, const float *ex, const float *ey, const float *ez,
                     long ix, long iy, long iz, int mx, int my, int mz, long m, float c)
{
#define A1(a, n) (a)[(n) - 1]
    const float s = .4142136f;
    const float t = c / (2.f * c + 1.f + s);
    for (long n = m + iy * (my - 1) + iz * (mz - 1) + ix, q = 0; q < mx - 2; q++, n += ix)
        A1(bx, n) = A1(bx, n - iy - iz) + (1.f - 4.f * t) * A1(bx, n) + (1.f - 2.f * t) * (A1(bx, n - iy) + A1(bx, n - iz))
            + s * t * (A1(by, n) - A1(by, n - ix) + A1(by, n - iz) - A1(by, n - ix - iz)
                       + A1(bz, n) - A1(bz, n - ix) + A1(bz, n - iy) - A1(bz, n - ix - iy));
    const float r = 4.f / (2.f * c + 2.f + s);
    for (long n = m + iz * (mz - 1), q = 0; q < my - 1; q++, n += iy)
        A1(bx, n) = (1.f - c * r) * (A1(bx, n) + A1(bx, n + ix)) + A1(bx, n - iz) + A1(bx, n + ix - iz) - r * (A1(bz, n)
            + c * ((1.f - s) * (A1(ex, n + iy) - A1(ex, n)) + (1.f + s) * .25f * (A1(ez, n + iy)
            - A1(ez, n) + A1(ez, n + ix + iy) - A1(ez, n + ix) + A1(ez, n + iy - iz) - A1(ez, n - iz)
            + A1(ez, n + ix + iy - iz) - A1(ez, n + ix - iz))));
    for (long n = m + iy * (my - 1), q = 0; q < mz - 1; q++, n += iz)
        A1(bx, n) = (1.f - c * r) * (A1(bx, n) + A1(bx, n + ix)) + A1(bx, n - iy) + A1(bx, n + ix - iy)
            - r * (A1(by, n) - c * ((1.f - s) * (A1(ex, n + iz) - A1(ex, n))
            + (1.f + s) * .25f * (A1(ey, n + iz) - A1(ey, n) + A1(ey, n + ix + iz)
            - A1(ey, n + ix) + A1(ey, n + iz - iy) - A1(ey, n - iy) + A1(ey, n + ix + iz - iy)
            - A1(ey, n + ix - iy))));
    const float p = (1.f + c) * 2.f / (1.f + 2.f * c * (1.f + c * s));
    const float q_ = c * s * 2.f / (1.f + 2.f * c * (1.f + c * s));
    for (long n = m + iy * (my - 1) + iz * (mz - 1), q = 0; q < mx - 1; q++, n += ix) {
        const float temp = A1(bz, n) - .5f * c * (1.f - s) * (A1(ey, n + ix) - A1(ey, n) + A1(ey, n + ix - iy) - A1(ey, n - iy));
        A1(bz, n) = A1(bz, n - iy) - A1(bz, n) + p * temp + q_ * A1(by, n);
        A1(by, n) = A1(by, n - iz) - A1(by, n) + p * A1(by, n) + q_ * temp;
    }
#undef A1
}
static void postedge(float *bx, float *by, float *bz, const float *ex, const float *ey, const float *ez,
                     long ix, long iy, long iz, int mx, int my, int mz, long m, float c)
{
#define A1(a, n) (a)[(n) - 1]
    const float s = .4142136f;
    const float p = (1.f + c) * 2.f / (1.f + 2.f * c * (1.f + c * s));
    const float q_ = c * s * 2.f / (1.f + 2.f * c * (1.f + c * s));
    for (long n = m + iy * (my - 1) + iz * (mz - 1), q = 0; q < mx - 1; q++, n += ix) {
        const float temp = A1(by, n - iz) - .5f * c * (1.f - s) * (A1(ez, n + ix) - A1(ez, n) + A1(ez, n + ix - iz) - A1(ez, n - iz));
        A1(bz, n) = A1(bz, n - iy) + A1(bz, n) - q_ * temp - p * A1(bz, n - iy);
        A1(by, n) = A1(by, n - iz) + A1(by, n) - q_ * A1(bz, n - iy) - p * temp;
    }
    const float t = c / (2.f * c + 1.f + s);
    for (long n = m + iy * (my - 1) + iz * (mz - 1) + ix, q = 0; q < mx - 2; q++, n += ix)
        A1(bx, n) = A1(bx, n) - (1.f - 4.f * t) * A1(bx, n - iy - iz) - (1.f - 2.f * t) * (A1(bx, n - iy) + A1(bx, n - iz))
            + s * t * (A1(by, n) - A1(by, n - ix) + A1(by, n - iz) - A1(by, n - ix - iz)
                       + A1(bz, n) - A1(bz, n - ix) + A1(bz, n - iy) - A1(bz, n - ix - iy));
    const float r = 4.f / (2.f * c + 2.f + s);
    for (long n = m + iz * (mz - 1), q = 0; q < my - 1; q++, n += iy)
        A1(bx, n) = A1(bx, n) - A1(bx, n + ix) - (1.f - c * r) * (A1(bx, n - iz) + A1(bx, n + ix - iz)) + r * A1(bz, n);
    for (long n = m + iy * (my - 1), q = 0; q < mz - 1; q++, n += iz)
        A1(bx, n) = A1(bx, n) - A1(bx, n + ix) - (1.f - c * r) * (A1(bx, n - iy) + A1(bx, n + ix - iy)) + r * A1(by, n);
#undef A1
}
/* which = 0 pre_bc_b, 1 post_bc_b, 2 pre_bc_e, 3 post_bc_e: the three edge calls (the ghost refresh that follows is
   orc_bc_fields).  No-op unless the box is 3D with all three axes open (fieldboundaries.F90:120, 154, 442, 468). */
void orc_edges(orc_rank *r, int which)
{
    if (r->P.dim != 3 || r->P.periodicx || r->P.periodicy || r->P.periodicz) return;
    float *ex = r->f[ORC_EX], *ey = r->f[ORC_EY], *ez = r->f[ORC_EZ];
    float *bx = r->f[ORC_BX], *by = r->f[ORC_BY], *bz = r->f[ORC_BZ];
    const long ix = 1, iy = r->mx, iz = (long)r->mx * r->my, lot = (long)r->lot;
    const int mx = r->mx, my = r->my, mz = r->mz;
    const float c = r->P.c;
    void (*f)(float *, float *, float *, const float *, const float *, const float *, long, long, long, int, int, int, long, float) =
        (which & 1) ? postedge : preledge;
    if (which < 2) {
        f(by, bz, bx, ey, ez, ex, iy, iz, ix, my, mz, mx, 1, c);
        f(bz, bx, by, ez, ex, ey, iz, ix, iy, mz, mx, my, 1, c);
        f(bx, by, bz, ex, ey, ez, ix, iy, iz, mx, my, mz, 1, c);
    } else {
        f(ey, ez, ex, by, bz, bx, -iy, -iz, -ix, my, mz, mx, lot, c);
        f(ez, ex, ey, bz, bx, by, -iz, -ix, -iy, mz, mx, my, lot, c);
        f(ex, ey, ez, bx, by, bz, -ix, -iy, -iz, mx, my, mz, lot, c);
    }
}
static int all_open(const orc_world *w) { return w->P.dim == 3 && !w->P.periodicx && !w->P.periodicy && !w->P.periodicz; }

/* ------------------------------------------------------------------------- */
/* exchange_current: fieldboundaries.F90:1768-2189                              */
/* high ghosts m-g..m (g+1 layers) are ADDED into the +neighbour's g+1..nghost; */
/* low ghosts 1..g are added into the -neighbour's m-nghost+1..m-g-1.           */
/* Order x, y, z; full extent of the other axes; each component in turn with    */
/* both receive buffers captured before the adds of that direction pair.        */
/* ------------------------------------------------------------------------- */
void orc_exchange_current(orc_world *w)
{
    int n = w->size0;
    int naxes = w->P.dim == 3 ? 3 : 2;
    float **sbuf = (float **)calloc(n, sizeof(float *));
    for (int axis = 0; axis < naxes; axis++) {
        orc_rank *r0 = w->r[0];
        int per = axis_per(r0, axis), sz = axis_size(r0, axis);
        if (sz == 1 && axis != 2 && !per) continue;          /* :1794, :1929 : nothing on open single-rank axes */
        int g = axis_g(r0, axis), ng = axis_ng(r0, axis);
        if (sz == 1 && axis != 2) {
            /* local periodic fold, :1796-1813 / :1931-1948: both buffers captured first */
            for (int rk = 0; rk < n; rk++) {
                orc_rank *r = w->r[rk]; int m = axis_m(r, axis);
                for (int c = 0; c < 3; c++) {
                    int lo[3], hi[3];
                    full_box(r, lo, hi); lo[axis] = m - g; hi[axis] = m;
                    float *b1 = (float *)malloc(box_count(lo, hi) * sizeof(float));
                    orc_box_get(r, ORC_CURX + c, lo, hi, b1);
                    full_box(r, lo, hi); lo[axis] = 1; hi[axis] = g;
                    float *b2 = (float *)malloc(box_count(lo, hi) * sizeof(float));
                    orc_box_get(r, ORC_CURX + c, lo, hi, b2);
                    full_box(r, lo, hi); lo[axis] = g + 1; hi[axis] = ng;
                    orc_box_add(r, ORC_CURX + c, lo, hi, b1);
                    full_box(r, lo, hi); lo[axis] = m - (ng - 1); hi[axis] = m - (g + 1);
                    orc_box_add(r, ORC_CURX + c, lo, hi, b2);
                    free(b1); free(b2);
                }
            }
            continue;
        }
        /* MPI path, e.g. :1990-2079 (y).  Component order in the reference is z,x,y for x-axis,
           x,y,z for y; components are independent so order is immaterial. */
        for (int c = 0; c < 3; c++) {
            for (int pass = 0; pass < 2; pass++) {
                for (int rk = 0; rk < n; rk++) {
                    orc_rank *r = w->r[rk]; int m = axis_m(r, axis); int lo[3], hi[3]; full_box(r, lo, hi);
                    if (pass == 0) { lo[axis] = m - g; hi[axis] = m; } else { lo[axis] = 1; hi[axis] = g; }
                    sbuf[rk] = (float *)malloc(box_count(lo, hi) * sizeof(float));
                    orc_box_get(r, ORC_CURX + c, lo, hi, sbuf[rk]);
                }
                for (int rk = 0; rk < n; rk++) {
                    orc_rank *r = w->r[rk]; int m = axis_m(r, axis); int lo[3], hi[3]; full_box(r, lo, hi);
                    int pos = axis_pos(r, axis);
                    int src = orc_neighbour(r, pass == 0 ? 2 * axis : 2 * axis + 1);
                    int iper = 1;
                    if (pass == 0 && pos == 0 && !per) iper = 0;
                    if (pass == 1 && pos == sz - 1 && !per) iper = 0;
                    if (iper) {
                        if (pass == 0) { lo[axis] = g + 1; hi[axis] = ng; }
                        else { lo[axis] = m - (ng - 1); hi[axis] = m - (g + 1); }
                        orc_box_add(r, ORC_CURX + c, lo, hi, sbuf[src]);
                    }
                }
                for (int rk = 0; rk < n; rk++) free(sbuf[rk]);
            }
        }
    }
    free(sbuf);
}

/* ------------------------------------------------------------------------- */
/* filter1: filter.F90:8-221                                                    */
/* ------------------------------------------------------------------------- */
void orc_filter1_pass(orc_rank *r)
{
    const int g = r->nghost / 2, gz = r->nghostz / 2;
    const int istr = g + 1, ifin = r->mx - (g + 1);
    const int three = r->P.dim == 3;
    const int k1 = three ? gz + 1 : 1, k2 = three ? r->mz - (gz + 1) : 1;
    /* weights filter.F90:33-64 */
    const float winv = three ? 1.f / 64.f : 1.f / 16.f;
    const float w1 = (three ? 4.f : 2.f) * winv;   /* wtl,wtr,wtu,wtb */
    const float w0 = (three ? 8.f : 4.f) * winv;   /* wt */
    const float w2 = (three ? 2.f : 1.f) * winv;   /* wtlt.. */
    const float wz1 = 2.f * winv, wz0 = 4.f * winv, wz2 = 1.f * winv;
    float *temp = r->temp;
    memset(temp, 0, r->lot * sizeof(float));                /* filter.F90:66 */
    for (int c = 0; c < 3; c++) {
        float *cu = r->f[ORC_CURX + c];
        int j1 = g + 1, j2 = r->my - (g + 1);
        if (c == 2 && (r->P.quirks & ORC_Q5_FILTER_CURZ_J)) j2 = r->my - g + 1;   /* filter.F90:186,211 */
        for (int k = k1; k <= k2; k++) for (int j = j1; j <= j2; j++) for (int i = istr; i <= ifin; i++) {
#define C(di, dj, dk) cu[IDX(r, i + (di), j + (dj), k + (dk))]
            /* filter.F90:105-108 : left-to-right sum */
            float t = w1 * C(-1, 0, 0) + w0 * C(0, 0, 0) + w1 * C(1, 0, 0) +
                      w1 * C(0, -1, 0) + w1 * C(0, 1, 0) +
                      w2 * C(-1, 1, 0) + w2 * C(1, 1, 0) +
                      w2 * C(-1, -1, 0) + w2 * C(1, -1, 0);
            if (three) {
                /* filter.F90:110-117 */
                t = t + wz1 * C(-1, 0, -1) + wz0 * C(0, 0, -1) + wz1 * C(1, 0, -1) +
                    wz1 * C(0, -1, -1) + wz1 * C(0, 1, -1) +
                    wz2 * C(-1, 1, -1) + wz2 * C(1, 1, -1) +
                    wz2 * C(-1, -1, -1) + wz2 * C(1, -1, -1) +
                    wz1 * C(-1, 0, 1) + wz0 * C(0, 0, 1) + wz1 * C(1, 0, 1) +
                    wz1 * C(0, -1, 1) + wz1 * C(0, 1, 1) +
                    wz2 * C(-1, 1, 1) + wz2 * C(1, 1, 1) +
                    wz2 * C(-1, -1, 1) + wz2 * C(1, -1, 1);
            }
#undef C
            temp[IDX(r, i, j, k)] = t;
        }
        for (int k = k1; k <= k2; k++) for (int j = j1; j <= j2; j++) for (int i = istr; i <= ifin; i++)
            cu[IDX(r, i, j, k)] = temp[IDX(r, i, j, k)];
    }
}

/* one-layer refresh of cur before each pass: filter.F90:71-99 : (lt,ls,nt,ns)=(g, m-g-1, m-g, g+1) */
static void filter1_refresh(orc_world *w)
{
    int n = w->size0;
    int naxes = w->P.dim == 3 ? 3 : 2;
    float **sbuf = (float **)calloc(n, sizeof(float *));
    for (int axis = 0; axis < naxes; axis++) {
        for (int pass = 0; pass < 2; pass++) {
            for (int rk = 0; rk < n; rk++) {
                orc_rank *r = w->r[rk]; int lo[3], hi[3]; full_box(r, lo, hi);
                int m = axis_m(r, axis), g = axis_g(r, axis);
                lo[axis] = hi[axis] = pass == 0 ? m - (g + 1) : g + 1;
                size_t cnt = box_count(lo, hi);
                sbuf[rk] = (float *)malloc(3 * cnt * sizeof(float));
                for (int c = 0; c < 3; c++) orc_box_get(r, ORC_CURX + c, lo, hi, sbuf[rk] + c * cnt);
            }
            for (int rk = 0; rk < n; rk++) {
                orc_rank *r = w->r[rk]; int lo[3], hi[3]; full_box(r, lo, hi);
                int m = axis_m(r, axis), g = axis_g(r, axis);
                int src = orc_neighbour(r, pass == 0 ? 2 * axis : 2 * axis + 1);
                int per = axis_per(r, axis), pos = axis_pos(r, axis), sz = axis_size(r, axis);
                int skip = 0;
                if (!per) { if (pass == 0 && pos == 0) skip = 1; if (pass == 1 && pos == sz - 1) skip = 1; }
                if (!skip) {
                    lo[axis] = hi[axis] = pass == 0 ? g : m - g;
                    size_t cnt = box_count(lo, hi);
                    for (int c = 0; c < 3; c++) orc_box_put(r, ORC_CURX + c, lo, hi, sbuf[src] + c * cnt);
                }
            }
            for (int rk = 0; rk < n; rk++) free(sbuf[rk]);
        }
    }
    free(sbuf);
}

void orc_apply_filter1(orc_world *w)
{
    for (int n = 1; n <= w->P.ntimes; n++) {
        filter1_refresh(w);
#pragma omp parallel for schedule(static)
        for (int rk = 0; rk < w->size0; rk++) orc_filter1_pass(w->r[rk]);
    }
}

/* ------------------------------------------------------------------------- */
/* filter2: optimized_filters.F90:9-227 (driver), :459-558 (filter_x; y,z same) */
/* The two-register in-place sweep of filter_x is, per pass,                    */
/*   new(i) = wtm1*old(i-1) + wt*old(i) + wtp1*old(i+1)  (left-to-right sum)    */
/* on the extended line [ghost(1:n) | cur(istr:ifin) | ghost(n+1:2n)] with the  */
/* two end points held fixed.  Restated with an explicit old/new pair.          */
/* ------------------------------------------------------------------------- */
void orc_filter2_line(float *line, int len, int ntimes)
{
    float *tmp = (float *)malloc((size_t)len * sizeof(float));
    for (int n = 0; n < ntimes; n++) {
        tmp[0] = line[0]; tmp[len - 1] = line[len - 1];
        for (int i = 1; i < len - 1; i++)
            tmp[i] = .25f * line[i - 1] + .5f * line[i] + .25f * line[i + 1];
        memcpy(line, tmp, (size_t)len * sizeof(float));
    }
    free(tmp);
}

/* deep_copy_layr{x,y,z}{1,2}: optimized_filters.F90:1387-1963.
   ghost(1:n) <- (-neighbour) cur(fin-n+1:fin); ghost(n+1:2n) <- (+neighbour) cur(str:str+n-1),
   rows restricted to the interior of the other axes; open edge replicates the edge value. */
/* per-rank part: filter every line of component c along `axis`, given the two nt-deep halo slabs in
   orc_box_get order over the restricted box (x fastest).  Open edges replicate cur(str) / cur(fin). */
void orc_filter2_rank(orc_rank *r, int c, int axis, const float *glo, const float *ghi)
{
    const int nt = r->P.ntimes;
    int lo[3], hi[3];
    for (int a = 0; a < 3; a++) { lo[a] = axis_g(r, a) + 1; hi[a] = axis_m(r, a) - (axis_g(r, a) + 1); }
    if (r->P.dim == 2) { lo[2] = hi[2] = 1; }
    int per = axis_per(r, axis), pos = axis_pos(r, axis), sz = axis_size(r, axis);
    int str = lo[axis], fin = hi[axis], ncell = fin - str + 1, len = ncell + 2 * nt;
    int a1 = (axis + 1) % 3, a2 = (axis + 2) % 3;
    if (a1 > a2) { int t = a1; a1 = a2; a2 = t; }           /* a1 < a2: a1 is the faster of the two */
    int n1 = hi[a1] - lo[a1] + 1, n2 = hi[a2] - lo[a2] + 1;
    float *line = (float *)malloc((size_t)len * sizeof(float));
    float *cu = r->f[ORC_CURX + c];
    int lowrep = !(per || pos != 0), highrep = !(per || pos != sz - 1);
    for (int q2 = 0; q2 < n2; q2++) for (int q1 = 0; q1 < n1; q1++) {
        int ijk[3]; ijk[a1] = lo[a1] + q1; ijk[a2] = lo[a2] + q2;
        for (int t = 0; t < nt; t++) {
            size_t off;
            if (axis == 0) off = (size_t)t + (size_t)nt * ((size_t)q1 + (size_t)n1 * q2);
            else if (axis == 1) off = (size_t)q1 + (size_t)n1 * ((size_t)t + (size_t)nt * q2);
            else off = (size_t)q1 + (size_t)n1 * ((size_t)q2 + (size_t)n2 * t);
            ijk[axis] = str; float elo = cu[IDX(r, ijk[0], ijk[1], ijk[2])];
            ijk[axis] = fin; float ehi = cu[IDX(r, ijk[0], ijk[1], ijk[2])];
            line[t] = lowrep ? elo : glo[off];
            line[nt + ncell + t] = highrep ? ehi : ghi[off];
        }
        for (int s = 0; s < ncell; s++) { ijk[axis] = str + s; line[nt + s] = cu[IDX(r, ijk[0], ijk[1], ijk[2])]; }
        orc_filter2_line(line, len, nt);
        for (int s = 0; s < ncell; s++) { ijk[axis] = str + s; cu[IDX(r, ijk[0], ijk[1], ijk[2])] = line[nt + s]; }
    }
    free(line);
}

/* boxes a rank contributes to / expects from its axis neighbours: what = 0 -> my last nt cells (sent up, becomes the
   + neighbour's low halo), what = 1 -> my first nt cells (sent down) */
void orc_filter2_send_box(const orc_rank *r, int axis, int what, int lo[3], int hi[3])
{
    for (int a = 0; a < 3; a++) { lo[a] = axis_g(r, a) + 1; hi[a] = axis_m(r, a) - (axis_g(r, a) + 1); }
    if (r->P.dim == 2) { lo[2] = hi[2] = 1; }
    int nt = r->P.ntimes, str = lo[axis], fin = hi[axis];
    if (what == 0) { lo[axis] = fin - nt + 1; hi[axis] = fin; } else { lo[axis] = str; hi[axis] = str + nt - 1; }
}

static void filter2_axis(orc_world *w, int c, int axis)
{
    int nr = w->size0;
    float **glo = (float **)calloc(nr, sizeof(float *)), **ghi = (float **)calloc(nr, sizeof(float *));
    for (int rk = 0; rk < nr; rk++) {
        orc_rank *r = w->r[rk];
        orc_rank *rm = w->r[orc_neighbour(r, 2 * axis)], *rp = w->r[orc_neighbour(r, 2 * axis + 1)];
        int lo[3], hi[3];
        orc_filter2_send_box(rm, axis, 0, lo, hi);
        glo[rk] = (float *)malloc(box_count(lo, hi) * sizeof(float));
        orc_box_get(rm, ORC_CURX + c, lo, hi, glo[rk]);
        orc_filter2_send_box(rp, axis, 1, lo, hi);
        ghi[rk] = (float *)malloc(box_count(lo, hi) * sizeof(float));
        orc_box_get(rp, ORC_CURX + c, lo, hi, ghi[rk]);
    }
#pragma omp parallel for schedule(static)
    for (int rk = 0; rk < nr; rk++) orc_filter2_rank(w->r[rk], c, axis, glo[rk], ghi[rk]);
    for (int rk = 0; rk < nr; rk++) { free(glo[rk]); free(ghi[rk]); }
    free(glo); free(ghi);
}

void orc_apply_filter2(orc_world *w)
{
    if (w->P.ntimes <= 0) return;
    int naxes = w->P.dim == 3 ? 3 : 2;
    for (int c = 0; c < 3; c++)                              /* optimized_filters.F90:49-222: curx, cury, curz */
        for (int axis = 0; axis < naxes; axis++) filter2_axis(w, c, axis);
}

/* filter2 needs an ntimes-deep halo from the axis neighbour, i.e. ntimes <= that rank's interior extent on every
   filtered axis.  The reference guards this (loosely) in tristanmainloop.F90:217-224 and otherwise runs filter1. */
int orc_filter2_fits(const orc_world *w)
{
    int naxes = w->P.dim == 3 ? 3 : 2;
    for (int rk = 0; rk < w->size0; rk++)
        for (int a = 0; a < naxes; a++)
            if (w->P.ntimes > axis_m(w->r[rk], a) - 2 * axis_g(w->r[rk], a) - 1) return 0;
    return 1;
}

void orc_apply_filter(orc_world *w)
{
    /* tristanmainloop.F90:213-229 */
    if (w->P.filter_kind == 2 && orc_filter2_fits(w)) orc_apply_filter2(w); else orc_apply_filter1(w);
}

/* ------------------------------------------------------------------------- */
/* shape weights: Appendix A.1 of SURVEY.md.                                    */
/* S is indexed 0..7; slots 1..6 are the reference's Sx(1:6); slot 3 <-> cell ip*/
/* particles_movedeposit.F90:447-467 (o1), :709-790 (o2), :1035-1128 (o3);      */
/* particles.F90:738-769, 928-982, 1175-1260 (same, with `shift`).              */
/* ------------------------------------------------------------------------- */
void orc_shape(int order, float d, int shift, float S[8], int *smin, int *smax)
{
    /* particles.F90:239-251 : fp32 module constants */
    const float half = 1.f / 2.f, quart = 1.f / 4.f, one = 1.f, two = 2.f, thhalf = 3.f / 2.f,
                nineighth = 9.f / 8.f, twoth = 2.f / 3.f, sixth = 1.f / 6.f, negsixth = -1.f / 6.f,
                negone = -1.f;
    for (int i = 0; i < 8; i++) S[i] = 0.f;
    if (order <= 1) {
        S[3 + shift] = 1.f - d; S[4 + shift] = d;
        *smin = 3 + shift; *smax = 4 + shift;
    } else if (order == 2) {
        if (d <= half) {
            S[2 + shift] = half * (d * d - d + quart);
            S[4 + shift] = S[2 + shift] + d;
            S[3 + shift] = one - S[4 + shift] - S[2 + shift];
            *smin = 2 + shift; *smax = 4 + shift;
        } else {
            S[3 + shift] = nineighth - thhalf * d + half * d * d;
            S[5 + shift] = S[3 + shift] - one + d;
            S[4 + shift] = one - S[5 + shift] - S[3 + shift];
            *smin = 3 + shift; *smax = 5 + shift;
        }
    } else {
        if (d <= half) {
            S[2 + shift] = negsixth * (d - one) * (d - one) * (d - one);
            S[3 + shift] = twoth + half * (d - two) * d * d;
            S[5 + shift] = sixth * d * d * d;
            S[4 + shift] = one - S[5 + shift] - S[3 + shift] - S[2 + shift];
        } else {
            S[5 + shift] = sixth * d * d * d;
            S[4 + shift] = twoth + half * (negone - d) * (one - d) * (one - d);
            S[2 + shift] = sixth * (one - d) * (one - d) * (one - d);
            S[3 + shift] = one - S[5 + shift] - S[4 + shift] - S[2 + shift];
        }
        *smin = 2 + shift; *smax = 5 + shift;
    }
}

/* ------------------------------------------------------------------------- */
/* Boris / Vay push, shared tail of all movers: particles_movedeposit.F90:834-929*/
/* ------------------------------------------------------------------------- */
static void push(const orc_rank *r, orc_particle *p, float ex0, float ey0, float ez0,
                 float bx0, float by0, float bz0)
{
    const float c = r->P.c, cinv = 1.f / c;
    float u0, v0, w0, u1, v1, w1, g, f;
    if (r->P.pusher == 1) {
        /* Vay 2008, :861-885 */
        g = 1.f / sqrtf(1.f + p->u * p->u + p->v * p->v + p->w * p->w);
        float vx0 = c * p->u * g, vy0 = c * p->v * g, vz0 = c * p->w * g;
        u1 = c * p->u + 2.f * ex0 + vy0 * bz0 - vz0 * by0;
        v1 = c * p->v + 2.f * ey0 + vz0 * bx0 - vx0 * bz0;
        w1 = c * p->w + 2.f * ez0 + vx0 * by0 - vy0 * bx0;
        float ustar = cinv * (u1 * bx0 + v1 * by0 + w1 * bz0);
        float sig = cinv * cinv * (c * c + u1 * u1 + v1 * v1 + w1 * w1) - (bx0 * bx0 + by0 * by0 + bz0 * bz0);
        g = 1.f / sqrtf(0.5f * (sig + sqrtf(sig * sig + 4.f * (bx0 * bx0 + by0 * by0 + bz0 * bz0 + ustar * ustar))));
        float tx = bx0 * g, ty = by0 * g, tz = bz0 * g;
        f = 1.f / (1.f + tx * tx + ty * ty + tz * tz);
        u0 = f * (u1 + (u1 * tx + v1 * ty + w1 * tz) * tx + v1 * tz - w1 * ty);
        v0 = f * (v1 + (u1 * tx + v1 * ty + w1 * tz) * ty + w1 * tx - u1 * tz);
        w0 = f * (w1 + (u1 * tx + v1 * ty + w1 * tz) * tz + u1 * ty - v1 * tx);
    } else {
        /* Boris, :889-905 */
        u0 = c * p->u + ex0; v0 = c * p->v + ey0; w0 = c * p->w + ez0;
        g = c / sqrtf(c * c + u0 * u0 + v0 * v0 + w0 * w0);
        bx0 = g * bx0; by0 = g * by0; bz0 = g * bz0;
        f = 2.f / (1.f + bx0 * bx0 + by0 * by0 + bz0 * bz0);
        u1 = (u0 + v0 * bz0 - w0 * by0) * f;
        v1 = (v0 + w0 * bx0 - u0 * bz0) * f;
        w1 = (w0 + u0 * by0 - v0 * bx0) * f;
        u0 = u0 + v1 * bz0 - w1 * by0 + ex0;
        v0 = v0 + w1 * bx0 - u1 * bz0 + ey0;
        w0 = w0 + u1 * by0 - v1 * bx0 + ez0;
    }
    /* :913-929 */
    p->u = u0 * cinv; p->v = v0 * cinv; p->w = w0 * cinv;
    g = c / sqrtf(c * c + u0 * u0 + v0 * v0 + w0 * w0);
    p->x = p->x + p->u * g * c;
    p->y = p->y + p->v * g * c;
    p->z = p->z + p->w * g * c;
}

/* mover (zigzag build): particles_movedeposit.F90:98-347, trilinear staggered gather :172-248 */
static void mover_zigzag(orc_rank *r, int n1, int n2, float qm)
{
    const float cinv = 1.f / r->P.c;
    const float *ex = r->f[ORC_EX], *ey = r->f[ORC_EY], *ez = r->f[ORC_EZ];
    const float *bx = r->f[ORC_BX], *by = r->f[ORC_BY], *bz = r->f[ORC_BZ];
    const long ix = 1, iy = r->iy, iz = r->iz;
    for (int n = n1; n <= n2; n++) {
        orc_particle *p = &r->p[n - 1];
        float qm1 = qm;
        if (p->ind < 0 && qm > 0) qm1 = fabsf(r->P.qme);       /* :125-127 positron hack */
        int i = (int)p->x; float dx = p->x - i;
        int j = (int)p->y; float dy = p->y - j;
        int k = (int)p->z; float dz = p->z - k;
        if (r->P.dim == 2) { k = 1; dz = 0; }
        long l = (i - 1) + iy * (j - 1) + iz * (k - 1);         /* 0-based linear index of (i,j,k) */
        float f, g, ex0, ey0, ez0, bx0, by0, bz0;
        f = ex[l] + ex[l - ix] + dx * (ex[l + ix] - ex[l - ix]);
        f = f + dy * (ex[l + iy] + ex[l - ix + iy] + dx * (ex[l + ix + iy] - ex[l - ix + iy]) - f);
        g = ex[l + iz] + ex[l - ix + iz] + dx * (ex[l + ix + iz] - ex[l - ix + iz]);
        g = g + dy * (ex[l + iy + iz] + ex[l - ix + iy + iz] + dx * (ex[l + ix + iy + iz] - ex[l - ix + iy + iz]) - g);
        ex0 = (f + dz * (g - f)) * (.25f * qm1);
        f = ey[l] + ey[l - iy] + dy * (ey[l + iy] - ey[l - iy]);
        f = f + dz * (ey[l + iz] + ey[l - iy + iz] + dy * (ey[l + iy + iz] - ey[l - iy + iz]) - f);
        g = ey[l + ix] + ey[l - iy + ix] + dy * (ey[l + iy + ix] - ey[l - iy + ix]);
        g = g + dz * (ey[l + iz + ix] + ey[l - iy + iz + ix] + dy * (ey[l + iy + iz + ix] - ey[l - iy + iz + ix]) - g);
        ey0 = (f + dx * (g - f)) * (.25f * qm1);
        f = ez[l] + ez[l - iz] + dz * (ez[l + iz] - ez[l - iz]);
        f = f + dx * (ez[l + ix] + ez[l - iz + ix] + dz * (ez[l + iz + ix] - ez[l - iz + ix]) - f);
        g = ez[l + iy] + ez[l - iz + iy] + dz * (ez[l + iz + iy] - ez[l - iz + iy]);
        g = g + dx * (ez[l + ix + iy] + ez[l - iz + ix + iy] + dz * (ez[l + iz + ix + iy] - ez[l - iz + ix + iy]) - g);
        ez0 = (f + dy * (g - f)) * (.25f * qm1);
        f = bx[l - iy] + bx[l - iy - iz] + dz * (bx[l - iy + iz] - bx[l - iy - iz]);
        f = bx[l] + bx[l - iz] + dz * (bx[l + iz] - bx[l - iz]) + f +
            dy * (bx[l + iy] + bx[l + iy - iz] + dz * (bx[l + iy + iz] - bx[l + iy - iz]) - f);
        g = bx[l + ix - iy] + bx[l + ix - iy - iz] + dz * (bx[l + ix - iy + iz] - bx[l + ix - iy - iz]);
        g = bx[l + ix] + bx[l + ix - iz] + dz * (bx[l + ix + iz] - bx[l + ix - iz]) + g +
            dy * (bx[l + ix + iy] + bx[l + ix + iy - iz] + dz * (bx[l + ix + iy + iz] - bx[l + ix + iy - iz]) - g);
        bx0 = (f + dx * (g - f)) * (.125f * qm1 * cinv);
        f = by[l - iz] + by[l - iz - ix] + dx * (by[l - iz + ix] - by[l - iz - ix]);
        f = by[l] + by[l - ix] + dx * (by[l + ix] - by[l - ix]) + f +
            dz * (by[l + iz] + by[l + iz - ix] + dx * (by[l + iz + ix] - by[l + iz - ix]) - f);
        g = by[l + iy - iz] + by[l + iy - iz - ix] + dx * (by[l + iy - iz + ix] - by[l + iy - iz - ix]);
        g = by[l + iy] + by[l + iy - ix] + dx * (by[l + iy + ix] - by[l + iy - ix]) + g +
            dz * (by[l + iy + iz] + by[l + iy + iz - ix] + dx * (by[l + iy + iz + ix] - by[l + iy + iz - ix]) - g);
        by0 = (f + dy * (g - f)) * (.125f * qm1 * cinv);
        f = bz[l - ix] + bz[l - ix - iy] + dy * (bz[l - ix + iy] - bz[l - ix - iy]);
        f = bz[l] + bz[l - iy] + dy * (bz[l + iy] - bz[l - iy]) + f +
            dx * (bz[l + ix] + bz[l + ix - iy] + dy * (bz[l + ix + iy] - bz[l + ix - iy]) - f);
        g = bz[l + iz - ix] + bz[l + iz - ix - iy] + dy * (bz[l + iz - ix + iy] - bz[l + iz - ix - iy]);
        g = bz[l + iz] + bz[l + iz - iy] + dy * (bz[l + iz + iy] - bz[l + iz - iy]) + g +
            dx * (bz[l + iz + ix] + bz[l + iz + ix - iy] + dy * (bz[l + iz + ix + iy] - bz[l + iz + ix - iy]) - g);
        bz0 = (f + dz * (g - f)) * (.125f * qm1 * cinv);
        if (r->P.external_fields) {                              /* :250-262 */
            bx0 = bx0 + r->P.ext[3] * 0.5f * qm1 * cinv;
            by0 = by0 + r->P.ext[4] * 0.5f * qm1 * cinv;
            bz0 = bz0 + r->P.ext[5] * 0.5f * qm1 * cinv;
            ex0 = ex0 + r->P.ext[0] * 0.5f * qm1;
            ey0 = ey0 + r->P.ext[1] * 0.5f * qm1;
            ez0 = ez0 + r->P.ext[2] * 0.5f * qm1;
        }
        push(r, p, ex0, ey0, ez0, bx0, by0, bz0);
    }
}

/* node-centred ("primal") fields, 3D only: particles_movedeposit.F90:395-404, 658-667, 982-991.
   cshift is circular. out[6] are mx*my*mz arrays. */
static void primal_grids(const orc_rank *r, float *out[6])
{
    const float *ex = r->f[ORC_EX], *ey = r->f[ORC_EY], *ez = r->f[ORC_EZ];
    const float *bx = r->f[ORC_BX], *by = r->f[ORC_BY], *bz = r->f[ORC_BZ];
    int q2 = (r->P.quirks & ORC_Q2_BXBY_NO_KAVG) != 0;
    for (int k = 1; k <= r->mz; k++) for (int j = 1; j <= r->my; j++) for (int i = 1; i <= r->mx; i++) {
        int im = i == 1 ? r->mx : i - 1, jm = j == 1 ? r->my : j - 1, km = k == 1 ? r->mz : k - 1;
        size_t l = IDX(r, i, j, k);
        out[0][l] = 0.5f * (ex[l] + ex[IDX(r, im, j, k)]);
        out[1][l] = 0.5f * (ey[l] + ey[IDX(r, i, jm, k)]);
        out[2][l] = 0.5f * (ez[l] + ez[IDX(r, i, j, km)]);
        float bxp = 0.5f * (bx[l] + bx[IDX(r, i, jm, k)]);
        float byp = 0.5f * (by[l] + by[IDX(r, im, j, k)]);
        if (!q2) {
            /* "fixed": complete the node centring in k as well */
            float bxk = 0.5f * (bx[IDX(r, i, j, km)] + bx[IDX(r, i, jm, km)]);
            float byk = 0.5f * (by[IDX(r, i, j, km)] + by[IDX(r, im, j, km)]);
            bxp = 0.5f * (bxp + bxk); byp = 0.5f * (byp + byk);
        }
        out[3][l] = bxp; out[4][l] = byp;
        out[5][l] = 0.5f * (0.5f * (bz[l] + bz[IDX(r, im, j, k)]) +
                            0.5f * (bz[IDX(r, i, jm, k)] + bz[IDX(r, im, jm, k)]));
    }
}

/* mover_1ord/2ord/3ord: particles_movedeposit.F90:356-610, 619-933, 943-1271 */
static void mover_shaped(orc_rank *r, int n1, int n2, float qm)
{
    const int order = r->P.order, three = r->P.dim == 3;
    const float cinv = 1.f / r->P.c, half = 0.5f;
    float *prim[6] = {0};
    if (three) { for (int a = 0; a < 6; a++) prim[a] = (float *)malloc(r->lot * sizeof(float)); primal_grids(r, prim); }
    const int q1 = order == 2 && (r->P.quirks & ORC_Q1_MOVER2_RANGE);
    for (int n = n1; n <= n2; n++) {
        orc_particle *p = &r->p[n - 1];
        float Sxp[8], Syp[8], Szp[8], Sxd[8], Syd[8], Szd[8];
        int pmin[3], pmax[3], dmin[3], dmax[3];
        int ip = (int)p->x; float dxp = p->x - ip;
        int jp = (int)p->y; float dyp = p->y - jp;
        int kp = (int)p->z; float dzp = p->z - kp;
        int id = (int)(p->x - half); float dxd = p->x - half - id;
        int jd = (int)(p->y - half); float dyd = p->y - half - jd;
        int kd = (int)(p->z - half); float dzd = (p->z - half) - kd;
        orc_shape(order, dxp, 0, Sxp, &pmin[0], &pmax[0]);
        orc_shape(order, dyp, 0, Syp, &pmin[1], &pmax[1]);
        orc_shape(order, dxd, 0, Sxd, &dmin[0], &dmax[0]);
        orc_shape(order, dyd, 0, Syd, &dmin[1], &dmax[1]);
        if (three) {
            orc_shape(order, dzp, 0, Szp, &pmin[2], &pmax[2]);
            orc_shape(order, dzd, 0, Szd, &dmin[2], &dmax[2]);
        }
        (void)kd; (void)Szd;
        int imin[3], imax[3];
        for (int a = 0; a < 3; a++) {
            if (q1) { imin[a] = dmin[a]; imax[a] = dmax[a]; }   /* Q1: dual branch overwrote the bounds */
            else if (order == 2 && !three) { imin[a] = 2; imax[a] = 5; } /* fixed 2D: cover both supports */
            else { imin[a] = pmin[a]; imax[a] = pmax[a]; }
        }
        float ex0 = 0, ey0 = 0, ez0 = 0, bx0 = 0, by0 = 0, bz0 = 0;
        if (three) {
            /* :801-815 : sum() over the x slice, then *Syp*Szp */
            for (int i3 = imin[2]; i3 <= imax[2]; i3++) for (int i2 = imin[1]; i2 <= imax[1]; i2++) {
                float s[6] = {0, 0, 0, 0, 0, 0};
                for (int i1 = imin[0]; i1 <= imax[0]; i1++) {
                    size_t l = IDX(r, ip - 3 + i1, jp - 3 + i2, kp - 3 + i3);
                    for (int a = 0; a < 6; a++) s[a] = s[a] + prim[a][l] * Sxp[i1];
                }
                ex0 = ex0 + s[0] * Syp[i2] * Szp[i3];
                ey0 = ey0 + s[1] * Syp[i2] * Szp[i3];
                ez0 = ez0 + s[2] * Syp[i2] * Szp[i3];
                bx0 = bx0 + s[3] * Syp[i2] * Szp[i3];
                by0 = by0 + s[4] * Syp[i2] * Szp[i3];
                bz0 = bz0 + s[5] * Syp[i2] * Szp[i3];
            }
        } else {
            /* :817-831 : raw Yee arrays, mixed primal/dual weights */
            const float *ex = r->f[ORC_EX], *ey = r->f[ORC_EY], *ez = r->f[ORC_EZ];
            const float *bx = r->f[ORC_BX], *by = r->f[ORC_BY], *bz = r->f[ORC_BZ];
            const long iy = r->iy;
            for (int i2 = imin[1]; i2 <= imax[1]; i2++) for (int i1 = imin[0]; i1 <= imax[0]; i1++) {
                long lpp = (ip - 3 + i1) + iy * (jp - 3 + i2 - 1) - 1;
                long lpd = (ip - 3 + i1) + iy * (jd - 3 + i2 - 1) - 1;
                long ldp = (id - 3 + i1) + iy * (jp - 3 + i2 - 1) - 1;
                long ldd = (id - 3 + i1) + iy * (jd - 3 + i2 - 1) - 1;
                ex0 = ex0 + ex[ldp] * Sxd[i1] * Syp[i2];
                ey0 = ey0 + ey[lpd] * Sxp[i1] * Syd[i2];
                ez0 = ez0 + ez[lpp] * Sxp[i1] * Syp[i2];
                bx0 = bx0 + bx[lpd] * Sxp[i1] * Syd[i2];
                by0 = by0 + by[ldp] * Sxd[i1] * Syp[i2];
                bz0 = bz0 + bz[ldd] * Sxd[i1] * Syd[i2];
            }
        }
        /* :834-839 */
        ex0 = 0.5f * ex0 * qm; ey0 = 0.5f * ey0 * qm; ez0 = 0.5f * ez0 * qm;
        bx0 = 0.5f * bx0 * qm * cinv; by0 = 0.5f * by0 * qm * cinv; bz0 = 0.5f * bz0 * qm * cinv;
        if (r->P.external_fields) {                              /* :841-853 */
            bx0 = bx0 + r->P.ext[3] * 0.5f * qm * cinv;
            by0 = by0 + r->P.ext[4] * 0.5f * qm * cinv;
            bz0 = bz0 + r->P.ext[5] * 0.5f * qm * cinv;
            ex0 = ex0 + r->P.ext[0] * 0.5f * qm;
            ey0 = ey0 + r->P.ext[1] * 0.5f * qm;
            ez0 = ez0 + r->P.ext[2] * 0.5f * qm;
        }
        push(r, p, ex0, ey0, ez0, bx0, by0, bz0);
    }
    if (three) for (int a = 0; a < 6; a++) free(prim[a]);
}

void orc_mover_range(orc_rank *r, int n1, int n2, float qm)
{
    if (n2 < n1) return;
    if (r->P.order == 0) mover_zigzag(r, n1, n2, qm); else mover_shaped(r, n1, n2, qm);
}

/* move_particles: particles_movedeposit.F90:63-88 */
void orc_move_particles(orc_rank *r)
{
    orc_mover_range(r, 1, r->ions, r->P.qmi);
    orc_mover_range(r, r->maxhlf + 1, r->maxhlf + r->lecs, r->P.qme);
}

/* ------------------------------------------------------------------------- */
/* deposits: zigzag particles.F90:550-669; densdecomp_{1,2,3}ord :678-1358      */
/* ------------------------------------------------------------------------- */
static void zigzag(orc_rank *r, float x2, float y2, float z2, float x1, float y1, float z1, float q)
{
    float *curx = r->f[ORC_CURX], *cury = r->f[ORC_CURY], *curz = r->f[ORC_CURZ];
    int three = r->P.dim == 3;
    int i1 = (int)x1, i2 = (int)x2, j1 = (int)y1, j2 = (int)y2, k1 = (int)z1, k2 = (int)z2;
#define FMIN(a, b) ((a) < (b) ? (a) : (b))
#define FMAX(a, b) ((a) > (b) ? (a) : (b))
    float xr = FMIN((float)(FMIN(i1, i2) + 1), FMAX((float)FMAX(i1, i2), .5f * (x1 + x2)));
    float yr = FMIN((float)(FMIN(j1, j2) + 1), FMAX((float)FMAX(j1, j2), .5f * (y1 + y2)));
    float zr = FMIN((float)(FMIN(k1, k2) + 1), FMAX((float)FMAX(k1, k2), .5f * (z1 + z2)));
    if (!three) { k1 = 1; k2 = 1; }
    float Fx1 = -q * (xr - x1), Fy1 = -q * (yr - y1), Fz1 = -q * (zr - z1);
    float Wx1 = .5f * (x1 + xr) - i1, Wy1 = .5f * (y1 + yr) - j1, Wz1 = three ? .5f * (z1 + zr) - k1 : 0.f;
    float Wx2 = .5f * (x2 + xr) - i2, Wy2 = .5f * (y2 + yr) - j2, Wz2 = three ? .5f * (z2 + zr) - k2 : 0.f;
    float Fx2 = -q * (x2 - xr), Fy2 = -q * (y2 - yr), Fz2 = -q * (z2 - zr);
#define ADD(arr, i, j, k, v) arr[IDX(r, i, j, k)] = arr[IDX(r, i, j, k)] + (v)
    ADD(curx, i1, j1, k1, Fx1 * (1.f - Wy1) * (1.f - Wz1));
    ADD(curx, i1, j1 + 1, k1, Fx1 * Wy1 * (1.f - Wz1));
    if (three) { ADD(curx, i1, j1, k1 + 1, Fx1 * (1 - Wy1) * Wz1); ADD(curx, i1, j1 + 1, k1 + 1, Fx1 * Wy1 * Wz1); }
    ADD(curx, i2, j2, k2, Fx2 * (1.f - Wy2) * (1.f - Wz2));
    ADD(curx, i2, j2 + 1, k2, Fx2 * Wy2 * (1.f - Wz2));
    if (three) { ADD(curx, i2, j2, k2 + 1, Fx2 * (1.f - Wy2) * Wz2); ADD(curx, i2, j2 + 1, k2 + 1, Fx2 * Wy2 * Wz2); }
    ADD(cury, i1, j1, k1, Fy1 * (1.f - Wx1) * (1.f - Wz1));
    ADD(cury, i1 + 1, j1, k1, Fy1 * Wx1 * (1.f - Wz1));
    if (three) { ADD(cury, i1, j1, k1 + 1, Fy1 * (1.f - Wx1) * Wz1); ADD(cury, i1 + 1, j1, k1 + 1, Fy1 * Wx1 * Wz1); }
    ADD(cury, i2, j2, k2, Fy2 * (1.f - Wx2) * (1.f - Wz2));
    ADD(cury, i2 + 1, j2, k2, Fy2 * Wx2 * (1.f - Wz2));
    if (three) { ADD(cury, i2, j2, k2 + 1, Fy2 * (1.f - Wx2) * Wz2); ADD(cury, i2 + 1, j2, k2 + 1, Fy2 * Wx2 * Wz2); }
    ADD(curz, i1, j1, k1, Fz1 * (1.f - Wx1) * (1.f - Wy1));
    ADD(curz, i1 + 1, j1, k1, Fz1 * Wx1 * (1.f - Wy1));
    ADD(curz, i1, j1 + 1, k1, Fz1 * (1.f - Wx1) * Wy1);
    ADD(curz, i1 + 1, j1 + 1, k1, Fz1 * Wx1 * Wy1);
    ADD(curz, i2, j2, k2, Fz2 * (1.f - Wx2) * (1.f - Wy2));
    ADD(curz, i2 + 1, j2, k2, Fz2 * Wx2 * (1.f - Wy2));
    ADD(curz, i2, j2 + 1, k2, Fz2 * (1.f - Wx2) * Wy2);
    ADD(curz, i2 + 1, j2 + 1, k2, Fz2 * Wx2 * Wy2);
#undef ADD
}

static void densdecomp(orc_rank *r, float x2, float y2, float z2, float x1, float y1, float z1, float q)
{
    const int order = r->P.order, three = r->P.dim == 3;
    const float half = 0.5f, third = 1.f / 3.f;
    float *curx = r->f[ORC_CURX], *cury = r->f[ORC_CURY], *curz = r->f[ORC_CURZ];
    float Sx1[8], Sy1[8], Sz1[8], Sx2[8], Sy2[8], Sz2[8];
    int i1 = (int)x1, j1 = (int)y1, k1 = (int)z1;
    int shifti = (int)x2 - (int)x1, shiftj = (int)y2 - (int)y1, shiftk = (int)z2 - (int)z1;
    float dx1 = x1 - (int)x1, dy1 = y1 - (int)y1, dz1 = z1 - (int)z1;
    float dx2 = x2 - (int)x2, dy2 = y2 - (int)y2, dz2 = z2 - (int)z2;
    float deltaz = z2 - z1;
    int a1, b1, a2, b2, xmin, xmax, ymin, ymax, zmin = 3, zmax = 3;
    orc_shape(order, dx1, 0, Sx1, &a1, &b1); orc_shape(order, dx2, shifti, Sx2, &a2, &b2);
    xmin = a1 < a2 ? a1 : a2; xmax = b1 > b2 ? b1 : b2;
    orc_shape(order, dy1, 0, Sy1, &a1, &b1); orc_shape(order, dy2, shiftj, Sy2, &a2, &b2);
    ymin = a1 < a2 ? a1 : a2; ymax = b1 > b2 ? b1 : b2;
    if (three) {
        orc_shape(order, dz1, 0, Sz1, &a1, &b1); orc_shape(order, dz2, shiftk, Sz2, &a2, &b2);
        zmin = a1 < a2 ? a1 : a2; zmax = b1 > b2 ? b1 : b2;
    } else { k1 = 1; }
    const int carry = (r->P.quirks & ORC_Q4_DEPOSIT_CARRY) != 0;
    float curx_add = 0.f, curx_add_prev = 0.f;              /* Q3: treated as zero-initialised */
    float cury_adds[8] = {0}, cury_add_prevs[8] = {0};
    float curz_adds[8][8], curz_add_prevs[8][8];
    memset(curz_adds, 0, sizeof curz_adds); memset(curz_add_prevs, 0, sizeof curz_add_prevs);
    if (three) {
        /* particles.F90:1024-1067 (2ord; 1ord :776-819 and 3ord :1280-1322 are the same loop) */
        for (int it2 = zmin; it2 <= zmax; it2++) {
            for (int it1 = ymin; it1 <= ymax; it1++) {
                for (int it = xmin; it <= xmax; it++) {
                    size_t l2 = IDX(r, i1 - 3 + it, j1 - 3 + it1, k1 - 3 + it2);
                    curx_add = q * ((Sx2[it] - Sx1[it]) *
                                    (Sy1[it1] * Sz1[it2] + half * (Sy2[it1] - Sy1[it1]) * Sz1[it2] +
                                     half * Sy1[it1] * (Sz2[it2] - Sz1[it2]) +
                                     third * (Sy2[it1] - Sy1[it1]) * (Sz2[it2] - Sz1[it2]))) + curx_add_prev;
                    cury_adds[it] = q * ((Sy2[it1] - Sy1[it1]) *
                                         (Sx1[it] * Sz1[it2] + half * (Sx2[it] - Sx1[it]) * Sz1[it2] +
                                          half * Sx1[it] * (Sz2[it2] - Sz1[it2]) +
                                          third * (Sx2[it] - Sx1[it]) * (Sz2[it2] - Sz1[it2]))) + cury_add_prevs[it];
                    curz_adds[it][it1] = q * ((Sz2[it2] - Sz1[it2]) *
                                              (Sx1[it] * Sy1[it1] + half * (Sx2[it] - Sx1[it]) * Sy1[it1] +
                                               half * Sx1[it] * (Sy2[it1] - Sy1[it1]) +
                                               third * (Sx2[it] - Sx1[it]) * (Sy2[it1] - Sy1[it1]))) + curz_add_prevs[it][it1];
                    curx[l2] = curx[l2] + curx_add;
                    cury[l2] = cury[l2] + cury_adds[it];
                    curz[l2] = curz[l2] + curz_adds[it][it1];
                    curx_add_prev = curx_add;
                }
                curx_add = 0.f;
                if (!carry) curx_add_prev = 0.f;
                memcpy(cury_add_prevs, cury_adds, sizeof cury_adds);
                memset(cury_adds, 0, sizeof cury_adds);
            }
            if (!carry) memset(cury_add_prevs, 0, sizeof cury_add_prevs);
            memcpy(curz_add_prevs, curz_adds, sizeof curz_adds);
            memset(curz_adds, 0, sizeof curz_adds);
        }
    } else {
        /* particles.F90:1070-1097 */
        for (int it1 = ymin; it1 <= ymax; it1++) {
            for (int it = xmin; it <= xmax; it++) {
                size_t l2 = IDX(r, i1 - 3 + it, j1 - 3 + it1, k1);
                curx_add = q * ((Sx2[it] - Sx1[it]) * (Sy1[it1] + half * (Sy2[it1] - Sy1[it1]))) + curx_add_prev;
                cury_adds[it] = q * ((Sy2[it1] - Sy1[it1]) * (Sx1[it] + half * (Sx2[it] - Sx1[it]))) + cury_add_prevs[it];
                float curz_add = -1 * q * deltaz *
                                 (Sx1[it] * Sy1[it1] + half * (Sx2[it] - Sx1[it]) * Sy1[it1] +
                                  half * Sx1[it] * (Sy2[it1] - Sy1[it1]) +
                                  third * (Sx2[it] - Sx1[it]) * (Sy2[it1] - Sy1[it1]));
                curx[l2] = curx[l2] + curx_add;
                cury[l2] = cury[l2] + cury_adds[it];
                curz[l2] = curz[l2] + curz_add;
                curx_add_prev = curx_add;
            }
            curx_add = 0.f;
            if (!carry) curx_add_prev = 0.f;
            memcpy(cury_add_prevs, cury_adds, sizeof cury_adds);
            memset(cury_adds, 0, sizeof cury_adds);
        }
    }
}

void orc_deposit_one(orc_rank *r, float x2, float y2, float z2, float x1, float y1, float z1, float q)
{
    if (r->P.order == 0) zigzag(r, x2, y2, z2, x1, y1, z1, q);
    else densdecomp(r, x2, y2, z2, x1, y1, z1, q);
}

/* loop A of deposit_particles: particles_movedeposit.F90:1381-1401 (ions), :1717-1737 (electrons) */
static void deposit_species(orc_rank *r, int first, int count, float qs)
{
    const float c = r->P.c;
    for (int n = 0; n < count; n++) {
        orc_particle *p = &r->p[first + n];
        float invgam = 1.f / sqrtf(1 + p->u * p->u + p->v * p->v + p->w * p->w);
        float x0 = p->x - p->u * invgam * c;
        float y0 = p->y - p->v * invgam * c;
        float z0 = p->z - p->w * invgam * c;
        float q = p->ch * qs;
        orc_deposit_one(r, p->x, p->y, p->z, x0, y0, z0, q);
    }
}
void orc_deposit_currents_only(orc_rank *r)
{
    deposit_species(r, 0, r->ions, r->P.qi);
    deposit_species(r, r->maxhlf, r->lecs, r->P.qe);
}

/* loops B and C of deposit_particles: particles_movedeposit.F90:1546-1705 / :1886-2040 */
static void classify_species(orc_rank *r, int first, int *count, int is_lec)
{
    const int three = r->P.dim == 3;
    const int nghost = r->nghost, nghostz = r->nghostz;
    const int sx = r->P.sizex, sy = r->P.sizey, sz = r->P.sizez, rank = r->rank;
    /* :1359-1374 */
    const float maxx = r->mx - 1.f * (nghost / 2), minx = 1.f * (nghost / 2 + 1);
    const float maxy = r->my - 1.f * (nghost / 2), miny = 1.f * (nghost / 2 + 1);
    const float midx = .5f * (maxx - minx), midy = .5f * (maxy - miny);
    float maxz, minz;
    if (three) { maxz = r->mz - 1.f * (nghostz / 2); minz = 1.f * (nghostz / 2 + 1); }
    else { minz = 1.f * (nghostz / 2 + 1); maxz = 1.f * (nghostz / 2 + 1) + 1; }
    const float midz = .5f * (maxz - minz);
    int np = *count;
    int32_t *pind = r->pind + first;
    orc_particle *P = r->p + first;
    for (int n = 0; n < np; n++) {
        int in = 1;
        float perx = 0.f, pery = 0.f, perz = 0.f;
        orc_particle *p = &P[n];
        if (p->x < minx || p->x > maxx) perx = fsign(midx, p->x - minx) + fsign(midx, p->x - maxx);
        if (p->y < miny || p->y > maxy) pery = fsign(midy, p->y - miny) + fsign(midy, p->y - maxy);
        if (p->z < minz || p->z > maxz) perz = fsign(midz, p->z - minz) + fsign(midz, p->z - maxz);
        if (r->P.periodicx == 0) in = (p->x + r->mxcum > r->x1in) && (p->x + r->mxcum < r->x2in);
        if (r->P.periodicy == 0 && in) in = (p->y + r->mycum > r->y1in) && (p->y + r->mycum < r->y2in);
        if (three && r->P.periodicz == 0 && in) in = (p->z + r->mzcum > r->z1in) && (p->z + r->mzcum < r->z2in);
        if (!in) { perx = 0; pery = 0; perz = 0; }
        if (perx != 0 && in && sx != 1) { in = 0; pery = 0; perz = 0; }
        if (perx < 0 && sx != 1) {
            int i1 = (rank / sx) * sx + imodulo(rank - 1, sx);
            perx = -(r->mxl[i1] - 1.f * nghost);
        }
        p->x = p->x - perx;
        if (pery != 0 && in && sy != 1) { in = 0; perx = 0; perz = 0; }
        if (pery < 0 && sy != 1) {
            int j1 = imodulo(rank / sx - 1, sy) * sx + rank / (sx * sy) * (sx * sy) + imodulo(rank, sx);
            pery = -(r->myl[j1] - 1.f * nghost);
        }
        p->y = p->y - pery;
        if (three) {
            if (perz != 0 && in) { in = 0; perx = 0; pery = 0; }
            if (perz < 0) {
                int k1 = imodulo(rank / (sx * sy) - 1, sz) * (sx * sy) + imodulo(rank, sx * sy);
                perz = -(r->mzl[k1] - 1.f * nghostz);
            }
        }
        p->z = p->z - perz;
        if (in) continue;
        int dir = -1;
        if (three) { if (perz < 0) dir = 4; if (perz > 0) dir = 5; }
        if (sy != 1) { if (pery < 0) dir = 2; if (pery > 0) dir = 3; }
        if (sx != 1) { if (perx < 0) dir = 0; if (perx > 0) dir = 1; }
        if (dir >= 0) {
            orc_box *b = &r->out[dir];
            if (b->nion + b->nlec >= r->P.buffsize) { fprintf(stderr, "oracle: outbox overflow\n"); abort(); }
            if (!is_lec) b->p[b->nion++] = *p;
            else b->p[b->nion + b->nlec++] = *p;
        }
        pind[n] = 1;
    }
    /* compaction, :1694-1705 */
    int cnt = np;
    for (int n = 0; n < np; n++) {
        if (pind[n] != 0) {
            while (pind[n] != 0) {
                P[n] = P[cnt - 1];
                pind[n] = pind[cnt - 1];
                pind[cnt - 1] = 0;
                cnt--;
            }
        }
    }
    *count = cnt;
}

/* deposit_particles: particles_movedeposit.F90:1281-2051 */
void orc_deposit_particles(orc_rank *r)
{
    for (int d = 0; d < 6; d++) { r->out[d].nion = 0; r->out[d].nlec = 0; }   /* :1326-1354 */
    if (r->ions > 0) { deposit_species(r, 0, r->ions, r->P.qi); classify_species(r, 0, &r->ions, 0); }
    if (r->lecs > 0) { deposit_species(r, r->maxhlf, r->lecs, r->P.qe); classify_species(r, r->maxhlf, &r->lecs, 1); }
}

/* exchange_particles: particles.F90:1865-2116 -- z pair, then y pair, then x pair; each rank's
   out[dir] becomes the dir-neighbour's in[opposite] (ions first, then electrons). */
void orc_exchange_particles(orc_world *w)
{
    for (int rk = 0; rk < w->size0; rk++)
        for (int d = 0; d < 6; d++) { w->r[rk]->in[d].nion = 0; w->r[rk]->in[d].nlec = 0; }
    for (int rk = 0; rk < w->size0; rk++) {
        orc_rank *r = w->r[rk];
        for (int d = 0; d < 6; d++) {
            int axis = d / 2;
            if (axis == 0 && r->P.sizex == 1) continue;         /* :2066 */
            if (axis == 1 && r->P.sizey == 1) continue;         /* :1974 */
            if (axis == 2 && w->P.dim == 2) continue;
            orc_rank *dst = w->r[orc_neighbour(r, d)];
            /* arrivals travelling in +dir are recorded as coming from the - side (in[d^1]) */
            orc_box *ib = &dst->in[d ^ 1], *ob = &r->out[d];
            memcpy(ib->p, ob->p, (size_t)(ob->nion + ob->nlec) * sizeof(orc_particle));
            ib->nion = ob->nion; ib->nlec = ob->nlec;
        }
    }
}

/* inject_others: particles.F90:1368-1852.  Appends arrivals; arrivals from y are re-tested in z
   (3D), arrivals from x re-tested in y, and re-queued for the second exchange round. */
static void append_arrival(orc_rank *r, const orc_particle *src, int is_lec, int retest_axis)
{
    const int sx = r->P.sizex, sy = r->P.sizey, sz = r->P.sizez, rank = r->rank;
    orc_particle q = *src;
    if (retest_axis == 2) {
        const int ngz = r->nghostz;
        /* :1431 : no "outside" guard here, unlike deposit_particles */
        float perz = fsign(.5f * (r->mz - 1.f * ngz), q.z - 1.f * (ngz / 2 + 1)) +
                     fsign(.5f * (r->mz - 1.f * ngz), q.z - r->mz + 1.f * (ngz / 2));
        if (perz < 0) {
            int k1 = imodulo(rank / (sx * sy) - 1, sz) * (sx * sy) + imodulo(rank, sx * sy);
            perz = -(r->mzl[k1] - 1.f * ngz);
        }
        q.z = q.z - perz;
        if (perz != 0) {
            orc_box *b = &r->out[perz < 0 ? 4 : 5];
            if (!is_lec) b->p[b->nion++] = q; else b->p[b->nion + b->nlec++] = q;
            return;
        }
    } else if (retest_axis == 1) {
        const int ng = r->nghost;
        float pery = fsign(.5f * (r->my - 1.f * ng), q.y - 1.f * (ng / 2 + 1)) +
                     fsign(.5f * (r->my - 1.f * ng), q.y - r->my + 1.f * (ng / 2));
        if (pery < 0) {
            int j1 = imodulo(rank / sx - 1, sy) * sx + rank / (sx * sy) * (sx * sy) + imodulo(rank, sx);
            pery = -(r->myl[j1] - 1.f * ng);
        }
        q.y = q.y - pery;
        if (pery != 0) {
            orc_box *b = &r->out[pery < 0 ? 2 : 3];
            if (!is_lec) b->p[b->nion++] = q; else b->p[b->nion + b->nlec++] = q;
            return;
        }
    }
    if (!is_lec) { if (r->ions >= r->maxhlf) { fprintf(stderr, "oracle: ion overflow\n"); abort(); } r->p[r->ions++] = q; }
    else { if (r->lecs >= r->maxhlf) { fprintf(stderr, "oracle: lec overflow\n"); abort(); } r->p[r->maxhlf + r->lecs++] = q; }
}

void orc_inject_others(orc_rank *r)
{
    const int three = r->P.dim == 3;
    for (int d = 0; d < 6; d++) { r->out[d].nion = 0; r->out[d].nlec = 0; }    /* :1379-1394 */
    /* in[5] = "abv" (came from above, travelling -z), in[4] = "blw" */
    if (three) {
        for (int s = 5; s >= 4; s--) {
            orc_box *b = &r->in[s];
            for (int n = 0; n < b->nion; n++) append_arrival(r, &b->p[n], 0, -1);
            for (int n = 0; n < b->nlec; n++) append_arrival(r, &b->p[b->nion + n], 1, -1);
        }
    }
    if (r->P.sizey != 1) {
        int rt = three ? 2 : -1;
        /* ions from rgt, ions from lft, lecs from rgt, lecs from lft : :1424-1640 */
        for (int n = 0; n < r->in[3].nion; n++) append_arrival(r, &r->in[3].p[n], 0, rt);
        for (int n = 0; n < r->in[2].nion; n++) append_arrival(r, &r->in[2].p[n], 0, rt);
        for (int n = 0; n < r->in[3].nlec; n++) append_arrival(r, &r->in[3].p[r->in[3].nion + n], 1, rt);
        for (int n = 0; n < r->in[2].nlec; n++) append_arrival(r, &r->in[2].p[r->in[2].nion + n], 1, rt);
    }
    if (r->P.sizex != 1) {
        int rt = r->P.sizey != 1 ? 1 : -1;
        for (int n = 0; n < r->in[1].nion; n++) append_arrival(r, &r->in[1].p[n], 0, rt);
        for (int n = 0; n < r->in[0].nion; n++) append_arrival(r, &r->in[0].p[n], 0, rt);
        for (int n = 0; n < r->in[1].nlec; n++) append_arrival(r, &r->in[1].p[r->in[1].nion + n], 1, rt);
        for (int n = 0; n < r->in[0].nlec; n++) append_arrival(r, &r->in[0].p[r->in[0].nion + n], 1, rt);
    }
    for (int d = 0; d < 6; d++) { r->in[d].nion = 0; r->in[d].nlec = 0; }
}

/* reorder_particles_: particles.F90:418-497 -- stable counting sort by cell key
   int(x) + int(y)*iy + int(z)*iz (1-based pieces exactly as the reference forms them) */
static void reorder_species(orc_rank *r, int first, int count)
{
    if (count <= 0) return;
    size_t lot = r->lot;
    int *pall = (int *)calloc(lot + 2, sizeof(int));
    orc_particle *P = r->p + first, *tmp = (orc_particle *)malloc((size_t)count * sizeof(orc_particle));
    const long iy = r->iy, iz = r->iz;
    for (int n = 0; n < count; n++) {
        long key = (long)(int)P[n].x + (long)(int)P[n].y * iy + (long)(int)P[n].z * iz;
        if (key < 0) key = 0; if ((size_t)key > lot) key = (long)lot;
        pall[key]++;
    }
    int acc = 0;
    for (size_t c = 0; c <= lot; c++) { int t = pall[c]; pall[c] = acc; acc += t; }
    for (int n = 0; n < count; n++) {
        long key = (long)(int)P[n].x + (long)(int)P[n].y * iy + (long)(int)P[n].z * iz;
        if (key < 0) key = 0; if ((size_t)key > lot) key = (long)lot;
        tmp[pall[key]++] = P[n];
    }
    memcpy(P, tmp, (size_t)count * sizeof(orc_particle));
    free(tmp); free(pall);
}
void orc_reorder_particles(orc_rank *r)
{
    reorder_species(r, 0, r->ions);
    reorder_species(r, r->maxhlf, r->lecs);
}

/* ------------------------------------------------------------------------- */
/* one lap: tristanmainloop.F90:107-344 (periodic / plain-open configurations;  */
/* user hooks and injectors are in orc_step_shock; pre/post_bc_* edge fixes of an    */
/* all-open box (fieldboundaries.F90:120,154,442,468) are outside this restatement) */
/* ------------------------------------------------------------------------- */
enum { PH_SURF_B = 100, PH_SURF_E, PH_PRE_B, PH_POST_B, PH_PRE_E, PH_POST_E };
enum { PH_BC_B1, PH_BC_E1, PH_BHALF, PH_MOVE, PH_EFULL, PH_RESET, PH_DEPOSIT, PH_EXCH_P, PH_EXCH_CUR,
       PH_FILTER, PH_ADD_CUR, PH_INJECT_OTHERS, PH_REORDER };

void orc_step_phase(orc_world *w, int phase)
{
    int n = w->size0;
    switch (phase) {
    case PH_BC_B1: orc_bc_fields(w, ORC_BX); break;
    case PH_BC_E1: orc_bc_fields(w, ORC_EX); break;
    case PH_EXCH_P: orc_exchange_particles(w); break;
    case PH_EXCH_CUR: orc_exchange_current(w); break;
    case PH_FILTER: orc_apply_filter(w); break;
    default:
#pragma omp parallel for schedule(static)
        for (int rk = 0; rk < n; rk++) {
            orc_rank *r = w->r[rk];
            switch (phase) {
            case PH_BHALF: orc_advance_b_halfstep(r); break;
            case PH_MOVE: orc_move_particles(r); break;
            case PH_EFULL: orc_advance_e_fullstep(r); break;
            case PH_RESET: orc_reset_currents(r); break;
            case PH_DEPOSIT: orc_deposit_particles(r); break;
            case PH_ADD_CUR: orc_add_current(r); break;
            case PH_INJECT_OTHERS: orc_inject_others(r); break;
            case PH_REORDER: orc_reorder_particles(r); break;
            case PH_SURF_B: orc_surface_b(r); break;
            case PH_SURF_E: orc_surface_e(r); break;
            case PH_PRE_B: orc_edges(r, 0); break;
            case PH_POST_B: orc_edges(r, 1); break;
            case PH_PRE_E: orc_edges(r, 2); break;
            case PH_POST_E: orc_edges(r, 3); break;
            }
        }
    }
}

void orc_step(orc_world *w)
{
    w->lap++;
    const int edges = all_open(w);        /* pre_/post_bc_* only act in an all-open 3D box */
    if (edges) { orc_step_phase(w, PH_PRE_B); orc_step_phase(w, PH_BC_B1); }     /* :114 pre_bc_b */
    orc_step_phase(w, PH_BC_B1);          /* :117 */
    orc_step_phase(w, PH_BC_E1);          /* :118 */
    orc_step_phase(w, PH_BHALF);          /* :119 */
    orc_step_phase(w, PH_BC_B1);          /* :122 */
    orc_step_phase(w, PH_MOVE);           /* :134 */
    orc_step_phase(w, PH_BHALF);          /* :139 */
    orc_step_phase(w, PH_BC_B1);          /* :140 */
    orc_step_phase(w, PH_SURF_B);         /* :145 bc_b2 = surface on every radiating axis ... */
    orc_step_phase(w, PH_BC_B1);          /*      ... then bc_b1 */
    if (edges) { orc_step_phase(w, PH_POST_B); orc_step_phase(w, PH_BC_B1); }    /* :155 post_bc_b */
    if (edges) { orc_step_phase(w, PH_PRE_E); orc_step_phase(w, PH_BC_E1); }     /* :157 pre_bc_e */
    orc_step_phase(w, PH_EFULL);          /* :159 */
    orc_step_phase(w, PH_SURF_E);         /* :164 bc_e2 = surface ... */
    orc_step_phase(w, PH_BC_E1);          /*      ... then bc_e1 */
    if (edges) { orc_step_phase(w, PH_POST_E); orc_step_phase(w, PH_BC_E1); }    /* :165 post_bc_e */
    orc_step_phase(w, PH_RESET);          /* :171 */
    orc_step_phase(w, PH_BC_E1);          /* :181 */
    orc_step_phase(w, PH_BC_B1);          /* :182 */
    orc_step_phase(w, PH_DEPOSIT);        /* :183 */
    orc_step_phase(w, PH_EXCH_P);         /* :190 */
    orc_step_phase(w, PH_EXCH_CUR);       /* :203 */
    orc_step_phase(w, PH_FILTER);         /* :213-229 */
    orc_step_phase(w, PH_ADD_CUR);        /* :242 */
    orc_step_phase(w, PH_INJECT_OTHERS);  /* :257 */
    orc_step_phase(w, PH_EXCH_P);         /* :268 */
    orc_step_phase(w, PH_INJECT_OTHERS);  /* :272 */
    if (w->lap % 10 == 0) orc_step_phase(w, PH_REORDER);   /* particles.F90:398-400 */
}

/* ------------------------------------------------------------------------- */
/* seeded loader: aux.F90:82-134; particles.F90:2126-2273, 2549-2938            */
/* ------------------------------------------------------------------------- */
float orc_random(double *dseed)
{
    double seed = fmod(16807.0 * (*dseed), 2147483647.0);
    *dseed = seed;
    return (float)(seed / 2147483648.0);
}
static float poisson(double *dseed, float numps)
{
    double Lps = exp(-(double)numps), pps = 1; float kps = 0;
    while (pps >= Lps) { kps = kps + 1; pps = pps * orc_random(dseed); }
    return kps - 1;
}
#define PDF_SZ 1000
void orc_init_maxw_table(int dim, int pcosthmult, float delgam, float *gamma_table, float *pdf_table)
{
    /* particles.F90:2126-2163; pdf_table has PDF_SZ+1 entries */
    float func[PDF_SZ];
    float maxg = delgam * 20 + 1.f;
    for (int i = 1; i <= PDF_SZ; i++) gamma_table[i - 1] = (maxg - 1.f) / (PDF_SZ - 1) * (i - 1);
    for (int i = 0; i < PDF_SZ; i++) {
        float g = gamma_table[i];
        if (dim == 3 || pcosthmult == 1) func[i] = (g + 1.f) * sqrtf(g * (g + 2.f)) * expf(-g / delgam);
        else func[i] = (g + 1.f) * expf(-g / delgam);
    }
    pdf_table[0] = 0.f;
    float acc = 0.f;
    for (int i = 1; i <= PDF_SZ; i++) { acc = acc + func[i - 1]; pdf_table[i] = acc; }
    float norm = pdf_table[PDF_SZ - 1];                        /* normalised by entry pdf_sz, not pdf_sz+1 */
    for (int i = 0; i <= PDF_SZ; i++) pdf_table[i] = pdf_table[i] / norm;
}

void orc_maxwell_dist(int dim, int pcosthmult, float sigma, float gamma0, float cd, double *dseed,
                      float *u, float *v, float *w, const float *gamma_table, const float *pdf_table)
{
    /* particles.F90:2176-2273 */
    const double pi = (double)3.1415927f;                      /* fp32 literal held in fp64, :292 */
    float gamma0mag = fabsf(gamma0);
    float gm = fabsf(gamma0) > 1.f ? fabsf(gamma0) : 1.f;
    float beta_drift = fsign(sqrtf(1.f - 1.f / (gm * gm)), gamma0);
    float rannum = orc_random(dseed);
    if (rannum == 1.0f) rannum = orc_random(dseed);
    int i = 1, flag = 1; float gam = 0.f;
    while (flag) {
        if (i == PDF_SZ) { gam = gamma_table[PDF_SZ - 1]; flag = 0; }
        if (rannum >= pdf_table[i - 1] && rannum < pdf_table[i]) {
            gam = gamma_table[i - 1] + (gamma_table[i < PDF_SZ ? i : PDF_SZ - 1] - gamma_table[i - 1]) /
                  (pdf_table[i] - pdf_table[i - 1]) * (rannum - pdf_table[i - 1]);
            flag = 0;
        }
        i++;
    }
    float pcosth;
    if (dim == 2) {
        pcosth = (2 * orc_random(dseed) - 1) * pcosthmult;
        if (sigma != 0.f) pcosth = 2 * orc_random(dseed) - 1;
    } else pcosth = 2 * orc_random(dseed) - 1;
    float pphi = (float)(orc_random(dseed) * 2 * pi);
    float psinth = sqrtf(1 - pcosth * pcosth);
    float v0t = cd * sqrtf(gam * (gam + 2.f)) / (1.f + gam);
    float ut1 = v0t * psinth * cosf(pphi), vt1 = v0t * psinth * sinf(pphi), wt1 = v0t * pcosth;
    float ptx = (1.f + gam) * ut1, pty = (1.f + gam) * vt1, ptz = (1.f + gam) * wt1;
    float X7 = orc_random(dseed);
    if (-beta_drift * ut1 / cd > X7) ptx = -ptx;
    float px1 = (ptx + cd * beta_drift * (gam + 1.f)) * gamma0mag;
    *u = px1 / cd; *v = pty / cd; *w = ptz / cd;
}

static void clamp_range(float lo_in, float hi_in, int gmin, int gmax, int cum, float lo_edge, float hi_edge,
                        float *lo, float *hi)
{
    /* particles.F90:2619-2641 */
    *lo = lo_edge; *hi = lo_edge;
    if (lo_in < gmin) *lo = lo_edge;
    if (hi_in < gmin) *hi = lo_edge;
    if (lo_in >= gmax) *lo = hi_edge;
    if (hi_in >= gmax) *hi = hi_edge;
    if (lo_in >= gmin && lo_in < gmax) *lo = lo_in - 1.f * cum;
    if (hi_in >= gmin && hi_in < gmax) *hi = hi_in - 1.f * cum;
}

void orc_inject_plasma_region(orc_rank *r, float x1, float x2, float y1, float y2, float z1, float z2,
                              float ppc, float gamma_drift_in, float delgam_i, float delgam_e,
                              float weight, int direction, int pcosthmult, float sigma)
{
    /* particles.F90:2549-2938 with upsamp_e = upsamp_i = 1, no density profile */
    static float gti[PDF_SZ], pti[PDF_SZ + 1], gte[PDF_SZ], pte[PDF_SZ + 1];
    const int dim = r->P.dim, g = r->nghost / 2, gz = r->nghostz / 2;
    float gamma_drift = gamma_drift_in;
    if (fabsf(gamma_drift_in) < 1) gamma_drift = fsign(sqrtf(1.f / (1.f - gamma_drift_in * gamma_drift_in)), gamma_drift_in);
    orc_init_maxw_table(dim, pcosthmult, delgam_i, gti, pti);
    orc_init_maxw_table(dim, pcosthmult, delgam_e, gte, pte);
    float minx, maxx, miny, maxy, minz, maxz;
    clamp_range(x1, x2, (g + 1) + r->mxcum, (r->mx - g) + r->mxcum, r->mxcum, 1.f * (g + 1), (float)(r->mx - g), &minx, &maxx);
    clamp_range(y1, y2, (g + 1) + r->mycum, (r->my - g) + r->mycum, r->mycum, 1.f * (g + 1), (float)(r->my - g), &miny, &maxy);
    if (dim == 2) { maxz = 1.f * (gz + 1) + 1.f; minz = 1.f * (gz + 1); }
    else clamp_range(z1, z2, (gz + 1) + r->mzcum, (r->mz - gz) + r->mzcum, r->mzcum, 1.f * (gz + 1), (float)(r->mz - gz), &minz, &maxz);
    float delta_x = maxx - minx, delta_y = maxy - miny, delta_z = maxz - minz;
    float numps = (.5f * ppc) * delta_z * delta_y * delta_x;
    if (numps < 10) { if (numps != 0.f) numps = poisson(&r->dseed, numps); else numps = 0.f; }
    else numps = ceilf(numps);
    int n = 0;
    while (n < (int)numps) {
        n++;
        if (r->ions >= r->maxhlf || r->lecs >= r->maxhlf) { fprintf(stderr, "oracle: loader overflow\n"); abort(); }
        orc_particle *pi = &r->p[r->ions++];
        pi->x = minx + delta_x * orc_random(&r->dseed);
        pi->y = miny + delta_y * orc_random(&r->dseed);
        pi->z = minz + delta_z * orc_random(&r->dseed);
        orc_maxwell_dist(dim, pcosthmult, sigma, gamma_drift, r->P.c, &r->dseed, &pi->u, &pi->v, &pi->w, gti, pti);
        pi->ch = weight;
        if (direction == 2) { float t = pi->u; pi->u = pi->v; pi->v = t; }
        if (direction == 3) { float t = pi->u; pi->u = pi->w; pi->w = t; }
        pi->ind = ++r->totalpartnum; pi->proc = r->rank; pi->splitlev = 1;
        orc_particle *pe = &r->p[r->maxhlf + r->lecs++];
        pe->x = pi->x; pe->y = pi->y; pe->z = pi->z; pe->ch = weight;
        orc_maxwell_dist(dim, pcosthmult, sigma, gamma_drift, r->P.c, &r->dseed, &pe->u, &pe->v, &pe->w, gte, pte);
        if (direction == 2) { float t = pe->u; pe->u = pe->v; pe->v = t; }
        if (direction == 3) { float t = pe->u; pe->u = pe->w; pe->w = t; }
        pe->ind = ++r->totalpartnum; pe->proc = r->rank; pe->splitlev = 1;
    }
}

/* inject_from_wall: particles.F90:2439-2538 (upsamp_e = upsamp_i = 1) -- the slab a plane source fills in one step */
void orc_inject_from_wall(orc_rank *r, float x1, float x2, float y1, float y2, float z1, float z2, float ppc, float gamma_drift,
                          float delgam_i, float delgam_e, float wall_speed, float weight, int pcosthmult, float sigma)
{
    const float c = r->P.c;
    float beta_wall, beta_inj, xt;
    int direction = 0;                                      /* undefined in the reference when no pair coincides */
    if (fabsf(wall_speed) >= 1) beta_wall = fsign(sqrtf(1 - 1 / (wall_speed * wall_speed)), wall_speed);
    else beta_wall = wall_speed;
    if (fabsf(gamma_drift) >= 1) beta_inj = fsign(sqrtf(1 - 1.f / (gamma_drift * gamma_drift)), gamma_drift);
    else beta_inj = gamma_drift;
    float x1n = x1, x2n = x2, y1n = y1, y2n = y2, z1n = z1, z2n = z2;
    if (x1 == x2) { x2n = x1 + (beta_inj - beta_wall) * c; direction = 1; if (x2n < x1n) { xt = x1n; x1n = x2n; x2n = xt; } }
    if (y1 == y2) { y2n = y1 + (beta_inj - beta_wall) * c; direction = 2; if (y2n < y1n) { xt = y1n; y1n = y2n; y2n = xt; } }
    if (z1 == z2) { z2n = z1 + (beta_inj - beta_wall) * c; direction = 3; if (z2n < z1n) { xt = z1n; z1n = z2n; z2n = xt; } }
    orc_inject_plasma_region(r, x1n, x2n, y1n, y2n, z1n, z2n, ppc, gamma_drift, delgam_i, delgam_e, weight, direction, pcosthmult, sigma);
}
/* inject_particles_user of the shock problem: user/user_shock.F90:303-331 (a non-receding plane source at x = mx0 - 2) */
void orc_inject_particles_shock(orc_world *w, float ppc0, float gamma0_in, float delgam, float me, float mi,
                                float temperature_ratio, int pcosthmult, float sigma)
{
    float gamma0 = gamma0_in;
    if (gamma0 < 1) gamma0 = sqrtf(1.f / (1.f - gamma0 * gamma0));                       /* particles.F90:213 */
    for (int rk = 0; rk < w->size0; rk++) {
        const float x1 = w->mx0g - 2.f, y1 = 3.f, y2 = w->my0g - 2.f, z1 = 3.f, z2 = w->mz0g - 2.f;
        orc_inject_from_wall(w->r[rk], x1, x1, y1, y2, z1, z2, ppc0, -gamma0, delgam, delgam * mi / me * temperature_ratio, 0.f, 1.f,
                             pcosthmult, sigma);
    }
}

/* read_input_particles: particles.F90:219-235 */
void orc_charge_normalisation(orc_params *P, float ppc0, float c_omp, float gamma0, float me, float mi)
{
    if (gamma0 < 1) gamma0 = sqrtf(1.f / (1.f - gamma0 * gamma0));
    float omp = P->c / c_omp;
    P->qe = -(omp * omp * gamma0) / ((ppc0 * .5f) * (1 + me / mi));
    P->qi = -P->qe;
    me = me * fabsf(P->qi); mi = mi * fabsf(P->qi);
    P->qme = P->qe / me; P->qmi = P->qi / mi;
}

/* user/user_weibel.F90:255-313 */
void orc_init_weibel(orc_world *w, float ppc0, float gamma0_in, float delgam, float me, float mi,
                     float temperature_ratio, int distr_dim)
{
    float gamma0 = gamma0_in;
    if (gamma0 < 1) gamma0 = sqrtf(1.f / (1.f - gamma0 * gamma0));
    int pcosthmult = (w->P.dim == 2 && distr_dim == 2) ? 0 : 1;   /* user_weibel.F90:160-167 */
    for (int rk = 0; rk < w->size0; rk++) {
        orc_rank *r = w->r[rk];
        int g = r->nghost / 2;
        float xinject = 1.f * (g + 1), xinject2 = w->mx0g - 1.f * g;
        float y1 = 1.f * (g + 1), y2 = w->my0g - 1.f * g;
        float z1 = 3.f, z2 = w->mz0g - 2.f;
        float de = delgam * mi / me * temperature_ratio;
        orc_inject_plasma_region(r, xinject, xinject2, y1, y2, z1, z2, ppc0 / 2.f, -gamma0, delgam, de, 1.f, 1, pcosthmult, 0.f);
        orc_inject_plasma_region(r, xinject, xinject2, y1, y2, z1, z2, ppc0 / 2.f, gamma0, delgam, de, 1.f, 1, pcosthmult, 0.f);
        orc_reorder_particles(r);
    }
}

/* user/user_twostream.F90:226-270 : electrons only (ions = 0 after loading), pcosthmult = 0 */
void orc_init_twostream(orc_world *w, float ppc0, float gamma0_in, float delgam, float me, float mi,
                        float temperature_ratio)
{
    float gamma0 = gamma0_in;
    if (gamma0 < 1) gamma0 = sqrtf(1.f / (1.f - gamma0 * gamma0));
    for (int rk = 0; rk < w->size0; rk++) {
        orc_rank *r = w->r[rk];
        float x1 = 3.f, x2 = w->mx0g - 2.f, y1 = 3.f, y2 = w->my0g - 2.f, z1 = 3.f, z2 = w->mz0g - 2.f;
        float de = delgam * mi / me * temperature_ratio;
        orc_inject_plasma_region(r, x1, x2, y1, y2, z1, z2, ppc0 / 2.f, -gamma0, delgam, de, 1.f, 1, 0, 0.f);
        orc_inject_plasma_region(r, x1, x2, y1, y2, z1, z2, ppc0 / 2.f, gamma0, delgam, de, 1.f, 1, 0, 0.f);
        r->ions = 0;
        orc_reorder_particles(r);
    }
}

/* Fast synthetic loader for benchmark-sized problems (not a reference routine): uniform positions in
   the local interior, two counter-streaming beams +-beta_drift in x with an isotropic spread uth.
   splitmix64 + Box-Muller-free (sum of uniforms) so that it is cheap at 1e8 particles. */
static uint64_t sm64(uint64_t *s) { uint64_t z = (*s += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
static float u01(uint64_t *s) { return (float)((sm64(s) >> 40) * (1.0 / 16777216.0)); }
void orc_init_uniform(orc_world *w, float ppc0, float beta_drift, float uth, uint64_t seed)
{
    float gdrift = 1.f / sqrtf(1.f - beta_drift * beta_drift);
    for (int rk = 0; rk < w->size0; rk++) {
        orc_rank *r = w->r[rk];
        uint64_t s = seed + 0x1234567ull * (uint64_t)(rk + 1);
        int g = r->nghost / 2, gz = r->nghostz / 2;
        float minx = g + 1, maxx = r->mx - g, miny = g + 1, maxy = r->my - g;
        float minz = gz + 1, maxz = r->P.dim == 3 ? r->mz - gz : gz + 2;
        double cells = (double)(maxx - minx) * (maxy - miny) * (r->P.dim == 3 ? (maxz - minz) : 1.0);
        long np = (long)(0.5 * ppc0 * cells);
        if (np > r->maxhlf) np = r->maxhlf;
        for (long n = 0; n < np; n++) {
            orc_particle a;
            a.x = minx + (maxx - minx) * u01(&s); if (a.x >= maxx) a.x = minx;
            a.y = miny + (maxy - miny) * u01(&s); if (a.y >= maxy) a.y = miny;
            a.z = minz + (maxz - minz) * u01(&s); if (a.z >= maxz) a.z = minz;
            float sgn = (n & 1) ? 1.f : -1.f;
            for (int sp = 0; sp < 2; sp++) {
                orc_particle b = a;
                float t0 = (u01(&s) + u01(&s) + u01(&s) - 1.5f) * 2.f, t1 = (u01(&s) + u01(&s) + u01(&s) - 1.5f) * 2.f,
                      t2 = (u01(&s) + u01(&s) + u01(&s) - 1.5f) * 2.f;
                b.u = sgn * gdrift * beta_drift + uth * t0; b.v = uth * t1; b.w = uth * t2;
                b.ch = 1.f; b.ind = ++r->totalpartnum; b.proc = r->rank; b.splitlev = 1;
                if (sp == 0) r->p[r->ions++] = b; else r->p[r->maxhlf + r->lecs++] = b;
            }
        }
    }
}

/* ------------------------------------------------------------------------- */
/* shock-problem user hooks: user/user_shock.F90                                */
/* ------------------------------------------------------------------------- */
static int iloc(const orc_rank *r, int iglob)
{
    /* fields.F90:384-390 */
    int i = iglob - r->mxcum;
    if (i < r->nghost / 2 + 1) i = 1;
    if (i > r->mx - (r->nghost / 2)) i = r->mx;
    return i;
}
static void fill_x(orc_rank *r, int which, int i1, int i2, float v)
{
    for (int k = 1; k <= r->mz; k++) for (int j = 1; j <= r->my; j++) for (int i = i1; i <= i2; i++) r->f[which][IDX(r, i, j, k)] = v;
}
/* field_bc_user, user_shock.F90:342-373: conductor behind the left wall, upstream fields clamped at the right edge */
void orc_field_bc_shock(orc_rank *r, float leftwall, float binit, float btheta, float bphi, float beta)
{
    float xmin = 1.f, xmax = leftwall - 10.f;
    int i1 = iloc(r, (int)xmin), i2 = iloc(r, (int)xmax);
    if (i1 != i2) { fill_x(r, ORC_EY, i1, i2, 0.f); fill_x(r, ORC_EZ, i1, i2, 0.f); }
    xmin = r->mx0g - 2.f; xmax = (float)r->mx0g;
    i1 = iloc(r, (int)xmin); i2 = iloc(r, (int)xmax);
    if (i1 != i2) {
        float bxv = binit * cosf(btheta), byv = binit * sinf(btheta) * sinf(bphi), bzv = binit * sinf(btheta) * cosf(bphi);
        fill_x(r, ORC_BX, i1, i2, bxv); fill_x(r, ORC_BY, i1, i2, byv); fill_x(r, ORC_BZ, i1, i2, bzv);
        fill_x(r, ORC_EX, i1, i2, 0.f);
        fill_x(r, ORC_EY, i1, i2, (-beta) * bzv);
        fill_x(r, ORC_EZ, i1, i2, -(-beta) * byv);
    }
}
/* particle_bc_user, user_shock.F90:377-457 with gammawall = 1, betawall = 0: specular wall at x = leftwall (global) */
void orc_particle_bc_wall(orc_rank *r, float leftwall)
{
    const float c = r->P.c, gammawall = 1.f, betawall = 0.f, walloc = leftwall;
    for (int iter = 1; iter <= 2; iter++) {
        int i0 = iter == 1 ? 0 : r->maxhlf, cnt = iter == 1 ? r->ions : r->lecs;
        float q0 = iter == 1 ? r->P.qi : r->P.qe;
        for (int n = 0; n < cnt; n++) {
            orc_particle *p = &r->p[i0 + n];
            if (p->x + r->mxcum < walloc) {
                float gamma = sqrtf(1 + (p->u * p->u + p->v * p->v + p->w * p->w));
                float x0 = p->x - p->u / gamma * c, y0 = p->y, z0 = p->z;
                float walloc0 = walloc - betawall * c - r->mxcum;
                float tfrac = fabsf((x0 - walloc0) / (betawall * c - p->u / gamma * c));
                float xcolis = x0 + p->u / gamma * c * tfrac, ycolis = y0, zcolis = z0;
                float q = p->ch * q0;
                zigzag(r, xcolis, ycolis, zcolis, x0, y0, z0, q);
                p->u = gammawall * gammawall * gamma * (2 * betawall - p->u / gamma * (1 + betawall * betawall));
                gamma = sqrtf(1 + (p->u * p->u + p->v * p->v + p->w * p->w));
                float den = fabsf(p->x - x0); if (den < 1e-9f) den = 1e-9f;
                tfrac = fabsf((p->x - xcolis) / den); if (tfrac > 1.f) tfrac = 1.f;
                p->x = xcolis + p->u / gamma * c * tfrac;
                p->y = ycolis; p->z = zcolis;
                q = -q;
                zigzag(r, xcolis, ycolis, zcolis, p->x - p->u / gamma * c, p->y - p->v / gamma * c, p->z - p->w / gamma * c, q);
            }
        }
    }
}
/* one lap with the shock hooks at mainloop's hook points (tristanmainloop.F90:146-243); injector not included */
void orc_step_shock(orc_world *w, float leftwall, float binit, float btheta, float bphi, float beta)
{
    int n = w->size0;
#define FBC() for (int rk = 0; rk < n; rk++) orc_field_bc_shock(w->r[rk], leftwall, binit, btheta, bphi, beta)
    w->lap++;
    orc_step_phase(w, PH_BC_B1); orc_step_phase(w, PH_BC_E1); orc_step_phase(w, PH_BHALF); orc_step_phase(w, PH_BC_B1);
    orc_step_phase(w, PH_MOVE); orc_step_phase(w, PH_BHALF); orc_step_phase(w, PH_BC_B1);
    orc_step_phase(w, PH_SURF_B); orc_step_phase(w, PH_BC_B1);   /* :145 bc_b2 (x radiates) */
    FBC();                                   /* :146 */
    orc_step_phase(w, PH_EFULL);
    FBC();                                   /* :160 */
    orc_step_phase(w, PH_SURF_E); orc_step_phase(w, PH_BC_E1);   /* :164 bc_e2 */
    FBC();                                   /* :166 */
    orc_step_phase(w, PH_RESET);
    for (int rk = 0; rk < n; rk++) orc_particle_bc_wall(w->r[rk], leftwall);   /* :177 */
    orc_step_phase(w, PH_BC_E1); orc_step_phase(w, PH_BC_B1);
    orc_step_phase(w, PH_DEPOSIT); orc_step_phase(w, PH_EXCH_P); orc_step_phase(w, PH_EXCH_CUR); orc_step_phase(w, PH_FILTER);
    orc_step_phase(w, PH_ADD_CUR);
    FBC();                                   /* :243 */
    orc_step_phase(w, PH_INJECT_OTHERS); orc_step_phase(w, PH_EXCH_P); orc_step_phase(w, PH_INJECT_OTHERS);
    if (w->lap % 10 == 0) orc_step_phase(w, PH_REORDER);
#undef FBC
}

/* ------------------------------------------------------------------------- */
/* output-side moments: meanq_fld_cur(totname), output.F90:5229-5486            */
/* Every live particle adds `addprtx` (and the weight `addprty`) to the box of   */
/* half-width idx = idy = idz = 2 (idz = 0 in 2D, output.F90:189-195) around its  */
/* cell, clipped to the local array; curx/cury are the scratch arrays, as in the  */
/* reference.  The per-rank part is here; exchange_current() (:5436) and the      */
/* normalisation (:5439-5479) follow in orc_meanq_fld_cur.                        */
/* ------------------------------------------------------------------------- */
static void meanq_terms(const char *name, const orc_particle *q, int is_ion, float *ax, float *ay)
{
    const float gam = 1.f / sqrtf(1.f + q->u * q->u + q->v * q->v + q->w * q->w);      /* gamprt, :5264 */
    const int is_lec = !is_ion;
    float x = 0.f, y = 0.f;
#define IS(s) (strncmp(name, s, 5) == 0)
    if (IS("tdens")) x = q->ch;
    else if (IS("idens")) { if (is_ion) x = q->ch; }
    else if (IS("hdens")) { if (is_lec && q->ind > 0) x = q->ch; }
    else if (IS("ldens")) { if (is_lec && q->ind < 0) x = 1.f; }
    else if (IS("btden")) { if (q->ind < 0) x = q->ch; }
    else if (IS("biden")) { if (is_ion && q->ind < 0) x = q->ch; }
    else if ((name[0] == 't' || name[0] == 'e' || name[0] == 'i') && strncmp(name + 1, "bet", 3) == 0) {
        const float uu = name[4] == 'x' ? q->u : name[4] == 'y' ? q->v : q->w;
        if (name[0] == 't' || (name[0] == 'e' && is_lec) || (name[0] == 'i' && is_ion)) { x = uu * gam * q->ch; y = q->ch; }
    } else if ((name[0] == 't' || name[0] == 'i') && strncmp(name + 1, "mom", 3) == 0) {
        const float uu = name[4] == 'x' ? q->u : name[4] == 'y' ? q->v : q->w;
        if (name[0] == 't' || is_ion) { x = uu * q->ch; y = q->ch; }
    } else if (IS("eener")) { if (is_lec) { x = (1.f / gam - 1.f) * q->ch; y = q->ch; } }
    else if (IS("iener")) { if (is_ion) { x = (1.f / gam - 1.f) * q->ch; y = q->ch; } }
    else if ((name[0] == 'e' || name[0] == 'i') && name[1] == 'e' && name[2] == 't' && name[4] == '2') {
        const float uu = name[3] == 'x' ? q->u : name[3] == 'y' ? q->v : q->w;
        if ((name[0] == 'e' && is_lec) || (name[0] == 'i' && is_ion)) { x = (uu * gam) * (uu * gam) * q->ch; y = q->ch; }
    }
#undef IS
    *ax = x; *ay = y;
}
static void meanq_accumulate(orc_rank *r, const char *name)
{
    float *cx = r->f[ORC_CURX], *cy = r->f[ORC_CURY];
    const int idx = 2, idy = 2, idz = r->P.dim == 3 ? 2 : 0;
    orc_reset_currents(r);                                                             /* :5257-5259 */
    for (int sp = 0; sp < 2; sp++) {
        const int first = sp ? r->maxhlf : 0, cnt = sp ? r->lecs : r->ions;
        for (int n = 0; n < cnt; n++) {
            const orc_particle *q = &r->p[first + n];
            float ax, ay; meanq_terms(name, q, sp == 0, &ax, &ay);
            const int i = (int)q->x, j = (int)q->y, k = (int)q->z;                     /* :5405-5407 */
            int lz1 = k - idz < 1 ? 1 : k - idz, lz2 = k + idz > r->mz ? r->mz : k + idz;
            if (r->P.dim == 2) { lz1 = 1; lz2 = 1; }
            const int ly1 = j - idy < 1 ? 1 : j - idy, ly2 = j + idy > r->my ? r->my : j + idy;
            const int lx1 = i - idx < 1 ? 1 : i - idx, lx2 = i + idx > r->mx ? r->mx : i + idx;
            for (int kk = lz1; kk <= lz2; kk++) for (int jj = ly1; jj <= ly2; jj++) for (int ii = lx1; ii <= lx2; ii++) {
                const size_t l = IDX(r, ii, jj, kk);
                cx[l] = cx[l] + ax; cy[l] = cy[l] + ay;
            }
        }
    }
}
static void meanq_normalise(orc_rank *r, const char *name)
{
    float *cx = r->f[ORC_CURX], *cy = r->f[ORC_CURY];
    const int idx = 2, idy = 2, idz = r->P.dim == 3 ? 2 : 0;
    for (int k = 1; k <= r->mz; k++) for (int j = 1; j <= r->my; j++) for (int i = 1; i <= r->mx; i++) {
        int lz1 = k - idz < 1 ? 1 : k - idz, lz2 = k + idz > r->mz ? r->mz : k + idz;
        if (r->P.dim == 2) { lz1 = 1; lz2 = 1; }
        const int ly1 = j - idy < 1 ? 1 : j - idy, ly2 = j + idy > r->my ? r->my : j + idy;
        const int lx1 = i - idx < 1 ? 1 : i - idx, lx2 = i + idx > r->mx ? r->mx : i + idx;
        const float vol = (float)((lx2 - lx1 + 1) * (ly2 - ly1 + 1) * (lz2 - lz1 + 1));
        const size_t l = IDX(r, i, j, k);
        cx[l] = cx[l] / vol; cy[l] = cy[l] / vol;                                      /* :5458-5459 */
    }
    if (strncmp(name, "tdens", 5) && strncmp(name, "idens", 5) && strncmp(name, "hdens", 5) && strncmp(name, "ldens", 5))
        for (size_t l = 0; l < r->lot; l++) cx[l] = cy[l] != 0.f ? cx[l] / cy[l] : 0.f;   /* :5470-5477 */
}
void orc_meanq_fld_cur(orc_world *w, const char *totname)
{
    for (int rk = 0; rk < w->size0; rk++) meanq_accumulate(w->r[rk], totname);
    orc_exchange_current(w);
    for (int rk = 0; rk < w->size0; rk++) meanq_normalise(w->r[rk], totname);
}

/* ------------------------------------------------------------------------- */
/* output-side spectra: the per-rank part of save_spectrum, output.F90:380-633   */
/* gamma range (:440-455, before the allreduce of :458-463), then per species:    */
/* slice-mean flow velocity (:477-497, 561-578), lab-frame spectrum (:503-511)    */
/* and flow-rest-frame spectrum (:513-537) on nbins x-slices x gambins log bins.   */
/* The sums are returned as the ranks hold them BEFORE mpi_allreduce and before    */
/* the division by xgamma (:539-552).  The electron loops of the reference start   */
/* one slot early (i = maxhlf, a dead record, :562, 583); that slot is not read.   */
/* spec arrays are Fortran order (xbin fastest).                                   */
/* ------------------------------------------------------------------------- */
void orc_spectrum_gamma_range(const orc_rank *r, float *gammin, float *gammax)
{
    float lo = 1.f, hi = 1.f;
    for (int sp = 0; sp < 2; sp++) {
        const int first = sp ? r->maxhlf : 0, cnt = sp ? r->lecs : r->ions;
        for (int n = 0; n < cnt; n++) {
            const orc_particle *q = &r->p[first + n];
            const float gam = sqrtf(1.f + (q->u * q->u + q->v * q->v + q->w * q->w));
            if (gam > hi) hi = gam;
            if (gam < lo) lo = gam;
        }
    }
    *gammin = lo; *gammax = hi;
}
void orc_spectrum(const orc_rank *r, float gammin, float gammax, int mx0, float splitratio, int nbins, int gambins,
                  float *specp, float *spece, float *specpprime, float *speceprime)
{
    const int mxmin = 3, mxmax = mx0 - 2;
    const float dxslice = 1.f * (mxmax - mxmin) / nbins;
    gammin = gammin > 1.f + 1e-6f ? gammin : 1.f + 1e-6f;                              /* :465 */
    const float lg0 = log10f(gammin - 1.f);
    const float dgam = (log10f(gammax - 1.f) - lg0) / gambins;                          /* :466 */
    float *um = (float *)calloc(4 * (size_t)nbins, sizeof(float)), *vm = um + nbins, *wm = vm + nbins, *nd = wm + nbins;
    for (int sp = 0; sp < 2; sp++) {
        const int first = sp ? r->maxhlf : 0, cnt = sp ? r->lecs : r->ions;
        float *spec = sp ? spece : specp, *specr = sp ? speceprime : specpprime;
        memset(um, 0, 4 * (size_t)nbins * sizeof(float));
        memset(spec, 0, (size_t)nbins * gambins * sizeof(float)); memset(specr, 0, (size_t)nbins * gambins * sizeof(float));
        for (int n = 0; n < cnt; n++) {
            const orc_particle *q = &r->p[first + n];
            const int xbin = (int)((q->x + r->mxcum - mxmin) / dxslice + 1);
            if (xbin < 1 || xbin > nbins) continue;
            const float wgt = powf(splitratio, 1.f - (float)q->splitlev) * q->ch;
            const float gam = sqrtf(1.f + (q->u * q->u + q->v * q->v + q->w * q->w));
            nd[xbin - 1] += wgt; um[xbin - 1] += q->u / gam * wgt; vm[xbin - 1] += q->v / gam * wgt; wm[xbin - 1] += q->w / gam * wgt;
        }
        for (int b = 0; b < nbins; b++) { um[b] /= nd[b]; vm[b] /= nd[b]; wm[b] /= nd[b]; }
        for (int n = 0; n < cnt; n++) {
            const orc_particle *q = &r->p[first + n];
            const int xbin = (int)((q->x + r->mxcum - mxmin) / dxslice + 1);
            if (xbin < 1 || xbin > nbins) continue;
            const float wgt = powf(splitratio, 1.f - (float)q->splitlev) * q->ch;
            const float gam = sqrtf(1.f + (q->u * q->u + q->v * q->v + q->w * q->w));
            int gbin = (int)((log10f(gam - 1.f) - lg0) / dgam + 1);
            if (gbin >= 1 && gbin <= gambins) spec[(xbin - 1) + (size_t)nbins * (gbin - 1)] += wgt;
            const float vx = um[xbin - 1], vy = vm[xbin - 1], vz = wm[xbin - 1];
            const float vr = sqrtf(vx * vx + vy * vy + vz * vz), gvr = 1.f / sqrtf(1.f - vr * vr);
            const float up = -vx * gvr * gam + (1 + (gvr - 1) * vx * vx / (vr * vr)) * q->u + (gvr - 1) * vx * vy / (vr * vr) * q->v
                             + (gvr - 1) * vx * vz / (vr * vr) * q->w;
            const float vp = -vy * gvr * gam + (gvr - 1) * vx * vy / (vr * vr) * q->u + (1 + (gvr - 1) * vy * vy / (vr * vr)) * q->v
                             + (gvr - 1) * vy * vz / (vr * vr) * q->w;
            const float wp = -vz * gvr * gam + (gvr - 1) * vx * vz / (vr * vr) * q->u + (gvr - 1) * vy * vz / (vr * vr) * q->v
                             + (1 + (gvr - 1) * vz * vz / (vr * vr)) * q->w;
            const float gp = sqrtf(1.f + (up * up + vp * vp + wp * wp));
            gbin = (int)((log10f(gp - 1.f) - lg0) / dgam + 1);
            if (gbin >= 1 && gbin <= gambins) specr[(xbin - 1) + (size_t)nbins * (gbin - 1)] += wgt;
        }
    }
    free(um);
}

/* ------------------------------------------------------------------------- */
/* diagnostics used by the known-answer tests                                   */
/* ------------------------------------------------------------------------- */
/* node charge density with the deposit's own shape function: rho(i,j,k) = sum q S(i) S(j) S(k),
   the quantity whose change the Esirkepov scheme balances against div(cur). */
void orc_charge_density(const orc_rank *r, float *rho)
{
    memset(rho, 0, r->lot * sizeof(float));
    const int three = r->P.dim == 3;
    int order = r->P.order == 0 ? 1 : r->P.order;
    for (int sp = 0; sp < 2; sp++) {
        int first = sp ? r->maxhlf : 0, cnt = sp ? r->lecs : r->ions;
        float qs = sp ? r->P.qe : r->P.qi;
        for (int n = 0; n < cnt; n++) {
            const orc_particle *p = &r->p[first + n];
            float Sx[8], Sy[8], Sz[8]; int a, b;
            int i1 = (int)p->x, j1 = (int)p->y, k1 = (int)p->z;
            orc_shape(order, p->x - i1, 0, Sx, &a, &b);
            orc_shape(order, p->y - j1, 0, Sy, &a, &b);
            if (three) orc_shape(order, p->z - k1, 0, Sz, &a, &b); else { for (int s = 0; s < 8; s++) Sz[s] = 0; Sz[3] = 1; k1 = 1; }
            float q = p->ch * qs;
            for (int c = 1; c <= 6; c++) { if (Sz[c] == 0.f) continue;
                for (int bq = 1; bq <= 6; bq++) { if (Sy[bq] == 0.f) continue;
                    for (int aq = 1; aq <= 6; aq++) { if (Sx[aq] == 0.f) continue;
                        int i = i1 - 3 + aq, j = j1 - 3 + bq, k = k1 - 3 + c;
                        if (i < 1 || i > r->mx || j < 1 || j > r->my || k < 1 || k > r->mz) continue;
                        rho[IDX(r, i, j, k)] += q * Sx[aq] * Sy[bq] * Sz[c];
                    } } }
        }
    }
}
double orc_sum_array(const orc_rank *r, int which)
{
    double s = 0; for (size_t l = 0; l < r->lot; l++) s += r->f[which][l]; return s;
}
