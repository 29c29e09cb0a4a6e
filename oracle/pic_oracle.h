/*
 * pic_oracle.h -- CPU restatement of the TRISTAN-MP per-timestep PIC hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (tristan_mp_pu_master_densdecomp_b200/,
 * libtristan_gpu.so) may include, link or call this.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
 *
 * PARITY PINNED AGAINST THE REFERENCE'S SOURCE TEXT, not against a compiled
 * reference: the reference ships no golden vectors or tests and cannot be
 * compiled in this image or on the GPU box (no Fortran compiler, no MPI).  Its
 * hot-path subroutines -- up to `mainloop` itself, on several ranks -- are
 * executed from their text by the Fortran-subset interpreter
 * tests/golden/f90run.py; tests/test_ref_golden.py holds this restatement
 * BIT-EXACT against those outputs (tests/golden/ref_*.npz; 147 cases), every
 * reference routine restated here included (DESIGN.md section 2 has the table).
 * Every function cites the reference file:line it follows (relative to the
 * reference checkout).
 *
 * Conventions: arrays are Fortran column-major (mx,my,mz), addressed here with
 * 1-based (i,j,k) through ORC_IDX so index arithmetic reads like the reference.
 * All arithmetic is fp32 unless the reference uses fp64 (RNG seed).
 */
#ifndef PIC_ORACLE_H
#define PIC_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* code/particles.F90:51-55 -- `type particle`, sequence, 40 bytes */
typedef struct {
    float x, y, z, u, v, w, ch;
    int32_t ind, proc, splitlev;
} orc_particle;

/* quirk switches (SURVEY.md section 8, Q-list); default ORC_Q_REFERENCE */
enum {
    ORC_Q1_MOVER2_RANGE = 1 << 0, /* mover_2ord uses the dual-branch loop bounds for primal weights */
    ORC_Q2_BXBY_NO_KAVG = 1 << 1, /* 3D bx_p, by_p omit the k average */
    ORC_Q4_DEPOSIT_CARRY = 1 << 2,/* deposit prefix carries are not reset at row/plane ends */
    ORC_Q5_FILTER_CURZ_J = 1 << 3,/* filter1 curz j-range runs to my-nghost/2+1 */
    ORC_Q_REFERENCE = 0xF
};

enum { ORC_EX = 0, ORC_EY, ORC_EZ, ORC_BX, ORC_BY, ORC_BZ, ORC_CURX, ORC_CURY, ORC_CURZ, ORC_NARR };

typedef struct {
    int dim;                 /* 2 (-DtwoD) or 3 */
    int order;               /* 0 = -Dzzag, 1/2/3 = -Ddd1/2/3 */
    int mx0, my0, mz0;       /* interior cells requested in the input file (<grid> mx0,my0,mz0) */
    int sizex, sizey, sizez; /* rank grid, communications.F90:163-181 */
    float c, corr;
    int ntimes;
    int filter_kind;         /* 1 = apply_filter1_opt, 2 = apply_filter2_opt */
    int periodicx, periodicy, periodicz;
    float qi, qe, qmi, qme;
    int maxptl;              /* per-rank particle capacity (ions in [0,maxhlf), electrons in [maxhlf,..)) */
    int buffsize;            /* outbox capacity per direction */
    int quirks;
    int pusher;              /* 0 Boris, 1 Vay (-Dvay) */
    int external_fields;     /* constant external field model of get_external_fields */
    float ext[6];            /* ex,ey,ez,bx,by,bz */
    int highorder;           /* <algorithm> highorder: 1 = 4th-order `_42` field solver (fields.F90:1039-1361) */
    int wall_i2;             /* `wall` clamp of the _42 solver's x range, int(xinject2)+10 (fields.F90:1062-1065); 0 = no wall */
} orc_params;

typedef struct {
    orc_particle *p;
    int nion, nlec;
} orc_box;

typedef struct orc_rank {
    orc_params P;
    int rank, size0;
    int nghost, nghostz;
    int mx, my, mz;
    int mxcum, mycum, mzcum;
    int iy, iz;                   /* strides; iz = 0 in 2D (fields.F90:330-338) */
    size_t lot;
    float *f[ORC_NARR];           /* ex..bz, curx..curz */
    float *temp;
    orc_particle *p;
    int32_t *pind;
    int ions, lecs, maxhlf;
    float x1in, x2in, y1in, y2in, z1in, z2in;
    /* out/in boxes: 0 minus(-x) 1 plus(+x) 2 lft(-y) 3 rgt(+y) 4 dwn(-z) 5 up(+z) */
    orc_box out[6], in[6];
    double dseed;
    int totalpartnum;
    const int *mxl, *myl, *mzl;   /* owned by the world */
    int lap;
    int mx0g;                     /* global mx0 (ghosts included), for the user hooks */
} orc_rank;

typedef struct {
    orc_params P;
    int size0;
    orc_rank **r;
    int *mxl, *myl, *mzl;
    int mx0g, my0g, mz0g;         /* global sizes incl. ghosts (the reference's mx0,my0,mz0 after read_input_grid) */
    int lap;
} orc_world;

orc_world *orc_world_create(const orc_params *P);
void orc_world_destroy(orc_world *w);
size_t orc_sizeof_particle(void);

/* accessors for ctypes */
orc_rank *orc_world_rank(orc_world *w, int r);
float *orc_rank_array(orc_rank *r, int which);
orc_particle *orc_rank_particles(orc_rank *r);
void orc_rank_dims(const orc_rank *r, int *out /* mx,my,mz,nghost,nghostz,mxcum,mycum,mzcum,maxhlf */);
void orc_rank_counts(const orc_rank *r, int *ions, int *lecs);
void orc_rank_set_counts(orc_rank *r, int ions, int lecs);
orc_particle *orc_rank_box(orc_rank *r, int which /* 0 out, 1 in */, int dir, int *nion, int *nlec);
void orc_rank_box_set_counts(orc_rank *r, int which, int dir, int nion, int nlec);

/* --- shape weights (Appendix A.1) --- */
void orc_shape(int order, float d, int shift, float S[8], int *smin, int *smax);

/* --- per-rank kernels --- */
void orc_advance_b_halfstep(orc_rank *r);
void orc_advance_e_fullstep(orc_rank *r);
void orc_reset_currents(orc_rank *r);
void orc_add_current(orc_rank *r);
void orc_move_particles(orc_rank *r);
void orc_mover_range(orc_rank *r, int n1, int n2, float qm);
void orc_deposit_one(orc_rank *r, float x2, float y2, float z2, float x1, float y1, float z1, float q);
void orc_deposit_currents_only(orc_rank *r);      /* loop A of deposit_particles only */
void orc_deposit_particles(orc_rank *r);
void orc_inject_others(orc_rank *r);
void orc_reorder_particles(orc_rank *r);
void orc_filter1_pass(orc_rank *r);               /* one 9/27-point pass, ghosts assumed fresh */
void orc_filter2_line(float *line, int len, int ntimes); /* ntimes fixed-end 1-2-1 passes on an extended line */
void orc_filter2_rank(orc_rank *r, int c, int axis, const float *glo, const float *ghi);
void orc_filter2_send_box(const orc_rank *r, int axis, int what, int lo[3], int hi[3]);

/* box primitives the exchanges are built from (also used by the gloo tests) */
void orc_box_get(const orc_rank *r, int which, const int lo[3], const int hi[3], float *buf);
void orc_box_put(orc_rank *r, int which, const int lo[3], const int hi[3], const float *buf);
void orc_box_add(orc_rank *r, int which, const int lo[3], const int hi[3], const float *buf);

/* neighbour ranks: dir 0..5 as in out[] */
int orc_neighbour(const orc_rank *r, int dir);

/* --- world-level (exchange) routines --- */
void orc_bc_fields(orc_world *w, int first /* ORC_EX or ORC_BX or ORC_CURX */);  /* bc_e1 / bc_b1 */
void orc_exchange_current(orc_world *w);
void orc_exchange_particles(orc_world *w);
void orc_apply_filter1(orc_world *w);
void orc_apply_filter2(orc_world *w);
void orc_apply_filter(orc_world *w);
/* moments of the particle distribution into curx (cury = weight): meanq_fld_cur(totname), output.F90:5229-5486 */
void orc_meanq_fld_cur(orc_world *w, const char *totname);
/* per-rank part of save_spectrum (output.F90:380-633): gamma range, then the four nbins x gambins histograms (xbin fastest) */
void orc_spectrum_gamma_range(const orc_rank *r, float *gammin, float *gammax);
void orc_spectrum(const orc_rank *r, float gammin, float gammax, int mx0, float splitratio, int nbins, int gambins,
                  float *specp, float *spece, float *specpprime, float *speceprime);
void orc_step(orc_world *w);                      /* one lap of tristanmainloop.F90:107-344 */
void orc_step_phase(orc_world *w, int phase);     /* single named phase, for A/B tests */

/* --- seeded loader --- */
float orc_random(double *dseed);
void orc_init_maxw_table(int dim, int pcosthmult, float delgam, float *gamma_table, float *pdf_table);
void orc_maxwell_dist(int dim, int pcosthmult, float sigma, float gamma0, float cd, double *dseed,
                      float *u, float *v, float *w, const float *gamma_table, const float *pdf_table);
void orc_inject_plasma_region(orc_rank *r, float x1, float x2, float y1, float y2, float z1, float z2,
                              float ppc, float gamma_drift_in, float delgam_i, float delgam_e,
                              float weight, int direction, int pcosthmult, float sigma);
/* plane source: inject_from_wall (particles.F90:2439-2538) and the shock problem's per-lap injector (user/user_shock.F90:303-331) */
void orc_inject_from_wall(orc_rank *r, float x1, float x2, float y1, float y2, float z1, float z2, float ppc, float gamma_drift,
                          float delgam_i, float delgam_e, float wall_speed, float weight, int pcosthmult, float sigma);
void orc_inject_particles_shock(orc_world *w, float ppc0, float gamma0_in, float delgam, float me, float mi,
                                float temperature_ratio, int pcosthmult, float sigma);
/* problem setups: user/user_weibel.F90:255-313, user/user_twostream.F90:226-270 */
void orc_charge_normalisation(orc_params *P, float ppc0, float c_omp, float gamma0, float me, float mi);
void orc_init_weibel(orc_world *w, float ppc0, float gamma0_in, float delgam, float me, float mi,
                     float temperature_ratio, int distr_dim);
void orc_init_twostream(orc_world *w, float ppc0, float gamma0_in, float delgam, float me, float mi,
                        float temperature_ratio);
/* fast synthetic loader for benchmarks: uniform positions, drifting Maxwellian-ish momenta */
void orc_init_uniform(orc_world *w, float ppc0, float beta_drift, float uth, uint64_t seed);

/* --- shock-problem user hooks (user/user_shock.F90) --- */
/* radiation boundary `surface` of bc_b2 / bc_e2: fieldboundaries.F90:274-295, 403-426, 493-606 (per rank; the
   ghost refresh that completes bc_b2 / bc_e2 is orc_bc_fields) */
int orc_range42_ok(const orc_params *P);   /* 0: highorder = 1 with index ranges that are out of bounds in the reference */
void orc_surface_b(orc_rank *r);
void orc_edges(orc_rank *r, int which);   /* 0 pre_bc_b, 1 post_bc_b, 2 pre_bc_e, 3 post_bc_e: preledge / postedge (fieldboundaries.F90:2200-2505) */
void orc_surface_e(orc_rank *r);
void orc_field_bc_shock(orc_rank *r, float leftwall, float binit, float btheta, float bphi, float beta);   /* :342-373 */
void orc_particle_bc_wall(orc_rank *r, float leftwall);                                                     /* :377-457 */
void orc_step_shock(orc_world *w, float leftwall, float binit, float btheta, float bphi, float beta);       /* lap with the hooks */

/* diagnostics */
void orc_charge_density(const orc_rank *r, float *rho /* lot floats */);
double orc_sum_array(const orc_rank *r, int which);

#ifdef __cplusplus
}
#endif
#endif
