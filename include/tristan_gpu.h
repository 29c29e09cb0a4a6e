/*
 * tristan_gpu.h -- C ABI of libtristan_gpu.so: the sm_100a implementation of TRISTAN-MP's
 * per-timestep PIC hot path (mover, Esirkepov deposit, Yee half/full steps, current filter,
 * ghost/current/particle exchanges).
 *
 * The reference has no plugin API for this path: mainloop() calls argument-less module
 * procedures that work on module-global arrays (code/tristanmainloop.F90:107-344).  Each entry
 * point below replaces one of those procedures (file:line given per function; all paths are
 * relative to the reference checkout) so that `call move_particles()` becomes
 * `ierr = tgpu_move_particles(h)` through the ISO_C_BINDING interface shown in INTEGRATION.md.
 *
 * Conventions
 *   - every function returns 0 on success, a TGPU_E* code otherwise; tgpu_last_error() gives text.
 *     (The reference prints and `stop`s, e.g. code/particles.F90:362-370; the Fortran wrapper does
 *     `if (ierr/=0) stop`.)
 *   - arrays are Fortran column-major (mx,my,mz) fp32, exactly the reference's allocatables
 *     (code/fields.F90:81-88); particles are the 40-byte `type particle` (code/particles.F90:51-55).
 *   - one context per MPI rank / GPU; calls are synchronous with respect to host-visible outputs.
 *   - no CPU fallback: every call fails with TGPU_ECUDA if no sm_100 device is usable.
 */
#ifndef TRISTAN_GPU_H
#define TRISTAN_GPU_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    TGPU_OK = 0,
    TGPU_EINVAL = 1,     /* bad argument / unsupported configuration */
    TGPU_ECUDA = 2,      /* CUDA runtime error or no device */
    TGPU_EOVERFLOW = 3,  /* particle or outbox capacity exceeded (check_overflow, particles.F90:362-385) */
    TGPU_ENCCL = 4,      /* NCCL error / communicator not initialised */
    TGPU_ESTATE = 5      /* call sequence error */
};

/* quirk switches: reproduce reference defects bit for bit */
enum {
    TGPU_Q1_MOVER2_RANGE = 1 << 0,
    TGPU_Q2_BXBY_NO_KAVG = 1 << 1,
    TGPU_Q4_DEPOSIT_CARRY = 1 << 2,   /* accepted, no effect on the GPU: the carries are 0 in exact arithmetic */
    TGPU_Q5_FILTER_CURZ_J = 1 << 3,
    TGPU_Q_REFERENCE = 0xF
};

/* 40-byte particle record: code/particles.F90:51-55 */
typedef struct tgpu_particle {
    float x, y, z, u, v, w, ch;
    int32_t ind, proc, splitlev;
} tgpu_particle;

/* Everything the reference keeps in module globals that the kernels need
 * (code/fields.F90:68-79, code/particles.F90:219-251,
 * code/communications.F90:105-106, code/fieldboundaries.F90:86-88). */
typedef struct tgpu_params {
    int32_t dim;                 /* 2 (-DtwoD) or 3 */
    int32_t order;               /* 0 = -Dzzag, 1/2/3 = -Ddd1/-Ddd2/-Ddd3 */
    int32_t mx, my, mz;          /* this rank's array sizes, ghosts included (mz = 1 in 2D) */
    int32_t nghost, nghostz;     /* 5 or 7 (fields.F90:166-184) */
    float c, corr;               /* <time> c ; <algorithm> Corr */
    float qi, qe, qmi, qme;      /* particles.F90:224-233 */
    int32_t ntimes;              /* <algorithm> ntimes */
    int32_t filter_kind;         /* 1 = apply_filter1_opt (filter.F90), 2 = apply_filter2_opt (optimized_filters.F90) */
    int32_t periodicx, periodicy, periodicz;
    float x1in, x2in, y1in, y2in, z1in, z2in;   /* particles.F90:339-344, global coordinates */
    int32_t rank, sizex, sizey, sizez;          /* communications.F90:163-181 */
    int32_t mxcum, mycum, mzcum;                /* fields.F90:316-328 */
    const int32_t *mxl, *myl, *mzl;             /* per-rank sizes, size0 entries each (may be NULL when size0 == 1) */
    int32_t maxptl;              /* per-rank particle capacity; maxhlf = maxptl/2 per species */
    int32_t buffsize;            /* migration outbox capacity per direction (particles.F90:309) */
    int32_t quirks;              /* TGPU_Q_* mask; TGPU_Q_REFERENCE reproduces the reference */
    int32_t pusher;              /* 0 Boris, 1 Vay (-Dvay) */
    int32_t external_fields;     /* constant external field model of get_external_fields */
    float ext[6];                /* ex,ey,ez,bx,by,bz */
    int32_t device;              /* CUDA device ordinal, or -1 for rank % device_count */
    int32_t highorder;           /* <algorithm> highorder: 1 = 4th-order `_42` field solver (fields.F90:1039-1361, dispatch :1407-1440) */
    int32_t wall_i2;             /* `wall` clamp of the _42 solver's x range, int(xinject2)+10 (fields.F90:1062-1065); 0 = no wall */
} tgpu_params;

typedef struct tgpu_ctx tgpu_ctx;

/* ---- lifetime ---------------------------------------------------------------------------- */
/* after initialize() (code/initialize.F90:113-175): allocates device state for this rank */
int tgpu_init(const tgpu_params *p, tgpu_ctx **out);
int tgpu_finalize(tgpu_ctx *h);
const char *tgpu_last_error(void);
int tgpu_device_count(void);                     /* 0 when no usable GPU */

/* ---- topology helpers (host only; usable without a GPU) ------------------------------------ */
/* neighbour rank in direction dir = 0:-x 1:+x 2:-y 3:+y 4:-z 5:+z
 * (fieldboundaries.F90:1170-1171, 1311-1314; particles.F90:1892-1895) */
int tgpu_neighbour(int rank, int sizex, int sizey, int sizez, int dir);
/* slab sizes and offsets of `rank` for a global interior mx0 x my0 x mz0 (fields.F90:259-328);
 * out = {mx,my,mz,mxcum,mycum,mzcum} */
int tgpu_decompose(int dim, int order, int mx0, int my0, int mz0, int sizex, int sizey, int sizez,
                   int rank, int32_t out[6]);
int tgpu_ghost_width(int dim, int order, int32_t *nghost, int32_t *nghostz);   /* fields.F90:166-184 */

/* ---- communicator (NCCL over NVLink; replaces mpif.h SendRecv, communications.F90:127-161) -- */
int tgpu_comm_unique_id(uint8_t id[128]);         /* rank 0 creates, host broadcasts (MPI_Bcast / torch.distributed) */
int tgpu_comm_init(tgpu_ctx *h, const uint8_t id[128]);

/* ---- state transfer (mirror / resident modes, SURVEY.md 8b "Ownership") --------------------- */
int tgpu_fields_h2d(tgpu_ctx *h, const float *ex, const float *ey, const float *ez,
                    const float *bx, const float *by, const float *bz);
int tgpu_fields_d2h(tgpu_ctx *h, float *ex, float *ey, float *ez, float *bx, float *by, float *bz);
int tgpu_currents_h2d(tgpu_ctx *h, const float *curx, const float *cury, const float *curz);
int tgpu_currents_d2h(tgpu_ctx *h, float *curx, float *cury, float *curz);
/* p is the reference's p(1:maxptl): ions at [0,ions), electrons at [maxhlf, maxhlf+lecs) */
int tgpu_particles_h2d(tgpu_ctx *h, const tgpu_particle *p, int ions, int lecs);
int tgpu_particles_d2h(tgpu_ctx *h, tgpu_particle *p, int *ions, int *lecs);
int tgpu_counts(tgpu_ctx *h, int *ions, int *lecs);
/* host injector output (inject_particles_user, user/user_shock.F90:303-331): n_ion ions then n_lec electrons */
int tgpu_append_particles(tgpu_ctx *h, const tgpu_particle *p, int n_ion, int n_lec);

/* ---- field solver: code/fields.F90 ---------------------------------------------------------- */
int tgpu_advance_b_halfstep(tgpu_ctx *h);         /* fields.F90:586-728 */
int tgpu_advance_e_fullstep(tgpu_ctx *h);         /* fields.F90:739-870 */
int tgpu_reset_currents(tgpu_ctx *h);             /* fields.F90:499-506 */
int tgpu_add_current(tgpu_ctx *h);                /* fields.F90:1372-1395 */

/* ---- field boundaries: code/fieldboundaries.F90 ---------------------------------------------- */
int tgpu_bc_b1(tgpu_ctx *h);                      /* fieldboundaries.F90:181-263 */
int tgpu_bc_e1(tgpu_ctx *h);                      /* fieldboundaries.F90:306-392 */
int tgpu_bc_b2(tgpu_ctx *h);                      /* :274-295, 493-606: radiation `surface` on radiating axes, then bc_b1 */
int tgpu_bc_e2(tgpu_ctx *h);                      /* :403-426: `surface` (low faces, E<->B), then bc_e1 */
int tgpu_pre_bc_b(tgpu_ctx *h);                   /* :114-131: preledge x3 + bc_b1 when all three axes of a 3D box radiate; else no-op */
int tgpu_post_bc_b(tgpu_ctx *h);                  /* :437-452: postedge x3 + bc_b1, same condition (bodies :2200-2505) */
int tgpu_pre_bc_e(tgpu_ctx *h);                   /* :148-163: preledge x3 (E <-> B, mirrored strides) + bc_e1 */
int tgpu_post_bc_e(tgpu_ctx *h);                  /* :463-482: postedge x3 + bc_e1 */
int tgpu_exchange_current(tgpu_ctx *h);           /* fieldboundaries.F90:1768-2189 */

/* ---- filter: code/filter.F90, code/optimized_filters.F90 ------------------------------------- */
int tgpu_apply_filter(tgpu_ctx *h);               /* dispatch of tristanmainloop.F90:213-229 on filter_kind */
int tgpu_apply_filter1_opt(tgpu_ctx *h);          /* filter.F90:8-221 */
int tgpu_apply_filter2_opt(tgpu_ctx *h);          /* optimized_filters.F90:9-227 */

/* ---- particles: code/particles_movedeposit.F90, code/particles.F90 --------------------------- */
int tgpu_move_particles(tgpu_ctx *h);             /* particles_movedeposit.F90:63-88 -> mover / mover_{1,2,3}ord */
int tgpu_deposit_particles(tgpu_ctx *h);          /* particles_movedeposit.F90:1281-2051 -> zigzag / densdecomp_{1,2,3}ord */
int tgpu_exchange_particles(tgpu_ctx *h);         /* particles.F90:1865-2116 + inject_others :1368-1852, both rounds */
int tgpu_inject_others(tgpu_ctx *h);              /* no-op: folded into tgpu_exchange_particles; kept for call-list parity */
int tgpu_reorder_particles(tgpu_ctx *h);          /* particles.F90:394-497 (unconditional here; the lap%10 test stays in the caller) */

/* ---- user hooks of the shock problem (user/user_shock.F90), SURVEY.md 8(f) row 1 ---------------------------- */
/* field_bc_user, user_shock.F90:342-373: conductor behind the left wall (ey = ez = 0 for x < leftwall-10) and upstream
 * fields clamped on the last three x cells.  beta is the upstream drift v/c (particles.F90:222). */
int tgpu_field_bc_user_shock(tgpu_ctx *h, float leftwall, float binit, float btheta, float bphi, float beta);
/* particle_bc_user, user_shock.F90:377-457: specular wall at global x = leftwall with the two zigzag deposits */
int tgpu_particle_bc_user_wall(tgpu_ctx *h, float leftwall);
/* make tgpu_step call the two hooks at mainloop's hook points (tristanmainloop.F90:146,160,166,177,243);
 * params = {leftwall, binit, btheta, bphi, beta}; kind 0 = none, 1 = shock */
int tgpu_set_user_hooks(tgpu_ctx *h, int kind, const float params[5]);

/* ---- mirror mode, whole lap --------------------------------------------------------------------- */
/* One lap with the state owned by the host (what a call-for-call GPU build of mainloop does per lap): fields and particles
 * are read from, and written back to, the host arrays of code/fields.F90:81-88 and code/particles.F90:100 (p = the full
 * array, ions at p[0..ions), electrons at p[maxhlf..maxhlf+lecs)).  On one rank with all axes periodic the 40-byte records
 * are streamed through the fused mover while both PCIe directions are busy; otherwise it is fields_h2d + particles_h2d +
 * tgpu_step + particles_d2h + fields_d2h.  Pinned host memory is needed for the copies to overlap. */
int tgpu_step_mirror(tgpu_ctx *h, float *ex, float *ey, float *ez, float *bx, float *by, float *bz,
                     tgpu_particle *p, int *ions, int *lecs);

/* ---- output-side reductions (device-resident state) ---------------- */
/* meanq_fld_cur(totname), code/output.F90:5229-5486: totname = 'tdens' 'idens' 'hdens' 'ldens' 'btden' 'biden',
 * '[tei]bet[xyz]', '[ti]mom[xyz]', 'eener' 'iener', '[ei]et[xyz]2'.  As in the reference the result is left in curx
 * (cury = weight, curz = 0): read it with tgpu_currents_d2h.  Destroys the currents, like the reference (output laps only). */
int tgpu_meanq_fld_cur(tgpu_ctx *h, const char *totname);
/* per-rank part of save_spectrum, code/output.F90:380-633.  tgpu_spectrum_gamma_range: local min / max of gamma (both start
 * at 1, :424-455); the host allreduces them (:458-463) and passes the global values to tgpu_spectrum, which fills the four
 * nbins x gambins histograms (Fortran order, x-slice fastest; nbins = max((mx0-5)/100, 1), gambins = 200 in the reference):
 * lab-frame ions / electrons and flow-rest-frame ions / electrons, as this rank's sums before mpi_allreduce and before the
 * division by xgamma (:539-552).  mx0 = global x size incl. ghosts; splitratio as in the input file. */
int tgpu_spectrum_gamma_range(tgpu_ctx *h, float *gammin, float *gammax);
int tgpu_spectrum(tgpu_ctx *h, float gammin, float gammax, int mx0, float splitratio, int nbins, int gambins,
                  float *specp, float *spece, float *specprest, float *specerest);
/* the prtl.tot sub-sample, code/output.F90:3526-3551: every particle with modulo(ind/2, stride) == 0, compacted on the
 * device; ions to out[0 .. *n_ion), electrons to out[capacity .. capacity + *n_lec).  Positions are rank-local (the host adds
 * mxcum/mycum/mzcum as output.F90 does).  TGPU_EOVERFLOW if a species selects more than `capacity`. */
int tgpu_select_particles(tgpu_ctx *h, int stride, tgpu_particle *out, int capacity, int *n_ion, int *n_lec);

/* ---- whole lap, resident mode: tristanmainloop.F90:107-344 with Appendix-B de-duplication ---- */
int tgpu_step(tgpu_ctx *h, int nlaps);

/* ---- instrumentation (print_timers, communications.F90:242-326) ------------------------------ */
/* device milliseconds accumulated per phase since the last reset; out has TGPU_NPHASE entries */
enum { TGPU_PH_FIELDS = 0, TGPU_PH_MOVER, TGPU_PH_DEPOSIT, TGPU_PH_PEXCH, TGPU_PH_CUREXCH, TGPU_PH_FILTER,
       TGPU_PH_SORT, TGPU_PH_BC, TGPU_NPHASE };
int tgpu_timers(tgpu_ctx *h, double *out_ms, int reset);
/* number of kernels this library has launched since init (all streams) */
int64_t tgpu_launch_count(tgpu_ctx *h);
/* the CUDA stream all kernels of this context are launched on (cudaStream_t as void*) */
void *tgpu_stream(tgpu_ctx *h);
/* options: "fused" (0 = generic per-particle kernels, 1 = cell-run fused kernels where available), "lazy", "overlap",
   "timing", "fast_push" (SFU reciprocals in the cell-run Boris push), "peer" (before tgpu_comm_init: 1 = field halos over
   cudaIpc peer memory, 0 = NCCL send/recv) */
int tgpu_set_option(tgpu_ctx *h, const char *name, int value);
/* which transport the field-side halo exchanges use after tgpu_comm_init: 1 = peer memory (the neighbours' arrays are
   mapped with cudaIpc and read over NVLink by the halo kernels), 0 = NCCL send/recv (replaces the MPI_SendRecv pairs of
   fieldboundaries.F90:1179-1204, 1319-1344, 1652-1677, 1990-2185 either way) */
int tgpu_halo_transport(tgpu_ctx *h);

#ifdef __cplusplus
}
#endif
#endif
