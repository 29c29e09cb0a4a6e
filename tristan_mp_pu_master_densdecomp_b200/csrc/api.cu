// C ABI of libtristan_gpu.so (include/tristan_gpu.h): context lifetime, state transfer, and the one-to-one
// replacements of the procedures mainloop() calls (code/tristanmainloop.F90:107-344).
#include <cub/device/device_scan.cuh>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "tgpu_internal.h"

static thread_local std::string g_err;
void tgpu_set_error(const std::string &s) { g_err = s; }
extern "C" const char *tgpu_last_error(void) { return g_err.c_str(); }

int comm_destroy(tgpu_ctx *h);
int prt_move_generic(tgpu_ctx *h);
int prt_deposit_generic(tgpu_ctx *h);
int cellrun_supported(const tgpu_ctx *h);
int cellrun_move_deposit(tgpu_ctx *h);      // fused gather + push + deposit into shadow[]
int cellrun_deposit(tgpu_ctx *h);           // deposit only, into cur
int cellrun3_supported(const tgpu_ctx *h);
int cellrun3_deposit(tgpu_ctx *h);          // 3rd-order deposit only, into cur
int cellrun3_move_deposit(tgpu_ctx *h);     // fused 3rd-order gather + push + deposit into shadow[]
static int fused_mover_supported(const tgpu_ctx *h) { return cellrun_supported(h) || (cellrun3_supported(h) && h->G.lot < (1ll << 30)); }

extern "C" int tgpu_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

template <typename T> static int dalloc(T **p, size_t n)
{
    CK(cudaMalloc((void **)p, (n ? n : 1) * sizeof(T)));
    CK(cudaMemset(*p, 0, (n ? n : 1) * sizeof(T)));
    return 0;
}
static int alloc_species(Species &S, size_t n)
{
    int rc = 0;
    rc |= dalloc(&S.x, n); rc |= dalloc(&S.y, n); rc |= dalloc(&S.z, n); rc |= dalloc(&S.u, n); rc |= dalloc(&S.v, n);
    rc |= dalloc(&S.w, n); rc |= dalloc(&S.ch, n); rc |= dalloc(&S.ind, n); rc |= dalloc(&S.tag, n);
    S.n = 0;
    return rc ? TGPU_ECUDA : 0;
}
static void free_species(Species &S)
{
    cudaFree(S.x); cudaFree(S.y); cudaFree(S.z); cudaFree(S.u); cudaFree(S.v); cudaFree(S.w); cudaFree(S.ch);
    cudaFree(S.ind); cudaFree(S.tag);
}

static int init_impl(tgpu_ctx *h, const tgpu_params *p, int ndev);

extern "C" int tgpu_init(const tgpu_params *p, tgpu_ctx **out)
{
    if (!p || !out) { tgpu_set_error("null argument"); return TGPU_EINVAL; }
    *out = nullptr;
    if ((p->dim != 2 && p->dim != 3) || p->order < 0 || p->order > 3) { tgpu_set_error("dim must be 2|3, order 0..3"); return TGPU_EINVAL; }
    // at least one interior cell per axis (user/input.twostream runs my0 = 2 under nghost = 7); the deep-halo filter needs
    // ntimes interior layers (tristanmainloop.F90:217-224 falls back to filter1 otherwise -- here the caller chooses)
    if (p->mx < p->nghost + 1 || p->my < p->nghost + 1 || (p->dim == 3 && p->mz < p->nghostz + 1)) { tgpu_set_error("grid smaller than its ghost zones"); return TGPU_EINVAL; }
    if (p->dim == 2 && p->mz != 1) { tgpu_set_error("2D needs mz = 1 (fields.F90:228-232)"); return TGPU_EINVAL; }
    if (p->dim == 3 && p->sizex != 1) { tgpu_set_error("3D never splits x (communications.F90:176-181)"); return TGPU_EINVAL; }
    if (p->highorder && ((p->dim == 3 && !p->periodicz) || (!p->periodicy && p->sizex * p->sizey * (p->dim == 3 ? p->sizez : 1) != 1))) {
        // fields.F90:1071-1079, 1092-1101, 1262-1277: those index ranges make the reference itself read outside its arrays
        tgpu_set_error("highorder = 1 with open z, or open y on more than one rank: index ranges undefined in the reference"); return TGPU_EINVAL;
    }
    if (p->sizex < 1 || p->sizey < 1 || p->sizez < 1 || p->maxptl < 2 || p->c <= 0.f || p->c >= 0.5f) { tgpu_set_error("bad sizes / c (need 0 < c < 0.5)"); return TGPU_EINVAL; }
    int ndev = tgpu_device_count();
    if (ndev <= 0) { tgpu_set_error("no CUDA device: libtristan_gpu has no CPU fallback"); return TGPU_ECUDA; }
    const int size0 = p->sizex * p->sizey * (p->dim == 3 ? p->sizez : 1);
    if (size0 > 1 && (!p->mxl || !p->myl || (p->dim == 3 && !p->mzl))) { tgpu_set_error("mxl/myl/mzl required when size0 > 1"); return TGPU_EINVAL; }
    if (size0 > 1 && p->buffsize < 1) { tgpu_set_error("buffsize must be positive when size0 > 1"); return TGPU_EINVAL; }
    tgpu_ctx *h = new tgpu_ctx();        // value-initialised: every pointer, stream and event starts out null
    const int rc = init_impl(h, p, ndev);
    if (rc) { const std::string why = g_err; tgpu_finalize(h); tgpu_set_error(why); return rc; }   // one cleanup path for every failure
    *out = h;
    return 0;
}

static int init_impl(tgpu_ctx *h, const tgpu_params *p, int ndev)
{
    h->P = *p;
    h->size0 = p->sizex * p->sizey * (p->dim == 3 ? p->sizez : 1);
    if (p->dim == 2) h->P.sizez = 1;
    for (int r = 0; r < h->size0; r++) {
        h->mxl.push_back(p->mxl ? p->mxl[r] : p->mx); h->myl.push_back(p->myl ? p->myl[r] : p->my);
        h->mzl.push_back(p->mzl ? p->mzl[r] : p->mz);
    }
    h->P.mxl = h->P.myl = h->P.mzl = nullptr;
    h->device = p->device >= 0 ? p->device : p->rank % ndev;
    CK(cudaSetDevice(h->device));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, h->device));
    if (prop.major < 10) { tgpu_set_error(std::string("device is not sm_100 class: ") + prop.name); return TGPU_ECUDA; }
    // the field kernels are short and latency-bound, the particle scatter is long and HBM-bound: giving the field stream
    // the higher priority lets its CTAs slip in between the scatter's instead of queueing behind them
    int prio_lo = 0, prio_hi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    CK(cudaStreamCreateWithPriority(&h->stream_main, cudaStreamNonBlocking, prio_hi));
    CK(cudaStreamCreateWithPriority(&h->stream_prt, cudaStreamNonBlocking, prio_lo));
    h->stream = h->stream_main;
    CK(cudaEventCreate(&h->ev0)); CK(cudaEventCreate(&h->ev1));
    CK(cudaEventCreateWithFlags(&h->ev_move, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&h->ev_prt, cudaEventDisableTiming));
    h->opt_lazy = getenv("TGPU_LAZY") ? atoi(getenv("TGPU_LAZY")) : 1;
    for (int b = 0; b < 2; b++) { CK(cudaEventCreateWithFlags(&h->ev_stage_full[b], cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&h->ev_stage_free[b], cudaEventDisableTiming)); }
    for (int b = 0; b < 2; b++) { CK(cudaEventCreateWithFlags(&h->ev_out_full[b], cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&h->ev_out_free[b], cudaEventDisableTiming)); }
    CK(cudaStreamCreateWithFlags(&h->stream_d2h, cudaStreamNonBlocking));
    h->presort = 0;
    h->prt_pending = 0; h->opt_overlap = getenv("TGPU_OVERLAP") ? atoi(getenv("TGPU_OVERLAP")) : 1; h->nccl_main = h->nccl_prt = nullptr;
    h->maxhlf = p->maxptl / 2;
    DevGeom &G = h->G;
    G.dim = p->dim; G.order = p->order; G.mx = p->mx; G.my = p->my; G.mz = p->mz;
    G.nghost = p->nghost; G.nghostz = p->nghostz; G.g = p->nghost / 2; G.gz = p->nghostz / 2;
    G.lot = (long long)p->mx * p->my * p->mz;
    G.rowblk = p->dim == 3 ? 1 : 0; G.nbj = (p->my + 7) / 8;
    G.nkeys = G.rowblk ? (long long)p->mx * 64 * G.nbj * ((p->mz + 7) / 8) : G.lot;
    if (G.nkeys >= (1ll << 31) - 64) { G.rowblk = 0; G.nkeys = G.lot; }
    G.c = p->c; G.corr = p->corr; G.cinv = 1.f / p->c; G.quirks = p->quirks; G.pusher = p->pusher; G.external_fields = p->external_fields;
    for (int i = 0; i < 6; i++) G.ext[i] = p->ext[i];
    // particles_movedeposit.F90:1359-1374
    G.minx = 1.f * (G.g + 1); G.maxx = p->mx - 1.f * G.g; G.miny = 1.f * (G.g + 1); G.maxy = p->my - 1.f * G.g;
    if (p->dim == 3) { G.minz = 1.f * (G.gz + 1); G.maxz = p->mz - 1.f * G.gz; }
    else { G.minz = 1.f * (G.gz + 1); G.maxz = 1.f * (G.gz + 1) + 1; }
    const int sx = h->P.sizex, sy = h->P.sizey, sz = h->P.sizez, rank = p->rank;
    G.shiftx_hi = G.maxx - G.minx; G.shifty_hi = G.maxy - G.miny; G.shiftz_hi = G.maxz - G.minz;
    G.shiftx_lo = sx != 1 ? h->mxl[topo_neighbour(rank, sx, sy, sz, 0)] - 1.f * G.nghost : G.shiftx_hi;   // :1589-1593
    G.shifty_lo = sy != 1 ? h->myl[topo_neighbour(rank, sx, sy, sz, 2)] - 1.f * G.nghost : G.shifty_hi;   // :1604-1610
    G.shiftz_lo = p->dim == 3 ? h->mzl[topo_neighbour(rank, sx, sy, sz, 4)] - 1.f * G.nghostz : G.shiftz_hi; // :1624-1629
    G.sendx = sx != 1; G.sendy = sy != 1; G.sendz = p->dim == 3 && sz != 1;
    G.perx = p->periodicx; G.pery = p->periodicy; G.perz = p->periodicz;
    G.x1in = p->x1in; G.x2in = p->x2in; G.y1in = p->y1in; G.y2in = p->y2in; G.z1in = p->z1in; G.z2in = p->z2in;
    G.mxcum = p->mxcum; G.mycum = p->mycum; G.mzcum = p->mzcum;

    size_t lot = (size_t)G.lot;
    int rc = 0;
    for (int a = 0; a < 9; a++) rc |= dalloc(&h->f[a], lot);
    h->nty = (p->my + 3) / 4; h->ntz = (p->mz + 3) / 4; h->shadow_floats = (size_t)h->nty * h->ntz * p->mx * 16;
    for (int a = 0; a < 3; a++) { rc |= dalloc(&h->ftmp[a], lot); rc |= dalloc(&h->shadow[a], h->shadow_floats); }
    h->prim8 = nullptr;
    if (p->dim == 3 && p->order > 0) rc |= dalloc(&h->prim8, 2 * lot);
    size_t plane = (size_t)p->mx * p->my;
    if ((size_t)p->mx * p->mz > plane) plane = (size_t)p->mx * p->mz;
    if ((size_t)p->my * p->mz > plane) plane = (size_t)p->my * p->mz;
    size_t per = 6 * (size_t)(G.g + 1); if ((size_t)12 * p->ntimes > per) per = 12 * (size_t)p->ntimes;   // filter2: 4 slabs x 3 components
    h->halo_floats = per * plane;
    rc |= dalloc(&h->halo, h->halo_floats);
    for (int s = 0; s < 2; s++) {
        rc |= alloc_species(h->sp[s], h->maxhlf); rc |= alloc_species(h->alt[s], h->maxhlf);
        rc |= dalloc(&h->key[s], (size_t)h->maxhlf); rc |= dalloc(&h->perm[s], (size_t)h->maxhlf);
        h->lazy[s] = 0; h->nphys[s] = 0;
    }
    rc |= dalloc(&h->slot, (size_t)2 * h->maxhlf);
    size_t nb = (size_t)G.nkeys + TGPU_NBIN_EXTRA;
    rc |= dalloc(&h->bincount, 2 * nb); rc |= dalloc(&h->binoff, 2 * (nb + 1));
    h->cub_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, h->cub_bytes, h->bincount, h->binoff, (int)(nb + 1), h->stream);
    CK(cudaMalloc(&h->cub_tmp, h->cub_bytes ? h->cub_bytes : 16));
    rc |= dalloc(&h->d_small, 128);
    CK(cudaMallocHost((void **)&h->h_small, 128 * sizeof(int32_t)));
    memset(h->h_small, 0, 128 * sizeof(int32_t));
    h->stage_particles = (size_t)h->maxhlf < ((size_t)1 << 24) ? (size_t)h->maxhlf : ((size_t)1 << 24);
    if (h->stage_particles < 2) h->stage_particles = 2;     // two halves (double buffering)
    rc |= dalloc(&h->stage, h->stage_particles);
    h->sendbuf = h->recvbuf = nullptr;
    if (h->size0 > 1) {
        rc |= dalloc(&h->sendbuf, (size_t)TGPU_NDIR * p->buffsize); rc |= dalloc(&h->recvbuf, (size_t)TGPU_NDIR * p->buffsize);
    }
    if (rc) { tgpu_set_error("device allocation failed: " + g_err); return TGPU_ECUDA; }
    h->need_prim = 1; h->fused_pending = 0; h->keys_valid = 0; h->hook_kind = 0; h->in_step = 0; h->opt_fused = 1; h->opt_fast_push = 1; h->opt_peer = 1; h->opt_graph = 1; h->f1_graph = nullptr; h->f1_graph_launches = 0; h->peer = nullptr; h->sig = nullptr; h->xseq = 0; h->nccl_comm = nullptr; h->lap = 0; h->launches = 0; h->timing = 0;
    for (int i = 0; i < TGPU_NPHASE; i++) h->phase_ms[i] = 0;
    CK(cudaDeviceSynchronize());
    return 0;
}

extern "C" int tgpu_finalize(tgpu_ctx *h)
{
    if (!h) return 0;
    cudaSetDevice(h->device);
    if (h->stream_main) cudaStreamSynchronize(h->stream_main);
    if (h->stream_prt) cudaStreamSynchronize(h->stream_prt);
    comm_destroy(h);
    if (h->f1_graph) { cudaGraphExecDestroy((cudaGraphExec_t)h->f1_graph); h->f1_graph = nullptr; }
    for (int a = 0; a < 9; a++) cudaFree(h->f[a]);
    for (int a = 0; a < 3; a++) { cudaFree(h->ftmp[a]); cudaFree(h->shadow[a]); }
    if (h->prim8) cudaFree(h->prim8);
    cudaFree(h->halo);
    for (int s = 0; s < 2; s++) { free_species(h->sp[s]); free_species(h->alt[s]); cudaFree(h->key[s]); cudaFree(h->perm[s]); }
    cudaFree(h->slot); cudaFree(h->bincount); cudaFree(h->binoff); cudaFree(h->cub_tmp); cudaFree(h->d_small);
    if (h->h_small) cudaFreeHost(h->h_small);
    cudaFree(h->stage);
    auto ev_free = [](cudaEvent_t e) { if (e) cudaEventDestroy(e); };
    auto st_free = [](cudaStream_t s) { if (s) cudaStreamDestroy(s); };
    for (int b = 0; b < 2; b++) { ev_free(h->ev_out_full[b]); ev_free(h->ev_out_free[b]); }
    st_free(h->stream_d2h);
    if (h->sendbuf) cudaFree(h->sendbuf);
    if (h->recvbuf) cudaFree(h->recvbuf);
    ev_free(h->ev0); ev_free(h->ev1); ev_free(h->ev_move); ev_free(h->ev_prt);
    for (int b = 0; b < 2; b++) { ev_free(h->ev_stage_full[b]); ev_free(h->ev_stage_free[b]); }
    st_free(h->stream_main); st_free(h->stream_prt);
    cudaGetLastError();
    delete h;
    return 0;
}

// Every public entry point runs on stream_main.  If tgpu_step left particle work in flight on stream_prt, order it first.
static int join_prt(tgpu_ctx *h)
{
    if (h->prt_pending) { CK(cudaStreamWaitEvent(h->stream_main, h->ev_prt, 0)); h->prt_pending = 0; }
    return 0;
}
#define ENTER(h) do { if (!(h)) { tgpu_set_error("null context"); return TGPU_EINVAL; } CK(cudaSetDevice((h)->device)); \
                      if (!(h)->in_step) { int rcj_ = join_prt(h); if (rcj_) return rcj_; } } while (0)

// per-phase device timing (print_timers analogue); only when enabled, because it synchronises
struct PhaseTimer {
    tgpu_ctx *h; int ph;
    PhaseTimer(tgpu_ctx *h_, int ph_) : h(h_), ph(ph_) { if (h->timing) cudaEventRecord(h->ev0, h->stream); }
    ~PhaseTimer() {
        if (!h->timing) return;
        cudaEventRecord(h->ev1, h->stream); cudaEventSynchronize(h->ev1);
        float ms = 0; cudaEventElapsedTime(&ms, h->ev0, h->ev1); h->phase_ms[ph] += ms;
    }
};

// ---- state transfer ----------------------------------------------------------------------------
static int arrays_copy(tgpu_ctx *h, int first, int n, const float *const *src, float *const *dst, bool h2d)
{
    size_t bytes = (size_t)h->G.lot * sizeof(float);
    for (int a = 0; a < n; a++) {
        if (h2d) { if (!src[a]) { tgpu_set_error("null array"); return TGPU_EINVAL; } CK(cudaMemcpyAsync(h->f[first + a], src[a], bytes, cudaMemcpyHostToDevice, h->stream)); }
        else { if (!dst[a]) { tgpu_set_error("null array"); return TGPU_EINVAL; } CK(cudaMemcpyAsync(dst[a], h->f[first + a], bytes, cudaMemcpyDeviceToHost, h->stream)); }
    }
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}
extern "C" int tgpu_fields_h2d(tgpu_ctx *h, const float *ex, const float *ey, const float *ez, const float *bx, const float *by, const float *bz)
{
    ENTER(h); const float *s[6] = {ex, ey, ez, bx, by, bz}; h->need_prim = 1;
    return arrays_copy(h, 0, 6, s, nullptr, true);
}
extern "C" int tgpu_fields_d2h(tgpu_ctx *h, float *ex, float *ey, float *ez, float *bx, float *by, float *bz)
{
    ENTER(h); float *d[6] = {ex, ey, ez, bx, by, bz};
    int rc = arrays_copy(h, 0, 6, nullptr, d, false);
    return rc ? rc : comm_peer_check(h);          // a halo kernel that gave up waiting for a neighbour is reported here
}
extern "C" int tgpu_currents_h2d(tgpu_ctx *h, const float *cx, const float *cy, const float *cz)
{
    ENTER(h); const float *s[3] = {cx, cy, cz};
    return arrays_copy(h, 6, 3, s, nullptr, true);
}
extern "C" int tgpu_currents_d2h(tgpu_ctx *h, float *cx, float *cy, float *cz)
{
    ENTER(h);
    // Between a fused tgpu_move_particles and tgpu_deposit_particles the lap's deposit sits in the shadow arrays; it is NOT
    // folded in here: the host sees cur exactly as the reference would at this point (e.g. zeros after reset_currents),
    // and tgpu_deposit_particles still finds its pending deposit.
    float *d[3] = {cx, cy, cz};
    return arrays_copy(h, 6, 3, nullptr, d, false);
}
extern "C" int tgpu_particles_h2d(tgpu_ctx *h, const tgpu_particle *p, int ions, int lecs)
{
    ENTER(h); if (!p) { tgpu_set_error("null particles"); return TGPU_EINVAL; }
    for (int i = 0; i < 32; i++) h->h_small[i] = 0;
    if (h->fused_pending) {
        // the uploaded records replace the ones whose motion was deposited into the shadow arrays: drop that deposit
        for (int a = 0; a < 3; a++) CK(cudaMemsetAsync(h->shadow[a], 0, h->shadow_floats * sizeof(float), h->stream));
        h->fused_pending = 0;
    }
    int rc = prt_h2d(h, p, ions, lecs);
    for (int s = 0; s < 2; s++) for (int c = 0; c < 11; c++) h->h_small[s * 16 + c] = h->sp[s].n;
    return rc;
}
extern "C" int tgpu_particles_d2h(tgpu_ctx *h, tgpu_particle *p, int *ions, int *lecs)
{
    ENTER(h); if (!p || !ions || !lecs) { tgpu_set_error("null argument"); return TGPU_EINVAL; }
    return prt_d2h(h, p, ions, lecs);
}
extern "C" int tgpu_counts(tgpu_ctx *h, int *ions, int *lecs)
{
    ENTER(h); if (!ions || !lecs) return TGPU_EINVAL;
    *ions = h->sp[0].n; *lecs = h->sp[1].n;
    return 0;
}
extern "C" int tgpu_append_particles(tgpu_ctx *h, const tgpu_particle *p, int n_ion, int n_lec)
{
    ENTER(h); if (!p || n_ion < 0 || n_lec < 0) return TGPU_EINVAL;
    int rc = prt_append(h, 0, p, n_ion, true); if (rc) return rc;
    rc = prt_append(h, 1, p + n_ion, n_lec, true); if (rc) return rc;
    CK(cudaStreamSynchronize(h->stream));
    for (int s = 0; s < 2; s++) for (int c = 0; c < 11; c++) h->h_small[s * 16 + c] = h->sp[s].n;
    return 0;
}

// ---- fields ------------------------------------------------------------------------------------
extern "C" int tgpu_advance_b_halfstep(tgpu_ctx *h) { ENTER(h); PhaseTimer t(h, TGPU_PH_FIELDS); return fld_bhalf(h); }
extern "C" int tgpu_advance_e_fullstep(tgpu_ctx *h) { ENTER(h); PhaseTimer t(h, TGPU_PH_FIELDS); return fld_efull(h); }
extern "C" int tgpu_reset_currents(tgpu_ctx *h) { ENTER(h); PhaseTimer t(h, TGPU_PH_FIELDS); return fld_reset(h); }
extern "C" int tgpu_add_current(tgpu_ctx *h) { ENTER(h); PhaseTimer t(h, TGPU_PH_FIELDS); return fld_add(h); }
extern "C" int tgpu_bc_b1(tgpu_ctx *h) { ENTER(h); PhaseTimer t(h, TGPU_PH_BC); return fld_bc(h, 3); }
extern "C" int tgpu_bc_e1(tgpu_ctx *h) { ENTER(h); PhaseTimer t(h, TGPU_PH_BC); return fld_bc(h, 0); }
// bc_b2 / bc_e2 = radiation `surface` on every radiating axis (no launch when all axes are periodic), then bc_b1 / bc_e1
extern "C" int tgpu_bc_b2(tgpu_ctx *h)
{
    ENTER(h);
    { PhaseTimer t(h, TGPU_PH_BC); int rc = fld_surface(h, 0); if (rc) return rc; }
    return tgpu_bc_b1(h);
}
extern "C" int tgpu_bc_e2(tgpu_ctx *h)
{
    ENTER(h);
    { PhaseTimer t(h, TGPU_PH_BC); int rc = fld_surface(h, 1); if (rc) return rc; }
    return tgpu_bc_e1(h);
}
// pre_bc_b / post_bc_b / pre_bc_e / post_bc_e (fieldboundaries.F90:114-163, 437-482): the preledge / postedge edge fixes
// of a 3D box whose three axes all radiate, each followed by bc_b1 / bc_e1; no-ops otherwise, as in the reference
static bool all_open3(const tgpu_ctx *h) { return h->P.dim == 3 && !h->P.periodicx && !h->P.periodicy && !h->P.periodicz; }
static int edges_then_bc(tgpu_ctx *h, int which)
{
    if (!all_open3(h)) return 0;
    { PhaseTimer t(h, TGPU_PH_BC); int rc = fld_edges(h, which); if (rc) return rc; }
    return which < 2 ? tgpu_bc_b1(h) : tgpu_bc_e1(h);
}
extern "C" int tgpu_pre_bc_b(tgpu_ctx *h) { ENTER(h); return edges_then_bc(h, 0); }
extern "C" int tgpu_post_bc_b(tgpu_ctx *h) { ENTER(h); return edges_then_bc(h, 1); }
extern "C" int tgpu_pre_bc_e(tgpu_ctx *h) { ENTER(h); return edges_then_bc(h, 2); }
extern "C" int tgpu_post_bc_e(tgpu_ctx *h) { ENTER(h); return edges_then_bc(h, 3); }
extern "C" int tgpu_exchange_current(tgpu_ctx *h) { ENTER(h); PhaseTimer t(h, TGPU_PH_CUREXCH); return fld_fold(h); }
extern "C" int tgpu_apply_filter1_opt(tgpu_ctx *h) { ENTER(h); PhaseTimer t(h, TGPU_PH_FILTER); return fld_filter1(h); }
extern "C" int tgpu_apply_filter2_opt(tgpu_ctx *h) { ENTER(h); PhaseTimer t(h, TGPU_PH_FILTER); return fld_filter2(h); }
extern "C" int tgpu_apply_filter(tgpu_ctx *h)
{
    ENTER(h);
    // tristanmainloop.F90:213-229: filter2 only if compiled in (filter_kind == 2) and its ntimes-deep halo fits inside the
    // neighbour's interior on every filtered axis (all ranks must take the same branch: test the smallest slab), else filter1
    bool f2 = h->P.filter_kind == 2;
    const int naxes = h->P.dim == 3 ? 3 : 2;
    for (int r = 0; r < h->size0 && f2; r++) {
        const int m[3] = {h->mxl[r], h->myl[r], h->mzl[r]}, g[3] = {h->P.nghost / 2, h->P.nghost / 2, h->P.nghostz / 2};
        for (int a = 0; a < naxes; a++) if (h->P.ntimes > m[a] - 2 * g[a] - 1) f2 = false;
    }
    return f2 ? tgpu_apply_filter2_opt(h) : tgpu_apply_filter1_opt(h);
}

// ---- particles -----------------------------------------------------------------------------------
extern "C" int tgpu_move_particles(tgpu_ctx *h)
{
    ENTER(h); PhaseTimer t(h, TGPU_PH_MOVER);
    if (h->fused_pending) { tgpu_set_error("move_particles called twice without deposit_particles"); return TGPU_ESTATE; }
    if (h->opt_fused && fused_mover_supported(h)) {
        int rc = cellrun_supported(h) ? cellrun_move_deposit(h) : cellrun3_move_deposit(h); if (rc) return rc;
        h->fused_pending = 1;
        return 0;
    }
    { int rc = prt_materialize(h); if (rc) return rc; }
    return prt_move_generic(h);
}
extern "C" int tgpu_deposit_particles(tgpu_ctx *h)
{
    ENTER(h);
    int rc;
    {
        PhaseTimer t(h, TGPU_PH_DEPOSIT);
        if (h->fused_pending) { rc = fld_add_shadow(h); h->fused_pending = 0; }     // currents were deposited by the fused mover
        else {
            rc = prt_materialize(h);
            if (!rc) rc = (h->opt_fused && cellrun_supported(h)) ? cellrun_deposit(h)
                        : (h->opt_fused && cellrun3_supported(h)) ? cellrun3_deposit(h) : prt_deposit_generic(h);
        }
        if (rc) return rc;
    }
    PhaseTimer t2(h, TGPU_PH_SORT);
    return prt_sort(h, false);     // loops B, C of deposit_particles (+ the counting sort)
}
extern "C" int tgpu_exchange_particles(tgpu_ctx *h) { ENTER(h); PhaseTimer t(h, TGPU_PH_PEXCH); return prt_exchange(h); }
extern "C" int tgpu_inject_others(tgpu_ctx *h) { ENTER(h); return 0; }
extern "C" int tgpu_reorder_particles(tgpu_ctx *h) { ENTER(h); PhaseTimer t(h, TGPU_PH_SORT); return prt_sort(h, false); }

// ---- shock-problem user hooks -----------------------------------------------------------------------
extern "C" int tgpu_field_bc_user_shock(tgpu_ctx *h, float leftwall, float binit, float btheta, float bphi, float beta)
{
    ENTER(h); PhaseTimer t(h, TGPU_PH_BC);
    return fld_bc_shock(h, leftwall, binit, btheta, bphi, beta);
}
extern "C" int tgpu_particle_bc_user_wall(tgpu_ctx *h, float leftwall)
{
    ENTER(h); PhaseTimer t(h, TGPU_PH_DEPOSIT);
    if (h->fused_pending) { tgpu_set_error("particle_bc_user needs the un-fused mover: set_option(\"fused\", 0) or tgpu_set_user_hooks"); return TGPU_ESTATE; }
    return prt_wall(h, leftwall);
}
extern "C" int tgpu_set_user_hooks(tgpu_ctx *h, int kind, const float params[5])
{
    if (!h || kind < 0 || kind > 1 || (kind && !params)) return TGPU_EINVAL;
    h->hook_kind = kind;
    for (int i = 0; i < 5; i++) h->hook[i] = kind ? params[i] : 0.f;
    return 0;
}

// ---- mirror mode, whole lap ----------------------------------------------------------------------------
// One lap with the state owned by the host: fields and particles come in from host arrays and go back to them, i.e. what
// a call-for-call GPU build of mainloop would do per lap (fields_h2d, particles_h2d, every tgpu_* of the lap,
// particles_d2h, fields_d2h).  When nobody can leave the rank (one rank, all axes periodic) and the fused mover applies,
// the particle array is STREAMED: inbound chunks, the fused mover + deposit and outbound chunks overlap, so the lap costs
// max(H2D, D2H) of the 40-byte records instead of their sum.  Otherwise the plain sequence runs.
// Returns the particles in the host's order (streamed) or cell-sorted (plain); compare by (proc, ind).
extern "C" int tgpu_step_mirror(tgpu_ctx *h, float *ex, float *ey, float *ez, float *bx, float *by, float *bz,
                                tgpu_particle *p, int *ions, int *lecs)
{
    ENTER(h);
    if (!ex || !ey || !ez || !bx || !by || !bz || !p || !ions || !lecs) { tgpu_set_error("null argument"); return TGPU_EINVAL; }
    const bool stream_ok = h->size0 == 1 && h->P.periodicx && h->P.periodicy && (h->P.dim == 2 || h->P.periodicz) &&
                           h->opt_fused && cellrun_supported(h) && h->hook_kind == 0 && h->stage_particles >= 4 && !h->timing;
    int rc;
    if (!stream_ok) {
        rc = tgpu_fields_h2d(h, ex, ey, ez, bx, by, bz); if (rc) return rc;
        rc = tgpu_particles_h2d(h, p, *ions, *lecs); if (rc) return rc;
        rc = tgpu_step(h, 1); if (rc) return rc;
        rc = tgpu_particles_d2h(h, p, ions, lecs); if (rc) return rc;
        return tgpu_fields_d2h(h, ex, ey, ez, bx, by, bz);
    }
    float *hf[6] = {ex, ey, ez, bx, by, bz};
    const size_t fbytes = (size_t)h->G.lot * sizeof(float);
    h->in_step = 1;
#define DO(x) do { rc = (x); if (rc) { h->in_step = 0; return rc; } } while (0)
#define CKS(x) do { if ((x) != cudaSuccess) { h->in_step = 0; tgpu_set_error(#x); return TGPU_ECUDA; } } while (0)
    for (int a = 0; a < 6; a++) CKS(cudaMemcpyAsync(h->f[a], hf[a], fbytes, cudaMemcpyHostToDevice, h->stream_main));
    h->need_prim = 1; h->fused_pending = 0; h->lap++;
    DO(tgpu_bc_e1(h));                 // :118
    DO(tgpu_advance_b_halfstep(h));    // :119
    DO(tgpu_bc_b1(h));                 // :122
    DO(fld_primal(h));
    DO(prt_mirror_stream(h, p, *ions, *lecs));       // :134 + particle part of :183, streamed
    DO(tgpu_advance_b_halfstep(h));    // :139
    DO(tgpu_bc_b1(h));                 // :140
    DO(tgpu_advance_e_fullstep(h));    // :159
    DO(tgpu_reset_currents(h));        // :171
    DO(fld_add_shadow(h));             // :183, current part
    DO(tgpu_exchange_current(h));      // :203
    DO(tgpu_apply_filter(h));          // :213-229
    DO(tgpu_add_current(h));           // :242
    for (int a = 0; a < 6; a++) CKS(cudaMemcpyAsync(hf[a], h->f[a], fbytes, cudaMemcpyDeviceToHost, h->stream_main));
    CKS(cudaStreamSynchronize(h->stream_d2h));
    CKS(cudaStreamSynchronize(h->stream_main));
#undef DO
#undef CKS
    h->in_step = 0;
    *ions = h->sp[0].n; *lecs = h->sp[1].n;
    for (int s = 0; s < 2; s++) for (int c = 0; c < 11; c++) h->h_small[s * 16 + c] = h->sp[s].n;
    return 0;
}

// ---- output-side reductions -------------------------------------------------------------------------
// meanq_fld_cur(totname), output.F90:5229-5486: the moment lands in curx (cury = weight), as in the reference; the host
// reads it with tgpu_currents_d2h instead of pulling every particle across PCIe on an output lap
extern "C" int tgpu_meanq_fld_cur(tgpu_ctx *h, const char *totname) { ENTER(h); return prt_meanq(h, totname); }
// per-rank part of save_spectrum, output.F90:380-633 (the host keeps the two allreduces and the division by xgamma)
extern "C" int tgpu_spectrum_gamma_range(tgpu_ctx *h, float *gammin, float *gammax)
{
    ENTER(h); if (!gammin || !gammax) return TGPU_EINVAL;
    return prt_gamma_range(h, gammin, gammax);
}
extern "C" int tgpu_spectrum(tgpu_ctx *h, float gammin, float gammax, int mx0, float splitratio, int nbins, int gambins,
                             float *specp, float *spece, float *specprest, float *specerest)
{
    ENTER(h); return prt_spectrum(h, gammin, gammax, mx0, splitratio, nbins, gambins, specp, spece, specprest, specerest);
}
// the prtl.tot sub-sample, output.F90:3526-3551: particles with modulo(ind/2, stride) == 0
extern "C" int tgpu_select_particles(tgpu_ctx *h, int stride, tgpu_particle *out, int capacity, int *n_ion, int *n_lec)
{
    ENTER(h); return prt_select(h, stride, out, capacity, n_ion, n_lec);
}

// ---- whole lap -------------------------------------------------------------------------------------
// Call order of tristanmainloop.F90:107-344 with the redundant ghost refreshes of Appendix B removed: three
// refreshes per lap instead of eight.  Results on the parity region are identical to the full call list.
extern "C" int tgpu_step(tgpu_ctx *h, int nlaps)
{
    ENTER(h);
    int rc = 0;
#define DO(x) do { rc = (x); if (rc) { h->in_step = 0; return rc; } } while (0)
    // Overlap: once the fused mover has written the sort keys, the scan + scatter + migration of the particles depends
    // on nothing the field phase does (and vice versa), so it runs on stream_prt while B-half/E-full/fold/filter/add run
    // on stream_main.  Needs the fused mover (keys in hand) and is skipped while per-phase timing is on.
    const bool overlap = h->opt_overlap && h->opt_fused && fused_mover_supported(h) && !h->timing;
    // bc_b2 / bc_e2 differ from bc_b1 / bc_e1 only when an axis radiates (fieldboundaries.F90:90-94, 274-295, 403-426)
    const bool rad = !h->P.periodicx || !h->P.periodicy || (h->P.dim == 3 && !h->P.periodicz);
    h->in_step = 1;
    for (int l = 0; l < nlaps && h->hook_kind == 1; l++) {
        // shock problem: the reflecting wall edits particles between the mover and the deposit, so the mover is not fused
        // and the full call list with its hook points is replayed (tristanmainloop.F90:117-243)
        const float *q = h->hook;
        const int fused0 = h->opt_fused;
        h->lap++;
        DO(tgpu_pre_bc_b(h));                                         // :114 (acts only in an all-open 3D box)
        DO(tgpu_bc_b1(h)); DO(tgpu_bc_e1(h)); DO(tgpu_advance_b_halfstep(h)); DO(tgpu_bc_b1(h));
        h->opt_fused = 0; rc = tgpu_move_particles(h); h->opt_fused = fused0; if (rc) { h->in_step = 0; return rc; }
        DO(tgpu_advance_b_halfstep(h)); DO(tgpu_bc_b1(h)); DO(tgpu_bc_b2(h));
        DO(fld_bc_shock(h, q[0], q[1], q[2], q[3], q[4]));            // :146
        DO(tgpu_post_bc_b(h)); DO(tgpu_pre_bc_e(h));                  // :155, :157
        DO(tgpu_advance_e_fullstep(h));
        DO(fld_bc_shock(h, q[0], q[1], q[2], q[3], q[4]));            // :160
        DO(tgpu_bc_e2(h));
        DO(tgpu_post_bc_e(h));                                        // :165
        DO(fld_bc_shock(h, q[0], q[1], q[2], q[3], q[4]));            // :166
        DO(tgpu_reset_currents(h));
        DO(prt_wall(h, q[0]));                                        // :177
        DO(tgpu_bc_e1(h)); DO(tgpu_bc_b1(h));
        DO(tgpu_deposit_particles(h)); DO(tgpu_exchange_particles(h)); DO(tgpu_exchange_current(h));
        DO(tgpu_apply_filter(h)); DO(tgpu_add_current(h));
        DO(fld_bc_shock(h, q[0], q[1], q[2], q[3], q[4]));            // :243
    }
    for (int l = 0; l < nlaps && h->hook_kind == 0; l++) {
        h->lap++;
        if (rad) DO(tgpu_pre_bc_b(h));     // :114 (acts only in an all-open 3D box)
        DO(tgpu_bc_e1(h));                 // :118 (E changed by add_current)
        DO(tgpu_advance_b_halfstep(h));    // :119
        DO(tgpu_bc_b1(h));                 // :122
        DO(join_prt(h));                   // particles of the previous lap are sorted and migrated
        DO(tgpu_move_particles(h));        // :134
        if (overlap) {
            DO(cudaEventRecord(h->ev_move, h->stream_main) == cudaSuccess ? 0 : (tgpu_set_error("cudaEventRecord(ev_move)"), TGPU_ECUDA));
            DO(tgpu_advance_b_halfstep(h));    // :139
            DO(tgpu_bc_b1(h));                 // :140
            if (rad) DO(tgpu_bc_b2(h));        // :145
            if (rad) { DO(tgpu_post_bc_b(h)); DO(tgpu_pre_bc_e(h)); }   // :155, :157
            DO(tgpu_advance_e_fullstep(h));    // :159
            if (rad) { DO(tgpu_bc_e2(h)); DO(tgpu_post_bc_e(h)); }      // :164, :165
            DO(tgpu_reset_currents(h));        // :171
            DO(fld_add_shadow(h)); h->fused_pending = 0;      // :183, current part of deposit_particles
            DO(tgpu_exchange_current(h));      // :203
            DO(tgpu_apply_filter(h));          // :213-229
            DO(tgpu_add_current(h));           // :242
            // particle side, concurrently
            h->stream = h->stream_prt; h->nccl_comm = h->nccl_prt;
            rc = cudaStreamWaitEvent(h->stream_prt, h->ev_move, 0) == cudaSuccess ? 0 : TGPU_ECUDA;
            if (!rc) rc = prt_sort(h, false);                  // :183 loops B, C (+ counting sort)
            if (!rc) rc = prt_exchange(h);                     // :190, :257-272
            if (!rc) rc = cudaEventRecord(h->ev_prt, h->stream_prt) == cudaSuccess ? 0 : TGPU_ECUDA;
            h->stream = h->stream_main; h->nccl_comm = h->nccl_main;
            if (rc) { h->in_step = 0; if (rc == TGPU_ECUDA) tgpu_set_error("stream overlap failed"); return rc; }
            h->prt_pending = 1;
        } else {
            DO(tgpu_advance_b_halfstep(h));    // :139
            DO(tgpu_bc_b1(h));                 // :140
            if (rad) DO(tgpu_bc_b2(h));        // :145
            if (rad) { DO(tgpu_post_bc_b(h)); DO(tgpu_pre_bc_e(h)); }   // :155, :157
            DO(tgpu_advance_e_fullstep(h));    // :159
            if (rad) { DO(tgpu_bc_e2(h)); DO(tgpu_post_bc_e(h)); }      // :164, :165
            DO(tgpu_reset_currents(h));        // :171
            DO(tgpu_deposit_particles(h));     // :183
            DO(tgpu_exchange_particles(h));    // :190, :257-272
            DO(tgpu_exchange_current(h));      // :203
            DO(tgpu_apply_filter(h));          // :213-229
            DO(tgpu_add_current(h));           // :242
        }
    }
    h->in_step = 0;
#undef DO
    // anything recorded on stream_main after this call (the caller's timing events) must also cover the particle stream
    return join_prt(h);
}

extern "C" int tgpu_timers(tgpu_ctx *h, double *out_ms, int reset)
{
    ENTER(h);
    if (out_ms) for (int i = 0; i < TGPU_NPHASE; i++) out_ms[i] = h->phase_ms[i];
    if (reset) for (int i = 0; i < TGPU_NPHASE; i++) h->phase_ms[i] = 0;
    return 0;
}
extern "C" int64_t tgpu_launch_count(tgpu_ctx *h) { return h ? h->launches : 0; }
extern "C" void *tgpu_stream(tgpu_ctx *h) { return h ? (void *)h->stream : nullptr; }
extern "C" int tgpu_halo_transport(tgpu_ctx *h) { return h && h->peer ? 1 : 0; }
extern "C" int tgpu_set_option(tgpu_ctx *h, const char *name, int value)
{
    if (!h || !name) return TGPU_EINVAL;
    if (!strcmp(name, "fused")) { h->opt_fused = value; return 0; }
    if (!strcmp(name, "fast_push")) { h->opt_fast_push = value; return 0; }
    if (!strcmp(name, "peer")) { h->opt_peer = value; return 0; }
    if (!strcmp(name, "graph")) { h->opt_graph = value; return 0; }
    if (!strcmp(name, "blocked_rows")) {      // 0: sort by the reference's key i + mx*(j + my*k); 1 (3D default): rows in 8 x 8 blocks
        if (value && (h->P.dim != 3 || (long long)h->P.mx * 64 * h->G.nbj * ((h->P.mz + 7) / 8) != h->G.nkeys)) return TGPU_EINVAL;
        h->G.rowblk = value ? 1 : 0; h->keys_valid = 0;
        return prt_materialize(h);
    }     // 0: launch the filter1 passes one by one       // before tgpu_comm_init: 0 = halos through NCCL send/recv
    if (!strcmp(name, "timing")) { h->timing = value; return 0; }
    if (!strcmp(name, "overlap")) { h->opt_overlap = value; return 0; }
    if (!strcmp(name, "lazy_sort")) { h->opt_lazy = value; return 0; }
    tgpu_set_error(std::string("unknown option ") + name);
    return TGPU_EINVAL;
}
