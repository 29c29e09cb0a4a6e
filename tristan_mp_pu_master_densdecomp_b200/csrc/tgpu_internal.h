// Internal declarations shared by the translation units of libtristan_gpu.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/tristan_gpu.h"

#ifndef TGPU_SHADOW_TILED
#define TGPU_SHADOW_TILED 1  // fused mover deposits into 4x4 (y,z)-tiled shadow arrays (cellrun.cu row_index); 0 = Fortran order (A/B)
#endif
#define TGPU_NDIR 9          // 3x3 neighbour codes over the two decomposed axes; code 4 = stay
#define TGPU_NBIN_EXTRA 10   // sort bins after the `lot` cell bins: 9 direction codes (4 unused) + discard

struct Species {
    float *x, *y, *z, *u, *v, *w, *ch;
    int32_t *ind, *tag;      // tag = proc | splitlev << 24
    int n;                   // live particles (host copy)
};

struct DevGeom {             // passed by value to kernels
    int dim, order;
    int mx, my, mz;
    int g, gz;               // nghost/2, nghostz/2
    int nghost, nghostz;
    long long lot;
    // sort keys of the cells (tgpu_internal.h cell_key): nkeys bins; rowblk != 0: the (y,z) rows are numbered in 8 x 8 blocks
    // (x stays the fastest index), so that rows that share field nodes and current cells are swept close together in time
    long long nkeys;
    int rowblk, nbj;
    float c, corr, cinv;
    int quirks, pusher, external_fields;
    float ext[6];
    // classification (particles_movedeposit.F90:1359-1374, 1546-1633)
    float minx, maxx, miny, maxy, minz, maxz;
    float shiftx_lo, shiftx_hi, shifty_lo, shifty_hi, shiftz_lo, shiftz_hi;  // amount subtracted when leaving low/high side
    int sendx, sendy, sendz;   // axis is split across ranks (leavers are sent, not wrapped)
    int perx, pery, perz;      // periodic flags
    float x1in, x2in, y1in, y2in, z1in, z2in;
    int mxcum, mycum, mzcum;
};

// Peer-memory halo transport (comm.cu comm_peer_setup, fields.cu halo_step): the nine field arrays and two signal words
// of every neighbouring rank are mapped into this process with cudaIpc, and the halo kernels read them over NVLink.
struct PeerRank {
    float *f[9];             // the neighbour's ex..bz, curx..curz
    uint32_t *sig;           // its signal words (TGPU_SIG_*)
    void *base[10];          // what cudaIpcOpenMemHandle returned (for cudaIpcCloseMemHandle)
    int open;
};
#define TGPU_SIG_READY 0     // "everything I produce before exchange number s is in memory"
#define TGPU_SIG_PULLED 1    // "I have finished reading my neighbours' arrays for exchange number s"
#define TGPU_SIG_TIMEOUT 2   // set by a kernel that gave up waiting (5 s): the job is broken, reported at the next host sync
#define TGPU_SIG_FIN 3       // "my streams are drained, I am about to unmap and free" (comm.cu peer_close)

struct tgpu_ctx {
    tgpu_params P;
    DevGeom G;
    int size0, device, maxhlf;
    std::vector<int> mxl, myl, mzl;
    float *f[9];             // ex..bz, curx..curz
    float4 *prim8;           // node-centred fields, 8 floats per node (3D shaped movers)
    float *ftmp[3];          // filter1 scratch (the reference's `temp`, one per component)
    float *halo;             // pack/unpack scratch for exchanges and filter2 deep halos
    size_t halo_floats;
    Species sp[2], alt[2];
    int32_t *perm[2];        // lazy sort: logical position d of species s lives at physical index perm[s][d] of sp[s]
    int lazy[2];             // perm[s] is in force (positions in sp[s] are still unwrapped)
    int nphys[2];            // physical records in sp[s] while lazy (stayers, leavers and appended arrivals)
    int opt_lazy;
    uint32_t *key[2];        // cell key per particle (per species)
    int32_t *slot;           // rank of the particle inside its bin
    int32_t *bincount, *binoff;   // lot + TGPU_NBIN_EXTRA (+1)
    void *cub_tmp; size_t cub_bytes;
    int32_t *d_small, *h_small;   // 128 ints device / pinned host
    tgpu_particle *stage;    // device AoS staging for h2d/d2h and migration (2*buffsize*9 .. maxhlf)
    size_t stage_particles;
    tgpu_particle *sendbuf, *recvbuf;   // TGPU_NDIR * buffsize each
    int need_prim;           // primal grids stale
    int presort;             // records were uploaded in host order: cell-sort them once before the next fused mover
    int fused_pending;       // currents of the last move already deposited into shadow
    float *shadow[3];        // tiled: [tz][ty][i][4x4 (y,z) tile] (cellrun.cu row_index), nty x ntz tiles
    int nty, ntz; size_t shadow_floats;
    int opt_fused;
    int opt_fast_push;       // cell-run movers: SFU rcp / rsqrt in the Boris push (default 1)
    int hook_kind; float hook[5];   // user hooks for tgpu_step (1 = shock: leftwall, binit, btheta, bphi, beta)
    int keys_valid;          // key[]/slot[]/bincount[] already hold this lap's sort keys (written by the fused mover)
    cudaStream_t stream;     // the stream every launch helper uses (normally == stream_main)
    cudaStream_t stream_main, stream_prt;   // tgpu_step overlaps the particle sort/migration (stream_prt) with the field phase
    cudaEvent_t ev0, ev1, ev_move, ev_prt;
    cudaEvent_t ev_stage_full[2], ev_stage_free[2];   // double-buffered AoS staging for h2d / d2h
    cudaEvent_t ev_out_full[2], ev_out_free[2];       // outbound staging of the streamed mirror lap
    cudaStream_t stream_d2h;                          // device -> host copies of the streamed mirror lap
    int in_step;
    int prt_pending;         // ev_prt must be waited for before the particle arrays are touched on stream_main
    int opt_overlap;
    void *nccl_comm;         // ncclComm_t used on `stream` (normally == nccl_main)
    void *nccl_main, *nccl_prt;   // two communicators: NCCL calls from two streams must not share one
    PeerRank *peer;          // [size0]; null = halo exchanges go through NCCL send/recv
    uint32_t *sig;           // own signal words (device memory, mapped by the neighbours)
    uint32_t xseq;           // exchange counter: every rank runs the same sequence of halo steps
    int opt_peer;
    int opt_graph;           // replay fixed launch sequences (filter1 passes on one rank) as CUDA graphs
    void *f1_graph; int f1_graph_launches;
    int lap;
    int64_t launches;
    double phase_ms[TGPU_NPHASE];
    int timing;
};

// Address of the lane's row (J, K) (0-based) at x-plane i0 (0-based).
//   TILED = false: the reference's Fortran-order arrays, element i0 + mx*(J + my*K); consecutive planes are 1 apart.
//   TILED = true : the shadow arrays of the fused mover.  The 16 lanes of a half-warp flush one x-plane of the 4x4 (y,z)
//     footprint; in Fortran order that is 16 different cache lines per RED instruction.  The shadow arrays therefore keep
//     every 4x4 (y,z) tile of an x-plane contiguous (64 B): element ((tz*nty + ty)*mx + i0)*16 + (K&3)*4 + (J&3), so a flush
//     touches at most 2x2 tiles (<= 4 short segments) and consecutive planes are 16 apart.  k_add_shadow_tiled (fields.cu)
//     folds them back into curx/cury/curz.
template <bool TILED>
__device__ __forceinline__ size_t row_index(int mx, int my, int nty, int J, int K, int i0)
{
    if (TILED) return ((size_t)(((K >> 2) * nty + (J >> 2)) * (long long)mx + i0) << 4) + (size_t)(((K & 3) << 2) | (J & 3));
    return (size_t)((long long)mx * (J + (long long)my * K) + i0);
}

// Sort key of cell (i, j, k) (1-based).  The reference's reorder key is i-1 + mx*((j-1) + my*(k-1)) (particles.F90:441).
// With rowblk the rows are enumerated block by block: key = i-1 + mx * (((kb*nbj + jb) << 6) | kl << 3 | jl), (jb, jl) =
// divmod(j-1, 8), (kb, kl) = divmod(k-1, 8).  Any bijection serves the counting sort; this one keeps a cell run contiguous
// along x (what the cell-run kernels need) while rows j+-1 and k+-1 -- which read the same node-centred field sectors and
// flush into the same current tiles -- are processed within ~64 rows of each other instead of my rows apart: their
// sectors are still in L2 (ncu: DRAM traffic of the fused kernel 150 -> 133 B per particle, 7.52 -> 7.36 ms per launch).
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t cell_key(const DevGeom &G, int i, int j, int k)
{
    if (!G.rowblk) return (uint32_t)((i - 1) + G.mx * ((j - 1) + G.my * (k - 1)));
    const int jj = j - 1, kk = k - 1;
    const int row = ((((kk >> 3) * G.nbj + (jj >> 3)) << 6) | ((kk & 7) << 3)) | (jj & 7);
    return (uint32_t)((i - 1) + G.mx * row);
}
#endif

void tgpu_set_error(const std::string &s);
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    tgpu_set_error(std::string(#call) + ": " + cudaGetErrorString(e_)); return TGPU_ECUDA; } } while (0)
#define CKK(h) do { cudaError_t e_ = cudaGetLastError(); (h)->launches++; if (e_ != cudaSuccess) { \
    tgpu_set_error(std::string("kernel launch: ") + cudaGetErrorString(e_) + " at " + __FILE__ + ":" + std::to_string(__LINE__)); \
    return TGPU_ECUDA; } } while (0)

static inline int cdiv(long long a, int b) { return (int)((a + b - 1) / b); }

// fields.cu
int fld_bhalf(tgpu_ctx *h);
int fld_efull(tgpu_ctx *h);
int fld_reset(tgpu_ctx *h);
int fld_add(tgpu_ctx *h);
int fld_primal(tgpu_ctx *h);
int fld_bc(tgpu_ctx *h, int first);
int fld_fold(tgpu_ctx *h);
int fld_filter1(tgpu_ctx *h);
int fld_filter2(tgpu_ctx *h);
int fld_add_shadow(tgpu_ctx *h);
int fld_surface(tgpu_ctx *h, int is_e);     // radiation `surface` of bc_b2 (0) / bc_e2 (1)
int fld_edges(tgpu_ctx *h, int which);      // preledge / postedge groups: 0 pre_bc_b, 1 post_bc_b, 2 pre_bc_e, 3 post_bc_e
int fld_bc_shock(tgpu_ctx *h, float leftwall, float binit, float btheta, float bphi, float beta);
// particles.cu
int prt_h2d(tgpu_ctx *h, const tgpu_particle *p, int ions, int lecs);
int prt_d2h(tgpu_ctx *h, tgpu_particle *p, int *ions, int *lecs);
int prt_append(tgpu_ctx *h, int s, const tgpu_particle *p, int n, bool host);
int prt_move(tgpu_ctx *h);
int prt_deposit(tgpu_ctx *h);
int prt_sort(tgpu_ctx *h, bool classify_only);
int prt_materialize(tgpu_ctx *h);     // apply a pending lazy permutation (+ wrap) physically
int prt_exchange(tgpu_ctx *h);
int prt_wall(tgpu_ctx *h, float leftwall);
int prt_meanq(tgpu_ctx *h, const char *totname);
int prt_gamma_range(tgpu_ctx *h, float *gammin, float *gammax);
int prt_spectrum(tgpu_ctx *h, float gammin, float gammax, int mx0, float splitratio, int nbins, int gambins,
                 float *specp, float *spece, float *specprest, float *specerest);   // save_spectrum, per-rank part
int prt_select(tgpu_ctx *h, int stride, tgpu_particle *out_host, int capacity, int *n_ion, int *n_lec);   // prtl.tot selection
int prt_mirror_stream(tgpu_ctx *h, tgpu_particle *p, int ions, int lecs);   // particle side of tgpu_step_mirror   // meanq_fld_cur, output.F90:5229-5486
// comm.cu
int comm_sendrecv(tgpu_ctx *h, const void *sbuf, size_t sbytes, int dst, void *rbuf, size_t rbytes, int src);
int comm_group_begin(tgpu_ctx *h);
int comm_group_end(tgpu_ctx *h);
int comm_send(tgpu_ctx *h, const void *buf, size_t bytes, int peer);
int comm_recv(tgpu_ctx *h, void *buf, size_t bytes, int peer);
int topo_neighbour(int rank, int sx, int sy, int sz, int dir);
int topo_neighbour2(const tgpu_ctx *h, int da, int db);   // neighbour over the two decomposed axes
int comm_peer_setup(tgpu_ctx *h);     // map the neighbours' arrays (cudaIpc); all ranks or none
int comm_peer_check(tgpu_ctx *h);     // TGPU_ENCCL if a halo kernel timed out waiting for a neighbour
