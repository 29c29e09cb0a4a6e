// Rank topology (host arithmetic) and the NCCL transport that replaces the reference's blocking
// MPI_SendRecv pairs (code/communications.F90:127-161; call sites in fieldboundaries.F90 / particles.F90).
// NCCL is bound at run time with dlopen so that the library loads (and its host-only entry points work)
// on machines without a GPU or without libnccl.
#include <dlfcn.h>
#include <stdlib.h>
#include <unistd.h>
#include <string.h>
#include "tgpu_internal.h"

static int imodulo(int a, int b) { int m = a % b; return m < 0 ? m + b : m; }

// fieldboundaries.F90:1170-1171 (x), :1311-1314 (y); particles.F90:1892-1895 (z)
int topo_neighbour(int rank, int sx, int sy, int sz, int dir)
{
    switch (dir) {
    case 0: return (rank / sx) * sx + imodulo(rank - 1, sx);
    case 1: return (rank / sx) * sx + imodulo(rank + 1, sx);
    case 2: return imodulo(rank / sx - 1, sy) * sx + rank / (sx * sy) * (sx * sy) + imodulo(rank, sx);
    case 3: return imodulo(rank / sx + 1, sy) * sx + rank / (sx * sy) * (sx * sy) + imodulo(rank, sx);
    case 4: return imodulo(rank / (sx * sy) - 1, sz) * (sx * sy) + imodulo(rank, sx * sy);
    default: return imodulo(rank / (sx * sy) + 1, sz) * (sx * sy) + imodulo(rank, sx * sy);
    }
}

// neighbour at offset (da, db) over the two decomposed axes: (y,z) in 3D, (x,y) in 2D
int topo_neighbour2(const tgpu_ctx *h, int da, int db)
{
    const tgpu_params &P = h->P;
    int r = P.rank;
    int ax_a = P.dim == 3 ? 1 : 0, ax_b = P.dim == 3 ? 2 : 1;
    if (da) r = topo_neighbour(r, P.sizex, P.sizey, P.sizez, 2 * ax_a + (da > 0));
    if (db) r = topo_neighbour(r, P.sizex, P.sizey, P.sizez, 2 * ax_b + (db > 0));
    return r;
}

extern "C" int tgpu_neighbour(int rank, int sizex, int sizey, int sizez, int dir)
{
    if (sizex < 1 || sizey < 1 || sizez < 1 || dir < 0 || dir > 5) return -1;
    return topo_neighbour(rank, sizex, sizey, sizez, dir);
}

extern "C" int tgpu_ghost_width(int dim, int order, int32_t *nghost, int32_t *nghostz)
{
    if ((dim != 2 && dim != 3) || order < 0 || order > 3) return TGPU_EINVAL;
    *nghost = order <= 1 ? 5 : 7; *nghostz = dim == 2 ? 5 : *nghost;     // fields.F90:166-184
    return 0;
}

extern "C" int tgpu_decompose(int dim, int order, int mx0, int my0, int mz0, int sx, int sy, int sz, int rank, int32_t out[6])
{
    int32_t ng, ngz;
    if (tgpu_ghost_width(dim, order, &ng, &ngz)) return TGPU_EINVAL;
    if (dim == 2) sz = 1;
    if (sx < 1 || sy < 1 || sz < 1 || rank < 0 || rank >= sx * sy * sz) return TGPU_EINVAL;
    if (dim == 3 && sx != 1) return TGPU_EINVAL;                         // communications.F90:176-181
    int gx = mx0 + ng, gy = my0 + ng, gz = dim == 2 ? 1 : mz0 + ngz;
    // fields.F90:259-280
    auto local = [&](int rk, int *mx, int *my, int *mz) {
        *mx = (gx - ng) / sx + ng; *my = (gy - ng) / sy + ng; *mz = dim == 2 ? 1 : (gz - ngz) / sz + ngz;
        if (rk % sx == sx - 1 && gx != (*mx - ng) * sx + ng) *mx = gx - (*mx - ng) * (sx - 1);
        if ((rk % (sx * sy)) / sx == sy - 1 && gy != (*my - ng) * sy + ng) *my = gy - (*my - ng) * (sy - 1);
        if (dim == 3 && rk / (sx * sy) == sz - 1 && gz != (*mz - ngz) * sz + ngz) *mz = gz - (*mz - ngz) * (sz - 1);
    };
    int mx, my, mz; local(rank, &mx, &my, &mz);
    out[0] = mx; out[1] = my; out[2] = mz;
    int cx = 0, cy = 0, cz = 0, a, b, c;
    for (int i = 0; i < rank % sx; i++) { local((rank / sx) * sx + i, &a, &b, &c); cx += a - ng; }       // fields.F90:316-328
    for (int j = 0; j < (rank % (sx * sy)) / sx; j++) { local(j * sx, &a, &b, &c); cy += b - ng; }
    if (dim == 3) for (int k = 0; k < rank / (sx * sy); k++) { local(k * sx * sy, &a, &b, &c); cz += c - ngz; }
    out[3] = cx; out[4] = cy; out[5] = cz;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// NCCL through dlopen
// ---------------------------------------------------------------------------------------------
typedef struct { char internal[128]; } nccl_uid;
typedef void *nccl_comm_t;
typedef int (*fn_GetUniqueId)(nccl_uid *);
typedef int (*fn_CommInitRank)(nccl_comm_t *, int, nccl_uid, int);
typedef int (*fn_CommDestroy)(nccl_comm_t);
typedef int (*fn_Send)(const void *, size_t, int, int, nccl_comm_t, cudaStream_t);
typedef int (*fn_Recv)(void *, size_t, int, int, nccl_comm_t, cudaStream_t);
typedef int (*fn_Group)(void);
typedef const char *(*fn_ErrStr)(int);
static struct {
    void *lib; fn_GetUniqueId GetUniqueId; fn_CommInitRank CommInitRank; fn_CommDestroy CommDestroy;
    fn_Send Send; fn_Recv Recv; fn_Group GroupStart, GroupEnd; fn_ErrStr ErrStr;
} N;

static int nccl_load()
{
    if (N.lib) return 0;
    const char *names[] = {getenv("TGPU_NCCL_LIB"), "libnccl.so.2", "libnccl.so", nullptr};
    for (int i = 0; i < 3 && !N.lib; i++) if (names[i] && names[i][0]) N.lib = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
    if (!N.lib) { tgpu_set_error(std::string("cannot load NCCL: ") + (dlerror() ? dlerror() : "?")); return TGPU_ENCCL; }
#define SYM(n) N.n = (fn_##n)dlsym(N.lib, "nccl" #n); if (!N.n) { tgpu_set_error("NCCL symbol missing: nccl" #n); return TGPU_ENCCL; }
    SYM(GetUniqueId) SYM(CommInitRank) SYM(CommDestroy) SYM(Send) SYM(Recv)
#undef SYM
    N.GroupStart = (fn_Group)dlsym(N.lib, "ncclGroupStart"); N.GroupEnd = (fn_Group)dlsym(N.lib, "ncclGroupEnd");
    N.ErrStr = (fn_ErrStr)dlsym(N.lib, "ncclGetErrorString");
    if (!N.GroupStart || !N.GroupEnd) { tgpu_set_error("NCCL group symbols missing"); return TGPU_ENCCL; }
    return 0;
}
#define NCK(call) do { int r_ = (call); if (r_ != 0) { tgpu_set_error(std::string(#call) + ": " + (N.ErrStr ? N.ErrStr(r_) : "nccl error")); return TGPU_ENCCL; } } while (0)

extern "C" int tgpu_comm_unique_id(uint8_t id[128])
{
    int rc = nccl_load(); if (rc) return rc;
    nccl_uid u; NCK(N.GetUniqueId(&u));
    memcpy(id, u.internal, 128);
    return 0;
}

typedef int (*fn_Bcast)(const void *, void *, size_t, int, int, nccl_comm_t, cudaStream_t);

extern "C" int tgpu_comm_init(tgpu_ctx *h, const uint8_t id[128])
{
    if (!h) return TGPU_EINVAL;
    if (h->size0 == 1) return 0;
    int rc = nccl_load(); if (rc) return rc;
    CK(cudaSetDevice(h->device));
    nccl_uid u; memcpy(u.internal, id, 128);
    nccl_comm_t c; NCK(N.CommInitRank(&c, h->size0, u, h->P.rank));
    h->nccl_main = c; h->nccl_comm = c;
    // second communicator for the particle stream: its id is made on rank 0 and broadcast over the first one
    fn_Bcast Bcast = (fn_Bcast)dlsym(N.lib, "ncclBroadcast");
    if (!Bcast) { tgpu_set_error("NCCL symbol missing: ncclBroadcast"); return TGPU_ENCCL; }
    nccl_uid u2; memset(&u2, 0, sizeof u2);
    if (h->P.rank == 0) NCK(N.GetUniqueId(&u2));
    uint8_t *d = (uint8_t *)h->d_small;                 // 512 bytes of device scratch
    CK(cudaMemcpyAsync(d, u2.internal, 128, cudaMemcpyHostToDevice, h->stream));
    NCK(Bcast(d, d, 128, /*ncclInt8*/ 0, 0, c, h->stream));
    CK(cudaMemcpyAsync(u2.internal, d, 128, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    nccl_comm_t c2; NCK(N.CommInitRank(&c2, h->size0, u2, h->P.rank));
    h->nccl_prt = c2;
    // NCCL opens its point-to-point connections lazily, at the first send / recv between two ranks: hundreds of
    // milliseconds that otherwise land in whichever lap (or phase timer) first migrates a particle to a diagonal
    // neighbour.  One 4-byte exchange with all eight neighbours on both communicators, here.
    for (int k = 0; k < 2; k++) {
        h->nccl_comm = k ? h->nccl_prt : h->nccl_main;
        rc = comm_group_begin(h); if (rc) return rc;
        for (int cdir = 0; cdir < 9; cdir++) {
            if (cdir == 4) continue;
            const int da = cdir % 3 - 1, db = cdir / 3 - 1;
            const int to = topo_neighbour2(h, da, db), from = topo_neighbour2(h, -da, -db);
            comm_send(h, h->d_small + cdir, 4, to);
            comm_recv(h, h->d_small + 16 + cdir, 4, from);
        }
        rc = comm_group_end(h); if (rc) return rc;
    }
    h->nccl_comm = h->nccl_main;
    CK(cudaStreamSynchronize(h->stream));
    return comm_peer_setup(h);
}

// ---------------------------------------------------------------------------------------------
// Peer-memory transport for the field-side halo exchanges (C1, C2, C3 of SURVEY 2b: ghost refresh, current fold,
// filter deep halo -- fieldboundaries.F90:1179-1204, 1319-1344, 1652-1677, 1990-2185; optimized_filters.F90:1665-1674,
// 1838-1847).  One process per GPU stays; every rank exports its nine field arrays and a few signal words with cudaIpc,
// the handles travel once over NCCL (all-gather), and each rank maps those of its axis neighbours.  After that a halo
// step needs no NCCL call and no pack / unpack buffer: the consumer's kernel reads the neighbour's layers over NVLink
// (fields.cu halo_step).  All ranks switch together or not at all (a second all-gather of the outcome).
// ---------------------------------------------------------------------------------------------
typedef int (*fn_AllGather)(const void *, void *, size_t, int, nccl_comm_t, cudaStream_t);

// Tear-down handshake.  A neighbour's last k_sig_sync may still be polling my signal words when my own last lap is
// complete (it sets its PULLED, which releases me, and only then reads mine).  So before anything is unmapped or freed:
// raise FIN in my words, wait (bounded: 3 s) until every mapped neighbour has raised its own -- it does so after
// synchronising its streams -- and then give the slower side 20 ms to finish the read that saw my FIN.
static void peer_close(tgpu_ctx *h)
{
    if (h->peer && h->sig) {
        const uint32_t one = 1;
        cudaMemcpy(h->sig + TGPU_SIG_FIN, &one, sizeof one, cudaMemcpyHostToDevice);
        for (int r = 0; r < h->size0; r++) {
            if (!h->peer[r].open || !h->peer[r].sig) continue;
            for (int tries = 0; tries < 6000; tries++) {
                uint32_t v = 0;
                if (cudaMemcpy(&v, h->peer[r].sig + TGPU_SIG_FIN, sizeof v, cudaMemcpyDeviceToHost) != cudaSuccess) { cudaGetLastError(); break; }
                if (v) break;
                usleep(500);
            }
        }
        usleep(20000);
    }
    if (h->peer) {
        for (int r = 0; r < h->size0; r++)
            if (h->peer[r].open) for (int a = 0; a < 10; a++) if (h->peer[r].base[a]) cudaIpcCloseMemHandle(h->peer[r].base[a]);
        delete[] h->peer; h->peer = nullptr;
    }
}

int comm_peer_setup(tgpu_ctx *h)
{
    h->peer = nullptr; h->xseq = 0;
    const char *off = getenv("TGPU_NO_PEER");
    int want = h->opt_peer && !(off && off[0] == '1');
    fn_AllGather AllGather = (fn_AllGather)dlsym(N.lib, "ncclAllGather");
    if (!AllGather) want = 0;
    const int n = h->size0;
    struct Pack { cudaIpcMemHandle_t hd[10]; int ok; int pad[15]; };
    Pack mine; memset(&mine, 0, sizeof mine);
    mine.ok = want;
    if (want && !h->sig) {
        if (cudaMalloc((void **)&h->sig, 64 * sizeof(uint32_t)) != cudaSuccess || cudaMemset(h->sig, 0, 64 * sizeof(uint32_t)) != cudaSuccess) { cudaGetLastError(); mine.ok = 0; }
    }
    for (int a = 0; a < 9 && mine.ok; a++) if (cudaIpcGetMemHandle(&mine.hd[a], h->f[a]) != cudaSuccess) { cudaGetLastError(); mine.ok = 0; }
    if (mine.ok && cudaIpcGetMemHandle(&mine.hd[9], h->sig) != cudaSuccess) { cudaGetLastError(); mine.ok = 0; }
    if (!AllGather) return 0;                                 // (every rank loads the same NCCL: same decision everywhere)
    Pack *dsend = nullptr, *drecv = nullptr;
    std::vector<Pack> all(n);
    CK(cudaMalloc((void **)&dsend, sizeof(Pack))); CK(cudaMalloc((void **)&drecv, sizeof(Pack) * n));
    auto gather = [&]() -> int {
        CK(cudaMemcpyAsync(dsend, &mine, sizeof(Pack), cudaMemcpyHostToDevice, h->stream));
        NCK(AllGather(dsend, drecv, sizeof(Pack), /*ncclInt8*/ 0, (nccl_comm_t)h->nccl_main, h->stream));
        CK(cudaMemcpyAsync(all.data(), drecv, sizeof(Pack) * n, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        return 0;
    };
    int rc = gather();
    if (rc) { cudaFree(dsend); cudaFree(drecv); return rc; }
    int all_ok = 1;
    for (int r = 0; r < n; r++) all_ok &= all[r].ok;
    if (all_ok) {
        h->peer = new PeerRank[n];
        memset(h->peer, 0, sizeof(PeerRank) * n);
        const tgpu_params &P = h->P;
        for (int dir = 0; dir < 6 && mine.ok; dir++) {
            const int axis = dir / 2;
            if (axis == 2 && P.dim == 2) continue;
            const int sz = axis == 0 ? P.sizex : axis == 1 ? P.sizey : P.sizez;
            if (sz == 1) continue;
            const int r = topo_neighbour(P.rank, P.sizex, P.sizey, P.sizez, dir);
            if (r == P.rank || h->peer[r].open) continue;
            for (int a = 0; a < 10 && mine.ok; a++) {
                void *ptr = nullptr;
                if (cudaIpcOpenMemHandle(&ptr, all[r].hd[a], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); mine.ok = 0; break; }
                h->peer[r].base[a] = ptr;
                if (a < 9) h->peer[r].f[a] = (float *)ptr; else h->peer[r].sig = (uint32_t *)ptr;
            }
            h->peer[r].open = 1;
        }
    } else mine.ok = 0;
    // second round: did every rank manage to map its neighbours?
    rc = gather();
    cudaFree(dsend); cudaFree(drecv);
    if (rc) return rc;
    all_ok = 1;
    for (int r = 0; r < n; r++) all_ok &= all[r].ok;
    if (!all_ok) peer_close(h);
    return 0;
}

int comm_peer_check(tgpu_ctx *h)
{
    if (!h->peer || !h->sig) return 0;
    uint32_t t = 0;
    CK(cudaMemcpy(&t, h->sig + TGPU_SIG_TIMEOUT, sizeof t, cudaMemcpyDeviceToHost));
    if (t) { tgpu_set_error("halo exchange: timed out waiting for a neighbouring rank"); return TGPU_ENCCL; }
    return 0;
}

int comm_destroy(tgpu_ctx *h)
{
    peer_close(h);
    if (h->sig) { cudaFree(h->sig); h->sig = nullptr; }
    if (h->nccl_main && N.CommDestroy) N.CommDestroy((nccl_comm_t)h->nccl_main);
    if (h->nccl_prt && N.CommDestroy) N.CommDestroy((nccl_comm_t)h->nccl_prt);
    h->nccl_comm = h->nccl_main = h->nccl_prt = nullptr;
    return 0;
}

int comm_group_begin(tgpu_ctx *h)
{
    if (!h->nccl_comm) { tgpu_set_error("communicator not initialised (tgpu_comm_init)"); return TGPU_ENCCL; }
    NCK(N.GroupStart());
    return 0;
}
int comm_group_end(tgpu_ctx *h) { NCK(N.GroupEnd()); h->launches++; return 0; }
int comm_send(tgpu_ctx *h, const void *buf, size_t bytes, int peer)
{
    NCK(N.Send(buf, bytes, /*ncclInt8*/ 0, peer, (nccl_comm_t)h->nccl_comm, h->stream));
    return 0;
}
int comm_recv(tgpu_ctx *h, void *buf, size_t bytes, int peer)
{
    NCK(N.Recv(buf, bytes, 0, peer, (nccl_comm_t)h->nccl_comm, h->stream));
    return 0;
}
int comm_sendrecv(tgpu_ctx *h, const void *sbuf, size_t sbytes, int dst, void *rbuf, size_t rbytes, int src)
{
    int rc = comm_group_begin(h); if (rc) return rc;
    rc = comm_send(h, sbuf, sbytes, dst); if (rc) return rc;
    rc = comm_recv(h, rbuf, rbytes, src); if (rc) return rc;
    return comm_group_end(h);
}
