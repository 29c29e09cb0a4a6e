"""world_size-2 test of the slab-exchange protocol on CPU (torch.distributed, gloo).

Each process owns ONE rank of a two-rank decomposition: it runs the per-rank phases of a lap on its slab and performs
every exchange (ghost refresh, current fold, filter halos, particle migration incl. the second corner round) through
tests/slabs.py plans over gloo send/recv.  The result must equal, bit for bit, the same
rank of the in-process multi-rank oracle world (which moves the same boxes with memcpy)."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle as O
import pic_testlib as T


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _i3(v):
    return (C.c_int * 3)(*v)


class Driver:
    def __init__(self, slabs, w, rank):
        self.w, self.r, self.rank = w, w.ranks[rank], rank
        P = w.P
        r = self.r
        self.slab = slabs.Slab(P.dim, rank, (P.sizex, P.sizey, P.sizez), (r.mx, r.my, r.mz), r.nghost, r.nghostz,
                               (P.periodicx, P.periodicy, P.periodicz), P.ntimes)
        self.L = O.lib()

    def sendrecv(self, sbuf, to, rbuf, frm, tag):
        if to == self.rank and frm == self.rank:
            rbuf[...] = sbuf
            return
        reqs = [dist.isend(torch.from_numpy(sbuf), to, tag=tag), dist.irecv(torch.from_numpy(rbuf), frm, tag=tag)]
        for q in reqs:
            q.wait()

    def box_get(self, arr, lo, hi):
        n = int(np.prod([h - l + 1 for l, h in zip(lo, hi)]))
        buf = np.empty(n, np.float32)
        self.L.orc_box_get(self.r.h, arr, _i3(lo), _i3(hi), buf.ctypes.data_as(C.POINTER(C.c_float)))
        return buf

    def run_plan(self, plan, tag0=0):
        for n, sh in enumerate(plan):
            to, frm = self.slab.neighbour(sh.axis, sh.direction), self.slab.neighbour(sh.axis, -sh.direction)
            sb = np.concatenate([self.box_get(a, sh.src_lo, sh.src_hi) for a in sh.arrays])
            rb = np.empty_like(sb)
            self.sendrecv(sb, to, rb, frm, tag0 + n)
            if sh.recv_ok:
                cnt = sb.size // len(sh.arrays)
                fn = self.L.orc_box_put if sh.mode == "put" else self.L.orc_box_add
                for i, a in enumerate(sh.arrays):
                    fn(self.r.h, a, _i3(sh.dst_lo), _i3(sh.dst_hi), rb[i * cnt:(i + 1) * cnt].ctypes.data_as(C.POINTER(C.c_float)))

    def filter(self):
        P = self.w.P
        if P.filter_kind == 1:
            for _ in range(P.ntimes):
                self.run_plan(self.slab.plan_filter1_refresh(), 500)
                self.r.call("filter1_pass")
            return
        for c in range(3):
            for axis in range(self.slab.naxes):
                (ulo, uhi), (dlo, dhi) = self.slab.filter2_halo_boxes(axis)
                up, dn = self.slab.neighbour(axis, +1), self.slab.neighbour(axis, -1)
                s_up, s_dn = self.box_get(6 + c, ulo, uhi), self.box_get(6 + c, dlo, dhi)
                glo, ghi = np.empty_like(s_up), np.empty_like(s_dn)
                self.sendrecv(s_up, up, glo, dn, 700)
                self.sendrecv(s_dn, dn, ghi, up, 701)
                fp = C.POINTER(C.c_float)
                self.L.orc_filter2_rank(self.r.h, c, axis, glo.ctypes.data_as(fp), ghi.ctypes.data_as(fp))

    def exchange_particles(self):
        for d in self.slab.migration_directions():
            axis, direction = d // 2, (1 if d % 2 else -1)
            to, frm = self.slab.neighbour(axis, direction), self.slab.neighbour(axis, -direction)
            ni, ne = C.c_int(), C.c_int()
            addr = self.L.orc_rank_box(self.r.h, 0, d, C.byref(ni), C.byref(ne))
            n = ni.value + ne.value
            out = np.frombuffer((C.c_char * (max(n, 1) * 40)).from_address(addr), dtype=np.uint8)[:n * 40].copy()
            cnt_s, cnt_r = np.array([ni.value, ne.value], np.int64), np.zeros(2, np.int64)
            self.sendrecv(cnt_s, to, cnt_r, frm, 900 + d)
            nin = int(cnt_r.sum())
            rb = np.empty(nin * 40, np.uint8)
            if to == self.rank:
                rb[...] = out
            else:
                reqs = []
                if n:
                    reqs.append(dist.isend(torch.from_numpy(out), to, tag=950 + d))
                if nin:
                    reqs.append(dist.irecv(torch.from_numpy(rb), frm, tag=950 + d))
                for q in reqs:
                    q.wait()
            iaddr = self.L.orc_rank_box(self.r.h, 1, d ^ 1, None, None)
            C.memmove(iaddr, rb.ctypes.data, nin * 40)
            self.L.orc_rank_box_set_counts(self.r.h, 1, d ^ 1, int(cnt_r[0]), int(cnt_r[1]))

    def lap(self, lap):
        r, s = self.r, self.slab
        bcb = lambda: self.run_plan(s.plan_ghost_refresh((3, 4, 5)), 100)
        bce = lambda: self.run_plan(s.plan_ghost_refresh((0, 1, 2)), 200)
        bcb(); bce(); r.call("advance_b_halfstep"); bcb(); r.call("move_particles"); r.call("advance_b_halfstep"); bcb(); bcb()
        r.call("advance_e_fullstep"); bce(); r.call("reset_currents"); bce(); bcb(); r.call("deposit_particles")
        self.exchange_particles(); self.run_plan(s.plan_current_fold(), 300); self.filter(); r.call("add_current")
        r.call("inject_others"); self.exchange_particles(); r.call("inject_others")
        if lap % 10 == 0:
            r.call("reorder_particles")


def _worker(rank, world, port, case, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import slabs
        kw = dict(ppc=3.0, ntimes=3, delgam=0.05)
        kw.update(case)
        mine = T.oracle_world(**kw)          # this process drives rank `rank` of this copy over gloo
        ref = T.oracle_world(**kw)           # in-process multi-rank oracle
        drv = Driver(slabs, mine, rank)
        for lap in range(1, 4):
            drv.lap(lap)
            ref.step()
        a, b = mine.ranks[rank], ref.ranks[rank]
        ok = a.counts == b.counts
        for k in range(9):
            ok = ok and np.array_equal(a.arr(k), b.arr(k))
        ok = ok and np.array_equal(T.sort_particles(a.ions().copy()), T.sort_particles(b.ions().copy()))
        ok = ok and np.array_equal(T.sort_particles(a.lecs().copy()), T.sort_particles(b.lecs().copy()))
        moved = int((a.ions()["proc"] != rank).sum())
        q.put((rank, bool(ok), moved))
    finally:
        dist.destroy_process_group()


CASES = [dict(dim=3, order=2, n=(8, 8, 12), sizes=(1, 1, 2), filter_kind=2),
         dict(dim=3, order=1, n=(8, 12, 8), sizes=(1, 2, 1), filter_kind=1),
         dict(dim=2, order=2, n=(12, 16, 1), sizes=(1, 2, 1), filter_kind=1),
         dict(dim=2, order=1, n=(16, 10, 1), sizes=(2, 1, 1), filter_kind=1)]


@pytest.mark.parametrize("case", CASES, ids=["3d-z-f2", "3d-y-f1", "2d-y", "2d-x"])
def test_two_rank_lap_over_gloo(case):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, case, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    assert sum(m for _, _, m in res) > 0, "no particle migrated between the two ranks"


# ---- the same protocol against WHOLE LAPS OF THE REFERENCE'S OWN MAINLOOP (tests/golden/ref_lap.npz, two-rank cases) ----
def _golden_worker(rank, world, port, case, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import slabs
        z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_lap.npz"))
        key = f"l{case}"
        dim, order, px, py, pz, nx, ny, nz = (int(v) for v in z[key + "_meta"])
        sx, sy, sz, maxhlf, nsp, laps, highorder, shock, fkind = (int(v) for v in z[key + "_geom"])
        par = z[key + "_par"]
        P = O.make_params(dim=dim, order=order, mx0=nx, my0=ny, mz0=nz, sizex=sx, sizey=sy, sizez=sz, ntimes=2, filter_kind=fkind,
                          periodic=(px, py, pz), maxptl=2 * maxhlf, highorder=highorder)
        P.qi, P.qe, P.qmi, P.qme = (float(v) for v in par[5:9])
        w = O.World(P)
        r = w.ranks[rank]                                   # this process drives only its own rank of the world
        for a in range(6):
            r.arr(a)[...] = z[f"{key}_r{rank}_in{a}"]
        for a in range(6, 9):
            r.arr(a)[...] = 0
        r.particles()[:] = z[f"{key}_r{rank}_pin"]
        r.set_counts(nsp, nsp)
        drv = Driver(slabs, w, rank)
        for lap in range(1, laps + 1):
            drv.lap(lap)
        ions, lecs = (int(v) for v in z[f"{key}_r{rank}_counts"])
        ok = r.counts == (ions, lecs)
        for a in range(9):
            ok = ok and np.array_equal(r.arr(a), z[f"{key}_r{rank}_out{a}"])
        pout, p = z[f"{key}_r{rank}_pout"], r.particles()
        ok = ok and np.array_equal(p[:ions], pout[:ions]) and np.array_equal(p[maxhlf:maxhlf + lecs], pout[maxhlf:maxhlf + lecs])
        q.put((rank, bool(ok), int((pout["proc"][:ions] != rank).sum())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case", [4, 15, 16], ids=["3d-o3-highorder-1x2x1", "2d-o2-2x1", "3d-o2-filter2-1x1x2"])
def test_two_rank_protocol_against_the_reference_mainloop(case):
    """the slab-exchange plan (tests/slabs.py = the protocol csrc/fields.cu and csrc/particles.cu run over NCCL / peer
    memory), driven over gloo by two processes, against two laps of the reference's mainloop run from its own text:
    every array and the particle arrays in order, BIT-EXACT"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_golden_worker, args=(r, 2, port, case, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    assert sum(m for _, _, m in res) > 0, "no particle migrated between the two ranks"
