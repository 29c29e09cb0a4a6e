"""Multi-GPU parity: N processes, one GPU each, slabs exchanged over NCCL inside libtristan_gpu.so; every rank is
compared with the same rank of the in-process multi-rank oracle.  Needs >= 2 GPUs (skipped otherwise)."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, case, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import tristan_mp_pu_master_densdecomp_b200 as tg
        import pic_testlib as T
        from oracle import oracle as O
        kw = dict(ppc=4.0, ntimes=3, delgam=0.05)
        kw.update(case)
        peer = kw.pop("peer", 1)
        w = T.oracle_world(**kw)
        r = w.ranks[rank]
        ctx = tg.Context(T.gpu_params(tg, w, rank=rank, device=rank))
        ctx.set_option("peer", peer)                 # 1: halos over cudaIpc peer memory; 0: NCCL send / recv
        ctx.comm_init_torch()
        if ctx.halo_transport() != peer:
            raise RuntimeError(f"halo transport {ctx.halo_transport()} != requested {peer} (cudaIpc mapping of the neighbours failed?)")
        T.upload(ctx, r)
        err = ""
        for lap in range(3):
            ctx.step(1); w.step()
            fg = ctx.fields_d2h()
            for a in range(6):
                e = T.max_rel(T.interior(r, fg[a]), T.interior(r, r.arr(a)))
                if e > 4e-4 * (lap + 1):
                    err += f"lap {lap} rank {rank} {O.ARR_NAMES[a]} err {e:.2e}; "
            if ctx.counts() != r.counts:
                err += f"lap {lap} rank {rank} counts {ctx.counts()} != {r.counts}; "
            else:
                gi, ge = T.gpu_particles(ctx)
                oi, oe = T.oracle_particles(r)
                try:
                    T.assert_particles_close(gi, oi, rtol_pos=3e-5 * (lap + 1), rtol_mom=3e-4 * (lap + 1))
                    T.assert_particles_close(ge, oe, rtol_pos=3e-5 * (lap + 1), rtol_mom=3e-4 * (lap + 1))
                except AssertionError as ex:
                    err += f"lap {lap} rank {rank} particles: {ex}; "
        moved = int((r.ions()["proc"] != rank).sum())
        ctx.close()
        q.put((rank, err, moved))
    except Exception as ex:  # noqa
        import traceback
        q.put((rank, "EXC " + traceback.format_exc()[-1500:], 0))
    finally:
        dist.destroy_process_group()


CASES = {
    2: [dict(dim=3, order=2, n=(12, 12, 16), sizes=(1, 1, 2), filter_kind=2),
        dict(dim=3, order=1, n=(12, 16, 12), sizes=(1, 2, 1), filter_kind=1),
        dict(dim=3, order=3, n=(12, 12, 16), sizes=(1, 1, 2), filter_kind=2),
        dict(dim=2, order=2, n=(16, 16, 1), sizes=(2, 1, 1), filter_kind=1),
        # open (radiating) x: bc_b2 / bc_e2 run `surface`, leavers through the x faces are discarded
        dict(dim=3, order=2, n=(16, 12, 16), sizes=(1, 1, 2), filter_kind=2, periodic=(0, 1, 1)),
        dict(dim=2, order=1, n=(24, 16, 1), sizes=(2, 1, 1), filter_kind=1, periodic=(0, 1, 1)),
        # 4th-order field solver across a slab boundary
        dict(dim=3, order=2, n=(12, 12, 16), sizes=(1, 1, 2), filter_kind=2, highorder=1),
        # the same exchanges through NCCL send / recv (peer memory switched off); uneven split (last rank takes the rest)
        dict(dim=3, order=2, n=(12, 12, 16), sizes=(1, 1, 2), filter_kind=2, peer=0),
        dict(dim=3, order=2, n=(12, 15, 12), sizes=(1, 2, 1), filter_kind=1),
        dict(dim=3, order=2, n=(12, 12, 17), sizes=(1, 1, 2), filter_kind=2)],
    4: [dict(dim=3, order=2, n=(12, 16, 16), sizes=(1, 2, 2), filter_kind=2),
        dict(dim=2, order=1, n=(16, 16, 1), sizes=(2, 2, 1), filter_kind=1),
        dict(dim=3, order=2, n=(16, 16, 16), sizes=(1, 2, 2), filter_kind=2, periodic=(0, 1, 1)),
        dict(dim=3, order=2, n=(12, 16, 16), sizes=(1, 2, 2), filter_kind=1, peer=0)],
    8: [dict(dim=3, order=2, n=(12, 16, 32), sizes=(1, 2, 4), filter_kind=2),
        dict(dim=3, order=3, n=(12, 16, 32), sizes=(1, 2, 4), filter_kind=2),
        dict(dim=3, order=1, n=(12, 16, 32), sizes=(1, 4, 2), filter_kind=1)],
}


def _run(world, case):
    import tristan_mp_pu_master_densdecomp_b200 as tg
    if tg.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, case, q)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        res = [q.get(timeout=180) for _ in range(world)]
    finally:
        for p in procs:
            p.join(timeout=20)
        for p in procs:
            if p.is_alive():                     # a rank that hangs must not take the whole session with it
                p.kill()
    errs = [e for _, e, _ in res if e]
    assert not errs, errs
    assert sum(m for _, _, m in res) > 0, "no particle migrated"


@pytest.mark.parametrize("case", CASES[2], ids=["3d-z", "3d-y", "3d-z-o3", "2d-x", "3d-z-openx", "2d-x-openx", "3d-z-highorder",
                                                  "3d-z-nccl", "3d-y-uneven", "3d-z-uneven"])
def test_two_gpus(tg, case):
    _run(2, case)


@pytest.mark.parametrize("case", CASES[4], ids=["3d-yz", "2d-xy", "3d-yz-openx", "3d-yz-nccl"])
def test_four_gpus(tg, case):
    _run(4, case)


@pytest.mark.parametrize("case", CASES[8], ids=["3d-2x4", "3d-2x4-o3", "3d-4x2-o1"])
def test_eight_gpus(tg, case):
    _run(8, case)


# ---------------------------------------------------------------------------------------------------------------
# the same N-process set-up against WHOLE LAPS OF THE REFERENCE'S OWN MAINLOOP (tests/golden/ref_lap.npz: every rank of the
# reference ran tristanmainloop.F90 from its source text, MPI_SendRecv as a rendezvous; see tests/test_ref_golden.py)
# ---------------------------------------------------------------------------------------------------------------
def _golden_worker(rank, world, port, case, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import tristan_mp_pu_master_densdecomp_b200 as tg
        import pic_testlib as T
        z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_lap.npz"))
        key = f"l{case}"
        dim, order, px, py, pz, nx, ny, nz = (int(v) for v in z[key + "_meta"])
        sx, sy, sz, maxhlf, nsp, laps, highorder, shock, fkind = (int(v) for v in z[key + "_geom"])
        assert sx * sy * sz == world
        par = z[key + "_par"]
        P = tg.make_params(dim=dim, order=order, mx0=nx, my0=ny, mz0=nz, sizex=sx, sizey=sy, sizez=sz, rank=rank, periodic=(px, py, pz),
                           maxptl=2 * maxhlf, device=rank, ntimes=2, filter_kind=fkind, highorder=highorder)
        P.qi, P.qe, P.qmi, P.qme = (float(v) for v in par[5:9])
        ctx = tg.Context(P)
        ctx.comm_init_torch()
        ctx.fields_h2d(*[np.ascontiguousarray(z[f"{key}_r{rank}_in{a}"]) for a in range(6)])
        zero = np.zeros_like(z[f"{key}_r{rank}_in0"])
        ctx.currents_h2d(zero, zero, zero)
        ctx.particles_h2d(np.ascontiguousarray(z[f"{key}_r{rank}_pin"]), nsp, nsp)
        if shock:
            ctx.set_user_hooks(1, [float(v) for v in par[:5]])
        ctx.step(laps)
        err = ""
        ions, lecs = (int(v) for v in z[f"{key}_r{rank}_counts"])
        got = ctx.fields_d2h()
        g, gz = P.nghost // 2, (P.nghostz // 2 if dim == 3 else 0)
        for a in range(6):
            ref = z[f"{key}_r{rank}_out{a}"]
            sl = (slice(gz, ref.shape[0] - gz - 1) if dim == 3 else slice(None), slice(g, ref.shape[1] - g - 1), slice(g, ref.shape[2] - g - 1))
            e = T.max_rel(got[a][sl], ref[sl])
            if e > 4e-4 * laps:
                err += f"rank {rank} field {a} err {e:.2e}; "
        if ctx.counts() != (ions, lecs):
            err += f"rank {rank} counts {ctx.counts()} != {(ions, lecs)}; "
        else:
            gp, _, _ = ctx.particles_d2h()
            pout = z[f"{key}_r{rank}_pout"]
            ext = float(max(P.mx, P.my, P.mz))
            try:
                T.assert_particles_close(T.sort_particles(gp[:ions].copy()), T.sort_particles(pout[:ions].copy()), rtol_pos=3e-5 * laps,
                                         rtol_mom=3e-4 * laps, what="ions", extent=ext)
                T.assert_particles_close(T.sort_particles(gp[maxhlf:maxhlf + lecs].copy()), T.sort_particles(pout[maxhlf:maxhlf + lecs].copy()),
                                         rtol_pos=3e-5 * laps, rtol_mom=3e-4 * laps, what="lecs", extent=ext)
            except AssertionError as ex:
                err += f"rank {rank} particles: {ex}; "
        moved = int((z[f"{key}_r{rank}_pout"]["proc"][:ions] != rank).sum())
        ctx.close()
        q.put((rank, err, moved))
    except Exception:  # noqa
        import traceback
        q.put((rank, "EXC " + traceback.format_exc()[-1500:], 0))
    finally:
        dist.destroy_process_group()


def _run_golden(world, case):
    import tristan_mp_pu_master_densdecomp_b200 as tg
    if tg.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_golden_worker, args=(r, world, port, case, q)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        res = [q.get(timeout=180) for _ in range(world)]
    finally:
        for p in procs:
            p.join(timeout=20)
        for p in procs:
            if p.is_alive():
                p.kill()
    errs = [e for _, e, _ in res if e]
    assert not errs, errs
    assert sum(m for _, _, m in res) > 0, "no particle migrated"


@pytest.mark.parametrize("case", [3, 4], ids=["2d-shock-2x1", "3d-o3-highorder-1x2x1"])
def test_two_gpus_against_the_reference_mainloop(tg, case):
    _run_golden(2, case)


@pytest.mark.parametrize("case", [0, 1, 10], ids=["2d-2x2-10laps-reorder", "3d-o2-1x2x2", "3d-o1-1x2x2-filter2"])
def test_four_gpus_against_the_reference_mainloop(tg, case):
    _run_golden(4, case)
