"""Shared helpers: build an oracle world and a GPU context from the same parameters and compare them.
The oracle (oracle/) is the checker; the GPU library is the thing under test."""
import ctypes as C

import numpy as np

from oracle import oracle as O


def oracle_world(dim=3, order=2, n=(16, 16, 16), sizes=(1, 1, 1), ppc=4.0, ntimes=0, filter_kind=1, delgam=1e-2,
                 quirks=O.Q_REFERENCE, init="weibel", seed_fields=1, periodic=(1, 1, 1), ext=None, pusher=0, gamma0=0.5,
                 highorder=0, wall_i2=0, charges=None):
    P = O.make_params(dim=dim, order=order, mx0=n[0], my0=n[1], mz0=n[2], sizex=sizes[0], sizey=sizes[1], sizez=sizes[2],
                      ppc0=ppc if ppc > 0 else 16.0, ntimes=ntimes, maxptl=None if ppc > 0 else 65536, filter_kind=filter_kind, quirks=quirks, periodic=periodic, ext=ext,
                      pusher=pusher, gamma0=gamma0, highorder=highorder, wall_i2=wall_i2)
    if charges is not None:
        P.qi, P.qe = charges
    w = O.World(P)
    if init == "weibel":
        w.init_weibel(ppc0=ppc, delgam=delgam, distr_dim=3 if dim == 3 else 2, gamma0=gamma0)
    elif init == "uniform":
        w.init_uniform(ppc0=ppc, beta=0.5, uth=0.2, seed=7)
    elif init == "twostream":
        w.init_twostream(ppc0=ppc, delgam=delgam)
    if seed_fields:
        rng = np.random.default_rng(seed_fields)
        for r in w.ranks:
            for a in range(6):
                r.arr(a)[...] = (rng.standard_normal(r.arr(a).shape) * 0.02).astype(np.float32)
    return w


def gpu_params(tg, w, rank=0, device=-1):
    """tgpu_params for oracle rank `rank` of world w (same numbers the Fortran globals would hold)."""
    P = w.P
    gp = tg.make_params(dim=P.dim, order=P.order, mx0=P.mx0, my0=P.my0, mz0=P.mz0, sizex=P.sizex, sizey=P.sizey,
                        sizez=P.sizez, rank=rank, c=P.c, corr=P.corr, ntimes=P.ntimes, filter_kind=P.filter_kind,
                        periodic=(P.periodicx, P.periodicy, P.periodicz), maxptl=P.maxptl, buffsize=P.buffsize,
                        quirks=P.quirks, pusher=P.pusher, ext=list(P.ext) if P.external_fields else None, device=device,
                        highorder=P.highorder, wall_i2=P.wall_i2)
    gp.qi, gp.qe, gp.qmi, gp.qme = P.qi, P.qe, P.qmi, P.qme
    r = w.ranks[rank]
    assert (gp.mx, gp.my, gp.mz, gp.mxcum, gp.mycum, gp.mzcum) == (r.mx, r.my, r.mz, r.mxcum, r.mycum, r.mzcum)
    return gp


def upload(ctx, r):
    ctx.fields_h2d(*[np.ascontiguousarray(a) for a in r.fields()])
    ctx.currents_h2d(*[np.ascontiguousarray(a) for a in r.currents()])
    ions, lecs = r.counts
    ctx.particles_h2d(r.particles(), ions, lecs)


def interior(r, a, extra=0):
    g, gz = r.nghost // 2, r.nghostz // 2
    if r.mz == 1:
        return a[:, g + extra:r.my - g - 1 - extra, g + extra:r.mx - g - 1 - extra]
    return a[gz + extra:r.mz - gz - 1 - extra, g + extra:r.my - g - 1 - extra, g + extra:r.mx - g - 1 - extra]


def refreshed(r, a):
    """interior plus the ghosts a bc_* call refreshes: 1..m-1 (index m is never refreshed)."""
    if r.mz == 1:
        return a[:, :r.my - 1, :r.mx - 1]
    return a[:r.mz - 1, :r.my - 1, :r.mx - 1]


def sort_particles(p):
    """order-independent comparison key: (proc, ind) is unique (particles.F90:2799-2801)."""
    idx = np.lexsort((p["ind"], p["proc"]))
    return p[idx]


def gpu_particles(ctx):
    p, ions, lecs = ctx.particles_d2h()
    return sort_particles(p[:ions].copy()), sort_particles(p[ctx.maxhlf:ctx.maxhlf + lecs].copy())


def oracle_particles(r):
    return sort_particles(r.ions().copy()), sort_particles(r.lecs().copy())


def max_rel(a, b):
    scale = max(float(np.abs(b).max()), 1e-30)
    return float(np.abs(a.astype(np.float64) - b.astype(np.float64)).max()) / scale


def assert_particles_close(pg, po, rtol_pos=2e-6, rtol_mom=2e-5, what="", extent=1.0):
    """extent: size of the box along the longest axis.  A particle that is wrapped after the push carries the round-off of
    its pre-wrap coordinate (1 ulp at x = 131 is 4e-6 of x = 3), so positions are judged against max(|x|, extent)."""
    assert pg.size == po.size, f"{what}: particle count {pg.size} != {po.size}"
    assert np.array_equal(pg["ind"], po["ind"]) and np.array_equal(pg["proc"], po["proc"]), f"{what}: identity mismatch"
    if pg.size == 0:
        return
    for k in ("x", "y", "z"):
        d = np.abs(pg[k].astype(np.float64) - po[k]) / np.maximum(np.abs(po[k]), max(extent, 1.0))
        assert d.max() <= rtol_pos, f"{what}: {k} rel err {d.max():.3e}"
    scale = max(np.abs(po["u"]).max(), np.abs(po["v"]).max(), np.abs(po["w"]).max(), 1e-30)
    for k in ("u", "v", "w"):
        d = np.abs(pg[k].astype(np.float64) - po[k]) / scale
        assert d.max() <= rtol_mom, f"{what}: {k} err/scale {d.max():.3e}"
    assert np.array_equal(pg["ch"], po["ch"]) and np.array_equal(pg["splitlev"], po["splitlev"])


def gross_current(w, rank=0):
    """Magnitude of ONE species' current per cell: |q| * (particles of a species per interior cell) * mean |u|/gamma * c.
    Counter-streaming beams (Weibel, two-stream) and co-drifting species cancel almost completely in the net current, so
    an error bar "relative to the max-norm of the net current" measures the cancellation, not the deposit.  The
    round-off of a deposit (different summation order, atomics) is proportional to the gross current."""
    r = w.ranks[rank]
    ions, lecs = r.counts
    g, gz = r.nghost // 2, r.nghostz // 2
    ncell = (r.mx - r.nghost) * (r.my - r.nghost) * ((r.mz - r.nghostz) if r.mz > 1 else 1)
    p = r.lecs() if lecs else r.ions()
    n = max(lecs if lecs else ions, 1)
    u, v, ww = (p[k].astype(np.float64) for k in ("u", "v", "w"))
    gam = np.sqrt(1.0 + u * u + v * v + ww * ww)
    vmean = max(float(np.abs(u / gam).mean()), float(np.abs(v / gam).mean()), float(np.abs(ww / gam).mean()))
    return abs(float(w.P.qe)) * (n / ncell) * vmean * float(w.P.c)


def max_abs_diff(a, b):
    return float(np.abs(a.astype(np.float64) - b.astype(np.float64)).max())
