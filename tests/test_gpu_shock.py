"""the shock problem's user hooks (user/user_shock.F90) on the device -- conductor / upstream
field clamps, the reflecting wall with its two zigzag deposits, open-x particle removal -- against the oracle."""
import numpy as np
import pytest

from oracle import oracle as O
import pic_testlib as T

pytestmark = pytest.mark.gpu

HOOK = (20.0, 0.05, 0.3, 1.2, 0.4)      # leftwall, binit, btheta, bphi, beta


def make(tg, dim, order, kind=1):
    n = (40, 12, 10) if dim == 3 else (48, 16, 1)
    w = T.oracle_world(dim=dim, order=order, n=n, ppc=4.0, ntimes=2, filter_kind=kind, periodic=(0, 1, 1), delgam=0.05, gamma0=0.4)
    # the shock setup has plasma only upstream of the wall (user_shock.F90:262-300): squeeze the load into [wall+0.5, x2in)
    r = w.ranks[0]
    g = r.nghost // 2
    lo, hi = HOOK[0] + 0.5, r.mx - g - 0.1
    for p in (r.ions(), r.lecs()):
        p["x"] = (lo + (p["x"] - (g + 1)) * (hi - lo) / (r.mx - 2 * g - 1)).astype(np.float32)
    ctx = tg.Context(T.gpu_params(tg, w, device=0))
    T.upload(ctx, r)
    return w, ctx


@pytest.mark.parametrize("dim", [2, 3])
def test_field_clamps_bit_exact(tg, dim):
    w, ctx = make(tg, dim, 2)
    r = w.ranks[0]
    ctx.field_bc_user_shock(*HOOK)
    r.call("field_bc_shock", *HOOK)
    fg = ctx.fields_d2h()
    for a in range(6):
        assert np.array_equal(fg[a], r.arr(a)), O.ARR_NAMES[a]
    assert np.all(fg[1][..., :int(HOOK[0] - 10)] == 0) and np.all(fg[2][..., :int(HOOK[0] - 10)] == 0)
    # (with nghost = 7 the right-edge clamp never fires: iloc(mx0-2) and iloc(mx0) both clip to mx, user_shock.F90:359-363)
    ctx.close()


@pytest.mark.parametrize("dim,order", [(2, 1), (3, 2), (3, 3)])
def test_reflecting_wall(tg, dim, order):
    w, ctx = make(tg, dim, order)
    ctx.set_option("fused", 0)
    r = w.ranks[0]
    # throw a third of the particles at the wall
    for p in (r.ions(), r.lecs()):
        sel = np.arange(p.size) % 3 == 0
        p["x"][sel] = HOOK[0] - 0.2 * np.random.default_rng(1).random(sel.sum()).astype(np.float32)
        p["u"][sel] = -np.abs(p["u"][sel]) - 0.3
    T.upload(ctx, r)
    ctx.reset_currents(); r.call("reset_currents")
    ctx.particle_bc_user_wall(HOOK[0]); r.call("particle_bc_wall", HOOK[0])
    cg = ctx.currents_d2h()
    for c in range(3):
        assert T.max_rel(cg[c], r.arr(6 + c)) < 2e-5
    gi, ge = T.gpu_particles(ctx)
    oi, oe = T.oracle_particles(r)
    assert (oi["u"] > 0).sum() > oi.size // 4          # they were reflected
    T.assert_particles_close(gi, oi, rtol_pos=2e-6, rtol_mom=2e-6)
    T.assert_particles_close(ge, oe, rtol_pos=2e-6, rtol_mom=2e-6)
    ctx.close()


@pytest.mark.parametrize("dim,order,kind", [(2, 1, 1), (3, 2, 2), (3, 3, 1)])
def test_shock_laps_open_x(tg, dim, order, kind):
    """open x boundaries (particles leaving [x1in, x2in] are removed), wall + clamps at mainloop's hook points"""
    w, ctx = make(tg, dim, order, kind)
    ctx.set_user_hooks(1, HOOK)
    r = w.ranks[0]
    n0 = sum(r.counts)
    for lap in range(4):
        ctx.step(1)
        w.call("step_shock", *HOOK)
        fg = ctx.fields_d2h()
        for a in range(6):
            err = T.max_rel(T.interior(r, fg[a]), T.interior(r, r.arr(a)))
            assert err < 4e-4 * (lap + 1), f"lap {lap} {O.ARR_NAMES[a]} err {err:.3e}"
        assert ctx.counts() == r.counts
    assert sum(r.counts) < n0, "no particle left through the open x boundary"
    gi, ge = T.gpu_particles(ctx)
    oi, oe = T.oracle_particles(r)
    T.assert_particles_close(gi, oi, rtol_pos=2e-4, rtol_mom=2e-3)
    ctx.close()
