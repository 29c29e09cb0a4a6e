"""The CUDA library against golden vectors produced by the REFERENCE'S OWN SOURCE TEXT (tests/golden/ref_*.npz, made by
tests/golden/make_ref_golden.py with the Fortran-subset interpreter f90run.py; see tests/test_ref_golden.py for the oracle's
side).  Nothing here touches the oracle: the GPU is compared with what the reference's Fortran computes.
Bars: stencils, radiation boundaries, edge fixes, ghost refresh, fold, filters -- BIT-EXACT; movers -- positions 2e-6 of
max(|x|, box), momenta 2e-5 of scale (FMA contraction on the device; SFU reciprocals in the cell-run movers)."""
import os

import numpy as np
import pytest

import pic_testlib as T

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLD, name))


def ctx_from_meta(tg, meta, **kw):
    dim, order, px, py, pz, nx, ny, nz = (int(v) for v in meta[:8])
    P = tg.make_params(dim=dim, order=order, mx0=nx, my0=ny, mz0=nz, periodic=(px, py, pz), maxptl=4096, device=0, **kw)
    return tg.Context(P), P


@pytest.mark.parametrize("case", range(7))
def test_yee_solver_against_the_reference_source(tg, case):
    z = load("ref_fields.npz")
    key = f"f{case}"
    ctx, P = ctx_from_meta(tg, z[key + "_meta"], ntimes=0)
    ctx.fields_h2d(*[np.ascontiguousarray(z[f"{key}_in{a}"]) for a in range(6)])
    ctx.currents_h2d(*[np.ascontiguousarray(z[f"{key}_in{a}"]) for a in range(6, 9)])
    ctx.advance_b_halfstep(); ctx.advance_e_fullstep(); ctx.advance_b_halfstep(); ctx.add_current()
    got = ctx.fields_d2h()
    for a in range(6):
        assert np.array_equal(got[a], z[f"{key}_out{a}"]), a
    ctx.close()


@pytest.mark.parametrize("case", range(6))
def test_radiation_boundaries_against_the_reference_source(tg, case):
    """surface + preledge / postedge; the GPU entry points also refresh the ghosts, the reference goldens do not -> compare the
    layers a refresh does not touch: everything when no axis is periodic-refreshed differently, i.e. the interior plus the faces
    the radiation routines write (whole array minus refreshed ghost layers)"""
    z = load("ref_radiation.npz")
    key = f"r{case}"
    meta = z[key + "_meta"]
    ctx, P = ctx_from_meta(tg, meta, ntimes=0)
    per = [int(v) for v in meta[2:5]]
    if any(per):
        ctx.close()
        pytest.skip("with a periodic axis tgpu_bc_b2 / _e2 refresh ghost layers the reference-side golden leaves alone; the all-open "
                    "cases compare whole arrays (the refresh of an open single-rank axis copies nothing)")
    ctx.fields_h2d(*[np.ascontiguousarray(z[f"{key}_in{a}"]) for a in range(6)])
    for si, name in enumerate(("pre_bc_b", "bc_b2", "post_bc_b", "pre_bc_e", "bc_e2", "post_bc_e")):
        getattr(ctx, name)()
        got = ctx.fields_d2h()
        for a in range(6):
            assert np.array_equal(got[a], z[f"{key}_s{si}_{a}"]), (name, a)
    ctx.close()


@pytest.mark.parametrize("case", range(4))
def test_ghost_refresh_and_fold_against_the_reference_source(tg, case):
    z = load("ref_halo.npz")
    key = f"h{case}"
    ctx, P = ctx_from_meta(tg, z[key + "_meta"], ntimes=0)
    ctx.fields_h2d(*[np.ascontiguousarray(z[f"{key}_in{a}"]) for a in range(6)])
    ctx.currents_h2d(*[np.ascontiguousarray(z[f"{key}_in{a}"]) for a in range(6, 9)])
    ctx.bc_b1(); ctx.bc_e1(); ctx.exchange_current()
    got = ctx.fields_d2h() + ctx.currents_d2h()
    for a in range(9):
        assert np.array_equal(got[a], z[f"{key}_out{a}"]), a
    ctx.close()


@pytest.mark.parametrize("case", range(4))
def test_filter1_against_the_reference_source(tg, case):
    z = load("ref_filter.npz")
    key = f"f1_{case}"
    meta = z[key + "_meta"]
    ctx, P = ctx_from_meta(tg, meta, ntimes=int(meta[8]), filter_kind=1)
    ctx.currents_h2d(*[np.ascontiguousarray(z[f"{key}_in{c}"]) for c in range(3)])
    ctx.apply_filter1_opt()
    got = ctx.currents_d2h()
    g, gz = P.nghost // 2, P.nghostz // 2
    for c in range(3):
        ref = z[f"{key}_out{c}"]
        sl = (slice(gz, P.mz - gz - 1) if P.dim == 3 else slice(None), slice(g, P.my - g - 1), slice(g, P.mx - g - 1))
        assert np.array_equal(got[c][sl], ref[sl]), c
    ctx.close()


@pytest.mark.parametrize("case", range(4))
def test_filter2_against_the_reference_source(tg, case):
    z = load("ref_filter.npz")
    key = f"f2_{case}"
    meta = z[key + "_meta"]
    ctx, P = ctx_from_meta(tg, meta, ntimes=int(meta[8]), filter_kind=2)
    zero = np.zeros_like(z[key + "_in"])
    ctx.currents_h2d(np.ascontiguousarray(z[key + "_in"]), zero, zero)
    ctx.apply_filter2_opt()
    got = ctx.currents_d2h()[0]
    ref = z[f"{key}_after{2 if P.dim == 3 else 1}"]
    g, gz = P.nghost // 2, P.nghostz // 2
    sl = (slice(gz, P.mz - gz - 1) if P.dim == 3 else slice(None), slice(g, P.my - g - 1), slice(g, P.mx - g - 1))
    assert np.array_equal(got[sl], ref[sl])
    ctx.close()


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("order", [0, 1, 2, 3])
@pytest.mark.parametrize("fused", [0, 1])
def test_movers_against_the_reference_source(tg, dim, order, fused):
    z = load("ref_mover.npz")
    key = f"m{dim}o{order}"
    ctx, P = ctx_from_meta(tg, z[key + "_meta"], ntimes=0)
    qm = float(z[key + "_qm"][0])
    P2 = ctx.P
    ctx.close()
    # the golden pushed one species with charge-to-mass qm: make the ions that species
    P2.qmi = qm
    ctx = tg.Context(P2)
    ctx.set_option("fused", fused)
    ctx.fields_h2d(*[np.ascontiguousarray(z[f"{key}_f{a}"]) for a in range(6)])
    pin, pout = z[key + "_pin"], z[key + "_pout"]
    host = np.zeros(P2.maxptl, tg.PARTICLE_DTYPE)
    n = pin.size
    for k in ("x", "y", "z", "u", "v", "w", "ch"):
        host[k][:n] = pin[k]
    host["ind"][:n] = np.arange(1, n + 1)
    host["splitlev"][:n] = 1
    ctx.particles_h2d(host, n, 0)
    ctx.move_particles()
    got, ions, lecs = ctx.particles_d2h()
    assert (ions, lecs) == (n, 0)
    got = T.sort_particles(got[:n].copy())
    ref = np.zeros(n, tg.PARTICLE_DTYPE)
    for k in ("x", "y", "z", "u", "v", "w", "ch"):
        ref[k] = pout[k]
    ref["ind"] = np.arange(1, n + 1); ref["splitlev"] = 1
    T.assert_particles_close(got, T.sort_particles(ref), what=key, extent=float(max(P2.mx, P2.my, P2.mz)))
    ctx.close()


@pytest.mark.parametrize("case", range(6))
def test_42_solver_against_the_reference_source(tg, case):
    """advance_{b_halfstep,e_fullstep}_42 (fields.F90:1039-1361) with the radiationx branches and the `wall` clamp: BIT-EXACT"""
    z = load("ref_fields42.npz")
    key = f"g{case}"
    ctx, P = ctx_from_meta(tg, z[key + "_meta"], ntimes=0, highorder=1, wall_i2=int(z[key + "_wall_i2"][0]))
    ctx.fields_h2d(*[np.ascontiguousarray(z[f"{key}_in{a}"]) for a in range(6)])
    ctx.advance_b_halfstep(); ctx.advance_e_fullstep(); ctx.advance_b_halfstep()
    got = ctx.fields_d2h()
    for a in range(6):
        assert np.array_equal(got[a], z[f"{key}_out{a}"]), a
    ctx.close()


@pytest.mark.parametrize("case", range(4))
def test_shock_hooks_against_the_reference_source(tg, case):
    """field_bc_user / particle_bc_user (user/user_shock.F90:342-457).  Fields: bit-exact but for the fp32 cos/sin of the clamp
    values (1 ulp: numpy's against the device's).  Reflected particles: the mover bars.  Wall currents (atomics): 1e-5 of the
    largest current."""
    z = load("ref_shock.npz")
    key = f"s{case}"
    dim, order, px, py, pz, nx, ny, nz = (int(v) for v in z[key + "_meta"])
    mx0, mxcum, maxhlf, nsp = (int(v) for v in z[key + "_geom"])
    par = z[key + "_par"]
    sx = mx0 // nx
    P = tg.make_params(dim=dim, order=order, mx0=mx0, my0=ny, mz0=nz, sizex=sx, rank=mxcum // nx, periodic=(px, py, pz), maxptl=4096,
                       device=0, ntimes=0)
    P.qi, P.qe = float(par[5]), float(par[6])
    ctx = tg.Context(P)
    assert ctx.P.mxcum == mxcum
    ctx.fields_h2d(*[np.ascontiguousarray(z[f"{key}_in{a}"]) for a in range(6)])
    ctx.currents_h2d(*[np.ascontiguousarray(z[f"{key}_in{a}"]) for a in range(6, 9)])
    pin, pout = z[key + "_pin"], z[key + "_pout"]
    host = np.zeros(P.maxptl, tg.PARTICLE_DTYPE)
    ref = np.zeros(2 * nsp, tg.PARTICLE_DTYPE)
    for s0, d0, sign in ((0, 0, 1), (maxhlf, ctx.maxhlf, -1)):
        for k in ("x", "y", "z", "u", "v", "w", "ch"):
            host[k][d0:d0 + nsp] = pin[k][s0:s0 + nsp]
            ref[k][(s0 > 0) * nsp:(s0 > 0) * nsp + nsp] = pout[k][s0:s0 + nsp]
        host["ind"][d0:d0 + nsp] = sign * np.arange(1, nsp + 1)
        host["splitlev"][d0:d0 + nsp] = 1
    ctx.particles_h2d(host, nsp, nsp)
    ctx.field_bc_user_shock(*[float(v) for v in par[:5]])
    ctx.particle_bc_user_wall(float(par[0]))
    got = ctx.fields_d2h()
    for a in range(6):
        refa = z[f"{key}_out{a}"]
        assert float(np.abs(got[a] - refa).max()) <= np.spacing(np.abs(refa).max()), a
    cur = ctx.currents_d2h()
    scale = max(float(np.abs(z[f"{key}_out{a}"]).max()) for a in range(6, 9))
    for a in range(3):
        assert T.max_abs_diff(cur[a], z[f"{key}_out{6 + a}"]) <= 1e-5 * scale, a
    gp, ions, lecs = ctx.particles_d2h()
    assert (ions, lecs) == (nsp, nsp)
    for sp, d0 in ((0, 0), (1, ctx.maxhlf)):
        g1, r1 = gp[d0:d0 + nsp], ref[sp * nsp:(sp + 1) * nsp]
        for k in ("x", "y", "z"):
            assert np.abs(g1[k] - r1[k]).max() <= 2e-6 * max(P.mx, P.my, P.mz), (sp, k)
        for k in ("u", "v", "w"):
            assert np.abs(g1[k] - r1[k]).max() <= 2e-5 * max(np.abs(r1[k]).max(), 1.0), (sp, k)
    ctx.close()


@pytest.mark.parametrize("case", range(7))
@pytest.mark.parametrize("fused", [0, 1])
def test_deposit_particles_against_the_reference_source(tg, case, fused):
    """deposit_particles as a whole (particles_movedeposit.F90:1281-2051) on single ranks and on ranks of split boxes: the
    currents within 1e-5 of the largest current (summation order, atomics), the survivors as a SET keyed by (proc, ind) with
    their wrapped positions within 1e-6 of the box and momenta untouched.  `fused` deposits from the cell-run path where the
    configuration has one (the old position is then recomputed the reference's way, CR_OLDPOS_REF)."""
    z = load("ref_depositp.npz")
    key = f"p{case}"
    dim, order, px, py, pz, nx, ny, nz = (int(v) for v in z[key + "_meta"])
    sx, sy, sz, rank, maxhlf, nsp = (int(v) for v in z[key + "_geom"])
    P = tg.make_params(dim=dim, order=order, mx0=nx, my0=ny, mz0=nz, sizex=sx, sizey=sy, sizez=sz, rank=rank, periodic=(px, py, pz),
                       maxptl=4096, device=0, ntimes=0)
    P.qi, P.qe = float(np.float32(0.07)), float(np.float32(-0.07))
    ctx = tg.Context(P)
    ctx.set_option("fused", fused)
    pin, pout = z[key + "_pin"], z[key + "_pout"]
    host = np.zeros(P.maxptl, tg.PARTICLE_DTYPE)
    host[:nsp] = pin[:nsp]
    host[ctx.maxhlf:ctx.maxhlf + nsp] = pin[maxhlf:maxhlf + nsp]
    ctx.particles_h2d(host, nsp, nsp)
    zero = np.zeros_like(z[f"{key}_cur0"])
    ctx.currents_h2d(zero, zero, zero)
    ctx.deposit_particles()
    ions, lecs = (int(v) for v in z[key + "_counts"][:2])
    ref_i, ref_e = pout[:ions].copy(), pout[maxhlf:maxhlf + lecs].copy()
    if dim == 3 and sz == 1:
        # With one rank along z the reference still routes z-crossers through its out-buffers and gets them back from itself
        # in exchange_particles / inject_others; the library wraps them in place.  Same particles, same shifted z: the
        # reference's survivors plus its two z buffers are the library's survivors.
        lens = z[key + "_boxlen"].reshape(6, 2)
        for d in (4, 5):
            box, ni, nl = z[f"{key}_box{d}"], int(lens[d, 0]), int(lens[d, 1])
            ref_i, ref_e = np.concatenate([ref_i, box[:ni]]), np.concatenate([ref_e, box[ni:ni + nl]])
        ions, lecs = ref_i.size, ref_e.size
    cur = ctx.currents_d2h()
    scale = max(float(np.abs(z[f"{key}_cur{a}"]).max()) for a in range(3))
    for a in range(3):
        assert T.max_abs_diff(cur[a], z[f"{key}_cur{a}"]) <= 1e-5 * scale, (a, T.max_abs_diff(cur[a], z[f"{key}_cur{a}"]) / scale)
    gp, gi, gl = ctx.particles_d2h()
    assert (gi, gl) == (ions, lecs)
    ext = float(max(P.mx, P.my, P.mz))
    T.assert_particles_close(T.sort_particles(gp[:ions].copy()), T.sort_particles(ref_i), rtol_pos=1e-6, rtol_mom=0.0,
                             what=key + " ions", extent=ext)
    T.assert_particles_close(T.sort_particles(gp[ctx.maxhlf:ctx.maxhlf + lecs].copy()), T.sort_particles(ref_e),
                             rtol_pos=1e-6, rtol_mom=0.0, what=key + " lecs", extent=ext)
    ctx.close()


@pytest.mark.parametrize("case", [2, 5, 6, 7, 8, 9])
def test_whole_laps_against_the_reference_mainloop(tg, case):
    """tgpu_step against the reference's `mainloop` run from its own text (tests/golden/ref_lap.npz, the one-rank boxes):
    all-open 3D zigzag box (radiation + edge fixes + exits), periodic 3D order 2, periodic 2D order 1, the 2D shock problem
    with its hooks, and the _42 solver at order 3.  Bars as for the oracle-vs-GPU lap tests: fields 3e-4 per lap of the
    largest interior value, positions 2e-5 per lap, momenta 2e-4 per lap, particle counts equal."""
    z = load("ref_lap.npz")
    key = f"l{case}"
    dim, order, px, py, pz, nx, ny, nz = (int(v) for v in z[key + "_meta"])
    sx, sy, sz, maxhlf, nsp, laps, highorder, shock, fkind = (int(v) for v in z[key + "_geom"])
    assert (sx, sy, sz) == (1, 1, 1)
    par = z[key + "_par"]
    P = tg.make_params(dim=dim, order=order, mx0=nx, my0=ny, mz0=nz, periodic=(px, py, pz), maxptl=2 * maxhlf, device=0, ntimes=2,
                       filter_kind=fkind, highorder=highorder)
    P.qi, P.qe, P.qmi, P.qme = (float(v) for v in par[5:9])
    ctx = tg.Context(P)
    assert ctx.maxhlf == maxhlf
    ctx.fields_h2d(*[np.ascontiguousarray(z[f"{key}_r0_in{a}"]) for a in range(6)])
    zero = np.zeros_like(z[f"{key}_r0_in0"])
    ctx.currents_h2d(zero, zero, zero)
    ctx.particles_h2d(np.ascontiguousarray(z[f"{key}_r0_pin"]), nsp, nsp)
    if shock:
        ctx.set_user_hooks(1, [float(v) for v in par[:5]])
    ctx.step(laps)
    ions, lecs = (int(v) for v in z[f"{key}_r0_counts"])
    assert ctx.counts() == (ions, lecs)
    got = ctx.fields_d2h()
    g, gz = P.nghost // 2, (P.nghostz // 2 if dim == 3 else 0)
    for a in range(6):
        ref = z[f"{key}_r0_out{a}"]
        sl = (slice(gz, ref.shape[0] - gz - 1) if dim == 3 else slice(None), slice(g, ref.shape[1] - g - 1), slice(g, ref.shape[2] - g - 1))
        err = T.max_rel(got[a][sl], ref[sl])
        assert err < 3e-4 * laps, (a, err)
    gp, gi, gl = ctx.particles_d2h()
    pout = z[f"{key}_r0_pout"]
    ext = float(max(P.mx, P.my, P.mz))
    T.assert_particles_close(T.sort_particles(gp[:ions].copy()), T.sort_particles(pout[:ions].copy()), rtol_pos=2e-5 * laps,
                             rtol_mom=2e-4 * laps, what=key + " ions", extent=ext)
    T.assert_particles_close(T.sort_particles(gp[maxhlf:maxhlf + lecs].copy()), T.sort_particles(pout[maxhlf:maxhlf + lecs].copy()),
                             rtol_pos=2e-5 * laps, rtol_mom=2e-4 * laps, what=key + " lecs", extent=ext)
    ctx.close()
