"""Golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py from the oracle).
CPU: the oracle still reproduces them (regression pin).  GPU: the CUDA library reproduces them without the oracle."""
import glob
import os

import numpy as np
import pytest

from oracle import oracle as O
import pic_testlib as T

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = sorted(glob.glob(os.path.join(HERE, "golden", "case_*.npz")))


def case_of(path):
    b = os.path.basename(path)
    dim, order = int(b[6]), int(b[9])
    n = (8, 8, 8) if dim == 3 else (12, 10, 1)
    kind = 2 if (dim == 3 and order >= 2) else 1
    return dict(dim=dim, order=order, n=n, kind=kind)


def test_fixtures_exist():
    assert len(FILES) == 8


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f) for f in FILES])
def test_oracle_reproduces_golden(path):
    c = case_of(path)
    g = np.load(path)
    w = T.oracle_world(dim=c["dim"], order=c["order"], n=c["n"], ppc=2.0, ntimes=3, filter_kind=c["kind"], delgam=0.05)
    r = w.ranks[0]
    for a in range(6):
        assert np.array_equal(r.arr(a), g["in_" + O.ARR_NAMES[a]])
    assert np.array_equal(T.sort_particles(r.ions().copy()), g["in_ions"])
    r.call("move_particles")
    assert np.array_equal(T.sort_particles(r.lecs().copy()), g["moved_lecs"])
    r.call("reset_currents"); r.call("deposit_currents_only")
    for a in range(6, 9):
        assert np.array_equal(r.arr(a), g["dep_" + O.ARR_NAMES[a]])
    w2 = T.oracle_world(dim=c["dim"], order=c["order"], n=c["n"], ppc=2.0, ntimes=3, filter_kind=c["kind"], delgam=0.05)
    w2.step(); w2.step()
    for a in range(6):
        assert np.array_equal(w2.ranks[0].arr(a), g["lap2_" + O.ARR_NAMES[a]])


@pytest.mark.gpu
@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f) for f in FILES])
def test_gpu_reproduces_golden(tg, path):
    c = case_of(path)
    g = np.load(path)
    P = tg.make_params(dim=c["dim"], order=c["order"], mx0=c["n"][0], my0=c["n"][1], mz0=c["n"][2], ntimes=3,
                       filter_kind=c["kind"], ppc0=2.0, maxptl=65536, device=0)
    # charge normalisation used when the fixtures were made
    Po = O.make_params(dim=c["dim"], order=c["order"], mx0=c["n"][0], my0=c["n"][1], mz0=c["n"][2], ppc0=2.0)
    P.qi, P.qe, P.qmi, P.qme = Po.qi, Po.qe, Po.qmi, Po.qme
    ctx = tg.Context(P)
    p = np.zeros(P.maxptl, tg.PARTICLE_DTYPE)
    ni, ne = g["in_ions"].size, g["in_lecs"].size
    p[:ni] = g["in_ions"]; p[ctx.maxhlf:ctx.maxhlf + ne] = g["in_lecs"]

    def load():
        ctx.fields_h2d(*[np.ascontiguousarray(g["in_" + O.ARR_NAMES[a]]) for a in range(6)])
        ctx.currents_h2d(*[np.zeros(ctx.shape, np.float32) for _ in range(3)])
        ctx.particles_h2d(p, ni, ne)
    load()
    ctx.move_particles()
    gi, ge = T.gpu_particles(ctx)
    T.assert_particles_close(gi, g["moved_ions"]); T.assert_particles_close(ge, g["moved_lecs"])
    # deposit from the golden post-move state
    q = np.zeros(P.maxptl, tg.PARTICLE_DTYPE)
    q[:ni] = g["moved_ions"]; q[ctx.maxhlf:ctx.maxhlf + ne] = g["moved_lecs"]
    ctx2 = tg.Context(P)
    ctx2.fields_h2d(*[np.ascontiguousarray(g["in_" + O.ARR_NAMES[a]]) for a in range(6)])
    ctx2.particles_h2d(q, ni, ne)
    ctx2.reset_currents(); ctx2.deposit_particles()
    cg = ctx2.currents_d2h()
    for k in range(3):
        assert T.max_rel(cg[k], g["dep_" + O.ARR_NAMES[6 + k]]) < 1e-5
    ctx2.close()
    ctx3 = tg.Context(P)
    ctx3.fields_h2d(*[np.ascontiguousarray(g["in_" + O.ARR_NAMES[a]]) for a in range(6)])
    ctx3.particles_h2d(p, ni, ne)
    ctx3.step(2)
    fg = ctx3.fields_d2h()
    w = T.oracle_world(dim=c["dim"], order=c["order"], n=c["n"], ppc=2.0, init="none")     # geometry only
    r = w.ranks[0]
    for a in range(6):
        assert T.max_rel(T.interior(r, fg[a]), T.interior(r, g["lap2_" + O.ARR_NAMES[a]])) < 5e-4
    gi, ge = T.gpu_particles(ctx3)
    T.assert_particles_close(gi, g["lap2_ions"], rtol_pos=4e-5, rtol_mom=4e-4)
    ctx3.close(); ctx.close()


# ------------------------------------------------------------------ feature fixtures: surface, _42 solver, moments
FEAT = {"d3": dict(dim=3, n=(12, 10, 8)), "d2": dict(dim=2, n=(14, 12, 1))}
MOMENTS = ["tdens", "idens", "ibetx", "ebetz", "tmomy", "eener", "iety2"]


def _feat_world(c, **kw):
    return T.oracle_world(dim=c["dim"], order=2, n=c["n"], ppc=2.0, delgam=0.05, seed_fields=4, **kw)


@pytest.mark.parametrize("name", sorted(FEAT))
def test_oracle_reproduces_feature_golden(name):
    c = FEAT[name]
    g = np.load(os.path.join(HERE, "golden", f"feat_{name}.npz"))
    w = _feat_world(c, periodic=(0, 1, 1)); r = w.ranks[0]
    for a in range(6):
        assert np.array_equal(r.arr(a), g["surf_in_" + O.ARR_NAMES[a]])
    for _ in range(2):
        for ph in (O.PH_SURF_B, O.PH_BC_B1, O.PH_SURF_E, O.PH_BC_E1):
            w.phase(ph)
    for a in range(6):
        assert np.array_equal(r.arr(a), g["surf_out_" + O.ARR_NAMES[a]])
    w = _feat_world(c, highorder=1); r = w.ranks[0]
    for nm in ("advance_b_halfstep", "advance_e_fullstep", "advance_b_halfstep"):
        r.call(nm)
    for a in range(6):
        assert np.array_equal(r.arr(a), g["s42_out_" + O.ARR_NAMES[a]])
    w = _feat_world(c); r = w.ranks[0]
    for m in MOMENTS:
        w.meanq_fld_cur(m)
        assert np.array_equal(r.arr(O.CURX), g["mom_" + m])


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(FEAT))
def test_gpu_reproduces_feature_golden(tg, name):
    """radiation `surface` and the 4th-order solver bit-exact, the moments to 2e-5 — from files, without the oracle's arithmetic"""
    c = FEAT[name]
    g = np.load(os.path.join(HERE, "golden", f"feat_{name}.npz"))

    def ctx_for(**kw):
        P = tg.make_params(dim=c["dim"], order=2, mx0=c["n"][0], my0=c["n"][1], mz0=c["n"][2], ntimes=0, ppc0=2.0, maxptl=65536,
                           device=0, **kw)
        Po = O.make_params(dim=c["dim"], order=2, mx0=c["n"][0], my0=c["n"][1], mz0=c["n"][2], ppc0=2.0)
        P.qi, P.qe, P.qmi, P.qme = Po.qi, Po.qe, Po.qmi, Po.qme
        return tg.Context(P)

    ctx = ctx_for(periodic=(0, 1, 1))
    ctx.fields_h2d(*[np.ascontiguousarray(g["surf_in_" + O.ARR_NAMES[a]]) for a in range(6)])
    for _ in range(2):
        ctx.bc_b2(); ctx.bc_e2()
    for a, f in enumerate(ctx.fields_d2h()):
        assert np.array_equal(f, g["surf_out_" + O.ARR_NAMES[a]]), O.ARR_NAMES[a]
    ctx.close()
    ctx = ctx_for(highorder=1)
    ctx.fields_h2d(*[np.ascontiguousarray(g["s42_in_" + O.ARR_NAMES[a]]) for a in range(6)])
    ctx.advance_b_halfstep(); ctx.advance_e_fullstep(); ctx.advance_b_halfstep()
    for a, f in enumerate(ctx.fields_d2h()):
        assert np.array_equal(f, g["s42_out_" + O.ARR_NAMES[a]]), O.ARR_NAMES[a]
    ctx.close()
    ctx = ctx_for()
    p = np.zeros(ctx.maxptl, tg.PARTICLE_DTYPE)
    ni, ne = g["mom_ions"].size, g["mom_lecs"].size
    p[:ni] = g["mom_ions"]; p[ctx.maxhlf:ctx.maxhlf + ne] = g["mom_lecs"]
    ctx.particles_h2d(p, ni, ne)
    w = T.oracle_world(dim=c["dim"], order=2, n=c["n"], ppc=2.0, init="none")     # geometry only
    r = w.ranks[0]
    for m in MOMENTS:
        ctx.meanq_fld_cur(m)
        assert T.max_rel(T.interior(r, ctx.currents_d2h()[0]), T.interior(r, g["mom_" + m])) < 2e-5, m
    ctx.close()
