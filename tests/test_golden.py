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
