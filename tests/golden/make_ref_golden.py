#!/usr/bin/env python
"""Golden vectors produced by the REFERENCE'S OWN SOURCE TEXT (run in THIS container, where /root/reference exists).

The reference cannot be compiled anywhere we can reach (no Fortran compiler in the image or on the GPU box), so its hot-path
subroutines are executed by the Fortran-subset interpreter tests/golden/f90run.py: the text of each subroutine is read from
/root/reference/code/*.F90, cpp-preprocessed with the build's defines, translated statement by statement and run in fp32.
The inputs (seeded, generated here) and the reference's outputs are stored in tests/golden/ref_*.npz; tests/test_ref_golden.py
checks the CPU oracle against them (bit for bit) and tests/test_gpu_ref_golden.py the CUDA library.  Re-run:

    python tests/golden/make_ref_golden.py            # needs /root/reference; ~10 minutes (a Python interpreter of Fortran)
    python tests/golden/make_ref_golden.py lap        # one group: deposit fields mover filter radiation halo fields42 shock
                                                      #            depositp halo_mr migrate_mr lap

G1-G9 run single routines on one rank (or on one rank of a split box); G10-G12 run SEVERAL RANKS as threads with MPI_SendRecv
as a rendezvous, up to whole laps of `mainloop` itself.  Seeds are fixed: re-running reproduces the committed files.
"""
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import f90run as R  # noqa: E402

F = np.float32
REF = os.environ.get("TRISTAN_REFERENCE", "/root/reference")
OUT = os.environ.get("TRISTAN_GOLDEN_OUT", HERE)          # where the .npz files go (the tests re-generate into a tmp dir)


def src(name):
    return open(os.path.join(REF, "code", name)).read()


def constants():
    """particles.F90:238-250, evaluated as Fortran does (3/2. etc. are fp32 divisions)"""
    return dict(three=F(3.), two=F(2.), thhalf=F(F(3) / F(2.)), nineighth=F(F(9) / F(8.)), one=F(1.), threeq=F(F(3) / F(4.)),
                twoth=F(F(2) / F(3.)), half=F(F(1) / F(2.)), third=F(F(1) / F(3.)), quart=F(F(1) / F(4.)), sixth=F(F(1) / F(6.)),
                negsixth=F(F(-1) / F(6.)), negone=F(-1.))


GINTS = {"ix", "iy", "iz", "mx", "my", "mz", "lot", "nghost", "nghostz", "periodicx", "periodicy", "periodicz", "size0", "sizex", "sizey",
         "sizez", "rank", "ions", "lecs", "maxhlf", "ntimes", "external_fields", "debug", "highorder"}
GARR = {"ex", "ey", "ez", "bx", "by", "bz", "curx", "cury", "curz", "temp", "p"}


def grid(dim, order, n):
    ng = 5 if order <= 1 else 7
    ngz = ng if dim == 3 else 5
    mx, my = n[0] + ng, n[1] + ng
    mz = n[2] + ngz if dim == 3 else 1
    return ng, ngz, mx, my, mz


# ------------------------------------------------------------------------------------------------------------
# G1: the deposit kernels -- zigzag (particles.F90:550-669), densdecomp_{1,2,3}ord (:678-1358), 2D and 3D branches
# ------------------------------------------------------------------------------------------------------------
def gen_deposit():
    out = {}
    names = {0: "zigzag", 1: "densdecomp_1ord", 2: "densdecomp_2ord", 3: "densdecomp_3ord"}
    text = src("particles.F90")
    for dim in (2, 3):
        for order in (0, 1, 2, 3):
            defines = {"MPI"} | ({"twoD"} if dim == 2 else set())
            sub = R.Sub(text, names[order], defines=defines, global_arrays=GARR, global_ints=GINTS).compile()
            n = (9, 8, 7)
            ng, ngz, mx, my, mz = grid(dim, order, n)
            g = R.Globals(mx=mx, my=my, mz=mz, ix=1, iy=mx, iz=mx * my if dim == 3 else 0, lot=mx * my * mz,
                          curx=R.FArr((mx, my, mz)), cury=R.FArr((mx, my, mz)), curz=R.FArr((mx, my, mz)), q=F(0), **constants())
            rng = np.random.default_rng(100 * dim + order)
            npart = 96
            lo = ng // 2 + 1
            x1 = (lo + rng.random(npart) * n[0]).astype(F)
            y1 = (lo + rng.random(npart) * n[1]).astype(F)
            z1 = ((ngz // 2 + 1) + rng.random(npart) * n[2]).astype(F) if dim == 3 else (3 + rng.random(npart)).astype(F)
            d = ((rng.random((3, npart)) - 0.5) * 0.88).astype(F)
            # a few particles exactly on cell boundaries / half cells, where the shape branches switch
            x1[:6] = np.array([lo + 2, lo + 2.5, lo + 3, lo + 3.5, lo + 1, lo + 4.5], F)
            d[0, :3] = np.array([0.25, -0.25, -0.4], F)
            x2, y2, z2 = (x1 + d[0]).astype(F), (y1 + d[1]).astype(F), (z1 + d[2]).astype(F)
            q = ((rng.random(npart) - 0.5) * 2).astype(F)
            for i in range(npart):
                g.q = F(q[i])
                if order == 0:
                    sub(g, x2[i], y2[i], z2[i], x1[i], y1[i], z1[i], False)
                else:
                    sub(g, x2[i], y2[i], z2[i], x1[i], y1[i], z1[i], False)
            key = f"d{dim}o{order}"
            for nm, a in (("x1", x1), ("y1", y1), ("z1", z1), ("x2", x2), ("y2", y2), ("z2", z2), ("q", q)):
                out[f"{key}_{nm}"] = a
            for nm in ("curx", "cury", "curz"):
                out[f"{key}_{nm}"] = getattr(g, nm).nd().transpose(2, 1, 0).copy()        # C order (mz, my, mx)
            out[f"{key}_n"] = np.array(n, np.int32)
            print("deposit", key, "sum|curx| =", float(np.abs(out[f"{key}_curx"]).sum()))
    np.savez_compressed(os.path.join(OUT, "ref_deposit.npz"), **out)


# ------------------------------------------------------------------------------------------------------------
# G2: the Yee solver -- advance_b_halfstep (fields.F90:586-728), advance_e_fullstep (:739-870), index ranges included
# ------------------------------------------------------------------------------------------------------------
def field_globals(dim, order, n, periodic, rng, c=0.45, corr=1.025):
    ng, ngz, mx, my, mz = grid(dim, order, n)
    g = R.Globals(mx=mx, my=my, mz=mz, ix=1, iy=mx, iz=mx * my if dim == 3 else 0, lot=mx * my * mz, nghost=ng, nghostz=ngz,
                  periodicx=periodic[0], periodicy=periodic[1], periodicz=periodic[2], size0=1, sizex=1, sizey=1, sizez=1, rank=0,
                  c=F(c), corr=F(corr), **constants())
    for nm in ("ex", "ey", "ez", "bx", "by", "bz", "curx", "cury", "curz"):
        a = R.FArr((mx, my, mz))
        a.flat[:] = (rng.standard_normal(a.flat.size) * 0.1).astype(F)
        setattr(g, nm, a)
    return g


def c_order(a):
    return a.nd().transpose(2, 1, 0).copy()


def gen_fields():
    out = {}
    text = src("fields.F90")
    cases = [(3, 2, (1, 1, 1)), (3, 1, (0, 1, 1)), (3, 2, (0, 0, 0)), (3, 2, (1, 0, 1)), (2, 1, (1, 1, 1)), (2, 2, (0, 1, 1)), (2, 2, (0, 0, 1))]
    for ci, (dim, order, per) in enumerate(cases):
        defines = {"MPI"} | ({"twoD"} if dim == 2 else set())
        bhalf = R.Sub(text, "advance_b_halfstep", defines=defines, global_arrays=GARR, global_ints=GINTS).compile()
        efull = R.Sub(text, "advance_e_fullstep", defines=defines, global_arrays=GARR, global_ints=GINTS).compile()
        addc = R.Sub(text, "add_current", defines=defines, global_arrays=GARR, global_ints=GINTS).compile()
        n = (8, 7, 6)
        g = field_globals(dim, order, n, per, np.random.default_rng(200 + ci))
        key = f"f{ci}"
        out[key + "_meta"] = np.array([dim, order, *per, *n], np.int32)
        for a, nm in enumerate(("ex", "ey", "ez", "bx", "by", "bz", "curx", "cury", "curz")):
            out[f"{key}_in{a}"] = c_order(getattr(g, nm))
        bhalf(g); efull(g); bhalf(g); addc(g)
        for a, nm in enumerate(("ex", "ey", "ez", "bx", "by", "bz")):
            out[f"{key}_out{a}"] = c_order(getattr(g, nm))
        print("fields", key, dim, order, per)
    np.savez_compressed(os.path.join(OUT, "ref_fields.npz"), **out)


# ------------------------------------------------------------------------------------------------------------
# G3: the movers -- mover (zigzag build), mover_1ord, mover_2ord, mover_3ord (particles_movedeposit.F90:98-1271): the
#     cshift pre-average, the shape weights incl. the loop-range quirk Q1 of mover_2ord, the gather, the Boris push
# ------------------------------------------------------------------------------------------------------------
PDT = np.dtype([("x", "f4"), ("y", "f4"), ("z", "f4"), ("u", "f4"), ("v", "f4"), ("w", "f4"), ("ch", "f4"), ("ind", "i4"), ("proc", "i4"),
                ("splitlev", "i4")])


def gen_mover():
    out = {}
    text = src("particles_movedeposit.F90")
    names = {0: "mover", 1: "mover_1ord", 2: "mover_2ord", 3: "mover_3ord"}
    EXT = tuple(F(v) for v in (0.01, -0.02, 0.015, 0.05, -0.03, 0.04))       # ex, ey, ez, bx, by, bz of get_external_fields
    for dim in (2, 3):
        for order in (0, 1, 2, 3):
            # variants: "" = Boris; "v" = the `vay` build (Vay 2008 pusher); "x" = external_fields through get_external_fields
            for variant in ("", "v", "x"):
                defines = {"MPI"} | ({"twoD"} if dim == 2 else set()) | ({"vay"} if variant == "v" else set())
                sub = R.Sub(text, names[order], defines=defines, global_arrays=GARR, global_ints=GINTS).compile()
                n = (9, 8, 7)
                rng = np.random.default_rng(300 + 10 * dim + order)
                g = field_globals(dim, order, n, (1, 1, 1), rng)
                ng, ngz, mx, my, mz = grid(dim, order, n)
                npart = 64
                p = np.zeros(npart, PDT)
                lo = ng // 2 + 1
                p["x"] = (lo + rng.random(npart) * n[0]).astype(F)
                p["y"] = (lo + rng.random(npart) * n[1]).astype(F)
                p["z"] = ((ngz // 2 + 1) + rng.random(npart) * n[2]).astype(F) if dim == 3 else (3 + rng.random(npart)).astype(F)
                # particles exactly on nodes and on half cells: the branch points of the shapes and of quirk Q1
                p["x"][:4] = np.array([lo + 2, lo + 2.5, lo + 3, lo + 4.5], F)
                p["y"][2:6] = np.array([lo + 1, lo + 1.5, lo + 2, lo + 3.5], F)
                for k in "uvw":
                    p[k] = (rng.standard_normal(npart) * 0.4).astype(F)
                p["ch"] = 1.0
                g.p = R.RecArr(p)
                g.external_fields = variant == "x"
                # the problem's get_external_fields(x, y, z, ex, ey, ez, bx, by, bz [, qm, n]): a uniform field here
                g.get_external_fields = lambda *a: tuple(a[:3]) + EXT + tuple(a[9:])
                g.delgam = F(1e-2)
                g.qme_abs = F(1.0)
                key = f"m{dim}o{order}{variant}"
                out[key + "_meta"] = np.array([dim, order, 1, 1, 1, *n], np.int32)
                for a, nm in enumerate(("ex", "ey", "ez", "bx", "by", "bz")):
                    out[f"{key}_f{a}"] = c_order(getattr(g, nm))
                out[key + "_pin"] = p.copy()
                qm = F(-0.8)
                out[key + "_qm"] = np.array([qm], F)
                out[key + "_ext"] = np.array(EXT, F)
                sub(g, 1, npart, qm)
                out[key + "_pout"] = p.copy()
                print("mover", key, "mean |dx| =", float(np.abs(out[key + "_pout"]["x"] - out[key + "_pin"]["x"]).mean()))
    np.savez_compressed(os.path.join(OUT, "ref_mover.npz"), **out)


# ------------------------------------------------------------------------------------------------------------
# G4: the digital filters -- apply_filter1_opt (filter.F90:8-221) and filter_x / filter_y / filter_z
#     (optimized_filters.F90:459-915).  The layer copies are the reference's own copy_layr{x,y,z}1_opt (fieldboundaries.F90:
#     1149-1677); their MPI_SendRecv to the rank itself is executed as "recvbuf = sendbuf" (f90run.py).
# ------------------------------------------------------------------------------------------------------------
def layer_copies(g, defines):
    """bind the reference's own layer-copy routines (local and MPI-to-self variants) into the globals"""
    fb = src("fieldboundaries.F90")
    g.statsize, g.mpi_comm_world, g.mpi_read = 5, 0, 0
    for nm in ("copylayrx", "copylayry", "copy_layrx1_opt", "copy_layry1_opt", "copy_layrz1_opt", "copy_layrx2_opt", "copy_layry2_opt",
               "copy_layrz2_opt"):
        f = R.Sub(fb, nm, defines=defines, global_arrays=GARR, global_ints=GINTS | {"statsize"}).compile()
        setattr(g, nm, (lambda f_: (lambda *a: f_(g, *a)))(f))


def gen_filter():
    out = {}
    for ci, (dim, order, ntimes) in enumerate([(2, 1, 3), (3, 2, 2), (3, 1, 3), (2, 2, 4)]):
        defines = {"MPI"} | ({"twoD"} if dim == 2 else set())
        f1 = R.Sub(src("filter.F90"), "apply_filter1_opt", defines=defines, global_arrays=GARR, global_ints=GINTS).compile()
        n = (8, 7, 6)
        g = field_globals(dim, order, n, (1, 1, 1), np.random.default_rng(400 + ci))
        g.ntimes = ntimes
        g.temp = R.FArr((g.mx, g.my, g.mz))
        layer_copies(g, defines)
        key = f"f1_{ci}"
        out[key + "_meta"] = np.array([dim, order, 1, 1, 1, *n, ntimes], np.int32)
        for c, nm in enumerate(("curx", "cury", "curz")):
            out[f"{key}_in{c}"] = c_order(getattr(g, nm))
        f1(g)
        for c, nm in enumerate(("curx", "cury", "curz")):
            out[f"{key}_out{c}"] = c_order(getattr(g, nm))
        print("filter1", key, dim, order, ntimes)
    # filter_x / filter_y / filter_z on one component with given ghost slabs
    text = src("optimized_filters.F90")
    for ci, (dim, order, ntimes) in enumerate([(3, 2, 3), (3, 1, 4), (2, 2, 2), (3, 2, 5)]):
        defines = {"MPI", "filter2"} | ({"twoD"} if dim == 2 else set())
        n = (9, 8, 7)
        rng = np.random.default_rng(500 + ci)
        g = field_globals(dim, order, n, (1, 1, 1), rng)
        g.ntimes = ntimes
        key = f"f2_{ci}"
        out[key + "_meta"] = np.array([dim, order, 1, 1, 1, *n, ntimes], np.int32)
        gh = g.nghost // 2
        ghz = g.nghostz // 2
        rng_lo = (gh + 1, gh + 1, ghz + 1 if dim == 3 else 1)
        rng_hi = (g.mx - gh - 1, g.my - gh - 1, g.mz - ghz - 1 if dim == 3 else 1)
        out[key + "_in"] = c_order(g.curx)
        for axis, nm in enumerate(("filter_x", "filter_y", "filter_z")):
            if dim == 2 and axis == 2:
                continue
            sub = R.Sub(text, nm, defines=defines, global_arrays=GARR, global_ints=GINTS)
            fn = sub.compile()
            shape = [g.mx, g.my, g.mz]
            shape[axis] = 2 * ntimes
            ghost = R.FArr(tuple(shape))
            # periodic single rank: low ghosts = my last ntimes interior cells, high ghosts = my first ntimes
            cur = g.curx.nd()
            sl_lo = [slice(None)] * 3; sl_hi = [slice(None)] * 3
            sl_lo[axis] = slice(rng_hi[axis] - ntimes, rng_hi[axis]); sl_hi[axis] = slice(rng_lo[axis] - 1, rng_lo[axis] - 1 + ntimes)
            gv = ghost.nd()
            d_lo = [slice(None)] * 3; d_hi = [slice(None)] * 3
            d_lo[axis] = slice(0, ntimes); d_hi[axis] = slice(ntimes, 2 * ntimes)
            gv[tuple(d_lo)] = cur[tuple(sl_lo)]; gv[tuple(d_hi)] = cur[tuple(sl_hi)]
            args = sub.args
            vals = dict(cur=g.curx, ghost=ghost, xghost=ghost, yghost=ghost, zghost=ghost, istr=rng_lo[0], ifin=rng_hi[0], jstr=rng_lo[1],
                        jfin=rng_hi[1], kstr=rng_lo[2], kfin=rng_hi[2], ntimes=ntimes, mx=g.mx, my=g.my, mz=g.mz)
            fn(g, *[vals[a] for a in args])
            out[f"{key}_after{axis}"] = c_order(g.curx)
        print("filter2", key, dim, order, ntimes)
    np.savez_compressed(os.path.join(OUT, "ref_filter.npz"), **out)


# ------------------------------------------------------------------------------------------------------------
# G5: radiation boundaries -- bc_b2 / bc_e2 -> surface (fieldboundaries.F90:274-295, 403-426, 493-606) and the edge fixes
#     pre_bc_b / post_bc_b / pre_bc_e / post_bc_e -> preledge / postedge (:114-163, 437-482, 2200-2505).  The callers are
#     executed too (their rotated argument lists are part of the source); the ghost refresh that follows each of them
#     (bc_b1 / bc_e1, MPI) is stubbed out and tested separately.
# ------------------------------------------------------------------------------------------------------------
def gen_radiation():
    out = {}
    fb = src("fieldboundaries.F90")
    cases = [(3, 1, (0, 0, 0)), (3, 2, (0, 1, 1)), (3, 2, (1, 0, 1)), (2, 1, (0, 1, 1)), (2, 2, (0, 0, 1)), (2, 1, (1, 0, 1))]
    for ci, (dim, order, per) in enumerate(cases):
        defines = {"MPI"} | ({"twoD"} if dim == 2 else set())
        n = (8, 7, 6)
        g = field_globals(dim, order, n, per, np.random.default_rng(600 + ci))
        g.radiationx, g.radiationy, g.radiationz = 1 - per[0], 1 - per[1], 1 - per[2]
        if dim == 2 and g.radiationy == 1:
            g.radiationz = 1                                           # fieldboundaries.F90:91-93 (#ifdef twoD)
        gi = GINTS | {"radiationx", "radiationy", "radiationz"}
        sub = {nm: R.Sub(fb, nm, defines=defines, global_arrays=GARR, global_ints=gi).compile()
               for nm in ("surface", "preledge", "postedge", "bc_b2", "bc_e2", "pre_bc_b", "post_bc_b", "pre_bc_e", "post_bc_e")}
        for nm in ("surface", "preledge", "postedge"):
            setattr(g, nm, (lambda f: (lambda *a: f(g, *a)))(sub[nm]))
        g.bc_b1 = lambda: None
        g.bc_e1 = lambda: None
        key = f"r{ci}"
        out[key + "_meta"] = np.array([dim, order, *per, *n], np.int32)
        for a, nm in enumerate(("ex", "ey", "ez", "bx", "by", "bz")):
            out[f"{key}_in{a}"] = c_order(getattr(g, nm))
        seq = ["pre_bc_b", "bc_b2", "post_bc_b", "pre_bc_e", "bc_e2", "post_bc_e"]
        for si, nm in enumerate(seq):
            sub[nm](g)
            for a, fn in enumerate(("ex", "ey", "ez", "bx", "by", "bz")):
                out[f"{key}_s{si}_{a}"] = c_order(getattr(g, fn))
        print("radiation", key, dim, order, per)
    np.savez_compressed(os.path.join(OUT, "ref_radiation.npz"), **out)


# ------------------------------------------------------------------------------------------------------------
# G6: ghost refresh and current fold on one periodic rank -- bc_b1 / bc_e1 (fieldboundaries.F90:181-263, 306-392) with
#     copylayrx / copylayry / copy_layrz1_opt, and exchange_current (:1768-2189)
# ------------------------------------------------------------------------------------------------------------
def gen_halo():
    out = {}
    fb = src("fieldboundaries.F90")
    for ci, (dim, order) in enumerate([(2, 1), (2, 2), (3, 1), (3, 3)]):
        defines = {"MPI"} | ({"twoD"} if dim == 2 else set())
        n = (8, 7, 6)
        g = field_globals(dim, order, n, (1, 1, 1), np.random.default_rng(700 + ci))
        layer_copies(g, defines)
        g.highorder = False
        ng = g.nghost
        g.bufferin1x, g.bufferin2x = R.FArr((ng // 2 + 1, g.my, g.mz)), R.FArr((ng // 2, g.my, g.mz))
        g.bufferin1y, g.bufferin2y = R.FArr((g.mx, ng // 2 + 1, g.mz)), R.FArr((g.mx, ng // 2, g.mz))
        g.bufferin1, g.bufferin2 = R.FArr((g.mx, g.my, g.nghostz // 2 + 1)), R.FArr((g.mx, g.my, g.nghostz // 2))
        ga = GARR | {"bufferin1x", "bufferin2x", "bufferin1y", "bufferin2y", "bufferin1", "bufferin2"}
        gi = GINTS | {"statsize"}
        bcb = R.Sub(fb, "bc_b1", defines=defines, global_arrays=ga, global_ints=gi).compile()
        bce = R.Sub(fb, "bc_e1", defines=defines, global_arrays=ga, global_ints=gi).compile()
        exc = R.Sub(fb, "exchange_current", defines=defines, global_arrays=ga, global_ints=gi).compile()
        key = f"h{ci}"
        out[key + "_meta"] = np.array([dim, order, 1, 1, 1, *n], np.int32)
        names = ("ex", "ey", "ez", "bx", "by", "bz", "curx", "cury", "curz")
        # the last layer (index m) of the currents never receives a deposit in a run; with it zero the shape typo of
        # fieldboundaries.F90:1938 (a stale layer of the previous component is added) is invisible, as it is in the reference
        for nm in ("curx", "cury", "curz"):
            v = getattr(g, nm).nd()
            v[-1, :, :] = 0; v[:, -1, :] = 0
            if dim == 3:
                v[:, :, -1] = 0
        for a, nm in enumerate(names):
            out[f"{key}_in{a}"] = c_order(getattr(g, nm))
        bcb(g); bce(g); exc(g)
        for a, nm in enumerate(names):
            out[f"{key}_out{a}"] = c_order(getattr(g, nm))
        print("halo", key, dim, order)
    np.savez_compressed(os.path.join(OUT, "ref_halo.npz"), **out)


# ------------------------------------------------------------------------------------------------------------
# G7: the 4th-order `_42` solver -- advance_b_halfstep_42 (fields.F90:1039-1212), advance_e_fullstep_42 (:1223-1361); with
#     and without the injector-side clamp of the x range (`wall`, xinject2)
# ------------------------------------------------------------------------------------------------------------
def gen_fields42():
    out = {}
    text = src("fields.F90")
    cases = [(3, 2, (1, 1, 1), 0), (3, 2, (0, 1, 1), 0), (3, 2, (0, 1, 1), 6), (2, 2, (1, 1, 1), 0), (2, 2, (0, 1, 1), 5), (2, 2, (0, 0, 1), 0)]
    for ci, (dim, order, per, xinj) in enumerate(cases):
        defines = {"MPI"} | ({"twoD"} if dim == 2 else set())
        bhalf = R.Sub(text, "advance_b_halfstep_42", defines=defines, global_arrays=GARR, global_ints=GINTS).compile()
        efull = R.Sub(text, "advance_e_fullstep_42", defines=defines, global_arrays=GARR, global_ints=GINTS).compile()
        n = (12, 9, 8)
        g = field_globals(dim, order, n, per, np.random.default_rng(800 + ci))
        g.wall = xinj > 0
        g.xinject2 = F(xinj + 0.25)
        g.radiationx, g.radiationy, g.radiationz = 1 - per[0], 1 - per[1], 1 - per[2]
        key = f"g{ci}"
        out[key + "_meta"] = np.array([dim, order, *per, *n], np.int32)
        out[key + "_wall_i2"] = np.array([xinj + 10 if xinj else 0], np.int32)
        for a, nm in enumerate(("ex", "ey", "ez", "bx", "by", "bz", "curx", "cury", "curz")):
            out[f"{key}_in{a}"] = c_order(getattr(g, nm))
        bhalf(g); efull(g); bhalf(g)
        for a, nm in enumerate(("ex", "ey", "ez", "bx", "by", "bz")):
            out[f"{key}_out{a}"] = c_order(getattr(g, nm))
        print("fields42", key, dim, order, per, xinj)
    np.savez_compressed(os.path.join(OUT, "ref_fields42.npz"), **out)


# ------------------------------------------------------------------------------------------------------------
# G8: the shock problem's hooks -- field_bc_user / particle_bc_user (user/user_shock.F90:342-457) with the reference's own
#     iloc / xglob (fields.F90:384-467) and zigzag (particles.F90:550-669); one rank, and the first / last rank of a 3-rank x split
# ------------------------------------------------------------------------------------------------------------
def gen_shock():
    out = {}
    user = open(os.path.join(REF, "user", "user_shock.F90")).read()
    ftext, ptext = src("fields.F90"), src("particles.F90")
    cases = [(3, 1, 40, 0, 40), (2, 2, 40, 0, 40), (2, 1, 60, 0, 20), (2, 1, 60, 40, 20)]     # dim, order, mx0, mxcum, local nx
    for ci, (dim, order, mx0, mxcum, nx) in enumerate(cases):
        defines = {"MPI"} | ({"twoD"} if dim == 2 else set())
        gi = GINTS | {"mxcum", "mx0"}
        n = (nx, 7, 6)
        rng = np.random.default_rng(900 + ci)
        g = field_globals(dim, order, n, (0, 1, 1), rng)
        ng, ngz, mx, my, mz = grid(dim, order, n)
        g.mx0, g.mxcum = mx0 + ng, mxcum                       # mx0 counts the ghost cells (fields.F90:292)
        g.leftwall, g.binit, g.btheta, g.bphi, g.beta = F(15.0), F(0.05), F(1.1), F(0.4), F(0.3)
        g.wall = True
        g.qi, g.qe = F(0.07), F(-0.07)
        for nm in ("iloc", "xglob"):
            f = R.Sub(ftext, nm, defines=defines, global_arrays=GARR, global_ints=gi).compile()
            setattr(g, nm, (lambda f: (lambda *a: f(g, *a)))(f))
        zz = R.Sub(ptext, "zigzag", defines=defines, global_arrays=GARR, global_ints=gi).compile()
        g.zigzag = lambda *a: zz(g, *a)
        fbc = R.Sub(user, "field_bc_user", defines=defines, global_arrays=GARR, global_ints=gi).compile()
        pbc = R.Sub(user, "particle_bc_user", defines=defines, global_arrays=GARR, global_ints=gi).compile()
        # particles: half of each species has just crossed the wall (global x < leftwall after a push with u < 0)
        nsp = 24
        maxhlf = 32
        p = np.zeros(2 * maxhlf, PDT)
        for s0 in (0, maxhlf):
            sl = slice(s0, s0 + nsp)
            for k in "uvw":
                p[k][sl] = (rng.standard_normal(nsp) * 0.5).astype(F)
            p["u"][s0:s0 + nsp // 2] = -np.abs(p["u"][s0:s0 + nsp // 2]) - F(0.05)
            gam = np.sqrt(1 + p["u"][sl].astype(np.float64) ** 2 + p["v"][sl] ** 2 + p["w"][sl] ** 2)
            back = (rng.random(nsp) * 0.95 * np.abs(p["u"][sl] / gam) * 0.45).astype(F)       # how far behind the wall it ended
            xg = np.where(np.arange(nsp) < nsp // 2, 15.0 - back, 15.0 + 0.01 + rng.random(nsp) * 8).astype(F)
            p["x"][sl] = (xg - mxcum).astype(F)
            p["y"][sl] = ((ng // 2 + 1) + rng.random(nsp) * n[1]).astype(F)
            p["z"][sl] = ((ngz // 2 + 1) + rng.random(nsp) * n[2]).astype(F) if dim == 3 else (3 + rng.random(nsp)).astype(F)
            p["ch"][sl] = 1.0
        if mxcum:                      # a rank that does not hold the wall: nothing may happen to its particles
            p["x"] = (p["x"] + F(mxcum) + F(6)).astype(F) - F(mxcum)
        g.p = R.RecArr(p)
        g.ions, g.lecs, g.maxhlf = nsp, nsp, maxhlf
        g.q = F(0)
        key = f"s{ci}"
        out[key + "_meta"] = np.array([dim, order, 0, 1, 1, *n], np.int32)
        out[key + "_geom"] = np.array([mx0, mxcum, maxhlf, nsp], np.int32)
        out[key + "_par"] = np.array([g.leftwall, g.binit, g.btheta, g.bphi, g.beta, g.qi, g.qe], F)
        for a, nm in enumerate(("ex", "ey", "ez", "bx", "by", "bz", "curx", "cury", "curz")):
            out[f"{key}_in{a}"] = c_order(getattr(g, nm))
        out[key + "_pin"] = p.copy()
        fbc(g)
        pbc(g)
        for a, nm in enumerate(("ex", "ey", "ez", "bx", "by", "bz", "curx", "cury", "curz")):
            out[f"{key}_out{a}"] = c_order(getattr(g, nm))
        out[key + "_pout"] = p.copy()
        moved = int((out[key + "_pin"]["u"] != p["u"]).sum())
        print("shock", key, dim, order, "reflected", moved, "field cells changed", int(sum((out[f"{key}_in{a}"] != out[f"{key}_out{a}"]).sum() for a in range(6))))
    np.savez_compressed(os.path.join(OUT, "ref_shock.npz"), **out)


# ------------------------------------------------------------------------------------------------------------
# G9: deposit_particles as a whole (particles_movedeposit.F90:1281-2051): the unwinding of the old position, the deposit
#     calls, the periodic wrap / domain-exit classification with the neighbour-size shifts, the six out-buffers and the
#     hole-filling compaction -- on single ranks and on one rank of a split box (no MPI inside this routine)
# ------------------------------------------------------------------------------------------------------------
def gen_depositp():
    out = {}
    mtext, ptext = src("particles_movedeposit.F90"), src("particles.F90")
    dirs = ("outup", "outdwn", "inblw", "inabv", "outlft", "outrgt", "inlft", "inrgt", "outminus", "outplus", "inminus", "inplus")
    gi = GINTS | {"mxcum", "mycum", "mzcum", "nionout", "nlecout", "lap"} | {f"len{s}{d}" for s in ("ion", "lec") for d in dirs}
    ga = GARR | {"pind", "mxl", "myl", "mzl", "poutup", "poutdwn", "poutlft", "poutrgt", "poutminus", "poutplus"}
    dep = {0: ("zzag", "zigzag"), 1: ("dd1", "densdecomp_1ord"), 2: ("dd2", "densdecomp_2ord"), 3: ("dd3", "densdecomp_3ord")}
    #        dim order global n      sizes      rank periodic
    cases = [(2, 1, (12, 10, 1), (1, 1, 1), 0, (1, 1, 1)),
             (2, 2, (12, 10, 1), (1, 1, 1), 0, (0, 1, 1)),
             (2, 0, (24, 20, 1), (2, 2, 1), 3, (1, 1, 1)),
             (3, 2, (8, 12, 12), (1, 2, 2), 3, (1, 1, 1)),           # 3D: sizex = 1 always (communications.F90:176-181)
             (3, 1, (10, 8, 8), (1, 1, 1), 0, (1, 1, 1)),
             (3, 3, (10, 16, 8), (1, 2, 1), 1, (1, 0, 1)),
             (3, 0, (8, 8, 16), (1, 1, 2), 0, (0, 1, 0))]
    for ci, (dim, order, ng_, sizes, rank, per) in enumerate(cases):
        flag, name = dep[order]
        defines = {"MPI", flag} | ({"twoD"} if dim == 2 else set())
        n = tuple(a // s for a, s in zip(ng_, sizes))
        rng = np.random.default_rng(1000 + ci)
        g = field_globals(dim, order, n, per, rng)
        ng, ngz, mx, my, mz = grid(dim, order, n)
        sx, sy, sz = sizes
        size0 = sx * sy * sz
        g.size0, g.sizex, g.sizey, g.sizez, g.rank = size0, sx, sy, sz, rank
        g.mxcum, g.mycum, g.mzcum = (rank % sx) * n[0], (rank // sx % sy) * n[1], (rank // (sx * sy)) * n[2]
        for nm, m_ in (("mxl", mx), ("myl", my), ("mzl", mz)):
            a = R.FArr((size0,), np.int64)
            a.flat[:] = m_
            setattr(g, nm, a)
        # particles.F90:339-344 (mx0, my0, mz0 count the ghost cells)
        g.x1in, g.x2in = F(1. * (ng // 2 + 1)), F(ng_[0] + ng - 1. * (ng // 2))
        g.y1in, g.y2in = F(1. * (ng // 2 + 1)), F(ng_[1] + ng - 1. * (ng // 2))
        g.z1in, g.z2in = F(1. * (ngz // 2 + 1)), F(ng_[2] + ngz - 1. * (ngz // 2))
        g.qi, g.qe = F(0.07), F(-0.07)
        g.debug, g.lap = False, 1
        for nm in ("curx", "cury", "curz"):
            getattr(g, nm).flat[:] = 0
        maxhlf, nsp = 96, 80
        p = np.zeros(2 * maxhlf, PDT)
        for s0 in (0, maxhlf):
            sl = slice(s0, s0 + nsp)
            # positions AFTER the push: most inside, a band of them up to 0.4 cells outside each face (they are what leaves)
            lo = np.array([ng // 2 + 1, ng // 2 + 1, ngz // 2 + 1], F)
            ext = np.array([n[0], n[1], n[2] if dim == 3 else 1], F)
            u01 = rng.random((3, nsp))
            pos = lo[:, None] + (u01 * 1.16 - 0.08) * ext[:, None] if False else lo[:, None] + u01 * ext[:, None]
            edge = rng.random((3, nsp))
            pos = np.where(edge < 0.12, lo[:, None] - 0.4 * rng.random((3, nsp)), pos)
            pos = np.where(edge > 0.88, lo[:, None] + ext[:, None] + 0.4 * rng.random((3, nsp)), pos)
            p["x"][sl], p["y"][sl], p["z"][sl] = pos.astype(F)
            for k in "uvw":
                p[k][sl] = (rng.standard_normal(nsp) * 0.6).astype(F)
            # a particle that is outside must have moved outwards: give its momentum the sign of its excursion
            for k, ax in (("u", 0), ("v", 1), ("w", 2)):
                below, above = pos[ax] < lo[ax], pos[ax] > lo[ax] + ext[ax]
                p[k][sl] = np.where(below, -np.abs(p[k][sl]) - F(1.2), np.where(above, np.abs(p[k][sl]) + F(1.2), p[k][sl]))
            p["ch"][sl] = (0.5 + rng.random(nsp)).astype(F)
            p["ind"][sl] = np.arange(1, nsp + 1) * (1 if s0 == 0 else -1)
            p["proc"][sl] = rank
            p["splitlev"][sl] = 1
        g.p = R.RecArr(p)
        g.ions, g.lecs, g.maxhlf = nsp, nsp, maxhlf
        g.pind = R.FArr((2 * maxhlf,), np.int64)
        bufs = {}
        for nm in ("poutup", "poutdwn", "poutlft", "poutrgt", "poutminus", "poutplus"):
            bufs[nm] = np.zeros(2 * nsp, PDT)
            setattr(g, nm, R.RecArr(bufs[nm]))
        g.q = F(0)

        def copyprt(a, b):                                  # particles.F90:263-276, field by field
            for k in PDT.names:
                setattr(b, k, getattr(a, k))
        g.copyprt = copyprt
        dsub = R.Sub(ptext, name, defines=defines, global_arrays=ga, global_ints=gi).compile()
        setattr(g, name, lambda *a: dsub(g, *a))
        sub = R.Sub(mtext, "deposit_particles", defines=defines, global_arrays=ga, global_ints=gi).compile()
        key = f"p{ci}"
        out[key + "_meta"] = np.array([dim, order, *per, *ng_], np.int32)
        out[key + "_geom"] = np.array([*sizes, rank, maxhlf, nsp], np.int32)
        out[key + "_pin"] = p.copy()
        sub(g)
        out[key + "_pout"] = p.copy()
        out[key + "_counts"] = np.array([g.ions, g.lecs, g.nionout, g.nlecout], np.int32)
        # out-buffers in the oracle's direction order: x-, x+, y-, y+, z-, z+; ions then electrons (the reference appends the
        # electrons behind the ions in the same buffer: LenLecOut* counts on from LenIonOut*, :1990-2030)
        lens = []
        for d, (nm, suf) in enumerate((("poutminus", "outminus"), ("poutplus", "outplus"), ("poutlft", "outlft"), ("poutrgt", "outrgt"),
                                       ("poutdwn", "outdwn"), ("poutup", "outup"))):
            ni, nl = getattr(g, "lenion" + suf), getattr(g, "lenlec" + suf)
            lens += [ni, nl]
            out[f"{key}_box{d}"] = bufs[nm].copy()
        out[key + "_boxlen"] = np.array(lens, np.int32)
        for a, nm in enumerate(("curx", "cury", "curz")):
            out[f"{key}_cur{a}"] = c_order(getattr(g, nm))
        print("deposit_particles", key, dim, order, sizes, rank, per, "left:", int(g.ions), int(g.lecs), "boxes:", lens)
    np.savez_compressed(os.path.join(OUT, "ref_depositp.npz"), **out)


# ------------------------------------------------------------------------------------------------------------
# G10: SEVERAL RANKS.  Every rank runs the reference's source in its own thread; MPI_SendRecv is a rendezvous between the
#      threads (f90run.Comm), so neighbour ranks, tags, message sizes and the z -> y -> x order are the reference's own.
#      bc_b1 / bc_e1 / exchange_current (fieldboundaries.F90:181-392, 1768-2189) and apply_filter1_opt (filter.F90:8-221)
# ------------------------------------------------------------------------------------------------------------
def rank_geometry(g, dim, order, nglob, sizes, rank):
    sx, sy, sz = sizes
    n = tuple(a // s_ for a, s_ in zip(nglob, sizes))
    ng, ngz, mx, my, mz = grid(dim, order, n)
    g.size0, g.sizex, g.sizey, g.sizez, g.rank = sx * sy * sz, sx, sy, sz, rank
    g.mxcum, g.mycum, g.mzcum = (rank % sx) * n[0], (rank // sx % sy) * n[1], (rank // (sx * sy)) * n[2]
    for nm, m_ in (("mxl", mx), ("myl", my), ("mzl", mz)):
        a = R.FArr((g.size0,), np.int64)
        a.flat[:] = m_
        setattr(g, nm, a)
    return n


MR_CASES = [(2, 1, (16, 12, 1), (2, 2, 1), (1, 1, 1)),
            (2, 2, (16, 12, 1), (2, 1, 1), (0, 1, 1)),
            (3, 2, (8, 12, 12), (1, 2, 2), (1, 1, 1)),              # 3D: sizex = 1 always (communications.F90:176-181)
            (3, 1, (8, 12, 12), (1, 2, 2), (1, 0, 1)),
            (3, 3, (8, 24, 8), (1, 3, 1), (0, 1, 1)),
            (2, 1, (24, 8, 1), (3, 1, 1), (0, 0, 1))]


def gen_halo_mr():
    out = {}
    fb, ft = src("fieldboundaries.F90"), src("filter.F90")
    names = ("ex", "ey", "ez", "bx", "by", "bz", "curx", "cury", "curz")
    for ci, (dim, order, nglob, sizes, per) in enumerate(MR_CASES):
        defines = {"MPI"} | ({"twoD"} if dim == 2 else set())
        size0 = sizes[0] * sizes[1] * sizes[2]
        comm = R.Comm(size0)
        ga = GARR | {"bufferin1x", "bufferin2x", "bufferin1y", "bufferin2y", "bufferin1", "bufferin2", "mxl", "myl", "mzl"}
        gi = GINTS | {"statsize", "mxcum", "mycum", "mzcum"}
        subs = {nm: R.Sub(fb, nm, defines=defines, global_arrays=ga, global_ints=gi).compile() for nm in ("bc_b1", "bc_e1", "exchange_current")}
        subs["apply_filter1_opt"] = R.Sub(ft, "apply_filter1_opt", defines=defines, global_arrays=ga, global_ints=gi).compile()
        gs = []
        key = f"x{ci}"
        out[key + "_meta"] = np.array([dim, order, *per, *nglob], np.int32)
        out[key + "_sizes"] = np.array(sizes, np.int32)
        for rank in range(size0):
            n = tuple(a // s_ for a, s_ in zip(nglob, sizes))
            g = field_globals(dim, order, n, per, np.random.default_rng(1100 + 16 * ci + rank))
            rank_geometry(g, dim, order, nglob, sizes, rank)
            g.comm = comm
            layer_copies(g, defines)
            g.highorder = False
            g.debug = False
            g.ntimes = 2
            g.temp = R.FArr((g.mx, g.my, g.mz))
            ng = g.nghost
            g.bufferin1x, g.bufferin2x = R.FArr((ng // 2 + 1, g.my, g.mz)), R.FArr((ng // 2, g.my, g.mz))
            g.bufferin1y, g.bufferin2y = R.FArr((g.mx, ng // 2 + 1, g.mz)), R.FArr((g.mx, ng // 2, g.mz))
            g.bufferin1, g.bufferin2 = R.FArr((g.mx, g.my, g.nghostz // 2 + 1)), R.FArr((g.mx, g.my, g.nghostz // 2))
            for nm in ("curx", "cury", "curz"):            # see gen_halo: the last layer never holds a deposit
                v = getattr(g, nm).nd()
                v[-1, :, :] = 0; v[:, -1, :] = 0
                if dim == 3:
                    v[:, :, -1] = 0
            for a, nm in enumerate(names):
                out[f"{key}_r{rank}_in{a}"] = c_order(getattr(g, nm))
            gs.append(g)

        def lap(g):
            subs["bc_b1"](g); subs["bc_e1"](g); subs["exchange_current"](g); subs["apply_filter1_opt"](g)
        R.run_ranks([(lambda g=g: lap(g)) for g in gs])
        for rank, g in enumerate(gs):
            for a, nm in enumerate(names):
                out[f"{key}_r{rank}_out{a}"] = c_order(getattr(g, nm))
        print("halo_mr", key, dim, order, nglob, sizes, per)
    np.savez_compressed(os.path.join(OUT, "ref_halo_mr.npz"), **out)


# ------------------------------------------------------------------------------------------------------------
# G11: particle migration over several ranks -- deposit_particles, exchange_particles, inject_others, exchange_particles,
#      inject_others (particles_movedeposit.F90:1281-2051, particles.F90:1368-2116; the call order of tristanmainloop.F90:
#      183-249), ranks as threads (see G10).  Corner crossers take two hops (x/y first, z on the second exchange).
# ------------------------------------------------------------------------------------------------------------
def gen_migrate_mr():
    out = {}
    mtext, ptext = src("particles_movedeposit.F90"), src("particles.F90")
    dirs = ("outup", "outdwn", "inblw", "inabv", "outlft", "outrgt", "inlft", "inrgt", "outminus", "outplus", "inminus", "inplus")
    gi = GINTS | {"mxcum", "mycum", "mzcum", "nionout", "nlecout", "lap", "statsize", "buffsize", "receivedions", "receivedlecs"} \
        | {f"len{s}{d}" for s in ("ion", "lec") for d in dirs}
    pbufs = ("poutup", "poutdwn", "poutlft", "poutrgt", "poutminus", "poutplus", "pinblw", "pinabv", "pinlft", "pinrgt", "pinminus", "pinplus")
    ga = GARR | {"pind", "mxl", "myl", "mzl"} | set(pbufs)
    dep = {0: ("zzag", "zigzag"), 1: ("dd1", "densdecomp_1ord"), 2: ("dd2", "densdecomp_2ord"), 3: ("dd3", "densdecomp_3ord")}
    cases = [(2, 1, (16, 12, 1), (2, 2, 1), (1, 1, 1)),
             (2, 2, (16, 12, 1), (2, 1, 1), (0, 1, 1)),
             (3, 2, (8, 12, 12), (1, 2, 2), (1, 1, 1)),
             (3, 0, (8, 12, 12), (1, 2, 2), (1, 0, 1)),
             (3, 1, (8, 24, 8), (1, 3, 1), (0, 1, 1)),
             (2, 0, (24, 8, 1), (3, 1, 1), (0, 0, 1))]
    for ci, (dim, order, nglob, sizes, per) in enumerate(cases):
        flag, name = dep[order]
        defines = {"MPI", flag} | ({"twoD"} if dim == 2 else set())
        size0 = sizes[0] * sizes[1] * sizes[2]
        comm = R.Comm(size0)
        subs = {nm: R.Sub(mtext if nm == "deposit_particles" else ptext, nm, defines=defines, global_arrays=ga, global_ints=gi).compile()
                for nm in ("deposit_particles", "exchange_particles", "inject_others", name)}
        key = f"y{ci}"
        out[key + "_meta"] = np.array([dim, order, *per, *nglob], np.int32)
        maxhlf, nsp = 160, 64
        out[key + "_geom"] = np.array([*sizes, maxhlf, nsp], np.int32)
        gs, ps = [], []
        for rank in range(size0):
            rng = np.random.default_rng(1200 + 16 * ci + rank)
            n = tuple(a // s_ for a, s_ in zip(nglob, sizes))
            g = field_globals(dim, order, n, per, rng)
            rank_geometry(g, dim, order, nglob, sizes, rank)
            ng, ngz, mx, my, mz = grid(dim, order, n)
            g.comm = comm
            g.x1in, g.x2in = F(1. * (ng // 2 + 1)), F(nglob[0] + ng - 1. * (ng // 2))
            g.y1in, g.y2in = F(1. * (ng // 2 + 1)), F(nglob[1] + ng - 1. * (ng // 2))
            g.z1in, g.z2in = F(1. * (ngz // 2 + 1)), F(nglob[2] + ngz - 1. * (ngz // 2))
            g.qi, g.qe = F(0.07), F(-0.07)
            g.debug, g.lap, g.statsize, g.buffsize = False, 1, 5, 4 * nsp
            g.mpi_comm_world = g.mpi_integer = g.particletype = 0
            g.mpi_wtime = lambda: 0.0                        # timers around the loops, not part of the result
            for nm in ("curx", "cury", "curz"):
                getattr(g, nm).flat[:] = 0
            p = np.zeros(2 * maxhlf, PDT)
            lo = np.array([ng // 2 + 1, ng // 2 + 1, ngz // 2 + 1], F)
            ext = np.array([n[0], n[1], n[2] if dim == 3 else 1], F)
            for s0 in (0, maxhlf):
                sl = slice(s0, s0 + nsp)
                pos = lo[:, None] + rng.random((3, nsp)) * ext[:, None]
                edge = rng.random((3, nsp))
                pos = np.where(edge < 0.15, lo[:, None] - 0.4 * rng.random((3, nsp)), pos)
                pos = np.where(edge > 0.85, lo[:, None] + ext[:, None] + 0.4 * rng.random((3, nsp)), pos)
                p["x"][sl], p["y"][sl], p["z"][sl] = pos.astype(F)
                for k, ax in (("u", 0), ("v", 1), ("w", 2)):
                    v = (rng.standard_normal(nsp) * 0.6).astype(F)
                    below, above = pos[ax] < lo[ax], pos[ax] > lo[ax] + ext[ax]
                    p[k][sl] = np.where(below, -np.abs(v) - F(1.2), np.where(above, np.abs(v) + F(1.2), v))
                p["ch"][sl] = (0.5 + rng.random(nsp)).astype(F)
                p["ind"][sl] = np.arange(1, nsp + 1) * (1 if s0 == 0 else -1)
                p["proc"][sl] = rank
                p["splitlev"][sl] = 1
            g.p = R.RecArr(p)
            g.ions, g.lecs, g.maxhlf = nsp, nsp, maxhlf
            g.pind = R.FArr((2 * maxhlf,), np.int64)
            for nm in pbufs:
                setattr(g, nm, R.RecArr(np.zeros(g.buffsize, PDT)))
            g.q = F(0)

            def copyprt(a, b):
                for k in PDT.names:
                    setattr(b, k, getattr(a, k))
            g.copyprt = copyprt
            setattr(g, name, (lambda g_: (lambda *a: subs[name](g_, *a)))(g))
            out[f"{key}_r{rank}_pin"] = p.copy()
            gs.append(g); ps.append(p)

        def lap(g):
            subs["deposit_particles"](g); subs["exchange_particles"](g); subs["inject_others"](g)
            subs["exchange_particles"](g); subs["inject_others"](g)
        R.run_ranks([(lambda g=g: lap(g)) for g in gs])
        tot = 0
        for rank, (g, p) in enumerate(zip(gs, ps)):
            out[f"{key}_r{rank}_pout"] = p.copy()
            out[f"{key}_r{rank}_counts"] = np.array([g.ions, g.lecs], np.int32)
            for a, nm in enumerate(("curx", "cury", "curz")):
                out[f"{key}_r{rank}_cur{a}"] = c_order(getattr(g, nm))
            tot += g.ions + g.lecs
        print("migrate_mr", key, dim, order, nglob, sizes, per, "particles", 2 * nsp * size0, "->", tot, [int(g.ions) for g in gs])
    np.savez_compressed(os.path.join(OUT, "ref_migrate_mr.npz"), **out)


# ------------------------------------------------------------------------------------------------------------
# G12: WHOLE LAPS.  `mainloop` itself (tristanmainloop.F90:60-330) is executed: the order of the calls is the reference's
#      text, every hot-path routine it calls is the reference's text, ranks are threads (G10).  Only the timers, the
#      diagnostics / output, the domain-rebalancing calls and the problem's injector are bound to no-ops.
# ------------------------------------------------------------------------------------------------------------
LAP_CASES = [  # dim order nglob       sizes      periodic   laps highorder shock ppc
    (2, 1, (16, 12, 1), (2, 2, 1), (1, 1, 1), 10, 0, 0, 2),
    (3, 2, (6, 12, 12), (1, 2, 2), (1, 1, 1), 3, 0, 0, 1),
    (3, 0, (7, 6, 6), (1, 1, 1), (0, 0, 0), 3, 0, 0, 2),
    (2, 1, (40, 8, 1), (2, 1, 1), (0, 1, 1), 3, 0, 1, 2),
    (3, 3, (6, 12, 6), (1, 2, 1), (1, 1, 1), 2, 1, 0, 1),
    # one-rank boxes: these are also run through the CUDA library (tests/test_gpu_ref_golden.py)
    (3, 2, (8, 8, 8), (1, 1, 1), (1, 1, 1), 3, 0, 0, 2),
    (2, 1, (16, 12, 1), (1, 1, 1), (1, 1, 1), 3, 0, 0, 4),
    (2, 2, (40, 8, 1), (1, 1, 1), (0, 1, 1), 3, 0, 1, 2),
    (3, 3, (8, 6, 6), (1, 1, 1), (1, 1, 1), 2, 1, 0, 2),
    # the `filter2` build (3D only, tristanmainloop.F90:213-229): one rank (also run on the GPU) and 1x2x2
    (3, 2, (8, 8, 8), (1, 1, 1), (1, 1, 1), 3, 0, 0, 2, 2),
    (3, 1, (6, 12, 12), (1, 2, 2), (1, 1, 1), 2, 0, 0, 1, 2),
    # remaining order x dimension builds and radiating axes that are split over ranks (oracle side only)
    (2, 3, (12, 10, 1), (1, 1, 1), (1, 1, 1), 3, 0, 0, 2),
    (2, 0, (12, 10, 1), (1, 1, 1), (1, 0, 1), 3, 0, 0, 2),
    (3, 2, (6, 12, 6), (1, 2, 1), (1, 0, 1), 3, 0, 0, 1),
    (2, 2, (16, 12, 1), (2, 2, 1), (0, 0, 1), 3, 0, 0, 2),
    # two-rank periodic boxes, also driven over gloo through the slab-exchange plan (tests/test_gloo_slabs.py)
    (2, 2, (16, 12, 1), (2, 1, 1), (1, 1, 1), 3, 0, 0, 2),
    (3, 2, (6, 6, 12), (1, 1, 2), (1, 1, 1), 2, 0, 0, 1, 2)]


def gen_lap():
    out = {}
    T_ = {nm: src(nm + ".F90") for nm in ("tristanmainloop", "fields", "fieldboundaries", "filter", "particles", "particles_movedeposit")}
    T_["user"] = open(os.path.join(REF, "user", "user_shock.F90")).read()
    main_text = "\n".join(l for l in T_["tristanmainloop"].split("\n") if not l.strip().lower().startswith("call timer"))
    dirs = ("outup", "outdwn", "inblw", "inabv", "outlft", "outrgt", "inlft", "inrgt", "outminus", "outplus", "inminus", "inplus")
    gi = GINTS | {"mxcum", "mycum", "mzcum", "nionout", "nlecout", "lap", "lapst", "last", "statsize", "buffsize", "receivedions",
                  "receivedlecs", "radiationx", "radiationy", "radiationz", "mx0", "lapreorder", "outcorner"} \
        | {f"len{s_}{d}" for s_ in ("ion", "lec") for d in dirs}
    pbufs = ("poutup", "poutdwn", "poutlft", "poutrgt", "poutminus", "poutplus", "pinblw", "pinabv", "pinlft", "pinrgt", "pinminus", "pinplus")
    ga = GARR | {"pind", "pall", "tempp", "mxl", "myl", "mzl", "bufferin1x", "bufferin2x", "bufferin1y", "bufferin2y", "bufferin1", "bufferin2"} | set(pbufs)
    mov = {0: ("zzag", "mover", "zigzag"), 1: ("dd1", "mover_1ord", "densdecomp_1ord"), 2: ("dd2", "mover_2ord", "densdecomp_2ord"),
           3: ("dd3", "mover_3ord", "densdecomp_3ord")}
    where = {"mainloop": "main", "advance_bhalfstep": "fields", "advance_efield": "fields", "advance_b_halfstep": "fields",
             "advance_e_fullstep": "fields", "advance_b_halfstep_42": "fields", "advance_e_fullstep_42": "fields", "add_current": "fields",
             "reset_currents": "fields", "iloc": "fields", "xglob": "fields",
             "move_particles": "particles_movedeposit", "deposit_particles": "particles_movedeposit",
             "exchange_particles": "particles", "inject_others": "particles", "reorder_particles": "particles",
             "reorder_particles_": "particles", "zigzag": "particles",
             "apply_filter1_opt": "filter", "field_bc_user": "user", "particle_bc_user": "user"}
    for nm in ("bc_b1", "bc_e1", "bc_b2", "bc_e2", "pre_bc_b", "post_bc_b", "pre_bc_e", "post_bc_e", "surface", "preledge", "postedge",
               "exchange_current", "copylayrx", "copylayry", "copy_layrx1_opt", "copy_layry1_opt", "copy_layrz1_opt", "copy_layrx2_opt",
               "copy_layry2_opt", "copy_layrz2_opt"):
        where[nm] = "fieldboundaries"
    noop = ("diagnostics", "inject_particles", "check_overflow", "enlarge_domain", "redist_x_domain", "redist_y_domain", "shift_domain",
            "redist_z_domain", "print_timers", "pause_simulation")
    names9 = ("ex", "ey", "ez", "bx", "by", "bz", "curx", "cury", "curz")
    otext = "\n".join(l for l in src("optimized_filters.F90").split("\n") if not l.strip().lower().startswith("call timer"))
    f2names = ["apply_filter2_opt", "filter_x", "filter_y", "filter_z", "deep_copylayrx", "nonper_copylayrx"] + \
              [f"deep_copy_layr{a}{v}" for a in "xyz" for v in "12"]
    ga = ga | {"xghost", "yghost", "zghost"}
    for ci, case in enumerate(LAP_CASES):
        dim, order, nglob, sizes, per, laps, highorder, shock, ppc = case[:9]
        fkind = case[9] if len(case) > 9 else 1
        flag, mover, dname = mov[order]
        defines = {"MPI", flag} | ({"twoD"} if dim == 2 else set()) | ({"filter2"} if fkind == 2 else set())
        size0 = sizes[0] * sizes[1] * sizes[2]
        comm = R.Comm(size0, timeout=600.0)
        need = dict(where)
        need[mover] = "particles_movedeposit"
        need[dname] = "particles"
        if not shock:
            del need["field_bc_user"], need["particle_bc_user"]
        subs = {nm: R.Sub(main_text if f == "main" else T_[f], nm, defines=defines, global_arrays=ga, global_ints=gi).compile()
                for nm, f in need.items()}
        if fkind == 2:
            for nm in f2names:
                subs[nm] = R.Sub(otext, nm, defines=defines, global_arrays=ga, global_ints=gi).compile()
            subs["create_mpi_filter_datatypes"] = R.Sub(T_["fields"], "create_mpi_filter_datatypes", defines=defines, global_arrays=ga,
                                                        global_ints=gi).compile()
        key = f"l{ci}"
        out[key + "_meta"] = np.array([dim, order, *per, *nglob], np.int32)
        n = tuple(a // s_ for a, s_ in zip(nglob, sizes))
        ng, ngz, mx, my, mz = grid(dim, order, n)
        ncell = n[0] * n[1] * (n[2] if dim == 3 else 1)
        nsp = ppc * ncell
        maxhlf = 2 * nsp + 64
        out[key + "_geom"] = np.array([*sizes, maxhlf, nsp, laps, highorder, shock, fkind], np.int32)
        gs, ps = [], []
        for rank in range(size0):
            rng = np.random.default_rng(1300 + 16 * ci + rank)
            g = field_globals(dim, order, n, per, rng)
            for nm in names9[:6]:                                   # small-amplitude fields
                getattr(g, nm).flat[:] *= F(0.2)
            rank_geometry(g, dim, order, nglob, sizes, rank)
            g.comm = comm
            g.radiationx, g.radiationy, g.radiationz = 1 - per[0], 1 - per[1], 1 - per[2]
            if dim == 2 and g.radiationy == 1:
                g.radiationz = 1
            g.mx0 = nglob[0] + ng
            g.x1in, g.x2in = F(1. * (ng // 2 + 1)), F(nglob[0] + ng - 1. * (ng // 2))
            g.y1in, g.y2in = F(1. * (ng // 2 + 1)), F(nglob[1] + ng - 1. * (ng // 2))
            g.z1in, g.z2in = F(1. * (ngz // 2 + 1)), F(nglob[2] + ngz - 1. * (ngz // 2))
            g.qi, g.qe, g.qmi, g.qme = F(0.07), F(-0.07), F(0.5), F(-1.0)
            g.debug, g.statsize, g.buffsize = False, 5, maxhlf
            g.lap, g.lapst, g.last, g.lapreorder, g.outcorner = 0, 1, laps, -5, 0
            g.highorder, g.external_fields, g.ntimes = bool(highorder), False, 2
            g.wall, g.user_part_bcs = bool(shock), bool(shock)
            g.delgam, g.qme_abs = F(1e-2), F(1.0)
            g.xinject2 = F(nglob[0] + ng)
            g.leftwall, g.binit, g.btheta, g.bphi, g.beta = F(8.0), F(0.02), F(1.1), F(0.4), F(0.3)
            g.mpi_comm_world = g.mpi_integer = g.particletype = g.mpi_read = 0
            g.mpi_wtime = lambda: 0.0
            g.temp = R.FArr((g.mx, g.my, g.mz))
            g.bufferin1x, g.bufferin2x = R.FArr((ng // 2 + 1, g.my, g.mz)), R.FArr((ng // 2, g.my, g.mz))
            g.bufferin1y, g.bufferin2y = R.FArr((g.mx, ng // 2 + 1, g.mz)), R.FArr((g.mx, ng // 2, g.mz))
            g.bufferin1, g.bufferin2 = R.FArr((g.mx, g.my, g.nghostz // 2 + 1)), R.FArr((g.mx, g.my, g.nghostz // 2))
            for nm in ("curx", "cury", "curz"):
                getattr(g, nm).flat[:] = 0
            p = np.zeros(2 * maxhlf, PDT)
            lo = np.array([ng // 2 + 1, ng // 2 + 1, ngz // 2 + 1], F)
            ext = np.array([n[0], n[1], n[2] if dim == 3 else 1], F)
            for s0 in (0, maxhlf):
                sl = slice(s0, s0 + nsp)
                pos = lo[:, None] + rng.random((3, nsp)) * ext[:, None]
                if shock and rank == 0:                              # nothing starts behind the wall
                    pos[0] = np.maximum(pos[0], 8.5)
                p["x"][sl], p["y"][sl], p["z"][sl] = pos.astype(F)
                for k in "uvw":
                    p[k][sl] = (rng.standard_normal(nsp) * 0.5).astype(F)
                if shock:
                    p["u"][sl] -= F(0.6)                             # a stream towards the wall
                p["ch"][sl] = 1.0
                p["ind"][sl] = np.arange(1, nsp + 1) * (1 if s0 == 0 else -1)
                p["proc"][sl] = rank
                p["splitlev"][sl] = 1
            g.p = R.RecArr(p)
            g.tempp = R.RecArr(np.zeros(maxhlf, PDT))
            g.pall = R.FArr((g.lot,), np.int64)
            g.ions, g.lecs, g.maxhlf = nsp, nsp, maxhlf
            g.pind = R.FArr((2 * maxhlf,), np.int64)
            for nm in pbufs:
                setattr(g, nm, R.RecArr(np.zeros(g.buffsize, PDT)))
            g.q = F(0)

            def copyprt(a, b):
                for k in PDT.names:
                    setattr(b, k, getattr(a, k))
            g.copyprt = copyprt
            for nm, f in subs.items():
                setattr(g, nm, (lambda f_, g_: (lambda *a: f_(g_, *a)))(f, g))
            for nm in noop:
                setattr(g, nm, lambda *a: None)
            if fkind == 2:                                          # allocate_fields, fields.F90:353, 363-365
                nt = g.ntimes
                g.mpi_order_fortran = 0
                g.xghost, g.yghost, g.zghost = R.FArr((2 * nt, g.my, g.mz)), R.FArr((g.mx, 2 * nt, g.mz)), R.FArr((g.mx, g.my, 2 * nt))
                g.create_mpi_filter_datatypes(ng // 2 + 1, g.mx - (ng // 2 + 1), ng // 2 + 1, g.my - (ng // 2 + 1), ngz // 2 + 1,
                                              g.mz - (ngz // 2 + 1))
            if not shock:
                g.field_bc_user = lambda: None
                g.particle_bc_user = lambda: None
            for a, nm in enumerate(names9[:6]):
                out[f"{key}_r{rank}_in{a}"] = c_order(getattr(g, nm))
            out[f"{key}_r{rank}_pin"] = p.copy()
            gs.append(g); ps.append(p)
        R.OVERLONG.clear()
        R.run_ranks([(lambda g=g: g.mainloop()) for g in gs])
        for rank, (g, p) in enumerate(zip(gs, ps)):
            for a, nm in enumerate(names9):
                out[f"{key}_r{rank}_out{a}"] = c_order(getattr(g, nm))
            out[f"{key}_r{rank}_pout"] = p.copy()
            out[f"{key}_r{rank}_counts"] = np.array([g.ions, g.lecs], np.int32)
        out[key + "_par"] = np.array([gs[0].leftwall, gs[0].binit, gs[0].btheta, gs[0].bphi, gs[0].beta, gs[0].qi, gs[0].qe, gs[0].qmi, gs[0].qme], F)
        print("lap", key, dim, order, nglob, sizes, per, "laps", laps, "counts", [(int(g.ions), int(g.lecs)) for g in gs], "overlong msgs", len(R.OVERLONG))
    np.savez_compressed(os.path.join(OUT, "ref_lap.npz"), **out)


# ------------------------------------------------------------------------------------------------------------
# G13: filter2 over several ranks -- apply_filter2_opt (optimized_filters.F90:9-227) with its deep ghost exchange
#      deep_copy_layr{x,y,z}{1,2} (:1387-1963) through the MPI derived datatypes the reference builds in
#      create_MPI_filter_datatypes (fields.F90:1449-1600; MPI_Type_create_subarray is executed as "remember this box")
# ------------------------------------------------------------------------------------------------------------
F2_CASES = [(3, 2, (8, 12, 12), (1, 2, 2), (1, 1, 1), 3),
            (3, 1, (8, 18, 6), (1, 3, 1), (0, 1, 1), 4),
            (3, 3, (6, 8, 16), (1, 1, 2), (1, 0, 0), 2),
            (3, 2, (9, 8, 7), (1, 1, 1), (1, 1, 1), 5)]


def gen_filter2_mr():
    out = {}
    ftext = src("fields.F90")
    otext = "\n".join(l for l in src("optimized_filters.F90").split("\n") if not l.strip().lower().startswith("call timer"))
    caps = {a + b + "cap" for a in ("x", "y", "z", "xghost", "yghost", "zghost") for b in ("high", "low")}
    ga = GARR | {"xghost", "yghost", "zghost", "mxl", "myl", "mzl"}
    gi = GINTS | {"statsize", "mxcum", "mycum", "mzcum"}
    onames = ["apply_filter2_opt", "filter_x", "filter_y", "filter_z", "deep_copylayrx", "nonper_copylayrx"] + \
             [f"deep_copy_layr{a}{v}" for a in "xyz" for v in "12"]
    for ci, (dim, order, nglob, sizes, per, ntimes) in enumerate(F2_CASES):
        defines = {"MPI", "filter2"}
        size0 = sizes[0] * sizes[1] * sizes[2]
        comm = R.Comm(size0)
        subs = {nm: R.Sub(otext, nm, defines=defines, global_arrays=ga, global_ints=gi).compile() for nm in onames}
        subs["create_mpi_filter_datatypes"] = R.Sub(ftext, "create_mpi_filter_datatypes", defines=defines, global_arrays=ga, global_ints=gi).compile()
        key = f"z{ci}"
        out[key + "_meta"] = np.array([dim, order, *per, *nglob, ntimes], np.int32)
        out[key + "_sizes"] = np.array(sizes, np.int32)
        gs = []
        for rank in range(size0):
            n = tuple(a // s_ for a, s_ in zip(nglob, sizes))
            g = field_globals(dim, order, n, per, np.random.default_rng(1400 + 16 * ci + rank))
            rank_geometry(g, dim, order, nglob, sizes, rank)
            g.comm, g.ntimes, g.debug, g.statsize = comm, ntimes, False, 5
            g.mpi_comm_world = g.mpi_read = g.mpi_order_fortran = 0
            for nm, f in subs.items():
                setattr(g, nm, (lambda f_, g_: (lambda *a: f_(g_, *a)))(f, g))
            # allocate_fields, fields.F90:353, 363-365
            g.xghost, g.yghost, g.zghost = R.FArr((2 * ntimes, g.my, g.mz)), R.FArr((g.mx, 2 * ntimes, g.mz)), R.FArr((g.mx, g.my, 2 * ntimes))
            ng, ngz = g.nghost, g.nghostz
            g.create_mpi_filter_datatypes(ng // 2 + 1, g.mx - (ng // 2 + 1), ng // 2 + 1, g.my - (ng // 2 + 1), ngz // 2 + 1, g.mz - (ngz // 2 + 1))
            for c, nm in enumerate(("curx", "cury", "curz")):
                out[f"{key}_r{rank}_in{c}"] = c_order(getattr(g, nm))
            gs.append(g)
        R.run_ranks([(lambda g=g: g.apply_filter2_opt()) for g in gs])
        for rank, g in enumerate(gs):
            for c, nm in enumerate(("curx", "cury", "curz")):
                out[f"{key}_r{rank}_out{c}"] = c_order(getattr(g, nm))
        print("filter2_mr", key, dim, order, nglob, sizes, per, ntimes)
    np.savez_compressed(os.path.join(OUT, "ref_filter2_mr.npz"), **out)


# ------------------------------------------------------------------------------------------------------------
# G14: output-side moments -- meanq_fld_cur(totname) (output.F90:5229-5486): box deposit of half-width idx = idy = idz = 2
#      (idz = 0 in 2D), exchange_current, normalisation by the clipped box volume, ratio to the weight; one and several ranks
# ------------------------------------------------------------------------------------------------------------
MEANQ_NAMES = ("tdens", "idens", "hdens", "ldens", "btden", "tbetx", "ebety", "ibetz", "tmomx", "imomy", "eener", "iener", "eetx2", "iety2")
MEANQ_CASES = [(2, 1, (12, 10, 1), (1, 1, 1)), (3, 2, (8, 8, 6), (1, 1, 1)), (2, 2, (16, 12, 1), (2, 2, 1)), (3, 1, (6, 12, 12), (1, 2, 2))]


def gen_meanq():
    out = {}
    fb, otext = src("fieldboundaries.F90"), src("output.F90")
    out["names"] = np.array(MEANQ_NAMES)
    for ci, (dim, order, nglob, sizes) in enumerate(MEANQ_CASES):
        defines = {"MPI"} | ({"twoD"} if dim == 2 else set())
        size0 = sizes[0] * sizes[1] * sizes[2]
        comm = R.Comm(size0)
        ga = GARR | {"bufferin1x", "bufferin2x", "bufferin1y", "bufferin2y", "bufferin1", "bufferin2", "mxl", "myl", "mzl"}
        gi = GINTS | {"statsize", "mxcum", "mycum", "mzcum", "maxptl", "idx", "idy", "idz"}
        exc = R.Sub(fb, "exchange_current", defines=defines, global_arrays=ga, global_ints=gi).compile()
        mq = R.Sub(otext, "meanq_fld_cur", defines=defines, global_arrays=ga, global_ints=gi).compile()
        key = f"q{ci}"
        n = tuple(a // s_ for a, s_ in zip(nglob, sizes))
        ncell = n[0] * n[1] * (n[2] if dim == 3 else 1)
        nsp, maxhlf = 3 * ncell, 3 * ncell + 32
        out[key + "_meta"] = np.array([dim, order, 1, 1, 1, *nglob], np.int32)
        out[key + "_geom"] = np.array([*sizes, maxhlf, nsp], np.int32)
        gs = []
        for rank in range(size0):
            rng = np.random.default_rng(1500 + 16 * ci + rank)
            g = field_globals(dim, order, n, (1, 1, 1), rng)
            rank_geometry(g, dim, order, nglob, sizes, rank)
            ng, ngz, mx, my, mz = grid(dim, order, n)
            g.comm, g.debug, g.statsize, g.mpi_comm_world, g.mpi_read = comm, False, 5, 0, 0
            g.idx, g.idy, g.idz = 2, 2, (2 if dim == 3 else 0)                 # output.F90:189-195
            g.bufferin1x, g.bufferin2x = R.FArr((ng // 2 + 1, g.my, g.mz)), R.FArr((ng // 2, g.my, g.mz))
            g.bufferin1y, g.bufferin2y = R.FArr((g.mx, ng // 2 + 1, g.mz)), R.FArr((g.mx, ng // 2, g.mz))
            g.bufferin1, g.bufferin2 = R.FArr((g.mx, g.my, g.nghostz // 2 + 1)), R.FArr((g.mx, g.my, g.nghostz // 2))
            p = np.zeros(2 * maxhlf, PDT)
            lo = np.array([ng // 2 + 1, ng // 2 + 1, ngz // 2 + 1], F)
            ext = np.array([n[0], n[1], n[2] if dim == 3 else 1], F)
            for s0 in (0, maxhlf):
                sl = slice(s0, s0 + nsp)
                pos = lo[:, None] + rng.random((3, nsp)) * ext[:, None]
                p["x"][sl], p["y"][sl], p["z"][sl] = pos.astype(F)
                for k in "uvw":
                    p[k][sl] = (rng.standard_normal(nsp) * 0.7).astype(F)
                p["ch"][sl] = (0.5 + rng.random(nsp)).astype(F)
                p["ind"][sl] = np.arange(1, nsp + 1) * np.where(rng.random(nsp) < 0.3, -1, 1)
                p["proc"][sl] = rank
                p["splitlev"][sl] = 1
            g.p = R.RecArr(p)
            g.ions, g.lecs, g.maxhlf, g.maxptl = nsp, nsp, maxhlf, 2 * maxhlf
            g.exchange_current = (lambda g_: (lambda: exc(g_)))(g)
            out[f"{key}_r{rank}_pin"] = p.copy()
            gs.append(g)
        for ni, name in enumerate(MEANQ_NAMES):
            R.run_ranks([(lambda g=g: mq(g, name)) for g in gs])
            for rank, g in enumerate(gs):
                out[f"{key}_r{rank}_{name}"] = c_order(g.curx)
        print("meanq", key, dim, order, nglob, sizes, "max tdens", float(out[f"{key}_r0_tdens"].max()))
    np.savez_compressed(os.path.join(OUT, "ref_meanq.npz"), **out)


# ------------------------------------------------------------------------------------------------------------
# G15: the seeded loader -- init_particle_distribution_user of user/user_weibel.F90:250-296 calling inject_plasma_region
#      (particles.F90:2549-2938), init_maxw_table (:2126-2163), maxwell_dist (:2176-2273), random / poisson (aux.F90:82-134,
#      fp64 MINSTD), reorder_particles (:394-497).  exp / cos / sin go through libm's float entry points (f90run.elementary),
#      as in a compiled build.  `dseed` is the module variable every call site passes (alias_globals).
# ------------------------------------------------------------------------------------------------------------
LOADER_CASES = [  # dim order nglob      sizes      ppc0 distr_dim delgam  gamma0
    (2, 1, (12, 10, 1), (1, 1, 1), 8.0, 2, 2e-5, 0.5),
    (3, 2, (6, 5, 4), (1, 1, 1), 4.0, 3, 1e-2, 0.5),
    (2, 2, (16, 12, 1), (2, 2, 1), 4.0, 3, 5e-3, 3.0),
    (2, 1, (3, 2, 1), (1, 1, 1), 3.0, 2, 1e-3, 0.3),            # numps < 10: the Poisson branch
    (3, 3, (6, 8, 8), (1, 2, 2), 2.0, 3, 2e-5, 0.5),
    # user/user_twostream.F90:226-270 (configs[1]: dd2, a 2-cell-wide y axis, electrons only)
    (2, 2, (32, 2, 1), (1, 1, 1), 16.0, 2, 2e-5, 0.5, "twostream")]


def gen_loader():
    out = {}
    aux, ptext = src("aux.F90"), src("particles.F90")
    users = {u: open(os.path.join(REF, "user", f"user_{u}.F90")).read() for u in ("weibel", "twostream")}
    gi = GINTS | {"mxcum", "mycum", "mzcum", "lap", "lapreorder", "totalpartnum", "injectedions", "injectedlecs", "pdf_sz", "mx0", "my0",
                  "mz0", "myall", "mzall"}
    ga = GARR | {"pall", "tempp"}
    for ci, case in enumerate(LOADER_CASES):
        dim, order, nglob, sizes, ppc0, distr_dim, delgam, gamma0 = case[:8]
        user = users[case[8] if len(case) > 8 else "weibel"]
        defines = {"MPI"} | ({"twoD"} if dim == 2 else set())
        kw = dict(defines=defines, global_arrays=ga, global_ints=gi, alias_globals={"dseed"})
        subs = {nm: R.Sub(aux, nm, **kw).compile() for nm in ("random", "poisson")}
        for nm in ("init_maxw_table", "maxwell_dist", "inject_plasma_region", "reorder_particles", "reorder_particles_"):
            subs[nm] = R.Sub(ptext, nm, **kw).compile()
        subs["init_particle_distribution_user"] = R.Sub(user, "init_particle_distribution_user", **kw).compile()
        size0 = sizes[0] * sizes[1] * sizes[2]
        key = f"w{ci}"
        n = tuple(a // s_ for a, s_ in zip(nglob, sizes))
        ncell = n[0] * n[1] * (n[2] if dim == 3 else 1)
        maxhlf = int(2 * ppc0 * ncell) + 64
        out[key + "_meta"] = np.array([dim, order, 1, 1, 1, *nglob], np.int32)
        out[key + "_geom"] = np.array([*sizes, maxhlf, distr_dim, 1 if len(case) > 8 else 0], np.int32)
        out[key + "_par"] = np.array([ppc0, gamma0, delgam], F)
        for rank in range(size0):
            g = field_globals(dim, order, n, (1, 1, 1), np.random.default_rng(1))
            rank_geometry(g, dim, order, nglob, sizes, rank)
            ng, ngz, mx, my, mz = grid(dim, order, n)
            g.mx0, g.my0, g.mz0 = nglob[0] + ng, nglob[1] + ng, (nglob[2] + ngz if dim == 3 else 1)
            g.myall, g.mzall = g.my0, g.mz0
            g.pdf_sz = 1000                                                  # particles.F90:45
            g.pi = np.float64(F(3.1415927))                                  # :73, 292 -- an fp32 literal in an fp64 variable
            g.dseed = np.float64(123457.0) + rank                            # communications.F90:228-229
            g.pcosthmult = F(0.0) if (dim == 2 and distr_dim == 2) else F(1.0)       # user_weibel.F90:160-167
            g.sigma, g.c_omp, g.ppc0, g.delgam = F(0.0), F(10.0), F(ppc0), F(delgam)
            g.me, g.mi, g.temperature_ratio = F(1.0), F(1.0), F(1.0)
            g0 = F(gamma0)
            g.gamma0 = F(np.sqrt(F(F(1.) / F(F(1.) - F(g0 * g0))))) if g0 < 1 else g0          # particles.F90:213
            g.xinject = g.xinject2 = F(0)
            g.debug, g.lap, g.lapreorder = False, 0, 0
            g.totalpartnum = g.injectedions = g.injectedlecs = 0
            p = np.zeros(2 * maxhlf, PDT)
            g.p, g.tempp, g.pall = R.RecArr(p), R.RecArr(np.zeros(maxhlf, PDT)), R.FArr((g.lot,), np.int64)
            g.ions, g.lecs, g.maxhlf = 0, 0, maxhlf
            for nm, f in subs.items():
                setattr(g, nm, (lambda f_, g_: (lambda *a, **k: f_(g_, *a, **k)))(f, g))
            g.check_overflow = g.check_overflow_num = lambda *a: None

            def copyprt(a, b):
                for k in PDT.names:
                    setattr(b, k, getattr(a, k))
            g.copyprt = copyprt
            g.init_particle_distribution_user()
            out[f"{key}_r{rank}_p"] = p.copy()
            out[f"{key}_r{rank}_counts"] = np.array([g.ions, g.lecs, g.totalpartnum], np.int64)
            out[f"{key}_r{rank}_dseed"] = np.array([g.dseed], np.float64)
            print("loader", key, "rank", rank, "ions", g.ions, "lecs", g.lecs, "dseed", g.dseed, "mean u", float(p["u"][:g.ions].mean()) if g.ions else 0)
    np.savez_compressed(os.path.join(OUT, "ref_loader.npz"), **out)


# ------------------------------------------------------------------------------------------------------------
# G16: save_spectrum (output.F90:380-633), the computing part: gamma range, lab-frame and flow-frame spectra per x slice.
#      The routine is executed up to the line where it starts writing files.  What each rank contributes is captured where it
#      crosses the MPI boundary (the send buffers of the routine's six MPI_Allreduce calls, f90run.Globals.allreduce_log).
# ------------------------------------------------------------------------------------------------------------
SPEC_CASES = [(2, 1, (260, 6, 1), (1, 1, 1)), (3, 2, (210, 4, 4), (1, 1, 1)), (2, 1, (440, 6, 1), (2, 1, 1))]


def gen_spectrum():
    out = {}
    text = src("output.F90")
    import re
    a = re.search(r"^[ \t]*subroutine save_spectrum", text, flags=re.I | re.M).start()
    b = text.lower().index('"done spec calculation', a)
    b = text.rfind("\n", 0, b)
    text = text[a:b] + "\n\tendif\nend subroutine save_spectrum\n"
    gi = GINTS | {"mxcum", "lap", "pltstart", "interval", "mx0", "mpi_status_size"}
    for ci, (dim, order, nglob, sizes) in enumerate(SPEC_CASES):
        defines = {"MPI"} | ({"twoD"} if dim == 2 else set())
        sub = R.Sub(text, "save_spectrum", defines=defines, global_arrays=GARR, global_ints=gi).compile()
        size0 = sizes[0] * sizes[1] * sizes[2]
        comm = R.Comm(size0) if size0 > 1 else None
        key = f"e{ci}"
        n = tuple(a_ // s_ for a_, s_ in zip(nglob, sizes))
        ng, ngz, mx, my, mz = grid(dim, order, n)
        nsp, maxhlf = 600, 640
        out[key + "_meta"] = np.array([dim, order, 1, 1, 1, *nglob], np.int32)
        out[key + "_geom"] = np.array([*sizes, maxhlf, nsp], np.int32)
        gs = []
        for rank in range(size0):
            rng = np.random.default_rng(1600 + 16 * ci + rank)
            g = field_globals(dim, order, n, (1, 1, 1), rng)
            rank_geometry(g, dim, order, nglob, sizes, rank)
            g.comm = comm
            g.mx0 = nglob[0] + ng
            g.lap, g.pltstart, g.interval, g.debug = 0, 0, 50, False
            g.splitratio = F(10.0)
            g.mpi_status_size, g.mpi_read, g.mpi_comm_world = 5, 0, 0
            g.mpi_max, g.mpi_min, g.mpi_sum = "max", "min", "sum"
            p = np.zeros(2 * maxhlf, PDT)
            for s0 in (0, maxhlf):
                sl = slice(s0, s0 + nsp)
                p["x"][sl] = ((ng // 2 + 1) + rng.random(nsp) * n[0]).astype(F)
                p["y"][sl] = ((ng // 2 + 1) + rng.random(nsp) * n[1]).astype(F)
                p["z"][sl] = ((ngz // 2 + 1) + rng.random(nsp) * (n[2] if dim == 3 else 1)).astype(F)
                spread = np.exp(rng.standard_normal(nsp) * 1.2) * 0.3              # a wide range of energies
                for k in "uvw":
                    p[k][sl] = (rng.standard_normal(nsp) * spread).astype(F)
                p["u"][sl] += F(0.4)                                                  # a mean flow to boost out of
                p["ch"][sl] = (0.5 + rng.random(nsp)).astype(F)
                p["splitlev"][sl] = np.where(rng.random(nsp) < 0.2, 2, 1)
                p["ind"][sl] = np.arange(1, nsp + 1)
                p["proc"][sl] = rank
            g.p = R.RecArr(p)
            g.ions, g.lecs, g.maxhlf = nsp, nsp, maxhlf
            out[f"{key}_r{rank}_pin"] = p.copy()
            gs.append(g)
        R.run_ranks([(lambda g=g: sub(g)) for g in gs])
        for rank, g in enumerate(gs):
            log = g.allreduce_log            # gammax, gammin, specp, specpprime, spece, speceprime
            assert len(log) == 6
            out[f"{key}_r{rank}_range"] = np.array([log[1], log[0]], F).reshape(2)
            for nm, v in zip(("specp", "specpprime", "spece", "speceprime"), log[2:]):
                out[f"{key}_r{rank}_{nm}"] = np.asarray(v, F)
        print("spectrum", key, dim, order, nglob, sizes, "range", out[f"{key}_r0_range"], "counts in spectrum", float(out[f"{key}_r0_specp"].sum()))
    np.savez_compressed(os.path.join(OUT, "ref_spectrum.npz"), **out)


# ------------------------------------------------------------------------------------------------------------
# G17: the shock problem's per-lap injector -- inject_particles_user (user/user_shock.F90:303-331) -> inject_from_wall
#      (particles.F90:2439-2538) -> inject_plasma_region; three consecutive calls append to the same arrays
# ------------------------------------------------------------------------------------------------------------
# With nghost = 7 (dd2, dd3) the hard-coded source plane x = mx0 - 2 (an nghost = 5 number) lies outside the interior, whose last
# node is mx0 - 3: inject_plasma_region clamps the slab to zero width and the injector adds nothing (the last case pins that).
INJ_CASES = [(2, 1, (40, 12, 1), (2, 1, 1), 8.0, 2e-3, 0.5, 0), (3, 1, (20, 6, 5), (1, 1, 1), 4.0, 1e-2, 4.0, 1), (2, 0, (30, 3, 1), (1, 1, 1), 2.0, 1e-3, 0.3, 0),
             (3, 3, (20, 6, 5), (1, 1, 1), 4.0, 1e-2, 0.5, 1)]


def gen_injector():
    out = {}
    aux, ptext = src("aux.F90"), src("particles.F90")
    user = open(os.path.join(REF, "user", "user_shock.F90")).read()
    gi = GINTS | {"mxcum", "mycum", "mzcum", "lap", "totalpartnum", "injectedions", "injectedlecs", "pdf_sz", "mx0", "my0", "mz0", "myall", "mzall"}
    for ci, (dim, order, nglob, sizes, ppc0, delgam, gamma0, pcm) in enumerate(INJ_CASES):
        defines = {"MPI"} | ({"twoD"} if dim == 2 else set())
        kw = dict(defines=defines, global_arrays=GARR, global_ints=gi, alias_globals={"dseed"})
        subs = {nm: R.Sub(aux, nm, **kw).compile() for nm in ("random", "poisson")}
        for nm in ("init_maxw_table", "maxwell_dist", "inject_plasma_region", "inject_from_wall"):
            subs[nm] = R.Sub(ptext, nm, **kw).compile()
        subs["inject_particles_user"] = R.Sub(user, "inject_particles_user", **kw).compile()
        size0 = sizes[0] * sizes[1] * sizes[2]
        key = f"j{ci}"
        n = tuple(a // s_ for a, s_ in zip(nglob, sizes))
        maxhlf = 2000
        out[key + "_meta"] = np.array([dim, order, 0, 1, 1, *nglob], np.int32)
        out[key + "_geom"] = np.array([*sizes, maxhlf, pcm], np.int32)
        out[key + "_par"] = np.array([ppc0, gamma0, delgam], F)
        for rank in range(size0):
            g = field_globals(dim, order, n, (0, 1, 1), np.random.default_rng(1))
            rank_geometry(g, dim, order, nglob, sizes, rank)
            ng, ngz, mx, my, mz = grid(dim, order, n)
            g.mx0, g.my0, g.mz0 = nglob[0] + ng, nglob[1] + ng, (nglob[2] + ngz if dim == 3 else 1)
            g.myall, g.mzall, g.pdf_sz = g.my0, g.mz0, 1000
            g.pi = np.float64(F(3.1415927))
            g.dseed = np.float64(123457.0) + rank
            g.pcosthmult, g.sigma, g.c_omp, g.ppc0, g.delgam = F(pcm), F(0.0), F(10.0), F(ppc0), F(delgam)
            g.me, g.mi, g.temperature_ratio = F(1.0), F(1.0), F(1.0)
            g0 = F(gamma0)
            g.gamma0 = F(np.sqrt(F(F(1.) / F(F(1.) - F(g0 * g0))))) if g0 < 1 else g0
            g.debug, g.lap = False, 1
            g.totalpartnum = g.injectedions = g.injectedlecs = 0
            p = np.zeros(2 * maxhlf, PDT)
            g.p = R.RecArr(p)
            g.ions, g.lecs, g.maxhlf = 0, 0, maxhlf
            for nm, f in subs.items():
                setattr(g, nm, (lambda f_, g_: (lambda *a, **k: f_(g_, *a, **k)))(f, g))
            g.check_overflow = g.check_overflow_num = lambda *a: None
            for call in range(3):
                g.inject_particles_user()
                out[f"{key}_r{rank}_counts{call}"] = np.array([g.ions, g.lecs], np.int64)
            out[f"{key}_r{rank}_p"] = p.copy()
            print("injector", key, "rank", rank, "ions", g.ions, "lecs", g.lecs, "x range", (float(p["x"][:g.ions].min()), float(p["x"][:g.ions].max())) if g.ions else None)
    np.savez_compressed(os.path.join(OUT, "ref_injector.npz"), **out)


# ------------------------------------------------------------------------------------------------------------
# G18: the domain decomposition -- the block of allocate_fields (fields.F90:246-330) that splits the box: local sizes with
#      the remainder on the last rank of an axis, MPI_Allgather of the sizes, cumulative offsets, strides.  The block is cut
#      out of the routine (which otherwise allocates and parses input) and wrapped in `subroutine decomp`; ranks are threads.
# ------------------------------------------------------------------------------------------------------------
DECOMP_CASES = [(2, 1, (23, 17, 1), (3, 2, 1)), (2, 2, (40, 9, 1), (4, 1, 1)), (3, 2, (10, 14, 9), (1, 3, 2)), (3, 1, (8, 12, 12), (1, 2, 2)),
                (3, 3, (6, 7, 23), (1, 1, 4)), (2, 0, (16, 16, 1), (1, 1, 1))]


def gen_decomp():
    import re
    out = {}
    text = src("fields.F90")
    a = re.search(r"^[ \t]*if \(irestart \.ne\. 1\) then\s*$", text, flags=re.M).start()
    b = text.index("#ifdef filter2", a)
    body = text[a:b]
    assert "mxcum=sum(mxl(i1:i2)-nghost)-(mxl(i2)-nghost)" in body and "lot=iz*mz" in body
    wrapped = "subroutine decomp()\n\timplicit none\n\tinteger :: i1, i2, j1, j2, k1, k2, error\n" + body + "\nend subroutine decomp\n"
    gi = GINTS | {"mxcum", "mycum", "mzcum", "mx0", "my0", "mz0", "irestart"}
    for ci, (dim, order, nglob, sizes) in enumerate(DECOMP_CASES):
        defines = {"MPI"} | ({"twoD"} if dim == 2 else set())
        sub = R.Sub(wrapped, "decomp", defines=defines, global_arrays={"mxl", "myl", "mzl"}, global_ints=gi).compile()
        size0 = sizes[0] * sizes[1] * sizes[2]
        comm = R.Comm(size0) if size0 > 1 else None
        ng = 5 if order <= 1 else 7
        ngz = ng if dim == 3 else 5
        gs = []
        for rank in range(size0):
            g = R.Globals(rank=rank, size0=size0, sizex=sizes[0], sizey=sizes[1], sizez=sizes[2], nghost=ng, nghostz=ngz, irestart=0,
                          mx0=nglob[0] + ng, my0=nglob[1] + ng, mz0=(nglob[2] + ngz if dim == 3 else 1),
                          mx=0, my=0, mz=1, mxcum=0, mycum=0, mzcum=0, mpi_integer=0, mpi_comm_world=0,
                          mxl=R.FArr((size0,), np.int64), myl=R.FArr((size0,), np.int64), mzl=R.FArr((size0,), np.int64))
            g.comm = comm
            gs.append(g)
        R.run_ranks([(lambda g=g: sub(g)) for g in gs])
        key = f"c{ci}"
        out[key + "_meta"] = np.array([dim, order, *nglob, *sizes], np.int32)
        out[key + "_ranks"] = np.array([[g.mx, g.my, g.mz, g.mxcum, g.mycum, g.mzcum, g.ix, g.iy, g.iz, g.lot] for g in gs], np.int64)
        print("decomp", key, dim, order, nglob, sizes, out[key + "_ranks"][:, :6].tolist())
    np.savez_compressed(os.path.join(OUT, "ref_decomp.npz"), **out)


# ------------------------------------------------------------------------------------------------------------
# G19: who talks to whom -- the neighbour ranks the reference's layer-copy routines address (copy_layrx1_opt /
#      copy_layry1_opt / copy_layrz1_opt, fieldboundaries.F90:1149-1677): each routine is run on every rank of a grid with
#      an MPI_SendRecv that only records (dest, source) of the first, "plus"-direction message
# ------------------------------------------------------------------------------------------------------------
def gen_neighbours():
    out = {}
    fb = src("fieldboundaries.F90")

    class Spy(R.Globals):
        def mpi_xchg(self, sendbuf, count, dest, sendtag, source, recvtag, sendtype=None, recvtype=None):
            self.seen.append((int(dest), int(source)))
            return R.as_payload(sendbuf, count)

    for ci, (dim, sizes) in enumerate([(2, (3, 2, 1)), (2, (4, 1, 1)), (3, (1, 3, 2)), (3, (1, 2, 4)), (3, (1, 1, 1))]):
        defines = {"MPI"} | ({"twoD"} if dim == 2 else set())
        subs = [R.Sub(fb, nm, defines=defines, global_arrays=GARR, global_ints=GINTS | {"statsize"}).compile()
                for nm in ("copy_layrx1_opt", "copy_layry1_opt", "copy_layrz1_opt")]
        size0 = sizes[0] * sizes[1] * sizes[2]
        table = np.zeros((size0, 6), np.int32)
        for rank in range(size0):
            g = Spy(rank=rank, size0=size0, sizex=sizes[0], sizey=sizes[1], sizez=sizes[2], statsize=5, mpi_comm_world=0, mpi_read=0)
            a = [R.FArr((4, 4, 4)) for _ in range(3)]
            for ax, f in enumerate(subs):
                g.seen = []
                f(g, *a, 4, 4, 4, 1, 3, 4, 2)
                if not g.seen:                        # 2D build: copy_layrz1_opt has no exchange
                    table[rank, 2 * ax], table[rank, 2 * ax + 1] = -1, -1
                    continue
                plus, minus = g.seen[0]
                table[rank, 2 * ax], table[rank, 2 * ax + 1] = minus, plus          # direction order: x-, x+, y-, y+, z-, z+
        out[f"n{ci}_sizes"] = np.array([dim, *sizes], np.int32)
        out[f"n{ci}_table"] = table
        print("neighbours", dim, sizes, table.tolist()[:4])
    np.savez_compressed(os.path.join(OUT, "ref_neighbours.npz"), **out)


# ------------------------------------------------------------------------------------------------------------
# G20: the scalars the kernels need -- read_input_particles (particles.F90:189-256): gamma0 from a velocity, the charge
#      normalisation qe, qi, the rescaled masses, qme, qmi, and the fp32 constants (3/2., 9/8., 2/3., 1/6. ...).  The input
#      parser calls are stand-ins that hand back the case's values.
# ------------------------------------------------------------------------------------------------------------
def gen_scalars():
    out = {}
    sub = R.Sub(src("particles.F90"), "read_input_particles", defines={"MPI"}, global_ints=GINTS | {"upsamp_e", "upsamp_i", "maxptl0"}).compile()
    cases = [dict(ppc0=16., c_omp=10., gamma0=.5, me=1., mi=1., sigma=0.), dict(ppc0=64., c_omp=10., gamma0=.5, me=1., mi=1., sigma=0.),
             dict(ppc0=4., c_omp=8., gamma0=15., me=1., mi=100., sigma=0.1), dict(ppc0=2., c_omp=5., gamma0=.1, me=1., mi=20., sigma=0.)]
    names = ("qe", "qi", "qme", "qmi", "me", "mi", "gamma0", "beta", "binit", "three", "two", "thhalf", "nineighth", "one", "threeq", "twoth",
             "half", "third", "quart", "sixth", "negsixth", "negone")
    out["names"] = np.array(names)
    for ci, vals in enumerate(cases):
        g = R.Globals(c=F(0.45), sigma=F(0), ppc0=F(0), delgam=F(0), me=F(0), mi=F(0), gamma0=F(0), c_omp=F(0), upsamp_e=0, upsamp_i=0)

        def getd(sec, name, default, var, vals=vals):
            return (sec, name, default, F(vals.get(name, default)))
        g.inputpar_getd_def = getd
        g.inputpar_geti_def = lambda sec, name, default, var: (sec, name, default, default)
        sub(g)
        out[f"k{ci}_in"] = np.array([vals[k] for k in ("ppc0", "c_omp", "gamma0", "me", "mi", "sigma")], F)
        out[f"k{ci}_out"] = np.array([getattr(g, nm) for nm in names], F)
        print("scalars", ci, dict(zip(names[:7], out[f"k{ci}_out"][:7])))
    np.savez_compressed(os.path.join(OUT, "ref_scalars.npz"), **out)


# ------------------------------------------------------------------------------------------------------------
# G21: the order of the calls in one lap of `mainloop` (tristanmainloop.F90:107-330), read from the text for the 2D build
#      and for the 3D filter2 build: what host/tristan_mainloop.cpp and INTEGRATION.md section 4 must reproduce
# ------------------------------------------------------------------------------------------------------------
def gen_calllist():
    out = {}
    text = src("tristanmainloop.F90")
    for tag, defines in (("2d", {"MPI", "twoD", "dd1"}), ("3d", {"MPI", "dd2"}), ("3d_filter2", {"MPI", "dd2", "filter2"})):
        st = R.statements(R.extract_subroutine(R.preprocess(text, defines), "mainloop"))
        i0 = next(i for i, s_ in enumerate(st) if s_.startswith("do lap"))
        calls = []
        for s_ in st[i0:]:
            m = re.search(r"\bcall\s+([a-z_]\w*)", s_)
            if m and m.group(1) not in ("timer", "mpi_barrier"):
                calls.append(m.group(1))
        out[tag] = np.array(calls)
        print("calllist", tag, len(calls), calls[:8], "...")
    np.savez_compressed(os.path.join(OUT, "ref_calllist.npz"), **out)


# ------------------------------------------------------------------------------------------------------------
# G22: the restart dump -- the two unformatted WRITE statements of output.F90:2194-2222 (restflds.*, restprtl.*), executed
#      from the text; the kind of every item comes from its declaration in the module headers (fields.F90, particles.F90,
#      aux.F90).  The byte streams are the golden; the package's restart.py must write exactly these and read them back.
# ------------------------------------------------------------------------------------------------------------
def gen_restart():
    out = {}
    text = R.preprocess(src("output.F90"), {"MPI"})            # LOCALRESTART (an older dump without walloc, :2112-2141) is not defined
    w7 = re.search(r"^[ \t]*write\(7\)mx,my,mz.*?(?=^[ \t]*close\(7\))", text, flags=re.M | re.S).group(0)
    w8 = re.search(r"^[ \t]*write\(8\)ions,lecs.*?(?=^[ \t]*close\(8\))", text, flags=re.M | re.S).group(0)
    wrapped = "subroutine dump()\n\timplicit none\n\tinteger :: n\n" + w7 + "\n" + w8 + "\nend subroutine dump\n"
    heads = {f: R.preprocess(src(f), {"MPI"}) for f in ("fields.F90", "particles.F90", "aux.F90")}

    def kind_of(name):
        for t in heads.values():
            try:
                return R.module_kind(t, name)
            except KeyError:
                pass
        raise KeyError(name)
    cast = {"int": int, "real": F, "real8": np.float64}
    rng = np.random.default_rng(1700)
    for ci, (dim, n) in enumerate([(2, (6, 5, 1)), (3, (5, 4, 3))]):
        ng, ngz, mx, my, mz = grid(dim, 2, n)
        vals = dict(mx=mx, my=my, mz=mz, dseed=123457.0 * 16807 % 2147483647, lap=1234 + ci, xinject=3.25, xinject2=mx - 2.5, xinject3=0.125,
                    leftwall=15.5, walloc=20.75, ions=7 + ci, lecs=5, maxptl=40, maxhlf=20, totalpartnum=99 + ci)
        g = R.Globals(**{k: cast[kind_of(k)](v) for k, v in vals.items()})
        sub = R.Sub(wrapped, "dump", defines={"MPI"}, global_arrays=GARR, global_ints={k for k in vals if kind_of(k) == "int"}).compile()
        for nm in ("ex", "ey", "ez", "bx", "by", "bz"):
            a = R.FArr((mx, my, mz))
            a.flat[:] = rng.standard_normal(a.flat.size).astype(F)
            setattr(g, nm, a)
        p = np.zeros(vals["maxptl"], PDT)
        for k in PDT.names:
            p[k] = (rng.standard_normal(p.size) * 3).astype(PDT[k]) if PDT[k].kind == "f" else rng.integers(-50, 50, p.size)
        g.p = R.RecArr(p)
        sub(g)
        key = f"t{ci}"
        out[key + "_kinds"] = np.array([f"{k}:{kind_of(k)}" for k in vals])
        out[key + "_scalars"] = np.array([float(v) for v in vals.values()], np.float64)
        out[key + "_names"] = np.array(list(vals))
        for a, nm in enumerate(("ex", "ey", "ez", "bx", "by", "bz")):
            out[f"{key}_f{a}"] = c_order(getattr(g, nm))
        out[key + "_p"] = p.copy()
        out[key + "_restflds"] = np.frombuffer(b"".join(g.units[7]), np.uint8)
        out[key + "_restprtl"] = np.frombuffer(b"".join(g.units[8]), np.uint8)
        print("restart", key, "restflds", out[key + "_restflds"].size, "bytes, restprtl", out[key + "_restprtl"].size, "bytes;", list(out[key + "_kinds"]))
    np.savez_compressed(os.path.join(OUT, "ref_restart.npz"), **out)


# ------------------------------------------------------------------------------------------------------------
# G23: which particles go into prtl.tot -- the index-building block of save_param / particle output (output.F90:3331-3393):
#      modulo(p(n)%ind/2, stride) == 0 with Fortran's truncating division and floor modulo, ions then electrons.  The block
#      is cut out of the HDF5 writer and wrapped in `subroutine pick`.
# ------------------------------------------------------------------------------------------------------------
def gen_select():
    out = {}
    text = R.preprocess(src("output.F90"), {"MPI"})
    a = text.index("     strd_ions = 0\n     tmp1 = ions")
    b = text.index("     all_ions_lng=all_ions", a)
    wrapped = ("subroutine pick()\n\timplicit none\n\tinteger :: n, tmp1, error\n\tinteger :: all_strd_ions(size0), all_strd_lecs(size0), all_ions(size0), "
               "all_lecs(size0)\n" + text[a:b] + "\nend subroutine pick\n")
    gi = GINTS | {"stride", "strd_ions", "strd_lecs", "ions_str_ind", "lecs_str_ind"}
    sub = R.Sub(wrapped, "pick", defines={"MPI"}, global_arrays={"p", "ions_str_ind", "lecs_str_ind"}, global_ints=gi).compile()
    rng = np.random.default_rng(1800)
    maxhlf, ions, lecs = 64, 50, 41
    p = np.zeros(2 * maxhlf, PDT)
    p["ind"] = rng.integers(-200, 200, p.size)
    out["p_ind"] = p["ind"].copy()
    out["geom"] = np.array([maxhlf, ions, lecs], np.int32)
    for stride in (1, 2, 3, 7, 20):
        g = R.Globals(p=R.RecArr(p), ions=ions, lecs=lecs, maxhlf=maxhlf, stride=stride, size0=1, rank=0, debug=False, mpi_integer=0,
                      mpi_comm_world=0, strd_ions=0, strd_lecs=0, ions_str_ind=None, lecs_str_ind=None)
        sub(g)
        out[f"s{stride}_ions"] = np.array(g.ions_str_ind.flat[:g.strd_ions], np.int64)
        out[f"s{stride}_lecs"] = np.array(g.lecs_str_ind.flat[:g.strd_lecs], np.int64)
        print("select stride", stride, g.strd_ions, g.strd_lecs)
    np.savez_compressed(os.path.join(OUT, "ref_select.npz"), **out)


if __name__ == "__main__":
    which = sys.argv[1:] or ["deposit", "fields", "mover", "filter", "radiation", "halo", "fields42", "shock", "depositp", "halo_mr", "migrate_mr", "lap", "filter2_mr", "meanq", "loader", "spectrum", "injector", "decomp", "neighbours", "scalars", "calllist", "restart", "select"]
    for w in which:
        globals()["gen_" + w]()
