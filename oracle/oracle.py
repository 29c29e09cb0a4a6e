"""ctypes binding of the CPU oracle (oracle/pic_oracle.c).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
PARITY: pinned against the reference's source text (tests/test_ref_golden.py); see the header of pic_oracle.h.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libpic_oracle.so")

EX, EY, EZ, BX, BY, BZ, CURX, CURY, CURZ = range(9)
ARR_NAMES = ["ex", "ey", "ez", "bx", "by", "bz", "curx", "cury", "curz"]
(PH_BC_B1, PH_BC_E1, PH_BHALF, PH_MOVE, PH_EFULL, PH_RESET, PH_DEPOSIT, PH_EXCH_P, PH_EXCH_CUR,
 PH_FILTER, PH_ADD_CUR, PH_INJECT_OTHERS, PH_REORDER) = range(13)
PH_SURF_B, PH_SURF_E = 100, 101   # the `surface` part of bc_b2 / bc_e2 (radiating axes)
PH_PRE_B, PH_POST_B, PH_PRE_E, PH_POST_E = 102, 103, 104, 105   # preledge / postedge groups of an all-open 3D box
Q_REFERENCE = 0xF

PARTICLE_DTYPE = np.dtype([("x", "f4"), ("y", "f4"), ("z", "f4"), ("u", "f4"), ("v", "f4"), ("w", "f4"),
                           ("ch", "f4"), ("ind", "i4"), ("proc", "i4"), ("splitlev", "i4")])
assert PARTICLE_DTYPE.itemsize == 40


class Params(C.Structure):
    _fields_ = [("dim", C.c_int), ("order", C.c_int),
                ("mx0", C.c_int), ("my0", C.c_int), ("mz0", C.c_int),
                ("sizex", C.c_int), ("sizey", C.c_int), ("sizez", C.c_int),
                ("c", C.c_float), ("corr", C.c_float),
                ("ntimes", C.c_int), ("filter_kind", C.c_int),
                ("periodicx", C.c_int), ("periodicy", C.c_int), ("periodicz", C.c_int),
                ("qi", C.c_float), ("qe", C.c_float), ("qmi", C.c_float), ("qme", C.c_float),
                ("maxptl", C.c_int), ("buffsize", C.c_int), ("quirks", C.c_int), ("pusher", C.c_int),
                ("external_fields", C.c_int), ("ext", C.c_float * 6),
                ("highorder", C.c_int), ("wall_i2", C.c_int)]


def build(force=False):
    src = os.path.join(_HERE, "pic_oracle.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB


_lib = None
_which = _LIB


def use_timing_build():
    """bench.py only: load the -O3 -march=native build (libpic_oracle_fast.so) instead of the bit-faithful -O2
    -ffp-contract=off one.  Must be called before the first use of the library in this process."""
    global _which
    if _lib is not None:
        raise RuntimeError("oracle library already loaded")
    # -march=native code must be compiled on the machine that runs it (the repo snapshot travels to the GPU box with its
    # built artefacts): one build per CPU model, keyed by a hash of the model name and flags
    import hashlib
    try:
        info = [l for l in open("/proc/cpuinfo") if l.startswith(("model name", "flags"))][:2]
    except OSError:
        info = []
    tag = hashlib.sha1("".join(info).encode()).hexdigest()[:10]
    fast = os.path.join(_HERE, f"libpic_oracle_fast_{tag}.so")
    src = os.path.join(_HERE, "pic_oracle.c")
    if not os.path.exists(fast) or os.path.getmtime(fast) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "libpic_oracle_fast.so"])
        os.replace(os.path.join(_HERE, "libpic_oracle_fast.so"), fast)
    _which = fast


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_which)
        vp, ci, cf = C.c_void_p, C.c_int, C.c_float
        L.orc_world_create.restype = vp
        L.orc_world_create.argtypes = [C.POINTER(Params)]
        L.orc_world_destroy.argtypes = [vp]
        L.orc_world_rank.restype = vp
        L.orc_world_rank.argtypes = [vp, ci]
        L.orc_rank_array.restype = C.POINTER(C.c_float)
        L.orc_rank_array.argtypes = [vp, ci]
        L.orc_rank_particles.restype = vp
        L.orc_rank_particles.argtypes = [vp]
        L.orc_rank_dims.argtypes = [vp, C.POINTER(ci)]
        L.orc_rank_counts.argtypes = [vp, C.POINTER(ci), C.POINTER(ci)]
        L.orc_rank_set_counts.argtypes = [vp, ci, ci]
        for name in ["orc_advance_b_halfstep", "orc_advance_e_fullstep", "orc_reset_currents",
                     "orc_add_current", "orc_move_particles", "orc_deposit_currents_only",
                     "orc_deposit_particles", "orc_inject_others", "orc_reorder_particles",
                     "orc_filter1_pass", "orc_surface_b", "orc_surface_e"]:
            getattr(L, name).argtypes = [vp]
        L.orc_edges.argtypes = [vp, ci]
        L.orc_deposit_one.argtypes = [vp] + [cf] * 7
        L.orc_mover_range.argtypes = [vp, ci, ci, cf]
        L.orc_bc_fields.argtypes = [vp, ci]
        for name in ["orc_exchange_current", "orc_exchange_particles", "orc_apply_filter1",
                     "orc_apply_filter2", "orc_apply_filter", "orc_step"]:
            getattr(L, name).argtypes = [vp]
        L.orc_step_phase.argtypes = [vp, ci]
        L.orc_meanq_fld_cur.argtypes = [vp, C.c_char_p]
        L.orc_spectrum_gamma_range.argtypes = [vp, C.POINTER(cf), C.POINTER(cf)]
        L.orc_spectrum.argtypes = [vp, cf, cf, ci, cf, ci, ci] + [C.POINTER(cf)] * 4
        L.orc_shape.argtypes = [ci, cf, ci, C.POINTER(cf), C.POINTER(ci), C.POINTER(ci)]
        L.orc_filter2_line.argtypes = [C.POINTER(cf), ci, ci]
        L.orc_filter2_rank.argtypes = [vp, ci, ci, C.POINTER(cf), C.POINTER(cf)]
        L.orc_filter2_send_box.argtypes = [vp, ci, ci, C.POINTER(ci), C.POINTER(ci)]
        L.orc_rank_box.restype = vp
        L.orc_rank_box.argtypes = [vp, ci, ci, C.POINTER(ci), C.POINTER(ci)]
        L.orc_rank_box_set_counts.argtypes = [vp, ci, ci, ci, ci]
        L.orc_neighbour.restype = ci
        L.orc_neighbour.argtypes = [vp, ci]
        i3 = C.POINTER(ci)
        for name in ["orc_box_get", "orc_box_put", "orc_box_add"]:
            getattr(L, name).argtypes = [vp, ci, i3, i3, C.POINTER(cf)]
        L.orc_random.restype = cf
        L.orc_random.argtypes = [C.POINTER(C.c_double)]
        L.orc_charge_normalisation.argtypes = [C.POINTER(Params), cf, cf, cf, cf, cf]
        L.orc_init_weibel.argtypes = [vp, cf, cf, cf, cf, cf, cf, ci]
        L.orc_init_twostream.argtypes = [vp, cf, cf, cf, cf, cf, cf]
        L.orc_init_uniform.argtypes = [vp, cf, cf, cf, C.c_uint64]
        L.orc_inject_particles_shock.argtypes = [vp, cf, cf, cf, cf, cf, cf, ci, cf]
        L.orc_field_bc_shock.argtypes = [vp, cf, cf, cf, cf, cf]
        L.orc_particle_bc_wall.argtypes = [vp, cf]
        L.orc_step_shock.argtypes = [vp, cf, cf, cf, cf, cf]
        L.orc_charge_density.argtypes = [vp, C.POINTER(cf)]
        L.orc_sum_array.restype = C.c_double
        L.orc_sum_array.argtypes = [vp, ci]
        _lib = L
    return _lib


def make_params(dim=2, order=1, mx0=32, my0=32, mz0=1, sizex=1, sizey=1, sizez=1, c=0.45, corr=1.025,
                ntimes=0, filter_kind=1, periodic=(1, 1, 1), ppc0=16.0, c_omp=10.0, gamma0=0.5, me=1.0,
                mi=1.0, maxptl=None, buffsize=None, quirks=Q_REFERENCE, pusher=0, ext=None, highorder=0, wall_i2=0):
    P = Params()
    P.highorder, P.wall_i2 = highorder, wall_i2
    P.dim, P.order = dim, order
    P.mx0, P.my0, P.mz0 = mx0, my0, (mz0 if dim == 3 else 1)
    P.sizex, P.sizey, P.sizez = sizex, sizey, (sizez if dim == 3 else 1)
    P.c, P.corr, P.ntimes, P.filter_kind = c, corr, ntimes, filter_kind
    P.periodicx, P.periodicy, P.periodicz = periodic
    ncell = mx0 * my0 * (mz0 if dim == 3 else 1)
    nrank = P.sizex * P.sizey * P.sizez
    if maxptl is None:
        maxptl = int(2.5 * ppc0 * ncell / nrank) + 4096
    P.maxptl = maxptl
    P.buffsize = buffsize if buffsize is not None else max(maxptl // 4, 10000)
    P.quirks, P.pusher = quirks, pusher
    P.external_fields = 0 if ext is None else 1
    for i in range(6):
        P.ext[i] = 0.0 if ext is None else ext[i]
    lib().orc_charge_normalisation(C.byref(P), ppc0, c_omp, gamma0, me, mi)
    return P


class Rank:
    def __init__(self, world, idx):
        self.world, self.idx = world, idx
        self.h = lib().orc_world_rank(world.h, idx)
        d = (C.c_int * 9)()
        lib().orc_rank_dims(self.h, d)
        (self.mx, self.my, self.mz, self.nghost, self.nghostz, self.mxcum, self.mycum, self.mzcum,
         self.maxhlf) = list(d)
        self.maxptl = world.P.maxptl

    def arr(self, which):
        """numpy view shaped (mz,my,mx) (C order) == Fortran (mx,my,mz)."""
        p = lib().orc_rank_array(self.h, which)
        return np.ctypeslib.as_array(p, shape=(self.mz, self.my, self.mx))

    def fields(self):
        return [self.arr(i) for i in range(6)]

    def currents(self):
        return [self.arr(i) for i in range(6, 9)]

    def particles(self):
        """structured view of the whole AoS p(:) array (ions at [0,ions), electrons at [maxhlf, ...))."""
        addr = lib().orc_rank_particles(self.h)
        buf = (C.c_char * (self.maxptl * 40)).from_address(addr)
        return np.frombuffer(buf, dtype=PARTICLE_DTYPE)

    @property
    def counts(self):
        a, b = C.c_int(), C.c_int()
        lib().orc_rank_counts(self.h, C.byref(a), C.byref(b))
        return a.value, b.value

    def set_counts(self, ions, lecs):
        lib().orc_rank_set_counts(self.h, ions, lecs)

    def ions(self):
        return self.particles()[: self.counts[0]]

    def lecs(self):
        return self.particles()[self.maxhlf: self.maxhlf + self.counts[1]]

    def charge_density(self):
        rho = np.zeros((self.mz, self.my, self.mx), np.float32)
        lib().orc_charge_density(self.h, rho.ctypes.data_as(C.POINTER(C.c_float)))
        return rho

    def call(self, name, *a):
        return getattr(lib(), "orc_" + name)(self.h, *a)

    def spectrum(self, mx0, splitratio=10.0, gambins=200, gamma_range=None):
        """per-rank part of save_spectrum (output.F90:380-633) -> (gammin, gammax, specp, spece, specprest, specerest),
        arrays shaped (gambins, nbins) (C order == Fortran (nbins, gambins))"""
        cf = C.c_float
        lo, hi = cf(), cf()
        lib().orc_spectrum_gamma_range(self.h, C.byref(lo), C.byref(hi))
        glo, ghi = (lo.value, hi.value) if gamma_range is None else gamma_range      # the allreduced range
        nbins = max((mx0 - 2 - 3) // 100, 1)
        out = [np.zeros((gambins, nbins), np.float32) for _ in range(4)]
        lib().orc_spectrum(self.h, glo, ghi, mx0, splitratio, nbins, gambins, *[a.ctypes.data_as(C.POINTER(cf)) for a in out])
        return (lo.value, hi.value, *out)


class World:
    def __init__(self, P):
        self.P = P
        self.h = lib().orc_world_create(C.byref(P))
        if not self.h:
            raise ValueError("unsupported configuration (highorder = 1 with open y on a split axis or open z)")
        self.n = P.sizex * P.sizey * P.sizez
        self.ranks = [Rank(self, i) for i in range(self.n)]

    def __del__(self):
        try:
            if self.h:
                lib().orc_world_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def call(self, name, *a):
        return getattr(lib(), "orc_" + name)(self.h, *a)

    def phase(self, ph):
        lib().orc_step_phase(self.h, ph)

    def step(self):
        lib().orc_step(self.h)

    def meanq_fld_cur(self, totname):
        """output.F90:5229-5486: moment `totname` ('tdens', 'ibetx', ...) into curx; cury holds the weight"""
        lib().orc_meanq_fld_cur(self.h, totname.encode())

    def init_weibel(self, ppc0=16.0, gamma0=0.5, delgam=2e-5, me=1.0, mi=1.0, tratio=1.0, distr_dim=2):
        lib().orc_init_weibel(self.h, ppc0, gamma0, delgam, me, mi, tratio, distr_dim)

    def inject_particles_shock(self, ppc0=16.0, gamma0=0.5, delgam=2e-5, me=1.0, mi=1.0, tratio=1.0, pcosthmult=0, sigma=0.0):
        """the shock problem's per-lap plane source on every rank (user/user_shock.F90:303-331 -> inject_from_wall)"""
        lib().orc_inject_particles_shock(self.h, ppc0, gamma0, delgam, me, mi, tratio, pcosthmult, sigma)

    def init_twostream(self, ppc0=64.0, gamma0=0.5, delgam=2e-5, me=1.0, mi=1.0, tratio=1.0):
        lib().orc_init_twostream(self.h, ppc0, gamma0, delgam, me, mi, tratio)

    def init_uniform(self, ppc0=16.0, beta=0.5, uth=0.05, seed=1):
        lib().orc_init_uniform(self.h, ppc0, beta, uth, seed)


def shape(order, d, shift=0):
    S = (C.c_float * 8)()
    a, b = C.c_int(), C.c_int()
    lib().orc_shape(order, d, shift, S, C.byref(a), C.byref(b))
    return np.array(S[:], np.float32), a.value, b.value
