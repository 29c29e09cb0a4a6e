"""Host-side mirror of the reference's hot-path call surface, over the C ABI of libtristan_gpu.so.

The reference (TRISTAN-MP, Fortran) exposes this path as argument-less module procedures called by
``mainloop`` (code/tristanmainloop.F90:107-344).  :class:`Context` carries one method per procedure with
the same name (``move_particles``, ``deposit_particles``, ``advance_b_halfstep`` ...), each a thin ctypes
call into ``include/tristan_gpu.h``.  There is no CPU fallback: importing works anywhere (so the ABI can be
inspected), creating a :class:`Context` raises unless the CUDA library and a B200-class GPU are present.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# TGPU_LIB selects an A/B build of the same library (csrc/Makefile: BUILD=... OUT=... CRFLAGS=...)
LIB_PATH = os.environ.get("TGPU_LIB") or os.path.join(_HERE, "libtristan_gpu.so")

PARTICLE_DTYPE = np.dtype([("x", "f4"), ("y", "f4"), ("z", "f4"), ("u", "f4"), ("v", "f4"), ("w", "f4"),
                           ("ch", "f4"), ("ind", "i4"), ("proc", "i4"), ("splitlev", "i4")])
Q_REFERENCE = 0xF
PHASES = ["fields", "mover", "deposit", "part_exch", "cur_exch", "filter", "sort", "bc"]


class TristanGPUError(RuntimeError):
    pass


class Params(C.Structure):
    """struct tgpu_params (include/tristan_gpu.h)."""
    _fields_ = [("dim", C.c_int32), ("order", C.c_int32),
                ("mx", C.c_int32), ("my", C.c_int32), ("mz", C.c_int32),
                ("nghost", C.c_int32), ("nghostz", C.c_int32),
                ("c", C.c_float), ("corr", C.c_float),
                ("qi", C.c_float), ("qe", C.c_float), ("qmi", C.c_float), ("qme", C.c_float),
                ("ntimes", C.c_int32), ("filter_kind", C.c_int32),
                ("periodicx", C.c_int32), ("periodicy", C.c_int32), ("periodicz", C.c_int32),
                ("x1in", C.c_float), ("x2in", C.c_float), ("y1in", C.c_float), ("y2in", C.c_float),
                ("z1in", C.c_float), ("z2in", C.c_float),
                ("rank", C.c_int32), ("sizex", C.c_int32), ("sizey", C.c_int32), ("sizez", C.c_int32),
                ("mxcum", C.c_int32), ("mycum", C.c_int32), ("mzcum", C.c_int32),
                ("mxl", C.POINTER(C.c_int32)), ("myl", C.POINTER(C.c_int32)), ("mzl", C.POINTER(C.c_int32)),
                ("maxptl", C.c_int32), ("buffsize", C.c_int32), ("quirks", C.c_int32), ("pusher", C.c_int32),
                ("external_fields", C.c_int32), ("ext", C.c_float * 6),
                ("device", C.c_int32),
                ("highorder", C.c_int32), ("wall_i2", C.c_int32)]


ABI_SYMBOLS = [
    "tgpu_init", "tgpu_finalize", "tgpu_last_error", "tgpu_device_count", "tgpu_neighbour", "tgpu_decompose",
    "tgpu_ghost_width", "tgpu_comm_unique_id", "tgpu_comm_init", "tgpu_fields_h2d", "tgpu_fields_d2h",
    "tgpu_currents_h2d", "tgpu_currents_d2h", "tgpu_particles_h2d", "tgpu_particles_d2h", "tgpu_counts",
    "tgpu_append_particles", "tgpu_advance_b_halfstep", "tgpu_advance_e_fullstep", "tgpu_reset_currents",
    "tgpu_add_current", "tgpu_bc_b1", "tgpu_bc_e1", "tgpu_bc_b2", "tgpu_bc_e2", "tgpu_exchange_current",
    "tgpu_apply_filter", "tgpu_apply_filter1_opt", "tgpu_apply_filter2_opt", "tgpu_move_particles",
    "tgpu_deposit_particles", "tgpu_exchange_particles", "tgpu_inject_others", "tgpu_reorder_particles",
    "tgpu_meanq_fld_cur", "tgpu_spectrum_gamma_range", "tgpu_spectrum", "tgpu_select_particles", "tgpu_step_mirror", "tgpu_field_bc_user_shock", "tgpu_particle_bc_user_wall", "tgpu_set_user_hooks", "tgpu_step", "tgpu_timers", "tgpu_launch_count", "tgpu_stream", "tgpu_set_option", "tgpu_halo_transport", "tgpu_pre_bc_b", "tgpu_post_bc_b", "tgpu_pre_bc_e", "tgpu_post_bc_e",
]

_lib = None


def load_library(path=None):
    """dlopen libtristan_gpu.so; raises if it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise TristanGPUError(f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(nvcc, sm_100a). There is no CPU fallback.")
    L = C.CDLL(path)
    vp, ci = C.c_void_p, C.c_int
    fp = C.POINTER(C.c_float)
    L.tgpu_init.argtypes = [C.POINTER(Params), C.POINTER(vp)]
    L.tgpu_finalize.argtypes = [vp]
    L.tgpu_last_error.restype = C.c_char_p
    L.tgpu_neighbour.argtypes = [ci] * 5
    L.tgpu_decompose.argtypes = [ci] * 9 + [C.POINTER(C.c_int32)]
    L.tgpu_ghost_width.argtypes = [ci, ci, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.tgpu_comm_unique_id.argtypes = [C.POINTER(C.c_uint8)]
    L.tgpu_comm_init.argtypes = [vp, C.POINTER(C.c_uint8)]
    L.tgpu_fields_h2d.argtypes = [vp] + [fp] * 6
    L.tgpu_fields_d2h.argtypes = [vp] + [fp] * 6
    L.tgpu_currents_h2d.argtypes = [vp] + [fp] * 3
    L.tgpu_currents_d2h.argtypes = [vp] + [fp] * 3
    L.tgpu_particles_h2d.argtypes = [vp, vp, ci, ci]
    L.tgpu_particles_d2h.argtypes = [vp, vp, C.POINTER(ci), C.POINTER(ci)]
    L.tgpu_counts.argtypes = [vp, C.POINTER(ci), C.POINTER(ci)]
    L.tgpu_append_particles.argtypes = [vp, vp, ci, ci]
    for name in ["tgpu_advance_b_halfstep", "tgpu_advance_e_fullstep", "tgpu_reset_currents", "tgpu_add_current",
                 "tgpu_bc_b1", "tgpu_bc_e1", "tgpu_bc_b2", "tgpu_bc_e2", "tgpu_exchange_current", "tgpu_apply_filter",
                 "tgpu_apply_filter1_opt", "tgpu_apply_filter2_opt", "tgpu_move_particles", "tgpu_deposit_particles",
                 "tgpu_exchange_particles", "tgpu_inject_others", "tgpu_reorder_particles",
                 "tgpu_pre_bc_b", "tgpu_post_bc_b", "tgpu_pre_bc_e", "tgpu_post_bc_e"]:
        getattr(L, name).argtypes = [vp]
    L.tgpu_field_bc_user_shock.argtypes = [vp] + [C.c_float] * 5
    L.tgpu_meanq_fld_cur.argtypes = [vp, C.c_char_p]
    L.tgpu_spectrum_gamma_range.argtypes = [vp, fp, fp]
    L.tgpu_spectrum.argtypes = [vp, C.c_float, C.c_float, ci, C.c_float, ci, ci] + [fp] * 4
    L.tgpu_select_particles.argtypes = [vp, ci, vp, ci, C.POINTER(ci), C.POINTER(ci)]
    L.tgpu_step_mirror.argtypes = [vp] + [fp] * 6 + [vp, C.POINTER(ci), C.POINTER(ci)]
    L.tgpu_particle_bc_user_wall.argtypes = [vp, C.c_float]
    L.tgpu_set_user_hooks.argtypes = [vp, ci, C.POINTER(C.c_float)]
    L.tgpu_step.argtypes = [vp, ci]
    L.tgpu_timers.argtypes = [vp, C.POINTER(C.c_double), ci]
    L.tgpu_launch_count.restype = C.c_int64
    L.tgpu_launch_count.argtypes = [vp]
    L.tgpu_stream.restype = vp
    L.tgpu_stream.argtypes = [vp]
    L.tgpu_halo_transport.argtypes = [vp]
    L.tgpu_set_option.argtypes = [vp, C.c_char_p, ci]
    if path == LIB_PATH:
        _lib = L
    return L


def ghost_width(dim, order):
    a, b = C.c_int32(), C.c_int32()
    if load_library().tgpu_ghost_width(dim, order, C.byref(a), C.byref(b)):
        raise TristanGPUError("bad dim/order")
    return a.value, b.value


def decompose(dim, order, mx0, my0, mz0, sizex, sizey, sizez, rank):
    """(mx,my,mz,mxcum,mycum,mzcum) of `rank` -- code/fields.F90:259-328."""
    out = (C.c_int32 * 6)()
    if load_library().tgpu_decompose(dim, order, mx0, my0, mz0, sizex, sizey, sizez, rank, out):
        raise TristanGPUError("bad decomposition")
    return tuple(out)


def neighbour(rank, sizex, sizey, sizez, direction):
    return load_library().tgpu_neighbour(rank, sizex, sizey, sizez, direction)


def charge_normalisation(c, c_omp, ppc0, gamma0, me, mi):
    """qe, qi, qme, qmi as read_input_particles computes them (code/particles.F90:219-235), in fp32."""
    f = np.float32
    gamma0 = f(gamma0)
    if gamma0 < 1:
        gamma0 = f(np.sqrt(f(1.) / (f(1.) - gamma0 * gamma0)))
    omp = f(c) / f(c_omp)
    qe = -(omp * omp * gamma0) / ((f(ppc0) * f(.5)) * (f(1) + f(me) / f(mi)))
    qi = -qe
    me2, mi2 = f(me) * abs(qi), f(mi) * abs(qi)
    return float(qe), float(qi), float(qe / me2), float(qi / mi2)


def make_params(dim=3, order=2, mx0=32, my0=32, mz0=32, sizex=1, sizey=1, sizez=1, rank=0, c=0.45, corr=1.025,
                ntimes=32, filter_kind=1, periodic=(1, 1, 1), ppc0=16.0, c_omp=10.0, gamma0=0.5, me=1.0, mi=1.0,
                maxptl=None, buffsize=None, quirks=Q_REFERENCE, pusher=0, ext=None, device=-1, highorder=0, wall_i2=0):
    """Build a tgpu_params the way initialize() fills the reference's module globals for one rank."""
    if dim == 2:
        sizez, mz0 = 1, 1
    ng, ngz = ghost_width(dim, order)
    size0 = sizex * sizey * sizez
    P = Params()
    P.dim, P.order = dim, order
    geo = [decompose(dim, order, mx0, my0, mz0, sizex, sizey, sizez, r) for r in range(size0)]
    P.mx, P.my, P.mz, P.mxcum, P.mycum, P.mzcum = geo[rank]
    P.nghost, P.nghostz = ng, ngz
    P.c, P.corr, P.ntimes, P.filter_kind = c, corr, ntimes, filter_kind
    P.periodicx, P.periodicy, P.periodicz = periodic
    P.qe, P.qi, P.qme, P.qmi = charge_normalisation(c, c_omp, ppc0, gamma0, me, mi)
    # particles.F90:339-344
    P.x1in, P.x2in = ng // 2 + 1, mx0 + ng - ng // 2
    P.y1in, P.y2in = ng // 2 + 1, my0 + ng - ng // 2
    P.z1in, P.z2in = ngz // 2 + 1, mz0 + ngz - ngz // 2
    P.rank, P.sizex, P.sizey, P.sizez = rank, sizex, sizey, sizez
    keep = []
    for name, idx in (("mxl", 0), ("myl", 1), ("mzl", 2)):
        arr = (C.c_int32 * size0)(*[g[idx] for g in geo])
        keep.append(arr)
        setattr(P, name, C.cast(arr, C.POINTER(C.c_int32)))
    P._keep = keep
    ncell = mx0 * my0 * (mz0 if dim == 3 else 1)
    if maxptl is None:
        maxptl = int(2.5 * ppc0 * ncell / size0) + 4096
    P.maxptl = maxptl
    P.buffsize = buffsize if buffsize is not None else max(maxptl // 8, 10000)
    P.quirks, P.pusher = quirks, pusher
    P.external_fields = 0 if ext is None else 1
    for i in range(6):
        P.ext[i] = 0.0 if ext is None else ext[i]
    P.device = device
    P.highorder, P.wall_i2 = highorder, wall_i2
    return P


def _fptr(a):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_float))


class Context:
    """One rank's GPU-resident state; methods are named after the reference procedures they replace."""

    def __init__(self, params):
        self.lib = load_library()
        self.P = params
        h = C.c_void_p()
        rc = self.lib.tgpu_init(C.byref(params), C.byref(h))
        if rc:
            raise TristanGPUError(f"tgpu_init failed ({rc}): {self.lib.tgpu_last_error().decode()}")
        self.h = h
        self.shape = (params.mz, params.my, params.mx)   # C-order view of Fortran (mx,my,mz)
        self.maxptl = params.maxptl
        self.maxhlf = params.maxptl // 2

    def close(self):
        if getattr(self, "h", None):
            self.lib.tgpu_finalize(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        if rc:
            raise TristanGPUError(f"{what} failed ({rc}): {self.lib.tgpu_last_error().decode()}")

    # --- state transfer -----------------------------------------------------------------------------
    def fields_h2d(self, ex, ey, ez, bx, by, bz):
        self._ck(self.lib.tgpu_fields_h2d(self.h, *[_fptr(a) for a in (ex, ey, ez, bx, by, bz)]), "fields_h2d")

    def fields_d2h(self, out=None):
        out = out or [np.empty(self.shape, np.float32) for _ in range(6)]
        self._ck(self.lib.tgpu_fields_d2h(self.h, *[_fptr(a) for a in out]), "fields_d2h")
        return out

    def currents_h2d(self, cx, cy, cz):
        self._ck(self.lib.tgpu_currents_h2d(self.h, _fptr(cx), _fptr(cy), _fptr(cz)), "currents_h2d")

    def currents_d2h(self, out=None):
        out = out or [np.empty(self.shape, np.float32) for _ in range(3)]
        self._ck(self.lib.tgpu_currents_d2h(self.h, *[_fptr(a) for a in out]), "currents_d2h")
        return out

    def particles_h2d(self, p, ions, lecs):
        """p: structured array of PARTICLE_DTYPE with maxptl entries (ions at 0.., electrons at maxhlf..)."""
        assert p.dtype == PARTICLE_DTYPE and p.size >= self.maxhlf + lecs
        self._ck(self.lib.tgpu_particles_h2d(self.h, p.ctypes.data, ions, lecs), "particles_h2d")

    def particles_d2h(self, out=None):
        out = out if out is not None else np.zeros(self.maxptl, PARTICLE_DTYPE)
        a, b = C.c_int(), C.c_int()
        self._ck(self.lib.tgpu_particles_d2h(self.h, out.ctypes.data, C.byref(a), C.byref(b)), "particles_d2h")
        return out, a.value, b.value

    def counts(self):
        a, b = C.c_int(), C.c_int()
        self._ck(self.lib.tgpu_counts(self.h, C.byref(a), C.byref(b)), "counts")
        return a.value, b.value

    def append_particles(self, p, n_ion, n_lec):
        assert p.dtype == PARTICLE_DTYPE and p.size >= n_ion + n_lec
        self._ck(self.lib.tgpu_append_particles(self.h, p.ctypes.data, n_ion, n_lec), "append_particles")

    # --- communicator -------------------------------------------------------------------------------
    def comm_init(self, unique_id):
        buf = (C.c_uint8 * 128).from_buffer_copy(bytes(unique_id))
        self._ck(self.lib.tgpu_comm_init(self.h, buf), "comm_init")

    def comm_init_torch(self):
        """Bootstrap NCCL through an already initialised torch.distributed group (any backend)."""
        import torch.distributed as dist
        ids = [unique_id() if dist.get_rank() == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        self.comm_init(ids[0])

    # --- instrumentation -----------------------------------------------------------------------------
    def timers(self, reset=False):
        out = (C.c_double * len(PHASES))()
        self._ck(self.lib.tgpu_timers(self.h, out, int(reset)), "timers")
        return dict(zip(PHASES, out))

    def launch_count(self):
        return int(self.lib.tgpu_launch_count(self.h))

    def stream(self):
        return self.lib.tgpu_stream(self.h)

    def halo_transport(self):
        """1 = field halos over cudaIpc peer memory, 0 = NCCL send/recv"""
        return int(self.lib.tgpu_halo_transport(self.h))

    def set_option(self, name, value):
        self._ck(self.lib.tgpu_set_option(self.h, name.encode(), int(value)), "set_option")

    def field_bc_user_shock(self, leftwall, binit, btheta, bphi, beta):
        self._ck(self.lib.tgpu_field_bc_user_shock(self.h, leftwall, binit, btheta, bphi, beta), "field_bc_user_shock")

    def step_mirror(self, fields, p, ions, lecs):
        """One mirror-mode lap: `fields` (six float32 arrays) and the particle array `p` (PARTICLE_DTYPE, length >= maxptl)
        are read from and written back to host memory; returns (ions, lecs).  tgpu_step_mirror in include/tristan_gpu.h."""
        assert p.dtype == PARTICLE_DTYPE and p.flags["C_CONTIGUOUS"]
        ni, ne = C.c_int(ions), C.c_int(lecs)
        self._ck(self.lib.tgpu_step_mirror(self.h, *[_fptr(a) for a in fields], p.ctypes.data_as(C.c_void_p),
                                           C.byref(ni), C.byref(ne)), "step_mirror")
        return ni.value, ne.value

    def spectrum(self, mx0, splitratio=10.0, gambins=200, gamma_range=None):
        """per-rank part of save_spectrum (output.F90:380-633) -> (gammin, gammax, specp, spece, specprest, specerest);
        arrays shaped (gambins, nbins).  `gamma_range` = the allreduced (gammin, gammax) when there are several ranks."""
        lo, hi = C.c_float(), C.c_float()
        self._ck(self.lib.tgpu_spectrum_gamma_range(self.h, C.byref(lo), C.byref(hi)), "spectrum_gamma_range")
        glo, ghi = (lo.value, hi.value) if gamma_range is None else gamma_range
        nbins = max((mx0 - 2 - 3) // 100, 1)
        out = [np.zeros((gambins, nbins), np.float32) for _ in range(4)]
        self._ck(self.lib.tgpu_spectrum(self.h, glo, ghi, mx0, splitratio, nbins, gambins, *[_fptr(a) for a in out]), "spectrum")
        return (lo.value, hi.value, *out)

    def select_particles(self, stride, capacity):
        """prtl.tot sub-sample (output.F90:3526-3551): returns (ions, electrons) with modulo(ind/2, stride) == 0."""
        out = np.zeros(2 * capacity, PARTICLE_DTYPE)
        a, b = C.c_int(), C.c_int()
        self._ck(self.lib.tgpu_select_particles(self.h, stride, out.ctypes.data_as(C.c_void_p), capacity, C.byref(a), C.byref(b)),
                 "select_particles")
        return out[:a.value].copy(), out[capacity:capacity + b.value].copy()

    def meanq_fld_cur(self, totname):
        """output.F90:5229-5486 on the device: moment `totname` into curx (read it with currents_d2h()[0])."""
        self._ck(self.lib.tgpu_meanq_fld_cur(self.h, totname.encode()), "meanq_fld_cur")

    def particle_bc_user_wall(self, leftwall):
        self._ck(self.lib.tgpu_particle_bc_user_wall(self.h, leftwall), "particle_bc_user_wall")

    def set_user_hooks(self, kind, params=None):
        arr = (C.c_float * 5)(*(params or [0] * 5))
        self._ck(self.lib.tgpu_set_user_hooks(self.h, kind, arr), "set_user_hooks")

    def step(self, nlaps=1):
        self._ck(self.lib.tgpu_step(self.h, nlaps), "step")


def _add_procedure(name):
    def method(self):
        self._ck(getattr(self.lib, "tgpu_" + name)(self.h), name)
    method.__name__ = name
    method.__doc__ = f"tgpu_{name} (include/tristan_gpu.h) -- replaces the reference's `call {name}()`."
    setattr(Context, name, method)


for _n in ["advance_b_halfstep", "advance_e_fullstep", "reset_currents", "add_current", "bc_b1", "bc_e1", "bc_b2",
           "bc_e2", "exchange_current", "apply_filter", "apply_filter1_opt", "apply_filter2_opt", "move_particles",
           "deposit_particles", "exchange_particles", "inject_others", "reorder_particles",
           "pre_bc_b", "post_bc_b", "pre_bc_e", "post_bc_e"]:
    _add_procedure(_n)


def unique_id():
    _preload_nccl()
    buf = (C.c_uint8 * 128)()
    L = load_library()
    rc = L.tgpu_comm_unique_id(buf)
    if rc:
        raise TristanGPUError(f"tgpu_comm_unique_id failed: {L.tgpu_last_error().decode()}")
    return bytes(buf)


def _preload_nccl():
    """Point the library's dlopen at the NCCL that ships with torch (nvidia-nccl wheel)."""
    if os.environ.get("TGPU_NCCL_LIB"):
        return
    try:
        import nvidia
        for base in list(nvidia.__path__):
            cand = os.path.join(base, "nccl", "lib", "libnccl.so.2")
            if os.path.exists(cand):
                os.environ["TGPU_NCCL_LIB"] = cand
                return
    except Exception:
        pass


def device_count():
    return load_library().tgpu_device_count()
