"""Reader / writer for the reference's restart dumps, so that device state can be exchanged with a real Fortran build.

The reference writes two Fortran *unformatted sequential* files per rank, one record each
(code/output.F90:2194-2226 writes them, code/restart.F90:209-253 reads them back):

  restflds.<job>.<rank>.d :  mx, my, mz (int32) | ex, ey, ez, bx, by, bz (float32, Fortran order (mx,my,mz)) |
                             dseed (float64) | lap (int32) | xinject, xinject2, xinject3 (float64: real(dprec),
                             fields.F90:58) | leftwall (float32: real(sprec), particles.F90:67-68) | walloc (float64,
                             particles.F90:74) -- a 36-byte tail, no padding inside a Fortran record
  restprtl.<job>.<rank>.d :  ions, lecs, maxptl, maxhlf, totalpartnum (int32) | then, attribute by attribute, the ions
                             followed by the electrons: x y z u v w ch (float32), ind proc splitlev (int32)

i.e. the particle dump is already structure-of-arrays — the layout the device keeps.  A record is framed by 4-byte
length markers (gfortran and ifort defaults); a record longer than `max_subrecord` bytes (2**31 - 9 in gfortran) is split
into sub-records whose leading marker is negative when another sub-record follows and whose trailing marker is negative
when one precedes (the gfortran convention).  This module is host-side glue: it is not on the hot path.
"""
import struct

import numpy as np

from . import PARTICLE_DTYPE

_FLOAT_ATTRS = ("x", "y", "z", "u", "v", "w", "ch")
_INT_ATTRS = ("ind", "proc", "splitlev")
GFORTRAN_MAX_SUBRECORD = 2 ** 31 - 9
_TAIL_FMT = "<3dfd"          # xinject, xinject2, xinject3 (dprec), leftwall (sprec), walloc (dprec): 36 bytes


def _write_record(f, payload, max_subrecord=GFORTRAN_MAX_SUBRECORD):
    """payload: bytes-like; written as one logical record, split into sub-records where needed."""
    mv = memoryview(payload).cast("B")
    n = len(mv)
    off, first = 0, True
    while True:
        chunk = min(n - off, max_subrecord)
        more = off + chunk < n
        lead = np.int32(-chunk if more else chunk)
        trail = np.int32(chunk if first else -chunk)
        f.write(lead.tobytes()); f.write(mv[off:off + chunk]); f.write(trail.tobytes())
        off += chunk
        first = False
        if not more:
            break


def _read_record(f):
    """one logical record -> bytes (sub-records concatenated)"""
    parts = []
    while True:
        head = f.read(4)
        if len(head) != 4:
            raise EOFError("truncated Fortran record")
        lead = int(np.frombuffer(head, np.int32)[0])
        n = abs(lead)
        data = f.read(n)
        tail = f.read(4)
        if len(data) != n or len(tail) != 4 or abs(int(np.frombuffer(tail, np.int32)[0])) != n:
            raise ValueError("corrupt Fortran record markers")
        parts.append(data)
        if lead >= 0:
            break
    return b"".join(parts) if len(parts) > 1 else parts[0]


def write_fields(path, fields, dseed=0.0, lap=0, xinject=0.0, xinject2=0.0, xinject3=0.0, leftwall=0.0, walloc=0.0,
                 max_subrecord=GFORTRAN_MAX_SUBRECORD):
    """fields: six float32 arrays shaped (mz, my, mx) (C order == Fortran (mx,my,mz)), ex ey ez bx by bz."""
    mz, my, mx = fields[0].shape
    blob = [np.array([mx, my, mz], np.int32).tobytes()]
    for a in fields:
        assert a.dtype == np.float32 and a.shape == (mz, my, mx)
        blob.append(np.ascontiguousarray(a).tobytes())
    blob.append(np.float64(dseed).tobytes())
    blob.append(np.int32(lap).tobytes())
    blob.append(struct.pack(_TAIL_FMT, xinject, xinject2, xinject3, leftwall, walloc))
    with open(path, "wb") as f:
        _write_record(f, b"".join(blob), max_subrecord)


def read_fields(path):
    """-> (fields [ex..bz] shaped (mz,my,mx), scalars dict)"""
    with open(path, "rb") as f:
        rec = _read_record(f)
    mx, my, mz = (int(v) for v in np.frombuffer(rec, np.int32, 3))
    n = mx * my * mz
    want = 12 + 24 * n + 8 + 4 + struct.calcsize(_TAIL_FMT)
    if len(rec) != want:
        raise ValueError(f"restflds record is {len(rec)} bytes, expected {want} for {mx}x{my}x{mz} "
                         "(output.F90:2198: 3 int32, 6 float32 arrays, dseed f64, lap i32, 3 f64, f32, f64)")
    off = 12
    fields = []
    for _ in range(6):
        fields.append(np.frombuffer(rec, np.float32, n, off).reshape(mz, my, mx).copy())
        off += 4 * n
    dseed = float(np.frombuffer(rec, np.float64, 1, off)[0]); off += 8
    lap = int(np.frombuffer(rec, np.int32, 1, off)[0]); off += 4
    tail = struct.unpack_from(_TAIL_FMT, rec, off)
    scal = dict(dseed=dseed, lap=lap, xinject=tail[0], xinject2=tail[1], xinject3=tail[2],
                leftwall=tail[3], walloc=tail[4])
    return fields, scal


def write_particles(path, p, ions, lecs, maxptl, totalpartnum=0, max_subrecord=GFORTRAN_MAX_SUBRECORD):
    """p: PARTICLE_DTYPE array in the reference's layout (ions at [0, ions), electrons at [maxhlf, maxhlf + lecs))."""
    assert p.dtype == PARTICLE_DTYPE
    maxhlf = maxptl // 2
    ion, lec = p[:ions], p[maxhlf:maxhlf + lecs]
    blob = [np.array([ions, lecs, maxptl, maxhlf, totalpartnum], np.int32).tobytes()]
    for k in _FLOAT_ATTRS + _INT_ATTRS:
        blob.append(np.ascontiguousarray(ion[k]).tobytes()); blob.append(np.ascontiguousarray(lec[k]).tobytes())
    with open(path, "wb") as f:
        _write_record(f, b"".join(blob), max_subrecord)


def read_particles(path, maxptl=None):
    """-> (p, ions, lecs, header dict).  `maxptl` = capacity of the returned array (restart.F90:232-233 recomputes it from
    the input file rather than trusting the dump); defaults to the dumped value."""
    with open(path, "rb") as f:
        rec = _read_record(f)
    ions, lecs, maxptl_d, maxhlf_d, total = (int(v) for v in np.frombuffer(rec, np.int32, 5))
    maxptl = maxptl_d if maxptl is None else maxptl
    maxhlf = maxptl // 2
    if ions > maxhlf or lecs > maxhlf:
        raise ValueError(f"dump holds {ions}+{lecs} particles, capacity maxhlf = {maxhlf}")
    p = np.zeros(maxptl, PARTICLE_DTYPE)
    off = 20
    n = ions + lecs
    for k in _FLOAT_ATTRS + _INT_ATTRS:
        dt = np.float32 if k in _FLOAT_ATTRS else np.int32
        col = np.frombuffer(rec, dt, n, off); off += 4 * n
        p[k][:ions] = col[:ions]
        p[k][maxhlf:maxhlf + lecs] = col[ions:]
    return p, ions, lecs, dict(maxptl=maxptl_d, maxhlf=maxhlf_d, totalpartnum=total)


def load(ctx, fld_path, prt_path):
    """restart(): read both dumps of this rank and upload them to the device context; returns the scalars"""
    fields, scal = read_fields(fld_path)
    p, ions, lecs, hdr = read_particles(prt_path, maxptl=ctx.maxptl)
    ctx.fields_h2d(*fields)
    ctx.particles_h2d(p, ions, lecs)
    scal.update(hdr)
    return scal


def save(ctx, fld_path, prt_path, **scalars):
    """write_restart(): dump the device state of this rank in the reference's format"""
    fields = ctx.fields_d2h()
    p, ions, lecs = ctx.particles_d2h()
    tot = scalars.pop("totalpartnum", 0)
    write_fields(fld_path, fields, **scalars)
    write_particles(prt_path, p, ions, lecs, ctx.maxptl, totalpartnum=tot)
