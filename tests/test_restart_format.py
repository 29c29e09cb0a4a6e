"""Reference restart dumps (code/output.F90:2194-2226, code/restart.F90:209-253): format round trips on the CPU, and the
device round trip (save -> load -> identical state) on the GPU."""
import struct

import numpy as np
import pytest

import tristan_mp_pu_master_densdecomp_b200 as tg
from tristan_mp_pu_master_densdecomp_b200 import restart as R


def _state(seed=0, n=(9, 7, 5), ions=37, lecs=29, maxptl=100):
    rng = np.random.default_rng(seed)
    fields = [rng.standard_normal((n[2], n[1], n[0])).astype(np.float32) for _ in range(6)]
    p = np.zeros(maxptl, tg.PARTICLE_DTYPE)
    for first, cnt in ((0, ions), (maxptl // 2, lecs)):
        for k in ("x", "y", "z", "u", "v", "w", "ch"):
            p[k][first:first + cnt] = rng.standard_normal(cnt).astype(np.float32)
        p["ind"][first:first + cnt] = rng.integers(-1000, 1000, cnt)
        p["proc"][first:first + cnt] = rng.integers(0, 8, cnt)
        p["splitlev"][first:first + cnt] = 1
    return fields, p, ions, lecs, maxptl


@pytest.mark.parametrize("max_sub", [R.GFORTRAN_MAX_SUBRECORD, 1000, 64])
def test_restart_round_trip(tmp_path, max_sub):
    fields, p, ions, lecs, maxptl = _state()
    ff, fp = tmp_path / "restflds.d", tmp_path / "restprtl.d"
    R.write_fields(ff, fields, dseed=123457.25, lap=42, xinject=20.0, xinject2=510.0, xinject3=3.0, leftwall=20.0, walloc=21.5,
                   max_subrecord=max_sub)
    R.write_particles(fp, p, ions, lecs, maxptl, totalpartnum=66, max_subrecord=max_sub)
    f2, scal = R.read_fields(ff)
    for a, b in zip(fields, f2):
        assert np.array_equal(a, b)
    assert scal == dict(dseed=123457.25, lap=42, xinject=20.0, xinject2=510.0, xinject3=3.0, leftwall=20.0, walloc=21.5)
    p2, i2, l2, hdr = R.read_particles(fp)
    assert (i2, l2) == (ions, lecs) and hdr == dict(maxptl=maxptl, maxhlf=maxptl // 2, totalpartnum=66)
    assert np.array_equal(p2, p)
    # a larger capacity on re-read (restart.F90:232-233 takes maxptl from the input file, not from the dump)
    p3, _, _, _ = R.read_particles(fp, maxptl=200)
    assert np.array_equal(p3[:ions], p[:ions]) and np.array_equal(p3[100:100 + lecs], p[maxptl // 2:maxptl // 2 + lecs])


def test_restart_record_layout_is_fortran_sequential(tmp_path):
    """byte-level check of the framing: 4-byte marker, payload in write(7) order, 4-byte marker"""
    fields, p, ions, lecs, maxptl = _state(n=(3, 2, 2))
    ff = tmp_path / "f.d"
    R.write_fields(ff, fields, dseed=1.5, lap=7, xinject=20.125, xinject2=510.5, xinject3=3.25, leftwall=15.0, walloc=21.0625)
    raw = ff.read_bytes()
    # tail after lap: xinject, xinject2, xinject3 real(dprec) (fields.F90:58), leftwall real(sprec) (particles.F90:67-68),
    # walloc real(dprec) (particles.F90:74) = 3*8 + 4 + 8 = 36 bytes, unpadded (output.F90:2198)
    nbytes = 12 + 6 * 12 * 4 + 8 + 4 + 36
    assert struct.unpack("<i", raw[:4])[0] == nbytes and struct.unpack("<i", raw[-4:])[0] == nbytes and len(raw) == nbytes + 8
    assert struct.unpack("<3i", raw[4:16]) == (3, 2, 2)
    assert np.array_equal(np.frombuffer(raw, np.float32, 12, 16), fields[0].ravel())      # ex, x fastest
    assert struct.unpack("<d", raw[16 + 288:16 + 296])[0] == 1.5 and struct.unpack("<i", raw[16 + 296:16 + 300])[0] == 7
    assert struct.unpack("<3d", raw[16 + 300:16 + 324]) == (20.125, 510.5, 3.25)            # float64 offsets
    assert struct.unpack("<f", raw[16 + 324:16 + 328])[0] == 15.0 and struct.unpack("<d", raw[16 + 328:16 + 336])[0] == 21.0625
    # a record of the wrong length (e.g. the five-float32 tail an earlier version wrote) is refused, not mis-decoded
    bad = tmp_path / "bad.d"
    body = raw[4:-4][:-36] + np.zeros(5, np.float32).tobytes()
    bad.write_bytes(struct.pack("<i", len(body)) + body + struct.pack("<i", len(body)))
    with pytest.raises(ValueError, match="restflds record"):
        R.read_fields(bad)
    # split records: leading marker negative while another sub-record follows, trailing marker negative when one precedes
    R.write_fields(ff, fields, max_subrecord=100)
    raw = ff.read_bytes()
    assert struct.unpack("<i", raw[:4])[0] == -100 and struct.unpack("<i", raw[104:108])[0] == 100
    assert struct.unpack("<i", raw[108:112])[0] == -100 and struct.unpack("<i", raw[212:216])[0] == -100


@pytest.mark.gpu
def test_restart_device_round_trip(tmp_path):
    import pic_testlib as T
    if tg.device_count() < 1:
        pytest.fail("no CUDA device visible")
    w = T.oracle_world(dim=3, order=2, n=(12, 10, 8), ppc=4.0)
    r = w.ranks[0]
    ctx = tg.Context(T.gpu_params(tg, w, device=0))
    T.upload(ctx, r)
    ctx.step(2)
    R.save(ctx, tmp_path / "restflds.d", tmp_path / "restprtl.d", lap=2, dseed=1.0)
    ctx2 = tg.Context(T.gpu_params(tg, w, device=0))
    scal = R.load(ctx2, tmp_path / "restflds.d", tmp_path / "restprtl.d")
    assert scal["lap"] == 2
    for a, b in zip(ctx.fields_d2h(), ctx2.fields_d2h()):
        assert np.array_equal(a, b)
    for a, b in zip(T.gpu_particles(ctx), T.gpu_particles(ctx2)):
        assert np.array_equal(a, b)
    ctx.step(1); ctx2.step(1)                      # and both continue identically
    for a, b in zip(ctx.fields_d2h(), ctx2.fields_d2h()):
        assert T.max_rel(a, b) < 1e-5
    ctx.close(); ctx2.close()
