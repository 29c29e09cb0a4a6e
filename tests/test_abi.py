"""CPU-side checks of the drop-in boundary: the library loads, exports every symbol include/tristan_gpu.h declares,
its host-only topology helpers agree with the oracle's restatement of the reference formulas, and it refuses to run
without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "tristan_gpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tgpu_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(tg):
    L = tg.load_library()
    names = header_functions()
    assert len(names) >= 35
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/tristan_gpu.h but not exported"
    assert sorted(tg.ABI_SYMBOLS) == names, "python ABI list out of sync with the header"


def test_params_struct_layout_matches_header(tg):
    # field order in the header == field order in the ctypes mirror
    src = open(os.path.join(ROOT, "include", "tristan_gpu.h")).read()
    body = src[src.index("typedef struct tgpu_params {"):src.index("} tgpu_params;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S).split("{", 1)[1]
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl or decl.startswith("typedef"):
            continue
        decl = re.sub(r"^(const\s+)?(int32_t|float)\s*", "", decl)
        for f in decl.split(","):
            f = f.strip().lstrip("*").strip()
            fields.append(re.sub(r"\[.*\]", "", f))
    assert fields == [f[0] for f in tg.Params._fields_]
    assert C.sizeof(tg.Params) % 8 == 0


@pytest.mark.parametrize("sizes", [(1, 1, 1), (1, 2, 4), (1, 4, 2), (1, 8, 1), (1, 1, 8), (1, 3, 2)])
def test_topology_matches_reference_formulas(tg, sizes):
    P = O.make_params(dim=3, order=2, mx0=16, my0=24, mz0=32, sizex=sizes[0], sizey=sizes[1], sizez=sizes[2], ppc0=1.0)
    w = O.World(P)
    for r in w.ranks:
        for d in range(6):
            assert tg.neighbour(r.idx, *sizes, d) == O.lib().orc_neighbour(r.h, d)
        assert tg.decompose(3, 2, 16, 24, 32, *sizes, r.idx) == (r.mx, r.my, r.mz, r.mxcum, r.mycum, r.mzcum)


def test_topology_2d_uneven_split(tg):
    P = O.make_params(dim=2, order=1, mx0=30, my0=26, sizex=4, sizey=3, ppc0=1.0)
    w = O.World(P)
    tot = 0
    for r in w.ranks:
        assert tg.decompose(2, 1, 30, 26, 1, 4, 3, 1, r.idx) == (r.mx, r.my, 1, r.mxcum, r.mycum, 0)
        tot += (r.mx - 5) * (r.my - 5)
    assert tot == 30 * 26
    assert tg.ghost_width(2, 2) == (7, 5) and tg.ghost_width(3, 1) == (5, 5) and tg.ghost_width(3, 3) == (7, 7)


def test_charge_normalisation_matches_oracle(tg):
    P = O.make_params(dim=3, order=2, ppc0=16.0, c_omp=10.0, gamma0=0.5)
    qe, qi, qme, qmi = tg.charge_normalisation(0.45, 10.0, 16.0, 0.5, 1.0, 1.0)
    assert np.float32(qe) == np.float32(P.qe) and np.float32(qmi) == np.float32(P.qmi) and np.float32(qme) == np.float32(P.qme)


def test_no_cpu_fallback(tg):
    if tg.device_count() > 0:
        pytest.skip("a GPU is visible; the refusal path is exercised on the CPU box")
    with pytest.raises(tg.TristanGPUError, match="no CUDA device"):
        tg.Context(tg.make_params(dim=3, order=2, mx0=8, my0=8, mz0=8))


def test_product_package_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "tristan_mp_pu_master_densdecomp_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(base, f)).read()
                assert "pic_oracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, f


def test_fortran_bridge_is_in_step_with_the_header():
    """host/gpu_bridge.F90 (the ISO_C_BINDING module the Fortran driver uses) is generated from include/tristan_gpu.h:
    regenerating it must reproduce the committed file, and every exported symbol must have an interface in it"""
    import subprocess
    import sys
    assert subprocess.call([sys.executable, os.path.join(ROOT, "scripts", "gen_gpu_bridge.py"), "--check"]) == 0, \
        "host/gpu_bridge.F90 is stale: run scripts/gen_gpu_bridge.py"
    f90 = open(os.path.join(ROOT, "host", "gpu_bridge.F90")).read()
    import tristan_mp_pu_master_densdecomp_b200 as tgm
    for sym in tgm.ABI_SYMBOLS:
        assert f'bind(C, name="{sym}")' in f90, sym
