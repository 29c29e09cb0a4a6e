"""host/tristan_mainloop.cpp: the C++ host side above the C ABI (mainloop's call list, tristanmainloop.F90:107-344)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "host", "tristan_mainloop")


def _build():
    import __graft_entry__ as g
    g.build()
    assert os.path.exists(BIN)


def test_host_driver_builds_and_fails_loudly_without_a_gpu():
    _build()
    import tristan_mp_pu_master_densdecomp_b200 as tg
    if tg.device_count() > 0:
        pytest.skip("a GPU is visible: the failure path is for CPU-only boxes")
    r = subprocess.run([BIN, "--laps", "1"], capture_output=True, text=True, timeout=60)
    assert r.returncode != 0 and "tristan_gpu error" in r.stderr          # no CPU fallback of any kind


@pytest.mark.gpu
def test_host_driver_modes_agree():
    """call-for-call, tgpu_step and tgpu_step_mirror run the same laps: same particle totals, same field energy"""
    _build()
    res = {}
    for mode in ("calls", "step", "mirror"):
        r = subprocess.run([BIN, "--laps", "4", "--mode", mode, "--n", "24", "16", "12", "--ppc", "8"],
                           capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr + r.stdout
        m = re.search(r"field energy ([0-9.e+-]+), particles (\d+)", r.stdout)
        res[mode] = (float(m.group(1)), int(m.group(2)))
        assert "lap    4" in r.stdout
    assert res["calls"][1] == res["step"][1] == res["mirror"][1] == 2 * int(0.5 * 8 * 24 * 16 * 12)
    for mode in ("step", "mirror"):
        assert abs(res[mode][0] - res["calls"][0]) <= 2e-3 * res["calls"][0], res
