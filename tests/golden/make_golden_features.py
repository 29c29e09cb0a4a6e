"""Generates tests/golden/feat_*.npz from the CPU oracle: radiation boundary `surface` (bc_b2 / bc_e2), the 4th-order `_42`
field solver and the meanq_fld_cur moments.  Same caveat as make_golden.py: these freeze the oracle's behaviour and give the
GPU a file-based target; they do not pin the oracle to the (unbuildable) reference.
Re-run with:  python tests/golden/make_golden_features.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
import pic_testlib as T  # noqa: E402
from oracle import oracle as O  # noqa: E402

CASES = {"d3": dict(dim=3, n=(12, 10, 8)), "d2": dict(dim=2, n=(14, 12, 1))}
MOMENTS = ["tdens", "idens", "ibetx", "ebetz", "tmomy", "eener", "iety2"]


def world(c, **kw):
    return T.oracle_world(dim=c["dim"], order=2, n=c["n"], ppc=2.0, delgam=0.05, seed_fields=4, **kw)


def main():
    for name, c in CASES.items():
        out = {}
        # radiation boundary: open x, two applications of bc_b2 + bc_e2
        w = world(c, periodic=(0, 1, 1))
        r = w.ranks[0]
        for a in range(6):
            out["surf_in_" + O.ARR_NAMES[a]] = r.arr(a).copy()
        for _ in range(2):
            for ph in (O.PH_SURF_B, O.PH_BC_B1, O.PH_SURF_E, O.PH_BC_E1):
                w.phase(ph)
        for a in range(6):
            out["surf_out_" + O.ARR_NAMES[a]] = r.arr(a).copy()
        # 4th-order solver: B half, E full, B half (periodic)
        w = world(c, highorder=1)
        r = w.ranks[0]
        for a in range(6):
            out["s42_in_" + O.ARR_NAMES[a]] = r.arr(a).copy()
        for nm in ("advance_b_halfstep", "advance_e_fullstep", "advance_b_halfstep"):
            r.call(nm)
        for a in range(6):
            out["s42_out_" + O.ARR_NAMES[a]] = r.arr(a).copy()
        # moments of the loaded plasma
        w = world(c)
        r = w.ranks[0]
        out["mom_ions"] = T.sort_particles(r.ions().copy()); out["mom_lecs"] = T.sort_particles(r.lecs().copy())
        for m in MOMENTS:
            w.meanq_fld_cur(m)
            out["mom_" + m] = r.arr(O.CURX).copy()
        np.savez_compressed(os.path.join(HERE, f"feat_{name}.npz"), **out)
        print(name, sum(v.nbytes for v in out.values()), "bytes")


if __name__ == "__main__":
    main()
