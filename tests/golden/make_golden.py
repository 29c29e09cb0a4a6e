"""Generates tests/golden/*.npz from the CPU oracle (oracle/pic_oracle.c).

The reference ships no golden vectors and cannot be built in this image (no Fortran/MPI), so these fixtures do NOT pin
the oracle to the reference; they freeze the oracle's current behaviour (regression pin) and give the GPU tests a
file-based target that does not need the oracle at run time.  Re-run with:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
import pic_testlib as T  # noqa: E402
from oracle import oracle as O  # noqa: E402

CASES = [dict(dim=2, order=1, n=(12, 10, 1), kind=1), dict(dim=2, order=2, n=(12, 10, 1), kind=1),
         dict(dim=3, order=0, n=(8, 8, 8), kind=1), dict(dim=3, order=1, n=(8, 8, 8), kind=1),
         dict(dim=3, order=2, n=(8, 8, 8), kind=2), dict(dim=3, order=3, n=(8, 8, 8), kind=2),
         dict(dim=2, order=0, n=(12, 10, 1), kind=1), dict(dim=2, order=3, n=(12, 10, 1), kind=1)]


def world(c):
    return T.oracle_world(dim=c["dim"], order=c["order"], n=c["n"], ppc=2.0, ntimes=3, filter_kind=c["kind"], delgam=0.05)


def snapshot(r):
    d = {O.ARR_NAMES[a]: r.arr(a).copy() for a in range(9)}
    d["ions"] = T.sort_particles(r.ions().copy())
    d["lecs"] = T.sort_particles(r.lecs().copy())
    return d


def main():
    for c in CASES:
        w = world(c)
        r = w.ranks[0]
        out = {}
        for k, v in snapshot(r).items():
            out["in_" + k] = v
        r.call("move_particles")
        for k, v in snapshot(r).items():
            if k in ("ions", "lecs"):
                out["moved_" + k] = v
        r.call("reset_currents"); r.call("deposit_currents_only")
        for a in range(6, 9):
            out["dep_" + O.ARR_NAMES[a]] = r.arr(a).copy()
        w2 = world(c)
        for _ in range(2):
            w2.step()
        for k, v in snapshot(w2.ranks[0]).items():
            out["lap2_" + k] = v
        name = f"case_d{c['dim']}_o{c['order']}.npz"
        np.savez_compressed(os.path.join(HERE, name), **out)
        print(name, os.path.getsize(os.path.join(HERE, name)))


if __name__ == "__main__":
    main()
