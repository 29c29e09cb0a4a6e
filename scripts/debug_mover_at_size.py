"""debug: worst-particle report for the mover at 128x64x64 (run inside gpurun)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import tristan_mp_pu_master_densdecomp_b200 as tg
import pic_testlib as T
from oracle import oracle as O
order = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = (128, 64, 64)
w = T.oracle_world(dim=3, order=order, n=n, ppc=16.0, ntimes=4, filter_kind=2, init="uniform", seed_fields=3)
r = w.ranks[0]
ctx = tg.Context(T.gpu_params(tg, w, device=0))
T.upload(ctx, r)
p0 = r.particles().copy()
ctx.bc_b1(); ctx.bc_e1(); ctx.advance_b_halfstep(); ctx.bc_b1()
ctx.move_particles()
for ph in (O.PH_BC_B1, O.PH_BC_E1, O.PH_BHALF, O.PH_BC_B1, O.PH_MOVE):
    w.phase(ph)
fg = ctx.fields_d2h()
for a in range(6):
    print("field", a, "bit-equal:", np.array_equal(fg[a], r.arr(a)))
gi, ge = T.gpu_particles(ctx)
oi, oe = T.oracle_particles(r)
ions, lecs = r.counts
b0 = T.sort_particles(p0[:ions].copy())
for k in "xyzuvw":
    d = np.abs(gi[k].astype(np.float64) - oi[k])
    i = int(d.argmax())
    print(k, "max abs diff", d.max(), "at", i, "count>1e-4:", int((d > 1e-4).sum()), "count>2e-5:", int((d > 2e-5).sum()))
d = np.abs(gi["x"].astype(np.float64) - oi["x"])
for i in np.argsort(-d)[:5]:
    print("before", b0[i], "\n gpu  ", gi[i], "\n orc  ", oi[i])
