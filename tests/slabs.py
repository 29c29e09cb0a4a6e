"""Slab-exchange plan: which boxes travel to which neighbour in each exchange phase of a lap.

This is the host-side description of the protocol that csrc/fields.cu + csrc/particles.cu execute over NCCL
(pack kernel -> ncclSend/ncclRecv -> unpack/add kernel).  It is pure index arithmetic on the reference's ghost maps
(code/fieldboundaries.F90:181-263, 1768-2189; code/filter.F90:71-99;
code/optimized_filters.F90:1387-1963) and needs no GPU, so the same plan can be driven over any transport --
tests/test_gloo_slabs.py runs it over torch.distributed/gloo with two processes.

A plan is a list of Shift(axis, direction, src_lo, src_hi, dst_lo, dst_hi, mode, recv_ok, arrays):
  every rank packs box [src_lo..src_hi] of `arrays`, sends it to its neighbour in `direction` along `axis`, receives
  the matching box from the opposite neighbour and (if recv_ok) writes ("put") or accumulates ("add") it into
  [dst_lo..dst_hi].  Boxes are 1-based inclusive (i,j,k) triples, Fortran style.
"""
from collections import namedtuple

from tristan_mp_pu_master_densdecomp_b200 import neighbour

Shift = namedtuple("Shift", "axis direction src_lo src_hi dst_lo dst_hi mode recv_ok arrays")

FIELDS_E, FIELDS_B, CURRENTS = (0, 1, 2), (3, 4, 5), (6, 7, 8)


class Slab:
    """geometry of one rank (the numbers the reference keeps in m_fields / m_communications)"""

    def __init__(self, dim, rank, sizes, dims, nghost, nghostz, periodic, ntimes=0):
        self.dim, self.rank, self.sizes = dim, rank, tuple(sizes)
        self.m = tuple(dims)
        self.ng = (nghost, nghost, nghostz)
        self.g = (nghost // 2, nghost // 2, nghostz // 2)
        self.periodic = tuple(periodic)
        self.ntimes = ntimes
        sx, sy, sz = self.sizes
        self.pos = (rank % sx, (rank % (sx * sy)) // sx, rank // (sx * sy))
        self.naxes = 3 if dim == 3 else 2

    def neighbour(self, axis, direction):
        return neighbour(self.rank, *self.sizes, 2 * axis + (1 if direction > 0 else 0))

    def full(self):
        return [1, 1, 1], list(self.m)

    def recv_ok(self, axis, direction):
        """open boundaries: an edge rank does not take the wrapped-around message (copy_layr*2_opt)"""
        if self.periodic[axis]:
            return True
        if direction > 0:      # message travelling up arrives from the - side: rank 0 on the axis has no - neighbour
            return self.pos[axis] != 0
        return self.pos[axis] != self.sizes[axis] - 1

    def active(self, axis):
        """axes on which anything is exchanged or wrapped"""
        return self.periodic[axis] or self.sizes[axis] > 1

    # ---- plans -------------------------------------------------------------------------------------------
    def plan_ghost_refresh(self, arrays):
        """bc_b1 / bc_e1: g layers per side, x then y then z, full extent of the other axes"""
        plan = []
        for axis in range(self.naxes):
            if not self.active(axis):
                continue
            m, g = self.m[axis], self.g[axis]
            lo, hi = self.full()
            s_lo, s_hi, d_lo, d_hi = list(lo), list(hi), list(lo), list(hi)
            s_lo[axis], s_hi[axis], d_lo[axis], d_hi[axis] = m - 2 * g, m - g - 1, 1, g
            plan.append(Shift(axis, +1, s_lo, s_hi, d_lo, d_hi, "put", self.recv_ok(axis, +1), arrays))
            s_lo, s_hi, d_lo, d_hi = list(lo), list(hi), list(lo), list(hi)
            s_lo[axis], s_hi[axis], d_lo[axis], d_hi[axis] = g + 1, 2 * g, m - g, m - 1
            plan.append(Shift(axis, -1, s_lo, s_hi, d_lo, d_hi, "put", self.recv_ok(axis, -1), arrays))
        return plan

    def plan_current_fold(self):
        """exchange_current: ghost deposits are added into the owning interior layers"""
        plan = []
        for axis in range(self.naxes):
            if not self.active(axis):
                continue
            m, g, ng = self.m[axis], self.g[axis], self.ng[axis]
            lo, hi = self.full()
            s_lo, s_hi, d_lo, d_hi = list(lo), list(hi), list(lo), list(hi)
            s_lo[axis], s_hi[axis], d_lo[axis], d_hi[axis] = m - g, m, g + 1, ng
            plan.append(Shift(axis, +1, s_lo, s_hi, d_lo, d_hi, "add", self.recv_ok(axis, +1), CURRENTS))
            s_lo, s_hi, d_lo, d_hi = list(lo), list(hi), list(lo), list(hi)
            s_lo[axis], s_hi[axis], d_lo[axis], d_hi[axis] = 1, g, m - ng + 1, m - g - 1
            plan.append(Shift(axis, -1, s_lo, s_hi, d_lo, d_hi, "add", self.recv_ok(axis, -1), CURRENTS))
        return plan

    def plan_filter1_refresh(self):
        """one ghost layer per side before every filter1 pass: (lt,ls,nt,ns) = (g, m-g-1, m-g, g+1)"""
        plan = []
        for axis in range(self.naxes):
            m, g = self.m[axis], self.g[axis]
            lo, hi = self.full()
            for direction, src, dst in ((+1, m - g - 1, g), (-1, g + 1, m - g)):
                s_lo, s_hi, d_lo, d_hi = list(lo), list(hi), list(lo), list(hi)
                s_lo[axis] = s_hi[axis] = src
                d_lo[axis] = d_hi[axis] = dst
                plan.append(Shift(axis, direction, s_lo, s_hi, d_lo, d_hi, "put", self.recv_ok(axis, direction), CURRENTS))
        return plan

    def filter2_halo_boxes(self, axis):
        """(box sent up = my last ntimes cells, box sent down = my first ntimes cells), interior of the other axes"""
        lo = [self.g[a] + 1 for a in range(3)]
        hi = [self.m[a] - self.g[a] - 1 for a in range(3)]
        if self.dim == 2:
            lo[2] = hi[2] = 1
        up_lo, up_hi, dn_lo, dn_hi = list(lo), list(hi), list(lo), list(hi)
        up_lo[axis] = hi[axis] - self.ntimes + 1
        dn_hi[axis] = lo[axis] + self.ntimes - 1
        return (up_lo, up_hi), (dn_lo, dn_hi)

    def migration_directions(self):
        """outbox directions that leave the rank: the reference's z, y, x order (particles.F90:1904-2112)"""
        dirs = []
        if self.dim == 3:
            dirs += [4, 5]
        if self.sizes[1] > 1:
            dirs += [2, 3]
        if self.sizes[0] > 1:
            dirs += [0, 1]
        return dirs
