#!/usr/bin/env python
"""bench.py -- particle-steps/s of the PIC hot path (one "step" = one lap of mainloop over one rank's particles).

Workload (config.workload): 3D Weibel, 2nd-order shapes (dd2, nghost 7), 16 ppc, filter2 with ntimes = 32, per-GPU slab
512x256x128 cells -- the per-GPU share of BASELINE.json configs[2] (512^3 over 8 B200 as sizey x sizez = 2 x 4).  The slab
per GPU is fixed as N grows (weak scaling); at N = 8 the job IS configs[2].

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path through the C ABI
  python bench.py --impl reference ...                     # the reference's CPU algorithm (oracle restatement; the
                                                           # Fortran build cannot be compiled in this image) on host cores
Prints ONE JSON line (see the task contract): value = resident throughput, e2e = mirror-mode throughput with the full
state crossing PCIe every lap, roofline for the dominant kernels, cpu_baseline on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GRID = {1: (1, 1), 2: (1, 2), 4: (2, 2), 8: (2, 4)}       # sizey, sizez
PPC = 16.0
ORDER = 2
NTIMES = 32
FILTER_KIND = 2
B_PER_PARTICLE = 52.0 + 36.0 / PPC                         # mover+deposit algorithmic bytes


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cells", type=int, nargs=3, default=[512, 256, 128], help="per-GPU slab (interior cells)")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-plain", action="store_true", help="mirror lap as separate h2d / step / d2h calls instead of tgpu_step_mirror")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--fused", type=int, default=1)
    ap.add_argument("--peer", type=int, default=1, help="field halos over cudaIpc peer memory (1) or NCCL send/recv (0)")
    ap.add_argument("--order", type=int, default=ORDER, help="shape order (headline = 2)")
    ap.add_argument("--ppc", type=float, default=PPC)
    ap.add_argument("--cpu-cells", type=int, nargs=3, default=[128, 32, 32], help="CPU sample: cells per host-thread slab")
    ap.add_argument("--config", default="default", choices=["default", "c0", "c1", "c3", "c4"],
                    help="BASELINE.json configs: default = per-GPU share of configs[2] (the headline); c0 = configs[0] 2D Weibel "
                         "128^2 dd1 (GPU against ONE CPU rank); c1 = configs[1] 2D two-stream 128x2 dd2; c3 = configs[3] 3D shock "
                         "512x128x128 dd3 with wall, clamps and a per-lap injector; c4 = configs[4] 3D Weibel 64 ppc "
                         "(256x256x128 per GPU, --order 1|2|3).  Only `default` is the driver's bench line.")
    a = ap.parse_args()
    if a.config == "c4":
        a.cells, a.ppc = [256, 256, 128], 64.0
    return a


# ------------------------------------------------------------------------------------------------------------
# The other BASELINE.json configs (single GPU): small 2D problems and the shock.  Same metric, same JSON shape; these are
# extra lines for profiles/, never the driver's headline.
# ------------------------------------------------------------------------------------------------------------
def beams(rng, n, lo, hi, zlo, zhi, drift, uth, x_sign=None):
    """n particles, uniform in the box, drifting along +-x (alternating, or all along x_sign) with a thermal spread"""
    import tristan_mp_pu_master_densdecomp_b200 as tg
    p = np.zeros(n, tg.PARTICLE_DTYPE)
    for k, (a, b) in zip("xyz", (lo[0:1] + hi[0:1], lo[1:2] + hi[1:2], [zlo, zhi])):
        p[k] = (a + (b - a) * rng.random(n)).astype(np.float32)
        np.minimum(p[k], np.nextafter(np.float32(b), np.float32(0)), out=p[k])
    sign = np.where((np.arange(n) & 1) == 0, 1.0, -1.0) if x_sign is None else x_sign
    gb = drift / np.sqrt(1 - drift * drift)
    p["u"] = (sign * gb + uth * rng.standard_normal(n)).astype(np.float32)
    p["v"] = (uth * rng.standard_normal(n)).astype(np.float32)
    p["w"] = (uth * rng.standard_normal(n)).astype(np.float32)
    p["ch"] = 1.0; p["splitlev"] = 1
    return p


def run_small_config(args):
    import torch
    import __graft_entry__ as ge
    ge.build()
    import tristan_mp_pu_master_densdecomp_b200 as tg
    rng = np.random.default_rng(77)
    cfg = args.config
    if cfg == "c0":       # user/input.weibel: 128 x 128, dd1, 16 ppc, ntimes = 32, filter1 (default build), cold counter-streaming beams
        kw = dict(dim=2, order=1, mx0=128, my0=128, ntimes=32, filter_kind=1, ppc0=16.0)
        name = "configs[0]: 2D Weibel (user/input.weibel) 128x128, dd1, 16 ppc, filter1 ntimes=32, 262144 particles"
        species, drift, uth, hooks = (1, 1), 0.5, 2e-3, None
    elif cfg == "c1":     # user/input.twostream: 128 x 2, dd2, 64 ppc, electrons only
        kw = dict(dim=2, order=2, mx0=128, my0=2, ntimes=32, filter_kind=1, ppc0=64.0)
        name = "configs[1]: 2D two-stream (user/input.twostream) 128x2, dd2, 64 ppc, electrons only, 8192 particles"
        species, drift, uth, hooks = (0, 1), 0.5, 2e-3, None
    else:                 # user/input.shock made 3D (sizex = 1): dd3, open x, reflecting wall at leftwall, clamps, injector
        kw = dict(dim=3, order=3, mx0=512, my0=128, mz0=128, ntimes=4, filter_kind=2, ppc0=16.0, periodic=(0, 1, 1))
        name = "configs[3]: 3D shock (user/user_shock.F90 hooks) 512x128x128, dd3, 16 ppc upstream, wall at x=20, injector at the right edge"
        species, drift, uth, hooks = (1, 1), 0.4, 0.05, (20.0, 0.05, 0.3, 1.2, 0.4)
    ncell = kw["mx0"] * kw["my0"] * kw.get("mz0", 1)
    nsp = int(0.5 * kw["ppc0"] * ncell)
    P = tg.make_params(maxptl=int(2 * nsp * 1.6) + 8192, device=0, **kw)
    ctx = tg.Context(P)
    g, gz = P.nghost // 2, P.nghostz // 2
    lo, hi = [g + 1.0, g + 1.0], [float(P.mx - g), float(P.my - g)]
    zlo, zhi = (gz + 1.0, float(P.mz - gz)) if P.dim == 3 else (3.0, 4.0)      # 2D: minz = 3, maxz = 4 (particles_movedeposit.F90:1371)
    if hooks:
        lo[0] = hooks[0] + 0.5                                     # plasma only upstream of the wall, flowing towards it
    maxhlf = P.maxptl // 2
    host = np.zeros(P.maxptl, tg.PARTICLE_DTYPE)
    n_i = nsp if species[0] else 0
    n_e = nsp if species[1] else 0
    for s, n, off in ((0, n_i, 0), (1, n_e, maxhlf)):
        if n:
            q = beams(rng, n, lo, hi, zlo, zhi, drift, uth, x_sign=(-1.0 if hooks else None))
            q["ind"] = np.arange(1, n + 1, dtype=np.int32) * 2 - s
            host[off:off + n] = q
    ctx.particles_h2d(host, n_i, n_e)
    zero = [np.zeros((P.mz, P.my, P.mx), np.float32) for _ in range(6)]
    ctx.fields_h2d(*zero)
    if hooks:
        ctx.set_user_hooks(1, hooks)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", 0))
    next_ind = [2 * nsp + 2]

    def inject():
        """A stand-in for the reference's host injector (user_shock.F90:303-331 -> inject_from_wall): the same number of
        particles per lap, ppc0/2 * c * beta per cell face and species, drawn with numpy in the slab the plane source fills in
        one step and uploaded with tgpu_append_particles.  Two deliberate differences: the draws are not the reference's
        MINSTD / Juttner sequence (the oracle has that: orc_inject_particles_shock, pinned in tests/test_ref_golden.py), and
        the plane sits at the last interior node mx - g, because the reference's hard-coded x = mx0 - 2 is outside the
        interior when nghost = 7 (dd2 / dd3) and its injector then adds nothing (same test, case 3)."""
        if not hooks:
            return 0
        n = int(0.5 * kw["ppc0"] * kw["my0"] * kw["mz0"] * 0.45 * drift)
        xr = float(P.mx - g)
        q = np.zeros(2 * n, tg.PARTICLE_DTYPE)
        for s in (0, 1):
            b = beams(rng, n, [xr - 0.45 * drift, lo[1]], [xr, hi[1]], zlo, zhi, drift, uth, x_sign=-1.0)
            b["ind"] = (next_ind[0] + 2 * np.arange(n, dtype=np.int32)) - s
            q[s * n:(s + 1) * n] = b
        next_ind[0] += 2 * n
        ctx.append_particles(q, n, n)
        return 2 * n

    steps, warm = (args.steps, args.warmup) if cfg == "c3" else (max(args.steps, 200), max(args.warmup, 20))
    for _ in range(warm):
        inject(); ctx.step(1)
    torch.cuda.synchronize()
    n0 = sum(ctx.counts())
    sampler = ClockSampler(0); sampler.start()
    l0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    psteps = 0
    injected = 0
    t_host0 = time.perf_counter()
    for _ in range(steps):
        injected += inject()
        ctx.step(1)
        psteps += sum(ctx.counts()) if hooks else n0
    e1.record(stream)
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - t_host0) * 1e3
    ms = max(e0.elapsed_time(e1), 1e-6)
    sampler.stop_flag = True
    launches = ctx.launch_count() - l0
    n1 = sum(ctx.counts())
    # sanity: counts (injected - removed through the open faces), finite fields, energy
    f = ctx.fields_d2h()
    fe = float(sum((a.astype(np.float64) ** 2).sum() for a in f))
    pp, ci, ce = ctx.particles_d2h()
    gam = lambda q: np.sqrt(1.0 + q["u"].astype(np.float64) ** 2 + q["v"].astype(np.float64) ** 2 + q["w"].astype(np.float64) ** 2)
    ke = float((gam(pp[:ci]) - 1).sum() + (gam(pp[maxhlf:maxhlf + ce]) - 1).sum())
    sanity = {"particles_start": n0, "particles_end": n1, "injected": injected, "field_energy_sum_sq": fe, "kinetic_sum_gamma_minus_1": ke,
              "finite": bool(np.isfinite(fe) and np.isfinite(ke))}
    ctx.close()
    value = psteps / (ms * 1e-3)
    cpu = None
    if cfg in ("c0", "c1") and not args.no_cpu:
        # ONE CPU rank, as BASELINE.json configs[0] says: the oracle restatement on one thread, same problem from the seeded loader
        os.environ["OMP_NUM_THREADS"] = "1"
        from oracle import oracle as O
        O.use_timing_build()
        Pc = O.make_params(dim=2, order=kw["order"], mx0=kw["mx0"], my0=kw["my0"], ppc0=kw["ppc0"], ntimes=32, filter_kind=1)
        w = O.World(Pc)
        (w.init_weibel(ppc0=16.0, delgam=2e-5, distr_dim=2) if cfg == "c0" else w.init_twostream(ppc0=64.0, delgam=2e-5))
        npc = sum(sum(r.counts) for r in w.ranks)
        w.step()
        t0 = time.perf_counter(); k = 0
        while time.perf_counter() - t0 < 10.0:
            w.step(); k += 1
        dt = (time.perf_counter() - t0) / k
        cpu = {"value": npc / dt, "unit": "particle-steps/s", "cores": 1, "kind": "port",
               "sample": f"the whole problem ({npc} particles), {k} laps in {k * dt:.1f} s on ONE thread (one CPU rank), oracle built -O3 -march=native",
               "pin": "same C source as the parity build, which is bit-exact against the reference's own source text on 147 cases (tests/test_ref_golden.py); this -O3 -march=native build is for timing only"}
    peak, peak_src = measured_peak()
    line = {"metric": "particle-steps/sec", "value": value, "unit": "particle-steps/s", "n_gpus": 1, "steps": steps, "warmup": warm,
            "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": name, "order": kw["order"], "ppc": kw["ppc0"],
                       "l2": "small problem, resident in L2 by nature" if cfg != "c3" else "inputs larger than L2"},
            "particles": n1, "gpu_launches": launches, "clocks": sampler.summary(), "host_wall_ms_per_step": wall_ms / steps,
            "sanity": sanity, "cpu_baseline": cpu,
            "roofline": {"bound": "hbm", "achieved": value * (52.0 + 36.0 / kw["ppc0"]) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": value * (52.0 + 36.0 / kw["ppc0"]) / 1e9 / peak, "traffic": None,
                         "note": "whole lap against the mover+deposit algorithmic bytes; small configs are launch-latency bound", "peak_source": peak_src},
            "e2e": None}
    print(json.dumps(line))




# ------------------------------------------------------------------------------------------------------------
# synthetic Weibel-like state on the host: two counter-streaming beams (+-0.5c in x) with a thermal spread,
# uniform positions, small random seed fields.  Same recipe for every rank (seeded by rank).
# ------------------------------------------------------------------------------------------------------------
def make_state(tg, P, rank, pinned):
    rng = np.random.default_rng(1234 + rank)
    g, gz = P.nghost // 2, P.nghostz // 2
    nx, ny, nz = P.mx - P.nghost, P.my - P.nghost, P.mz - P.nghostz
    nhalf = int(0.5 * PPC * nx * ny * nz)                  # per species
    maxhlf = P.maxptl // 2
    assert nhalf <= maxhlf
    if pinned:
        import torch
        buf = torch.empty(P.maxptl * 40, dtype=torch.uint8, pin_memory=True)
        p = buf.numpy().view(tg.PARTICLE_DTYPE)
        keep = buf
    else:
        p = np.zeros(P.maxptl, tg.PARTICLE_DTYPE)
        keep = None
    blk = 1 << 22
    base = {k: rng.random(blk, dtype=np.float32) for k in "xyz"}
    mom = {k: (rng.standard_normal(blk).astype(np.float32) * np.float32(0.05)) for k in "uvw"}
    gam_beta = np.float32(0.5 / np.sqrt(1 - 0.25))
    for s, lo in ((0, 0), (1, maxhlf)):
        done = 0
        while done < nhalf:
            n = min(blk, nhalf - done)
            sl = slice(lo + done, lo + done + n)
            sh = rng.random(3).astype(np.float32)
            p["x"][sl] = np.float32(g + 1) + np.float32(nx) * ((base["x"][:n] + sh[0]) % np.float32(1.0))
            p["y"][sl] = np.float32(g + 1) + np.float32(ny) * ((base["y"][:n] + sh[1]) % np.float32(1.0))
            p["z"][sl] = np.float32(gz + 1) + np.float32(nz) * ((base["z"][:n] + sh[2]) % np.float32(1.0))
            sign = np.where((np.arange(n) & 1) == 0, np.float32(1), np.float32(-1))
            p["u"][sl] = sign * gam_beta + mom["u"][:n]
            p["v"][sl] = mom["v"][:n]
            p["w"][sl] = mom["w"][:n]
            p["ch"][sl] = 1.0
            p["ind"][sl] = np.arange(done + 1, done + n + 1, dtype=np.int32) * 2 - s
            p["proc"][sl] = rank
            p["splitlev"][sl] = 1
            done += n
        # keep strictly inside the interior
        for k, hi in (("x", P.mx - g), ("y", P.my - g), ("z", P.mz - gz)):
            v = p[k][lo:lo + nhalf]
            np.minimum(v, np.nextafter(np.float32(hi), np.float32(0)), out=v)
    shape = (P.mz, P.my, P.mx)
    fields = [(rng.standard_normal(shape).astype(np.float32) * np.float32(1e-3)) for _ in range(6)]
    if pinned:
        import torch
        tf = [torch.from_numpy(f).pin_memory() for f in fields]
        fields = [t.numpy() for t in tf]
        keep = (keep, tf)
    return p, nhalf, fields, keep


class ClockSampler(threading.Thread):
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def _run_nvml(self):
        """NVML in-process: a sample every 10 ms, so that even a 0.2 s timed region is covered by a dozen samples"""
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(self.index)
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        bits = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))
        while not self.stop_flag:
            r = get_reasons(h)
            self.rows.append([str(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), str(mx), "0"] +
                             ["Active" if r & b else "Not Active" for _, b in bits])
            time.sleep(0.01)

    def run(self):
        try:
            self._run_nvml()
            return
        except Exception:
            pass
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def bind_to_gpu_numa_node(local):
    """Pin this rank's host threads (and with them the first-touch placement of its pinned staging buffers) to the NUMA
    node its GPU hangs off.  With 8 ranks copying at once, buffers on the wrong socket share one inter-socket link: round 1
    measured 15-18 GB/s per GPU at N = 8 against 55 GB/s at N = 1.  Best effort: silently does nothing where sysfs does
    not tell (containers without /sys/bus/pci, single-node hosts)."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local)
        bus = "%04x:%02x:%02x.0" % (getattr(pr, "pci_domain_id", 0), pr.pci_bus_id, pr.pci_device_id)
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------------------
# CPU leg: the oracle restatement (kind "port"), one slab per OpenMP thread, bounded sample of the same workload
# ------------------------------------------------------------------------------------------------------------
def cpu_leg(cells, steps, warmup):
    """cells = (nx, ny_per_slab, nz_per_slab): one y/z slab per host thread (filter2 needs slabs at least ntimes thick)"""
    from oracle import oracle as O
    O.use_timing_build()                                   # -O3 -march=native (timing only; parity uses the -O2 no-contraction build)
    cores = os.cpu_count() or 1
    sy = sz = 1
    while sy * sz * 2 <= cores:
        if sz <= sy:
            sz *= 2
        else:
            sy *= 2
    os.environ["OMP_NUM_THREADS"] = str(sy * sz)
    cells = (cells[0], cells[1] * sy, cells[2] * sz)
    P = O.make_params(dim=3, order=ORDER, mx0=cells[0], my0=cells[1], mz0=cells[2], sizey=sy, sizez=sz, ppc0=PPC,
                      ntimes=NTIMES, filter_kind=FILTER_KIND)
    w = O.World(P)
    w.init_uniform(ppc0=PPC, beta=0.5, uth=0.05, seed=3)
    npart = sum(sum(r.counts) for r in w.ranks)
    for _ in range(warmup):
        w.step()
    t0 = time.perf_counter()
    for _ in range(steps):
        w.step()
    dt = time.perf_counter() - t0
    return npart * steps / dt, sy * sz, npart, dt / steps, cells


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # same step and warm-up counts as the GPU arm (a lap of the sample takes about a second on 16 cores); capped so that
    # an unusually long request still ends within minutes
    steps, warmup = max(1, min(args.steps, 40)), max(0, min(args.warmup, 10))
    val, cores, npart, sec, cc = cpu_leg(args.cpu_cells, steps, warmup)
    sample = f"3D Weibel dd{ORDER} {PPC:g} ppc filter2 ntimes={NTIMES}, {cc[0]}x{cc[1]}x{cc[2]} cells " \
             f"({npart} particles), {steps} laps after {warmup} warm-up, one y/z slab per OpenMP thread, {cores} threads, " \
             f"gcc -O3 -march=native"
    cfg = workload_config(args, args.gpus)
    # the workload is the GPU arm's; what this arm actually steps is a bounded sample of it, stated here and in cpu_baseline
    cfg["sample"] = f"bounded sample: {cc[0]}x{cc[1]}x{cc[2]} cells, {npart} particles (throughput is per particle-step, size-independent)"
    line = {"impl": "reference", "metric": "particle-steps/sec", "value": val, "unit": "particle-steps/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": sec * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg,
            "cpu_baseline": {"value": val, "unit": "particle-steps/s", "cores": cores, "kind": "port", "sample": sample,
                             "pin": "same C source as the parity build, which is bit-exact against the reference's own source text on 147 cases (tests/test_ref_golden.py); this -O3 -march=native build is for timing only"},
            "e2e": {"value": val, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "reference = CPU oracle restatement of the Fortran routines (no Fortran/MPI toolchain in this image)"}
    print(json.dumps(line))


def workload_config(args, n):
    sy, sz = GRID[n]
    which = "configs[4] weak-scaling sweep" if getattr(args, "config", "default") == "c4" else "configs[2] share"
    return {"workload": f"3D Weibel, dd{ORDER} (order-{ORDER} Esirkepov), {PPC:g} ppc, filter2 ntimes={NTIMES}, per-GPU slab "
                        f"{args.cells[0]}x{args.cells[1]}x{args.cells[2]} cells ({which}), global "
                        f"{args.cells[0]}x{args.cells[1] * sy}x{args.cells[2] * sz}",
            "decomposition": f"sizey={sy} sizez={sz}", "ppc": PPC, "order": ORDER, "c": 0.45,
            "l2": "inputs (>= 9 GB of particle SoA per GPU) larger than L2; no flush needed"}


def main():
    global ORDER, PPC, B_PER_PARTICLE
    args = parse()
    ORDER, PPC = args.order, args.ppc
    B_PER_PARTICLE = 52.0 + 36.0 / PPC
    if args.impl == "reference":
        return run_reference(args)
    if args.config in ("c0", "c1", "c3"):
        return run_small_config(args)
    import __graft_entry__ as ge
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    n = args.gpus
    assert n in GRID and world in (1, n), "--gpus must be 1,2,4,8 and match WORLD_SIZE under torchrun"
    import torch
    torch.cuda.set_device(local)
    numa_node = bind_to_gpu_numa_node(local) if world > 1 else None
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if rank == 0:
        ge.build()
    if dist:
        dist.barrier()
    import tristan_mp_pu_master_densdecomp_b200 as tg
    sy, sz = GRID[n] if world > 1 else (1, 1)
    cx, cy, cz = args.cells
    nhalf_est = int(0.5 * PPC * cx * cy * cz)
    P = tg.make_params(dim=3, order=ORDER, mx0=cx, my0=cy * sy, mz0=cz * sz, sizey=sy, sizez=sz, rank=rank, ntimes=NTIMES,
                       filter_kind=FILTER_KIND, ppc0=PPC, maxptl=int(2 * nhalf_est * 1.25) + 8192,
                       buffsize=max(int(nhalf_est * 0.05), 100000), device=local)
    ctx = tg.Context(P)
    ctx.set_option("fused", args.fused)
    ctx.set_option("peer", args.peer)
    if world > 1:
        ctx.comm_init_torch()
    halo_transport = ("peer memory (cudaIpc, halo kernels read the neighbours' arrays over NVLink)" if ctx.halo_transport()
                      else "NCCL send/recv") if world > 1 else "none (one rank)"
    do_e2e = not args.no_e2e
    p, nhalf, fields, keep = make_state(tg, P, rank, pinned=do_e2e)
    ctx.fields_h2d(*fields)
    ctx.particles_h2d(p, nhalf, nhalf)
    npart = 2 * nhalf
    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        ctx.step(1)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        ctx.step(1)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ctx.launch_count() - l0
    sampler.stop_flag = True
    counts = ctx.counts()
    t = torch.tensor([ms, float(sum(counts))], dtype=torch.float64, device="cuda")
    if dist:
        tm = t.clone()
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ts = t.clone()
        dist.all_reduce(ts, op=dist.ReduceOp.SUM)
        ms, total_particles = float(tm[0]), float(ts[1])
    else:
        total_particles = float(t[1])
    value = total_particles * args.steps / (ms * 1e-3)

    # per-phase device time (CUDA events on the library's stream around each phase), 3 extra laps
    ctx.set_option("timing", 1)
    ctx.timers(reset=True)
    nphase = 3
    for _ in range(nphase):
        ctx.step(1)
    ph = {k: v / nphase for k, v in ctx.timers(reset=True).items()}
    ctx.set_option("timing", 0)
    peak, peak_src = measured_peak()
    md_ms = ph["mover"] + ph["deposit"]
    alg_bytes = sum(counts) * B_PER_PARTICLE
    achieved = alg_bytes / (md_ms * 1e-3) / 1e9 if md_ms > 0 else 0.0
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "r2_traffic.json")) as f:
            tj = json.load(f)
        if list(args.cells) == [512, 256, 128] and args.fused and ORDER == 2 and PPC == 16.0:
            traffic = 2 * tj["traffic_per_launch"]            # two launches (ions, electrons) per lap
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": (f"k_cellrun<{ORDER},fused,lazy> " if ORDER < 3 else "k_cellrun3<fused,lazy> ") +
                "(gather + Boris push + Esirkepov deposit + sort keys + lazy-sort gather) + k_primal + k_add_shadow_tiled: "
                "the mover and deposit phases of one lap (two cell-run launches, ions and electrons)",
                "algorithmic_bytes_per_launch": alg_bytes, "ms_per_launch": md_ms, "peak_source": peak_src,
                "phase_ms": ph}

    e2e = None
    if do_e2e:
        # mirror mode: the host owns the state; every lap the full state crosses PCIe in both directions
        outp = p
        barrier()
        t0 = time.perf_counter()
        ni, ne = nhalf, nhalf
        ci, ce = ctx.counts()
        # untimed warm-up: the host copy becomes what a mirror-mode host would hold (the state the device returned)
        _, ci, ce = ctx.particles_d2h(outp)
        ctx.fields_d2h(fields)
        barrier()
        t0 = time.perf_counter()
        parts = {"h2d": 0.0, "lap": 0.0, "d2h": 0.0}
        for _ in range(args.e2e_steps):
            ta = time.perf_counter()
            if args.e2e_plain:
                # the call-for-call sequence: upload everything, run the lap, download everything
                ctx.fields_h2d(*fields)
                ctx.particles_h2d(outp, ci, ce)
                tb = time.perf_counter()
                ctx.step(1)
                ctx.counts()
                torch.cuda.synchronize()
                tc = time.perf_counter()
                _, ci, ce = ctx.particles_d2h(outp)
                ctx.fields_d2h(fields)
                td = time.perf_counter()
                parts["h2d"] += tb - ta; parts["lap"] += tc - tb; parts["d2h"] += td - tc
            else:
                # tgpu_step_mirror: the same lap with host arrays in and out; on one periodic rank the particle records are
                # streamed through the fused mover with both PCIe directions busy (elsewhere it is the sequence above)
                ci, ce = ctx.step_mirror(fields, outp, ci, ce)
                parts["lap"] += time.perf_counter() - ta
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / args.e2e_steps
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if dist:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt[0])
        fbytes = 6 * P.mx * P.my * P.mz * 4
        pbytes = (ci + ce) * 40
        # what the link itself gives (1 GiB pinned copies), to read the mirror number against
        probe = torch.empty(1 << 30, dtype=torch.uint8, pin_memory=True)
        dev = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
        link = {}
        for name, (dst, src) in (("h2d", (dev, probe)), ("d2h", (probe, dev))):
            dst.copy_(src, non_blocking=True); torch.cuda.synchronize()
            t1 = time.perf_counter(); dst.copy_(src, non_blocking=True); torch.cuda.synchronize()
            link[name] = (1 << 30) / (time.perf_counter() - t1) / 1e9
        del probe, dev
        e2e = {"value": total_particles / dt, "unit": "particle-steps/s", "h2d_bytes_per_step": fbytes + pbytes,
               "pcie_gbs_measured": link, "host_numa_node": numa_node,
               "d2h_bytes_per_step": fbytes + pbytes,
               "mode": ("mirror: fields+particles H2D, one lap, fields+particles D2H as separate tgpu_* calls, pinned host buffers"
                        if args.e2e_plain else
                        "mirror: tgpu_step_mirror, host arrays in and out every lap (pinned); " +
                        ("particle records streamed through the fused mover, both PCIe directions busy" if world == 1 else
                         "plain h2d + lap + d2h sequence (streaming needs a single periodic rank)")),
               "ms_per_step": dt * 1e3,
               "ms_parts": {k: v / args.e2e_steps * 1e3 for k, v in parts.items()}}
    ctx.close()
    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return
    cpu = None
    if not args.no_cpu and world == 1:
        val, cores, npc, sec, cc = cpu_leg(args.cpu_cells, 2, 1)
        cpu = {"value": val, "unit": "particle-steps/s", "cores": cores, "kind": "port",
               "sample": f"{cc[0]}x{cc[1]}x{cc[2]} cells, {npc} particles, 2 laps after 1 warm-up, same "
                         f"physics (dd{ORDER}, {PPC:g} ppc, filter2 ntimes={NTIMES}); oracle restatement built -O3 -march=native, "
                         f"one y/z slab per OpenMP thread, {cores} threads",
               "pin": "same C source as the parity build, which is bit-exact against the reference's own source text on 147 cases (tests/test_ref_golden.py); this -O3 -march=native build is for timing only"}
    line = {"metric": "particle-steps/sec", "value": value, "unit": "particle-steps/s", "n_gpus": n, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "ns_per_particle_step": 1e9 / value * n,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(args, n), halo_transport=halo_transport), "particles": total_particles, "gpu_launches": launches,
            "clocks": sampler.summary(), "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
            "hbm_roofline_frac_whole_step": (total_particles / n * B_PER_PARTICLE + 144.0 * cx * cy * cz) / (ms / args.steps * 1e-3) / 1e9 / peak}
    print(json.dumps(line))
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
