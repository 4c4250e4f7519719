#!/usr/bin/env python
"""bench.py -- osinco3d Chorin-projection time step on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--n 256] [--bc freeslip|periodic]
                    [--les] [--impl reference] [--legs all|none|name,...] [--strong]

One "step" = predict_velocity + correct_pression (divergence + Poisson solve) + correct_velocity
[+ transeq], AB3, on synthetic flow fields.  The HEADLINE workload (`value`,
`roofline`, `e2e`, `cpu_baseline`) is BASELINE.json configs[1]: TGV Re=1600 DNS at 256^3 on one
B200 with the shipped example's settings (free-slip on [0,pi]^3, dt = 0.05 dx, omega = 1.887,
eps = 1e-4, idyn = 0).  N > 1 (torchrun): z-slab decomposition, weak scaling -- the box is
replicated in z (nz = N (n-1) + 1 planes on [0, N pi], still an exact free-slip TGV).

The same JSON line carries, under `configs`, the other configurations of BASELINE.json / the north
star, each with its own value, Poisson iterations per step K and per-stage roofline:
  N = 1 : 512^3 DNS (north-star target), 512^3 LES (configs[2]), periodic 256^3 (K ~ 18) and 257^3
          (odd extents: seam SOR), coplanar jet 257 x 513 x 129 (configs[4], K ~ 11), mixing layer 241 x 241 x
          81 LES + scalar with SOR (K ~ 84) and with multigrid (configs[3])
  N > 1 : 512 x 512 x 511 planes per GPU DNS (N = 8: the 1024^3 class), 512^3 LES cut into N z
          slabs (configs[2]: the N = 1 leg strong-scaled), periodic 256 x 256 x 256 N
          (K ~ 18, wrap link rank 0 <-> N-1) and the coplanar jet replicated in z (odd x / y
          extents: seam classes across slabs)
and a `parity` object: N = 1 -- a 64^3 side problem against the CPU oracle; N > 1 -- the N-rank
fields against the same steps on ONE GPU (rank 0), compared bitwise through digests.

Printed JSON (rank 0, one line):
  value        Mpts*steps/s, state resident in HBM, device time (CUDA events on the session
               stream), max over ranks
  e2e          same metric through the reference-facing host-pointer module procedures
               (o3d_predict_velocity / o3d_correct_pression / o3d_correct_velocity with HOST
               arrays, pinned): every H2D / D2H copy is inside the timed region
  roofline     the dominant kernel (largest share of device time): algorithmic bytes per launch
               / mean launch duration, against the measured HBM copy peak
  cpu_baseline the CPU oracle (C restatement of the reference; the Fortran reference cannot be
               compiled in this image) timed on a bounded sample, 1 thread -- the reference is
               serial
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PI = 3.141592653589793

# algorithmic bytes per grid point and launch (SURVEY.md 8d / DESIGN.md "Kernels")
B_RHS = {1: 72.0, 2: 96.0, 3: 120.0}     # Euler / AB2 / AB3 (+8 with LES nu_t)
B_DIV = 32.0
B_SOR_HALF = 16.0     # in-place colour half-sweep: half of 32 B/pt/iteration
B_SOR_FUSED = 24.0    # fused red+black pass with ping-pong: read pp + read rhs + write pp
B_CORR = 56.0
B_TRANSEQ = 88.0
# --impl reference: wall-clock budget of the whole run and the port's nominal throughput, used
# only to decide whether K + W steps of the full grid fit (else a smaller sample grid is used)
REF_BUDGET_S = 150.0
REF_RATE_PTS_S = 6.0e6
# z chunks of the e2e leg (csrc/pipeline.cu: the library default); O3D_PIPELINE overrides
try:
    E2E_PIPELINE_CHUNKS = max(0, int(os.environ.get("O3D_PIPELINE", "16")))
except ValueError:
    E2E_PIPELINE_CHUNKS = 16


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--n", "--grid", dest="n", type=int, default=256)
    ap.add_argument("--bc", default="freeslip", choices=["freeslip", "periodic"])
    ap.add_argument("--les", action="store_true")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-n", type=int, default=0, help="grid of the CPU sample (0 = same n)")
    ap.add_argument("--pipeline", type=int, default=E2E_PIPELINE_CHUNKS,
                    help="z chunks of the pipelined host-pointer procedures in the e2e leg "
                         "(o3d_set_pipeline; 0 = unpipelined)")
    ap.add_argument("--strong", action="store_true",
                    help="N > 1: keep the grid at n^3 (strong scaling) instead of replicating "
                         "the box in z")
    ap.add_argument("--legs", default=None,
                    help="extra configurations reported under `configs`: all | none | comma list "
                         "(default: all for the default headline workload, none otherwise)")
    ap.add_argument("--no-parity", action="store_true")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------
# workloads: grid, closures, physics and a synthetic initial state per z slab (numpy restatements
# of the reference's initial profiles, src/initial_conditions.f90 -- input generation only)
# ----------------------------------------------------------------------------------------------
def tgv_fields(nx, ny, nz_local, z0, d, perturb=False):
    """TGV (src/initial_conditions.f90:141-153) for planes [z0, z0+nz_local)"""
    x = (d[0] * np.arange(nx))[:, None, None]
    y = (d[1] * np.arange(ny))[None, :, None]
    z = (d[2] * np.arange(z0, z0 + nz_local))[None, None, :]
    ux = np.sin(x) * np.cos(y) * np.cos(z)
    uy = -np.cos(x) * np.sin(y) * np.cos(z)
    uz = np.zeros((nx, ny, nz_local))
    pp = 0.0625 * (np.cos(2 * x) + np.cos(2 * y)) * (np.cos(2 * z) + 2.0)
    if perturb:   # parity legs: break the symmetries and make uz != 0 so every halo plane matters
        ux = ux + 0.1 * np.sin(2 * x) * np.cos(3 * y) * np.cos(2 * z)
        uy = uy + 0.05 * np.cos(x) * np.sin(2 * y) * np.cos(3 * z)
        uz = uz + 0.2 * np.cos(x) * np.cos(y) * np.sin(z)
    f = np.asfortranarray
    return {"ux": f(ux), "uy": f(uy), "uz": f(uz), "pp": f(pp)}


def tgv_slab(n, nz_local, z0, d):
    w = tgv_fields(n, n, nz_local, z0, (d, d, d))
    return w["ux"], w["uy"], w["uz"], w["pp"]


def _u_base(profile, dy, base):
    """calcul_u_base + normalize1D (src/utils.f90:9-45); dery1D replaced by numpy.gradient"""
    g = np.gradient(profile, dy)
    lo, hi = g.min(), g.max()
    return np.full_like(g, base) if hi - lo < 1e-12 else base + (1.0 - base) * (g - lo) / (hi - lo)


def shear_fields(kind, nx, ny, nz_local, z0, d, origin_y, xlx):
    """mixing layer (typesim 5, :366-391) / coplanar jet (typesim 4, :284-323) profiles plus the
    deterministic ici = 1 oscillations (:554-629); fields do not depend on z"""
    x = d[0] * np.arange(nx)
    y = origin_y + d[1] * np.arange(ny)
    if kind == "mixing_layer":
        t3 = y * np.log(2.0) / (2.0 / 13.0)
        prof = 0.5 * np.tanh(t3)                    # u0 = 1, ratio = -1
        phi = 0.5 - 0.5 * np.tanh(t3)
        ub = _u_base(prof, d[1], 0.0)
        sx = np.sin(8 * PI * x / xlx) + np.sin(4 * PI * x / xlx) / 8 + np.sin(2 * PI * x / xlx) / 16
        cxs = np.cos(8 * PI * x / xlx) + np.cos(4 * PI * x / xlx) / 8 + np.cos(2 * PI * x / xlx) / 16
        ux2 = prof[None, :] + 0.03 * ub[None, :] * sx[:, None]
        uy2 = 0.12 * ub[None, :] * cxs[:, None]
        uz2 = np.zeros((nx, ny))
        pp0 = 1.0
    else:
        u2, u1, u3 = 1.0, 1.0 / 3.0, 0.0            # u0 = 1, ratio = 3
        h1, h2 = 0.5, 1.0
        th1, th2 = h1 / 10.0, h2 / 25.0
        ay = np.abs(y)
        inner = 0.5 * (u1 + u2) + 0.5 * (u2 - u1) * np.tanh((ay - h1) / (2 * th1))
        outer = 0.5 * (u2 + u3) + 0.5 * (u3 - u2) * np.tanh((ay - h2) / (2 * th2))
        prof = np.where(ay < 0.5 * (h1 + h2), inner, outer)
        phi = None
        ub = _u_base(prof, d[1], -1.0)
        s9 = np.sin(2 * PI * 9 * x / xlx)
        ux2 = prof[None, :] + 0.03 * ub[None, :] * s9[:, None]
        uy2 = 0.03 * ub[None, :] * s9[:, None]
        uz2 = 0.03 * ub[None, :] * s9[:, None]
        pp0 = 0.0
    one = np.ones((1, 1, nz_local))
    f = np.asfortranarray
    out = {"ux": f(ux2[:, :, None] * one), "uy": f(uy2[:, :, None] * one),
           "uz": f(uz2[:, :, None] * one), "pp": f(np.full((nx, ny, nz_local), pp0))}
    if phi is not None:
        out["phi"] = f(np.ones((nx, 1, 1)) * phi[None, :, None] * one)
    return out


def make_workload(kind, nranks=1, n=256, bc="freeslip", les=False, strong=False, multigrid=0,
                  perturb=False):
    """-> dict(name, grid, d, bc, phys, nscr, multigrid, init(z0, nzl))"""
    if kind == "tgv":
        b = (1, 1, 1) if bc == "freeslip" else (0, 0, 0)
        L = PI if bc == "freeslip" else 2 * PI
        dd = L / (n - 1)                      # dx = xlx/(nx-1), src/initialization.f90:182-184
        # weak scaling: the box replicated in z -- free-slip: N (n-1) + 1 planes on [0, N pi];
        # periodic: N n planes (the reference's period is n dx, SURVEY finding 5)
        nz = n
        if nranks > 1 and not strong:
            nz = nranks * (n - 1) + 1 if bc == "freeslip" else nranks * n
        if les:    # examples/tgv_re2500_les, dt scaled with dx from 5e-4 @ 129^3
            phys = dict(re=2500.0, dt=5e-4 * 128.0 / (n - 1), omega=1.999, eps=1e-6, idyn=1,
                        iles=1, cs=0.17, kmax=10000)
        else:      # examples/tgv_re1600_dns (dt = cfl*dx/u0)
            phys = dict(re=1600.0, dt=0.05 * dd, omega=1.887, eps=1e-4, idyn=0, iles=0, cs=0.0,
                        kmax=10000)
        name = "tgv_re%d_%s_%s_%dx%dx%d_ab3_sor" % (int(phys["re"]), "les" if les else "dns", bc,
                                                     n, n, nz)
        d3 = (dd, dd, dd)
        return dict(name=name, grid=(n, n, nz), d=d3, bc=b, phys=phys, nscr=0, multigrid=0, n=n,
                    init=lambda z0, nzl: tgv_fields(n, n, nzl, z0, d3, perturb))
    if kind == "mixing_layer":   # examples/mixing_layer_re3000_les, ici = 1 instead of ici = 2
        nx, ny, nz = 241, 241, 81
        d3 = (12.0 / (nx - 1), 12.0 / (ny - 1), 4.0 / (nz - 1))
        phys = dict(re=3000.0, dt=1.5e-3, omega=1.999, eps=1e-5, idyn=1, iles=1, cs=0.15, kmax=5000)
        name = "mixing_layer_re3000_les_scalar_%dx%dx%d_ab3_%s" % (
            nx, ny, nz, "multigrid" if multigrid else "sor")
        return dict(name=name, grid=(nx, ny, nz), d=d3, bc=(0, 1, 0), phys=phys, nscr=1,
                    multigrid=multigrid, n=nx,
                    init=lambda z0, nzl: shear_fields("mixing_layer", nx, ny, nzl, z0, d3, -6.0,
                                                      12.0))
    if kind == "cojet":          # examples/coplanar_jet_re2200 (all periodic, odd extents)
        nx, ny, nz1 = 257, 513, 129
        d3 = (5.5 / (nx - 1), 11.0 / (ny - 1), 2.75 / (nz1 - 1))
        nz = nz1 * max(1, nranks)             # weak scaling: the box replicated in z
        phys = dict(re=2200.0, dt=0.07 * min(d3), omega=1.35, eps=1e-5, idyn=0, iles=0, cs=0.0,
                    kmax=1000)
        name = "coplanar_jet_re2200_periodic_%dx%dx%d_ab3_sor" % (nx, ny, nz)
        return dict(name=name, grid=(nx, ny, nz), d=d3, bc=(0, 0, 0), phys=phys, nscr=0,
                    multigrid=0, n=nx,
                    init=lambda z0, nzl: shear_fields("cojet", nx, ny, nzl, z0, d3, -5.5, 5.5))
    raise ValueError(kind)


def leg_specs(world):
    """the `configs` legs of the JSON line: (key, make_workload arguments)"""
    if world == 1:
        return [("tgv512_dns", dict(kind="tgv", n=512)),
                ("tgv512_les", dict(kind="tgv", n=512, les=True)),
                ("tgv256_periodic", dict(kind="tgv", n=256, bc="periodic")),
                ("tgv257_periodic", dict(kind="tgv", n=257, bc="periodic")),
                ("cojet", dict(kind="cojet")),
                ("mixing_layer_sor", dict(kind="mixing_layer")),
                ("mixing_layer_multigrid", dict(kind="mixing_layer", multigrid=1))]
    return [("tgv512_dns", dict(kind="tgv", n=512)),
            # BASELINE configs[2], "LES at 512^3, 1/2/4/8 GPUs": the SAME 512^3 grid as the N = 1
            # leg `tgv512_les`, cut into z slabs (strong scaling; same workload name)
            ("tgv512_les_strong", dict(kind="tgv", n=512, les=True, strong=True)),
            ("tgv256_periodic", dict(kind="tgv", n=256, bc="periodic")),
            ("cojet", dict(kind="cojet"))]


def workload(args, nranks):
    """the headline workload from the command line (kept for --impl reference and the tests)"""
    w = make_workload("tgv", nranks, args.n, args.bc, args.les, args.strong)
    return dict(w, nz=w["grid"][2], d=w["d"][0])


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index),
                                       "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            c = [t.strip() for t in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                smax.append(float(c[2]))
                power.append(float(c[3]))
            except ValueError:
                continue
            for nm, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        self.f.close()
        os.unlink(self.f.name)
        if sm:
            # "under load": drop the idle samples before/after the region
            load = [s for s, p in zip(sm, power) if p >= 0.5 * max(power)] or sm
            out = {"sm_mhz": float(np.median(load)), "sm_max_mhz": float(max(smax)),
                   "power_w_max": float(max(power)), "reasons": sorted(reasons),
                   "samples": len(sm)}
        return out


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(kernel, grid):
    """per-launch DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of a kernel from the
    committed ncu --set full summaries (profiles/traffic.json), keyed by kernel AND grid: a capture
    at another size says nothing about this launch, so anything else is null"""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            t = json.load(open(p))
            return t.get("%s@%dx%dx%d" % ((kernel,) + tuple(grid)))
        except Exception:
            return None
    return None


# ----------------------------------------------------------------------------------------------
# CPU baseline: the oracle (C restatement of the reference's serial Fortran), 1 thread
# ----------------------------------------------------------------------------------------------
def cpu_sample(args, steps, n, opt=False):
    from oracle import oracle_py as O      # checker / baseline only -- never the product path
    O.build()
    bc = (1, 1, 1) if args.bc == "freeslip" else (0, 0, 0)
    L = PI if args.bc == "freeslip" else 2 * PI
    d = L / (n - 1)
    w = workload(args, 1)
    ph = dict(w["phys"])
    if not args.les:
        ph["dt"] = 0.05 * d
    else:
        ph["dt"] = 5e-4 * 128.0 / (n - 1)
    if opt:
        O.use_fast_build(True)
    try:
        g = O.grid(n, n, n, d, d, d, bc)
        ux, uy, uz, pp, _ = O.init_tgv(g)
        sim = O.Sim(g, re=ph["re"], dt=ph["dt"], itscheme=3, iles=ph["iles"], cs=ph["cs"],
                    omega=ph["omega"], eps=ph["eps"], kmax=10000, idyn=ph["idyn"])
        sim.set(ux=ux, uy=uy, uz=uz, pp=pp)
        times, iters = [], []
        for _ in range(steps):
            t0 = time.perf_counter()
            iters.append(sim.step())
            times.append(time.perf_counter() - t0)
        sim.close()
    finally:
        if opt:
            O.use_fast_build(False)
    return times, iters


def guarded(fn):
    """a secondary single-rank leg (e2e, parity, cpu_baseline) that fails is reported in its own
    object; it does not take the headline line down with it.  Never used around collectives."""
    try:
        return fn()
    except Exception as e:
        import traceback
        traceback.print_exc()
        return {"error": "%s: %s" % (type(e).__name__, e)}


def cpu_baseline(args, n):
    ncpu = args.cpu_n or n
    nsteps = 5 if ncpu >= 200 else 8      # ~10-15 s of CPU work at 256^3
    times, its = cpu_sample(args, nsteps, ncpu, opt=True)
    tt = times[1:] if len(times) > 1 else times
    ms = 1e3 * sum(tt) / len(tt)
    t2, _ = cpu_sample(args, 3, ncpu, opt=False)
    return {
        "value": (ncpu ** 3) / 1e6 / (ms / 1e3), "unit": "Mpts*steps/s", "cores": 1,
        "kind": "port", "build": "gcc -O3 -ffp-contract=off (the reference builds -O3, "
                                 "src/Makefile:15)",
        "value_strict_O2_build": (ncpu ** 3) / 1e6 / (sum(t2[1:]) / len(t2[1:])),
        "sample": "%d steps of the same workload at %d^3 (first step dropped), SOR "
                  "iters/step %s; oracle = C restatement of the serial Fortran reference "
                  "(gfortran absent)" % (len(tt), ncpu, its[1:] if len(its) > 1 else its)}


def oracle_parity(o3d, n=64, steps=3):
    """N = 1 `parity`: a small side problem (free-slip TGV + perturbation, LES + scalar, shipped-like
    Poisson settings) through the SAME library and session API that was just timed, against the CPU
    oracle: bitwise with the reference's sweep order (LEXI_WAVEFRONT), to solver tolerance with the
    red-black fast path.  The oracle is the checker here, nothing it computes is timed."""
    from oracle import oracle_py as O
    O.build()
    d = PI / (n - 1)
    g = O.grid(n, n, n, d, d, d, (1, 1, 1))
    f = tgv_fields(n, n, n, 0, (d, d, d), perturb=True)
    phi = np.asfortranarray(0.5 + 0.4 * np.cos(f["ux"]))
    kw = dict(re=1600.0, dt=0.05 * d, omega=1.887, eps=1e-6, kmax=2000, idyn=0)
    sim = O.Sim(g, itscheme=3, iles=1, cs=0.17, nscr=1, **kw)
    sim.set(phi=phi, **f)
    it_o = [sim.step() for _ in range(steps)]
    out = {"vs": "cpu oracle (oracle/o3d_oracle.c)", "grid": [n, n, n], "steps": steps,
           "fields": ["ux", "uy", "uz", "pp", "phi"]}
    for order, key in ((o3d.SOR_LEXI_WAVEFRONT, "wavefront"), (o3d.SOR_RED_BLACK, "red_black")):
        cfg = o3d.make_config(n, n, n, d, d, d, bc=(1, 1, 1), itscheme=3, iles=1, cs=0.17, nscr=1,
                              sor_order=order, **kw)
        ses = o3d.Session(cfg)
        ses.set(phi=phi, **f)
        it_g = [ses.step() for _ in range(steps)]
        scale = max(float(np.max(np.abs(sim.field(k)))) for k in ("ux", "uy", "uz"))
        err = max(float(np.max(np.abs(ses.download(k) - sim.field(k)))) for k in ("ux", "uy", "uz"))
        if key == "wavefront":
            out["bitwise_equal"] = bool(
                it_g == it_o and all(np.array_equal(ses.download(k), sim.field(k))
                                     for k in ("ux", "uy", "uz", "pp")) and
                float(np.max(np.abs(ses.download("phi") - sim.field("phi")))) < 1e-13)
            out["sor_iters"] = {"oracle": it_o, "gpu_wavefront": it_g}
        else:
            out["red_black_max_rel_velocity_error"] = err / scale
            out["sor_iters"]["gpu_red_black"] = it_g
        ses.close()
    sim.close()
    return out


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  The Fortran
    source cannot be compiled here (no Fortran compiler / FFTW3 in the image), so this times the
    oracle port, single thread: the reference is serial code.  The timed build is
    `gcc -O3 -ffp-contract=off` (the reference's own optimisation level, src/Makefile:15); the
    strict -O2 parity build is timed beside it on a few steps and reported in `cpu_baseline`.

    Exactly W warm-up and K timed steps are run.  One step = one full time step of the workload;
    when K + W steps of the full grid would not end within a few minutes (REF_BUDGET_S at the
    port's ~7 Mpts*steps/s), each step is a bounded sample instead: the same workload on a
    smaller grid (the metric is per grid point, so it carries over), named in `sample`."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    K, W = max(1, args.steps), max(0, args.warmup)
    n_full = args.n
    n = args.cpu_n or n_full
    if not args.cpu_n:
        per_step = n ** 3 / REF_RATE_PTS_S
        if (K + W) * per_step > REF_BUDGET_S:
            n = max(32, int((REF_BUDGET_S / (K + W) * REF_RATE_PTS_S) ** (1.0 / 3.0)))
            n = min(n, n_full)
    # the arm is launched like the b200 arm: at N > 1 its workload is the N-rank grid (weak scaling:
    # the n^3 box replicated N times in z; --strong: the n^3 box itself), and each step times a
    # bounded sample of it -- ONE box (or a smaller one) -- since the metric is per grid point
    world = max(int(os.environ.get("WORLD_SIZE", "1")), int(args.gpus or 1))
    w = workload(args, 1)
    w_full = workload(args, world)
    ph = dict(w["phys"])
    if n != n_full:    # the bounded sample keeps the CFL number: dt follows the grid (cpu_sample)
        L = PI if args.bc == "freeslip" else 2 * PI
        ph["dt"] = 0.05 * L / (n - 1) if not args.les else 5e-4 * 128.0 / (n - 1)
    times, iters = cpu_sample(args, W + K, n, opt=True)
    t = times[W:]
    ms = 1e3 * sum(t) / len(t)
    val = (n ** 3) / 1e6 / (ms / 1e3)
    t2, _ = cpu_sample(args, 3, n, opt=False)
    val_o2 = (n ** 3) / 1e6 / (sum(t2[1:]) / len(t2[1:]))
    sample = ("%d timed steps (after %d warm-up) of %s at %d^3%s%s, SOR iters/step %s"
              % (len(t), W, w["name"], n,
                 "" if n == n_full else " (bounded sample of the %d^3 workload)" % n_full,
                 "" if w_full["name"] == w["name"] else
                 " = one of the %d z replicas of the weak-scaled grid %s" % (world, w_full["name"]),
                 iters[W:]))
    line = {"impl": "reference", "metric": "Mpts*steps/s (full AB3+SOR time step)",
            "value": val, "unit": "Mpts*steps/s", "n_gpus": args.gpus, "steps": len(t),
            "warmup": W, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            # the keys of the b200 arm's `config`, with this arm's own values where they differ
            "config": {"workload": w_full["name"], "grid": [n, n, n],
                       "workload_grid": list(w_full["grid"]), "dt": ph["dt"], "re": ph["re"],
                       "omega": ph["omega"], "eps": ph["eps"], "idyn": ph["idyn"],
                       "iles": ph["iles"], "sor_order": "lexicographic (src/poisson.f90)",
                       "sor_iters_per_step": float(np.mean(iters[W:])),
                       "sor_path": {"persistent": False, "peer_memory": False},
                       "parallelism": "serial cpu x1",
                       "l2": "host arrays (18 fields x %.0f MB), no device" % (n ** 3 * 8 / 1e6),
                       "host": "cpu"},
            "cpu_baseline": {"value": val, "unit": "Mpts*steps/s", "cores": 1, "kind": "port",
                             "sample": sample, "build": "gcc -O3 -ffp-contract=off",
                             "value_strict_O2_build": val_o2,
                             "note": "C restatement of the reference (oracle/), not the gfortran "
                                     "build; the reference is single-threaded"},
            "e2e": {"value": val, "unit": "Mpts*steps/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------
# end-to-end through the host-pointer module procedures
# ----------------------------------------------------------------------------------------------
def run_e2e(o3d, w, steps, warmup, chunks=0):
    from osinco3d_b200 import modules as M
    chunks_before = M.get_pipeline()
    M.set_pipeline(chunks)
    n, d, bc, ph = w["n"], w["d"][0], w["bc"], w["phys"]
    shape = (n, n, n)
    N = n ** 3
    pool = o3d.PinnedPool()
    ux0, uy0, uz0, pp0 = tgv_slab(n, n, 0, d)
    u = [pool.array(a) for a in (ux0, uy0, uz0)]
    pp = pool.array(pp0)
    up = [pool.empty(shape) for _ in range(3)]
    nu_t = pool.empty(shape)
    f = [pool.empty(shape + (3,)) for _ in range(3)]
    for a in f:
        a[...] = 0.0
    M.schemes(bc[0], bc[0], bc[1], bc[1], bc[2], bc[2])
    adt, bdt, cdt = M.ab_coefficients(ph["dt"])
    delta = (d * d * d) ** (1.0 / 3.0)
    lib = o3d.lib()
    import ctypes as C
    dp = o3d._lib.dp

    def P(a):
        return a.ctypes.data_as(dp)

    v3 = lambda v: (C.c_double * 3)(*v)  # noqa: E731
    omega = C.c_double(ph["omega"])
    it, dmax = C.c_int(0), C.c_double(0.0)
    A, B, Cc = v3(adt), v3(bdt), v3(cdt)
    cd = C.c_double

    call_s = [0.0, 0.0, 0.0]     # wall time inside each of the three (synchronous) procedures

    def step(itime):
        t = [time.perf_counter()]
        o3d._lib.check(lib.o3d_predict_velocity(
            P(up[0]), P(up[1]), P(up[2]), P(u[0]), P(u[1]), P(u[2]), P(f[0]), P(f[1]), P(f[2]),
            cd(ph["re"]), A, B, Cc, itime, 3, cd(d), cd(d), cd(d), n, n, n, ph["iles"],
            cd(ph["cs"]), cd(delta), P(nu_t)))
        t.append(time.perf_counter())
        o3d._lib.check(lib.o3d_correct_pression(
            P(pp), P(up[0]), P(up[1]), P(up[2]), cd(d), cd(d), cd(d), n, n, n, cd(ph["dt"]),
            C.byref(omega), cd(ph["eps"]), 10000, ph["idyn"], 0, C.byref(it), C.byref(dmax)))
        t.append(time.perf_counter())
        o3d._lib.check(lib.o3d_correct_velocity(
            P(u[0]), P(u[1]), P(u[2]), P(up[0]), P(up[1]), P(up[2]), P(pp), cd(ph["dt"]), cd(d),
            cd(d), cd(d), n, n, n))
        t.append(time.perf_counter())
        for q in range(3):
            call_s[q] += t[q + 1] - t[q]
        return it.value

    itime = 0
    for _ in range(warmup):
        itime += 1
        step(itime)
    call_s[:] = [0.0, 0.0, 0.0]
    t0 = time.perf_counter()
    iters = []
    for _ in range(steps):
        itime += 1
        iters.append(step(itime))
    dt_wall = time.perf_counter() - t0      # every call returns after its D2H copies completed
    pool.close()
    M.set_pipeline(chunks_before)
    # bytes per step, counted from the arrays the three calls copy (see modules.cu)
    hostshift = bool(chunks) and bool(M.get_hostshift())
    h2d, d2h = M.e2e_bytes_per_step(N, ph["iles"])
    ms = 1e3 * dt_wall / steps
    return {"value": N / 1e6 / (ms / 1e3), "unit": "Mpts*steps/s", "ms_per_step": ms,
            "steps": steps, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
            "sor_iters_per_step": float(np.mean(iters)), "pipeline_chunks": chunks,
            "host_shift": hostshift,
            "ms_per_call": {"predict_velocity": 1e3 * call_s[0] / steps,
                            "correct_pression": 1e3 * call_s[1] / steps,
                            "correct_velocity": 1e3 * call_s[2] / steps},
            "path": "o3d_predict_velocity + o3d_correct_pression + o3d_correct_velocity with "
                    "pinned HOST arrays (stateless drop-in procedures), wall clock"
                    + ("; predict / correct_velocity pipelined over %d z chunks (upload, kernel "
                       "and download of successive chunks overlap)" % chunks if chunks else "")
                    + ("; history levels 2 and 3 and the DNS nu_t = 0 are made in the host "
                       "arrays by worker threads instead of being downloaded "
                       "(src/integration.f90:176-188)" if hostshift else "")}


def run_e2e_resident(o3d, ses, steps):
    """the production integration mode (`o3d_resident = .true.`): state
    stays in HBM, and every step mirrors ux, uy, uz back into pinned host arrays for the driver's
    prints / output.  Reported next to `e2e`; it has no per-step H2D, so it is NOT the e2e line."""
    pool = o3d.PinnedPool()
    host = [pool.empty(ses.shape) for _ in range(3)]
    ses.sync()
    t0 = time.perf_counter()
    for _ in range(steps):
        ses.step()
        for a, name in zip(host, ("ux", "uy", "uz")):
            ses.download_ptr(name, a.ctypes.data)
    dt_wall = time.perf_counter() - t0
    pool.close()
    n = float(np.prod(ses.shape))
    ms = 1e3 * dt_wall / steps
    # the loop a resident driver actually runs: the step plus all the
    # per-step prints of src/osinco3d_main.f90:116-128 as device reductions, nothing mirrored
    ses.sync()
    t0 = time.perf_counter()
    for _ in range(steps):
        ses.step()
        ses.step_diagnostics()
    ms_diag = 1e3 * (time.perf_counter() - t0) / steps
    return {"value": n / 1e6 / (ms / 1e3), "unit": "Mpts*steps/s", "ms_per_step": ms,
            "steps": steps, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": int(3 * n * 8),
            "path": "Session.step() + D2H mirror of ux, uy, uz into pinned host arrays every "
                    "step (resident integration mode), wall clock",
            "with_device_diagnostics": {
                "value": n / 1e6 / (ms_diag / 1e3), "unit": "Mpts*steps/s", "ms_per_step": ms_diag,
                "path": "Session.step() + o3d_s_step_diagnostics (divergence statistics of u* and "
                        "u, min/max, CFL on the device; 23 doubles D2H), wall clock"}}


def run_e2e_slabs(o3d, ses, steps, dist):
    """N > 1: the host-pointer module procedures are single-device (the reference has no domain
    decomposition to drop into), so the end-to-end figure of a z-slab run goes through the public
    session API instead: every step each rank uploads its slab of the flow state (ux, uy, uz, pp)
    from pinned host memory, runs o3d_step, and reads the new state back into pinned host memory.
    Wall clock between barriers, max over ranks."""
    import torch
    pool = o3d.PinnedPool()
    names = ("ux", "uy", "uz", "pp")
    host = {k: pool.empty(ses.shape) for k in names}
    for k in names:
        ses.download_ptr(k, host[k].ctypes.data)
    ses.sync()
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        for k in names:
            ses.upload_ptr(k, host[k].ctypes.data)
        ses.step()
        for k in names:
            ses.download_ptr(k, host[k].ctypes.data)
    ses.sync()
    dt_wall = time.perf_counter() - t0
    t = torch.tensor([dt_wall], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    pool.close()
    nloc = float(np.prod(ses.shape))
    return float(t.item()) / steps, int(4 * nloc * 8)


# ----------------------------------------------------------------------------------------------
# one device-resident leg: session, warm-up, timed steps, per-stage roofline
# ----------------------------------------------------------------------------------------------
class Env:
    """process-wide context of a run: library, rank layout, torch.distributed"""

    def __init__(self, o3d, rank, world, local, dist):
        self.o3d, self.rank, self.world, self.local, self.dist = o3d, rank, world, local, dist

    def nccl_id(self):
        if self.world <= 1:
            return None
        import torch
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if self.rank == 0:
            idt = torch.tensor(list(self.o3d.nccl_unique_id()), dtype=torch.uint8, device="cuda")
        self.dist.broadcast(idt, 0)
        return bytes(idt.cpu().tolist())

    def barrier(self, ses=None):
        if ses is not None:
            ses.sync()
        if self.dist is not None:
            import torch
            self.dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(self, v):
        if self.dist is None:
            return v
        import torch
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())


def open_session(env, w, nranks=None, rank=None, nccl_id=None):
    o3d = env.o3d
    nranks = env.world if nranks is None else nranks
    rank = env.rank if rank is None else rank
    nx, ny, nz = w["grid"]
    ph = w["phys"]
    cfg = o3d.make_config(nx, ny, nz, *w["d"], bc=w["bc"], re=ph["re"], cs=ph["cs"], dt=ph["dt"],
                          itscheme=3, iles=ph["iles"], nscr=w["nscr"], sc=1.0, omega=ph["omega"],
                          eps=ph["eps"], kmax=ph["kmax"], idyn=ph["idyn"], multigrid=w["multigrid"],
                          rank=rank, nranks=nranks,
                          nccl_id=(env.nccl_id() if nccl_id is None and nranks > 1 else nccl_id))
    ses = o3d.Session(cfg)
    plane = nx * ny
    if plane * ses.nz_local <= (96 << 20):
        ses.set(**w["init"](ses.z0, ses.nz_local))
    else:
        # large slabs (512^3 and up) are generated and uploaded in z chunks of <= 512 MB, so that
        # neither the host nor the device ever holds a second whole-field copy (1024^3: 8.6 GB)
        step = max(1, (64 << 20) // plane)
        for k0 in range(0, ses.nz_local, step):
            nk = min(step, ses.nz_local - k0)
            for name, arr in w["init"](ses.z0 + k0, nk).items():
                ses.upload_planes(name, arr, k0)
    return ses


def stage_table(w, ses, tm, K, ms_dev, iters, peak):
    """per-stage roofline of this rank's slab from the session's CUDA-event timers"""
    nx, ny, nz = w["grid"]
    ph, bc = w["phys"], w["bc"]
    nloc = float(nx) * ny * ses.nz_local
    b_rhs = B_RHS[3] + (8.0 if ph["iles"] else 0.0)
    stages = {}
    alg = {"rhs": b_rhs, "div": B_DIV, "corr": B_CORR, "transeq": B_TRANSEQ}
    for k, b in alg.items():
        ms, cnt = tm[k]
        if cnt and ms > 0:
            stages[k] = {"launches": int(cnt), "ms_per_launch": ms / cnt, "bytes_per_pt": b,
                         "gbs": b * nloc / (ms / cnt * 1e-3) / 1e9}
    ms_sor, sweeps = tm["sor"]
    seams = any(bc[a] == 0 and (w["grid"][a] % 2) for a in range(3))
    inplace = (seams and os.environ.get("O3D_SOR_SEAM") == "inplace") or \
        os.environ.get("O3D_SOR_FUSED") == "off"
    persistent, peer = ses.sor_path()
    if w["multigrid"]:
        if ms_sor > 0:
            stages["multigrid"] = {"launches": int(max(1, sweeps)), "ms_per_launch": ms_sor / max(1, sweeps),
                                   "note": "V(5,4) cycles; launch-latency-bound coarse levels, no "
                                           "single roofline (DESIGN.md section 6)"}
    elif sweeps:
        nl = 2 * sweeps if inplace else sweeps
        bpp = B_SOR_HALF if inplace else B_SOR_FUSED
        kern = "sor_rb_kernel" if inplace else (
            ("sor_persist_kernel<seam>" if seams else "sor_persist_kernel") if persistent else
            ("sor_tma_kernel<seam> + seam classes" if seams else "sor_tma_kernel"))
        stages["sor"] = {"launches": int(nl), "ms_per_launch": ms_sor / nl, "bytes_per_pt": bpp,
                         "gbs": bpp * nloc / (ms_sor / nl * 1e-3) / 1e9,
                         "iterations_per_step": sweeps / K, "kernel": kern,
                         "launch_unit": "one red+black iteration (a pass of the persistent kernel: "
                                        "span of the whole solve / iterations)" if persistent else
                                        "one launch"}
    for k in stages:
        if "gbs" in stages[k]:
            stages[k]["frac"] = stages[k]["gbs"] / peak
        stages[k]["share_of_step"] = stages[k]["ms_per_launch"] * stages[k]["launches"] / ms_dev
    k_mean = float(np.mean(iters))
    b_step = b_rhs + B_DIV + B_CORR + (B_TRANSEQ if w["nscr"] else 0.0)
    whole = None
    if not w["multigrid"]:
        b_step += (32.0 if inplace else B_SOR_FUSED) * k_mean
        whole = {"bytes_per_pt_step": b_step, "gbs": b_step * nloc / (ms_dev / K * 1e-3) / 1e9}
        whole["frac"] = whole["gbs"] / peak
    return stages, whole, {"persistent": persistent, "peer_memory": peer}


def run_leg(env, w, K, W, sampler=None, keep=False):
    """-> (result dict, session or None).  Timed with CUDA events on the session stream between
    barriers, max over ranks; W >= 3 untimed warm-up steps pass the Euler / AB2 start-up."""
    o3d = env.o3d
    ses = open_session(env, w)
    W = max(3, W)
    for _ in range(W):
        ses.step()
    env.barrier(ses)
    ses.enable_timers(True)
    ses.timers(reset=True)
    if sampler:
        sampler.start()
        time.sleep(0.3)
    launches0 = o3d.kernel_launches()
    env.barrier(ses)
    t_wall0 = time.perf_counter()
    ses.stopwatch_start()
    iters = [ses.step() for _ in range(K)]
    ms_dev = ses.stopwatch_stop()
    env.barrier(ses)
    t_wall = time.perf_counter() - t_wall0
    launches = o3d.kernel_launches() - launches0
    clocks = None
    if sampler:
        time.sleep(0.2)
        clocks = sampler.stop()
    tm = ses.timers(reset=True)
    ses.enable_timers(False)
    ms_max = env.max_over_ranks(ms_dev)
    nx, ny, nz = w["grid"]
    npts = float(nx) * ny * nz
    peak, peak_src = measured_peak()
    stages, whole, path = stage_table(w, ses, tm, K, ms_dev, iters, peak)
    res = {"workload": w["name"], "grid": [nx, ny, nz], "n_gpus": env.world,
           "planes_per_gpu": ses.nz_local, "value": npts / 1e6 / (ms_max / K / 1e3),
           "unit": "Mpts*steps/s", "ms_per_step": ms_max / K, "steps": K, "warmup": W,
           "solver": "multigrid" if w["multigrid"] else "sor",
           "poisson_iterations_per_step": float(np.mean(iters)), "sor_path": path,
           "stages": stages, "whole_step_roofline": whole,
           "wall_ms_per_step": 1e3 * t_wall / K, "gpu_launches": int(launches),
           "settings": {k: w["phys"][k] for k in ("re", "dt", "omega", "eps", "idyn", "iles")},
           "peak": peak, "peak_source": peak_src}
    if clocks is not None:
        res["clocks"] = clocks
    if keep:
        return res, ses
    ses.close()
    return res, None


def digest(a):
    return hashlib.blake2b(np.ascontiguousarray(a).tobytes(), digest_size=16).hexdigest()


def parity_multi(env, w, steps=3):
    """N > 1 `parity`: the N-rank run (default paths: persistent SOR, peer-memory or NCCL halos)
    against the SAME steps on ONE GPU -- rank 0 runs the whole grid in a single-rank session on
    its own device.  Per point the arithmetic is identical and the halo planes carry the
    neighbour's values verbatim, so the fields must agree bit for bit: compared through 128-bit
    digests of every rank's slab of ux, uy, uz, pp, plus the SOR iteration counts."""
    from osinco3d_b200 import slab
    ses = open_session(env, w)
    iters = [ses.step() for _ in range(steps)]
    path = ses.sor_path()
    mine = {k: digest(ses.download(k)) for k in ("ux", "uy", "uz", "pp")}
    mine["iters"] = iters
    ses.close()
    gathered = [None] * env.world
    env.dist.all_gather_object(gathered, mine)
    out = None
    if env.rank == 0:
        one = open_session(env, w, nranks=1, rank=0)
        it1 = [one.step() for _ in range(steps)]
        ok = all(g["iters"] == it1 for g in gathered)
        bad = []
        nz = w["grid"][2]
        for k in ("ux", "uy", "uz", "pp"):
            full = one.download(k)
            for r in range(env.world):
                z0, nzl = slab.slab_range(nz, r, env.world)
                if digest(full[:, :, z0:z0 + nzl]) != gathered[r][k]:
                    ok = False
                    bad.append("%s@rank%d" % (k, r))
            del full
        one.close()
        out = {"workload": w["name"], "grid": list(w["grid"]), "steps": steps,
               "bitwise_equal": bool(ok), "mismatches": bad, "sor_iters": it1,
               "sor_path": {"persistent": path[0], "peer_memory": path[1]}}
    env.barrier()
    return out


# ----------------------------------------------------------------------------------------------
def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    import osinco3d_b200 as o3d     # raises if libo3d_b200.so is not built: no fallback
    if o3d.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: libo3d_b200 has no CPU fallback")
    o3d._lib.check(o3d.lib().o3d_set_device(local))

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    env = Env(o3d, rank, world, local, dist)

    default_headline = (args.n == 256 and args.bc == "freeslip" and not args.les and
                        not args.strong)
    legs = args.legs if args.legs is not None else ("all" if default_headline else "none")
    w = make_workload("tgv", world, args.n, args.bc, args.les, args.strong)
    n, nz = w["n"], w["grid"][2]
    ph = w["phys"]

    K, W = args.steps, max(3, args.warmup)
    sampler = ClockSampler(local) if rank == 0 else None
    head, ses = run_leg(env, w, K, W, sampler=sampler, keep=True)
    stages = head["stages"]
    peak, peak_src = head["peak"], head["peak_source"]
    nloc = float(n) * n * ses.nz_local
    kernel_of = {"rhs": ("march_kernel<3,0,1,RhsEpi<les>>" if ph["iles"] else
                         "march_kernel<0,3,2,RhsEpi<dns>,split ring>"),
                 "div": "march_kernel<0,2,2,DivEpi,split ring>",
                 "sor": stages.get("sor", {}).get("kernel", "sor_persist_kernel"),
                 "corr": "march_kernel<1,0,3,CorrEpi,3 stream fields>"}
    dom = max(stages, key=lambda k: stages[k]["share_of_step"]) if stages else None
    roofline = None
    if dom:
        s = stages[dom]
        roofline = {"bound": "hbm", "kernel": kernel_of[dom], "achieved": s["gbs"], "peak": peak,
                    "unit": "GB/s", "frac": s["frac"],
                    "traffic": ncu_traffic(kernel_of[dom], (n, n, ses.nz_local)),
                    "peak_source": peak_src, "bytes_per_launch": s["bytes_per_pt"] * nloc,
                    "ms_per_launch": s["ms_per_launch"], "share_of_step": s["share_of_step"],
                    "stages": stages}

    e2e_res = None
    if rank == 0 and world == 1 and not args.no_e2e:
        e2e_res = guarded(lambda: run_e2e_resident(o3d, ses, max(3, K // 2)))
    e2e_slabs = None
    if world > 1 and not args.no_e2e:
        sec, nbytes = run_e2e_slabs(o3d, ses, max(3, K // 4), dist)
        npts = float(n) * n * nz
        e2e_slabs = {"value": npts / 1e6 / sec, "unit": "Mpts*steps/s", "ms_per_step": 1e3 * sec,
                     "steps": max(3, K // 4), "h2d_bytes_per_step": nbytes * world,
                     "d2h_bytes_per_step": nbytes * world,
                     "path": "per rank and step: o3d_upload of ux, uy, uz, pp from pinned host slabs "
                             "+ o3d_step + o3d_download of the same four fields (session API; the "
                             "host-pointer module procedures are single-device), wall clock, max "
                             "over ranks; bytes are totals over all ranks"}
    ses.close()

    line = None
    if rank == 0:
        line = {"metric": "Mpts*steps/s (full AB3+SOR time step)", "value": head["value"],
                "unit": "Mpts*steps/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": head["ms_per_step"], "higher_is_better": True,
                "scaling": "strong" if (args.strong and world > 1) else "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": w["name"], "grid": [n, n, nz], "dt": ph["dt"],
                           "re": ph["re"], "omega": ph["omega"], "eps": ph["eps"],
                           "idyn": ph["idyn"], "iles": ph["iles"], "sor_order": "red_black",
                           "sor_iters_per_step": head["poisson_iterations_per_step"],
                           "sor_path": head["sor_path"], "parallelism": "z-slab x%d" % world,
                           "l2": "working set (18 fields x %.0f MB) >> 126 MB L2, no flush needed"
                                 % (nloc * 8 / 1e6)},
                "wall_ms_per_step": head["wall_ms_per_step"],
                "whole_step_roofline": head["whole_step_roofline"],
                "roofline": roofline, "clocks": head.get("clocks"),
                "gpu_launches": head["gpu_launches"]}
    if rank == 0 and world == 1:
        if not args.no_e2e:
            e2e_steps = max(3, K // 4)
            line["e2e"] = guarded(lambda: run_e2e(o3d, w, e2e_steps, 3, args.pipeline))
            line["e2e_resident"] = e2e_res
    elif rank == 0:
        line["e2e"] = e2e_slabs

    # ---- the other configurations (each a full leg of its own) ----
    specs = leg_specs(world)
    if legs == "none":
        specs = []
    elif legs != "all":
        want = set(legs.split(","))
        specs = [s for s in specs if s[0] in want]
    configs = []
    for key, kw in specs:
        wl = make_workload(nranks=world, **kw)
        try:
            res, _ = run_leg(env, wl, max(6, K // 2), 4)
            res["key"] = key
        except Exception as e:      # a leg that fails is reported, it does not sink the headline
            res = {"key": key, "workload": wl["name"], "error": "%s: %s" % (type(e).__name__, e)}
        configs.append(res)
    if rank == 0 and configs:
        line["configs"] = configs

    # ---- parity ----
    if not args.no_parity:
        if world > 1:
            par = []
            wp = make_workload("tgv", world, args.n if args.n <= 256 else 256, args.bc, args.les,
                               args.strong, perturb=True)
            par.append(parity_multi(env, wp))
            if legs != "none":
                par.append(parity_multi(env, make_workload("cojet", world)))
            if rank == 0:
                line["parity"] = {"vs": "the same steps on ONE GPU (rank 0, single-rank session)",
                                  "bitwise_equal": all(p["bitwise_equal"] for p in par),
                                  "legs": par}
        elif rank == 0 and not args.no_cpu:
            line["parity"] = guarded(lambda: oracle_parity(o3d))

    if rank == 0 and world == 1 and not args.no_cpu:
        line["cpu_baseline"] = guarded(lambda: cpu_baseline(args, n))
    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
