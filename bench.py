#!/usr/bin/env python
"""bench.py -- osinco3d Chorin-projection time step on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--n 256] [--bc freeslip|periodic]
                    [--les] [--impl reference]

One "step" = predict_velocity + correct_pression (divergence + SOR) + correct_velocity, AB3,
on the synthetic Taylor-Green vortex.  Default workload = BASELINE.json
configs[1]: TGV Re=1600 DNS at 256^3 on one B200, with the shipped example's settings
(free-slip on [0,pi]^3, dt = 0.05 dx, omega = 1.887, eps = 1e-4, idyn = 0).

N > 1 (torchrun): z-slab decomposition, weak scaling -- the box is replicated in z
(nz = N (n-1) + 1 planes on [0, N pi], where the TGV is still an exact free-slip solution), halo
planes and the SOR residual go over NCCL.

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
B_SOR_HALF = 16.0     # in-place colour half-sweep (odd periodic grids): half of 32 B/pt/iteration
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
    return ap.parse_args()


def workload(args, nranks):
    n = args.n
    bc = (1, 1, 1) if args.bc == "freeslip" else (0, 0, 0)
    L = PI if args.bc == "freeslip" else 2 * PI
    d = L / (n - 1)                       # dx = xlx/(nx-1), src/initialization.f90:182-184
    nz = nranks * (n - 1) + 1 if (nranks > 1 and not args.strong) else n
    if args.les:   # examples/tgv_re2500_les, dt scaled with dx from 5e-4 @ 129^3
        phys = dict(re=2500.0, dt=5e-4 * 128.0 / (n - 1), omega=1.999, eps=1e-6, idyn=1, iles=1,
                    cs=0.17)
    else:          # examples/tgv_re1600_dns (dt = cfl*dx/u0)
        phys = dict(re=1600.0, dt=0.05 * d, omega=1.887, eps=1e-4, idyn=0, iles=0, cs=0.0)
    name = "tgv_re%d_%s_%s_%dx%dx%d_ab3_sor" % (int(phys["re"]), "les" if args.les else "dns",
                                                 args.bc, n, n, nz)
    return dict(n=n, nz=nz, d=d, bc=bc, phys=phys, name=name)


def tgv_slab(n, nz_local, z0, d):
    """TGV initial fields (src/initial_conditions.f90:141-153) for planes [z0, z0+nz_local)"""
    x = (d * np.arange(n))[:, None, None]
    y = (d * np.arange(n))[None, :, None]
    z = (d * np.arange(z0, z0 + nz_local))[None, None, :]
    ux = np.asfortranarray(np.sin(x) * np.cos(y) * np.cos(z))
    uy = np.asfortranarray(-np.cos(x) * np.sin(y) * np.cos(z))
    uz = np.asfortranarray(np.zeros((n, n, nz_local)))
    pp = np.asfortranarray(0.0625 * (np.cos(2 * x) + np.cos(2 * y)) * (np.cos(2 * z) + 2.0))
    return ux, uy, uz, pp


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


def ncu_traffic(kernel):
    """per-launch DRAM bytes of a kernel from the committed ncu --set full summary, if any"""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(kernel)
        except Exception:
            return None
    return None


# ----------------------------------------------------------------------------------------------
# CPU baseline: the oracle (C restatement of the reference's serial Fortran), 1 thread
# ----------------------------------------------------------------------------------------------
def cpu_sample(args, steps, n):
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
    return times, iters


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  The Fortran
    source cannot be compiled here (no Fortran compiler / FFTW3 in the image), so this times the
    oracle port (gcc -O2, strict IEEE), single thread: the reference is serial code.

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
    w = workload(args, 1)
    times, iters = cpu_sample(args, W + K, n)
    t = times[W:]
    ms = 1e3 * sum(t) / len(t)
    val = (n ** 3) / 1e6 / (ms / 1e3)
    sample = ("%d timed steps (after %d warm-up) of %s at %d^3%s, SOR iters/step %s"
              % (len(t), W, w["name"], n,
                 "" if n == n_full else " (bounded sample of the %d^3 workload)" % n_full,
                 iters[W:]))
    line = {"impl": "reference", "metric": "Mpts*steps/s (full AB3+SOR time step)",
            "value": val, "unit": "Mpts*steps/s", "n_gpus": args.gpus, "steps": len(t),
            "warmup": W, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": w["name"], "grid": [n, n, n], "host": "cpu"},
            "cpu_baseline": {"value": val, "unit": "Mpts*steps/s", "cores": 1, "kind": "port",
                             "sample": sample,
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
    n, d, bc, ph = w["n"], w["d"], w["bc"], w["phys"]
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

    def step(itime):
        o3d._lib.check(lib.o3d_predict_velocity(
            P(up[0]), P(up[1]), P(up[2]), P(u[0]), P(u[1]), P(u[2]), P(f[0]), P(f[1]), P(f[2]),
            cd(ph["re"]), A, B, Cc, itime, 3, cd(d), cd(d), cd(d), n, n, n, ph["iles"],
            cd(ph["cs"]), cd(delta), P(nu_t)))
        o3d._lib.check(lib.o3d_correct_pression(
            P(pp), P(up[0]), P(up[1]), P(up[2]), cd(d), cd(d), cd(d), n, n, n, cd(ph["dt"]),
            C.byref(omega), cd(ph["eps"]), 10000, ph["idyn"], 0, C.byref(it), C.byref(dmax)))
        o3d._lib.check(lib.o3d_correct_velocity(
            P(u[0]), P(u[1]), P(u[2]), P(up[0]), P(up[1]), P(up[2]), P(pp), cd(ph["dt"]), cd(d),
            cd(d), cd(d), n, n, n))
        return it.value

    itime = 0
    for _ in range(warmup):
        itime += 1
        step(itime)
    t0 = time.perf_counter()
    iters = []
    for _ in range(steps):
        itime += 1
        iters.append(step(itime))
    dt_wall = time.perf_counter() - t0      # every call returns after its D2H copies completed
    pool.close()
    M.set_pipeline(chunks_before)
    # bytes per step, counted from the arrays the three calls copy (see modules.cu)
    h2d = (3 + 6) * N * 8 + 4 * N * 8 + 4 * N * 8
    d2h = (3 + 1 + 9) * N * 8 + 1 * N * 8 + 3 * N * 8
    ms = 1e3 * dt_wall / steps
    return {"value": N / 1e6 / (ms / 1e3), "unit": "Mpts*steps/s", "ms_per_step": ms,
            "steps": steps, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
            "sor_iters_per_step": float(np.mean(iters)), "pipeline_chunks": chunks,
            "path": "o3d_predict_velocity + o3d_correct_pression + o3d_correct_velocity with "
                    "pinned HOST arrays (stateless drop-in procedures), wall clock"
                    + ("; predict / correct_velocity pipelined over %d z chunks (upload, kernel "
                       "and download of successive chunks overlap)" % chunks if chunks else "")}


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
    nccl_id = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt = torch.tensor(list(o3d.nccl_unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(idt, 0)
        nccl_id = bytes(idt.cpu().tolist())

    w = workload(args, world)
    n, nz, d, bc, ph = w["n"], w["nz"], w["d"], w["bc"], w["phys"]
    cfg = o3d.make_config(n, n, nz, d, d, d, bc=bc, re=ph["re"], cs=ph["cs"], dt=ph["dt"],
                          itscheme=3, iles=ph["iles"], nscr=0, omega=ph["omega"], eps=ph["eps"],
                          kmax=10000, idyn=ph["idyn"], rank=rank, nranks=world, nccl_id=nccl_id)
    ses = o3d.Session(cfg)
    ux, uy, uz, pp = tgv_slab(n, ses.nz_local, ses.z0, d)
    ses.set(ux=ux, uy=uy, uz=uz, pp=pp)
    del ux, uy, uz, pp

    def barrier():
        ses.sync()
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    K, W = args.steps, max(3, args.warmup)
    for _ in range(W):
        ses.step()
    barrier()
    ses.enable_timers(True)
    ses.timers(reset=True)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    launches0 = o3d.kernel_launches()
    barrier()
    t_wall0 = time.perf_counter()
    ses.stopwatch_start()
    iters = []
    for _ in range(K):
        iters.append(ses.step())
    ms_dev = ses.stopwatch_stop()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = o3d.kernel_launches() - launches0
    if sampler:
        time.sleep(0.2)
    clocks = sampler.stop() if sampler else None
    tm = ses.timers(reset=True)
    ses.enable_timers(False)

    ms_max = ms_dev
    if dist is not None:
        import torch
        t = torch.tensor([ms_dev], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_max = float(t.item())
    npts = float(n) * n * nz
    ms_per_step = ms_max / K
    value = npts / 1e6 / (ms_per_step / 1e3)

    # per-stage roofline (this rank's slab)
    peak, peak_src = measured_peak()
    nloc = float(n) * n * ses.nz_local
    b_rhs = B_RHS[3] + (8.0 if ph["iles"] else 0.0)
    stages = {}
    alg = {"rhs": b_rhs, "div": B_DIV, "corr": B_CORR}
    for k, b in alg.items():
        ms, cnt = tm[k]
        if cnt:
            stages[k] = {"launches": int(cnt), "ms_per_launch": ms / cnt, "bytes_per_pt": b,
                         "gbs": b * nloc / (ms / cnt * 1e-3) / 1e9}
    ms_sor, sweeps = tm["sor"]
    # fused single-pass red+black kernel; with an odd periodic extent (grid not 2-colourable) the
    # pass is followed by two thin seam-class launches (inside the same span), unless the in-place
    # 4-class sweeps are forced with O3D_SOR_SEAM=inplace
    seams = bc[0] == 0 and (n % 2 or nz % 2)
    fused = not (seams and os.environ.get("O3D_SOR_SEAM") == "inplace")
    if sweeps:
        nl = sweeps if fused else 2 * sweeps
        bpp = B_SOR_FUSED if fused else B_SOR_HALF
        stages["sor"] = {"launches": int(nl), "ms_per_launch": ms_sor / nl, "bytes_per_pt": bpp,
                         "gbs": bpp * nloc / (ms_sor / nl * 1e-3) / 1e9,
                         "iterations_per_step": sweeps / K,
                         "kernel": ("sor_tma_kernel<seam> + 2 x sor_seam_kernel" if seams else
                                    "sor_tma_kernel") if fused else "sor_rb_kernel"}
    for k in stages:
        stages[k]["frac"] = stages[k]["gbs"] / peak
        stages[k]["share_of_step"] = stages[k]["ms_per_launch"] * stages[k]["launches"] / ms_dev
    kernel_of = {"rhs": ("march_kernel<3,0,1,RhsEpi<les>>" if ph["iles"] else
                         "march_kernel<0,3,2,RhsEpi<dns>,split ring>"),
                 "div": "march_kernel<0,2,2,DivEpi,split ring>",
                 "sor": "sor_tma_kernel" if fused else "sor_rb_kernel",
                 "corr": "march_kernel<1,0,3,CorrEpi,3 stream fields>"}
    dom = max(stages, key=lambda k: stages[k]["share_of_step"]) if stages else None
    roofline = None
    if dom:
        s = stages[dom]
        roofline = {"bound": "hbm", "kernel": kernel_of[dom], "achieved": s["gbs"], "peak": peak,
                    "unit": "GB/s", "frac": s["frac"], "traffic": ncu_traffic(kernel_of[dom]),
                    "peak_source": peak_src, "bytes_per_launch": s["bytes_per_pt"] * nloc,
                    "ms_per_launch": s["ms_per_launch"], "share_of_step": s["share_of_step"],
                    "stages": stages}
    k_mean = float(np.mean(iters))
    b_step = b_rhs + B_DIV + B_CORR + (B_SOR_FUSED if fused else 32.0) * k_mean
    whole = {"bytes_per_pt_step": b_step, "gbs": b_step * nloc / (ms_dev / K * 1e-3) / 1e9}
    whole["frac"] = whole["gbs"] / peak

    e2e_res = None
    if rank == 0 and world == 1 and not args.no_e2e:
        e2e_res = run_e2e_resident(o3d, ses, max(3, K // 2))
    e2e_slabs = None
    if world > 1 and not args.no_e2e:
        sec, nbytes = run_e2e_slabs(o3d, ses, max(3, K // 4), dist)
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
        line = {"metric": "Mpts*steps/s (full AB3+SOR time step)", "value": value,
                "unit": "Mpts*steps/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "strong" if (args.strong and world > 1) else "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": w["name"], "grid": [n, n, nz], "dt": ph["dt"],
                           "re": ph["re"], "omega": ph["omega"], "eps": ph["eps"],
                           "idyn": ph["idyn"], "iles": ph["iles"], "sor_order": "red_black",
                           "sor_iters_per_step": k_mean, "parallelism": "z-slab x%d" % world,
                           "l2": "working set (18 fields x %.0f MB) >> 126 MB L2, no flush needed"
                                 % (nloc * 8 / 1e6)},
                "wall_ms_per_step": 1e3 * t_wall / K, "whole_step_roofline": whole,
                "roofline": roofline, "clocks": clocks, "gpu_launches": int(launches)}
    if rank == 0 and world == 1:
        if not args.no_e2e:
            e2e_steps = max(3, K // 4)
            line["e2e"] = run_e2e(o3d, w, e2e_steps, 3, args.pipeline)
            line["e2e_resident"] = e2e_res
        if not args.no_cpu:
            ncpu = args.cpu_n or n
            nsteps = 5 if ncpu >= 200 else 8      # ~10-15 s of CPU work at 256^3
            times, its = cpu_sample(args, nsteps, ncpu)
            tt = times[1:] if len(times) > 1 else times
            ms = 1e3 * sum(tt) / len(tt)
            line["cpu_baseline"] = {
                "value": (ncpu ** 3) / 1e6 / (ms / 1e3), "unit": "Mpts*steps/s", "cores": 1,
                "kind": "port",
                "sample": "%d steps of the same workload at %d^3 (first step dropped), SOR "
                          "iters/step %s; oracle = C restatement of the serial Fortran reference "
                          "(gfortran absent)" % (len(tt), ncpu, its[1:] if len(its) > 1 else its)}
    elif rank == 0:
        line["e2e"] = e2e_slabs
    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
