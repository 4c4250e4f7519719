#!/usr/bin/env python
"""Golden vectors for the whole hot path, produced FROM THE REFERENCE SOURCE: every routine of
SURVEY.md section 8(a) -- the 20 `derivation` routines, schemes(), divergence / rotational /
calculate_Q_criterion, calculate_nu_t, predict_velocity, the three poisson_solver variants,
correct_pression, correct_velocity, transeq, function_stats -- is read from /root/reference/src,
translated statement by statement into NumPy by f90np.py (see its header for why this is
operation-for-operation what gfortran computes) and executed on seeded inputs for three boundary
configurations.  The reference cannot be compiled in the build image (no Fortran compiler); this
is the closest thing to running it.

    python tests/golden/make_hotpath_golden.py            # writes tests/golden/hotpath.npz
    python tests/golden/make_hotpath_golden.py --check    # regenerate and compare with the file

Run in the build container (where /root/reference exists).  Tests read only the .npz.
"""
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import f90np  # noqa: E402

REF = "/root/reference/src"
OUT = os.path.join(HERE, "hotpath.npz")
FILES = ["derivation.f90", "differential_operators.f90", "les_turbulence.f90", "poisson.f90",
         "integration.f90", "functions.f90", "initialization.f90", "utils.f90",
         "initial_conditions.f90"]
DER = ["derx_00", "derxp_11", "derxi_11", "dery_00", "deryp_11", "deryi_11", "derz_00", "derzp_11",
       "derzi_11", "derxx_00", "derxxp_11", "derxxi_11", "deryy_00", "deryyp_11", "deryyi_11",
       "derzz_00", "derzzp_11", "derzzi_11", "derz_2dsim", "derzz_2dsim"]
NAMES = DER + ["contains_nan", "function_stats", "schemes", "divergence", "rotational",
               "calculate_q_criterion", "calculate_nu_t", "predict_velocity",
               "poisson_solver_0000", "poisson_solver_0011", "poisson_solver_111111",
               "correct_pression", "correct_velocity", "transeq",
               "average_3d_array", "statistics_calc", "calculate_residuals", "old_values",
               "compute_cfl",
               "compute_velocity_magnitude",
               "initialize_taylor_green_vortex", "initialize_mixing_layer",
               "initialize_coplanar_jet", "dery1d", "calcul_u_base", "normalize1d",
               "add_oscillations_init"]

SHAPE = (8, 7, 9)                    # all extents different, >= 7 (small: the file is committed)
D = (0.0371, 0.0412, 0.0293)
# (nbcx, nbcy, nbcz, sim2d): the three Poisson variants + a 2-D run
# + three mixed ones: schemes() picks the Poisson variant from the x and y flags alone, so "ffp"
# runs poisson_solver_111111 (z mirrored in the solver, periodic in the stencils) and "ppf"
# poisson_solver_0000; "fpf" leaves the poisson_solver pointer unbound (no solve)
CONFIGS = {"ppp": (0, 0, 0, 0), "fff": (1, 1, 1, 0), "pfp": (0, 1, 0, 0), "pfp2d": (0, 1, 0, 1),
           "ffp": (1, 1, 0, 0), "ppf": (0, 0, 1, 0), "fpf": (1, 0, 1, 0)}
HAS_SOLVER = {"ppp", "fff", "pfp", "ffp", "ppf"}
RE, SC, CS, DT = 1600.0, 0.7, 0.17, 1.3e-3


def farr(a):
    return np.asfortranarray(np.array(a, dtype=np.float64))


def smooth(seed, shape=SHAPE):
    """smooth, non-symmetric O(1) field"""
    rng = np.random.default_rng(seed)
    x = np.linspace(0.0, 1.0, shape[0])[:, None, None]
    y = np.linspace(0.0, 1.0, shape[1])[None, :, None]
    z = np.linspace(0.0, 1.0, shape[2])[None, None, :]
    a = rng.uniform(0.5, 3.0, 6)
    return farr(np.sin(a[0] * x + 0.3) * np.cos(a[1] * y - 0.2) * np.sin(a[2] * z + 0.7) +
                0.5 * np.cos(a[3] * x * y) + 0.25 * np.sin(a[4] * y * z + a[5] * x))


def ab_coefficients(dt):
    """adt / bdt / cdt exactly as src/initialization.f90 assigns them (lines matched by pattern,
    right-hand sides evaluated by the translator)"""
    c = {"adt": [None] * 3, "bdt": [None] * 3, "cdt": [None] * 3}
    for st in f90np.logical_lines(open(os.path.join(REF, "initialization.f90")).read()):
        m = re.match(r"^([abc]dt)\((\d)\)\s*=\s*(.+)$", st)
        if m:
            c[m.group(1)][int(m.group(2)) - 1] = float(eval(f90np.expr_py(m.group(3), set()),
                                                            {"__builtins__": {}, "dt": dt}))
    assert all(v is not None for k in c for v in c[k]), c
    return farr(c["adt"]), farr(c["bdt"]), farr(c["cdt"])


def namespace():
    ns = dict(f90np.RUNTIME)
    ns.update(periodic=0, free_slip=1, huge=lambda x: np.finfo(np.float64).max)
    text = open(os.path.join(REF, "initialization.f90")).read()
    # the two named constants, from the source
    for nm in ("PERIODIC", "FREE_SLIP"):
        m = re.search(r"integer,\s*parameter\s*::\s*%s\s*=\s*(\d+)" % nm, text)
        ns[nm.lower()] = int(m.group(1))
    # the three output routines the diagnostics end in: capture what they are handed
    ns["_captured"] = {}
    for nm in ("write_statistics", "print_residuals", "save_residu"):
        ns[nm] = (lambda key: (lambda *a: ns["_captured"].__setitem__(key, [float(v) for v in a])))(nm)
    # x, y, z are module-level arrays of `initialization` for the routines that do not get them
    # as dummy arguments (add_oscillations_init)
    f90np.load([os.path.join(REF, f) for f in FILES], NAMES, ns, module_arrays=("x", "y", "z"))
    return ns


def generate():
    ns = namespace()
    out = {}
    nx, ny, nz = SHAPE
    dx, dy, dz = D
    delta = (dx * dy * dz) ** (1.0 / 3.0)
    adt, bdt, cdt = ab_coefficients(DT)
    out["adt"], out["bdt"], out["cdt"] = adt, bdt, cdt
    inp = {"ux": smooth(1), "uy": smooth(2), "uz": smooth(3), "pp": smooth(4),
           "rhs": smooth(5) - np.mean(smooth(5)),
           "phi": farr(np.clip(0.5 + 0.6 * smooth(6), 0.0, 1.0)), "nu_t_in": farr(1e-3 * np.abs(smooth(7)))}
    rng = np.random.default_rng(77)
    for c in "xyz":
        inp["fu" + c] = farr(0.3 * rng.standard_normal(SHAPE + (3,)))
    inp["fphi"] = farr(0.3 * rng.standard_normal(SHAPE + (3,)))
    for k, v in inp.items():
        out["in_" + k] = v
    new = lambda: np.full(SHAPE, np.nan, order="F")  # noqa: E731

    # ---- the 20 stencil routines (also on a minimum-size grid: overlapping boundary planes) ----
    small = farr(np.random.default_rng(9).standard_normal((7, 7, 7)))
    out["in_small"] = small
    for name in DER:
        d = {"x": dx, "y": dy, "z": dz}[name[3]]
        for tag, f in (("", inp["pp"]), ("_small", small)):
            df = np.full(f.shape, np.nan, order="F")
            ns[name](df, f, d)
            assert not np.isnan(df).any(), name
            out["%s%s" % (name, tag)] = df
    out["function_stats"] = farr(ns["function_stats"](inp["pp"], nx, ny, nz))

    for cfg, (bx, by, bz, sim2d) in CONFIGS.items():
        ns.update(nbcx1=bx, nbcxn=bx, nbcy1=by, nbcyn=by, nbcz1=bz, nbczn=bz, sim2d=sim2d)
        ns["poisson_solver"] = None      # unbound unless schemes() binds it
        ns["schemes"]()
        assert (ns["poisson_solver"] is not None) == (cfg in HAS_SOLVER or cfg == "pfp2d"), cfg
        P = cfg + "_"
        ux, uy, uz = (inp[k].copy(order="F") for k in ("ux", "uy", "uz"))
        for odd in (0, 1):
            o = new()
            ns["divergence"](o, ux, uy, uz, dx, dy, dz, nx, ny, nz, odd)
            out[P + "divergence_odd%d" % odd] = o
        r = [new() for _ in range(3)]
        ns["rotational"](r[0], r[1], r[2], ux, uy, uz, dx, dy, dz, nx, ny, nz)
        for c, a in zip("xyz", r):
            out[P + "rot" + c] = a
        q = new()
        ns["calculate_q_criterion"](q, ux, uy, uz, dx, dy, dz, nx, ny, nz)
        out[P + "q"] = q
        nu = new()
        ns["calculate_nu_t"](nu, ux, uy, uz, dx, dy, dz, CS, delta)
        out[P + "nu_t"] = nu

        # predict_velocity: the Euler / AB2 / AB3 start-up of an AB3 run, DNS and LES
        for iles in (0, 1):
            f = [inp["fu" + c].copy(order="F") for c in "xyz"]
            for itime in (1, 2, 3):
                up = [new() for _ in range(3)]
                nut = inp["nu_t_in"].copy(order="F")
                ns["predict_velocity"](up[0], up[1], up[2], ux, uy, uz, f[0], f[1], f[2], RE, adt,
                                       bdt, cdt, itime, 3, dx, dy, dz, nx, ny, nz, iles, CS,
                                       delta, nut)
                for c, a in zip("xyz", up):
                    out[P + "pred_les%d_it%d_u%s" % (iles, itime, c)] = a
                out[P + "pred_les%d_it%d_nu_t" % (iles, itime)] = nut
            for c, a in zip("xyz", f):
                out[P + "pred_les%d_fu%s" % (iles, c)] = a
        # itscheme = 2 history shift
        f = [inp["fu" + c].copy(order="F") for c in "xyz"]
        up = [new() for _ in range(3)]
        nut = inp["nu_t_in"].copy(order="F")
        ns["predict_velocity"](up[0], up[1], up[2], ux, uy, uz, f[0], f[1], f[2], RE, adt, bdt, cdt,
                               4, 2, dx, dy, dz, nx, ny, nz, 0, CS, delta, nut)
        out[P + "pred_sch2_ux"] = up[0]
        out[P + "pred_sch2_fux"] = f[0]

        if cfg in HAS_SOLVER:
            # the bound poisson_solver: (a) kmax sweeps without convergence (loop runs out: iter =
            # kmax + 1), (b) dynamic omega until one of the two exits fires
            for tag, (omega, eps, kmax, idyn) in (("fixed", (1.6, 1e-30, 12, 0)),
                                                  ("dyn", (1.9, 2e-3, 400, 1))):
                pp = inp["pp"].copy(order="F")
                loc = ns["poisson_solver"](pp, inp["rhs"], dx, dy, dz, nx, ny, nz, omega, eps,
                                           kmax, idyn)
                out[P + "sor_%s_pp" % tag] = pp
                out[P + "sor_%s_scalars" % tag] = farr([loc["iter"], loc["omega"], loc["dmax"]])
            # correct_pression on u* = u (divergence / dt -> rhs -> SOR)
            pp = inp["pp"].copy(order="F")
            loc = ns["correct_pression"](pp, ux, uy, uz, dx, dy, dz, nx, ny, nz, DT, 1.7, 1e-4, 300,
                                         1, 0)
            out[P + "pression_pp"] = pp
            out[P + "pression_omega"] = farr([loc["omega"]])

        u = [new() for _ in range(3)]
        ns["correct_velocity"](u[0], u[1], u[2], ux, uy, uz, inp["pp"], DT, dx, dy, dz, nx, ny, nz)
        for c, a in zip("xyz", u):
            out[P + "corr_u" + c] = a

        # transeq: three steps of an AB3 run with clipping at both ends, DNS and LES diffusivity
        for iles in (0, 1):
            phi = inp["phi"].copy(order="F")
            fphi = inp["fphi"].copy(order="F")
            src = farr(np.zeros(SHAPE))
            big = farr([40.0 * v for v in adt])      # large steps: the clip / redistribution acts
            for itime in (1, 2, 3):
                ns["transeq"](phi, ux, uy, uz, src, fphi, RE, SC, big, bdt, cdt, itime, 3, dx, dy,
                              dz, nx, ny, nz, iles, out[P + "nu_t"])
                out[P + "transeq_les%d_it%d_phi" % (iles, itime)] = phi.copy(order="F")
            out[P + "transeq_les%d_fphi" % iles] = fphi
        # statistics_calc: the 17 stats.dat columns as handed to write_statistics
        ns["statistics_calc"](ux, uy, uz, nx, ny, nz, dx, dy, dz, RE, 0.25)
        out[P + "statistics"] = farr(ns["_captured"]["write_statistics"])

    # ---- whole time steps: the call sequence of src/osinco3d_main.f90:105-115, four steps from
    # rest histories (Euler -> AB2 -> AB3 -> AB3), SOR with the omega carried from step to step
    for cfg, iles, nscr, idyn in (("fff", 1, 1, 1), ("pfp", 0, 1, 0), ("ppp", 1, 0, 1)):
        bx, by, bz, sim2d = CONFIGS[cfg]
        ns.update(nbcx1=bx, nbcxn=bx, nbcy1=by, nbcyn=by, nbcz1=bz, nbczn=bz, sim2d=sim2d)
        ns["schemes"]()
        ux, uy, uz, pp, phi = (inp[k].copy(order="F") for k in ("ux", "uy", "uz", "pp", "phi"))
        ux *= 0.2
        uy *= 0.2
        uz *= 0.2
        fu = [farr(np.zeros(SHAPE + (3,))) for _ in range(3)]
        fphi = farr(np.zeros(SHAPE + (3,)))
        nut = farr(np.zeros(SHAPE))
        src = farr(np.zeros(SHAPE))
        up = [new() for _ in range(3)]
        omega, eps, kmax = 1.8, 1e-4, 300
        log = []
        for itime in (1, 2, 3, 4):
            ns["predict_velocity"](up[0], up[1], up[2], ux, uy, uz, fu[0], fu[1], fu[2], RE, adt,
                                   bdt, cdt, itime, 3, dx, dy, dz, nx, ny, nz, iles, CS, delta, nut)
            loc = ns["correct_pression"](pp, up[0], up[1], up[2], dx, dy, dz, nx, ny, nz, DT, omega,
                                         eps, kmax, idyn, 0)
            omega = loc["omega"]
            ns["correct_velocity"](ux, uy, uz, up[0], up[1], up[2], pp, DT, dx, dy, dz, nx, ny, nz)
            if nscr == 1:
                ns["transeq"](phi, ux, uy, uz, src, fphi, RE, SC, adt, bdt, cdt, itime, 3, dx, dy,
                              dz, nx, ny, nz, iles, nut)
            log.append(omega)
        for k, a in (("ux", ux), ("uy", uy), ("uz", uz), ("pp", pp), ("phi", phi), ("nu_t", nut)):
            out["steps_%s_%s" % (cfg, k)] = a
        out["steps_%s_omega" % cfg] = farr(log)

    # ---- initial conditions (inputs of the benchmarks / examples, SURVEY 8c "inputs"):
    # src/initial_conditions.f90:103-174 (TGV + scalar blob), :329-394 (mixing layer), :244-327
    # (coplanar jet), :554-629 + src/utils.f90:9-45 + src/derivation.f90:950-992 (ici = 1
    # oscillations); coordinates as src/initialization.f90:211-219
    u0, l0 = 1.3, 0.8
    origin = (-0.4, -0.15, 0.2)
    xs = [farr([o + float(i) * d for i in range(n)]) for o, d, n in zip(origin, D, SHAPE)]
    out["init_params"] = farr([u0, l0, origin[0], origin[1], origin[2]])
    ns.update(u0=u0, ici=0, delta=delta, x=xs[0], y=xs[1], z=xs[2],
              xlx=D[0] * (nx - 1))
    for nm, fn, ratio in (("tgv", "initialize_taylor_green_vortex", 1.0),
                          ("mixing", "initialize_mixing_layer", 0.0),
                          ("mixing_r", "initialize_mixing_layer", 0.25),
                          ("jet", "initialize_coplanar_jet", 3.0)):
        f5 = [new() for _ in range(5)]
        ns[fn](f5[0], f5[1], f5[2], f5[3], f5[4], xs[0], xs[1], xs[2], nx, ny, nz, l0, ratio, 1)
        for k, a in zip(("ux", "uy", "uz", "pp", "phi"), f5):
            out["init_%s_%s" % (nm, k)] = a
        if nm in ("mixing", "jet"):
            # typesim: 3 = mixing layer in the reference's numbering of the oscillation branch,
            # 0 = the generic branch (coplanar jet)
            for typesim in (3, 0):
                o = [a.copy(order="F") for a in f5[:3]]
                ns["add_oscillations_init"](o[0], o[1], o[2], nx, ny, nz, dy, u0, 0.05, 0.03, 0.02,
                                            typesim)
                for k, a in zip(("ux", "uy", "uz"), o):
                    out["init_%s_osc%d_%s" % (nm, typesim, k)] = a

    # per-step driver diagnostics: residuals, old_values, CFL
    old = [smooth(11), smooth(12), smooth(13)]
    ns["calculate_residuals"](inp["ux"], inp["uy"], inp["uz"], old[0], old[1], old[2], DT, 3.1, 0.9,
                              nx, ny, nz, 7)
    for c, a in zip("uvw", old):
        out["in_old_" + c] = a
    out["residuals"] = farr(ns["_captured"]["print_residuals"])
    out["residuals_saved"] = farr(ns["_captured"]["save_residu"])
    o3 = [new() for _ in range(3)]
    ns["old_values"](inp["ux"], inp["uy"], inp["uz"], o3[0], o3[1], o3[2], nx, ny, nz)
    assert all(np.array_equal(a, inp[k]) for a, k in zip(o3, ("ux", "uy", "uz")))
    loc = ns["compute_cfl"](0.0, 0.0, 0.0, inp["ux"], inp["uy"], inp["uz"], dx, dy, dz, DT)
    out["cfl"] = farr([loc["cflx"], loc["cfly"], loc["cflz"]])
    # the divergence guard: stop on NaN or > 1000
    bad = inp["ux"].copy(order="F")
    bad[3, 4, 5] = 2000.0
    try:
        ns["correct_velocity"](new(), new(), new(), bad, inp["uy"], inp["uz"], farr(np.zeros(SHAPE)),
                               DT, dx, dy, dz, nx, ny, nz)
        out["corr_guard_stops"] = farr([0.0])
    except f90np.FortranStop:
        out["corr_guard_stops"] = farr([1.0])
    out["params"] = farr([RE, SC, CS, DT, delta, dx, dy, dz])
    return out


def main():
    new = generate()
    if "--check" in sys.argv:
        old = np.load(OUT)
        bad = [k for k in new if k not in old or not np.array_equal(old[k], new[k], equal_nan=True)]
        print("entries: %d, mismatching: %s" % (len(new), bad))
        sys.exit(1 if bad else 0)
    np.savez_compressed(OUT, **new)
    print("wrote %s: %d arrays, %d bytes" % (OUT, len(new), os.path.getsize(OUT)))


if __name__ == "__main__":
    main()
