"""The oracle against the reference SOURCE.  tests/golden/hotpath.npz holds the outputs of every
hot-path routine of SURVEY.md 8(a), obtained by translating the reference's Fortran statements
mechanically into NumPy and executing them (tests/golden/make_hotpath_golden.py + f90np.py; the
reference cannot be compiled in the build image).  The oracle -- the C restatement every GPU
parity test compares against -- must reproduce them BIT FOR BIT: the 20 stencil routines, the
procedure-pointer binding of schemes(), divergence / curl / Q, nu_t, predict_velocity with its
history shifts, the three lexicographic SOR solvers (iterates, exit iteration, dynamic omega),
correct_pression, correct_velocity and its guard, transeq with clipping and redistribution.
"""
import ctypes as C
import os

import numpy as np
import pytest

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "hotpath.npz"))
SHAPE = GOLD["in_ux"].shape
RE, SC, CS, DT, DELTA, DX, DY, DZ = [float(v) for v in GOLD["params"]]
CONFIGS = {"ppp": ((0, 0, 0), 0), "fff": ((1, 1, 1), 0), "pfp": ((0, 1, 0), 0),
           "pfp2d": ((0, 1, 0), 1), "ffp": ((1, 1, 0), 0), "ppf": ((0, 0, 1), 0),
           "fpf": ((1, 0, 1), 0)}
DER = ["derx_00", "derxp_11", "derxi_11", "dery_00", "deryp_11", "deryi_11", "derz_00", "derzp_11",
       "derzi_11", "derxx_00", "derxxp_11", "derxxi_11", "deryy_00", "deryyp_11", "deryyi_11",
       "derzz_00", "derzzp_11", "derzzi_11", "derz_2dsim", "derzz_2dsim"]


def inp(name):
    return np.asfortranarray(GOLD["in_" + name]).copy(order="F")


def same(a, b):
    return np.array_equal(np.asarray(a), np.asarray(b))


def closure_of(name):
    axis = "xyz".index(name[3])
    order = 2 if name[4] == name[3] else 1
    if name.endswith("2dsim"):
        return axis, order, 3
    if name.endswith("_00"):
        return axis, order, 0
    return axis, order, 1 if "p_11" in name else 2


@pytest.mark.parametrize("name", DER)
def test_stencil_routine(O, name):
    """src/derivation.f90: all 18 + 2 routines, explicit boundary planes included"""
    axis, order, closure = closure_of(name)
    d = (DX, DY, DZ)[axis]
    for tag, f in (("", inp("pp")), ("_small", inp("small"))):
        assert same(O.der(axis, order, closure, f, d), GOLD[name + tag]), (name, tag)


def test_ab_coefficients_and_function_stats(O):
    """src/initialization.f90:194-202, src/functions.f90:27-63"""
    a, b, c = O.ab_coefficients(DT)
    assert same(a, GOLD["adt"]) and same(b, GOLD["bdt"]) and same(c, GOLD["cdt"])
    assert same(O.function_stats(inp("pp")), GOLD["function_stats"])


@pytest.mark.parametrize("cfg", list(CONFIGS))
def test_operators(O, cfg):
    """src/differential_operators.f90:7-108, src/les_turbulence.f90:10-97 under the pointers
    schemes() binds (src/initialization.f90:226-281)"""
    bc, sim2d = CONFIGS[cfg]
    g = O.grid(*SHAPE, DX, DY, DZ, bc, sim2d)
    u = [inp(k) for k in ("ux", "uy", "uz")]
    for odd in (0, 1):
        assert same(O.divergence(g, *u, odd), GOLD["%s_divergence_odd%d" % (cfg, odd)]), odd
    for c, a in zip("xyz", O.rotational(g, *u)):
        assert same(a, GOLD["%s_rot%s" % (cfg, c)]), c
    assert same(O.q_criterion(g, *u), GOLD[cfg + "_q"])
    assert same(O.calculate_nu_t(g, *u, CS, DELTA), GOLD[cfg + "_nu_t"])


@pytest.mark.parametrize("cfg", list(CONFIGS))
def test_predict_velocity(O, cfg):
    """src/integration.f90:14-197: Euler -> AB2 -> AB3 start-up, DNS and LES, history shifts of
    itscheme 3 and 2"""
    bc, sim2d = CONFIGS[cfg]
    g = O.grid(*SHAPE, DX, DY, DZ, bc, sim2d)
    u = [inp(k) for k in ("ux", "uy", "uz")]
    for iles in (0, 1):
        f = [inp("fu" + c) for c in "xyz"]
        for itime in (1, 2, 3):
            got = O.predict_velocity(g, *u, *f, RE, DT, itime, 3, iles, CS, DELTA)
            for c, a in zip("xyz", got[:3]):
                assert same(a, GOLD["%s_pred_les%d_it%d_u%s" % (cfg, iles, itime, c)]), (iles, itime, c)
            assert same(got[3], GOLD["%s_pred_les%d_it%d_nu_t" % (cfg, iles, itime)]), (iles, itime)
        for c, a in zip("xyz", f):
            assert same(a, GOLD["%s_pred_les%d_fu%s" % (cfg, iles, c)]), (iles, c)
    f = [inp("fu" + c) for c in "xyz"]
    got = O.predict_velocity(g, *u, *f, RE, DT, 4, 2, 0, CS, DELTA)
    assert same(got[0], GOLD[cfg + "_pred_sch2_ux"]) and same(f[0], GOLD[cfg + "_pred_sch2_fux"])


@pytest.mark.parametrize("cfg", ["ppp", "fff", "pfp", "ffp", "ppf"])
def test_sor_solvers_and_correct_pression(O, cfg):
    """src/poisson.f90:6,132,257 (the variant schemes() binds): iterates, the value of `iter`
    after the loop (kmax + 1 when it runs out), the dynamic relaxation factor;
    src/integration.f90:199-255"""
    bc, _ = CONFIGS[cfg]
    g = O.grid(*SHAPE, DX, DY, DZ, bc)
    for tag, (omega, eps, kmax, idyn) in (("fixed", (1.6, 1e-30, 12, 0)),
                                          ("dyn", (1.9, 2e-3, 400, 1))):
        pp = inp("pp")
        it, om, dmax = O.poisson_solver(g, pp, inp("rhs"), omega, eps, kmax, idyn)
        ref = GOLD["%s_sor_%s_scalars" % (cfg, tag)]
        assert (it, om, dmax) == (int(ref[0]), float(ref[1]), float(ref[2])), (tag, it, om, dmax, ref)
        assert same(pp, GOLD["%s_sor_%s_pp" % (cfg, tag)]), tag
    assert int(GOLD[cfg + "_sor_fixed_scalars"][0]) == 13          # ran out: kmax + 1
    assert 1 < int(GOLD[cfg + "_sor_dyn_scalars"][0]) <= 400       # one of the exits fired
    pp = inp("pp")
    it, om, dmax, rhs = O.correct_pression(g, pp, inp("ux"), inp("uy"), inp("uz"), DT, 1.7, 1e-4,
                                           300, 1)
    assert same(pp, GOLD[cfg + "_pression_pp"])
    assert om == float(GOLD[cfg + "_pression_omega"][0])


@pytest.mark.parametrize("cfg", list(CONFIGS))
def test_correct_velocity_and_transeq(O, cfg):
    """src/integration.f90:257-330 and :332-468"""
    bc, sim2d = CONFIGS[cfg]
    g = O.grid(*SHAPE, DX, DY, DZ, bc, sim2d)
    u = [inp(k) for k in ("ux", "uy", "uz")]
    got = O.correct_velocity(g, *u, inp("pp"), DT)
    for c, a in zip("xyz", got[:3]):
        assert same(a, GOLD["%s_corr_u%s" % (cfg, c)]), c
    assert got[3] == 0
    v3 = lambda v: (C.c_double * 3)(*[float(x) for x in v])  # noqa: E731
    big = [40.0 * v for v in GOLD["adt"]]
    for iles in (0, 1):
        phi, fphi = inp("phi"), inp("fphi")
        nu_t = np.asfortranarray(GOLD[cfg + "_nu_t"]).copy(order="F")
        for itime in (1, 2, 3):
            rc = O.lib().orc_transeq(C.byref(g), O._p(phi), O._p(u[0]), O._p(u[1]), O._p(u[2]), None,
                                     O._p(fphi), C.c_double(RE), C.c_double(SC), v3(big),
                                     v3(GOLD["bdt"]), v3(GOLD["cdt"]), itime, 3, iles, O._p(nu_t))
            assert rc == 0
            assert same(phi, GOLD["%s_transeq_les%d_it%d_phi" % (cfg, iles, itime)]), (iles, itime)
        assert same(fphi, GOLD["%s_transeq_les%d_fphi" % (cfg, iles)]), iles
        # the steps were large enough for the clip and the redistribution to act
        assert phi.min() == 0.0 or phi.max() == 1.0


def test_divergence_guard(O):
    """src/integration.f90:309-325: the reference stops on NaN or a value > 1000"""
    assert GOLD["corr_guard_stops"][0] == 1.0
    g = O.grid(*SHAPE, DX, DY, DZ, (1, 1, 1))
    bad = inp("ux")
    bad[3, 4, 5] = 2000.0
    assert O.correct_velocity(g, bad, inp("uy"), inp("uz"), np.asfortranarray(np.zeros(SHAPE)),
                              DT)[3] != 0


@pytest.mark.parametrize("cfg", list(CONFIGS))
def test_statistics_calc(O, cfg):
    """src/utils.f90:243-375: the 17 stats.dat columns as handed to write_statistics, including
    average_3d_array's i-outer / k-inner summation order (src/functions.f90:123-155)"""
    bc, sim2d = CONFIGS[cfg]
    g = O.grid(*SHAPE, DX, DY, DZ, bc, sim2d)
    got = O.statistics_calc(g, inp("ux"), inp("uy"), inp("uz"), RE, 0.25)
    ref = GOLD[cfg + "_statistics"]
    assert got.shape == ref.shape == (17,)
    assert same(got, ref), np.max(np.abs(got - ref) / np.abs(ref))


def test_residuals_and_cfl(O):
    """src/utils.f90:93-160 (with its single-precision real(nx*ny*nz) and the last-match arg-max
    scan), :165-176, :178-205"""
    got = O.calculate_residuals(inp("ux"), inp("uy"), inp("uz"), inp("old_u"), inp("old_v"),
                                inp("old_w"), DT, 3.1, 0.9)
    r = GOLD["residuals"]     # print_residuals(res_u,res_v,res_w, aa,ia,ja,ka, bb,ib,jb,kb, cc,ic,jc,kc)
    ref = np.array([r[0], r[1], r[2], r[3], r[7], r[11], r[4], r[5], r[6], r[8], r[9], r[10],
                    r[12], r[13], r[14]])
    assert same(got, ref), (got, ref)
    assert same(GOLD["residuals_saved"], [7.0, r[0], r[1], r[2]])
    cfl = [np.max(np.abs(inp(k))) * DT / d for k, d in (("ux", DX), ("uy", DY), ("uz", DZ))]
    assert same(GOLD["cfl"], cfl)


@pytest.mark.parametrize("cfg,iles,nscr,idyn", [("fff", 1, 1, 1), ("pfp", 0, 1, 0), ("ppp", 1, 0, 1)])
def test_whole_time_steps(O, cfg, iles, nscr, idyn):
    """four steps of the call sequence of src/osinco3d_main.f90:105-115 (Euler -> AB2 -> AB3 ->
    AB3 from rest histories, SOR with omega carried from step to step, scalar transport) through
    the oracle's main-loop restatement (orc_sim_step)"""
    bc, _ = CONFIGS[cfg]
    g = O.grid(*SHAPE, DX, DY, DZ, bc)
    sim = O.Sim(g, re=RE, dt=DT, itscheme=3, iles=iles, cs=CS, delta=DELTA, nscr=nscr, sc=SC,
                omega=1.8, eps=1e-4, kmax=300, idyn=idyn)
    sim.set(ux=0.2 * inp("ux"), uy=0.2 * inp("uy"), uz=0.2 * inp("uz"), pp=inp("pp"))
    if nscr:
        sim.set(phi=inp("phi"))
    omegas = []
    for _ in range(4):
        sim.step()
        omegas.append(sim.omega)
    assert same(omegas, GOLD["steps_%s_omega" % cfg]), (omegas, GOLD["steps_%s_omega" % cfg])
    for k in ("ux", "uy", "uz", "pp") + (("phi",) if nscr else ()):
        assert same(sim.field(k), GOLD["steps_%s_%s" % (cfg, k)]), k
    sim.close()


def test_initial_conditions(O):
    """src/initial_conditions.f90:103-174, :244-327, :329-394 with the coordinates of
    src/initialization.f90:211-219 -- the inputs of the TGV / mixing-layer / coplanar-jet
    benchmarks and example tests"""
    u0, l0, x0, y0, z0 = [float(v) for v in GOLD["init_params"]]
    g = O.grid(*SHAPE, DX, DY, DZ, (1, 1, 1))
    cases = {"tgv": O.init_tgv(g, nscr=1, delta=DELTA, u0=u0, l0=l0, ratio=1.0, origin=(x0, y0, z0)),
             "mixing": O.init_mixing_layer(g, nscr=1, u0=u0, l0=l0, ratio=0.0, origin=(x0, y0, z0)),
             "mixing_r": O.init_mixing_layer(g, nscr=1, u0=u0, l0=l0, ratio=0.25,
                                             origin=(x0, y0, z0)),
             "jet": O.init_coplanar_jet(g, nscr=1, u0=u0, l0=l0, ratio=3.0, origin=(x0, y0, z0))}
    for nm, fields in cases.items():
        for k, a in zip(("ux", "uy", "uz", "pp", "phi"), fields):
            assert same(a, GOLD["init_%s_%s" % (nm, k)]), (nm, k, np.max(np.abs(a - GOLD["init_%s_%s" % (nm, k)])))


@pytest.mark.parametrize("nm", ["mixing", "jet"])
@pytest.mark.parametrize("typesim", [3, 0])
def test_deterministic_oscillations(O, nm, typesim):
    """src/initial_conditions.f90:554-629 + src/utils.f90:9-45 + src/derivation.f90:950-992
    (ici = 1).  The oracle's version is vectorised NumPy (np.sin / np.cos instead of libm), so
    this one is compared to 4 ulp of the velocity scale instead of bitwise."""
    u0, l0, x0, y0, z0 = [float(v) for v in GOLD["init_params"]]
    g = O.grid(*SHAPE, DX, DY, DZ, (1, 1, 1))
    base = [np.asfortranarray(GOLD["init_%s_%s" % (nm, k)]).copy(order="F") for k in ("ux", "uy", "uz")]
    got = O.add_oscillations_init(g, *base, u0, (0.05, 0.03, 0.02), typesim, x0,
                                  DX * (SHAPE[0] - 1))
    for k, a in zip(("ux", "uy", "uz"), got):
        ref = GOLD["init_%s_osc%d_%s" % (nm, typesim, k)]
        assert np.max(np.abs(a - ref)) <= 4 * np.finfo(float).eps * u0, (k, np.max(np.abs(a - ref)))
    assert np.max(np.abs(got[0] - base[0])) > 1e-3       # the perturbation is really there
