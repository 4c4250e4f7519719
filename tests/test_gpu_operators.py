"""GPU parity, operator level: every module procedure of the drop-in C ABI (section A of
include/o3d_b200.h, called through osinco3d_b200.modules) against the CPU oracle on the same
seeded inputs.

Bar (BASELINE.json north_star): derivative / divergence / RHS fields <= 1e-12 relative
max-norm.  The kernels evaluate the reference's expressions in the reference's order with FMA
contraction off, so the stencil-only outputs are in fact asserted BIT-EXACT here.
"""
import numpy as np
import pytest

from conftest import rand_field, smooth_field, rel_max

pytestmark = pytest.mark.gpu

# odd, non-tile-multiple extents; one > 32+ in x and > 8 in y so several CTAs + partial tiles
SHAPES = [(37, 29, 23), (70, 19, 9), (7, 7, 7), (33, 8, 40)]
BCS = [(1, 1, 1), (0, 0, 0), (0, 1, 0), (1, 1, 0), (0, 0, 1)]


def bc_flags(bc):
    return (bc[0], bc[0], bc[1], bc[1], bc[2], bc[2])


@pytest.mark.parametrize("shape", SHAPES)
def test_all_18_derivative_routines_bit_exact(gpu, O, shape):
    """src/derivation.f90: der{x,y,z}_00 / p_11 / i_11 and der{xx,yy,zz}_00 / p_11 / i_11"""
    from osinco3d_b200 import modules as M
    f = rand_field(shape, 11)
    d = 0.0371
    for axis, ax in enumerate("xyz"):
        for order in (1, 2):
            for closure, suffix in ((0, "_00"), (1, "p_11"), (2, "i_11")):
                name = "der" + ax * order + suffix
                got = getattr(M, name)(f, d)
                ref = O.der(axis, order, closure, f, d)
                assert np.array_equal(got, ref), (name, shape, rel_max(got, ref))
                got2 = M.der(axis, order, closure, f, d)
                assert np.array_equal(got2, ref)


def test_2dsim_routines_are_zero(gpu):
    from osinco3d_b200 import modules as M
    f = rand_field((12, 9, 8), 5)
    assert not M.derz_2dsim(f, 0.1).any()
    assert not M.derzz_2dsim(f, 0.1).any()


@pytest.mark.parametrize("bc", BCS)
def test_pointer_dispatch_follows_schemes(gpu, O, bc):
    """schemes() binds derxp..derzzi from the BC flags (src/initialization.f90:226-281)"""
    from osinco3d_b200 import modules as M
    M.schemes(*bc_flags(bc))
    shape = (19, 23, 17)
    g = O.grid(*shape, 0.1, 0.2, 0.3, bc)
    f = smooth_field(shape, 3)
    for axis, ax in enumerate("xyz"):
        d = (0.1, 0.2, 0.3)[axis]
        for order in (1, 2):
            for parity, sfx in ((0, "p"), (1, "i")):
                got = getattr(M, "der" + ax * order + sfx)(f, d)
                ref = O.der(axis, order, O.lib().orc_closure(O.C.byref(g), axis, parity), f, d)
                assert np.array_equal(got, ref), (ax, order, sfx, bc)


def test_even_closure_first_derivative_vanishes_on_walls(gpu):
    """discrete identity of der?p_11 (src/derivation.f90:87,105)"""
    from osinco3d_b200 import modules as M
    f = rand_field((21, 18, 15), 8)
    assert not M.derxp_11(f, 0.1)[[0, -1], :, :].any()
    assert not M.deryp_11(f, 0.1)[:, [0, -1], :].any()
    assert not M.derzp_11(f, 0.1)[:, :, [0, -1]].any()


def test_derivative_order_of_accuracy(gpu):
    """6th-order first / 4th-order second derivative on sin(x) with the reference's periodic
    quirk: period = n*dx (SURVEY finding 5)."""
    from osinco3d_b200 import modules as M
    errs1, errs2 = [], []
    for n in (32, 64):
        dx = 2 * np.pi / n
        x = dx * np.arange(n)
        f = np.asfortranarray(np.sin(x)[:, None, None] * np.ones((n, 8, 8)))
        errs1.append(np.max(np.abs(M.derx_00(f, dx)[:, 0, 0] - np.cos(x))))
        errs2.append(np.max(np.abs(M.derxx_00(f, dx)[:, 0, 0] + np.sin(x))))
    assert 5.7 < np.log2(errs1[0] / errs1[1]) < 6.3
    assert 3.7 < np.log2(errs2[0] / errs2[1]) < 4.3


@pytest.mark.parametrize("bc", BCS)
@pytest.mark.parametrize("odd", [0, 1])
def test_divergence_bit_exact(gpu, O, bc, odd):
    """src/differential_operators.f90:7-38"""
    from osinco3d_b200 import modules as M
    M.schemes(*bc_flags(bc))
    shape = (41, 27, 13)
    g = O.grid(*shape, 0.05, 0.07, 0.11, bc)
    fx, fy, fz = (rand_field(shape, s) for s in (1, 2, 3))
    got = M.divergence(fx, fy, fz, g.dx, g.dy, g.dz, odd)
    ref = O.divergence(g, fx, fy, fz, odd)
    assert np.array_equal(got, ref), rel_max(got, ref)


@pytest.mark.parametrize("bc", BCS)
def test_rotational_and_q_criterion_bit_exact(gpu, O, bc):
    """src/differential_operators.f90:40-108"""
    from osinco3d_b200 import modules as M
    M.schemes(*bc_flags(bc))
    shape = (35, 21, 19)
    g = O.grid(*shape, 0.05, 0.07, 0.11, bc)
    u = [rand_field(shape, s) for s in (4, 5, 6)]
    got = M.rotational(*u, g.dx, g.dy, g.dz)
    ref = O.rotational(g, *u)
    for a, b in zip(got, ref):
        assert np.array_equal(a, b), rel_max(a, b)
    q = M.calculate_Q_criterion(*u, g.dx, g.dy, g.dz)
    assert np.array_equal(q, O.q_criterion(g, *u))


@pytest.mark.parametrize("bc", BCS)
def test_nu_t_bit_exact(gpu, O, bc):
    """src/les_turbulence.f90:10-97 (+ function_stats print values, :89-90)"""
    from osinco3d_b200 import modules as M
    M.schemes(*bc_flags(bc))
    shape = (34, 17, 21)
    g = O.grid(*shape, 0.05, 0.07, 0.11, bc)
    u = [smooth_field(shape, s) for s in (7, 8, 9)]
    delta = (g.dx * g.dy * g.dz) ** (1.0 / 3.0)
    got, st = M.calculate_nu_t(*u, g.dx, g.dy, g.dz, 0.17, delta, want_stats=True)
    ref = O.calculate_nu_t(g, *u, 0.17, delta)
    assert np.array_equal(got, ref), rel_max(got, ref)
    rs = O.function_stats(ref)
    assert st[0] == rs[0] and st[1] == rs[1] and st[3:] == rs[3:]
    assert abs(st[2] - rs[2]) <= 1e-14 * abs(rs[2])


@pytest.mark.parametrize("bc", BCS)
@pytest.mark.parametrize("iles", [0, 1])
def test_predict_velocity_bit_exact(gpu, O, bc, iles):
    """src/integration.f90:14-197: Euler (itime 1), AB2 (itime 2), AB3, incl. history shift"""
    from osinco3d_b200 import modules as M
    M.schemes(*bc_flags(bc))
    shape = (37, 21, 15)
    g = O.grid(*shape, 0.05, 0.07, 0.11, bc)
    delta = (g.dx * g.dy * g.dz) ** (1.0 / 3.0)
    dt, re, cs = 1.3e-3, 1600.0, 0.17
    adt, bdt, cdt = M.ab_coefficients(dt)
    assert (adt, bdt, cdt) == tuple(O.ab_coefficients(dt))
    u = [smooth_field(shape, s) for s in (1, 2, 3)]
    for itscheme in (1, 2, 3):
        fo = [np.asfortranarray(rand_field(shape + (3,), 20 + c)) for c in range(3)]
        fg = [f.copy(order="F") for f in fo]
        for itime in (1, 2, 3):
            ref = O.predict_velocity(g, *u, *fo, re, dt, itime, itscheme, iles, cs, delta)
            got = M.predict_velocity(*u, *fg, re, adt, bdt, cdt, itime, itscheme, g.dx, g.dy,
                                     g.dz, iles, cs, delta)
            for a, b, nm in zip(got, ref, ("ux_pred", "uy_pred", "uz_pred", "nu_t")):
                assert np.array_equal(a, b), (nm, itscheme, itime, rel_max(a, b))
            for a, b in zip(fg, fo):
                assert np.array_equal(a, b), ("history", itscheme, itime)


def test_predict_velocity_itscheme_unrecognized(gpu):
    """print + stop at src/integration.f90:99-104 -> O3D_ERR_ITSCHEME"""
    from osinco3d_b200 import modules as M
    M.schemes(1, 1, 1, 1, 1, 1)
    shape = (9, 9, 9)
    u = [rand_field(shape, s) for s in (1, 2, 3)]
    f = [np.asfortranarray(np.zeros(shape + (3,))) for _ in range(3)]
    adt, bdt, cdt = M.ab_coefficients(1e-3)
    with pytest.raises(gpu.O3DError) as e:
        M.predict_velocity(*u, *f, 100.0, adt, bdt, cdt, 3, 4, 0.1, 0.1, 0.1, 0, 0.0, 0.1)
    assert e.value.code == gpu._lib.ERR_ITSCHEME


@pytest.mark.parametrize("bc", BCS)
def test_correct_velocity_bit_exact(gpu, O, bc):
    """src/integration.f90:257-330"""
    from osinco3d_b200 import modules as M
    M.schemes(*bc_flags(bc))
    shape = (33, 25, 11)
    g = O.grid(*shape, 0.05, 0.07, 0.11, bc)
    up = [rand_field(shape, s) for s in (1, 2, 3)]
    pp = smooth_field(shape, 4)
    got = M.correct_velocity(*up, pp, 2e-3, g.dx, g.dy, g.dz)
    ref = O.correct_velocity(g, *up, pp, 2e-3)
    for a, b in zip(got[:3], ref[:3]):
        assert np.array_equal(a, b), rel_max(a, b)
    assert got[3] is False and ref[3] == 0


def test_correct_velocity_divergence_guard(gpu, O):
    """NaN or max(u) > 1000 -> write_velocity_diverged + stop (src/integration.f90:309-325)"""
    from osinco3d_b200 import modules as M
    M.schemes(1, 1, 1, 1, 1, 1)
    shape = (12, 10, 9)
    up = [rand_field(shape, s) for s in (1, 2, 3)]
    pp = rand_field(shape, 4)
    bad = [a.copy(order="F") for a in up]
    bad[1][3, 4, 5] = 2000.0
    assert M.correct_velocity(*bad, np.zeros_like(pp, order="F"), 1e-3, 0.1, 0.1, 0.1)[3] is True
    bad = [a.copy(order="F") for a in up]
    bad[2][0, 0, 0] = np.nan
    assert M.correct_velocity(*bad, pp, 1e-3, 0.1, 0.1, 0.1)[3] is True
    # large NEGATIVE values do not trip maxval (the reference tests maxval, not maxval(abs))
    neg = [a.copy(order="F") for a in up]
    neg[0][1, 1, 1] = -5000.0
    assert M.correct_velocity(*neg, np.zeros_like(pp, order="F"), 1e-3, 0.1, 0.1, 0.1)[3] is False


@pytest.mark.parametrize("bc", [(1, 1, 1), (0, 0, 0), (0, 1, 0)])
@pytest.mark.parametrize("iles", [0, 1])
def test_transeq(gpu, O, bc, iles):
    """src/integration.f90:332-468.  The RHS is stencil-only (bit-exact); phi goes through three
    global sums whose association differs on the GPU (tree vs sequential): <= 1e-13."""
    from osinco3d_b200 import modules as M
    M.schemes(*bc_flags(bc))
    shape = (36, 19, 14)
    g = O.grid(*shape, 0.05, 0.07, 0.11, bc)
    dt, re, sc = 2e-3, 500.0, 0.7
    adt, bdt, cdt = M.ab_coefficients(dt)
    u = [smooth_field(shape, s) for s in (1, 2, 3)]
    nu_t = np.asfortranarray(np.abs(smooth_field(shape, 4)) * 1e-3)
    rng = np.random.default_rng(5)
    phi0 = np.asfortranarray(rng.uniform(-0.05, 1.05, shape))   # some values get clipped
    for itscheme in (2, 3):
        fo = np.asfortranarray(rand_field(shape + (3,), 30) * 0.1)
        fg = fo.copy(order="F")
        po, pg = phi0.copy(order="F"), phi0.copy(order="F")
        for itime in (1, 2, 3):
            O.transeq(g, po, *u, fo, re, sc, dt, itime, itscheme, iles, nu_t)
            M.transeq(pg, *u, fg, re, sc, adt, bdt, cdt, itime, itscheme, g.dx, g.dy, g.dz, iles,
                      nu_t=nu_t)
            assert np.array_equal(fg, fo), (itscheme, itime, rel_max(fg, fo))
            assert rel_max(pg, po) < 1e-13, (itscheme, itime, rel_max(pg, po))
            assert pg.min() >= 0.0 and pg.max() <= 1.0
            pg = po.copy(order="F")   # same inputs for the next step


@pytest.mark.parametrize("bc", BCS)
def test_statistics_calc(gpu, O, bc):
    """src/utils.f90:243-375: 17 stats.dat columns (sums re-associated: <= 1e-12)"""
    from osinco3d_b200 import modules as M
    M.schemes(*bc_flags(bc))
    shape = (31, 26, 22)
    g = O.grid(*shape, 0.05, 0.07, 0.11, bc)
    u = [smooth_field(shape, s) for s in (1, 2, 3)]
    got = M.statistics_calc(*u, g.dx, g.dy, g.dz, 1600.0, 0.25)
    ref = O.statistics_calc(g, *u, 1600.0, 0.25)
    assert got[0] == 0.25
    for c in range(1, 17):
        assert abs(got[c] - ref[c]) <= 1e-12 * abs(ref[c]) + 1e-300, (c, got[c], ref[c])


def test_function_stats(gpu, O):
    """src/functions.f90:27: min / max / mean / first argmax in i-fastest order"""
    from osinco3d_b200 import modules as M
    f = rand_field((23, 17, 31), 9)
    f[5, 6, 7] = 10.0
    f[4, 6, 9] = 10.0   # later in memory order: the first occurrence must win
    got = M.function_stats(f)
    ref = O.function_stats(f)
    assert got[0] == ref[0] and got[1] == ref[1] and got[3:] == ref[3:] == [6.0, 7.0, 8.0]
    assert abs(got[2] - ref[2]) < 1e-15


def test_tgv_golden_row1_through_cuda_statistics(gpu, O):
    """The reference's own golden vector (tgv_stats_re1600_dns.dat row 1, t = 0) reproduced by
    the CUDA statistics kernel at 185^3: all 17 columns to the 13 printed digits."""
    import json
    import os
    from osinco3d_b200 import modules as M
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden",
                                       "reference_stats.json")))
    n = 185
    d = 3.141592653589793 / (n - 1)
    g = O.grid(n, n, n, d, d, d, (1, 1, 1))
    ux, uy, uz, pp, phi = O.init_tgv(g)
    M.schemes(1, 1, 1, 1, 1, 1)
    st = M.statistics_calc(ux, uy, uz, d, d, d, 1600.0, 0.0)
    ref = np.array(gold["tgv_re1600_dns"]["rows"][0])
    # 13 printed digits; eps2 (col 3) is a cancelling sum whose association differs on the GPU
    for c in range(17):
        assert abs(st[c] - ref[c]) <= 1e-12 * max(abs(ref[c]), 1e-30) + 1e-300, (c, st[c], ref[c])


def test_invalid_arguments(gpu):
    from osinco3d_b200 import modules as M
    with pytest.raises(gpu.O3DError) as e:
        M.derx_00(np.asfortranarray(np.zeros((6, 8, 8))), 0.1)   # n >= 7 (stencil radius 3)
    assert e.value.code == gpu._lib.ERR_INVALID
