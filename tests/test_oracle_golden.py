"""Pin the CPU oracle against the reference's own golden statistics histories
(tests/golden/reference_stats.json, extracted verbatim by tests/golden/make_golden.py from
examples/tgv_re1600_dns/tgv_stats_re1600_dns.dat and examples/tgv_re2500_les/...).

The reference has no test-suite; these two files are the only outputs it ships."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "reference_stats.json")))
XLX = 3.141592653589793


def tgv_grid(O, n):
    d = XLX / float(n - 1)            # src/initialization.f90:182-184
    return O.grid(n, n, n, d, d, d, (1, 1, 1)), d


def test_dns_row1_all_columns(O):
    """t = 0: TGV init + all free-slip first/second derivative closures + statistics_calc.
    The file prints 13 significant digits; every one of the 17 columns must agree."""
    g, d = tgv_grid(O, 185)
    ux, uy, uz, pp, phi = O.init_tgv(g, nscr=0)
    st = O.statistics_calc(g, ux, uy, uz, 1600.0, 0.0)
    ref = np.array(GOLD["tgv_re1600_dns"]["rows"][0])
    for c in range(17):
        assert abs(st[c] - ref[c]) <= 6e-13 * max(abs(ref[c]), 1e-30) + 1e-300, (c, st[c], ref[c])


def test_dns_row2_25_steps(O):
    """25 full steps (Euler -> AB2 -> AB3, divergence(odd), poisson_solver_111111 with fixed
    omega = 1.887, eps = 1e-4, correction).  dt = cfl*dmin/u0 = 0.05*pi/184.
    Agreement is at Poisson-tolerance noise: E_k, enstrophy 2e-8; dissipation 2e-7."""
    g, d = tgv_grid(O, 185)
    dt = 0.05 * d / 1.0
    ux, uy, uz, pp, phi = O.init_tgv(g, nscr=0)
    s = O.Sim(g, re=1600.0, dt=dt, itscheme=3, omega=1.887, eps=1e-4, kmax=10000, idyn=0)
    s.set(ux=ux, uy=uy, uz=uz, pp=pp)
    for _ in range(25):
        s.step()
    st = s.stats()
    ref = np.array(GOLD["tgv_re1600_dns"]["rows"][1])
    assert abs(st[0] - ref[0]) < 1e-13                      # time = 25*dt
    assert abs(st[1] - ref[1]) / ref[1] < 2e-8              # E_k
    assert abs(st[4] - ref[4]) / ref[4] < 2e-8              # enstrophy
    assert abs(st[2] - ref[2]) / ref[2] < 2e-7              # eps
    assert abs(st[3] - ref[3]) / ref[3] < 2e-7              # eps2
    s.close()


def test_les_row1_25_steps(O):
    """129^3 Smagorinsky LES, dynamic omega (idyn = 1): the golden file predates the current
    LES source: agreement 4.5e-7, inside the 1e-6 history tolerance."""
    n = 129
    g, d = tgv_grid(O, n)
    ux, uy, uz, pp, phi = O.init_tgv(g, nscr=0)
    s = O.Sim(g, re=2500.0, dt=5e-4, itscheme=3, iles=1, cs=0.17, omega=1.999, eps=1e-6,
              kmax=10000, idyn=1)
    s.set(ux=ux, uy=uy, uz=uz, pp=pp)
    for _ in range(25):
        s.step()
    st = s.stats()
    ref = np.array(GOLD["tgv_re2500_les"]["rows"][0])
    assert abs(st[0] - ref[0]) < 1e-13
    for c in (1, 2, 4):
        assert abs(st[c] - ref[c]) / ref[c] < 1e-6, (c, st[c], ref[c])
    s.close()


def test_calculate_residuals_restatement_against_numpy():
    """orc_calculate_residuals (src/utils.f90:93-160) against an independent numpy evaluation:
    interior points only, L2 with the single-precision real(nx*ny*nz) divisor, Linf, and the
    LAST point (array order) attaining each maximum."""
    import numpy as np
    from oracle import oracle_py as O
    O.build()
    rng = np.random.default_rng(7)
    n = (13, 11, 9)
    new = [np.asfortranarray(rng.standard_normal(n)) for _ in range(3)]
    old = [np.asfortranarray(rng.standard_normal(n)) for _ in range(3)]
    old[1][5, 5, 5], new[1][5, 5, 5] = 40.0, 0.0      # the maximum, twice: the later one is kept
    old[1][7, 6, 5], new[1][7, 6, 5] = 0.0, 40.0
    old[2][0, 3, 3] = 1e3                              # boundary point: not scanned
    dt, t_ref, u_ref = 0.25, 2.0, 4.0
    got = O.calculate_residuals(*new, *old, dt, t_ref, u_ref)
    cnt = float(np.float32(n[0] * n[1] * n[2]))
    for c in range(3):
        a = np.abs(old[c] - new[c])[1:-1, 1:-1, 1:-1] / (2.0 * dt)
        assert abs(got[c] - (t_ref / u_ref) * np.sqrt(np.sum(a * a) / cnt)) < 1e-13 * got[c]
        assert got[3 + c] == (t_ref / u_ref) * a.max()
        where = np.argwhere(a == a.max())
        last = max(where, key=lambda q: (q[2], q[1], q[0]))    # k slowest, i fastest
        assert list(got[6 + 3 * c: 9 + 3 * c]) == [last[0] + 2, last[1] + 2, last[2] + 2]
    assert list(got[9:12]) == [8, 7, 6]
