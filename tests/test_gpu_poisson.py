"""GPU parity of the pressure Poisson solvers (src/poisson.f90) through the C ABI.

The reference sweeps lexicographically (loop-carried dependence).  Two GPU orderings:
  * LEXI_WAVEFRONT reproduces the reference's iterates bit for bit -> asserted BIT-EXACT,
    identical iteration counts, identical dynamic-omega trajectory;
  * RED_BLACK (fast path) is a different sweep ordering, stated explicitly: it converges to the
    same solution only within eps and up to an additive constant (singular Neumann/periodic
    operator) -> compared after mean removal against the residual bound, iteration counts
    reported side by side.
"""
import numpy as np
import pytest

from conftest import rand_field, smooth_field, rel_max

pytestmark = pytest.mark.gpu

VARIANTS = {"0000": (0, 0, 0), "0011": (0, 1, 0), "111111": (1, 1, 1)}


def nbr(n, mirror):
    idx = np.arange(n)
    m1, p1 = idx - 1, idx + 1
    if mirror:
        m1[0], p1[-1] = 1, n - 2
    else:
        m1[0], p1[-1] = n - 1, 0
    return m1, p1


def laplacian(p, d, mirror):
    """the 7-point operator of src/poisson.f90:95-98 (numpy, test helper)"""
    out = np.zeros_like(p)
    for ax in range(3):
        m1, p1 = nbr(p.shape[ax], mirror[ax])
        out += (np.take(p, m1, axis=ax) + np.take(p, p1, axis=ax) - 2.0 * p) / d[ax] ** 2
    return np.asfortranarray(out)


def consistent_problem(shape, d, mirror, seed):
    p_true = smooth_field(shape, seed)
    return np.asfortranarray(laplacian(p_true, d, mirror)), p_true


@pytest.mark.parametrize("variant", list(VARIANTS))
@pytest.mark.parametrize("idyn", [0, 1])
def test_wavefront_order_reproduces_reference_sweep_bitwise(gpu, O, variant, idyn):
    from osinco3d_b200 import modules as M
    bc = VARIANTS[variant]
    shape = (19, 15, 13)
    d = (0.11, 0.13, 0.17)
    g = O.grid(*shape, *d, bc)
    rhs, _ = consistent_problem(shape, d, bc, 3)
    p0 = rand_field(shape, 4, 0.1)
    po, pg = p0.copy(order="F"), p0.copy(order="F")
    it_o, om_o, dm_o = O.poisson_solver(g, po, rhs, 1.7, 1e-7, 60, idyn)
    M.set_sor_order(gpu.SOR_LEXI_WAVEFRONT)
    try:
        fn = getattr(M, "poisson_solver_" + variant)
        it_g, om_g, dm_g = fn(pg, rhs, *d, 1.7, 1e-7, 60, idyn)
    finally:
        M.set_sor_order(gpu.SOR_RED_BLACK)
    assert it_g == it_o, (it_g, it_o)
    assert om_g == om_o and dm_g == dm_o
    assert np.array_equal(pg, po), rel_max(pg, po)


@pytest.mark.parametrize("variant", list(VARIANTS))
def test_wavefront_runs_to_convergence_like_reference(gpu, O, variant):
    from osinco3d_b200 import modules as M
    bc = VARIANTS[variant]
    shape = (17, 14, 12)
    d = (0.1, 0.1, 0.1)
    g = O.grid(*shape, *d, bc)
    rhs, _ = consistent_problem(shape, d, bc, 6)
    po = np.asfortranarray(np.zeros(shape))
    pg = po.copy(order="F")
    it_o, om_o, dm_o = O.poisson_solver(g, po, rhs, 1.6, 1e-8, 5000, 0)
    M.set_sor_order(gpu.SOR_LEXI_WAVEFRONT)
    try:
        M.schemes(bc[0], bc[0], bc[1], bc[1], bc[2], bc[2])
        it_g, om_g, dm_g = M.poisson_solver(pg, rhs, *d, 1.6, 1e-8, 5000, 0)
    finally:
        M.set_sor_order(gpu.SOR_RED_BLACK)
    assert it_o < 5000 and it_g == it_o and dm_g == dm_o
    assert np.array_equal(pg, po)


@pytest.mark.parametrize("variant", list(VARIANTS))
@pytest.mark.parametrize("shape", [(21, 17, 15), (20, 16, 14), (33, 18, 9)])
def test_red_black_converges_to_the_reference_solution(gpu, O, variant, shape):
    """odd extents exercise the periodic seam classes, even ones the plain two-colour sweep"""
    from osinco3d_b200 import modules as M
    bc = VARIANTS[variant]
    d = (0.11, 0.13, 0.17)
    g = O.grid(*shape, *d, bc)
    rhs, _ = consistent_problem(shape, d, bc, 5)
    eps = 1e-11
    po = np.asfortranarray(np.zeros(shape))
    pg = po.copy(order="F")
    it_o, _, dm_o = O.poisson_solver(g, po, rhs, 1.7, eps, 20000, 0)
    fn = getattr(M, "poisson_solver_" + variant)
    it_g, om_g, dm_g = fn(pg, rhs, *d, 1.7, eps, 20000, 0)
    assert it_o <= 20000 and it_g <= 20000, (it_o, it_g)
    assert dm_g < eps and om_g == 1.7
    # residual bound: dmax = max|rhs - L p| / |A|  
    A = 2.0 * sum(1.0 / x ** 2 for x in d)
    res = np.max(np.abs(rhs - laplacian(pg, d, bc))) / A
    assert res < 20 * eps, res
    a = pg - pg.mean()
    b = po - po.mean()
    scale = np.max(np.abs(b))
    assert np.max(np.abs(a - b)) / scale < 1e-7, (np.max(np.abs(a - b)) / scale, it_o, it_g)
    print("SOR iterations lexicographic(oracle)=%d red-black(gpu)=%d" % (it_o, it_g))


def test_red_black_kmax_and_exit_codes(gpu, O):
    """loop exhausted -> Fortran iter == kmax + 1 (src/poisson.f90:53)"""
    from osinco3d_b200 import modules as M
    shape = (16, 16, 16)
    d = (0.1, 0.1, 0.1)
    rhs, _ = consistent_problem(shape, d, (1, 1, 1), 2)
    pg = np.asfortranarray(np.zeros(shape))
    it, om, dm = M.poisson_solver_111111(pg, rhs, *d, 1.5, 1e-14, 7, 0)
    assert it == 8 and dm > 1e-14
    g = O.grid(*shape, *d, (1, 1, 1))
    po = np.asfortranarray(np.zeros(shape))
    it_o, _, _ = O.poisson_solver(g, po, rhs, 1.5, 1e-14, 7, 0)
    assert it_o == 8


def test_null_poisson_pointer(gpu):
    """(free-slip x, periodic y) leaves poisson_solver null (src/initialization.f90:283-301)"""
    from osinco3d_b200 import modules as M
    M.schemes(1, 1, 0, 0, 0, 0)
    p = np.asfortranarray(np.zeros((8, 8, 8)))
    with pytest.raises(gpu.O3DError) as e:
        M.poisson_solver(p, p.copy(order="F"), 0.1, 0.1, 0.1, 1.5, 1e-6, 10, 0)
    assert e.value.code == gpu._lib.ERR_BC
    M.schemes(1, 1, 1, 1, 1, 1)


@pytest.mark.parametrize("bc", [(1, 1, 1), (0, 0, 0), (0, 1, 0)])
def test_correct_pression_rhs_and_solve(gpu, O, bc):
    """src/integration.f90:199-255: rhs = div(u*)/dt bit-exact (wavefront -> pp bit-exact)"""
    from osinco3d_b200 import modules as M
    M.schemes(bc[0], bc[0], bc[1], bc[1], bc[2], bc[2])
    shape = (22, 17, 13)
    d = (0.1, 0.12, 0.14)
    g = O.grid(*shape, *d, bc)
    up = [smooth_field(shape, s) for s in (1, 2, 3)]
    p0 = rand_field(shape, 7, 0.01)
    po, pg = p0.copy(order="F"), p0.copy(order="F")
    it_o, om_o, dm_o, rhs_o = O.correct_pression(g, po, *up, 1e-2, 1.6, 1e-6, 40, 1)
    M.set_sor_order(gpu.SOR_LEXI_WAVEFRONT)
    try:
        it_g, om_g, dm_g = M.correct_pression(pg, *up, *d, 1e-2, 1.6, 1e-6, 40, 1)
    finally:
        M.set_sor_order(gpu.SOR_RED_BLACK)
    assert (it_g, om_g, dm_g) == (it_o, om_o, dm_o)
    assert np.array_equal(pg, po)


@pytest.mark.parametrize("variant,shape", [
    ("0000", (21, 17, 15)), ("0000", (33, 18, 9)), ("0000", (20, 17, 14)), ("0000", (65, 49, 37)),
    ("0000", (71, 34, 41)), ("0011", (21, 17, 15)), ("0011", (67, 40, 35)), ("0011", (32, 21, 33)),
])
@pytest.mark.parametrize("idyn", [0, 1])
def test_fused_seam_pass_equals_inplace_class_sweeps_bitwise(gpu, variant, shape, idyn, monkeypatch):
    """Odd periodic extents (not 2-colourable): the fused TMA pass (even seam classes) followed by
    sor_seam_kernel (odd classes) is the same class order as the four in-place half-sweeps
    (O3D_SOR_SEAM=inplace) -> identical iterates, iteration counts, dmax and omega history.
    Shapes cover several 32 x 16 tiles, partial tiles and seams on one, two or three axes."""
    from osinco3d_b200 import modules as M
    bc = VARIANTS[variant]
    d = (0.11, 0.13, 0.17)
    rhs, _ = consistent_problem(shape, d, bc, 11)
    p0 = rand_field(shape, 12, 0.1)
    fn = getattr(M, "poisson_solver_" + variant)
    out = {}
    for mode in ("inplace", "split", "fused"):
        # inplace: four in-place half-sweeps; split: fused pass + the two odd classes as separate
        # launches + control kernel; fused (default): pass + ONE cooperative launch that sweeps
        # both odd classes behind a grid barrier and evaluates the exit tests / dynamic omega
        if mode == "fused":
            monkeypatch.delenv("O3D_SOR_SEAM", raising=False)
        else:
            monkeypatch.setenv("O3D_SOR_SEAM", mode)
        for kmax in (1, 2, 7, 400):
            p = p0.copy(order="F")
            out[mode, kmax] = (fn(p, rhs, *d, 1.8, 1e-9, kmax, idyn), p)
    for kmax in (1, 2, 7, 400):
        for mode in ("split", "fused"):
            (ra, pa), (rb, pb) = out["inplace", kmax], out[mode, kmax]
            assert ra == rb, (mode, kmax, ra, rb)
            assert np.array_equal(pa, pb), (mode, kmax, rel_max(pa, pb))
    assert out["fused", 400][0][2] < out["fused", 7][0][2]   # dmax keeps falling


@pytest.mark.parametrize("variant,shape", [
    ("111111", (40, 33, 29)), ("111111", (64, 48, 40)), ("111111", (21, 17, 15)),
    ("0000", (32, 48, 20)), ("0000", (66, 34, 38)), ("0011", (64, 37, 24)), ("0011", (34, 20, 36)),
])
@pytest.mark.parametrize("idyn", [0, 1])
def test_fused_pass_equals_inplace_half_sweeps_bitwise(gpu, variant, shape, idyn, monkeypatch):
    """2-colourable grids: one fused TMA pass (red update of plane k+2, black update of plane k,
    ping-pong buffers, ghost-cell closures) == a red in-place half-sweep followed by a black one
    (sor_rb_kernel with the reference's neighbour-index rule, O3D_SOR_FUSED=off): identical
    iterates, iteration counts, dmax and omega history, for mirrored, periodic and mixed axes,
    full and partial tiles, several z chunks."""
    from osinco3d_b200 import modules as M
    bc = VARIANTS[variant]
    d = (0.11, 0.13, 0.17)
    rhs, _ = consistent_problem(shape, d, bc, 21)
    p0 = rand_field(shape, 22, 0.1)
    fn = getattr(M, "poisson_solver_" + variant)
    out = {}
    for mode in ("off", "fused"):
        if mode == "fused":
            monkeypatch.delenv("O3D_SOR_FUSED", raising=False)
        else:
            monkeypatch.setenv("O3D_SOR_FUSED", "off")
        for kmax in (1, 2, 5, 300):
            p = p0.copy(order="F")
            out[mode, kmax] = (fn(p, rhs, *d, 1.8, 1e-9, kmax, idyn), p)
    for kmax in (1, 2, 5, 300):
        (ra, pa), (rb, pb) = out["off", kmax], out["fused", kmax]
        assert ra == rb, (kmax, ra, rb)
        assert np.array_equal(pa, pb), (kmax, rel_max(pa, pb))


@pytest.mark.parametrize("variant,shape", [
    ("111111", (96, 80, 70)), ("111111", (40, 33, 29)), ("111111", (130, 20, 9)),
    ("0000", (64, 48, 40)), ("0000", (65, 49, 37)), ("0000", (71, 34, 41)),
    ("0011", (64, 37, 24)), ("0011", (67, 40, 35)), ("0011", (33, 16, 8)),
])
@pytest.mark.parametrize("idyn", [0, 1])
def test_persistent_solve_equals_launch_per_pass_bitwise(gpu, variant, shape, idyn, monkeypatch):
    """The persistent kernel (all iterations of a solve in one cooperative launch: grid barrier
    between passes, exit tests and dynamic omega of src/poisson.f90:110-122 evaluated by the last
    CTA to arrive, odd seam classes folded in) against the launch-per-pass path (O3D_SOR_PERSIST=0:
    one sor_tma_kernel launch + control / seam kernel per iteration, host polls): identical
    iterates, Fortran `iter`, dmax and omega, for every exit (kmax 1 / 2 / odd / converged),
    2-colourable and odd-periodic grids, one and several items per CTA."""
    from osinco3d_b200 import modules as M
    bc = VARIANTS[variant]
    d = (0.11, 0.13, 0.17)
    rhs, _ = consistent_problem(shape, d, bc, 31)
    p0 = rand_field(shape, 32, 0.1)
    fn = getattr(M, "poisson_solver_" + variant)
    out = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("O3D_SOR_PERSIST", mode)
        for kmax in (1, 2, 7, 500):
            p = p0.copy(order="F")
            out[mode, kmax] = (fn(p, rhs, *d, 1.8, 1e-9, kmax, idyn), p)
    for kmax in (1, 2, 7, 500):
        (ra, pa), (rb, pb) = out["0", kmax], out["1", kmax]
        assert ra == rb, (kmax, ra, rb)
        assert np.array_equal(pa, pb), (kmax, rel_max(pa, pb))
    assert out["1", 500][0][2] < out["1", 7][0][2]


@pytest.mark.parametrize("bc,n", [((1, 1, 1), 64), ((0, 0, 0), 49), ((0, 1, 0), 48)])
def test_persistent_steps_equal_launch_per_pass_steps_bitwise(gpu, O, bc, n, monkeypatch):
    """whole time steps (o3d_step: the correction queued behind the persistent kernel, gated on its
    outcome, ping-pong parity resolved on the device) == the same steps with one launch per pass
    and host polls; the session reports which path the last solve took"""
    PI = 3.141592653589793
    d = tuple((PI if b else 2 * PI) / (n - 1) for b in bc)
    g = O.grid(n, n, n, *d, bc)
    ux, uy, uz, pp, phi = O.init_tgv(g, nscr=1)
    res = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("O3D_SOR_PERSIST", mode)
        cfg = gpu.make_config(n, n, n, *d, bc=bc, re=400.0, dt=0.02 * d[0], itscheme=3, iles=1,
                              cs=0.17, nscr=1, omega=1.7, eps=1e-6, kmax=500, idyn=1)
        ses = gpu.Session(cfg)
        ses.set(ux=ux, uy=uy, uz=uz, pp=pp, phi=phi)
        iters = [ses.step() for _ in range(6)]
        assert ses.sor_path() == (mode == "1", False)
        res[mode] = (iters, ses.omega, ses.last_dmax,
                     {k: ses.download(k) for k in ("ux", "uy", "uz", "pp", "phi")})
        ses.close()
    assert res["0"][:3] == res["1"][:3], (res["0"][:3], res["1"][:3])
    assert max(res["1"][0]) > 1
    for k, a in res["0"][3].items():
        assert np.array_equal(a, res["1"][3][k]), (k, rel_max(a, res["1"][3][k]))
