"""GPU tests of the multigrid Poisson solver behind solve_poisson_multigrid
(src/poisson_multigrid.f90:10, called from src/integration.f90:244).

The reference routine is undefined behaviour as called, so no reference MG
result exists.  Parity is pinned where SURVEY 8c puts it: against the SOR ORACLE's converged
solution on the same 7-point operator and neighbour rule, modulo the additive constant of the
singular Neumann/periodic operator, and through the residual bound max|rhs - L p|/|A| < tol.
oracle/mg_model.py (NumPy model of the same hierarchy) supplies the expected cycle counts.
"""
import numpy as np
import pytest

from conftest import smooth_field
from test_gpu_poisson import VARIANTS, consistent_problem, laplacian

pytestmark = pytest.mark.gpu

# odd periodic (non-nested tables + seam classes), even (nested), mixed, anisotropic spacing
SHAPES = [(33, 33, 33), (32, 32, 32), (41, 37, 21), (40, 36, 22), (61, 30, 17)]


def bind(M, bc):
    M.schemes(bc[0], bc[0], bc[1], bc[1], bc[2], bc[2])


@pytest.mark.parametrize("variant", list(VARIANTS))
@pytest.mark.parametrize("shape", SHAPES)
def test_multigrid_converges_to_the_sor_oracle_solution(gpu, O, variant, shape):
    from osinco3d_b200 import modules as M
    from oracle import mg_model
    bc = VARIANTS[variant]
    d = (0.11, 0.12, 0.10)
    rhs, _ = consistent_problem(shape, d, bc, 11)
    A = 2.0 * sum(1.0 / x ** 2 for x in d)
    tol = 1e-10 * np.max(np.abs(rhs)) / A
    bind(M, bc)
    pg = np.asfortranarray(np.zeros(shape))
    cyc, dmax = M.solve_poisson_multigrid(pg, rhs, *d, 10000, 5, 4, tol)
    assert dmax < tol and cyc <= 14, (cyc, dmax, tol)
    # the reported measure is the true residual of the returned field
    res = np.max(np.abs(rhs - laplacian(pg, d, bc))) / A
    assert abs(res - dmax) <= 1e-3 * tol + 1e-14 * np.max(np.abs(rhs)) / A, (res, dmax)
    # SOR oracle on the same operator (src/poisson.f90 restatement), converged tightly
    g = O.grid(*shape, *d, bc)
    po = np.asfortranarray(np.zeros(shape))
    it_o, _, dm_o = O.poisson_solver(g, po, rhs, 1.7, tol, 50000, 0)
    assert dm_o < tol
    a, b = pg - pg.mean(), po - po.mean()
    err = np.max(np.abs(a - b)) / np.max(np.abs(b))
    assert err < 2e-6, err
    # same hierarchy and transfer tables as the NumPy model: same number of V-cycles (+-1)
    _, hist, _ = mg_model.solve(np.zeros(shape), np.array(rhs), d, bc, 5, 4, tol, 50)
    assert abs(cyc - (len(hist) - 1)) <= 1, (cyc, len(hist) - 1)
    print("MG cycles gpu=%d model=%d; SOR oracle iterations=%d" % (cyc, len(hist) - 1, it_o))


def test_multigrid_matches_numpy_model_first_cycle(gpu):
    """one V-cycle from a zero guess: the device transfer operators, smoother classes and
    coarse solve reproduce the NumPy model to round-off"""
    from osinco3d_b200 import modules as M
    from oracle import mg_model
    for shape, bc in [((33, 25, 17), (1, 1, 1)), ((21, 26, 15), (0, 1, 0)), ((24, 20, 18), (0, 0, 0))]:
        d = (0.1, 0.1, 0.1)
        rhs, _ = consistent_problem(shape, d, bc, 3)
        A = 2.0 * sum(1.0 / x ** 2 for x in d)
        bind(M, bc)
        # tolerance between the initial and the after-one-cycle residual: exactly one cycle runs
        r0 = np.max(np.abs(rhs)) / A
        pg = np.asfortranarray(np.zeros(shape))
        cyc, dmax = M.solve_poisson_multigrid(pg, rhs, *d, 10000, 5, 4, 0.5 * r0)
        assert cyc == 1
        levels = mg_model.build_hierarchy(shape, d, bc)
        pm = mg_model.vcycle(levels, 0, np.zeros(shape), np.array(rhs), 5, 4)
        scale = np.max(np.abs(pm))
        assert np.max(np.abs(pg - pm)) / scale < 1e-11, np.max(np.abs(pg - pm)) / scale


def test_multigrid_warm_start_and_level_cap(gpu):
    """a converged initial guess returns after 0 cycles (the time loop warm-starts from the
    previous pressure); nlevels caps the depth (the reference passes nlevels = kmax)"""
    from osinco3d_b200 import modules as M
    shape, bc, d = (33, 33, 33), (1, 1, 1), (0.1, 0.1, 0.1)
    rhs, _ = consistent_problem(shape, d, bc, 4)
    A = 2.0 * sum(1.0 / x ** 2 for x in d)
    tol = 1e-9 * np.max(np.abs(rhs)) / A
    bind(M, bc)
    pg = np.asfortranarray(np.zeros(shape))
    c1, _ = M.solve_poisson_multigrid(pg, rhs, *d, 10000, 5, 4, tol)
    c2, dm2 = M.solve_poisson_multigrid(pg, rhs, *d, 10000, 5, 4, tol)
    assert c1 >= 3 and c2 == 0 and dm2 < tol
    p2 = np.asfortranarray(np.zeros(shape))
    c3, dm3 = M.solve_poisson_multigrid(p2, rhs, *d, 2, 5, 4, tol)     # two-grid cap
    assert dm3 < tol and c3 >= c1
    a, b = pg - pg.mean(), p2 - p2.mean()
    assert np.max(np.abs(a - b)) / np.max(np.abs(a)) < 1e-6


@pytest.mark.parametrize("bc", [(1, 1, 1), (0, 1, 0)])
def test_time_steps_with_multigrid_match_sor_oracle(gpu, O, bc):
    """correct_pression with multigrid = 1 (src/integration.f90:244) inside whole time steps:
    with both solvers converged tightly the velocity agrees with the SOR-based oracle."""
    import osinco3d_b200 as o3d
    n = 33
    L = np.pi if bc == (1, 1, 1) else 2 * np.pi
    d = L / (n - 1)
    g = O.grid(n, n, n, d, d, d, bc)
    ux, uy, uz, pp, _ = O.init_tgv(g)
    kw = dict(re=400.0, dt=0.02 * d, omega=1.7, eps=1e-11)
    sim = O.Sim(g, itscheme=3, kmax=50000, idyn=0, **kw)
    sim.set(ux=ux, uy=uy, uz=uz, pp=pp)
    cfg = o3d.make_config(n, n, n, d, d, d, bc=bc, itscheme=3, kmax=10000, idyn=0, multigrid=1,
                          **kw)
    ses = o3d.Session(cfg)
    ses.set(ux=ux, uy=uy, uz=uz, pp=pp)
    cycles = []
    for _ in range(4):
        sim.step()
        cycles.append(ses.step())
    scale = max(np.max(np.abs(sim.field(k))) for k in ("ux", "uy", "uz"))
    for k in ("ux", "uy", "uz"):
        err = np.max(np.abs(ses.download(k) - sim.field(k))) / scale
        assert err < 1e-8, (k, err)      # north star: fields after N steps <= 1e-8 relative
    pg, po = ses.download("pp"), sim.field("pp")
    assert np.max(np.abs((pg - pg.mean()) - (po - po.mean()))) / np.max(np.abs(po - po.mean())) < 1e-5
    print("MG V-cycles per step:", cycles)
    ses.close()


def test_multigrid_full_size_residual_property(gpu):
    """257 x 129 x 65 mixed-BC problem (size-independent property: the residual bound holds)"""
    from osinco3d_b200 import modules as M
    shape, bc, d = (257, 129, 65), (0, 1, 0), (0.05, 0.05, 0.05)
    rhs, _ = consistent_problem(shape, d, bc, 9)
    A = 2.0 * sum(1.0 / x ** 2 for x in d)
    tol = 1e-9 * np.max(np.abs(rhs)) / A
    bind(M, bc)
    pg = np.asfortranarray(np.zeros(shape))
    cyc, dmax = M.solve_poisson_multigrid(pg, rhs, *d, 10000, 5, 4, tol)
    res = np.max(np.abs(rhs - laplacian(pg, d, bc))) / A
    assert dmax < tol and res < 1.01 * tol + 1e-15 and cyc <= 12, (cyc, dmax, res)


@pytest.mark.parametrize("variant,shape", [("111111", (33, 41, 37)), ("0000", (32, 48, 40)),
                                           ("0000", (33, 35, 41)), ("0011", (65, 33, 34))])
def test_fused_level0_smoother_equals_inplace_smoother_bitwise(gpu, variant, shape, monkeypatch):
    """On a single rank the level-0 smoother is the SOR solver's fused TMA pass with omega = 1
    (ping-pong buffers, ghost-cell closures, odd seam classes on odd periodic extents); it must
    give the same V-cycles, bit for bit, as the in-place class sweeps (O3D_MG_SMOOTHER=inplace)
    that z-slab runs and the coarse levels use."""
    from osinco3d_b200 import modules as M
    bc = {"0000": (0, 0, 0), "0011": (0, 1, 0), "111111": (1, 1, 1)}[variant]
    M.schemes(bc[0], bc[0], bc[1], bc[1], bc[2], bc[2])
    d = (0.11, 0.13, 0.17)
    rng = np.random.default_rng(3)
    rhs = np.asfortranarray(rng.standard_normal(shape))
    rhs -= rhs.mean()
    out = {}
    for mode in ("inplace", "fused"):
        # "inplace" also replays the coarse levels launch by launch instead of as a CUDA graph
        if mode == "inplace":
            monkeypatch.setenv("O3D_MG_SMOOTHER", "inplace")
            monkeypatch.setenv("O3D_MG_GRAPH", "0")
        else:
            monkeypatch.delenv("O3D_MG_SMOOTHER", raising=False)
            monkeypatch.delenv("O3D_MG_GRAPH", raising=False)
        for npre, npost, tol in ((5, 4, 1e-9), (2, 1, 1e-6), (1, 0, 1e-3)):
            p = np.asfortranarray(np.zeros(shape))
            out[mode, npre] = (M.solve_poisson_multigrid(p, rhs, *d, 10000, npre, npost, tol), p)
    for npre in (5, 2, 1):
        (ra, pa), (rb, pb) = out["inplace", npre], out["fused", npre]
        assert ra == rb, (npre, ra, rb)
        assert np.array_equal(pa, pb), (npre, np.max(np.abs(pa - pb)))
    M.schemes(1, 1, 1, 1, 1, 1)
