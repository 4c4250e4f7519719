"""z-chunk copy pipelining of the host-pointer procedures (csrc/pipeline.cu, o3d_set_pipeline):
o3d_predict_velocity / o3d_correct_velocity with the host arrays cut into z chunks whose upload,
kernels and download overlap must return BITWISE the arrays of the unpipelined call -- every
output, for every closure combination, time scheme and chunk count, with pageable and with
pinned (truly asynchronous) host memory.  The unpipelined procedures are pinned against the
oracle in test_gpu_operators.py / test_gpu_step.py.
"""
import ctypes as C

import numpy as np
import pytest

from conftest import smooth_field

pytestmark = pytest.mark.gpu

PI = 3.141592653589793


def _inputs(shape, seed, pool=None):
    mk = (lambda a: pool.array(a)) if pool else (lambda a: a)
    u = [mk(smooth_field(shape, seed + k)) for k in range(3)]
    f = []
    for k in range(3):
        a = np.empty(shape + (3,), dtype=np.float64, order="F")
        for l in range(3):
            a[..., l] = 0.3 * smooth_field(shape, seed + 10 + 3 * k + l)
        f.append(mk(a))
    pp = mk(smooth_field(shape, seed + 30))
    return u, f, pp


def _predict(gpu, u, f, d, itime, itscheme, iles, pool=None):
    """o3d_predict_velocity straight through the C ABI (outputs in pinned memory when a pool is
    given); f is modified in place"""
    L = gpu._lib
    lib = gpu.lib()
    from osinco3d_b200 import modules as M
    shape = u[0].shape
    new = (lambda: pool.empty(shape)) if pool else (lambda: np.empty(shape, order="F"))
    up = [new() for _ in range(3)]
    nu_t = new()
    for a in up + [nu_t]:
        a[...] = np.nan
    adt, bdt, cdt = M.ab_coefficients(2e-3)
    v3 = lambda v: (C.c_double * 3)(*v)  # noqa: E731
    P = lambda a: a.ctypes.data_as(L.dp)  # noqa: E731
    cd = C.c_double
    delta = (d[0] * d[1] * d[2]) ** (1.0 / 3.0)
    L.check(lib.o3d_predict_velocity(
        P(up[0]), P(up[1]), P(up[2]), P(u[0]), P(u[1]), P(u[2]), P(f[0]), P(f[1]), P(f[2]),
        cd(1600.0), v3(adt), v3(bdt), v3(cdt), itime, itscheme, cd(d[0]), cd(d[1]), cd(d[2]),
        shape[0], shape[1], shape[2], iles, cd(0.17), cd(delta), P(nu_t)))
    return up, nu_t


def _correct(gpu, up, pp, d, dt, pool=None):
    L = gpu._lib
    lib = gpu.lib()
    shape = pp.shape
    new = (lambda: pool.empty(shape)) if pool else (lambda: np.empty(shape, order="F"))
    u = [new() for _ in range(3)]
    for a in u:
        a[...] = np.nan
    P = lambda a: a.ctypes.data_as(L.dp)  # noqa: E731
    cd = C.c_double
    rc = L.check(lib.o3d_correct_velocity(
        P(u[0]), P(u[1]), P(u[2]), P(up[0]), P(up[1]), P(up[2]), P(pp), cd(dt), cd(d[0]),
        cd(d[1]), cd(d[2]), shape[0], shape[1], shape[2]), allow=(L.ERR_DIVERGED,))
    return u, rc == L.ERR_DIVERGED


@pytest.fixture(params=[0, 1], ids=["download", "hostshift"])
def pipeline(gpu, request):
    """every test runs twice: history levels 2, 3 (and the DNS nu_t) downloaded, and made in the
    host arrays by the worker threads of csrc/host_copier.h (o3d_set_hostshift)"""
    from osinco3d_b200 import modules as M
    before, hs_before = M.get_pipeline(), M.get_hostshift()
    M.set_hostshift(request.param)
    yield M
    M.set_pipeline(before)
    M.set_hostshift(hs_before)


CASES = [
    # shape,            bc,        sim2d, iles, itscheme, itime, chunks
    ((40, 24, 33), (1, 1, 1), 0, 0, 3, 3, 4),
    ((40, 24, 33), (1, 1, 1), 0, 1, 3, 5, 3),
    ((33, 40, 48), (0, 0, 0), 0, 0, 3, 3, 4),     # periodic z: chunk 0 runs last
    ((33, 40, 48), (0, 0, 0), 0, 1, 3, 4, 6),
    ((37, 21, 32), (0, 1, 0), 0, 1, 3, 3, 2),
    ((24, 37, 41), (1, 0, 1), 0, 0, 3, 3, 5),
    ((40, 24, 33), (1, 1, 1), 0, 0, 3, 1, 4),     # Euler start-up step of an AB3 run
    ((40, 24, 33), (1, 1, 1), 0, 0, 3, 2, 4),     # AB2 start-up step
    ((40, 24, 33), (1, 1, 0), 0, 0, 2, 4, 4),     # itscheme = 2
    ((40, 24, 33), (0, 0, 1), 0, 1, 1, 4, 4),     # itscheme = 1
    ((40, 40, 24), (0, 1, 0), 1, 0, 3, 3, 3),     # sim2d: z derivatives are zero
    ((70, 45, 130), (1, 1, 1), 0, 0, 3, 3, 8),    # several tiles, 8 chunks of 16-17 planes
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "%dx%dx%d-bc%d%d%d-s%d-les%d-sch%d-it%d-c%d"
                         % (c[0] + c[1] + c[2:]))
def test_pipelined_predict_velocity_is_bitwise_the_plain_call(gpu, pipeline, case):
    M = pipeline
    shape, bc, sim2d, iles, itscheme, itime, chunks = case
    d = (PI / (shape[0] - 1), 0.9 * PI / (shape[1] - 1), 1.1 * PI / (shape[2] - 1))
    M.schemes(bc[0], bc[0], bc[1], bc[1], bc[2], bc[2], sim2d)
    u, f0, _ = _inputs(shape, 100)
    M.set_pipeline(0)
    fa = [a.copy(order="F") for a in f0]
    upa, nua = _predict(gpu, u, fa, d, itime, itscheme, iles)
    M.set_pipeline(chunks)
    fb = [a.copy(order="F") for a in f0]
    upb, nub = _predict(gpu, u, fb, d, itime, itscheme, iles)
    for k in range(3):
        assert np.array_equal(upa[k], upb[k]), "u* component %d" % k
        for l in range(3):
            assert np.array_equal(fa[k][..., l], fb[k][..., l]), "f component %d level %d" % (k, l + 1)
    assert np.array_equal(nua, nub)
    assert not np.isnan(upb[0]).any() and not np.isnan(nub).any()
    if iles:
        assert np.max(nub) > 0.0
    else:
        assert not nub.any()          # nu_t = 0.0d0, src/integration.f90:112
    M.schemes(1, 1, 1, 1, 1, 1)


@pytest.mark.parametrize("case", [c for c in CASES if c[4] == 3 and c[5] == 3],
                         ids=lambda c: "%dx%dx%d-bc%d%d%d-s%d-c%d" % (c[0] + c[1] + (c[2], c[6])))
def test_pipelined_correct_velocity_is_bitwise_the_plain_call(gpu, pipeline, case):
    M = pipeline
    shape, bc, sim2d, _, _, _, chunks = case
    d = (PI / (shape[0] - 1), 0.9 * PI / (shape[1] - 1), 1.1 * PI / (shape[2] - 1))
    M.schemes(bc[0], bc[0], bc[1], bc[1], bc[2], bc[2], sim2d)
    up, _, pp = _inputs(shape, 200)
    M.set_pipeline(0)
    ua, da = _correct(gpu, up, pp, d, 2e-3)
    M.set_pipeline(chunks)
    ub, db = _correct(gpu, up, pp, d, 2e-3)
    assert not da and not db
    for k in range(3):
        assert np.array_equal(ua[k], ub[k]), "u component %d" % k
        assert not np.isnan(ub[k]).any()
    M.schemes(1, 1, 1, 1, 1, 1)


def test_pipelined_calls_with_pinned_arrays_and_a_whole_projection_step(gpu, pipeline):
    """pinned host memory makes every copy asynchronous (the overlap the pipeline exists for);
    three consecutive steps predict -> pression -> correct, pipelined vs plain, bit for bit"""
    M = pipeline
    shape = (64, 48, 96)
    d = (PI / (shape[0] - 1),) * 3
    dt = 0.05 * d[0]
    M.schemes(1, 1, 1, 1, 1, 1)
    results = []
    for chunks in (0, 6):
        M.set_pipeline(chunks)
        pool = gpu.PinnedPool()
        u, f, pp = _inputs(shape, 300, pool)
        for a in f:
            a[...] = 0.0
        for itime in (1, 2, 3):
            up, _ = _predict(gpu, u, f, d, itime, 3, 0, pool)
            M.correct_pression(pp, up[0], up[1], up[2], d[0], d[1], d[2], dt, 1.887, 1e-6, 500, 0)
            unew, div = _correct(gpu, up, pp, d, dt, pool)
            assert not div
            for k in range(3):
                u[k][...] = unew[k]
        results.append([a.copy() for a in u + f + [pp]])
        pool.close()
    for a, b in zip(*results):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("itscheme,itime,iles", [(3, 3, 0), (3, 1, 1), (2, 4, 0), (1, 4, 0),
                                                 (5, 1, 0)])
def test_hostshift_leaves_untouched_levels_alone_and_fills_the_rest(gpu, pipeline, itscheme,
                                                                    itime, iles):
    """src/integration.f90:176-188 on pinned arrays pre-filled so that every level is
    distinguishable: itscheme 3 -> (new, new, old 2); 2 -> (new, new, old 3); anything else ->
    (new, old 2, old 3) (`select case` without a default; itscheme = 5 is accepted on the Euler
    start-up step, src/integration.f90:84-105)"""
    M = pipeline
    shape = (40, 24, 64)
    d = (PI / (shape[0] - 1),) * 3
    M.schemes(1, 1, 1, 1, 1, 1)
    M.set_pipeline(5)
    pool = gpu.PinnedPool()
    u, f, _ = _inputs(shape, 600, pool)
    old = [a.copy() for a in f]
    up, nu_t = _predict(gpu, u, f, d, itime, itscheme, iles, pool)
    for k in range(3):
        new = f[k][..., 0]
        assert not np.array_equal(new, old[k][..., 0])
        exp2 = new if itscheme in (2, 3) else old[k][..., 1]
        exp3 = old[k][..., 1] if itscheme == 3 else old[k][..., 2]
        assert np.array_equal(f[k][..., 1], exp2), (k, "level 2")
        assert np.array_equal(f[k][..., 2], exp3), (k, "level 3")
    assert (np.max(nu_t) > 0.0) if iles else (not nu_t.any())
    pool.close()


def test_pipelined_correct_velocity_reports_divergence(gpu, pipeline):
    """the NaN / > 1000 guard (src/integration.f90:309-325) fires in whichever chunk holds the
    offending point; the outputs are complete all the same"""
    M = pipeline
    shape = (32, 32, 40)
    d = (0.1, 0.1, 0.1)
    M.schemes(1, 1, 1, 1, 1, 1)
    up, _, pp = _inputs(shape, 400)
    up[1][5, 7, 33] = 5000.0
    M.set_pipeline(0)
    ua, da = _correct(gpu, up, pp, d, 1e-3)
    M.set_pipeline(4)
    ub, db = _correct(gpu, up, pp, d, 1e-3)
    assert da and db
    for k in range(3):
        assert np.array_equal(ua[k], ub[k])


def test_thin_grids_fall_back_to_the_plain_path(gpu, pipeline):
    """fewer than 8 planes per chunk: fewer chunks, or no pipelining at all"""
    M = pipeline
    shape = (20, 20, 12)
    d = (0.1, 0.1, 0.1)
    M.schemes(1, 1, 1, 1, 1, 1)
    u, f0, _ = _inputs(shape, 500)
    M.set_pipeline(0)
    fa = [a.copy(order="F") for a in f0]
    upa, _ = _predict(gpu, u, fa, d, 3, 3, 0)
    M.set_pipeline(8)
    fb = [a.copy(order="F") for a in f0]
    upb, _ = _predict(gpu, u, fb, d, 3, 3, 0)
    for k in range(3):
        assert np.array_equal(upa[k], upb[k])
        assert np.array_equal(fa[k], fb[k])


@pytest.mark.parametrize("bc", [(1, 1, 1), (0, 0, 0), (0, 1, 0), (1, 1, 0)])
def test_pipelined_procedures_bit_exact_against_the_oracle(gpu, O, pipeline, bc):
    """the pipelined calls directly against the CPU restatement of src/integration.f90:14-197 and
    :257-330 (three time levels of an AB3 start-up, LES on, history shift included)"""
    from conftest import rand_field
    M = pipeline
    M.schemes(bc[0], bc[0], bc[1], bc[1], bc[2], bc[2])
    M.set_pipeline(4)
    shape = (37, 21, 35)
    g = O.grid(*shape, 0.05, 0.07, 0.11, bc)
    delta = (g.dx * g.dy * g.dz) ** (1.0 / 3.0)
    dt, re, cs = 1.3e-3, 1600.0, 0.17
    adt, bdt, cdt = M.ab_coefficients(dt)
    u = [smooth_field(shape, s) for s in (1, 2, 3)]
    fo = [np.asfortranarray(rand_field(shape + (3,), 20 + c)) for c in range(3)]
    fg = [f.copy(order="F") for f in fo]
    for itime in (1, 2, 3):
        ref = O.predict_velocity(g, *u, *fo, re, dt, itime, 3, 1, cs, delta)
        got = M.predict_velocity(*u, *fg, re, adt, bdt, cdt, itime, 3, g.dx, g.dy, g.dz, 1, cs,
                                 delta)
        for a, b, nm in zip(got, ref, ("ux_pred", "uy_pred", "uz_pred", "nu_t")):
            assert np.array_equal(a, b), (nm, itime)
        for a, b in zip(fg, fo):
            assert np.array_equal(a, b), ("history", itime)
    up = [rand_field(shape, s) for s in (4, 5, 6)]
    pp = smooth_field(shape, 7)
    got = M.correct_velocity(*up, pp, 2e-3, g.dx, g.dy, g.dz)
    ref = O.correct_velocity(g, *up, pp, 2e-3)
    for a, b in zip(got[:3], ref[:3]):
        assert np.array_equal(a, b)
    assert got[3] is False and ref[3] == 0
    M.schemes(1, 1, 1, 1, 1, 1)
