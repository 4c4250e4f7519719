"""old_values / calculate_residuals on the device (src/utils.f90:165-176, :93-160)
against the oracle's literal restatement.  Sums are reduced in a different association than
the reference's triple loop -> res_? within round-off (1e-13 relative, stated here); the Linf
values and the (last-occurrence) indices of the maxima are exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
PI = 3.141592653589793


@pytest.mark.parametrize("bc,n", [((1, 1, 1), (40, 33, 29)), ((0, 0, 0), (33, 32, 36))])
def test_residuals_match_oracle(gpu, O, bc, n):
    d = tuple((PI if b else 2 * PI) / (m - 1) for b, m in zip(bc, n))
    g = O.grid(*n, *d, bc)
    ux, uy, uz, pp, _ = O.init_tgv(g)
    dt = 0.02 * d[0]
    cfg = gpu.make_config(*n, *d, bc=bc, re=400.0, dt=dt, itscheme=3, omega=1.7, eps=1e-7,
                          kmax=2000)
    ses = gpu.Session(cfg)
    ses.set(ux=ux, uy=uy, uz=uz, pp=pp)
    for _ in range(2):
        ses.step()
    ses.old_values()                       # src/osinco3d_main.f90:104
    old = [ses.download(k) for k in ("ux", "uy", "uz")]
    for k, a in zip(("old_ux", "old_uy", "old_uz"), old):
        assert np.array_equal(ses.download(k), a)
    ses.step()
    new = [ses.download(k) for k in ("ux", "uy", "uz")]
    got = ses.calculate_residuals(dt, t_ref=2.5, u_ref=1.5)
    exp = O.calculate_residuals(*new, *old, dt, 2.5, 1.5)
    assert np.all(exp[:3] > 0)
    assert np.allclose(got[:3], exp[:3], rtol=1e-13, atol=0), (got[:3], exp[:3])
    assert np.array_equal(got[3:], exp[3:]), (got[3:], exp[3:])
    ses.close()


def test_residuals_tie_keeps_last_point_and_excludes_boundary(gpu, O):
    """the reference's second loop leaves the LAST interior point equal to the maximum
    (src/utils.f90:125-145); boundary planes are not scanned (do k = 2, nz-1 ...)"""
    n = (12, 10, 9)
    cfg = gpu.make_config(*n, 0.1, 0.1, 0.1, dt=0.5)
    ses = gpu.Session(cfg)
    z = np.zeros(n, order="F")
    old = [z.copy(order="F") for _ in range(3)]
    new = [z.copy(order="F") for _ in range(3)]
    new[0][3, 4, 5] = 2.0
    new[0][7, 2, 6] = -2.0       # same magnitude, later in array order -> reported
    new[0][0, 0, 0] = 50.0       # boundary: ignored
    new[1][5, 5, 1] = 1.0
    new[2][n[0] - 1, 3, 3] = 9.0  # boundary: ignored -> all-zero interior: last interior point
    ses.set(ux=old[0], uy=old[1], uz=old[2])
    ses.old_values()
    ses.set(ux=new[0], uy=new[1], uz=new[2])
    got = ses.calculate_residuals(0.5, 1.0, 1.0)
    exp = O.calculate_residuals(*new, *old, 0.5, 1.0, 1.0)
    assert np.array_equal(got, exp), (got, exp)
    assert list(got[6:9]) == [8, 3, 7] and list(got[9:12]) == [6, 6, 2]
    assert list(got[12:15]) == [n[0] - 1, n[1] - 1, n[2] - 1] and got[5] == 0.0
    ses.close()


@pytest.mark.parametrize("bc,n,nscr", [((1, 1, 1), (40, 33, 29), 1), ((0, 0, 0), (33, 32, 36), 0),
                                       ((0, 1, 0), (64, 20, 41), 1)])
def test_step_diagnostics_match_oracle(gpu, O, bc, n, nscr):
    """o3d_s_step_diagnostics = divergence + function_stats of u* and u, minval/maxval, CFL
    (src/osinco3d_main.f90:116-128) in two fused passes.  Divergence values are bit-identical to
    the oracle's, so min / max / position are exact; the mean is a differently associated sum."""
    d = tuple((PI if b else 2 * PI) / (m - 1) for b, m in zip(bc, n))
    g = O.grid(*n, *d, bc)
    ux, uy, uz, pp, phi = O.init_tgv(g, nscr=1)
    dt = 0.02 * d[0]
    cfg = gpu.make_config(*n, *d, bc=bc, re=400.0, dt=dt, itscheme=3, nscr=nscr, omega=1.7,
                          eps=1e-7, kmax=2000)
    ses = gpu.Session(cfg)
    ses.set(ux=ux, uy=uy, uz=uz, pp=pp, phi=phi)
    for _ in range(3):
        ses.step()
    got = ses.step_diagnostics()
    f = {k: ses.download(k) for k in ("ux", "uy", "uz", "ux_pred", "uy_pred", "uz_pred", "phi")}
    for key, names in (("divu_pred", ("ux_pred", "uy_pred", "uz_pred")), ("divu", ("ux", "uy", "uz"))):
        div = O.divergence(g, *[f[k] for k in names], 1)
        exp = O.function_stats(div)
        assert got[key][0] == exp[0] and got[key][1] == exp[1], (key, got[key], exp)
        assert list(got[key][3:]) == list(exp[3:]), (key, got[key], exp)
        assert abs(got[key][2] - exp[2]) <= 1e-12 * np.max(np.abs(div)), (key, got[key][2], exp[2])
    for q, k in enumerate(("ux", "uy", "uz")):
        assert got["umin"][q] == f[k].min() and got["umax"][q] == f[k].max()
        assert got["cfl"][q] == np.abs(f[k]).max() * dt / d[q]        # src/utils.f90:199-201
    if nscr:
        assert got["phi"] == [f["phi"].min(), f["phi"].max()]
    # the diagnostics must not disturb the state: the next steps equal an undisturbed run
    ref = gpu.Session(cfg)
    ref.set(ux=ux, uy=uy, uz=uz, pp=pp, phi=phi)
    for _ in range(4):
        ref.step()
    ses.step()
    for k in ("ux", "uy", "uz", "pp"):
        assert np.array_equal(ses.download(k), ref.download(k)), k
    ref.close()
    ses.close()
