"""Session-API contracts that no reference routine states but a driver relies on (review
findings of round 1): history uploads address LOGICAL levels whatever the internal buffer
rotation, and a diverged state is never handed out silently."""
import numpy as np
import pytest

from conftest import rand_field

pytestmark = pytest.mark.gpu
PI = 3.141592653589793


def _session(gpu, O, n=24, **kw):
    d = PI / (n - 1)
    g = O.grid(n, n, n, d, d, d, (1, 1, 1))
    ux, uy, uz, pp, phi = O.init_tgv(g)
    cfg = gpu.make_config(n, n, n, d, d, d, re=400.0, dt=0.02 * d, omega=1.7, eps=1e-6, kmax=500,
                          **kw)
    ses = gpu.Session(cfg)
    ses.set(ux=ux, uy=uy, uz=uz, pp=pp)
    return ses, (ux, uy, uz, pp)


@pytest.mark.parametrize("order", [(1, 2, 3), (2, 3, 1), (3, 1, 2)])
@pytest.mark.parametrize("itscheme", [2, 3])
def test_history_upload_in_any_order_after_rotation(gpu, O, order, itscheme):
    """after a step, levels 1 and 2 of fu? share one physical buffer (the reference's copy
    fu(:,:,:,2) = fu(:,:,:,1), src/integration.f90:176-188, is a pointer rotation): a restart
    that uploads the three levels in any order must read back exactly what it uploaded"""
    ses, _ = _session(gpu, O, itscheme=itscheme)
    for _ in range(3):
        ses.step()
    shape = ses.shape
    new = {lvl: rand_field(shape, 100 + lvl) for lvl in (1, 2, 3)}
    for lvl in order:
        ses.upload("fuy%d" % lvl, new[lvl])
    for lvl in (1, 2, 3):
        assert np.array_equal(ses.download("fuy%d" % lvl), new[lvl]), (order, lvl)
    # and the next predictor consumes levels 2 and 3 as uploaded: same result as a fresh session
    # primed with the same state
    ref, (ux, uy, uz, pp) = _session(gpu, O, itscheme=itscheme)
    for k in ("ux", "uy", "uz"):
        ref.upload(k, ses.download(k))
    for comp in ("fux", "fuy", "fuz"):
        for lvl in (2, 3):
            ref.upload("%s%d" % (comp, lvl), ses.download("%s%d" % (comp, lvl)))
    ses.predict_velocity(7)
    ref.predict_velocity(7)
    for k in ("ux_pred", "uy_pred", "uz_pred", "fuy1", "fuy2", "fuy3"):
        assert np.array_equal(ses.download(k), ref.download(k)), k
    ses.close()
    ref.close()


def test_download_after_step_reports_divergence(gpu, O):
    """o3d_step defers the NaN / >1000 guard of correct_velocity (src/integration.f90:309-325); a
    download right after it must report O3D_ERR_DIVERGED instead of returning silently"""
    ses, (ux, uy, uz, pp) = _session(gpu, O)
    bad = ux.copy(order="F")
    bad[5, 6, 7] = 5.0e4
    ses.upload("ux", bad)
    ses.step()
    with pytest.raises(gpu.O3DError) as e:
        ses.download("ux")
    assert e.value.code == gpu._lib.ERR_DIVERGED
    ses.close()


def test_output_of_a_diverged_state_is_refused(gpu, O, tmp_path):
    """the reference stops inside correct_velocity, before any save_fields / write_all_data"""
    ses, (ux, uy, uz, pp) = _session(gpu, O)
    bad = uy.copy(order="F")
    bad[3, 3, 3] = np.nan
    ses.upload("uy", bad)
    ses.step()
    n = ses.shape[0]
    xyz = [np.arange(n, dtype=float)] * 3
    with pytest.raises(gpu.O3DError) as e:
        ses.save_fields(str(tmp_path / "fields.bin"), 0.0, *xyz)
    assert e.value.code == gpu._lib.ERR_DIVERGED
    assert not (tmp_path / "fields.bin").exists()
    ses.close()


def test_plane_range_upload_download_equal_whole_field_copies(gpu, O):
    """o3d_upload_planes / o3d_download_planes (drivers that fill a 1024^3 slab in pieces) move
    exactly the planes they name; a field assembled from ragged chunks equals the whole-field
    upload bit for bit, ghost state included (a step from either gives the same result)"""
    ses, (ux, uy, uz, pp) = _session(gpu, O, n=40)
    ref, _ = _session(gpu, O, n=40)
    rng = np.random.default_rng(5)
    f = {k: np.asfortranarray(v + 0.01 * rng.standard_normal(v.shape))
         for k, v in (("ux", ux), ("uy", uy), ("uz", uz), ("pp", pp))}
    ref.set(**f)
    for k, v in f.items():
        for k0, nk in ((0, 7), (7, 1), (8, 19), (27, 13)):
            ses.upload_planes(k, v[:, :, k0:k0 + nk], k0)
    for k, v in f.items():
        assert np.array_equal(ses.download(k), v), k
        assert np.array_equal(ses.download_planes(k, 11, 6), v[:, :, 11:17]), k
    for _ in range(3):
        assert ses.step() == ref.step()
    for k in ("ux", "uy", "uz", "pp"):
        assert np.array_equal(ses.download(k), ref.download(k)), k
    with pytest.raises(gpu.O3DError):
        ses.upload_planes("ux", f["ux"][:, :, :5], 38)
    ses.close()
    ref.close()
