"""GPU parity AT THE SIZES AND LAUNCH GEOMETRIES bench.py RUNS (VERDICT r1 "What's weak" 1).

The operator tests compare with the oracle on grids of <= 70 points per axis: one or two tiles,
one z chunk.  Here the device-resident session -- the path bench.py times -- is driven stage by
stage at BASELINE.json's sizes (256^3 free-slip DNS = configs[1]; the shipped 129^3 LES and
241 x 241 x 81 mixing-layer grids; a 512 x 512 slab with the tile count of the 512^3 north-star
grid), so the 8 x 32 x (8..16)-CTA launches with many z chunks from pick_zchunk, the split ring and
the producer-written ghost images are compared FIELD BY FIELD, bit for bit, with the oracle:

  predict_velocity  src/integration.f90:14-197   u*, fu*(:,:,:,1), nu_t          bitwise
  divergence / dt   src/integration.f90:234-239  Poisson right-hand side         bitwise
  correct_velocity  src/integration.f90:257-330  u^{n+1} from the GPU's own pp   bitwise
  transeq           src/integration.f90:332-468  fphi bitwise, phi <= 1e-13 (three global sums)

The Poisson iterate itself is ordering-dependent (red-black vs lexicographic, stated in
DESIGN.md section 2) and is pinned at these sizes by the converged-solution and golden-history
tests; here it is checked to satisfy the reference's own exit criterion.
"""
import numpy as np
import pytest

from conftest import rel_max

pytestmark = pytest.mark.gpu
PI = 3.141592653589793


def perturbed_tgv(O, g, nscr):
    """TGV plus a smooth non-symmetric perturbation with uz != 0: a pure TGV has uz = 0 and is
    symmetric about many planes, which would hide parity / halo mistakes"""
    ux, uy, uz, pp, phi = O.init_tgv(g, nscr=nscr)
    x = (g.dx * np.arange(g.nx))[:, None, None]
    y = (g.dy * np.arange(g.ny))[None, :, None]
    z = (g.dz * np.arange(g.nz))[None, None, :]
    ux = np.asfortranarray(ux + 0.1 * np.sin(2 * x + 0.3) * np.cos(y) * np.cos(3 * z + 0.1))
    uy = np.asfortranarray(uy + 0.05 * np.cos(x) * np.sin(2 * y + 0.2) * np.cos(z))
    uz = np.asfortranarray(uz + 0.2 * np.cos(x + 0.1) * np.cos(2 * y) * np.sin(z + 0.4))
    return ux, uy, uz, pp, phi


CASES = {
    # BASELINE configs[1]: what `python bench.py` times
    "tgv_dns_256_freeslip": dict(shape=(256, 256, 256), bc=(1, 1, 1), re=1600.0, iles=0, cs=0.0,
                                 nscr=0, omega=1.887, eps=1e-4, idyn=0, dtf=0.05),
    # examples/tgv_re2500_les as shipped
    "tgv_les_129_freeslip": dict(shape=(129, 129, 129), bc=(1, 1, 1), re=2500.0, iles=1, cs=0.17,
                                 nscr=0, omega=1.999, eps=1e-6, idyn=1, dt=5e-4),
    # examples/mixing_layer_re3000_les grid and closures (x, z periodic; y free-slip), LES + scalar
    "mixing_layer_241x241x81_pfp": dict(shape=(241, 241, 81), bc=(0, 1, 0), re=3000.0, iles=1,
                                        cs=0.15, nscr=1, omega=1.999, eps=1e-5, idyn=1, dt=1.5e-3),
    # 16 x 64 tiles per plane as at 512^3 (the north-star grid), thin in z to keep the oracle fast
    "tgv_dns_512x512x40_freeslip": dict(shape=(512, 512, 40), bc=(1, 1, 1), re=1600.0, iles=0,
                                        cs=0.0, nscr=0, omega=1.887, eps=1e-4, idyn=0, dtf=0.05),
    # all-periodic odd extents (coplanar-jet closures): wrap ghosts + the seam SOR path
    "periodic_257x129x65": dict(shape=(257, 129, 65), bc=(0, 0, 0), re=2200.0, iles=0, cs=0.0,
                                nscr=0, omega=1.35, eps=1e-5, idyn=0, dtf=0.07),
}


@pytest.mark.parametrize("case", sorted(CASES))
def test_session_stages_bit_exact_at_bench_geometry(gpu, O, case):
    c = CASES[case]
    shape, bc = c["shape"], c["bc"]
    L = [PI if b else 2 * PI for b in bc]
    d = [L[a] / (shape[a] - 1) for a in range(3)]
    g = O.grid(*shape, *d, bc)
    dt = c["dt"] if "dt" in c else c["dtf"] * min(d)
    nscr, iles = c["nscr"], c["iles"]
    delta = (d[0] * d[1] * d[2]) ** (1.0 / 3.0)
    ux, uy, uz, pp, phi = perturbed_tgv(O, g, nscr)
    rng = np.random.default_rng(1234)
    # AB3 history: levels 2 and 3 hold smooth-ish data of the size of a real right-hand side
    hist = [np.asfortranarray(0.3 * rng.standard_normal(shape + (3,))) for _ in range(4)]
    cfg = gpu.make_config(*shape, *d, bc=bc, re=c["re"], cs=c["cs"], dt=dt, itscheme=3, iles=iles,
                          nscr=nscr, sc=1.0, omega=c["omega"], eps=c["eps"], kmax=400,
                          idyn=c["idyn"])
    ses = gpu.Session(cfg)
    ses.set(ux=ux, uy=uy, uz=uz, pp=pp)
    for comp, nm in enumerate(("fux", "fuy", "fuz")):
        for lvl in (2, 3):
            ses.upload("%s%d" % (nm, lvl), hist[comp][..., lvl - 1])
    if nscr:
        ses.upload("phi", phi)
        for lvl in (2, 3):
            ses.upload("fphi%d" % lvl, hist[3][..., lvl - 1])
    itime = 5      # AB3 branch of src/integration.f90:84-98

    # ---- predict_velocity ----
    ses.predict_velocity(itime)
    fo = [h.copy(order="F") for h in hist[:3]]
    ref = O.predict_velocity(g, ux, uy, uz, *fo, c["re"], dt, itime, 3, iles, c["cs"], delta)
    up_gpu = [ses.download(k) for k in ("ux_pred", "uy_pred", "uz_pred")]
    for a, b, nm in zip(up_gpu, ref[:3], ("ux_pred", "uy_pred", "uz_pred")):
        assert np.array_equal(a, b), (case, nm, rel_max(a, b))
    for comp, nm in enumerate(("fux", "fuy", "fuz")):
        for lvl in (1, 2, 3):      # new f and the shifted history, src/integration.f90:176-188
            got = ses.download("%s%d" % (nm, lvl))
            assert np.array_equal(got, fo[comp][..., lvl - 1]), (case, nm, lvl)
    if iles:
        assert np.array_equal(ses.download("nu_t"), ref[3]), (case, "nu_t")
    del fo

    # ---- correct_pression: right-hand side bitwise; iterate meets the reference's exit test ----
    iters, dmax = ses.correct_pression()
    rhs_ref = np.asfortranarray(O.divergence(g, *ref[:3], 1) / dt)   # :234-239
    got = ses.download("rhs")
    assert np.array_equal(got, rhs_ref), (case, "rhs", rel_max(got, rhs_ref))
    # Fortran `iter` after the loop: sweeps done, or kmax + 1 when the loop ran out (:53,:123)
    assert 1 <= iters <= 401 and np.isfinite(dmax)
    if iters <= 400:
        assert dmax < c["eps"] or iters > 1      # dmax < eps exit (:110) or the stall exit (:111)
    pp_gpu = ses.download("pp")
    assert np.isfinite(pp_gpu).all()
    del rhs_ref, got

    # ---- correct_velocity from the GPU's own pressure ----
    ses.correct_velocity()
    ur = O.correct_velocity(g, *up_gpu, pp_gpu, dt)
    assert ur[3] == 0
    u_gpu = [ses.download(k) for k in ("ux", "uy", "uz")]
    for a, b, nm in zip(u_gpu, ur[:3], ("ux", "uy", "uz")):
        assert np.array_equal(a, b), (case, nm, rel_max(a, b))

    # ---- transeq with the corrected velocity (src/osinco3d_main.f90:112-115) ----
    if nscr:
        ses.transeq(itime)
        fphi = hist[3].copy(order="F")
        phi_o = phi.copy(order="F")
        O.transeq(g, phi_o, *u_gpu, fphi, c["re"], 1.0, dt, itime, 3, iles,
                  ref[3] if iles else np.zeros(shape, order="F"))
        for lvl in (1, 2, 3):
            assert np.array_equal(ses.download("fphi%d" % lvl), fphi[..., lvl - 1]), (case, lvl)
        assert np.max(np.abs(ses.download("phi") - phi_o)) < 1e-13
    ses.close()
