"""The CUDA path against the reference SOURCE: the stencil-only outputs of tests/golden/hotpath.npz
(the reference's Fortran statements translated mechanically into NumPy and executed, see
tests/golden/make_hotpath_golden.py) through the drop-in C ABI, BIT FOR BIT -- the 20 `derivation`
routines, divergence / curl / Q, nu_t, predict_velocity with its history shifts, correct_velocity.
(The SOR solvers and transeq are pinned to the same vectors through the oracle on the CPU,
tests/test_oracle_reference_source.py, and to the oracle on the GPU in test_gpu_poisson.py /
test_gpu_operators.py: their device reductions / sweep order are not bitwise those of a serial
loop.)  Named zz so that it runs last."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "hotpath.npz"))
SHAPE = GOLD["in_ux"].shape
RE, SC, CS, DT, DELTA, DX, DY, DZ = [float(v) for v in GOLD["params"]]
CONFIGS = {"ppp": (0, 0, 0), "fff": (1, 1, 1), "pfp": (0, 1, 0), "ffp": (1, 1, 0), "ppf": (0, 0, 1)}
DER = ["derx_00", "derxp_11", "derxi_11", "dery_00", "deryp_11", "deryi_11", "derz_00", "derzp_11",
       "derzi_11", "derxx_00", "derxxp_11", "derxxi_11", "deryy_00", "deryyp_11", "deryyi_11",
       "derzz_00", "derzzp_11", "derzzi_11", "derz_2dsim", "derzz_2dsim"]


def inp(name):
    return np.asfortranarray(GOLD["in_" + name]).copy(order="F")


@pytest.fixture
def M(gpu):
    from osinco3d_b200 import modules
    yield modules
    modules.schemes(1, 1, 1, 1, 1, 1)


@pytest.mark.parametrize("name", DER)
def test_stencil_routine_from_the_reference_source(M, name):
    d = {"x": DX, "y": DY, "z": DZ}[name[3]]
    for tag in ("", "_small"):
        f = inp("pp") if not tag else inp("small")
        assert np.array_equal(getattr(M, name)(f, d), GOLD[name + tag]), (name, tag)


@pytest.mark.parametrize("cfg", list(CONFIGS))
def test_operators_and_predictor_from_the_reference_source(M, cfg):
    bc = CONFIGS[cfg]
    M.schemes(bc[0], bc[0], bc[1], bc[1], bc[2], bc[2])
    u = [inp(k) for k in ("ux", "uy", "uz")]
    for odd in (0, 1):
        assert np.array_equal(M.divergence(*u, DX, DY, DZ, odd),
                              GOLD["%s_divergence_odd%d" % (cfg, odd)]), odd
    for c, a in zip("xyz", M.rotational(*u, DX, DY, DZ)):
        assert np.array_equal(a, GOLD["%s_rot%s" % (cfg, c)]), c
    assert np.array_equal(M.calculate_Q_criterion(*u, DX, DY, DZ), GOLD[cfg + "_q"])
    assert np.array_equal(M.calculate_nu_t(*u, DX, DY, DZ, CS, DELTA), GOLD[cfg + "_nu_t"])
    adt, bdt, cdt = (list(GOLD[k]) for k in ("adt", "bdt", "cdt"))
    assert (adt, bdt, cdt) == tuple(M.ab_coefficients(DT))
    for iles in (0, 1):
        f = [inp("fu" + c) for c in "xyz"]
        for itime in (1, 2, 3):
            got = M.predict_velocity(*u, *f, RE, adt, bdt, cdt, itime, 3, DX, DY, DZ, iles, CS,
                                     DELTA)
            for c, a in zip("xyz", got[:3]):
                assert np.array_equal(a, GOLD["%s_pred_les%d_it%d_u%s" % (cfg, iles, itime, c)]), \
                    (iles, itime, c)
            assert np.array_equal(got[3], GOLD["%s_pred_les%d_it%d_nu_t" % (cfg, iles, itime)])
        for c, a in zip("xyz", f):
            assert np.array_equal(a, GOLD["%s_pred_les%d_fu%s" % (cfg, iles, c)]), (iles, c)
    got = M.correct_velocity(*u, inp("pp"), DT, DX, DY, DZ)
    for c, a in zip("xyz", got[:3]):
        assert np.array_equal(a, GOLD["%s_corr_u%s" % (cfg, c)]), c
    assert got[3] is False


def test_cpp_host_mirror_on_the_device(gpu, tmp_path):
    """include/o3d_b200.hpp from a compiled C++ program (tests/cpu/hpp_mirror_test.cpp): the 18
    derivative routines, divergence and correct_velocity through the namespaces that mirror the
    reference's modules, against the reference-source vectors, bit for bit"""
    from test_hpp_mirror_cpu import build_hpp_test
    import subprocess
    exe = build_hpp_test(tmp_path)
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    fields = [inp("pp"), inp("ux"), inp("uy"), inp("uz"), inp("pp")]
    np.concatenate([a.ravel(order="F") for a in fields]).tofile(fin)
    r = subprocess.run([exe, "run", fin, fout] + [str(n) for n in SHAPE] +
                       [repr(DX), repr(DY), repr(DZ), repr(DT)], capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0 and "hpp mirror run OK" in r.stdout, r.stdout + r.stderr
    N = int(np.prod(SHAPE))
    out = [o.reshape(SHAPE, order="F") for o in np.fromfile(fout).reshape((22, N))]
    q = 0
    for axis in "xyz":
        for order in (1, 2):
            for closure in ("_00", "p_11", "i_11"):
                assert np.array_equal(out[q], GOLD["der" + axis * order + closure]), (axis, order, closure)
                q += 1
    assert np.array_equal(out[18], GOLD["fff_divergence_odd1"])
    for c, a in zip("xyz", out[19:22]):
        assert np.array_equal(a, GOLD["fff_corr_u" + c]), c
