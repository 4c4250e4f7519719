"""Logic the kernels share with the host, checked on the CPU: the .cu files under tests/cpu/ are
compiled by nvcc (host code; they make no CUDA call) and run.

host_rules_test.cu   csrc/o3d_common.cuh (map_index, image_offsets, padded layout) and
                     csrc/kernels.h (pick_zchunk): the ghost images a producer kernel stores with
                     each interior point are, cell for cell, what the ghost-fill rule of the
                     closures of src/derivation.f90 would have written.
sor_classes_test.cu  csrc/sor_kernels.cu (nbr_idx, seam_pop): every colour / seam class of the
                     red-black SOR sweeps is an independent set under the neighbour rule of
                     src/poisson.f90, for all three boundary variants and all extent parities.
"""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name,ok", [("host_rules_test", "host rules OK"),
                                     ("sor_classes_test", "sor classes OK")])
def test_host_compiled_rules(tmp_path, name, ok):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / name)
    src = os.path.join(ROOT, "tests", "cpu", name + ".cu")
    r = subprocess.run([nvcc, "-std=c++17", "-O1", "-fmad=false", "-gencode",
                        "arch=compute_100a,code=sm_100a", "-o", exe, src], capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and ok in r.stdout, r.stdout + r.stderr


def test_multigrid_transfer_tables_equal_the_design_model(tmp_path, built_lib):
    """csrc/multigrid.cu (coarse_extent, axis_tables: non-nested vertex-centred coarsening for
    mirrored / periodic axes of any parity) against oracle/mg_model.py, entry for entry --
    including the exact 0 / 0.5 weights of the nested cases -- for the shipped extents"""
    import json
    import sys

    import numpy as np
    sys.path.insert(0, ROOT)
    from oracle import mg_model as mm
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "mg_tables_dump")
    libdir = os.path.join(ROOT, "osinco3d_b200", "lib")
    r = subprocess.run([nvcc, "-std=c++17", "-O1", "-fmad=false", "-gencode",
                        "arch=compute_100a,code=sm_100a", "-o", exe,
                        os.path.join(ROOT, "tests", "cpu", "mg_tables_dump.cu"), "-L" + libdir,
                        "-lo3d_b200", "-Xlinker", "-rpath", "-Xlinker", libdir],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = [json.loads(ln) for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 32
    for t in lines:
        n, mode = t["n"], t["mode"]
        assert t["nc"] == mm.coarse_extent(n, mode), (n, mode)
        if not t["nc"]:
            continue
        D, c0, w, ridx, rw = mm.axis_tables(n, 0.0371, mode, t["nc"])
        assert t["D"] == D, (n, mode)
        assert t["c0"] == list(c0) and t["w"] == list(w), (n, mode)
        assert t["ridx"] == list(ridx.ravel()), (n, mode)
        assert np.array_equal(np.array(t["rw"]), rw.ravel()), (n, mode)


def test_kernel_stencil_expressions_reproduce_the_reference_routines(tmp_path):
    """d1_expr / d2_expr + make_coef + map_index (the arithmetic every stencil kernel performs,
    csrc/o3d_common.cuh) evaluated on the host == the 18 routines of src/derivation.f90 as
    executed from the reference source (tests/golden/hotpath.npz), bit for bit, boundary planes
    included"""
    import numpy as np
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "stencil_rules_test")
    r = subprocess.run([nvcc, "-std=c++17", "-O1", "-fmad=false", "-Xcompiler", "-ffp-contract=off",
                        "-gencode", "arch=compute_100a,code=sm_100a", "-o", exe,
                        os.path.join(ROOT, "tests", "cpu", "stencil_rules_test.cu")],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr
    gold = np.load(os.path.join(ROOT, "tests", "golden", "hotpath.npz"))
    dx, dy, dz = [float(v) for v in gold["params"][5:8]]
    for tag, key in (("", "in_pp"), ("_small", "in_small")):
        f = np.asfortranarray(gold[key])
        fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
        f.ravel(order="F").tofile(fin)
        r = subprocess.run([exe, fin] + [str(n) for n in f.shape] + [repr(dx), repr(dy), repr(dz), fout],
                           capture_output=True, text=True, timeout=120)
        assert r.returncode == 0, r.stdout + r.stderr
        out = np.fromfile(fout).reshape((18,) + f.shape[::-1])
        q = 0
        for axis in "xyz":
            for order in (1, 2):
                for closure in ("_00", "p_11", "i_11"):
                    name = "der" + axis * order + closure
                    got = out[q].transpose(2, 1, 0)            # file is (k, j, i) C order
                    assert np.array_equal(got, gold[name + tag]), (name, tag)
                    q += 1


def test_kernel_rhs_arithmetic_reproduces_predict_velocity(tmp_path):
    """the per-point sequence of the fused RHS + nu_t + predictor kernel -- d1_expr / d2_expr on
    natural-parity ghost cells, smagorinsky(), rhs_expr(), predictor_expr(), all from
    csrc/o3d_common.cuh -- on the host == predict_velocity (src/integration.f90:14-197) and
    calculate_nu_t (src/les_turbulence.f90:10-97) as executed from the reference source, bit for
    bit, for every boundary configuration, DNS and LES, Euler and AB2 coefficients"""
    import numpy as np
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "step_rules_test")
    r = subprocess.run([nvcc, "-std=c++17", "-O1", "-fmad=false", "-Xcompiler", "-ffp-contract=off",
                        "-gencode", "arch=compute_100a,code=sm_100a", "-o", exe,
                        os.path.join(ROOT, "tests", "cpu", "step_rules_test.cu")],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr
    gold = np.load(os.path.join(ROOT, "tests", "golden", "hotpath.npz"))
    re_, sc, cs, dt, delta, dx, dy, dz = [float(v) for v in gold["params"]]
    shape = gold["in_ux"].shape
    N = int(np.prod(shape))
    u = [np.asfortranarray(gold["in_u" + c]) for c in "xyz"]
    fu = [np.asfortranarray(gold["in_fu" + c]) for c in "xyz"]
    blob = np.concatenate([a.ravel(order="F") for a in u] +
                          [a[..., 1].ravel(order="F") for a in fu] +
                          [a[..., 2].ravel(order="F") for a in fu])
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    blob.tofile(fin)
    configs = {"ppp": (0, 0, 0, 0), "fff": (1, 1, 1, 0), "pfp": (0, 1, 0, 0), "pfp2d": (0, 1, 0, 1),
               "ffp": (1, 1, 0, 0), "ppf": (0, 0, 1, 0), "fpf": (1, 0, 1, 0)}

    def run(cfg, iles, level):
        bx, by, bz, sim2d = configs[cfg]
        coef = [repr(float(gold[k][level])) for k in ("adt", "bdt", "cdt")]
        r = subprocess.run([exe, fin, fout] + [str(n) for n in shape] +
                           [repr(dx), repr(dy), repr(dz), str(bx), str(by), str(bz), str(sim2d),
                            str(iles), repr(re_), repr(cs), repr(delta)] + coef,
                           capture_output=True, text=True, timeout=120)
        assert r.returncode == 0, r.stdout + r.stderr
        out = np.fromfile(fout).reshape((11, N))
        return [o.reshape(shape, order="F") for o in out]

    for cfg in configs:
        for iles in (0, 1):                       # itime = 1: Euler coefficients
            out = run(cfg, iles, 0)
            for c, a in zip("xyz", out[:3]):
                assert np.array_equal(a, gold["%s_pred_les%d_it1_u%s" % (cfg, iles, c)]), (cfg, iles, c)
            assert np.array_equal(out[3], gold["%s_pred_les%d_it1_nu_t" % (cfg, iles)]), (cfg, iles)
        # rotational / calculate_Q_criterion (src/differential_operators.f90:40-108)
        assert np.array_equal(out[7], gold[cfg + "_q"]), cfg
        for c, a in zip("xyz", out[8:11]):
            assert np.array_equal(a, gold["%s_rot%s" % (cfg, c)]), (cfg, c)
        out = run(cfg, 0, 1)                      # itscheme = 2: AB2 coefficients, f1 stored
        assert np.array_equal(out[0], gold[cfg + "_pred_sch2_ux"]), cfg
        assert np.array_equal(out[4], gold[cfg + "_pred_sch2_fux"][..., 0]), cfg


def test_kernel_sor_functions_reproduce_the_reference_solvers(tmp_path):
    """nbr_idx + sor_pnew_ref + sor_relax + sor_control_step (csrc/sor_kernels.cu: what
    sor_wavefront_kernel and sor_control_kernel execute) in the reference's loop order on the host
    == poisson_solver_0000 / _0011 / _111111 as executed from the reference source: iterates,
    the loop variable after the loop, the dynamic omega and dmax, bit for bit"""
    import numpy as np
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "sor_sweep_test")
    r = subprocess.run([nvcc, "-std=c++17", "-O1", "-fmad=false", "-Xcompiler", "-ffp-contract=off",
                        "-gencode", "arch=compute_100a,code=sm_100a", "-o", exe,
                        os.path.join(ROOT, "tests", "cpu", "sor_sweep_test.cu")],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr
    gold = np.load(os.path.join(ROOT, "tests", "golden", "hotpath.npz"))
    dx, dy, dz = [float(v) for v in gold["params"][5:8]]
    shape = gold["in_pp"].shape
    N = int(np.prod(shape))
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    np.concatenate([np.asfortranarray(gold["in_pp"]).ravel(order="F"),
                    np.asfortranarray(gold["in_rhs"]).ravel(order="F")]).tofile(fin)
    # configuration -> the variant schemes() binds from the x and y flags
    variants = {"ppp": 0, "ppf": 0, "pfp": 1, "fff": 2, "ffp": 2}
    for cfg, variant in variants.items():
        for tag, (omega, eps, kmax, idyn) in (("fixed", (1.6, 1e-30, 12, 0)),
                                              ("dyn", (1.9, 2e-3, 400, 1))):
            r = subprocess.run([exe, fin, fout] + [str(n) for n in shape] +
                               [repr(dx), repr(dy), repr(dz), str(variant), repr(omega), repr(eps),
                                str(kmax), str(idyn)], capture_output=True, text=True, timeout=300)
            assert r.returncode == 0, r.stdout + r.stderr
            out = np.fromfile(fout)
            assert np.array_equal(out[:N].reshape(shape, order="F"), gold["%s_sor_%s_pp" % (cfg, tag)]), \
                (cfg, tag)
            assert np.array_equal(out[N:], gold["%s_sor_%s_scalars" % (cfg, tag)]), \
                (cfg, tag, out[N:], gold["%s_sor_%s_scalars" % (cfg, tag)])


def test_kernel_divergence_correction_and_transeq_arithmetic(tmp_path):
    """div_expr / corr_expr / transeq_* / clip01 / predictor_expr (csrc/o3d_common.cuh: what
    DivEpi, CorrEpi, TranseqEpi and transeq_clip_kernel call) on ghost cells on the host ==
    divergence, correct_velocity and transeq (clip + redistribution included) as executed from
    the reference source, bit for bit, on every boundary configuration"""
    import numpy as np
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "scalar_rules_test")
    r = subprocess.run([nvcc, "-std=c++17", "-O1", "-fmad=false", "-Xcompiler", "-ffp-contract=off",
                        "-gencode", "arch=compute_100a,code=sm_100a", "-o", exe,
                        os.path.join(ROOT, "tests", "cpu", "scalar_rules_test.cu")],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr
    gold = np.load(os.path.join(ROOT, "tests", "golden", "hotpath.npz"))
    re_, sc, cs, dt, delta, dx, dy, dz = [float(v) for v in gold["params"]]
    shape = gold["in_ux"].shape
    N = int(np.prod(shape))
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    flat = lambda a: np.asfortranarray(a).ravel(order="F")  # noqa: E731
    configs = {"ppp": (0, 0, 0, 0), "fff": (1, 1, 1, 0), "pfp": (0, 1, 0, 0), "pfp2d": (0, 1, 0, 1),
               "ffp": (1, 1, 0, 0), "ppf": (0, 0, 1, 0), "fpf": (1, 0, 1, 0)}

    def run(mode, fields, cfg, extra, nout):
        np.concatenate([flat(a) for a in fields]).tofile(fin)
        r = subprocess.run([exe, mode, fin, fout] + [str(n) for n in shape] +
                           [repr(dx), repr(dy), repr(dz)] + [str(v) for v in configs[cfg]] + extra,
                           capture_output=True, text=True, timeout=120)
        assert r.returncode == 0, (mode, cfg, r.stdout + r.stderr)
        return [o.reshape(shape, order="F") for o in np.fromfile(fout).reshape((nout, N))]

    u = [gold["in_u" + c] for c in "xyz"]
    for cfg in configs:
        for odd in (0, 1):
            got = run("div", u, cfg, [str(odd)], 1)[0]
            assert np.array_equal(got, gold["%s_divergence_odd%d" % (cfg, odd)]), (cfg, odd)
        got = run("corr", u + [gold["in_pp"]], cfg, [repr(dt)], 3)
        for c, a in zip("xyz", got):
            assert np.array_equal(a, gold["%s_corr_u%s" % (cfg, c)]), (cfg, c)
        for iles in (0, 1):      # first step of the golden sequence: Euler with the enlarged adt
            coef = [repr(40.0 * float(gold["adt"][0])), repr(float(gold["bdt"][0])),
                    repr(float(gold["cdt"][0]))]
            fphi = gold["in_fphi"]
            got = run("transeq", [gold["in_phi"]] + u + [gold[cfg + "_nu_t"], fphi[..., 1], fphi[..., 2]],
                      cfg, [str(iles), repr(re_), repr(sc)] + coef, 2)
            assert np.array_equal(got[0], gold["%s_transeq_les%d_it1_phi" % (cfg, iles)]), (cfg, iles)
