"""include/o3d_b200.hpp -- the C++ host-side mirror of the reference's module interfaces -- in a
compiled program (tests/cpu/hpp_mirror_test.cpp, g++ -Wall -Wextra, linked against
libo3d_b200.so): it compiles cleanly, schemes() throws O3D_ERR_BC where the reference stops, and
without a CUDA device every module procedure throws O3D_ERR_NO_DEVICE (no CPU fallback).
The device half of the program runs in tests/test_gpu_zz_reference_source.py."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_hpp_test(tmp_path):
    gxx = shutil.which("g++")
    if not gxx:
        pytest.skip("g++ not available")
    exe = str(tmp_path / "hpp_mirror_test")
    libdir = os.path.join(ROOT, "osinco3d_b200", "lib")
    r = subprocess.run([gxx, "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-o", exe,
                        os.path.join(ROOT, "tests", "cpu", "hpp_mirror_test.cpp"), "-L" + libdir,
                        "-lo3d_b200", "-Wl,-rpath," + libdir], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    return exe


def test_cpp_mirror_compiles_and_fails_loudly_without_a_device(tmp_path, built_lib):
    exe = build_hpp_test(tmp_path)
    r = subprocess.run([exe, "nodevice"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "hpp mirror OK" in r.stdout, r.stdout + r.stderr


def test_cpp_mirror_covers_the_reference_procedures():
    """one wrapper per hot-path procedure of the replaced modules, under the reference's names"""
    import json
    import re
    sig = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_signatures.json")))
    txt = open(os.path.join(ROOT, "include", "o3d_b200.hpp")).read()
    skip = {"calculate_tau_ij", "calculate_dtau_ij_dxj", "v_cycle", "gauss_seidel", "compute_residual",
            "restrict_full_weighting", "prolongation_add", "dery1d"}
    for mod, ref in sig.items():
        if mod.startswith("_"):     # not a module -> namespace entry (output_b200 procedures)
            continue
        m = re.search(r"namespace %s \{(.*?)\}  // namespace %s" % (mod, mod), txt, re.S)
        assert m, mod
        body = m.group(1).lower()
        for proc in ref["procedures"]:
            if proc in skip:
                continue
            assert re.search(r"\b%s\b" % proc, body), (mod, proc)
