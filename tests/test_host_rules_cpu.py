"""The index rules shared by every kernel (csrc/o3d_common.cuh: map_index, image_offsets, the
padded layout) and the host-side launch planning (csrc/kernels.h: pick_zchunk), checked on the
CPU: tests/cpu/host_rules_test.cu is compiled by nvcc (host code only, no CUDA call) and run.
Main invariant: the ghost images a producer kernel stores with each interior point are, cell for
cell, what the ghost-fill rule of the closures of src/derivation.f90 would have written."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_rules(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "host_rules_test")
    src = os.path.join(ROOT, "tests", "cpu", "host_rules_test.cu")
    r = subprocess.run([nvcc, "-std=c++17", "-O1", "-gencode", "arch=compute_100a,code=sm_100a",
                        "-o", exe, src], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "host rules OK" in r.stdout, r.stdout + r.stderr
