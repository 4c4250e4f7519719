"""Host logic of the copy pipeline (csrc/pipeline.cu) without a device: the schedule
o3d_pipeline_plan reports -- the one the pipelined o3d_predict_velocity / o3d_correct_velocity
execute -- is replayed plane by plane: every plane a chunk's kernels read (its own planes +- 3,
through the z closure) must already have been uploaded, or be a ghost plane that was filled from
uploaded planes, when the chunk is issued; every chunk is issued exactly once."""
import ctypes as C

import pytest


def plan(lib, nz, periodic):
    z = (C.c_int * 65)()
    after = (C.c_int * 64)()
    zfill = (C.c_int * 64)()
    n = lib.o3d_pipeline_plan(nz, periodic, z, after, zfill)
    return n, list(z[:n + 1]), list(after[:n]), list(zfill[:n])


@pytest.fixture
def plib(built_lib):
    lib = built_lib.lib()
    before = lib.o3d_get_pipeline()
    yield lib
    lib.o3d_set_pipeline(before)


@pytest.mark.parametrize("periodic", [0, 1])
@pytest.mark.parametrize("chunks", [2, 3, 4, 8, 16, 64, 1000])
@pytest.mark.parametrize("nz", [16, 17, 33, 64, 81, 129, 256, 257, 1024])
def test_schedule_only_reads_planes_that_have_landed(plib, nz, chunks, periodic):
    assert plib.o3d_set_pipeline(chunks) == 0
    n, z, after, zfill = plan(plib, nz, periodic)
    assert 2 <= n <= min(chunks, 64)
    assert z[0] == 0 and z[n] == nz
    assert all(z[c + 1] - z[c] >= 8 for c in range(n))
    # the count is the largest one <= the setting that keeps 8 planes per chunk
    assert n == min(chunks, 64, nz // 8)
    lo = hi = False
    issued = []
    for j in range(n):
        landed = z[j + 1]                       # planes [0, landed) are on the device
        for c in range(n):
            if c in issued or after[c] > j:
                continue
            if zfill[c] & 1:                    # low ghosts: wrap copies the last planes,
                src = range(nz - 3, nz) if periodic else range(1, 4)   # mirror planes 1..3
                assert all(q < landed for q in src), (c, j)
                lo = True
            if zfill[c] & 2:
                src = range(0, 3) if periodic else range(nz - 4, nz - 1)
                assert all(q < landed for q in src), (c, j)
                hi = True
            for q in range(z[c] - 3, z[c + 1] + 3):
                if q < 0:
                    assert lo, (c, j, q)
                elif q >= nz:
                    assert hi, (c, j, q)
                else:
                    assert q < landed, (c, j, q)
            issued.append(c)
    assert sorted(issued) == list(range(n))
    if not periodic:
        # nothing waits longer than it has to: chunk c runs right after upload c+1
        assert after == [min(c + 1, n - 1) for c in range(n)]
    else:
        assert after[0] == n - 1 and after[1:] == [min(c + 1, n - 1) for c in range(1, n)]


def test_thin_grids_and_the_off_switch(plib):
    plib.o3d_set_pipeline(8)
    assert plan(plib, 15, 0)[0] == 0            # < 2 chunks of 8 planes: plain path
    assert plan(plib, 7, 1)[0] == 0
    assert plan(plib, 16, 0)[0] == 2
    assert plan(plib, 40, 0)[0] == 5
    plib.o3d_set_pipeline(0)
    assert plib.o3d_get_pipeline() == 0
    assert plan(plib, 256, 0)[0] == 0
    plib.o3d_set_pipeline(1)
    assert plan(plib, 256, 0)[0] == 0           # one chunk = no pipeline
    assert plib.o3d_set_pipeline(-1) != 0       # O3D_ERR_INVALID
    assert plib.o3d_get_pipeline() == 1


def test_hostshift_switch(plib):
    before = plib.o3d_get_hostshift()
    assert before in (0, 1)
    assert plib.o3d_set_hostshift(1) == 0 and plib.o3d_get_hostshift() == 1
    assert plib.o3d_set_hostshift(0) == 0 and plib.o3d_get_hostshift() == 0
    assert plib.o3d_set_hostshift(2) != 0 and plib.o3d_set_hostshift(-1) != 0
    assert plib.o3d_get_hostshift() == 0
    plib.o3d_set_hostshift(before)


def test_e2e_byte_accounting_follows_the_switches(plib):
    from osinco3d_b200 import modules as M
    before = plib.o3d_get_hostshift()
    N = 1000
    plib.o3d_set_pipeline(16)
    plib.o3d_set_hostshift(0)
    assert M.e2e_bytes_per_step(N) == (17 * N * 8, 17 * N * 8)
    plib.o3d_set_hostshift(1)
    # predict: ux_pred, uy_pred, uz_pred + level 1 of fux, fuy, fuz (+ nu_t when LES);
    # correct_pression: pp;  correct_velocity: ux, uy, uz
    assert M.e2e_bytes_per_step(N) == (17 * N * 8, 10 * N * 8)
    assert M.e2e_bytes_per_step(N, iles=1) == (17 * N * 8, 11 * N * 8)
    plib.o3d_set_pipeline(0)                     # unpipelined calls download everything
    assert M.e2e_bytes_per_step(N) == (17 * N * 8, 17 * N * 8)
    plib.o3d_set_hostshift(before)


@pytest.mark.parametrize("tsan", [True, False])
def test_host_copier_order_protocol(tmp_path, tsan):
    """csrc/host_copier.h with the device replaced by threads (tests/cpu/host_copier_test.cpp):
    every history level and nu_t end up as src/integration.f90:176-188 leaves them, for
    itscheme 1 / 2 / 3, mirror and periodic z schedules, 1 - 5 workers; under ThreadSanitizer an
    unordered access of a worker and a "DMA" thread to the same host array fails the test."""
    import os
    import shutil
    import subprocess
    gxx = shutil.which("g++")
    if not gxx:
        pytest.skip("g++ not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "host_copier_test")
    cmd = [gxx, "-std=c++17", "-O1", "-g", "-Wall", "-Wextra", "-Werror", "-o", exe,
           os.path.join(root, "tests", "cpu", "host_copier_test.cpp"), "-lpthread"]
    if tsan:
        cmd.insert(1, "-fsanitize=thread")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    if tsan and r.returncode != 0 and "sanitize" in (r.stdout + r.stderr):
        pytest.skip("ThreadSanitizer runtime not available")
    assert r.returncode == 0, r.stdout + r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    if tsan and "FATAL: ThreadSanitizer" in r.stderr:
        pytest.skip("ThreadSanitizer cannot run in this sandbox: " + r.stderr.strip()[:120])
    assert r.returncode == 0 and "host copier OK" in r.stdout, r.stdout + r.stderr
    assert "ThreadSanitizer" not in r.stderr, r.stderr
