"""osinco3d_b200/host/o3d_mainloop.cpp -- the reference's time loop (src/osinco3d_main.f90:97-188)
over the C++ mirror -- builds warning-free and, like everything else, fails loudly without a CUDA
device.  Its device run is tests/test_gpu_mainloop.py."""
import os
import subprocess


def test_mainloop_builds_and_fails_loudly_without_a_device(built_lib, tmp_path):
    from osinco3d_b200 import build as b
    exe = b.build_mainloop()
    assert exe and os.path.exists(exe)
    if built_lib.device_count() > 0:
        return
    r = subprocess.run([exe, "--n", "16", "--steps", "1", "--out", str(tmp_path / "run")],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 1 and "O3D" not in r.stdout
    assert "no CUDA device" in r.stderr or "device" in r.stderr.lower(), r.stderr


def test_mainloop_rejects_unknown_options(built_lib):
    from osinco3d_b200 import build as b
    r = subprocess.run([b.build_mainloop(), "--bogus"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 2 and "unknown option" in r.stderr
