"""The main-loop harness (osinco3d_b200/host/o3d_mainloop.cpp: the cadence of
src/osinco3d_main.f90:97-188 over include/o3d_b200.hpp) on the device: its outputs/stats.dat --
written in the reference's '(17es21.12)' record format (src/IOfunctions.f90:552) every 25 steps --
against the reference's shipped history, and its output files against the session API."""
import json
import os
import re
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "reference_stats.json")))
PI = 3.141592653589793


def run(tmp_path, *opts):
    from osinco3d_b200 import build as b
    exe = b.build_mainloop()
    out = str(tmp_path / "run")
    r = subprocess.run([exe, "--out", out] + [str(o) for o in opts], capture_output=True, text=True,
                       timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return out, r.stdout


def test_mainloop_stats_dat_matches_the_shipped_dns_history(gpu, tmp_path):
    """examples/tgv_re1600_dns (185^3, cfl 0.05, omega 1.887, eps 1e-4): 100 steps -> 5 rows of
    stats.dat (t = 0 written by --stats-at-start, as the shipped file has it) against
    examples/tgv_re1600_dns/tgv_stats_re1600_dns.dat:18-22, E_k / enstrophy <= 1e-6 (north star),
    in fact <= 1e-7; and the text format is the reference's 17es21.12"""
    out, log = run(tmp_path, "--n", 185, "--steps", 100, "--stats-at-start", "--nsve", 100,
                   "--nfre", 50)
    lines = open(os.path.join(out, "outputs", "stats.dat")).read().splitlines()
    rows = GOLD["tgv_re1600_dns"]["rows"]
    assert len(lines) == 5 == len(rows)
    num = r"  [ -]\d\.\d{12}E[+-]\d{2}"      # es21.12: 21 columns
    for ln, ref in zip(lines, rows):
        assert len(ln) == 17 * 21 and re.fullmatch("(%s){17}" % num, ln), ln
        v = np.array([float(t) for t in ln.split()])
        assert abs(v[0] - ref[0]) <= 1e-12 * max(1.0, abs(ref[0]))
        for c in (1, 4):
            assert abs(v[c] - ref[c]) / ref[c] < 1e-7, (c, v[c], ref[c])
        for c in (2, 3):
            assert abs(v[c] - ref[c]) / ref[c] < 5e-7, (c, v[c], ref[c])
    # the cadence: per-step prints, residuals every 25 steps, files every nfre / nsve steps
    assert log.count("Iteration:") == 100 and log.count("* residuals") == 4
    assert log.count("* CFL") == 100
    n = 185
    for f in ("ux_0.bin", "ux_1.bin", "ux_2.bin", "pp_2.bin", "vort_2.bin", "qcrit_2.bin"):
        assert os.path.getsize(os.path.join(out, "outputs", f)) == 8 * n ** 3, f
    hdr = 8 + 12 + 3 * 8 * n
    assert os.path.getsize(os.path.join(out, "fields_000100.bin")) == hdr + 5 * 8 * n ** 3
    t = np.fromfile(os.path.join(out, "fields_000100.bin"), dtype=np.float64, count=1)[0]
    assert abs(t - rows[4][0]) < 1e-12


def test_mainloop_les_dynamic_omega_reference_sweep_order(gpu, tmp_path):
    """examples/tgv_re2500_les at 129^3, 25 steps in the reference's sweep order: row 1 of the
    shipped LES history within the 1e-6 history tolerance"""
    out, log = run(tmp_path, "--n", 129, "--steps", 25, "--re", 2500, "--dt", 5e-4, "--omega",
                   1.999, "--eps", 1e-6, "--idyn", 1, "--les", 0.17, "--wavefront", "--quiet")
    lines = open(os.path.join(out, "outputs", "stats.dat")).read().splitlines()
    assert len(lines) == 1
    v = np.array([float(t) for t in lines[0].split()])
    ref = GOLD["tgv_re2500_les"]["rows"][0]
    for c in (1, 2, 4):
        assert abs(v[c] - ref[c]) / ref[c] < 1e-6, (c, v[c], ref[c])
