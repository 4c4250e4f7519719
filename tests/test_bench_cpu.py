"""bench.py contract checks that need no GPU: the reference arm (`--impl reference`, the oracle
port on the host) runs exactly W warm-up + K timed steps and prints the agreed JSON line; when
the full grid would not fit the time budget each step becomes a bounded sample (smaller grid);
non-zero ranks of a torchrun launch print nothing."""
import argparse
import io
import json
import os
import subprocess
import sys
from contextlib import redirect_stdout

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step",
        "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline",
        "e2e"}


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                        "--n", "24", "--steps", "3", "--warmup", "1"], capture_output=True,
                       text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert KEYS <= set(d)
    assert d["impl"] == "reference" and d["steps"] == 3 and d["warmup"] == 1
    assert d["unit"] == "Mpts*steps/s" and d["higher_is_better"] is True and d["dtype"] == "f64"
    assert d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["config"]["workload"].startswith("tgv_re1600_dns_freeslip_24x24x24")
    assert d["config"]["grid"] == [24, 24, 24]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == 1 and cb["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}
    assert d["value"] > 0 and abs(d["value"] - 24 ** 3 / 1e6 / (d["ms_per_step"] / 1e3)) < 1e-9


def test_reference_arm_other_ranks_are_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                        "--gpus", "2", "--n", "24", "--steps", "1", "--warmup", "0"], env=env,
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_reference_arm_bounds_the_sample(monkeypatch):
    sys.path.insert(0, ROOT)
    import bench
    monkeypatch.setattr(bench, "REF_BUDGET_S", 0.01)     # 40^3 x 3 steps no longer fits
    monkeypatch.delenv("RANK", raising=False)
    args = argparse.Namespace(n=40, bc="freeslip", les=False, strong=False, steps=2, warmup=1,
                              cpu_n=0, gpus=1)
    buf = io.StringIO()
    with redirect_stdout(buf):
        bench.run_reference(args)
    d = json.loads(buf.getvalue())
    n = d["config"]["grid"][0]
    assert 32 <= n < 40 and d["steps"] == 2 and d["warmup"] == 1
    assert "bounded sample of the 40^3 workload" in d["cpu_baseline"]["sample"]
    assert d["config"]["workload"].startswith("tgv_re1600_dns_freeslip_40x40x40")


def test_workload_names_and_weak_strong_grids():
    sys.path.insert(0, ROOT)
    import bench
    a = argparse.Namespace(n=256, bc="freeslip", les=False, strong=False)
    assert bench.workload(a, 1)["nz"] == 256
    assert bench.workload(a, 8)["nz"] == 8 * 255 + 1          # weak: the box replicated in z
    a.strong = True
    assert bench.workload(a, 8)["nz"] == 256                   # strong: the grid stays n^3
    a = argparse.Namespace(n=512, bc="freeslip", les=True, strong=False)
    w = bench.workload(a, 1)
    assert w["name"] == "tgv_re2500_les_freeslip_512x512x512_ab3_sor" and w["phys"]["iles"] == 1
