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


def test_configs_legs_workloads_match_the_shipped_examples():
    """the extra legs of bench.py (`configs`): grids, closures and Poisson settings of the shipped
    examples (examples/*/parameters_*.o3d), weak-scaling replication in z, and initial slabs that
    are consistent pieces of one global field (slab [z0, z0+nk) == the same planes of the whole)"""
    sys.path.insert(0, ROOT)
    import numpy as np
    import bench
    ml = bench.make_workload("mixing_layer")
    assert ml["grid"] == (241, 241, 81) and ml["bc"] == (0, 1, 0) and ml["nscr"] == 1
    assert ml["phys"]["iles"] == 1 and ml["phys"]["idyn"] == 1 and ml["phys"]["eps"] == 1e-5
    assert bench.make_workload("mixing_layer", multigrid=1)["multigrid"] == 1
    cj1, cj8 = bench.make_workload("cojet"), bench.make_workload("cojet", nranks=8)
    assert cj1["grid"] == (257, 513, 129) and cj8["grid"] == (257, 513, 8 * 129)
    assert cj1["bc"] == (0, 0, 0) and cj1["phys"]["omega"] == 1.35 and cj1["phys"]["kmax"] == 1000
    assert abs(cj1["phys"]["dt"] - 0.07 * min(cj1["d"])) < 1e-18
    per = bench.make_workload("tgv", nranks=4, n=256, bc="periodic")
    assert per["grid"] == (256, 256, 4 * 256) and per["bc"] == (0, 0, 0)
    fs = bench.make_workload("tgv", nranks=4, n=256)
    assert fs["grid"] == (256, 256, 4 * 255 + 1)
    # slabs are windows of one global field, whatever the chunking
    for w in (bench.make_workload("tgv", n=24, perturb=True), bench.make_workload("tgv", n=24, les=True)):
        whole = w["init"](0, 24)
        part = w["init"](7, 5)
        for k in whole:
            assert np.array_equal(whole[k][:, :, 7:12], part[k]), k
            assert whole[k].flags["F_CONTIGUOUS"] and whole[k].dtype == np.float64
    f = ml["init"](0, 3)
    assert set(f) == {"ux", "uy", "uz", "pp", "phi"} and f["phi"].min() >= 0.0 and f["phi"].max() <= 1.0
    assert abs(f["ux"]).max() <= 0.5 + 0.03 * 1.2 and not f["uz"].any()
    g = cj1["init"](0, 2)
    assert "phi" not in g and 0.99 < g["ux"].max() <= 1.03 and abs(g["uy"]).max() <= 0.03 + 1e-12


def test_ncu_traffic_is_keyed_by_kernel_and_grid():
    sys.path.insert(0, ROOT)
    import bench
    k = "march_kernel<0,3,2,RhsEpi<dns>,split ring>"
    assert bench.ncu_traffic(k, (256, 256, 256)) > 1.9e9
    assert bench.ncu_traffic(k, (512, 512, 512)) is None       # no capture at that size: null
    assert bench.ncu_traffic("no_such_kernel", (256, 256, 256)) is None


def test_digest_is_a_bitwise_fingerprint():
    sys.path.insert(0, ROOT)
    import numpy as np
    import bench
    a = np.asfortranarray(np.random.default_rng(1).standard_normal((5, 4, 3)))
    b = a.copy(order="F")
    assert bench.digest(a) == bench.digest(b)
    b[2, 1, 1] = np.nextafter(b[2, 1, 1], 1.0)
    assert bench.digest(a) != bench.digest(b)


def test_leg_specs_cover_the_baseline_configs():
    """N = 1: every single-GPU configuration of BASELINE.json; N > 1: the weak-scaled legs plus the
    512^3 LES grid of configs[2] cut into z slabs (same workload as the N = 1 leg)"""
    sys.path.insert(0, ROOT)
    import bench
    one = dict(bench.leg_specs(1))
    assert set(one) == {"tgv512_dns", "tgv512_les", "tgv256_periodic", "tgv257_periodic", "cojet",
                        "mixing_layer_sor", "mixing_layer_multigrid"}
    for world in (2, 4, 8):
        many = dict(bench.leg_specs(world))
        assert set(many) == {"tgv512_dns", "tgv512_les_strong", "tgv256_periodic", "cojet"}
        les1 = bench.make_workload(nranks=1, **one["tgv512_les"])
        lesn = bench.make_workload(nranks=world, **many["tgv512_les_strong"])
        assert lesn["name"] == les1["name"] and lesn["grid"] == (512, 512, 512)
        assert lesn["phys"] == les1["phys"] and lesn["phys"]["iles"] == 1
        assert bench.make_workload(nranks=world, **many["tgv512_dns"])["grid"] == \
            (512, 512, world * 511 + 1)


def test_reference_arm_names_the_n_rank_workload():
    """launched like the b200 arm at N > 1, rank 0 reports the arm's workload (the box replicated
    N times in z) and says that each step timed one replica of it"""
    env = dict(os.environ, RANK="0", WORLD_SIZE="4", LOCAL_RANK="0")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                        "--gpus", "4", "--n", "24", "--steps", "1", "--warmup", "0"], env=env,
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    d = json.loads(r.stdout.strip())
    assert d["n_gpus"] == 4 and d["config"]["workload"] == "tgv_re1600_dns_freeslip_24x24x93_ab3_sor"
    assert d["config"]["workload_grid"] == [24, 24, 93] and d["config"]["grid"] == [24, 24, 24]
    assert "one of the 4 z replicas" in d["cpu_baseline"]["sample"]


def test_secondary_legs_are_guarded_and_cpu_baseline_object(capsys):
    sys.path.insert(0, ROOT)
    import bench
    r = bench.guarded(lambda: {}["missing"])
    assert set(r) == {"error"} and r["error"].startswith("KeyError")
    capsys.readouterr()                               # the traceback goes to stderr
    a = argparse.Namespace(n=24, bc="freeslip", les=False, strong=False, cpu_n=0)
    cb = bench.cpu_baseline(a, 24)
    assert cb["kind"] == "port" and cb["cores"] == 1 and cb["unit"] == "Mpts*steps/s"
    assert cb["value"] > 0 and cb["value_strict_O2_build"] > 0 and "24^3" in cb["sample"]
