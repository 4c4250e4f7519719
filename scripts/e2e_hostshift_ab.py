"""A/B of the end-to-end leg of bench.py (o3d_predict_velocity + o3d_correct_pression +
o3d_correct_velocity on pinned HOST arrays, 256^3 TGV) with and without the host-side history
shift (o3d_set_hostshift, csrc/host_copier.h), interleaved so that box-to-box and run-to-run
drift does not decide it.  One JSON line per run + a summary line.

    python scripts/e2e_hostshift_ab.py [--rounds 2] [--steps 5] [--threads 0,4,8]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rounds", type=int, default=2)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--n", type=int, default=256)
    ap.add_argument("--chunks", default="16")
    args = ap.parse_args()
    import bench
    import osinco3d_b200 as o3d
    from osinco3d_b200 import modules as M
    o3d._lib.check(o3d.lib().o3d_set_device(0))
    w = bench.make_workload("tgv", 1, args.n, "freeslip", False, False)
    res = {}
    for r in range(args.rounds):
        for chunks in [int(c) for c in args.chunks.split(",")]:
            for hs in (0, 1):
                M.set_hostshift(hs)
                e = bench.run_e2e(o3d, w, args.steps, 2, chunks)
                e.pop("path")
                e["round"] = r
                print(json.dumps(e), flush=True)
                res.setdefault((chunks, hs), []).append(e["ms_per_step"])
    summ = {"%d chunks, hostshift %d" % k: {"ms_per_step_min": min(v), "ms_per_step_all": v}
            for k, v in res.items()}
    summ["host_cores"] = os.cpu_count()
    summ["hostshift_threads_env"] = os.environ.get("O3D_HOSTSHIFT_THREADS")
    print(json.dumps({"summary": summ}))


if __name__ == "__main__":
    main()
