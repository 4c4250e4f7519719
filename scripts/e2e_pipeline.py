#!/usr/bin/env python
"""End-to-end time of the host-pointer procedures (bench.py `e2e` leg) against the number of z
chunks of the copy pipeline (csrc/pipeline.cu): one JSON line per setting.

    python scripts/e2e_pipeline.py [--n 256] [--steps 3] [--chunks 0,4,8,16]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=256)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--chunks", default="0,4,8,16")
    a = ap.parse_args()
    import osinco3d_b200 as o3d
    args = argparse.Namespace(n=a.n, bc="freeslip", les=False, strong=False)
    w = bench.workload(args, 1)
    for c in [int(x) for x in a.chunks.split(",")]:
        r = bench.run_e2e(o3d, w, a.steps, a.warmup, c)
        print(json.dumps({"n": a.n, "chunks": c, "ms_per_step": r["ms_per_step"],
                          "mpts_steps_per_s": r["value"],
                          "gb_per_s_each_way": r["h2d_bytes_per_step"] / r["ms_per_step"] / 1e6}),
              flush=True)


if __name__ == "__main__":
    main()
