#!/usr/bin/env python
"""Extract the reference's golden statistics rows into small committed fixtures.

Run in the build container (where /root/reference exists); the GPU box never reads
/root/reference.  Source files (reference checkout):
  examples/tgv_re1600_dns/tgv_stats_re1600_dns.dat   (185^3 DNS, free-slip, AB3, SOR)
  examples/tgv_re2500_les/tgv_stats_re2500_les.dat   (129^3 LES, Smagorinsky, dynamic omega)
Only the first rows are kept (one row per 25 time steps, 17 columns, format 17es21.12,
reference src/IOfunctions.f90:552).
"""
import json
import os
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def rows(path, n):
    out = []
    with open(path) as fh:
        for line in fh:
            if line.lstrip().startswith("#") or not line.strip():
                continue
            out.append([float(v) for v in line.split()])
            if len(out) == n:
                break
    return out


golden = {
    "source": "jojoledemago/osinco3d examples/*/tgv_stats_*.dat (first rows, verbatim values)",
    "columns": ["t", "e_k", "eps", "eps2", "dzeta", "ux2", "uy2", "uz2", "duxdx2", "duxdy2",
                "duxdz2", "duydx2", "duydy2", "duydz2", "duzdx2", "duzdy2", "duzdz2"],
    "tgv_re1600_dns": {
        "config": {"n": 185, "re": 1600.0, "cfl": 0.05, "omega": 1.887, "eps": 1e-4,
                   "kmax": 10000, "idyn": 0, "itscheme": 3, "iles": 0, "nscr": 1, "sc": 1.0},
        "rows": rows(os.path.join(REF, "examples/tgv_re1600_dns/tgv_stats_re1600_dns.dat"), 5),
    },
    "tgv_re2500_les": {
        "config": {"n": 129, "re": 2500.0, "dt": 5e-4, "omega": 1.999, "eps": 1e-6,
                   "kmax": 10000, "idyn": 1, "itscheme": 3, "iles": 1, "cs": 0.17, "nscr": 0},
        "rows": rows(os.path.join(REF, "examples/tgv_re2500_les/tgv_stats_re2500_les.dat"), 5),
    },
}
with open(os.path.join(HERE, "reference_stats.json"), "w") as fh:
    json.dump(golden, fh, indent=1)
print("wrote", os.path.join(HERE, "reference_stats.json"))
