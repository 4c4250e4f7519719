#!/usr/bin/env python
"""Generate tests/golden/oracle_history.json: the 17 stats.dat columns (src/utils.f90:243-361,
src/IOfunctions.f90:504-552) of the CPU oracle at every row the reference's shipped histories
hold in tests/golden/reference_stats.json (every 25 steps, src/osinco3d_main.f90:167-181), for

  examples/tgv_re1600_dns  185^3, dt = 0.05 pi/184, omega 1.887, eps 1e-4      rows t = 0 .. 100 dt
  examples/tgv_re2500_les  129^3, dt = 5e-4, cs 0.17, omega0 1.999 dynamic, eps 1e-6  25 .. 125 dt

    python tests/golden/make_oracle_history.py        (~5 min of CPU, two processes)

The GPU history tests compare with these rows (same source version as the product) AND with the
reference's own file.  Measured drift of the oracle against the reference file, recorded in the
JSON: DNS <= 3e-8 on E_k / enstrophy at all five rows (Poisson-tolerance noise); LES grows
linearly, 4.5e-7 per 25 steps (2.3e-6 at step 125) -- the shipped LES file predates the current
les_turbulence.f90, eps 1e-6 -> 1e-9 moves it by 1e-9 only."""
import json
import os
import sys
from concurrent.futures import ProcessPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
PI = 3.141592653589793


def run(which):
    from oracle import oracle_py as O
    O.build()
    gold = json.load(open(os.path.join(HERE, "reference_stats.json")))
    if which == "tgv_re1600_dns":
        n = 185
        d = PI / (n - 1)
        g = O.grid(n, n, n, d, d, d, (1, 1, 1))
        s = O.Sim(g, re=1600.0, dt=0.05 * d, itscheme=3, omega=1.887, eps=1e-4, kmax=10000, idyn=0)
        first = 0
    else:
        n = 129
        d = PI / (n - 1)
        g = O.grid(n, n, n, d, d, d, (1, 1, 1))
        s = O.Sim(g, re=2500.0, dt=5e-4, itscheme=3, iles=1, cs=0.17, omega=1.999, eps=1e-6,
                  kmax=10000, idyn=1)
        first = 1
    ux, uy, uz, pp, phi = O.init_tgv(g, nscr=0)
    s.set(ux=ux, uy=uy, uz=uz, pp=pp)
    rows, iters, step = [], [], 0
    for r, ref in enumerate(gold[which]["rows"]):
        while step < 25 * (r + first):
            iters.append(s.step())
            step += 1
        st = [float(v) for v in s.stats()]
        rows.append({"step": step, "columns": st,
                     "rel_vs_reference_file": [(st[c] - ref[c]) / ref[c] for c in (1, 2, 3, 4)]})
    return which, {"grid": n, "rows": rows, "sor_iters_per_step": iters}


if __name__ == "__main__":
    out = {"source": "oracle/o3d_oracle.c (gcc -O2 -ffp-contract=off), lexicographic SOR",
           "rel_columns": ["E_k", "eps", "eps2", "enstrophy"]}
    with ProcessPoolExecutor(2) as ex:
        for k, v in ex.map(run, ["tgv_re1600_dns", "tgv_re2500_les"]):
            out[k] = v
    json.dump(out, open(os.path.join(HERE, "oracle_history.json"), "w"), indent=1)
    print("wrote oracle_history.json")
