#!/usr/bin/env python
"""profiles/r2_sass_summary.txt: per-kernel SASS instruction-class counts of the shipped library
(cuobjdump -sass osinco3d_b200/lib/libo3d_b200.so): UTMALDG (TMA tile loads), SYNCS (mbarrier),
SHFL (warp reductions), FP64 arithmetic, shared-memory loads, global stores, tensor-core ops.

    python scripts/sass_summary.py > profiles/r2_sass_summary.txt
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "osinco3d_b200", "lib", "libo3d_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)[1:]
rows = []
for f in funcs:
    name = f.split("\n", 1)[0].strip()
    ins = re.findall(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", f, re.M)
    c = collections.Counter(i.split(".")[0] for i in ins)
    rows.append((name, len(ins), c["UTMALDG"], c["SYNCS"], c["SHFL"],
                 c["DADD"] + c["DMUL"] + c["DFMA"] + c["DSETP"], c["LDS"], c["STG"],
                 c["HMMA"] + c["UTCMMA"] + c["IMMA"] + c["DMMA"]))
names = subprocess.run(["c++filt"] + [r[0] for r in rows], capture_output=True,
                       text=True).stdout.splitlines()
print("Round 2 -- SASS evidence for the shipped libo3d_b200.so (cuobjdump -sass, sm_100a), per kernel:")
print("instruction counts by class.  UTMALDG = cp.async.bulk.tensor (TMA tile load), SYNCS = mbarrier ops;")
print("no HMMA / UTCMMA / tcgen05 anywhere: nothing on this path is a dense contraction (FP64 stencils).")
print()
fmt = "%-112s %6s %7s %6s %5s %6s %6s %5s %4s"
print(fmt % ("kernel", "instr", "UTMALDG", "SYNCS", "SHFL", "FP64", "LDS", "STG", "MMA"))
for r, n in zip(rows, names):
    n = re.sub(r"\(anonymous namespace\)::|o3d::", "", n).split("(")[0][:110]
    print(fmt % ((n,) + r[1:]))
tot = [sum(r[i] for r in rows) for i in range(1, 9)]
print(fmt % (("TOTAL (%d kernels)" % len(rows),) + tuple(tot)))
