"""time o3d_s_step_diagnostics alone (device events) at n^3; under ncu this gives its launch list"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import osinco3d_b200 as o3d
from bench import tgv_slab
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
d = 3.141592653589793 / (n - 1)
cfg = o3d.make_config(n, n, n, d, d, d, bc=(1, 1, 1), re=1600.0, dt=0.05 * d, omega=1.887, eps=1e-4)
ses = o3d.Session(cfg)
ux, uy, uz, pp = tgv_slab(n, n, 0, d)
ses.set(ux=ux, uy=uy, uz=uz, pp=pp)
for _ in range(3):
    ses.step()
ses.step_diagnostics()
ses.sync()
ses.stopwatch_start()
for _ in range(10):
    ses.step_diagnostics()
print("step_diagnostics: %.4f ms per call (device events)" % (ses.stopwatch_stop() / 10))
t0 = time.perf_counter()
for _ in range(10):
    ses.step()
    ses.step_diagnostics()
print("step + diagnostics: %.4f ms wall" % (1e3 * (time.perf_counter() - t0) / 10))
