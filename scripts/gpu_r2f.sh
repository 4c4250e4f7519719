# round 2, pass f: persistent SOR with the launch-per-pass body + ticket scheduling, all K > 1 grids
TAG=${1:-r2f}
mkdir -p gpurun_out
show='import sys,json
l=json.loads(sys.stdin.read().strip().splitlines()[-1])
s=l["roofline"]["stages"]["sor"]; out=["256: %.4f (%.2f)" % (s["ms_per_launch"], s["frac"])]
for c in l.get("configs", []):
    if "stages" in c and "sor" in c["stages"]:
        s=c["stages"]["sor"]; out.append("%s: %.4f (%.2f) K=%.1f step %.3f ms" % (c["key"], s["ms_per_launch"], s["frac"], c["poisson_iterations_per_step"], c["ms_per_step"]))
print("  |  ".join(out))'
B="python bench.py --steps 10 --warmup 4 --no-e2e --no-cpu --no-parity --legs tgv512_dns,tgv256_periodic,tgv257_periodic,cojet,mixing_layer_sor"
for V in "A=1" "O3D_PERSIST_DYN=0" "O3D_SOR_PERSIST=0" "O3D_PERSIST_ONESHOT=1" "O3D_PERSIST_CTAS=2"; do
  echo "== $V"; env $V timeout 300 $B 2>> gpurun_out/${TAG}_sweep.err | python -c "$show"
done
timeout 900 python -m pytest tests/test_gpu_poisson.py tests/test_gpu_step.py tests/test_gpu_multigrid.py -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -4 gpurun_out/${TAG}_pytest_gpu.log
