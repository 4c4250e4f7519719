# round 2, pass c: z-stagger sweep of the persistent SOR pass + full GPU test suite
TAG=${1:-r2c}
mkdir -p gpurun_out
show='import sys,json
l=json.loads(sys.stdin.read().strip().splitlines()[-1])
s=l["roofline"]["stages"]["sor"]; out=["256: %.4f (%.2f)" % (s["ms_per_launch"], s["frac"])]
for c in l.get("configs", []):
    if "stages" in c and "sor" in c["stages"]:
        s=c["stages"]["sor"]; out.append("%s: %.4f (%.2f) K=%.1f step %.3f ms" % (c["key"], s["ms_per_launch"], s["frac"], c["poisson_iterations_per_step"], c["ms_per_step"]))
print("  |  ".join(out))'
B="python bench.py --steps 10 --warmup 4 --no-e2e --no-cpu --no-parity --legs tgv512_dns,tgv257_periodic,cojet,mixing_layer_sor"
for S in 0 2 4 8 16; do
  echo "== stagger $S"; O3D_PERSIST_STAGGER=$S timeout 300 $B 2>> gpurun_out/${TAG}_sweep.err | python -c "$show"
done
echo "== launch-per-pass"; O3D_SOR_PERSIST=0 timeout 300 $B 2>> gpurun_out/${TAG}_sweep.err | python -c "$show"
for N in 20 40; do
  echo "== stagger 4 nch $N"; O3D_NCH_P=$N timeout 300 $B 2>> gpurun_out/${TAG}_sweep.err | python -c "$show"
done
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -5 gpurun_out/${TAG}_pytest_gpu.log
