# round 2, pass d: isolate the persistent-pass slowdown: (1) ncu of the launch-per-pass kernel at
# 512^3 on a REAL pass, (2) the persistent kernel's hot loop launched one CTA per item (oneshot)
TAG=${1:-r2d}
mkdir -p gpurun_out
B="python bench.py --n 512 --steps 1 --warmup 3 --no-e2e --no-cpu --legs none --no-parity"
O3D_SOR_PERSIST=0 timeout 600 ncu --set full --clock-control none -k regex:'sor_tma' --launch-skip 8 -c 1 -o gpurun_out/${TAG}_sor512_old $B > gpurun_out/${TAG}_ncu_old.log 2>&1
ncu -i gpurun_out/${TAG}_sor512_old.ncu-rep --page raw --csv > gpurun_out/${TAG}_sor512_old_raw.csv 2>/dev/null
python profiles/ncu_summary.py gpurun_out/${TAG}_sor512_old_raw.csv > gpurun_out/${TAG}_sor512_old_summary.txt 2>&1
rm -f gpurun_out/${TAG}_sor512_old.ncu-rep
cat gpurun_out/${TAG}_sor512_old_summary.txt
show='import sys,json
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=l["roofline"]["stages"]["sor"]; print("sor %.4f ms/iter frac %.3f K=%.1f" % (s["ms_per_launch"], s["frac"], s["iterations_per_step"]))'
B2="python bench.py --n 512 --steps 6 --warmup 3 --no-e2e --no-cpu --legs none --no-parity"
for N in 0 6 8 13 26; do
  echo "== oneshot nch $N"; O3D_PERSIST_ONESHOT=1 O3D_NCH_P=$N timeout 300 $B2 2>> gpurun_out/${TAG}.err | python -c "$show"
done
echo "== persistent default"; timeout 300 $B2 2>> gpurun_out/${TAG}.err | python -c "$show"
echo "== launch per pass"; O3D_SOR_PERSIST=0 timeout 300 $B2 2>> gpurun_out/${TAG}.err | python -c "$show"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -5 gpurun_out/${TAG}_pytest_gpu.log
