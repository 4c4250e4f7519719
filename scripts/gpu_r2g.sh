# round 2, pass g (1 GPU): full GPU suite, LES RHS variants at 512^3, the 1024^3 single-GPU base
TAG=${1:-r2g}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -4 gpurun_out/${TAG}_pytest_gpu.log
show='import sys,json
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); st=l["roofline"]["stages"]; print("%.1f Mpts/s %.3f ms/step  " % (l["value"], l["ms_per_step"]) + "  ".join("%s %.4f (%.2f)" % (k, s["ms_per_launch"], s.get("frac",0)) for k,s in st.items()))'
B="python bench.py --n 512 --les --steps 8 --warmup 4 --no-e2e --no-cpu --legs none --no-parity"
for V in "A=1" "O3D_RHS_LES=general" "O3D_RHS_RING=classic" "O3D_RHS_RING=split3"; do
  echo "== LES 512 $V"; env $V timeout 300 $B 2>> gpurun_out/${TAG}.err | python -c "$show"
done
echo "== 1024^3 on one GPU"
timeout 600 python bench.py --n 1024 --strong --steps 6 --warmup 3 --no-e2e --no-cpu --legs none --no-parity > gpurun_out/${TAG}_bench_1024_n1.json 2> gpurun_out/${TAG}_bench_1024_n1.err
tail -c 400 gpurun_out/${TAG}_bench_1024_n1.err; python -c "$show" < gpurun_out/${TAG}_bench_1024_n1.json
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
