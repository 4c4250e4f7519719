# One GPU pass for the profiles/ evidence of a round: tests, bench lines, ncu launch list + full capture.
# usage: bash scripts/gpu_round.sh <tag>
TAG=${1:-rX}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_256.json 2> gpurun_out/${TAG}_bench_256.err; tail -c 1500 gpurun_out/${TAG}_bench_256.json
timeout 600 python bench.py --n 512 --les --steps 10 --warmup 4 --no-cpu > gpurun_out/${TAG}_bench_512les.json 2> gpurun_out/${TAG}_bench_512les.err
timeout 600 python bench.py --bc periodic --n 257 --steps 10 --warmup 4 --no-cpu --no-e2e > gpurun_out/${TAG}_bench_257per.json 2> gpurun_out/${TAG}_bench_257per.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_256.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/${TAG}_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'march_kernel|sor_' --launch-skip 15 -c 8 -o gpurun_out/${TAG}_full python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/${TAG}_ncu_full.log 2>&1
ncu -i gpurun_out/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_raw.csv 2>/dev/null
python profiles/ncu_summary.py gpurun_out/${TAG}_full_raw.csv > gpurun_out/${TAG}_ncu_full_summary.txt 2>&1
ls -la gpurun_out
