# final evidence of a round on a short GPU budget: full suite, bench lines, ncu launch list
TAG=${1:-rX}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -4 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_256.json 2> gpurun_out/${TAG}_bench_256.err
timeout 300 python bench.py --n 512 --les --steps 10 --warmup 4 --no-cpu --no-e2e > gpurun_out/${TAG}_bench_512les.json 2>/dev/null
timeout 300 python bench.py --n 512 --steps 10 --warmup 4 --no-cpu --no-e2e > gpurun_out/${TAG}_bench_512.json 2>/dev/null
timeout 300 python bench.py --bc periodic --n 257 --steps 10 --warmup 4 --no-cpu --no-e2e > gpurun_out/${TAG}_bench_257per.json 2>/dev/null
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_256.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1
python scripts/show_bench.py gpurun_out/${TAG}_bench_256.json gpurun_out/${TAG}_bench_512les.json gpurun_out/${TAG}_bench_512.json gpurun_out/${TAG}_bench_257per.json
