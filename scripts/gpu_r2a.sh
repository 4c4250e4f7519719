# round 2, first GPU pass: full GPU test suite with the persistent SOR kernel as default, then
# A/B timings persistent vs launch-per-pass on the K > 1 configurations.
TAG=${1:-r2a}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -5 gpurun_out/${TAG}_pytest_gpu.log
for P in 1 0; do
  O3D_SOR_PERSIST=$P O3D_TIMING_OUT=gpurun_out/${TAG}_examples_persist$P.jsonl timeout 600 python -m pytest tests/test_gpu_examples.py -q -k full_size_steps_and_timing > gpurun_out/${TAG}_examples_persist$P.log 2>&1
  O3D_SOR_PERSIST=$P timeout 300 python bench.py --bc periodic --n 257 --steps 10 --warmup 4 --no-cpu --no-e2e > gpurun_out/${TAG}_bench_257per_persist$P.json 2> gpurun_out/${TAG}_bench_257per_persist$P.err
  O3D_SOR_PERSIST=$P timeout 300 python bench.py --n 512 --les --steps 10 --warmup 4 --no-cpu --no-e2e > gpurun_out/${TAG}_bench_512les_persist$P.json 2> gpurun_out/${TAG}_bench_512les_persist$P.err
  O3D_SOR_PERSIST=$P timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e > gpurun_out/${TAG}_bench_256_persist$P.json 2> gpurun_out/${TAG}_bench_256_persist$P.err
done
cat gpurun_out/${TAG}_examples_persist*.jsonl
python scripts/show_bench.py gpurun_out/${TAG}_bench_*.json
