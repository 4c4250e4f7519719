# Multi-GPU evidence in ONE gpurun call (charged N x box time: keep it short).
# usage (on the box, via gpurun --gpus N): bash scripts/gpu_scale.sh <tag> <N>
#   - tests/mgpu_check.py on N GPUs: every case bitwise equal to one GPU
#   - bench.py weak scaling at 1, 2, 4 ... N GPUs (256^2 x 255 planes per GPU)
#   - N == 8: the north-star strong-scaling point, 1024^3 on 8 GPUs (--strong), and the same
#     per-GPU work on one GPU (1024 x 1024 x 128) for the speed-up
TAG=${1:-rX}
N=${2:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node $N --master-port 29511 tests/mgpu_check.py > gpurun_out/${TAG}_mgpu${N}_parity.log 2>&1
echo "mgpu_check exit $?" >> gpurun_out/${TAG}_mgpu${N}_parity.log
tail -4 gpurun_out/${TAG}_mgpu${N}_parity.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/${TAG}_bench_256_n1.json 2> gpurun_out/${TAG}_bench_256_n1.err
for P in 2 4 8; do
  [ $P -le $N ] || continue
  timeout 300 $TR --nproc-per-node $P --master-port $((29520 + P)) bench.py --gpus $P --steps 20 --warmup 5 \
      > gpurun_out/${TAG}_bench_256_n${P}.json 2> gpurun_out/${TAG}_bench_256_n${P}.err
done
if [ $N -ge 8 ]; then
  timeout 600 $TR --nproc-per-node 8 --master-port 29540 bench.py --gpus 8 --n 1024 --strong --steps 10 --warmup 4 --no-e2e \
      > gpurun_out/${TAG}_bench_1024_strong_n8.json 2> gpurun_out/${TAG}_bench_1024_strong_n8.err
fi
python scripts/show_bench.py gpurun_out/${TAG}_bench_*_n*.json
