# round 2, multi-GPU pass (gpurun --gpus N): P-GPU == 1-GPU bitwise on every path, then the
# bench lines (weak scaling + legs + parity) with peer-memory SOR and, for comparison, NCCL halos.
TAG=${1:-r2m}
N=${2:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node $N --master-port 29511 tests/mgpu_check.py > gpurun_out/${TAG}_mgpu${N}_parity.log 2>&1
echo "mgpu_check exit $?" >> gpurun_out/${TAG}_mgpu${N}_parity.log
grep -c BITWISE-EQUAL gpurun_out/${TAG}_mgpu${N}_parity.log; grep "MISMATCH\|Error\|error\|exit" gpurun_out/${TAG}_mgpu${N}_parity.log | head -20
timeout 600 $TR --nproc-per-node $N --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_n${N}.json 2> gpurun_out/${TAG}_bench_n${N}.err
tail -c 1500 gpurun_out/${TAG}_bench_n${N}.err
O3D_SOR_PEER=0 timeout 600 $TR --nproc-per-node $N --master-port 29522 bench.py --gpus $N --steps 20 --warmup 5 --no-e2e --legs tgv256_periodic,tgv512_dns > gpurun_out/${TAG}_bench_n${N}_nccl.json 2> gpurun_out/${TAG}_bench_n${N}_nccl.err
O3D_SOR_PERSIST=0 timeout 600 $TR --nproc-per-node $N --master-port 29523 bench.py --gpus $N --steps 20 --warmup 5 --no-e2e --legs tgv256_periodic --no-parity > gpurun_out/${TAG}_bench_n${N}_r1path.json 2> gpurun_out/${TAG}_bench_n${N}_r1path.err
python - <<PY
import json
for f in ("gpurun_out/${TAG}_bench_n${N}.json", "gpurun_out/${TAG}_bench_n${N}_nccl.json", "gpurun_out/${TAG}_bench_n${N}_r1path.json"):
    try:
        l = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f, "value %.1f ms/step %.4f" % (l["value"], l["ms_per_step"]), l["config"]["sor_path"], "e2e", (l.get("e2e") or {}).get("value"))
    for k, s in l["roofline"]["stages"].items():
        print("    %-5s %.4f ms/launch frac %.3f" % (k, s["ms_per_launch"], s.get("frac", 0)))
    for c in l.get("configs", []):
        if "error" in c: print("   ", c["key"], c["error"]); continue
        print("   ", c["key"], c["grid"], "%.1f Mpts/s %.3f ms/step K=%.1f" % (c["value"], c["ms_per_step"], c["poisson_iterations_per_step"]), c["sor_path"], {k: round(v["ms_per_launch"], 4) for k, v in c["stages"].items()})
    print("    parity:", json.dumps(l.get("parity"))[:600])
PY
