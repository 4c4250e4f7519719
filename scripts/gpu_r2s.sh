# round 2, 8-GPU pass (gpurun --gpus 8; charged 8x -- keep it short): parity on 8 slabs, the weak
# scaling line with all legs + parity object, strong scaling of 1024^3, one step timeline.
TAG=${1:-r2s}
N=${2:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node $N --master-port 29511 tests/mgpu_check.py > gpurun_out/${TAG}_mgpu${N}_parity.log 2>&1
echo "mgpu_check exit $?" >> gpurun_out/${TAG}_mgpu${N}_parity.log
grep -c BITWISE-EQUAL gpurun_out/${TAG}_mgpu${N}_parity.log; grep "MISMATCH\|exit" gpurun_out/${TAG}_mgpu${N}_parity.log | head
timeout 600 $TR --nproc-per-node $N --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_n${N}.json 2> gpurun_out/${TAG}_bench_n${N}.err
tail -c 600 gpurun_out/${TAG}_bench_n${N}.err
timeout 600 $TR --nproc-per-node $N --master-port 29522 bench.py --gpus $N --grid 1024 --strong --steps 10 --warmup 4 --no-e2e --legs none > gpurun_out/${TAG}_bench_1024_strong_n${N}.json 2> gpurun_out/${TAG}_bench_1024_strong_n${N}.err
tail -c 600 gpurun_out/${TAG}_bench_1024_strong_n${N}.err
O3D_TRACE=8 timeout 300 $TR --nproc-per-node $N --master-port 29523 bench.py --gpus $N --grid 512 --steps 6 --warmup 4 --no-e2e --legs none --no-parity > gpurun_out/${TAG}_trace512_n${N}.json 2> gpurun_out/${TAG}_trace512_n${N}.err
grep "o3d trace" gpurun_out/${TAG}_trace512_n${N}.err | grep " r0 \|rank 0" > gpurun_out/${TAG}_trace_n${N}_rank0.txt
O3D_SOR_PEER=0 timeout 600 $TR --nproc-per-node $N --master-port 29524 bench.py --gpus $N --steps 20 --warmup 5 --no-e2e --legs tgv256_periodic --no-parity > gpurun_out/${TAG}_bench_n${N}_nccl.json 2> gpurun_out/${TAG}_bench_n${N}_nccl.err
python - <<PY
import json
for f in ("gpurun_out/${TAG}_bench_n${N}.json", "gpurun_out/${TAG}_bench_1024_strong_n${N}.json", "gpurun_out/${TAG}_trace512_n${N}.json", "gpurun_out/${TAG}_bench_n${N}_nccl.json"):
    try:
        l = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f, l["config"]["workload"], "value %.1f ms/step %.4f" % (l["value"], l["ms_per_step"]), l["config"]["sor_path"], "e2e", (l.get("e2e") or {}).get("value"))
    for k, s in l["roofline"]["stages"].items():
        print("    %-5s %.4f ms/launch frac %.3f" % (k, s["ms_per_launch"], s.get("frac", 0)))
    for c in l.get("configs", []):
        if "error" in c: print("   ", c["key"], c["error"]); continue
        print("   ", c["key"], c["grid"], "%.1f Mpts/s %.3f ms/step K=%.1f" % (c["value"], c["ms_per_step"], c["poisson_iterations_per_step"]), c["sor_path"], {k: round(v["ms_per_launch"], 4) for k, v in c["stages"].items()})
    print("    parity:", json.dumps(l.get("parity"))[:400])
PY
